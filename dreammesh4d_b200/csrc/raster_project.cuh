// Projection math shared by the preprocess forward (compiled with -fmad=false: the bit-exact arithmetic spec of
// DESIGN.md §4) and the preprocess backward (its own translation unit with default FMA contraction: gradients carry no
// integer state).  Algorithm: SURVEY.md Appendix A.2 steps 1-9.
#pragma once
#include "raster_internal.cuh"

namespace {

struct PreArgs {
    int P, H, W, n_views, n_sets, channels, gx, gy, tiles, rec, acc;
    const float* means3D;   long long means3D_stride;
    const float* scales;    long long scales_stride;
    const float* rotations; long long rotations_stride;
    const float* opacities; long long opacities_stride;
    const float* colors;    long long colors_stride;
    const float* colors2;   long long colors2_stride;
    const float* cov3D;     long long cov3D_stride;     // optional: precomputed 3D covariances instead of scales / rotations
    const float* view_params;
    float* g_rec;
    unsigned int* g_rect;
    unsigned int* tile_count;
    int32_t* radii;
};

// quaternion rows are 16 bytes: one vector load when the row is aligned (always for torch tensors)
__device__ __forceinline__ float4 load_quat(const float* __restrict__ q) {
    if ((reinterpret_cast<uintptr_t>(q) & 15) == 0) return *reinterpret_cast<const float4*>(q);
    return make_float4(q[0], q[1], q[2], q[3]);
}

__device__ __forceinline__ float ndc2pix(float v, int S) { return ((v + 1.0f) * (float)S - 1.0f) * 0.5f; }

__device__ __forceinline__ void quat_to_R(float r, float x, float y, float z, float Rm[3][3]) {
    Rm[0][0] = 1.f - 2.f * (y * y + z * z); Rm[0][1] = 2.f * (x * y - r * z); Rm[0][2] = 2.f * (x * z + r * y);
    Rm[1][0] = 2.f * (x * y + r * z); Rm[1][1] = 1.f - 2.f * (x * x + z * z); Rm[1][2] = 2.f * (y * z - r * x);
    Rm[2][0] = 2.f * (x * z - r * y); Rm[2][1] = 2.f * (y * z + r * x); Rm[2][2] = 1.f - 2.f * (x * x + y * y);
}

// Shared forward math: everything up to (a, b, c) of the 2D covariance. Returns false if culled
// by the near plane.
struct Proj {
    float tx, ty, tz;          // view-space point
    float hx, hy, hw, p_w;     // clip-space and 1/(w+eps)
    float S[3][3];             // 3D covariance
    float A[2][3], B[2][3];    // A = J W3, B = A Sigma
    float a, b, c;             // 2D covariance incl. the 0.3 low-pass
    float cx, cy;              // clamped view-space x, y
    float xmul, ymul;          // 0 where the 1.3 tanfov clamp was active
};

// 3D covariance Sigma = L L^T, L = R(q) diag(s)
__device__ __forceinline__ void gaussian_sigma(float s0, float s1, float s2, float qr, float qx, float qy, float qz, float (&S)[3][3]) {
    float Rm[3][3], L[3][3];
    quat_to_R(qr, qx, qy, qz, Rm);
    const float s[3] = {s0, s1, s2};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) L[a][k] = Rm[a][k] * s[k];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a; b < 3; ++b) {
            S[a][b] = L[a][0] * L[b][0] + L[a][1] * L[b][1] + L[a][2] * L[b][2];
            S[b][a] = S[a][b];
        }
}
// cov3D_precomp layout of the replaced module: xx, xy, xz, yy, yz, zz
__device__ __forceinline__ void load_sigma(const float* __restrict__ c, float (&S)[3][3]) {
    S[0][0] = c[0]; S[0][1] = S[1][0] = c[1]; S[0][2] = S[2][0] = c[2];
    S[1][1] = c[3]; S[1][2] = S[2][1] = c[4]; S[2][2] = c[5];
}

// Everything from the 3D covariance o.S (set by the caller) to (a, b, c) of the 2D covariance.  Returns false if culled by
// the near plane.
__device__ __forceinline__ bool project_with_sigma(const float* __restrict__ vp, float px, float py, float pz,
                                                   float focal_x, float focal_y, Proj& o) {
    const float* V = vp;
    const float* PV = vp + 16;
    o.tx = V[0] * px + V[4] * py + V[8] * pz + V[12];
    o.ty = V[1] * px + V[5] * py + V[9] * pz + V[13];
    o.tz = V[2] * px + V[6] * py + V[10] * pz + V[14];
    if (o.tz <= 0.2f) return false;
    o.hx = PV[0] * px + PV[4] * py + PV[8] * pz + PV[12];
    o.hy = PV[1] * px + PV[5] * py + PV[9] * pz + PV[13];
    o.hw = PV[3] * px + PV[7] * py + PV[11] * pz + PV[15];
    o.p_w = 1.0f / (o.hw + 0.0000001f);
    const float limx = 1.3f * vp[DM4D_VIEW_TANFOVX], limy = 1.3f * vp[DM4D_VIEW_TANFOVY];
    const float txtz = o.tx / o.tz, tytz = o.ty / o.tz;
    o.xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    o.ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    o.cx = fminf(limx, fmaxf(-limx, txtz)) * o.tz;
    o.cy = fminf(limy, fmaxf(-limy, tytz)) * o.tz;
    const float J00 = focal_x / o.tz, J02 = -(focal_x * o.cx) / (o.tz * o.tz);
    const float J11 = focal_y / o.tz, J12 = -(focal_y * o.cy) / (o.tz * o.tz);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        o.A[0][j] = J00 * V[0 + 4 * j] + J02 * V[2 + 4 * j];
        o.A[1][j] = J11 * V[1 + 4 * j] + J12 * V[2 + 4 * j];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j) o.B[r][j] = o.A[r][0] * o.S[0][j] + o.A[r][1] * o.S[1][j] + o.A[r][2] * o.S[2][j];
    o.a = o.B[0][0] * o.A[0][0] + o.B[0][1] * o.A[0][1] + o.B[0][2] * o.A[0][2];
    o.b = o.B[0][0] * o.A[1][0] + o.B[0][1] * o.A[1][1] + o.B[0][2] * o.A[1][2];
    o.c = o.B[1][0] * o.A[1][0] + o.B[1][1] * o.A[1][1] + o.B[1][2] * o.A[1][2];
    o.a += 0.3f;
    o.c += 0.3f;
    return true;
}

// View parameters of a block: a block of consecutive (view, Gaussian) indices touches at most two views when
// P >= blockDim, so both parameter rows are staged in shared memory once (the kernels read ~40 of the 48 floats per
// thread; as global loads they were 48 LDG per thread).  Returns the row of view v (shared or global).
struct ViewCache {
    float rows[2][DM4D_VIEW_STRIDE];       // the only shared state; (v_first, n) are block-uniform register values
};
struct ViewRows {
    const ViewCache* vc;
    const float* view_params;
    int v_first, n;
    // call from every thread of the block, then __syncthreads()
    __device__ __forceinline__ void fill(ViewCache* cache, const float* __restrict__ vp, int v_first_, int n_views) {
        vc = cache; view_params = vp; v_first = v_first_;
        n = min(2, n_views - v_first_);
        for (int i = threadIdx.x; i < n * DM4D_VIEW_STRIDE; i += blockDim.x) cache->rows[0][i] = vp[(size_t)v_first_ * DM4D_VIEW_STRIDE + i];
    }
    __device__ __forceinline__ const float* row(int v) const {
        const int k = v - v_first;
        return k < n ? vc->rows[k] : view_params + (size_t)v * DM4D_VIEW_STRIDE;
    }
};

// (view, Gaussian) of a thread without a 64-bit division per thread: the block's first index is divided once
// (block-uniform), the thread walks forward from there.
__device__ __forceinline__ void split_index(long long first, int tid, int P, int& v, int& g) {
    const int v0 = (int)(first / P);
    long long gg = first - (long long)v0 * P + tid;
    v = v0;
    while (gg >= P) { gg -= P; ++v; }
    g = (int)gg;
}

PreArgs make_pre_args(const dm4d_raster_desc* d, const RasterLayout& L, int32_t* radii) {
    PreArgs a;
    a.P = L.P; a.H = L.H; a.W = L.W; a.n_views = L.n_views; a.n_sets = d->n_sets; a.channels = L.channels;
    a.gx = L.gx; a.gy = L.gy; a.tiles = L.tiles; a.rec = L.rec; a.acc = L.acc;
    a.means3D = d->means3D; a.means3D_stride = d->means3D_stride;
    a.scales = d->scales; a.scales_stride = d->scales_stride;
    a.rotations = d->rotations; a.rotations_stride = d->rotations_stride;
    a.opacities = d->opacities; a.opacities_stride = d->opacities_stride;
    a.colors = d->colors; a.colors_stride = d->colors_stride;
    a.colors2 = d->colors2; a.colors2_stride = d->colors2_stride;
    a.cov3D = d->cov3D; a.cov3D_stride = d->cov3D_stride;
    a.view_params = d->view_params;
    a.g_rec = L.g_rec; a.g_rect = L.g_rect; a.tile_count = L.tile_count; a.radii = radii;
    return a;
}

}  // namespace
