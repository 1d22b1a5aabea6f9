// Per-(view, Gaussian) projection ("preprocess") forward and backward.
//
// Compiled with -fmad=false: every + - * / sqrt below is a separately rounded IEEE binary32
// operation evaluated left to right, which is the arithmetic spec of DESIGN.md §4 and what
// makes the integer state (radii, tile rectangles, depth keys) bit-exact against the CPU
// oracle.  Replaces preprocessCUDA / computeCov2DCUDA of the un-vendored rasterizer the
// reference binds at custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_temporal.py:169-178
// (algorithm: SURVEY.md Appendix A.2 steps 1-9, A.3 "preprocess-bwd").
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
    const float* view_params;
    float* g_rec;
    unsigned int* g_rect;
    unsigned int* tile_count;
    int32_t* radii;
};

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

__device__ __forceinline__ bool project_gaussian(const float* __restrict__ vp, float px, float py, float pz,
                                                 float s0, float s1, float s2, float qr, float qx, float qy,
                                                 float qz, float focal_x, float focal_y, Proj& o) {
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
            o.S[a][b] = L[a][0] * L[b][0] + L[a][1] * L[b][1] + L[a][2] * L[b][2];
            o.S[b][a] = o.S[a][b];
        }
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

constexpr int SMEM_TILES = 4096;   // per-view tile histogram kept in shared memory (<= 1024x1024 px)

// Projects one (view, Gaussian); returns the packed tile rect (0 = culled).
__device__ __forceinline__ unsigned int preprocess_one(const PreArgs& a, long long idx) {
    const int v = (int)(idx / a.P);
    const int g = (int)(idx - (long long)v * a.P);
    const float* __restrict__ vp = a.view_params + (size_t)v * DM4D_VIEW_STRIDE;
    const long long set = min(max((long long)vp[DM4D_VIEW_SET], 0ll), (long long)a.n_sets - 1);   // never index outside the sets

    a.radii[idx] = 0;
    a.g_rect[idx] = 0u;

    const float* m = a.means3D + set * a.means3D_stride + (size_t)g * 3;
    const float* sc = a.scales + set * a.scales_stride + (size_t)g * 3;
    const float* q = a.rotations + set * a.rotations_stride + (size_t)g * 4;
    const float mod = vp[DM4D_VIEW_SCALE_MOD];
    const float focal_x = (float)a.W / (2.0f * vp[DM4D_VIEW_TANFOVX]);
    const float focal_y = (float)a.H / (2.0f * vp[DM4D_VIEW_TANFOVY]);

    Proj pr;
    if (!project_gaussian(vp, m[0], m[1], m[2], mod * sc[0], mod * sc[1], mod * sc[2], q[0], q[1], q[2], q[3],
                          focal_x, focal_y, pr))
        return 0u;

    const float det = pr.a * pr.c - pr.b * pr.b;
    if (det == 0.0f) return 0u;
    const float det_inv = 1.0f / det;
    const float conx = pr.c * det_inv, cony = -pr.b * det_inv, conz = pr.a * det_inv;

    const float mid = 0.5f * (pr.a + pr.c);
    const float sq = sqrtf(fmaxf(0.1f, mid * mid - det));
    const float lambda1 = mid + sq, lambda2 = mid - sq;
    const float my_radius = ceilf(3.0f * sqrtf(fmaxf(lambda1, lambda2)));
    const float ix = ndc2pix(pr.hx * pr.p_w, a.W), iy = ndc2pix(pr.hy * pr.p_w, a.H);

    const int rminx = min(a.gx, max(0, (int)((ix - my_radius) / (float)DM4D_TILE)));
    const int rminy = min(a.gy, max(0, (int)((iy - my_radius) / (float)DM4D_TILE)));
    const int rmaxx = min(a.gx, max(0, (int)((ix + my_radius + (float)(DM4D_TILE - 1)) / (float)DM4D_TILE)));
    const int rmaxy = min(a.gy, max(0, (int)((iy + my_radius + (float)(DM4D_TILE - 1)) / (float)DM4D_TILE)));
    if ((rmaxx - rminx) * (rmaxy - rminy) == 0) return 0u;

    a.radii[idx] = (int32_t)my_radius;
    const unsigned int rect = (unsigned)rminx | ((unsigned)rminy << 8) | ((unsigned)rmaxx << 16) | ((unsigned)rmaxy << 24);
    a.g_rect[idx] = rect;

    const float op = a.opacities[set * a.opacities_stride + g];
    // Contribution threshold on the quadratic form q(d) = conic.x dx^2 + 2 conic.y dx dy + conic.z dy^2:
    // alpha = op*exp(-q/2) can reach 1/255 only where q <= 2 ln(255 op).  Stored with a safety margin that
    // covers the render kernels' rounding; <= 0 means the Gaussian can never contribute.  sort_pack turns it
    // into the per-(instance, tile) cell mask.
    float ey = 0.f;
    {
        const float t = logf(255.0f * op);
        if (t > 0.f) ey = 2.0f * (t + 1e-4f) * 1.002f;
        if (!(ey == ey)) ey = 1e30f;
    }
    float4* rec = reinterpret_cast<float4*>(a.g_rec + (size_t)idx * a.rec);
    rec[0] = make_float4(ix, iy, conx, cony);
    rec[1] = make_float4(conz, op, ey, __int_as_float(g));
    const float* c0 = a.colors + set * a.colors_stride + (size_t)g * 3;
    if (a.channels <= 3) {
        rec[2] = make_float4(c0[0], c0[1], c0[2], pr.tz);
    } else {
        const float* c1 = a.colors2 + set * a.colors2_stride + (size_t)g * 3;
        rec[2] = make_float4(c0[0], c0[1], c0[2], c1[0]);
        rec[3] = make_float4(c1[1], c1[2], pr.tz, 0.f);
    }

    return rect;
}

// One thread per (view, Gaussian).  Instances per tile are counted in a per-block shared-memory histogram
// (a block's 256 consecutive Gaussians are mesh-coherent and hit a few dozen tiles) and flushed with one
// global atomic per touched tile; blocks that straddle two views or very large images count globally.
__global__ void __launch_bounds__(DM4D_BLOCK) preprocess_kernel(PreArgs a) {
    __shared__ unsigned int hist[SMEM_TILES];
    const long long total = (long long)a.n_views * a.P;
    const long long first = (long long)blockIdx.x * blockDim.x;
    const long long last = min(first + blockDim.x, total) - 1;
    const long long idx = first + threadIdx.x;
    const int v_first = (int)(first / a.P);
    const bool use_smem = a.tiles <= SMEM_TILES && v_first == (int)(last / a.P);
    if (use_smem) {
        for (int i = threadIdx.x; i < a.tiles; i += blockDim.x) hist[i] = 0u;
        __syncthreads();
    }
    const unsigned int rect = idx < total ? preprocess_one(a, idx) : 0u;
    if (rect) {
        const int minx = rect & 0xff, miny = (rect >> 8) & 0xff, maxx = (rect >> 16) & 0xff, maxy = rect >> 24;
        if (use_smem) {
            for (int y = miny; y < maxy; ++y)
                for (int x = minx; x < maxx; ++x) atomicAdd(&hist[y * a.gx + x], 1u);
        } else {
            unsigned int* cnt = a.tile_count + (size_t)(idx / a.P) * a.tiles;
            for (int y = miny; y < maxy; ++y)
                for (int x = minx; x < maxx; ++x) atomicAdd(&cnt[y * a.gx + x], 1u);
        }
    }
    if (use_smem) {
        __syncthreads();
        unsigned int* cnt = a.tile_count + (size_t)v_first * a.tiles;
        for (int i = threadIdx.x; i < a.tiles; i += blockDim.x) {
            const unsigned int c = hist[i];
            if (c) atomicAdd(&cnt[i], c);
        }
    }
}

// ------------------------------------------------------------------------------------------------

struct PreBwdArgs {
    PreArgs f;
    const float* accum;
    float* dmeans3D; int dmeans3D_atomic;
    float* dmeans2D;
    float* dcolors;  int dcolors_atomic;
    float* dcolors2; int dcolors2_atomic;
    float* dopac;    int dopac_atomic;
    float* dscales;  int dscales_atomic;
    float* drots;    int drots_atomic;
};

__device__ __forceinline__ void emit(float* p, float v, int atomic) {
    if (atomic) atomicAdd(p, v);
    else *p = v;
}

__global__ void __launch_bounds__(DM4D_BLOCK) preprocess_backward_kernel(PreBwdArgs b) {
    const PreArgs& a = b.f;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)a.n_views * a.P) return;
    const int v = (int)(idx / a.P);
    const int g = (int)(idx - (long long)v * a.P);
    const float* __restrict__ vp = a.view_params + (size_t)v * DM4D_VIEW_STRIDE;
    const long long set = min(max((long long)vp[DM4D_VIEW_SET], 0ll), (long long)a.n_sets - 1);
    const bool live = a.g_rect[idx] != 0u;
    const float* acc = b.accum + (size_t)idx * a.acc;

    float dm[3] = {0.f, 0.f, 0.f}, ds[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};
    float g2x = 0.f, g2y = 0.f, gop = 0.f;
    float gcol[DM4D_MAX_CHANNELS] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

    if (live) {
        const float4 a0 = reinterpret_cast<const float4*>(acc)[0];
        const float4 a1 = reinterpret_cast<const float4*>(acc)[1];
        g2x = a0.x; g2y = a0.y;
        const float gx = a0.z, gy = a0.w, gz = a1.x;
        gop = a1.y;
        const float gd = a1.z;
        for (int ch = 0; ch < a.channels; ++ch) gcol[ch] = acc[8 + ch];

        const float* m = a.means3D + set * a.means3D_stride + (size_t)g * 3;
        const float* sc = a.scales + set * a.scales_stride + (size_t)g * 3;
        const float* q = a.rotations + set * a.rotations_stride + (size_t)g * 4;
        const float mod = vp[DM4D_VIEW_SCALE_MOD];
        const float fx = (float)a.W / (2.0f * vp[DM4D_VIEW_TANFOVX]);
        const float fy = (float)a.H / (2.0f * vp[DM4D_VIEW_TANFOVY]);
        const float px = m[0], py = m[1], pz = m[2];
        const float s[3] = {mod * sc[0], mod * sc[1], mod * sc[2]};
        Proj pr;
        project_gaussian(vp, px, py, pz, s[0], s[1], s[2], q[0], q[1], q[2], q[3], fx, fy, pr);
        const float* V = vp;
        const float* PV = vp + 16;

        // (1) conic -> cov2D -> (Sigma, t)
        const float A_ = pr.a, B_ = pr.b, C_ = pr.c;
        const float denom = A_ * C_ - B_ * B_;
        const float d2 = 1.0f / (denom * denom + 0.0000001f);
        const float dL_da = d2 * (-C_ * C_ * gx + 2.f * B_ * C_ * gy + (denom - A_ * C_) * gz);
        const float dL_dc = d2 * (-A_ * A_ * gz + 2.f * A_ * B_ * gy + (denom - A_ * C_) * gx);
        const float dL_db = d2 * 2.f * (B_ * C_ * gx - (denom + 2.f * B_ * B_) * gy + A_ * B_ * gz);
        const float hb = 0.5f * dL_db;
        float GS[3][3], dA[2][3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                GS[j][k] = dL_da * pr.A[0][j] * pr.A[0][k] + hb * (pr.A[0][j] * pr.A[1][k] + pr.A[1][j] * pr.A[0][k]) +
                           dL_dc * pr.A[1][j] * pr.A[1][k];
            dA[0][j] = 2.f * (dL_da * pr.B[0][j] + hb * pr.B[1][j]);
            dA[1][j] = 2.f * (hb * pr.B[0][j] + dL_dc * pr.B[1][j]);
        }
        float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            dJ00 += dA[0][j] * V[0 + 4 * j];
            dJ02 += dA[0][j] * V[2 + 4 * j];
            dJ11 += dA[1][j] * V[1 + 4 * j];
            dJ12 += dA[1][j] * V[2 + 4 * j];
        }
        const float tz1 = 1.f / pr.tz, tz2 = tz1 * tz1, tz3 = tz2 * tz1;
        const float dL_dtx = pr.xmul * -fx * tz2 * dJ02;
        const float dL_dty = pr.ymul * -fy * tz2 * dJ12;
        const float dL_dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.f * fx * pr.cx) * tz3 * dJ02 +
                             (2.f * fy * pr.cy) * tz3 * dJ12;
#pragma unroll
        for (int j = 0; j < 3; ++j) dm[j] = V[0 + 4 * j] * dL_dtx + V[1 + 4 * j] * dL_dty + V[2 + 4 * j] * dL_dtz;

        // (2) 2D mean -> 3D mean, (3) depth -> 3D mean
        const float m_w = pr.p_w;
        const float mul1 = pr.hx * m_w * m_w, mul2 = pr.hy * m_w * m_w;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            dm[k] += (PV[0 + 4 * k] * m_w - PV[3 + 4 * k] * mul1) * g2x + (PV[1 + 4 * k] * m_w - PV[3 + 4 * k] * mul2) * g2y;
            dm[k] += (V[2 + 4 * k] - V[3 + 4 * k] * pr.tz) * gd;
        }

        // (5) Sigma = L L^T, L = Rm diag(s)
        float Rm[3][3], L[3][3], dLm[3][3], dR[3][3];
        quat_to_R(q[0], q[1], q[2], q[3], Rm);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) L[i][k] = Rm[i][k] * s[k];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) dLm[i][k] = 2.f * (GS[i][0] * L[0][k] + GS[i][1] * L[1][k] + GS[i][2] * L[2][k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < 3; ++i) { d += dLm[i][k] * Rm[i][k]; dR[i][k] = dLm[i][k] * s[k]; }
            ds[k] = mod * d;
        }
        const float r = q[0], x = q[1], y = q[2], z = q[3];
        dq[0] = 2.f * (z * (dR[1][0] - dR[0][1]) + y * (dR[0][2] - dR[2][0]) + x * (dR[2][1] - dR[1][2]));
        dq[1] = 2.f * (y * (dR[0][1] + dR[1][0]) + z * (dR[0][2] + dR[2][0]) + r * (dR[2][1] - dR[1][2])) - 4.f * x * (dR[1][1] + dR[2][2]);
        dq[2] = 2.f * (x * (dR[0][1] + dR[1][0]) + r * (dR[0][2] - dR[2][0]) + z * (dR[1][2] + dR[2][1])) - 4.f * y * (dR[0][0] + dR[2][2]);
        dq[3] = 2.f * (r * (dR[1][0] - dR[0][1]) + x * (dR[0][2] + dR[2][0]) + y * (dR[1][2] + dR[2][1])) - 4.f * z * (dR[0][0] + dR[1][1]);
    }

    if (b.dmeans2D) {
        float* o = b.dmeans2D + (size_t)idx * 3;
        o[0] = g2x; o[1] = g2y; o[2] = 0.f;
    }
    // With atomics the outputs were zero-filled by the launcher; dead Gaussians add nothing.
    if (b.dmeans3D && (live || !b.dmeans3D_atomic)) {
        float* o = b.dmeans3D + set * a.means3D_stride + (size_t)g * 3;
        for (int k = 0; k < 3; ++k) emit(o + k, dm[k], b.dmeans3D_atomic);
    }
    if (b.dscales && (live || !b.dscales_atomic)) {
        float* o = b.dscales + set * a.scales_stride + (size_t)g * 3;
        for (int k = 0; k < 3; ++k) emit(o + k, ds[k], b.dscales_atomic);
    }
    if (b.drots && (live || !b.drots_atomic)) {
        float* o = b.drots + set * a.rotations_stride + (size_t)g * 4;
        for (int k = 0; k < 4; ++k) emit(o + k, dq[k], b.drots_atomic);
    }
    if (b.dopac && (live || !b.dopac_atomic)) emit(b.dopac + set * a.opacities_stride + g, gop, b.dopac_atomic);
    if (b.dcolors && (live || !b.dcolors_atomic)) {
        float* o = b.dcolors + set * a.colors_stride + (size_t)g * 3;
        for (int k = 0; k < 3; ++k) emit(o + k, gcol[k], b.dcolors_atomic);
    }
    if (b.dcolors2 && a.channels > 3 && (live || !b.dcolors2_atomic)) {
        float* o = b.dcolors2 + set * a.colors2_stride + (size_t)g * 3;
        for (int k = 0; k < 3; ++k) emit(o + k, gcol[3 + k], b.dcolors2_atomic);
    }
}

PreArgs make_args(const dm4d_raster_desc* d, const RasterLayout& L, int32_t* radii) {
    PreArgs a;
    a.P = L.P; a.H = L.H; a.W = L.W; a.n_views = L.n_views; a.n_sets = d->n_sets; a.channels = L.channels;
    a.gx = L.gx; a.gy = L.gy; a.tiles = L.tiles; a.rec = L.rec; a.acc = L.acc;
    a.means3D = d->means3D; a.means3D_stride = d->means3D_stride;
    a.scales = d->scales; a.scales_stride = d->scales_stride;
    a.rotations = d->rotations; a.rotations_stride = d->rotations_stride;
    a.opacities = d->opacities; a.opacities_stride = d->opacities_stride;
    a.colors = d->colors; a.colors_stride = d->colors_stride;
    a.colors2 = d->colors2; a.colors2_stride = d->colors2_stride;
    a.view_params = d->view_params;
    a.g_rec = L.g_rec; a.g_rect = L.g_rect; a.tile_count = L.tile_count; a.radii = radii;
    return a;
}

}  // namespace

int launch_preprocess(const dm4d_raster_desc* d, const RasterLayout& L, int32_t* radii, cudaStream_t s) {
    const long long n = (long long)L.n_views * L.P;
    if (n == 0) return DM4D_OK;
    PreArgs a = make_args(d, L, radii);
    const unsigned blocks = (unsigned)((n + DM4D_BLOCK - 1) / DM4D_BLOCK);
    { KernelTimer kt(DM4D_K_PREPROCESS, s); preprocess_kernel<<<blocks, DM4D_BLOCK, 0, s>>>(a); }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

int launch_preprocess_backward(const dm4d_raster_desc* d, const RasterLayout& L, float* dL_dmeans3D,
                               float* dL_dmeans2D, float* dL_dcolors, float* dL_dcolors2, float* dL_dopacities,
                               float* dL_dscales, float* dL_drotations, cudaStream_t s) {
    const long long n = (long long)L.n_views * L.P;
    if (n == 0) return DM4D_OK;
    PreBwdArgs b;
    b.f = make_args(d, L, nullptr);
    b.accum = L.accum;
    // Outputs are fully overwritten. A set-strided attribute can be stored directly when the
    // caller promises a view<->set bijection (or there is a single view and a single set);
    // otherwise the output is zero-filled here and accumulated with atomics.
    const bool bijection = (d->flags & DM4D_RASTER_VIEWS_DISTINCT_SETS) && d->n_sets == d->n_views;
    auto mode = [&](float* p, long long stride, size_t elems_per_set) -> int {
        if (!p) return 0;
        const bool direct = (stride != 0 && bijection) || (d->n_views == 1 && (stride == 0 || d->n_sets == 1));
        if (!direct) {
            const size_t sets = stride == 0 ? 1 : (size_t)d->n_sets;
            cudaMemsetAsync(p, 0, sets * elems_per_set * sizeof(float), s);
        }
        return direct ? 0 : 1;
    };
    const size_t P = (size_t)L.P;
    b.dmeans3D = dL_dmeans3D; b.dmeans3D_atomic = mode(dL_dmeans3D, d->means3D_stride, P * 3);
    b.dmeans2D = dL_dmeans2D;
    b.dcolors = dL_dcolors;   b.dcolors_atomic = mode(dL_dcolors, d->colors_stride, P * 3);
    b.dcolors2 = dL_dcolors2; b.dcolors2_atomic = mode(dL_dcolors2, d->colors2_stride, P * 3);
    b.dopac = dL_dopacities;  b.dopac_atomic = mode(dL_dopacities, d->opacities_stride, P);
    b.dscales = dL_dscales;   b.dscales_atomic = mode(dL_dscales, d->scales_stride, P * 3);
    b.drots = dL_drotations;  b.drots_atomic = mode(dL_drotations, d->rotations_stride, P * 4);
    const unsigned blocks = (unsigned)((n + DM4D_BLOCK - 1) / DM4D_BLOCK);
    { KernelTimer kt(DM4D_K_PREPROCESS_BWD, s); preprocess_backward_kernel<<<blocks, DM4D_BLOCK, 0, s>>>(b); }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}
