// Fused per-view image post-ops that follow the rasterizer, forward and backward.
//
// Replaces ~30 element-wise / convolution / boolean-index launches per view of DiffGaussian.forward
// (custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_temporal.py:180-193,212-218,229; Depth2Normal
// :25-54; static twin diff_sugar_rasterizer_normal.py:172-206) and the [B,H,W,C] stack/permute of
// GaussianBatchRenderer.batch_forward (renderer/gaussian_batch_renderer.py:78-122):
//
//   mask  = alpha > 0.99
//   depth = depth, gradient only inside the mask
//   xyz   = rays_o + depth * rays_d;  n = -cross(xyz(x+1) - xyz(x-1), xyz(y+1) - xyz(y-1))   (zero padding)
//   normal_from_dist = normalize(n) * 0.5 * alpha + 0.5,       gradient only inside the mask
//   normal           = normalize(rendered normals) * 0.5 * alpha + 0.5,  gradient only inside the mask
//   render           = clamp(rgb, 0, 1)
//
// Inputs are the rasterizer's own tensors (planar [B,C,H,W]); outputs are written directly in the [B,H,W,C]
// layout the system consumes.  One thread per pixel, HBM-bound: forward reads 4*(6+1+1) + 24 B and writes
// 4*(3+3+3+1+1) B per pixel (100 B); the backward is two passes (per-pixel terms + the stencil's cross-pixel
// gather through a 24 B/pixel scratch) so that it needs no atomics and is bit-reproducible.
#include "raster_internal.cuh"

namespace {

struct PostArgs {
    int B, H, W, flags;
    const float* color;     // [B,6,H,W]  rgb + rendered normals
    const float* depth;     // [B,1,H,W]
    const float* alpha;     // [B,1,H,W]
    const float* rays_o;    // [B,H,W,3]
    const float* rays_d;    // [B,H,W,3]
};

__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float3 ld3(const float* p) { return f3(p[0], p[1], p[2]); }

constexpr float NORM_EPS = 1e-12f;      // torch.nn.functional.normalize default

// y = x / max(|x|, eps)
__device__ __forceinline__ float3 normalize3(float3 x, float& len) {
    len = sqrtf(dot3(x, x));
    return (1.0f / fmaxf(len, NORM_EPS)) * x;
}
// gradient of normalize3 w.r.t. x given g = dL/dy and y
__device__ __forceinline__ float3 normalize3_bwd(float3 y, float len, float3 g) {
    if (len > NORM_EPS) return (1.0f / len) * (g - dot3(y, g) * y);
    return (1.0f / NORM_EPS) * g;
}

// position of pixel (x, y) of view b, or 0 outside the image (the convolution's zero padding)
__device__ __forceinline__ float3 xyz_at(const PostArgs& a, int b, int x, int y) {
    if (x < 0 || y < 0 || x >= a.W || y >= a.H) return f3(0.f, 0.f, 0.f);
    const size_t p = ((size_t)b * a.H + y) * a.W + x;
    return ld3(a.rays_o + 3 * p) + a.depth[p] * ld3(a.rays_d + 3 * p);
}

struct Stencil { float3 dx, dy, n, nn; float len; };
__device__ __forceinline__ Stencil stencil_at(const PostArgs& a, int b, int x, int y) {
    Stencil s;
    s.dx = xyz_at(a, b, x + 1, y) - xyz_at(a, b, x - 1, y);
    s.dy = xyz_at(a, b, x, y + 1) - xyz_at(a, b, x, y - 1);
    s.n = -1.0f * cross3(s.dx, s.dy);
    s.nn = normalize3(s.n, s.len);
    return s;
}

__global__ void __launch_bounds__(DM4D_BLOCK) postops_forward_kernel(PostArgs a, float* __restrict__ comp_rgb,
                                                                     float* __restrict__ comp_normal,
                                                                     float* __restrict__ comp_nfd,
                                                                     float* __restrict__ comp_depth,
                                                                     float* __restrict__ comp_mask) {
    const size_t npix = (size_t)a.H * a.W;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)a.B * npix) return;
    const int b = (int)(p / npix);
    const int r = (int)(p - (size_t)b * npix), y = r / a.W, x = r - y * a.W;
    const float* col = a.color + (size_t)b * 6 * npix + r;
    const float al = a.alpha[p];
    comp_rgb[3 * p + 0] = fminf(fmaxf(col[0], 0.f), 1.f);
    comp_rgb[3 * p + 1] = fminf(fmaxf(col[npix], 0.f), 1.f);
    comp_rgb[3 * p + 2] = fminf(fmaxf(col[2 * npix], 0.f), 1.f);
    float len;
    const float3 nn = normalize3(f3(col[3 * npix], col[4 * npix], col[5 * npix]), len);
    const float h = 0.5f * al;
    comp_normal[3 * p + 0] = nn.x * h + 0.5f;
    comp_normal[3 * p + 1] = nn.y * h + 0.5f;
    comp_normal[3 * p + 2] = nn.z * h + 0.5f;
    comp_depth[p] = a.depth[p];
    comp_mask[p] = al;
    if (comp_nfd) {
        const Stencil s = stencil_at(a, b, x, y);
        comp_nfd[3 * p + 0] = s.nn.x * h + 0.5f;
        comp_nfd[3 * p + 1] = s.nn.y * h + 0.5f;
        comp_nfd[3 * p + 2] = s.nn.z * h + 0.5f;
    }
}

// Pass A: everything that stays inside the pixel; the stencil's gradients w.r.t. its two difference vectors go
// to `scratch` [B,H,W,6] for pass B.
__global__ void __launch_bounds__(DM4D_BLOCK) postops_backward_a_kernel(
    PostArgs a, const float* __restrict__ g_rgb, const float* __restrict__ g_normal, const float* __restrict__ g_nfd,
    const float* __restrict__ g_depth, const float* __restrict__ g_mask, float* __restrict__ scratch,
    float* __restrict__ d_color, float* __restrict__ d_depth, float* __restrict__ d_alpha) {
    const size_t npix = (size_t)a.H * a.W;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)a.B * npix) return;
    const int b = (int)(p / npix);
    const int r = (int)(p - (size_t)b * npix), y = r / a.W, x = r - y * a.W;
    const float* col = a.color + (size_t)b * 6 * npix + r;
    float* dcol = d_color + (size_t)b * 6 * npix + r;
    const float al = a.alpha[p];
    const bool m = al > 0.99f;
    float dal = g_mask ? g_mask[p] : 0.f;

    // clamp(0,1): gradient passes where 0 <= x <= 1
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = col[c * npix];
        dcol[c * npix] = (g_rgb && v >= 0.f && v <= 1.f) ? g_rgb[3 * p + c] : 0.f;
    }
    // rendered normals
    float3 dn = f3(0.f, 0.f, 0.f);
    if (m && g_normal) {
        float len;
        const float3 nn = normalize3(f3(col[3 * npix], col[4 * npix], col[5 * npix]), len);
        const float3 g = ld3(g_normal + 3 * p);
        dal += 0.5f * dot3(nn, g);
        dn = normalize3_bwd(nn, len, (0.5f * al) * g);
    }
    dcol[3 * npix] = dn.x; dcol[4 * npix] = dn.y; dcol[5 * npix] = dn.z;
    // normal from distance: n = -cross(dx, dy)  =>  dL/d dx = g x dy,  dL/d dy = dx x g
    if (scratch) {
        float3 gdx = f3(0.f, 0.f, 0.f), gdy = gdx;
        if (m && g_nfd) {
            const Stencil s = stencil_at(a, b, x, y);
            const float3 g = ld3(g_nfd + 3 * p);
            dal += 0.5f * dot3(s.nn, g);
            const float3 gn = normalize3_bwd(s.nn, s.len, (0.5f * al) * g);
            gdx = cross3(gn, s.dy);
            gdy = cross3(s.dx, gn);
        }
        float* sp = scratch + 6 * p;
        sp[0] = gdx.x; sp[1] = gdx.y; sp[2] = gdx.z; sp[3] = gdy.x; sp[4] = gdy.y; sp[5] = gdy.z;
    } else {
        d_depth[p] = (m && g_depth) ? g_depth[p] : 0.f;
    }
    d_alpha[p] = dal;
}

// Pass B: xyz(q) enters d/dx of its left/right neighbours and d/dy of its upper/lower neighbours.
__global__ void __launch_bounds__(DM4D_BLOCK) postops_backward_b_kernel(PostArgs a, const float* __restrict__ g_depth,
                                                                        const float* __restrict__ scratch,
                                                                        float* __restrict__ d_depth) {
    const size_t npix = (size_t)a.H * a.W;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)a.B * npix) return;
    const int b = (int)(p / npix);
    const int r = (int)(p - (size_t)b * npix), y = r / a.W, x = r - y * a.W;
    const bool m = a.alpha[p] > 0.99f;
    float d = (m && g_depth) ? g_depth[p] : 0.f;
    // temporal renderer: the depth is detached outside the mask BEFORE the position map is built (:181,187);
    // static renderer: after (normal.py:176,203), so the stencil gradient reaches unmasked pixels too
    if (m || (a.flags & DM4D_POSTOPS_STATIC)) {
        float3 gx = f3(0.f, 0.f, 0.f);
        if (x > 0) gx = gx + ld3(scratch + 6 * (p - 1));                 // + d/dx of the left neighbour
        if (x + 1 < a.W) gx = gx - ld3(scratch + 6 * (p + 1));           // - d/dx of the right neighbour
        if (y > 0) gx = gx + ld3(scratch + 6 * (p - a.W) + 3);           // + d/dy of the upper neighbour
        if (y + 1 < a.H) gx = gx - ld3(scratch + 6 * (p + a.W) + 3);     // - d/dy of the lower neighbour
        d += dot3(gx, ld3(a.rays_d + 3 * p));
    }
    d_depth[p] = d;
}

int check_desc(const dm4d_postops_desc* d, PostArgs* a) {
    if (!d || d->n_views <= 0 || d->H <= 0 || d->W <= 0 || !d->color6 || !d->depth || !d->alpha) {
        dm4d_set_error("dm4d_postops: bad descriptor");
        return DM4D_EINVAL;
    }
    if ((d->flags & DM4D_POSTOPS_NORMAL_FROM_DIST) && (!d->rays_o || !d->rays_d)) {
        dm4d_set_error("dm4d_postops: normal_from_dist needs rays_o and rays_d");
        return DM4D_EINVAL;
    }
    a->B = d->n_views; a->H = d->H; a->W = d->W; a->flags = d->flags;
    a->color = d->color6; a->depth = d->depth; a->alpha = d->alpha; a->rays_o = d->rays_o; a->rays_d = d->rays_d;
    return DM4D_OK;
}

}  // namespace

extern "C" int dm4d_postops_forward(const dm4d_postops_desc* d, float* comp_rgb, float* comp_normal,
                                    float* comp_normal_from_dist, float* comp_depth, float* comp_mask, void* stream) {
    PostArgs a;
    if (int rc = check_desc(d, &a)) return rc;
    if (!comp_rgb || !comp_normal || !comp_depth || !comp_mask ||
        (((d->flags & DM4D_POSTOPS_NORMAL_FROM_DIST) != 0) != (comp_normal_from_dist != nullptr))) {
        dm4d_set_error("dm4d_postops_forward: output pointers do not match the flags");
        return DM4D_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)a.B * a.H * a.W;
    {
        KernelTimer kt(DM4D_K_POSTOPS_FWD, s);
        postops_forward_kernel<<<(unsigned)((n + DM4D_BLOCK - 1) / DM4D_BLOCK), DM4D_BLOCK, 0, s>>>(
            a, comp_rgb, comp_normal, comp_normal_from_dist, comp_depth, comp_mask);
    }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

extern "C" int dm4d_postops_backward(const dm4d_postops_desc* d, const float* g_rgb, const float* g_normal,
                                     const float* g_normal_from_dist, const float* g_depth, const float* g_mask,
                                     float* scratch, float* d_color6, float* d_depth, float* d_alpha, void* stream) {
    PostArgs a;
    if (int rc = check_desc(d, &a)) return rc;
    const bool nfd = (d->flags & DM4D_POSTOPS_NORMAL_FROM_DIST) != 0;
    if (!d_color6 || !d_depth || !d_alpha || (nfd && !scratch)) {
        dm4d_set_error("dm4d_postops_backward: NULL output (scratch [B,H,W,6] is required with normal_from_dist)");
        return DM4D_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)a.B * a.H * a.W;
    const unsigned blocks = (unsigned)((n + DM4D_BLOCK - 1) / DM4D_BLOCK);
    {
        KernelTimer kt(DM4D_K_POSTOPS_BWD, s);
        postops_backward_a_kernel<<<blocks, DM4D_BLOCK, 0, s>>>(a, g_rgb, g_normal, nfd ? g_normal_from_dist : nullptr,
                                                                g_depth, g_mask, nfd ? scratch : nullptr, d_color6,
                                                                d_depth, d_alpha);
        if (nfd) postops_backward_b_kernel<<<blocks, DM4D_BLOCK, 0, s>>>(a, g_depth, scratch, d_depth);
    }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}
