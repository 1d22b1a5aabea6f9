// Per-(view, Gaussian) projection ("preprocess") forward.
//
// Compiled with -fmad=false: every + - * / sqrt below is a separately rounded IEEE binary32
// operation evaluated left to right, which is the arithmetic spec of DESIGN.md §4 and what
// makes the integer state (radii, tile rectangles, depth keys) bit-exact against the CPU
// oracle.  Replaces preprocessCUDA / computeCov2DCUDA of the un-vendored rasterizer the
// reference binds at custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_temporal.py:169-178
// (algorithm: SURVEY.md Appendix A.2 steps 1-9).  The backward lives in raster_preprocess_bwd.cu.
#include "raster_project.cuh"

namespace {

constexpr int SMEM_TILES = 4096;   // per-view tile histogram kept in shared memory (<= 1024x1024 px)

// Projects one (view, Gaussian); returns the packed tile rect (0 = culled).
template <bool COV>
__device__ __forceinline__ unsigned int preprocess_one(const PreArgs& a, long long idx, int v, int g, const float* __restrict__ vp) {
    const long long set = min(max((long long)vp[DM4D_VIEW_SET], 0ll), (long long)a.n_sets - 1);   // never index outside the sets

    a.radii[idx] = 0;
    a.g_rect[idx] = 0u;

    const float* m = a.means3D + set * a.means3D_stride + (size_t)g * 3;
    const float focal_x = (float)a.W / (2.0f * vp[DM4D_VIEW_TANFOVX]);
    const float focal_y = (float)a.H / (2.0f * vp[DM4D_VIEW_TANFOVY]);

    Proj pr;
    if (COV) {
        load_sigma(a.cov3D + set * a.cov3D_stride + (size_t)g * 6, pr.S);
    } else {
        const float* sc = a.scales + set * a.scales_stride + (size_t)g * 3;
        const float4 q = load_quat(a.rotations + set * a.rotations_stride + (size_t)g * 4);
        const float mod = vp[DM4D_VIEW_SCALE_MOD];
        gaussian_sigma(mod * sc[0], mod * sc[1], mod * sc[2], q.x, q.y, q.z, q.w, pr.S);
    }
    if (!project_with_sigma(vp, m[0], m[1], m[2], focal_x, focal_y, pr)) return 0u;

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
#ifndef DM4D_PRE_MIN_BLOCKS
#define DM4D_PRE_MIN_BLOCKS 6      // register budget (42): measured 4: 0.0895 ms, 5: 0.0827, 6: 0.0812 at C3
#endif
template <bool COV>
__global__ void __launch_bounds__(DM4D_BLOCK, DM4D_PRE_MIN_BLOCKS) preprocess_kernel(PreArgs a) {
    __shared__ unsigned int hist[SMEM_TILES];
    __shared__ ViewCache vcache;
    ViewRows vc;
    const long long total = (long long)a.n_views * a.P;
    const long long first = (long long)blockIdx.x * blockDim.x;
    const long long last = min(first + blockDim.x, total) - 1;
    const long long idx = first + threadIdx.x;
    const int v_first = (int)(first / a.P);
    const bool use_smem = a.tiles <= SMEM_TILES && v_first == (int)(last / a.P);
    vc.fill(&vcache, a.view_params, v_first, a.n_views);
    if (use_smem)
        for (int i = threadIdx.x; i < a.tiles; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    int v = 0, g = 0;
    if (idx < total) split_index(first, threadIdx.x, a.P, v, g);
    const unsigned int rect = idx < total ? preprocess_one<COV>(a, idx, v, g, vc.row(v)) : 0u;
    if (rect) {
        const int minx = rect & 0xff, miny = (rect >> 8) & 0xff, maxx = (rect >> 16) & 0xff, maxy = rect >> 24;
        if (use_smem) {
            for (int y = miny; y < maxy; ++y)
                for (int x = minx; x < maxx; ++x) atomicAdd(&hist[y * a.gx + x], 1u);
        } else {
            unsigned int* cnt = a.tile_count + (size_t)v * a.tiles;
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

}  // namespace

int launch_preprocess(const dm4d_raster_desc* d, const RasterLayout& L, int32_t* radii, cudaStream_t s) {
    const long long n = (long long)L.n_views * L.P;
    if (n == 0) return DM4D_OK;
    PreArgs a = make_pre_args(d, L, radii);
    const unsigned blocks = (unsigned)((n + DM4D_BLOCK - 1) / DM4D_BLOCK);
    {
        KernelTimer kt(DM4D_K_PREPROCESS, s);
        if (a.cov3D) preprocess_kernel<true><<<blocks, DM4D_BLOCK, 0, s>>>(a);
        else preprocess_kernel<false><<<blocks, DM4D_BLOCK, 0, s>>>(a);
    }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

