// "Reference-equivalent" GPU baseline — OUR transcription of the STRUCTURE of the rasterizer the
// reference binds (un-vendored diff-gaussian-rasterization, ashawkey fork; algorithm as restated in SURVEY.md
// Appendix A), NOT the upstream code and NOT part of the product: bench.py times it next to libdm4d.so on the same
// GPU so that the speed-up of the B200-native design can be stated against the classic pipeline.
//
// What follows upstream's pipeline, per view, one launch sequence per view:
//   preprocess -> InclusiveSum(tiles_touched) -> D2H read-back of num_rendered (host sync) -> duplicateWithKeys
//   -> ONE global 64-bit radix sort of (tile << 32 | depth bits, id) pairs (CUB) -> identifyTileRanges
//   -> renderCUDA forward: CTA per 16x16 tile, 256 instances fetched cooperatively per round, every thread walks
//      every instance of its tile (no sub-tile culling)
//   -> renderCUDA backward: same walk back to front, TEN global atomicAdds per (pixel, contributing instance)
//   -> preprocess backward.
// The per-Gaussian projection math (preprocess forward/backward, < 5 % of the time of either pipeline) is shared with
// the product (launch_preprocess / launch_preprocess_backward of libdm4d.so) so both arms render identical inputs;
// scratch buffers are persistent here (upstream re-sizes torch tensors per call) — both choices favour this baseline.
#include <cub/cub.cuh>
#include <cstdio>
#include "../dreammesh4d_b200/csrc/raster_internal.cuh"

namespace {

template <typename T>
struct Buf {                     // grow-only device buffer
    T* p = nullptr;
    size_t cap = 0;
    int reserve(size_t need) {
        if (need <= cap) return 0;
        if (p) cudaFree(p);
        cap = need + need / 4 + 1024;
        return cudaMalloc(&p, cap * sizeof(T)) == cudaSuccess ? 0 : -2;
    }
};

struct Scratch {
    Buf<unsigned int> tiles_touched, offsets, vals, vals_sorted;
    Buf<unsigned long long> keys, keys_sorted;
    Buf<uint2> ranges;
    Buf<unsigned char> cub_tmp;
} g;

#define REF_CHECK(expr)                                                                            \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) { fprintf(stderr, "ref_equiv: %s: %s\n", #expr, cudaGetErrorString(_e)); return -2; } \
    } while (0)

__global__ void tiles_touched_kernel(const unsigned int* __restrict__ rect, int P, unsigned int* __restrict__ tt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const unsigned int r = rect[i];
    tt[i] = ((r >> 16 & 0xff) - (r & 0xff)) * ((r >> 24) - (r >> 8 & 0xff));
}

__global__ void duplicate_with_keys(int P, const unsigned int* __restrict__ rect, const float* __restrict__ g_rec,
                                    int rec, int depth_idx, const unsigned int* __restrict__ offsets,
                                    unsigned long long* __restrict__ keys, unsigned int* __restrict__ vals, int gx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const unsigned int r = rect[i];
    if (!r) return;
    unsigned int off = i == 0 ? 0u : offsets[i - 1];
    const unsigned int dbits = __float_as_uint(g_rec[(size_t)i * rec + depth_idx]);
    for (int y = r >> 8 & 0xff; y < (int)(r >> 24); ++y)
        for (int x = r & 0xff; x < (int)(r >> 16 & 0xff); ++x) {
            keys[off] = ((unsigned long long)(y * gx + x) << 32) | dbits;
            vals[off] = (unsigned int)i;
            ++off;
        }
}

__global__ void identify_tile_ranges(int R, const unsigned long long* __restrict__ keys, uint2* __restrict__ ranges) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const unsigned int cur = (unsigned int)(keys[i] >> 32);
    if (i == 0) ranges[cur].x = 0;
    else {
        const unsigned int prev = (unsigned int)(keys[i - 1] >> 32);
        if (cur != prev) { ranges[prev].y = i; ranges[cur].x = i; }
    }
    if (i == R - 1) ranges[cur].y = R;
}

constexpr int BS = 256;

__global__ void __launch_bounds__(BS) render_fwd(const uint2* __restrict__ ranges, const unsigned int* __restrict__ point_list,
                                                 int W, int H, const float* __restrict__ g_rec, const float* __restrict__ bg,
                                                 unsigned int* __restrict__ n_contrib, float* __restrict__ out_color,
                                                 float* __restrict__ out_depth, float* __restrict__ out_alpha) {
    const int gx = (W + 15) / 16;
    const int px = blockIdx.x * 16 + threadIdx.x, py = blockIdx.y * 16 + threadIdx.y;
    const int tid = threadIdx.y * 16 + threadIdx.x;
    const bool inside = px < W && py < H;
    const float pfx = (float)px, pfy = (float)py;
    bool done = !inside;
    const uint2 range = ranges[blockIdx.y * gx + blockIdx.x];
    const int rounds = ((int)(range.y - range.x) + BS - 1) / BS;
    int todo = (int)(range.y - range.x);
    __shared__ int c_id[BS];
    __shared__ float2 c_xy[BS];
    __shared__ float4 c_co[BS];
    float T = 1.f, C[3] = {0.f, 0.f, 0.f}, D = 0.f, Wg = 0.f;
    unsigned int contributor = 0, last = 0;
    for (int i = 0; i < rounds; ++i, todo -= BS) {
        if (__syncthreads_count(done) == BS) break;
        const int progress = i * BS + tid;
        if (range.x + progress < range.y) {
            const int id = (int)point_list[range.x + progress];
            const float* r = g_rec + (size_t)id * 12;
            c_id[tid] = id;
            c_xy[tid] = make_float2(r[0], r[1]);
            c_co[tid] = make_float4(r[2], r[3], r[4], r[5]);
        }
        __syncthreads();
        for (int j = 0; !done && j < min(BS, todo); ++j) {
            ++contributor;
            const float2 xy = c_xy[j];
            const float4 co = c_co[j];
            const float dx = xy.x - pfx, dy = xy.y - pfy;
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.f) continue;
            const float alpha = fminf(0.99f, co.w * expf(power));
            if (alpha < 1.f / 255.f) continue;
            const float test_T = T * (1.f - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
            const float* r = g_rec + (size_t)c_id[j] * 12;
            const float w = alpha * T;
            C[0] += r[8] * w; C[1] += r[9] * w; C[2] += r[10] * w;
            D += r[11] * w;
            Wg += w;
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        const size_t npix = (size_t)H * W, pix = (size_t)py * W + px;
        n_contrib[pix] = last;
        for (int ch = 0; ch < 3; ++ch) out_color[ch * npix + pix] = C[ch] + T * bg[ch];
        out_depth[pix] = D;
        out_alpha[pix] = Wg;
    }
}

__global__ void __launch_bounds__(BS) render_bwd(const uint2* __restrict__ ranges, const unsigned int* __restrict__ point_list,
                                                 int W, int H, const float* __restrict__ g_rec, const float* __restrict__ bg,
                                                 const unsigned int* __restrict__ n_contrib, const float* __restrict__ out_alpha,
                                                 const float* __restrict__ dL_dC, const float* __restrict__ dL_dD,
                                                 const float* __restrict__ dL_dA, float* __restrict__ accum) {
    const int gx = (W + 15) / 16;
    const int px = blockIdx.x * 16 + threadIdx.x, py = blockIdx.y * 16 + threadIdx.y;
    const int tid = threadIdx.y * 16 + threadIdx.x;
    const bool inside = px < W && py < H;
    const float pfx = (float)px, pfy = (float)py;
    const size_t npix = (size_t)H * W, pix = (size_t)py * W + px;
    const uint2 range = ranges[blockIdx.y * gx + blockIdx.x];
    const int rounds = ((int)(range.y - range.x) + BS - 1) / BS;
    int todo = (int)(range.y - range.x);
    bool done = !inside;
    __shared__ int c_id[BS];
    __shared__ float2 c_xy[BS];
    __shared__ float4 c_co[BS];
    __shared__ float c_col[3 * BS];
    __shared__ float c_dep[BS];
    const float T_final = inside ? 1.f - out_alpha[pix] : 0.f;
    float T = T_final;
    unsigned int contributor = todo;
    const unsigned int last_contributor = inside ? n_contrib[pix] : 0u;
    float gC[3] = {0.f, 0.f, 0.f}, gD = 0.f, gA = 0.f;
    if (inside) {
        for (int ch = 0; ch < 3; ++ch) gC[ch] = dL_dC[ch * npix + pix];
        if (dL_dD) gD = dL_dD[pix];
        if (dL_dA) gA = dL_dA[pix];
    }
    float accum_rec[3] = {0.f, 0.f, 0.f}, last_color[3] = {0.f, 0.f, 0.f};
    float accum_d = 0.f, last_depth = 0.f, accum_a = 0.f, last_alpha = 0.f;
    const float bg_dot = bg[0] * gC[0] + bg[1] * gC[1] + bg[2] * gC[2];
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    for (int i = 0; i < rounds; ++i, todo -= BS) {
        __syncthreads();
        const int progress = i * BS + tid;
        if (range.x + progress < range.y) {
            const int id = (int)point_list[range.y - progress - 1];
            const float* r = g_rec + (size_t)id * 12;
            c_id[tid] = id;
            c_xy[tid] = make_float2(r[0], r[1]);
            c_co[tid] = make_float4(r[2], r[3], r[4], r[5]);
            c_col[tid] = r[8]; c_col[BS + tid] = r[9]; c_col[2 * BS + tid] = r[10];
            c_dep[tid] = r[11];
        }
        __syncthreads();
        for (int j = 0; !done && j < min(BS, todo); ++j) {
            --contributor;
            if (contributor >= last_contributor) continue;
            const float2 xy = c_xy[j];
            const float4 co = c_co[j];
            const float dx = xy.x - pfx, dy = xy.y - pfy;
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.f) continue;
            const float G = expf(power);
            const float alpha = fminf(0.99f, co.w * G);
            if (alpha < 1.f / 255.f) continue;
            T = T / (1.f - alpha);
            const float w = alpha * T;
            float* row = accum + (size_t)c_id[j] * 12;
            float dL_dalpha = 0.f;
            for (int ch = 0; ch < 3; ++ch) {
                const float c = c_col[ch * BS + j];
                accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                last_color[ch] = c;
                dL_dalpha += (c - accum_rec[ch]) * gC[ch];
                atomicAdd(row + 8 + ch, w * gC[ch]);
            }
            accum_d = last_alpha * last_depth + (1.f - last_alpha) * accum_d;
            last_depth = c_dep[j];
            dL_dalpha += (last_depth - accum_d) * gD;
            atomicAdd(row + 6, w * gD);
            accum_a = last_alpha + (1.f - last_alpha) * accum_a;
            dL_dalpha += (1.f - accum_a) * gA;
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
            const float dL_dG = co.w * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_ddelx = -gdx * co.x - gdy * co.y;
            const float dG_ddely = -gdy * co.z - gdx * co.y;
            atomicAdd(row + 0, dL_dG * dG_ddelx * ddelx_dx);
            atomicAdd(row + 1, dL_dG * dG_ddely * ddely_dy);
            atomicAdd(row + 2, -0.5f * gdx * dx * dL_dG);
            atomicAdd(row + 3, -0.5f * gdx * dy * dL_dG);
            atomicAdd(row + 4, -0.5f * gdy * dy * dL_dG);
            atomicAdd(row + 5, G * dL_dalpha);
        }
    }
}

int check_single_view(const dm4d_raster_desc* d) {
    if (!d || d->n_views != 1 || d->channels != 3) { fprintf(stderr, "ref_equiv: one view, 3 channels per call\n"); return -1; }
    return 0;
}

}  // namespace

// One view forward.  `d` is a product descriptor with n_views == 1 (geom / img / bwd workspaces as for libdm4d; the
// bin workspace is only used for its tile_count section).  Returns num_rendered through the host pointer after a
// stream synchronisation, as upstream does.
extern "C" int refeq_forward(const dm4d_raster_desc* d, float* out_color, float* out_depth, float* out_alpha,
                             int32_t* radii, int64_t* num_rendered_host, void* stream) {
    if (check_single_view(d)) return -1;
    RasterLayout L;
    if (int rc = raster_make_layout(d, &L)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int P = L.P;
    REF_CHECK(cudaMemsetAsync(L.tile_count, 0, (size_t)L.tiles * sizeof(unsigned int), s));
    if (int rc = launch_preprocess(d, L, radii, s)) return rc;
    if (g.tiles_touched.reserve(P) || g.offsets.reserve(P) || g.ranges.reserve(L.tiles)) return -2;
    tiles_touched_kernel<<<(P + 255) / 256, 256, 0, s>>>(L.g_rect, P, g.tiles_touched.p);
    size_t need = 0;
    cub::DeviceScan::InclusiveSum(nullptr, need, g.tiles_touched.p, g.offsets.p, P, s);
    if (g.cub_tmp.reserve(need)) return -2;
    REF_CHECK(cub::DeviceScan::InclusiveSum(g.cub_tmp.p, need, g.tiles_touched.p, g.offsets.p, P, s));
    unsigned int R = 0;
    REF_CHECK(cudaMemcpyAsync(&R, g.offsets.p + P - 1, 4, cudaMemcpyDeviceToHost, s));
    REF_CHECK(cudaStreamSynchronize(s));                                  // upstream's num_rendered read-back
    if (num_rendered_host) *num_rendered_host = R;
    REF_CHECK(cudaMemsetAsync(g.ranges.p, 0, (size_t)L.tiles * sizeof(uint2), s));
    if (R > 0) {
        if (g.keys.reserve(R) || g.keys_sorted.reserve(R) || g.vals.reserve(R) || g.vals_sorted.reserve(R)) return -2;
        duplicate_with_keys<<<(P + 255) / 256, 256, 0, s>>>(P, L.g_rect, L.g_rec, L.rec, rec_depth_index(3), g.offsets.p,
                                                          g.keys.p, g.vals.p, L.gx);
        int bits = 0;
        while ((1 << bits) < L.tiles) ++bits;
        need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, g.keys.p, g.keys_sorted.p, g.vals.p, g.vals_sorted.p, (int)R, 0, 32 + bits, s);
        if (g.cub_tmp.reserve(need)) return -2;
        REF_CHECK(cub::DeviceRadixSort::SortPairs(g.cub_tmp.p, need, g.keys.p, g.keys_sorted.p, g.vals.p, g.vals_sorted.p,
                                                  (int)R, 0, 32 + bits, s));
        identify_tile_ranges<<<(R + 255) / 256, 256, 0, s>>>((int)R, g.keys_sorted.p, g.ranges.p);
    } else if (g.vals_sorted.reserve(1)) return -2;
    const float* bg = d->view_params + DM4D_VIEW_BG;
    render_fwd<<<dim3(L.gx, L.gy), dim3(16, 16), 0, s>>>(g.ranges.p, g.vals_sorted.p, L.W, L.H, L.g_rec, bg, L.n_contrib,
                                                        out_color, out_depth, out_alpha);
    REF_CHECK(cudaGetLastError());
    return 0;
}

// Backward of the view rendered by the LAST refeq_forward call (its sorted list and ranges are still in the scratch).
extern "C" int refeq_backward(const dm4d_raster_desc* d, const float* out_alpha, const float* dL_dcolor,
                              const float* dL_ddepth, const float* dL_dalpha, float* dL_dmeans3D, float* dL_dmeans2D,
                              float* dL_dcolors, float* dL_dopacities, float* dL_dscales, float* dL_drotations,
                              void* stream) {
    if (check_single_view(d)) return -1;
    RasterLayout L;
    if (int rc = raster_make_layout(d, &L)) return rc;
    if (!d->bwd) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    REF_CHECK(cudaMemsetAsync(L.accum, 0, (size_t)L.P * 12 * sizeof(float), s));
    const float* bg = d->view_params + DM4D_VIEW_BG;
    render_bwd<<<dim3(L.gx, L.gy), dim3(16, 16), 0, s>>>(g.ranges.p, g.vals_sorted.p, L.W, L.H, L.g_rec, bg, L.n_contrib, out_alpha,
                                                        dL_dcolor, dL_ddepth, dL_dalpha, L.accum);
    REF_CHECK(cudaGetLastError());
    return launch_preprocess_backward(d, L, dL_dmeans3D, dL_dmeans2D, dL_dcolors, nullptr, dL_dopacities, dL_dscales,
                                      dL_drotations, s);
}
