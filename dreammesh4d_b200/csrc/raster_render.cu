// Per-tile alpha compositing: forward and backward.
//
// Replaces renderCUDA (forward.cu / backward.cu) of the un-vendored rasterizer bound by the
// reference at custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_temporal.py:169-178,
// 202-211 (algorithm: SURVEY.md Appendix A.2 "render", A.3 "render-bwd"), with ashawkey's depth
// and alpha outputs.  One CTA per (view, tile), one thread per pixel, all views in one launch.
//
// B200 design: the depth-sorted instances of a tile are a contiguous run of packed 48/64-byte
// records (written by sort_pack_kernel).  Every WARP owns an 8x4 pixel block made of two 4x4 CELLS, one
// per HALF-WARP, and streams the tile's records through its own small shared-memory ring with TMA bulk
// copies (cp.async.bulk + mbarrier complete_tx, issued by the warp's lane 0): warps never meet at a
// block barrier, so a warp whose block is light (or saturated) runs ahead / retires while its
// neighbours keep compositing.  Surface-bound Gaussians are tiny (about 10 contributing pixels per
// (Gaussian, tile) instance at C3), so the unit of culling is the 4x4 cell: per chunk of 64 staged
// instances the lanes ballot the instances' 16-bit cell masks and each half-warp walks its OWN queue of
// instances that can reach its cell — the two halves run in lockstep on different instances, which cuts
// the walked (warp, instance) iterations from 1.91 (16x2 strips) to about 1.2 per instance.  The
// backward walks the same stream back to front, reduces the per-pixel partial gradients across the 16
// lanes of the half-warp with a recursive-halving shuffle reduction (11 shuffles for the 10 sums of the
// 3-channel pass, 15 for 6 channels; only when a lane contributes) and issues ONE coalesced global
// reduction (RED.ADD.F32, one accumulator row) per (half-warp, instance) instead of upstream's ten
// global atomics per (pixel, instance).
#include "raster_internal.cuh"

namespace {

// One thread per pixel of a 16x16 tile: 8 warps = 8 blocks of 8x4 = 16 cells of 4x4.  The warps of a tile never
// synchronise, so a tile may be split over PARTS CTAs of RWARPS warps each: smaller CTAs free their SM slot as soon
// as their own warps retire instead of waiting for the slowest block of the tile (measured at C3, step time:
// 8 warps 2.42 ms, 4: 2.32, 2: 2.32, 1: 2.26 -> one warp per CTA).
#ifndef DM4D_RENDER_WARPS
#define DM4D_RENDER_WARPS 1
#endif
constexpr int WARPS = DM4D_RENDER_WARPS;      // warps per CTA
constexpr int THREADS = 32 * WARPS;
constexpr int PARTS = 8 / WARPS;              // CTAs per tile
static_assert(WARPS == 1 || WARPS == 2 || WARPS == 4 || WARPS == 8, "DM4D_RENDER_WARPS must divide 8");
#ifndef DM4D_WCHUNK
#define DM4D_WCHUNK 64
#endif
#ifndef DM4D_WSTAGES
#define DM4D_WSTAGES 2
#endif
#ifndef DM4D_BWD_MIN_WARPS
#define DM4D_BWD_MIN_WARPS 32     // resident warps per SM the backward's register budget is sized for
#endif
#define DM4D_BWD_MIN_BLOCKS (DM4D_BWD_MIN_WARPS / DM4D_RENDER_WARPS)
constexpr int WCHUNK = DM4D_WCHUNK;     // instances per per-warp stage
constexpr int WSTAGES = DM4D_WSTAGES;   // per-warp ring depth
#ifndef DM4D_FWD_UNROLL
#define DM4D_FWD_UNROLL 4
#endif
constexpr int FWD_UNROLL = DM4D_FWD_UNROLL;
static_assert(WCHUNK == 64, "the per-half-warp candidate queue covers one staged chunk of 64 instances");

// exp of the Gaussian exponent.  expf (2 ulp) by default: DM4D_FAST_EXP switches to ex2.approx(x * log2 e), 6
// instructions shorter (-4 % step time) but 2 + |1.17 x| ulp; at C4 (about 100 blended Gaussians per pixel) the
// cancellation in (colour - accumulated colour) amplifies that to ~1e-3 of the largest gradient, the edge of the parity
// tolerance (scripts/c4_errors.py), so it stays opt-in.  Forward and backward use the same function, so their
// contribution decisions agree bit for bit either way.
#ifdef DM4D_FAST_EXP
__device__ __forceinline__ float gauss_exp(float x) { return __expf(x); }
#else
__device__ __forceinline__ float gauss_exp(float x) { return expf(x); }
#endif
__device__ __forceinline__ float fast_rcp(float x) {
    float r;      // MUFU.RCP, 1 ulp: no measurable effect on the gradient error (scripts/c4_errors.py)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

#if DM4D_CELL_ROWS == 4
// Pixel owned by a thread: warp w covers the 8x4 block (w & 1, w >> 1); its half-warp h covers the 4x4 cell
// (cx, cy) = (2 (w & 1) + h, w >> 1), whose bit in an instance's cell mask is cy * 4 + cx = 2 w + h.
struct PixelMap {
    int px, py, cell_bit, half, li;
    __device__ __forceinline__ PixelMap(int tile_x, int tile_y, int warp, int lane) {
        half = lane >> 4;
        li = lane & 15;
        px = tile_x * DM4D_TILE + (((warp & 1) << 1 | half) << 2) + (li & 3);
        py = tile_y * DM4D_TILE + ((warp >> 1) << 2) + (li >> 2);
        cell_bit = 2 * warp + half;
    }
};
#else
// EXPERIMENTAL 4x2 cells: warp w covers the 8x4 block (w & 1, w >> 1); its quarter-warp g = lane / 8 covers the 4x2
// cell (cx, cy8) = (2 (w & 1) + (g & 1), 2 (w >> 1) + (g >> 1)) of the tile's 4 x 8 cells; mask bit 4 cy8 + cx.
// `half` holds the group index g, `li` the lane within the group (0..7).
struct PixelMap {
    int px, py, cell_bit, half, li;
    __device__ __forceinline__ PixelMap(int tile_x, int tile_y, int warp, int lane) {
        half = lane >> 3;
        li = lane & 7;
        const int cx = ((warp & 1) << 1) | (half & 1), cy8 = ((warp >> 1) << 1) | (half >> 1);
        px = tile_x * DM4D_TILE + (cx << 2) + (li & 3);
        py = tile_y * DM4D_TILE + (cy8 << 1) + (li >> 2);
        cell_bit = 4 * cy8 + cx;
    }
};
#endif

// Candidate queue of this lane's half-warp for one staged chunk of 64 instances: bit j set <=> instance j of the chunk
// can reach the half-warp's cell.  Four warp ballots (two cells x two 32-instance groups); uniform within a half-warp.
// Kept as two 32-bit words: a pop is FLO + shift + mask on one word instead of 64-bit arithmetic.
struct CellQueue {
    unsigned int lo, hi;
    __device__ __forceinline__ bool empty() const { return (lo | hi) == 0u; }
    __device__ __forceinline__ void clear() { lo = hi = 0u; }
    // front to back (forward): lowest set bit, or -1
    __device__ __forceinline__ int pop_front() {
        const bool in_lo = lo != 0u;
        unsigned int cur = in_lo ? lo : hi;
        const int j = cur ? (__ffs((int)cur) - 1 + (in_lo ? 0 : 32)) : -1;
        cur &= cur - 1u;
        if (in_lo) lo = cur; else hi = cur;
        return j;
    }
    // back to front (backward): highest set bit; the queue must not be empty
    __device__ __forceinline__ int pop_back() {
        const bool in_hi = hi != 0u;
        unsigned int cur = in_hi ? hi : lo;
        const int b = 31 - __clz((int)cur);
        cur &= ~(1u << b);
        if (in_hi) hi = cur; else lo = cur;
        return b + (in_hi ? 32 : 0);
    }
};

#if DM4D_CELL_ROWS == 4
template <int R4>
__device__ __forceinline__ CellQueue cell_queue(const float4* r, int cnt, int lane, int warp, int half) {
    const unsigned int m0 = lane < cnt ? __float_as_uint(r[lane * R4 + 1].z) : 0u;
    const unsigned int m1 = lane + 32 < cnt ? __float_as_uint(r[(lane + 32) * R4 + 1].z) : 0u;
    const int sh = 2 * warp;
    const unsigned int a0 = __ballot_sync(0xffffffffu, (m0 >> sh) & 1u), b0 = __ballot_sync(0xffffffffu, (m0 >> (sh + 1)) & 1u);
    const unsigned int a1 = __ballot_sync(0xffffffffu, (m1 >> sh) & 1u), b1 = __ballot_sync(0xffffffffu, (m1 >> (sh + 1)) & 1u);
    CellQueue q;
    q.lo = half ? b0 : a0;
    q.hi = half ? b1 : a1;
    return q;
}
#else
// EXPERIMENTAL 4x2 cells: four queues per warp (one per quarter-warp), eight ballots per chunk.
template <int R4>
__device__ __forceinline__ CellQueue cell_queue(const float4* r, int cnt, int lane, int warp, int grp) {
    const unsigned int m0 = lane < cnt ? __float_as_uint(r[lane * R4 + 1].z) : 0u;
    const unsigned int m1 = lane + 32 < cnt ? __float_as_uint(r[(lane + 32) * R4 + 1].z) : 0u;
    CellQueue q;
    q.lo = q.hi = 0u;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int bit = 4 * ((((warp >> 1) << 1) | (g >> 1))) + (((warp & 1) << 1) | (g & 1));
        const unsigned int lo = __ballot_sync(0xffffffffu, (m0 >> bit) & 1u), hi = __ballot_sync(0xffffffffu, (m1 >> bit) & 1u);
        if (g == grp) { q.lo = lo; q.hi = hi; }
    }
    return q;
}
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

template <int C>
struct RecTraits {
    static constexpr int REC = C <= 3 ? 12 : 16;   // floats per record
    static constexpr int R4 = REC / 4;
    static constexpr int ACC = C <= 3 ? 12 : 16;   // floats per accumulator row
};

// features + depth of one record (layout: raster_internal.cuh)
template <int C>
__device__ __forceinline__ void load_features(const float4* r, float (&f)[C], float& depth) {
    const float4 c0 = r[2];
    f[0] = c0.x; f[1] = c0.y; f[2] = c0.z;
    if constexpr (C > 3) {
        const float4 c1 = r[3];
        f[3] = c0.w; f[4] = c1.x; f[5] = c1.y;
        depth = c1.z;
    } else {
        depth = c0.w;
    }
}

// Per-warp streaming ring -------------------------------------------------------------------------
template <int C>
struct WarpRing {
    using TR = RecTraits<C>;
    float4* buf;        // [WSTAGES][WCHUNK * R4]
    uint64_t* full;     // [WSTAGES]
    const float* stream;
    int n;              // instances available in the tile
    __device__ __forceinline__ void init(unsigned char* smem_raw, int warp, int lane, const float* stream_, int n_) {
        buf = reinterpret_cast<float4*>(smem_raw) + (size_t)warp * WSTAGES * WCHUNK * TR::R4;
        full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)WARPS * WSTAGES * WCHUNK * TR::REC * sizeof(float)) + warp * WSTAGES;
        stream = stream_;
        n = n_;
        if (lane == 0) {
            for (int s = 0; s < WSTAGES; ++s) mbar_init(&full[s], 1);
            fence_mbar_init();
        }
        __syncwarp();
    }
    // lane 0 only: stage chunk `c` (instances [c*WCHUNK, ...)) into slot `slot`
    __device__ __forceinline__ void issue(int c, int slot) {
        const int cnt = min(WCHUNK, n - c * WCHUNK);
        const uint32_t bytes = (uint32_t)cnt * TR::REC * sizeof(float);
        mbar_expect_tx(&full[slot], bytes);
        bulk_g2s(buf + (size_t)slot * WCHUNK * TR::R4, stream + (size_t)c * WCHUNK * TR::REC, bytes, &full[slot]);
    }
    __device__ __forceinline__ const float4* wait(int slot, int use) {
        mbar_wait(&full[slot], (uint32_t)use & 1u);
        return buf + (size_t)slot * WCHUNK * TR::R4;
    }
    static constexpr size_t smem_bytes() {
        return (size_t)WARPS * WSTAGES * WCHUNK * TR::REC * sizeof(float) + (size_t)WARPS * WSTAGES * sizeof(uint64_t);
    }
};

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(THREADS) render_forward_kernel(RasterLayout L, const float* __restrict__ view_params,
                                                                  float* __restrict__ out_color,
                                                                  float* __restrict__ out_depth,
                                                                  float* __restrict__ out_alpha) {
    using TR = RecTraits<C>;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int gt = (int)L.tile_order[blockIdx.x / PARTS];
    const int v = gt / L.tiles, t = gt - v * L.tiles;
    const int tile_x = t % L.gx, tile_y = t / L.gx;
    const int tid = threadIdx.x, lane = tid & 31, wl = tid >> 5;     // wl: warp within the CTA (ring slot)
    const int warp = (int)(blockIdx.x % PARTS) * WARPS + wl;          // warp within the tile (pixel block)
    const PixelMap pm(tile_x, tile_y, warp, lane);
    const int px = pm.px, py = pm.py;
    const bool inside = px < L.W && py < L.H;
    const float pfx = (float)px, pfy = (float)py;

    const unsigned int beg = L.tile_offset[gt];
    const int n = L.hdr->overflow ? 0 : (int)(L.tile_offset[gt + 1] - beg);
    const int nchunks = (n + WCHUNK - 1) / WCHUNK;

    bool done = !inside;
    bool warp_done = __all_sync(0xffffffffu, done);
    float T = 1.0f, D = 0.f, Wg = 0.f;
    float Cacc[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) Cacc[ch] = 0.f;
    unsigned int last = 0;

    if (!warp_done && nchunks > 0) {
        WarpRing<C> ring;
        ring.init(smem_raw, wl, lane, L.stream + (size_t)beg * TR::REC, n);
        if (lane == 0)
            for (int c = 0; c < min(WSTAGES, nchunks); ++c) ring.issue(c, c);
        int c = 0;
        for (; c < nchunks; ++c) {
            const int slot = c % WSTAGES;
            const float4* r = ring.wait(slot, c / WSTAGES);
            const int cnt = min(WCHUNK, n - c * WCHUNK);
            CellQueue q = cell_queue<TR::R4>(r, cnt, lane, warp, pm.half);
            if (done) q.clear();
            while (__any_sync(0xffffffffu, !q.empty())) {
                // Take up to FWD_UNROLL candidates of this half-warp's queue at once: their loads, power and exp
                // are independent, only the blend below is sequential.
                int js[FWD_UNROLL];
                float al[FWD_UNROLL];
#pragma unroll
                for (int u = 0; u < FWD_UNROLL; ++u) js[u] = q.pop_front();
#pragma unroll
                for (int u = 0; u < FWD_UNROLL; ++u) {
                    al[u] = 0.f;
                    if (js[u] >= 0 && !done) {
                        const float4* rp = r + js[u] * TR::R4;
                        const float4 a = rp[0];
                        const float4 b = rp[1];
                        const float dx = a.x - pfx, dy = a.y - pfy;
                        const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
                        const float alpha = fminf(0.99f, b.y * gauss_exp(power));
                        al[u] = (power > 0.0f || alpha < 1.0f / 255.0f) ? 0.f : alpha;
                    }
                }
#pragma unroll
                for (int u = 0; u < FWD_UNROLL; ++u) {
                    if (al[u] > 0.f && !done) {
                        const float alpha = al[u];
                        const float test_T = T * (1.0f - alpha);
                        if (test_T < 0.0001f) { done = true; continue; }
                        float f[C], dep;
                        load_features<C>(r + js[u] * TR::R4, f, dep);
                        const float w = alpha * T;
#pragma unroll
                        for (int ch = 0; ch < C; ++ch) Cacc[ch] += f[ch] * w;
                        D += dep * w;
                        Wg += w;
                        T = test_T;
                        last = (unsigned int)(c * WCHUNK + js[u] + 1);
                    }
                }
                if (done) q.clear();
            }
            warp_done = __all_sync(0xffffffffu, done);
            __syncwarp();
            if (warp_done) break;
            if (lane == 0 && c + WSTAGES < nchunks) ring.issue(c + WSTAGES, slot);
        }
        // drain bulk copies still in flight before this warp (and eventually the CTA's shared memory) retires
        if (c < nchunks)
            for (int c2 = c + 1; c2 < min(nchunks, c + WSTAGES); ++c2) ring.wait(c2 % WSTAGES, c2 / WSTAGES);
    }

    if (inside) {
        const float* bg = view_params + (size_t)v * DM4D_VIEW_STRIDE + DM4D_VIEW_BG;
        const size_t npix = (size_t)L.H * L.W;
        const size_t pix = (size_t)py * L.W + px;
        L.n_contrib[(size_t)v * npix + pix] = last;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) out_color[((size_t)v * C + ch) * npix + pix] = Cacc[ch] + T * bg[ch];
        out_depth[(size_t)v * npix + pix] = D;
        out_alpha[(size_t)v * npix + pix] = Wg;
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// Recursive-halving reductions over the 16 lanes of a half-warp (xor distances 8, 4, 2, 1 never leave the half).
// After the call lane `li` holds the half-warp total of ONE slot; `slot_of` gives that slot (or -1 for idle lanes).
//
// N = 16 slots: 8 + 4 + 2 + 1 = 15 shuffles, lane li ends with slot li.
__device__ __forceinline__ float half_reduce_scatter16(float (&v)[16], int li) {
    const bool b3 = li & 8, b2 = li & 4, b1 = li & 2, b0 = li & 1;
    float w8[8], w4[4], w2[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = b3 ? v[i] : v[i + 8], keep = b3 ? v[i + 8] : v[i];
        w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b2 ? w8[i] : w8[i + 4], keep = b2 ? w8[i + 4] : w8[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b1 ? w4[i] : w4[i + 2], keep = b1 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const float send = b0 ? w2[0] : w2[1], keep = b0 ? w2[1] : w2[0];
    return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}
// N = 10 slots (the 3-channel pass: 2 dmean2D, 3 dconic, dopacity, ddepth, 3 dfeatures): 10 -> 5 -> 3 -> 2 -> 1 with
// 5 + 3 + 2 + 1 = 11 shuffles.  Lane li ends with slot 5*b3 + 3*b2 + (2*b1 + b0) when 2*b1 + b0 <= 2 and
// 3*b2 + 2*b1 + b0 <= 4; the other six lanes of the half hold padding.
__device__ __forceinline__ float half_reduce_scatter10(float (&v)[10], int li) {
    const bool b3 = li & 8, b2 = li & 4, b1 = li & 2, b0 = li & 1;
    float w[6], x[4], y[2];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float send = b3 ? v[i] : v[i + 5], keep = b3 ? v[i + 5] : v[i];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    w[5] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float send = b2 ? w[i] : w[i + 3], keep = b2 ? w[i + 3] : w[i];
        x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    x[3] = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b1 ? x[i] : x[i + 2], keep = b1 ? x[i + 2] : x[i];
        y[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const float send = b0 ? y[0] : y[1], keep = b0 ? y[1] : y[0];
    return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}
__device__ __forceinline__ int slot_of10(int li) {
    const int t = li & 3, u = ((li & 4) ? 3 : 0) + t;
    return (t <= 2 && u <= 4) ? ((li & 8) ? 5 : 0) + u : -1;
}

#if DM4D_CELL_ROWS == 2
// EXPERIMENTAL quarter-warp (8 lanes, xor distances 4, 2, 1) reductions: every lane ends with up to TWO slots.
// N = 10: 10 -> 5 -> 3 -> 2 with 5 + 3 + 2 = 10 shuffles; lane li (bits b2 b1 b0) holds out[k], k = 0, 1, for slot
// 5 b2 + 3 b1 + (2 b0 + k) when 2 b0 + k <= 2 and 3 b1 + 2 b0 + k <= 4 (quarter_slot10), else padding.
__device__ __forceinline__ void quarter_reduce_scatter10(float (&v)[10], int li, float (&out)[2]) {
    const bool b2 = li & 4, b1 = li & 2, b0 = li & 1;
    float w[6], x[4];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float send = b2 ? v[i] : v[i + 5], keep = b2 ? v[i + 5] : v[i];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    w[5] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float send = b1 ? w[i] : w[i + 3], keep = b1 ? w[i + 3] : w[i];
        x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    x[3] = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b0 ? x[i] : x[i + 2], keep = b0 ? x[i + 2] : x[i];
        out[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
}
__device__ __forceinline__ int quarter_slot10(int li, int k) {
    const int t = ((li & 1) ? 2 : 0) + k, u = ((li & 2) ? 3 : 0) + t;
    return (t <= 2 && u <= 4) ? ((li & 4) ? 5 : 0) + u : -1;
}
// N = 16: 16 -> 8 -> 4 -> 2 with 8 + 4 + 2 = 14 shuffles; lane li holds slots 2 li and 2 li + 1.
__device__ __forceinline__ void quarter_reduce_scatter16(float (&v)[16], int li, float (&out)[2]) {
    const bool b2 = li & 4, b1 = li & 2, b0 = li & 1;
    float w8[8], w4[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = b2 ? v[i] : v[i + 8], keep = b2 ? v[i + 8] : v[i];
        w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b1 ? w8[i] : w8[i + 4], keep = b1 ? w8[i + 4] : w8[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b0 ? w4[i] : w4[i + 2], keep = b0 ? w4[i + 2] : w4[i];
        out[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
}
#endif

template <int C>
__global__ void __launch_bounds__(THREADS, DM4D_BWD_MIN_BLOCKS) render_backward_kernel(RasterLayout L, const float* __restrict__ view_params,
                                                                   const float* __restrict__ out_alpha,
                                                                   const float* __restrict__ dL_dcolor,
                                                                   const float* __restrict__ dL_ddepth,
                                                                   const float* __restrict__ dL_dalpha_img) {
    using TR = RecTraits<C>;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int gt = (int)L.tile_order[blockIdx.x / PARTS];
    const int v = gt / L.tiles, t = gt - v * L.tiles;
    const int tile_x = t % L.gx, tile_y = t / L.gx;
    const int tid = threadIdx.x, lane = tid & 31, wl = tid >> 5;     // wl: warp within the CTA (ring slot)
    const int warp = (int)(blockIdx.x % PARTS) * WARPS + wl;          // warp within the tile (pixel block)
    const PixelMap pm(tile_x, tile_y, warp, lane);
    const int px = pm.px, py = pm.py;
    const bool inside = px < L.W && py < L.H;
    const float pfx = (float)px, pfy = (float)py;
    const size_t npix = (size_t)L.H * L.W;
    const size_t pix = (size_t)py * L.W + px;

    const unsigned int beg = L.tile_offset[gt];
    const int n = L.hdr->overflow ? 0 : (int)(L.tile_offset[gt + 1] - beg);
    if (n == 0) return;

    const unsigned int last_contributor = inside ? L.n_contrib[(size_t)v * npix + pix] : 0u;
    unsigned int warp_last = last_contributor;      // last contributor over the warp's 32 pixels
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));
    if (warp_last == 0) return;                      // nothing composited in this block
    const int nlive = min(n, (int)warp_last);        // this warp only needs instances in front of its last contributor
    const int nchunks = (nlive + WCHUNK - 1) / WCHUNK;

    WarpRing<C> ring;
    ring.init(smem_raw, wl, lane, L.stream + (size_t)beg * TR::REC, nlive);
    // k-th chunk in processing order = chunk index nchunks-1-k (back to front)
    if (lane == 0)
        for (int k = 0; k < min(WSTAGES, nchunks); ++k) ring.issue(nchunks - 1 - k, k);

    const float* vp = view_params + (size_t)v * DM4D_VIEW_STRIDE;
    float gC[C];
    float bg_dot = 0.f;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
        gC[ch] = inside ? dL_dcolor[((size_t)v * C + ch) * npix + pix] : 0.f;
        bg_dot += vp[DM4D_VIEW_BG + ch] * gC[ch];
    }
    const float gD = (inside && dL_ddepth) ? dL_ddepth[(size_t)v * npix + pix] : 0.f;
    const float gA = (inside && dL_dalpha_img) ? dL_dalpha_img[(size_t)v * npix + pix] : 0.f;
    const float T_final = inside ? 1.0f - out_alpha[(size_t)v * npix + pix] : 0.f;
    float T = T_final;
    float accum_rec[C], last_color[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) { accum_rec[ch] = 0.f; last_color[ch] = 0.f; }
    float accum_d = 0.f, last_depth = 0.f, accum_a = 0.f, last_alpha = 0.f;
    const float ddelx_dx = 0.5f * (float)L.W, ddely_dy = 0.5f * (float)L.H;
    float* accum_view = L.accum + (size_t)v * L.P * TR::ACC;
    // accumulator-row offset this lane adds its reduced slot to (rows: raster_internal.cuh); -1 = idle lane
#if DM4D_CELL_ROWS == 4
    int acc_off;
    if constexpr (C <= 3) {
        const int sl = slot_of10(pm.li);          // reduction slots 0..6 = row 0..6, slots 7..9 = features at row 8..10
        acc_off = sl < 0 ? -1 : (sl < 7 ? sl : sl + 1);
    } else {
        acc_off = (pm.li != 7 && pm.li < 8 + C) ? pm.li : -1;
    }
    const unsigned int half_lanes = 0xffffu << (pm.half * 16);
#else
    int acc_off2[2];                              // EXPERIMENTAL: two reduced slots per lane of a quarter-warp
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if constexpr (C <= 3) {
            const int sl = quarter_slot10(pm.li, k);
            acc_off2[k] = sl < 0 ? -1 : (sl < 7 ? sl : sl + 1);
        } else {
            const int sl = 2 * pm.li + k;         // lane li of the 16 -> 8 -> 4 -> 2 halving holds slots 2 li, 2 li + 1
            acc_off2[k] = (sl != 7 && sl < 8 + C) ? sl : -1;
        }
    }
    const unsigned int half_lanes = 0xffu << (pm.half * 8);
#endif

    for (int k = 0; k < nchunks; ++k) {
        const int c = nchunks - 1 - k;
        const int sl = k % WSTAGES;
        const float4* r = ring.wait(sl, k / WSTAGES);
        const int cnt = min(WCHUNK, nlive - c * WCHUNK);
        CellQueue q = cell_queue<TR::R4>(r, cnt, lane, warp, pm.half);
        while (__any_sync(0xffffffffu, !q.empty())) {
            // next instance of this half-warp's queue, back to front (the two halves walk different instances)
            const bool have = !q.empty();
            const int j = have ? q.pop_back() : 0;
            const unsigned int gi = (unsigned int)(c * WCHUNK + j);
            const float4* rp = r + j * TR::R4;
            bool valid = have && gi < last_contributor;
            float4 a, b;
            float dx = 0.f, dy = 0.f, G = 0.f, alpha = 0.f;
            if (valid) {
                a = rp[0];
                b = rp[1];
                dx = a.x - pfx; dy = a.y - pfy;
                const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
                valid = !(power > 0.0f);
                if (valid) {
                    G = gauss_exp(power);
                    alpha = fminf(0.99f, b.y * G);
                    valid = !(alpha < 1.0f / 255.0f);
                }
            }
            const unsigned int vb = __ballot_sync(0xffffffffu, valid);
            if (vb == 0u) continue;

            // reduction slots: 0-1 dmean2D, 2-4 dconic, 5 dopacity, 6 ddepth, then the C feature gradients
            // (3 channels: slots 7..9 of a 10-slot reduction; 6 channels: slots 8..13 of a 16-slot one, 7 unused)
            constexpr int NS = C <= 3 ? 10 : 16;
            constexpr int F0 = C <= 3 ? 7 : 8;
            float gv[NS];
#pragma unroll
            for (int i = 0; i < NS; ++i) gv[i] = 0.f;
            if (valid) {
                const float inv_1ma = fast_rcp(1.0f - alpha);       // MUFU.RCP (1 ulp), shared by both uses below
                T = T * inv_1ma;
                const float w = alpha * T;
                float f[C], dep;
                load_features<C>(rp, f, dep);
                float dL_dalpha = 0.f;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                    last_color[ch] = f[ch];
                    dL_dalpha += (f[ch] - accum_rec[ch]) * gC[ch];
                    gv[F0 + ch] = w * gC[ch];
                }
                accum_d = last_alpha * last_depth + (1.f - last_alpha) * accum_d;
                last_depth = dep;
                dL_dalpha += (dep - accum_d) * gD;
                gv[6] = w * gD;
                accum_a = last_alpha + (1.f - last_alpha) * accum_a;
                dL_dalpha += (1.f - accum_a) * gA;
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final * inv_1ma) * bg_dot;
                const float dL_dG = b.y * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                const float dG_ddelx = -gdx * a.z - gdy * a.w;
                const float dG_ddely = -gdy * b.x - gdx * a.w;
                gv[0] = dL_dG * dG_ddelx * ddelx_dx;
                gv[1] = dL_dG * dG_ddely * ddely_dy;
                const float hx = -0.5f * dL_dG * gdx, hy = -0.5f * dL_dG * gdy;   // shared factors of the conic terms
                gv[2] = hx * dx;
                gv[3] = hx * dy;
                gv[4] = hy * dy;
                gv[5] = G * dL_dalpha;
            }
#if DM4D_CELL_ROWS == 4
            float tot;
            if constexpr (C <= 3) tot = half_reduce_scatter10(gv, pm.li);
            else tot = half_reduce_scatter16(gv, pm.li);
            // one coalesced reduction per (half-warp, instance) into the instance's accumulator row
            if (acc_off >= 0 && (vb & half_lanes)) {
                const int id = __float_as_int(rp[1].w);
                atomicAdd(accum_view + (size_t)id * TR::ACC + acc_off, tot);
            }
#else
            float tot2[2];
            if constexpr (C <= 3) quarter_reduce_scatter10(gv, pm.li, tot2);
            else quarter_reduce_scatter16(gv, pm.li, tot2);
            if (vb & half_lanes) {                // this quarter-warp's instance had a contributing lane
                float* row = accum_view + (size_t)__float_as_int(rp[1].w) * TR::ACC;
                if (acc_off2[0] >= 0) atomicAdd(row + acc_off2[0], tot2[0]);
                if (acc_off2[1] >= 0) atomicAdd(row + acc_off2[1], tot2[1]);
            }
#endif
        }
        __syncwarp();
        if (lane == 0 && k + WSTAGES < nchunks) ring.issue(nchunks - 1 - (k + WSTAGES), sl);
    }
}

template <int C>
int launch_fwd_t(const dm4d_raster_desc* d, const RasterLayout& L, float* out_color, float* out_depth,
                 float* out_alpha, cudaStream_t s) {
    const size_t smem = WarpRing<C>::smem_bytes();
    static bool configured = false;
    if (!configured) {
        DM4D_CUDA_CHECK(cudaFuncSetAttribute(render_forward_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    {
        KernelTimer kt(DM4D_K_RENDER_FWD, s);
        render_forward_kernel<C><<<(unsigned)(L.n_views * L.tiles * PARTS), THREADS, smem, s>>>(L, d->view_params, out_color,
                                                                                        out_depth, out_alpha);
    }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

template <int C>
int launch_bwd_t(const dm4d_raster_desc* d, const RasterLayout& L, const float* out_alpha, const float* dL_dcolor,
                 const float* dL_ddepth, const float* dL_dalpha, cudaStream_t s) {
    const size_t smem = WarpRing<C>::smem_bytes();
    static bool configured = false;
    if (!configured) {
        DM4D_CUDA_CHECK(cudaFuncSetAttribute(render_backward_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    DM4D_CUDA_CHECK(cudaMemsetAsync(L.accum, 0, (size_t)L.n_views * L.P * L.acc * sizeof(float), s));
    {
        KernelTimer kt(DM4D_K_RENDER_BWD, s);
        render_backward_kernel<C><<<(unsigned)(L.n_views * L.tiles * PARTS), THREADS, smem, s>>>(L, d->view_params, out_alpha,
                                                                                         dL_dcolor, dL_ddepth, dL_dalpha);
    }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

}  // namespace

int launch_render_forward(const dm4d_raster_desc* d, const RasterLayout& L, float* out_color, float* out_depth,
                          float* out_alpha, cudaStream_t s) {
    if (L.n_views * L.tiles == 0) return DM4D_OK;
    return L.channels <= 3 ? launch_fwd_t<3>(d, L, out_color, out_depth, out_alpha, s)
                           : launch_fwd_t<6>(d, L, out_color, out_depth, out_alpha, s);
}

int launch_render_backward(const dm4d_raster_desc* d, const RasterLayout& L, const float* out_alpha,
                           const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha, cudaStream_t s) {
    if (L.n_views * L.tiles == 0) return DM4D_OK;
    return L.channels <= 3 ? launch_bwd_t<3>(d, L, out_alpha, dL_dcolor, dL_ddepth, dL_dalpha, s)
                           : launch_bwd_t<6>(d, L, out_alpha, dL_dcolor, dL_ddepth, dL_dalpha, s);
}
