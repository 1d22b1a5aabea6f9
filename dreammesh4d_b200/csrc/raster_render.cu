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
// the walked (warp, instance) iterations from 1.91 (16x2 strips) to about 1.2 per instance.
//
// Backward.  Three ideas, each measured (profiles/):
//  * SEGMENTS.  The forward checkpoints every pixel's compositing state (C, D, T) at each DM4D_SEG-th instance of its
//    tile's list; the backward processes (tile, segment, 8x4 block) work items independently, so the serial chain of
//    a heavy tile (6000 instances at C3: 0.4 ms for ONE warp = the whole duration of a single-view launch) is bounded
//    by one segment.  A segment starts at its END with the state "behind" it: for pixels whose contributors stop
//    inside the segment that is (T_final, background) exactly as in the replaced rasterizer; for the others it
//    follows from the next segment's checkpoint and the forward's output images.
//  * ONE SCALAR RECURRENCE.  With s_k = c_k . dL/dC + depth_k dL/dD + dL/dA the back-to-front recurrences of the
//    replaced rasterizer (accum_rec[C], accum_d, accum_a, last_*) collapse into
//    Q_{j-1} = alpha_j s_j + (1 - alpha_j) Q_j,   dL/dalpha_j = T_j (s_j - Q_j),   Q_end = bg . dL/dC.
//  * NO CROSS-LANE REDUCTION.  ncu on the round-1 kernel: 175 warp instructions per (warp, instance) step, 40 % of
//    them the shuffle / select / add tree that sums ten partial gradients over the 16 pixels of a cell, with a third
//    of the lanes contributing.  Now the pixel-parallel sweep only evaluates alpha, advances (T, Q) and drops two
//    numbers per pixel — h = G dL/dalpha and w = alpha T — into shared memory; after up to 16 steps the warp switches
//    roles: every LANE takes one (cell, instance) pair and sums its 16 pixels in registers (all gradients of an
//    instance are moments of h and w over the pixel offsets), then issues three 16-byte vector reductions
//    (REDG.E.ADD.F32x4) on the instance's accumulator row.
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
#define DM4D_BWD_MIN_WARPS 24     // resident warps per SM the backward's register budget is sized for (measured at C3: no
#endif                            // cap (90 registers) 0.841 ms, 20: 0.825, 24 (85 registers): 0.805, 28: 0.863, 32: 0.904)
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

// Candidate queue of this lane's half-warp for one staged chunk of 64 instances: bit j set <=> instance j of the chunk
// can reach the half-warp's cell.  Four warp ballots (two cells x two 32-instance groups); uniform within a half-warp.
// Kept as two 32-bit words: a pop is FLO + shift + mask on one word instead of 64-bit arithmetic.
struct CellQueue {
    unsigned int lo, hi;
    __device__ __forceinline__ bool empty() const { return (lo | hi) == 0u; }
    __device__ __forceinline__ void clear() { lo = hi = 0u; }
    // back to front (backward): highest set bit; the queue must not be empty
    __device__ __forceinline__ int pop_back() {
        const bool in_hi = hi != 0u;
        unsigned int cur = in_hi ? hi : lo;
        const int b = 31 - __clz((int)cur);
        cur &= ~(1u << b);
        if (in_hi) hi = cur; else lo = cur;
        return b + (in_hi ? 32 : 0);
    }
    // front to back: lowest set bit, or -1
    __device__ __forceinline__ int pop_front() {
        const bool in_lo = lo != 0u;
        unsigned int cur = in_lo ? lo : hi;
        const int j = cur ? (__ffs((int)cur) - 1 + (in_lo ? 0 : 32)) : -1;
        cur &= cur - 1u;
        if (in_lo) lo = cur; else hi = cur;
        return j;
    }
};

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
template <int C, int CHUNK = WCHUNK>
struct WarpRing {
    using TR = RecTraits<C>;
    float4* buf;        // [WSTAGES][CHUNK * R4]
    uint64_t* full;     // [WSTAGES]
    const float* stream;
    int n;              // instances available in the tile
    __device__ __forceinline__ void init(unsigned char* smem_raw, int warp, int lane, const float* stream_, int n_) {
        buf = reinterpret_cast<float4*>(smem_raw) + (size_t)warp * WSTAGES * CHUNK * TR::R4;
        full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)WARPS * WSTAGES * CHUNK * TR::REC * sizeof(float)) + warp * WSTAGES;
        stream = stream_;
        n = n_;
        if (lane == 0) {
            for (int s = 0; s < WSTAGES; ++s) mbar_init(&full[s], 1);
            fence_mbar_init();
        }
        __syncwarp();
    }
    // lane 0 only: stage chunk `c` (instances [c*CHUNK, ...)) into slot `slot`
    __device__ __forceinline__ void issue(int c, int slot) {
        const int cnt = min(CHUNK, n - c * CHUNK);
        const uint32_t bytes = (uint32_t)cnt * TR::REC * sizeof(float);
        mbar_expect_tx(&full[slot], bytes);
        bulk_g2s(buf + (size_t)slot * CHUNK * TR::R4, stream + (size_t)c * CHUNK * TR::REC, bytes, &full[slot]);
    }
    __device__ __forceinline__ const float4* wait(int slot, int use) {
        mbar_wait(&full[slot], (uint32_t)use & 1u);
        return buf + (size_t)slot * CHUNK * TR::R4;
    }
    static constexpr size_t smem_bytes() {
        return (size_t)WARPS * WSTAGES * CHUNK * TR::REC * sizeof(float) + (size_t)WARPS * WSTAGES * sizeof(uint64_t);
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
            if (c > 0 && (c * WCHUNK) % DM4D_SEG == 0) {
                // first instance of a backward segment: checkpoint this warp's compositing state (coalesced rows of 32)
                float* ck = L.ckpt + ((size_t)(L.seg_offset[gt] + (unsigned)(c * WCHUNK / DM4D_SEG)) * (C + 2)) * 256 + warp * 32 + lane;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) ck[ch * 256] = Cacc[ch];
                ck[C * 256] = D;
                ck[(C + 1) * 256] = T;
            }
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
#ifndef DM4D_BWD_CHUNK
#define DM4D_BWD_CHUNK 64      // instances per staged chunk of the backward (32: smaller ring, more resident warps, but +10 % time)
#endif
constexpr int BCHUNK = DM4D_BWD_CHUNK;
static_assert(BCHUNK == 32 || BCHUNK == 64, "DM4D_BWD_CHUNK must be 32 or 64");
static_assert(DM4D_SEG % 64 == 0, "segments are whole chunks of the forward and of the backward");

// candidate queue of a 32-instance chunk: one ballot per cell
template <int R4>
__device__ __forceinline__ CellQueue cell_queue32(const float4* r, int cnt, int lane, int warp, int half) {
    const unsigned int m0 = lane < cnt ? __float_as_uint(r[lane * R4 + 1].z) : 0u;
    const int sh = 2 * warp;
    const unsigned int a0 = __ballot_sync(0xffffffffu, (m0 >> sh) & 1u), b0 = __ballot_sync(0xffffffffu, (m0 >> (sh + 1)) & 1u);
    CellQueue q;
    q.lo = half ? b0 : a0;
    q.hi = 0u;
    return q;
}

// Shared memory of one backward warp behind its record ring: the (h, w) rows of up to 2 x 16 (cell, instance) pairs,
// the per-pixel loss gradients of the warp's two cells, and the instance slot of every pair.
template <int C>
struct BwdScratch {
#ifndef DM4D_BWD_SLOTS
#define DM4D_BWD_SLOTS 16
#endif
    static constexpr int SLOTS = DM4D_BWD_SLOTS;   // pairs per half-warp per batch
    static constexpr int ROW = 34;                 // floats per pair: 16 pixels x (h, w) + 2 pad -> conflict-free 8-byte accesses
    static constexpr int PC = C <= 3 ? 4 : 8;      // per-pixel constants: dL/dC[0..C), dL/dD (+ pad)
    float hw[2 * SLOTS * ROW];
    float pc[32 * PC];
    int slot_j[2 * SLOTS];
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int C>
// the 6-channel instantiation needs 96 registers: its budget stays at 20 warps per SM (no spills)
__global__ void __launch_bounds__(THREADS, C <= 3 ? DM4D_BWD_MIN_BLOCKS : (20 / DM4D_RENDER_WARPS)) render_backward_kernel(RasterLayout L, const float* __restrict__ view_params,
                                                                   const float* __restrict__ out_color,
                                                                   const float* __restrict__ out_depth,
                                                                   const float* __restrict__ out_alpha,
                                                                   const float* __restrict__ dL_dcolor,
                                                                   const float* __restrict__ dL_ddepth,
                                                                   const float* __restrict__ dL_dalpha_img) {
    using TR = RecTraits<C>;
    using SC = BwdScratch<C>;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    // work item = (segment of a tile's instance list, 8x4 pixel block): one warp
    const unsigned int seg = blockIdx.x / PARTS;
    if (seg >= L.hdr->total_segs) return;           // also covers overflow (total_segs = 0)
    const int gt = (int)L.seg_tile[seg];
    const int ks = (int)(seg - L.seg_offset[gt]);   // segment index within the tile
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
    const int n = (int)(L.tile_offset[gt + 1] - beg);
    const int m0 = ks * DM4D_SEG;                   // first instance of this segment
    const int seg_end = m0 + DM4D_SEG;              // first instance of the next one

    const unsigned int last_contributor = inside ? L.n_contrib[(size_t)v * npix + pix] : 0u;
    unsigned int warp_last = last_contributor;      // last contributor over the warp's 32 pixels
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));
    if ((unsigned int)m0 >= warp_last) return;       // nothing composited by this block from here on
    const int m1 = min(min(n, seg_end), (int)warp_last);
    const int c0 = m0 / BCHUNK, nchunks = (m1 - m0 + BCHUNK - 1) / BCHUNK;

    WarpRing<C, BCHUNK> ring;
    ring.init(smem_raw, wl, lane, L.stream + (size_t)beg * TR::REC, m1);
    // k-th chunk in processing order = chunk c0 + nchunks-1-k (back to front)
    if (lane == 0)
        for (int k = 0; k < min(WSTAGES, nchunks); ++k) ring.issue(c0 + nchunks - 1 - k, k);
    SC& sc = *reinterpret_cast<SC*>(smem_raw + WarpRing<C, BCHUNK>::smem_bytes() + (size_t)wl * sizeof(SC));

    const float* vp = view_params + (size_t)v * DM4D_VIEW_STRIDE;
    float gC[C];
    float bg_dot = 0.f, F = 0.f;
    const float T_final = inside ? 1.0f - out_alpha[(size_t)v * npix + pix] : 0.f;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
        gC[ch] = inside ? dL_dcolor[((size_t)v * C + ch) * npix + pix] : 0.f;
        const float bgc = vp[DM4D_VIEW_BG + ch];
        bg_dot += bgc * gC[ch];
        // composited feature without the background term: out_color = sum_k w_k c_k + T_final bg
        F += gC[ch] * (inside ? out_color[((size_t)v * C + ch) * npix + pix] - T_final * bgc : 0.f);
        sc.pc[lane * SC::PC + ch] = gC[ch];
    }
    const float gD = (inside && dL_ddepth) ? dL_ddepth[(size_t)v * npix + pix] : 0.f;
    const float gA = (inside && dL_dalpha_img) ? dL_dalpha_img[(size_t)v * npix + pix] : 0.f;
    sc.pc[lane * SC::PC + C] = gD;
    if constexpr (C > 3) sc.pc[lane * SC::PC + C + 1] = 0.f;
    if (inside) F += gD * out_depth[(size_t)v * npix + pix] + gA * (1.0f - T_final);

    // State "behind" the segment.  Pixels whose contributors end inside it: (T_final, background), as in the replaced
    // rasterizer.  Pixels that continue past seg_end: T from the next segment's checkpoint, Q = (what is composited
    // behind the boundary, background included) / T — the totals come from the forward's output images.
    float T = T_final, Q = bg_dot;
    if (last_contributor > (unsigned int)seg_end) {
        const float* ck = L.ckpt + ((size_t)(seg + 1) * (C + 2)) * 256 + warp * 32 + lane;
        const float Tb = ck[(C + 1) * 256];
        float Pb = gD * ck[C * 256] + gA * (1.0f - Tb);
#pragma unroll
        for (int ch = 0; ch < C; ++ch) Pb += gC[ch] * ck[ch * 256];
        T = Tb;
        Q = (F + T_final * bg_dot - Pb) / Tb;
    }
    const float kx = 0.5f * (float)L.W, ky = 0.5f * (float)L.H;
    float* accum_view = L.accum + (size_t)v * L.P * TR::ACC;
    const unsigned int half_lanes = 0xffffu << (pm.half * 16);
    __syncwarp();

    for (int k = 0; k < nchunks; ++k) {
        const int c = c0 + nchunks - 1 - k;
        const int sl = k % WSTAGES;
        const float4* r = ring.wait(sl, k / WSTAGES);
        const int cnt = min(BCHUNK, m1 - c * BCHUNK);
        CellQueue q = BCHUNK == 64 ? cell_queue<TR::R4>(r, cnt, lane, warp, pm.half) : cell_queue32<TR::R4>(r, cnt, lane, warp, pm.half);
        while (__any_sync(0xffffffffu, !q.empty())) {
            // ---- sweep 1: pixel-parallel, back to front, up to SLOTS instances per half-warp; straight-line code ------
            int n_mine = 0;                           // pairs of this lane's half in the batch (uniform within the half)
            // everything of one candidate that does not depend on the running (T, Q): two candidates are evaluated per
            // step so their shared-memory loads, exponent and exp overlap; only the three-instruction (T, Q) update is serial
            struct Cand { float G, al, sj; bool valid; int slot_val; };
            auto eval = [&](bool have, int j) -> Cand {
                const float4* rp = r + j * TR::R4;
                const float4 a = rp[0];
                const float4 b = rp[1];
                float f[C], dep;
                load_features<C>(rp, f, dep);
                const float dx = a.x - pfx, dy = a.y - pfy;
                const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
                Cand o;
                o.G = gauss_exp(fminf(power, 0.0f));
                const float alpha = fminf(0.99f, b.y * o.G);
                o.valid = have && (unsigned int)(c * BCHUNK + j) < last_contributor && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                o.al = o.valid ? alpha : 0.0f;                           // alpha = 0 leaves (T, Q) untouched
                o.sj = gA + dep * gD;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) o.sj += f[ch] * gC[ch];
                const unsigned int vb = __ballot_sync(0xffffffffu, o.valid);
                o.slot_val = (have && (vb & half_lanes)) ? j : -1;
                return o;
            };
            auto advance = [&](const Cand& cd, int it) {
                T *= fast_rcp(1.0f - cd.al);                             // T_j (MUFU.RCP, 1 ulp)
                const float d = cd.sj - Q;
                const float hval = cd.valid ? cd.G * (T * d) : 0.0f;     // G dL/dalpha
                Q += cd.al * d;                                          // Q_{j-1} = alpha s + (1 - alpha) Q
                *reinterpret_cast<float2*>(&sc.hw[(pm.half * SC::SLOTS + it) * SC::ROW + 2 * pm.li]) = make_float2(hval, cd.al * T);
                if (pm.li == 0) sc.slot_j[pm.half * SC::SLOTS + it] = cd.slot_val;
            };
            for (int it = 0; it < SC::SLOTS && __any_sync(0xffffffffu, !q.empty()); it += 2) {
                const bool have0 = !q.empty();
                const int j0 = have0 ? q.pop_back() : 0;                 // record 0 of the chunk is always readable
                const bool have1 = !q.empty();
                const int j1 = have1 ? q.pop_back() : 0;
                const Cand ca = eval(have0, j0);
                const Cand cb = eval(have1, j1);
                advance(ca, it);
                advance(cb, it + 1);
                n_mine += (have0 ? 1 : 0) + (have1 ? 1 : 0);
            }
            const int n0 = __shfl_sync(0xffffffffu, n_mine, 0), n1 = __shfl_sync(0xffffffffu, n_mine, 16);
            __syncwarp();
            // ---- sweep 2: pair-parallel — lane p owns one (cell, instance) pair and sums its 16 pixels ----------------
            if (lane < n0 + n1) {
                const int h2 = lane < n0 ? 0 : 1, slot = lane < n0 ? lane : lane - n0;
                const int j = sc.slot_j[h2 * SC::SLOTS + slot];
                if (j >= 0) {
                    const float4* rp = r + j * TR::R4;
                    const float4 a = rp[0];
                    const float4 b = rp[1];
                    // pixel (i, jrow) of the cell sits at (ox + i, oy + jrow); offsets rounded once, exactly as in sweep 1
                    const int ox = tile_x * DM4D_TILE + ((((warp & 1) << 1) | h2) << 2), oy = tile_y * DM4D_TILE + ((warp >> 1) << 2);
                    float dxs[4], dys[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { dxs[i] = a.x - (float)(ox + i); dys[i] = a.y - (float)(oy + i); }
                    const float* row = &sc.hw[(h2 * SC::SLOTS + slot) * SC::ROW];
                    const float* pcs = &sc.pc[h2 * 16 * SC::PC];
                    float S0 = 0.f, Sx = 0.f, Sy = 0.f, Sxx = 0.f, Sxy = 0.f, Syy = 0.f, Sd = 0.f;
                    float Sf[C];
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) Sf[ch] = 0.f;
#pragma unroll
                    for (int p = 0; p < 16; ++p) {
                        const float2 hwv = *reinterpret_cast<const float2*>(row + 2 * p);
                        const float dx = dxs[p & 3], dy = dys[p >> 2];
                        const float tx = hwv.x * dx, ty = hwv.x * dy;
                        S0 += hwv.x; Sx += tx; Sy += ty;
                        Sxx += tx * dx; Sxy += tx * dy; Syy += ty * dy;
                        if constexpr (C <= 3) {
                            const float4 g4 = *reinterpret_cast<const float4*>(pcs + p * SC::PC);
                            Sf[0] += hwv.y * g4.x; Sf[1] += hwv.y * g4.y; Sf[2] += hwv.y * g4.z; Sd += hwv.y * g4.w;
                        } else {
                            const float4 g4 = *reinterpret_cast<const float4*>(pcs + p * SC::PC);
                            const float4 g5 = *reinterpret_cast<const float4*>(pcs + p * SC::PC + 4);
                            Sf[0] += hwv.y * g4.x; Sf[1] += hwv.y * g4.y; Sf[2] += hwv.y * g4.z; Sf[3] += hwv.y * g4.w;
                            Sf[4] += hwv.y * g5.x; Sf[5] += hwv.y * g5.y; Sd += hwv.y * g5.z;
                        }
                    }
                    // accumulator row: [0..1] dmean2D, [2..4] dconic (x, y(half), z), [5] dopacity, [6] ddepth, [7] pad, [8..] dfeatures
                    const float o = b.y;
                    float* acc = accum_view + (size_t)__float_as_int(b.w) * TR::ACC;
                    red_add_v4(acc, -o * kx * (a.z * Sx + a.w * Sy), -o * ky * (b.x * Sy + a.w * Sx), -0.5f * o * Sxx, -0.5f * o * Sxy);
                    red_add_v4(acc + 4, -0.5f * o * Syy, S0, Sd, 0.f);
                    if constexpr (C <= 3) {
                        red_add_v4(acc + 8, Sf[0], Sf[1], Sf[2], 0.f);
                    } else {
                        red_add_v4(acc + 8, Sf[0], Sf[1], Sf[2], Sf[3]);
                        red_add_v4(acc + 12, Sf[4], Sf[5], 0.f, 0.f);
                    }
                }
            }
            __syncwarp();
        }
        if (lane == 0 && k + WSTAGES < nchunks) ring.issue(c0 + nchunks - 1 - (k + WSTAGES), sl);
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
int launch_bwd_t(const dm4d_raster_desc* d, const RasterLayout& L, const float* out_color, const float* out_depth,
                 const float* out_alpha, const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha, cudaStream_t s) {
    const size_t smem = WarpRing<C, BCHUNK>::smem_bytes() + (size_t)WARPS * sizeof(BwdScratch<C>);
    static bool configured = false;
    if (!configured) {
        DM4D_CUDA_CHECK(cudaFuncSetAttribute(render_backward_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    DM4D_CUDA_CHECK(cudaMemsetAsync(L.accum, 0, (size_t)L.n_views * L.P * L.acc * sizeof(float), s));
    {
        KernelTimer kt(DM4D_K_RENDER_BWD, s);
        // one CTA group per POSSIBLE segment (host-side bound); CTAs past the device-side segment count exit at once
        render_backward_kernel<C><<<(unsigned)(L.seg_cap * PARTS), THREADS, smem, s>>>(L, d->view_params, out_color, out_depth, out_alpha,
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

int launch_render_backward(const dm4d_raster_desc* d, const RasterLayout& L, const float* out_color, const float* out_depth,
                           const float* out_alpha, const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                           cudaStream_t s) {
    if (L.n_views * L.tiles == 0) return DM4D_OK;
    return L.channels <= 3 ? launch_bwd_t<3>(d, L, out_color, out_depth, out_alpha, dL_dcolor, dL_ddepth, dL_dalpha, s)
                           : launch_bwd_t<6>(d, L, out_color, out_depth, out_alpha, dL_dcolor, dL_ddepth, dL_dalpha, s);
}
