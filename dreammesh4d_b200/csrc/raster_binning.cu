// Tile binning without a global sort and without a host read-back.
//
// The replaced rasterizer (un-vendored diff-gaussian-rasterization; bound by the reference at
// custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_temporal.py:169-178) does
// InclusiveSum(tiles_touched) -> D2H num_rendered -> duplicateWithKeys -> one global 64-bit
// radix sort -> identifyTileRanges (SURVEY.md Appendix A.2 "binning").  Here:
//   1. preprocess already counted instances per (view, tile) with atomics,
//   2. scan_tiles_kernel: exclusive prefix sum over the (view, tile) counts -> tile ranges directly,
//   3. scatter_kernel: every (view, Gaussian) drops (depth bits << 32 | id) into its tiles' segments,
//   4. sort_pack_kernel: one CTA per (view, tile) sorts its segment on the full 64-bit key in shared
//      memory (bitonic network; chunked with global merge steps for segments > 4096) and writes the
//      depth-sorted instance stream of packed records (with their 4x4-cell masks) that the render kernels pull
//      with TMA bulk copies.
// Order = ascending (tile, depth bits, Gaussian id) — exactly the order of upstream's stable radix
// sort fed in ascending-id emission order, independent of atomics ordering.
#include <algorithm>
#include "raster_internal.cuh"

namespace {

constexpr int SCAN_THREADS = 1024;

__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(RasterLayout L) {
    __shared__ unsigned int warp_sums[SCAN_THREADS / 32];
    __shared__ unsigned int carry_s;
    __shared__ unsigned int bucket_cnt[33], bucket_pos[33];
    const int n = L.n_views * L.tiles;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (n + SCAN_THREADS - 1) / SCAN_THREADS;
    const int beg = min(n, tid * per), end = min(n, beg + per);
    unsigned int local = 0;
    for (int i = beg; i < end; ++i) local += L.tile_count[i];
    // block-wide exclusive scan of `local`
    unsigned int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        unsigned int w = warp_sums[lane];
        unsigned int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        warp_sums[lane] = wi - w;   // exclusive
        if (lane == 31) carry_s = wi;
    }
    __syncthreads();
    unsigned int run = warp_sums[wid] + (incl - local);
    for (int i = beg; i < end; ++i) {
        L.tile_offset[i] = run;
        run += L.tile_count[i];
        L.tile_cursor[i] = 0u;
    }
    if (tid == 0) {
        const unsigned int total = carry_s;
        L.tile_offset[n] = total;
        L.hdr->total = total;
        L.hdr->overflow = ((long long)total > L.capacity) ? 1u : 0u;
    }
    // Segment tables of the backward: exclusive scan of ceil(count / DM4D_SEG) and the segment -> tile map.
    __syncthreads();
    {
        unsigned int lseg = 0;
        for (int i = beg; i < end; ++i) lseg += (L.tile_count[i] + DM4D_SEG - 1) / DM4D_SEG;
        unsigned int sincl = lseg;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int t = __shfl_up_sync(0xffffffffu, sincl, o);
            if (lane >= o) sincl += t;
        }
        if (lane == 31) warp_sums[wid] = sincl;
        __syncthreads();
        if (wid == 0) {
            unsigned int w = warp_sums[lane];
            unsigned int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sums[lane] = wi - w;
            if (lane == 31) carry_s = wi;
        }
        __syncthreads();
        unsigned int srun = warp_sums[wid] + (sincl - lseg);
        const bool fits = (long long)carry_s <= L.seg_cap && !((long long)L.tile_offset[n] > L.capacity);
        for (int i = beg; i < end; ++i) {
            L.seg_offset[i] = srun;
            const unsigned int ns = (L.tile_count[i] + DM4D_SEG - 1) / DM4D_SEG;
            if (fits)
                for (unsigned int k = 0; k < ns; ++k) L.seg_tile[srun + k] = (unsigned int)i;
            srun += ns;
        }
        if (tid == 0) {
            L.seg_offset[n] = carry_s;
            L.hdr->total_segs = fits ? carry_s : 0u;
        }
        __syncthreads();
    }
    // Launch order of the per-tile CTAs: heaviest tiles first (counting sort on floor(log2(count))), so the
    // long limb/centre tiles do not form the tail of the render kernels.
    if (tid < 33) bucket_cnt[tid] = 0u;
    __syncthreads();
    for (int i = beg; i < end; ++i) atomicAdd(&bucket_cnt[32 - __clz(L.tile_count[i])], 1u);
    __syncthreads();
    if (tid == 0) {
        unsigned int run2 = 0;
        for (int b = 32; b >= 0; --b) { bucket_pos[b] = run2; run2 += bucket_cnt[b]; }
    }
    __syncthreads();
    for (int i = beg; i < end; ++i) L.tile_order[atomicAdd(&bucket_pos[32 - __clz(L.tile_count[i])], 1u)] = (unsigned int)i;
}

constexpr int SMEM_TILES = 4096;

// One thread per (view, Gaussian): drops (depth bits << 32 | id) into the segments of the tiles it touches.
// Slots are reserved per block: shared-memory count per tile -> one global atomic per touched tile for the
// block's base -> shared-memory atomics hand out the slots.  (Order inside a segment is irrelevant: the
// per-tile sort orders on the full 64-bit key.)
__global__ void __launch_bounds__(DM4D_BLOCK) scatter_kernel(RasterLayout L) {
    __shared__ unsigned int hist[SMEM_TILES];
    if (L.hdr->overflow) return;
    const long long total = (long long)L.n_views * L.P;
    const long long first = (long long)blockIdx.x * blockDim.x;
    const long long last = min(first + blockDim.x, total) - 1;
    const long long idx = first + threadIdx.x;
    const int v_first = (int)(first / L.P);
    const bool use_smem = L.tiles <= SMEM_TILES && v_first == (int)(last / L.P);
    const unsigned int rect = idx < total ? L.g_rect[idx] : 0u;
    const int minx = rect & 0xff, miny = (rect >> 8) & 0xff, maxx = (rect >> 16) & 0xff, maxy = rect >> 24;
    unsigned long long key = 0ull;
    size_t tbase = 0;
    if (rect) {
        const int v = (int)(idx / L.P);
        const unsigned int g = (unsigned int)(idx - (long long)v * L.P);
        const float depth = L.g_rec[(size_t)idx * L.rec + rec_depth_index(L.channels)];
        key = ((unsigned long long)__float_as_uint(depth) << 32) | g;
        tbase = (size_t)v * L.tiles;
    }
    if (!use_smem) {
        for (int y = miny; y < maxy; ++y)
            for (int x = minx; x < maxx; ++x) {
                const size_t t = tbase + (size_t)y * L.gx + x;
                const unsigned int slot = atomicAdd(&L.tile_cursor[t], 1u);
                L.keys[L.tile_offset[t] + slot] = key;
            }
        return;
    }
    for (int i = threadIdx.x; i < L.tiles; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    for (int y = miny; y < maxy; ++y)
        for (int x = minx; x < maxx; ++x) atomicAdd(&hist[y * L.gx + x], 1u);
    __syncthreads();
    const size_t vb = (size_t)v_first * L.tiles;
    for (int i = threadIdx.x; i < L.tiles; i += blockDim.x) {
        const unsigned int c = hist[i];
        if (c) hist[i] = L.tile_offset[vb + i] + atomicAdd(&L.tile_cursor[vb + i], c);   // absolute position of this block's run
    }
    __syncthreads();
    for (int y = miny; y < maxy; ++y)
        for (int x = minx; x < maxx; ++x) L.keys[atomicAdd(&hist[y * L.gx + x], 1u)] = key;
}

// ---- per-tile sort ------------------------------------------------------------------------------
constexpr unsigned long long KEY_INF = 0xffffffffffffffffull;

// Single-direction bitonic network on `n` real keys padded virtually with +inf up to `npow2`:
// every compare-exchange puts the smaller key at the lower index, so the virtual +inf entries
// (indices >= n) never move and can simply be skipped.
__device__ __forceinline__ void cmpx(unsigned long long* k, int i, int j) {
    const unsigned long long a = k[i], b = k[j];
    if (a > b) { k[i] = b; k[j] = a; }
}

// All network steps whose span stays inside one SORT_CHUNK-aligned chunk held in shared memory.
// kbeg..kend: merge sizes to run (powers of two); for merge size k > cn only the half-cleaner steps
// with stride < cn are executed here.  Steps whose span is <= 64 elements are executed warp-locally
// (each warp owns 64-element spans, __syncwarp between steps), so only the steps with stride >= 64 cost a
// block barrier: 20 instead of 66 barriers for a 2048-key segment.
__device__ void smem_network(unsigned long long* sk, int cn /* pow2 <= SORT_CHUNK */, int kbeg, int kend) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int S = min(64, cn);                    // warp-local span
    for (int k = kbeg; k <= kend; k <<= 1) {
        const bool flip_local = k <= S;
        int jwarp;
        if (flip_local) {
            jwarp = k >> 2;
        } else {
            int jtop;
            if (k <= cn) {
                const int hm = (k >> 1) - 1;      // k, j are powers of two: index math with masks, no divisions
                for (int t = threadIdx.x; t < cn / 2; t += blockDim.x) {
                    const int off = t & hm, base = (t & ~hm) << 1;
                    cmpx(sk, base + off, base + (k - 1 - off));
                }
                __syncthreads();
                jtop = k >> 2;
            } else {
                jtop = cn >> 1;
            }
            for (int j = jtop; j >= 64; j >>= 1) {
                for (int t = threadIdx.x; t < cn / 2; t += blockDim.x) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    cmpx(sk, i, i + j);
                }
                __syncthreads();
            }
            jwarp = min(jtop, 32);
        }
        for (int sp = warp; sp < cn / S; sp += nwarps) {
            unsigned long long* base = sk + sp * S;
            if (flip_local) {
                if (lane < S / 2) {
                    const int hm = (k >> 1) - 1;
                    const int off = lane & hm, b0 = (lane & ~hm) << 1;
                    cmpx(base, b0 + off, b0 + (k - 1 - off));
                }
                __syncwarp();
            }
            for (int j = jwarp; j > 0; j >>= 1) {
                if (lane < S / 2) {
                    const int i = ((lane & ~(j - 1)) << 1) | (lane & (j - 1));
                    cmpx(base, i, i + j);
                }
                __syncwarp();
            }
        }
        if (2 * k > S || k == kend) __syncthreads();
    }
}

// Block-level merge sort of `n` keys held in shared memory (n <= E * blockDim.x): every thread first sorts E
// consecutive keys in registers, then log2(n / E) merge passes in which each thread produces E consecutive
// outputs of its pair of runs (merge-path binary search for the split, then a sequential merge).  About 4x
// fewer instructions per key than the bitonic network; keys are unique, so the result is the unique sorted order.
constexpr int MS_E = 8;

__device__ __forceinline__ void cmpswap(unsigned long long& a, unsigned long long& b) {
    const unsigned long long lo = a < b ? a : b, hi = a < b ? b : a;
    a = lo; b = hi;
}

// Sorting network on E (power of two) keys in registers; every index is a compile-time constant after unrolling.
// E = 8: Batcher's odd-even merge sort (19 comparators); larger E: bitonic (simple loop nest that always unrolls).
template <int E>
__device__ __forceinline__ void sort_registers(unsigned long long (&r)[E]) {
    if constexpr (E == 8) {
        cmpswap(r[0], r[1]); cmpswap(r[2], r[3]); cmpswap(r[4], r[5]); cmpswap(r[6], r[7]);
        cmpswap(r[0], r[2]); cmpswap(r[1], r[3]); cmpswap(r[4], r[6]); cmpswap(r[5], r[7]);
        cmpswap(r[1], r[2]); cmpswap(r[5], r[6]);
        cmpswap(r[0], r[4]); cmpswap(r[1], r[5]); cmpswap(r[2], r[6]); cmpswap(r[3], r[7]);
        cmpswap(r[2], r[4]); cmpswap(r[3], r[5]);
        cmpswap(r[1], r[2]); cmpswap(r[3], r[4]); cmpswap(r[5], r[6]);
    } else {
#pragma unroll
        for (int k = 2; k <= E; k <<= 1) {
#pragma unroll
            for (int i = 0; i < E; ++i) {                       // flip step: i <-> its mirror inside the k-block
                const int l = (i & ~(k - 1)) + (k - 1 - (i & (k - 1)));
                if (l > i) cmpswap(r[i], r[l]);
            }
#pragma unroll
            for (int j = k >> 2; j > 0; j >>= 1) {
#pragma unroll
                for (int i = 0; i < E; ++i)
                    if ((i & j) == 0) cmpswap(r[i], r[i | j]);
            }
        }
    }
}

template <int E>
__device__ void smem_merge_sort(unsigned long long* sk, int n) {
    const int npad = (n + E - 1) / E * E;
    const int base = threadIdx.x * E;
    const bool active = base < npad;
    unsigned long long r[E];
    if (active) {
#pragma unroll
        for (int e = 0; e < E; ++e) r[e] = sk[base + e];
        sort_registers<E>(r);
#pragma unroll
        for (int e = 0; e < E; ++e) sk[base + e] = r[e];
    }
    __syncthreads();
    for (int w = E; w < npad; w <<= 1) {
        if (active) {
            const int a0 = base & ~(2 * w - 1);            // start of this thread's pair of runs
            const int b0 = a0 + w;
            const int lenA = min(w, npad - a0);
            const int lenB = max(0, min(w, npad - b0));
            const unsigned long long* A = sk + a0;
            const unsigned long long* B = sk + b0;
            const int diag = base - a0;                      // outputs [diag, diag + E) of the merged pair
            int lo = max(0, diag - lenB), hi = min(diag, lenA);
            while (lo < hi) {                                // merge path: number of A elements among the first `diag`
                const int mid = (lo + hi) >> 1;
                if (A[mid] <= B[diag - 1 - mid]) lo = mid + 1; else hi = mid;
            }
            int i = lo, j = diag - lo;
            unsigned long long av = i < lenA ? A[i] : KEY_INF, bv = j < lenB ? B[j] : KEY_INF;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const bool takeA = (j >= lenB) || (i < lenA && av <= bv);
                r[e] = takeA ? av : bv;
                if (takeA) { ++i; av = i < lenA ? A[i] : KEY_INF; }
                else { ++j; bv = j < lenB ? B[j] : KEY_INF; }
            }
        }
        __syncthreads();
        if (active) {
#pragma unroll
            for (int e = 0; e < E; ++e) sk[base + e] = r[e];
        }
        __syncthreads();
    }
}

// Copies one projected record into the sorted stream, replacing the contribution threshold by the tile-local
// 16-bit CELL MASK: bit (4 cy + cx) set <=> some pixel of the 4x4 cell (cx, cy) of the tile can reach
// alpha >= 1/255, i.e. the cell's rectangle meets the ellipse E = {q(x,y) = cx x^2 + 2 cy x y + cz y^2 <= thr}.
// Exact test, one cell ROW at a time: E cut by the row's horizontal strip y0 <= y <= y1 is convex, so it meets a
// cell [x0,x1] x [y0,y1] iff its projection [xl, xr] on the x axis overlaps [x0, x1].  xr is the ellipse's
// rightmost point +X if that point lies inside the strip, else the larger right end of the two chords cut by the
// strip's edges (clamped to the ellipse's own y range); xl likewise.  About 160 instructions per instance instead of
// the 1300 of testing the four edges of all 16 rectangles.  Conservative by the 0.01 px margins (the arithmetic error
// is below 1e-4 px); never drops a contributing pixel.  The render kernels walk, per half-warp, only the instances
// whose bit for its cell is set.
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ unsigned int cell_mask(const float4 a, const float4 b, float tile_x0, float tile_y0) {
    const float thr = b.z;
    unsigned int mask = 0u;
    if (thr > 0.f) {
        const float cx = a.z, cy = a.w, cz = b.x;
        const float det = cx * cz - cy * cy;
        if (!(cx > 0.f) || !(cz > 0.f) || !(det > 0.f) || !(thr < 1e29f)) {
            mask = DM4D_CELL_ROWS == 4 ? 0xffffu : 0xffffffffu;
        } else {
            const float icx = 1.f / cx, idet = 1.f / det;
            const float X = sqrt_approx(thr * cz * idet), Y = sqrt_approx(thr * cx * idet);   // half extents of E
            const float yX = -cy * X / cz;                 // y of the rightmost point (+X, yX); leftmost is (-X, -yX)
            const float tcx = thr * cx;
            float x0[4];                                   // left edges of the cell columns (pixel - centre, with margin)
#pragma unroll
            for (int i = 0; i < 4; ++i) x0[i] = tile_x0 + (float)(4 * i) - a.x - 0.01f;
            constexpr int CR = DM4D_CELL_ROWS;                              // cell height: 4 (16-bit mask) or 2 (32-bit mask)
#pragma unroll
            for (int j = 0; j < 16 / CR; ++j) {
                const float y0 = tile_y0 + (float)(CR * j) - a.y - 0.01f, y1 = y0 + (CR == 4 ? 3.02f : 1.02f);
                if (y0 > Y || y1 < -Y) continue;                                   // the strip misses E
                const float ya = fminf(fmaxf(y0, -Y), Y), yb = fminf(fmaxf(y1, -Y), Y);
                const float da = sqrt_approx(fmaxf(tcx - det * ya * ya, 0.f)), db = sqrt_approx(fmaxf(tcx - det * yb * yb, 0.f));
                float xr = fmaxf((da - cy * ya) * icx, (db - cy * yb) * icx);
                float xl = fminf((-da - cy * ya) * icx, (-db - cy * yb) * icx);
                if (yX >= y0 && yX <= y1) xr = X;
                if (-yX >= y0 && -yX <= y1) xl = -X;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (xl <= x0[i] + 3.02f && xr >= x0[i]) mask |= 1u << (4 * j + i);
            }
        }
    }
    return mask;
}

// Gathers the records of sorted instances [0, n) of a tile (ids in `keys`) into the packed stream.  Two instances per
// thread and iteration with DM4D_PACK_UNROLL=2 (both gathers in flight before either mask is computed); measured
// no better than 1 (the extra staging registers cost as much occupancy as the overlap gains), so 1 is the default.
#ifndef DM4D_PACK_UNROLL
#define DM4D_PACK_UNROLL 1
#endif
template <int R4>
__device__ __forceinline__ void pack_tile(const unsigned long long* keys, int n, const float4* __restrict__ grec,
                                          float4* __restrict__ srec, float tile_x0, float tile_y0) {
    constexpr int U = DM4D_PACK_UNROLL;
    for (int i0 = threadIdx.x; i0 < n; i0 += U * blockDim.x) {
        float4 r[U][R4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = min(i0 + u * (int)blockDim.x, n - 1);
            const float4* src = grec + (size_t)(unsigned int)(keys[i] & 0xffffffffull) * R4;
#pragma unroll
            for (int q = 0; q < R4; ++q) r[u][q] = src[q];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * (int)blockDim.x;
            if (i < n) {
                r[u][1].z = __uint_as_float(cell_mask(r[u][0], r[u][1], tile_x0, tile_y0));
                float4* dst = srec + (size_t)i * R4;
#pragma unroll
                for (int q = 0; q < R4; ++q) dst[q] = r[u][q];
            }
        }
    }
}

// Several CTA shapes share the tiles by segment length, all sorting in shared memory: 128 threads x 8 keys up to 1024
// keys, 256 x 8 up to 2048, 512 x 8 up to 4096, 1024 x 8 up to 8192, 1024 x 16 up to 16384 (128 KB of shared memory; the limb / centre tiles of C3 hold ~6000
// instances, C4's densest ~10^4).  Small CTAs keep more tiles resident per SM and their pass barriers span 4 warps.
// Only segments above 16384 keys fall back to chunked bitonic merging through global memory — that path took 0.25 ms for
// a 6000-key tile and was the whole tail of the launch when the largest in-memory tier was 4096.
#ifndef DM4D_SORT_M_BLOCKS
#define DM4D_SORT_M_BLOCKS 3      // resident CTAs per SM the 512-thread tier's register budget is sized for
#endif
#ifndef DM4D_SORT_S_BLOCKS
#define DM4D_SORT_S_BLOCKS 10     // ... and the 128-thread tier's
#endif
// Finer tiers: a 256 x 8 tier for 1025..2048 keys (the 512-thread CTAs ran such a tile with half of their threads idle
// and 16-warp barriers; on by default: -2 % step time at C4, neutral at C3) and an optional 64 x 8 tier for tiles of at
// most 512 keys (no gain measured, off; profiles/r2aa_tune_sort_tiers.txt).
#ifndef DM4D_SORT_TIER256
#define DM4D_SORT_TIER256 1
#endif
#ifndef DM4D_SORT_TIER64
#define DM4D_SORT_TIER64 0
#endif
#ifndef DM4D_SORT_H_BLOCKS
#define DM4D_SORT_H_BLOCKS 6      // resident CTAs per SM the 256-thread tier's register budget is sized for
#endif
#ifndef DM4D_SORT_T_BLOCKS
#define DM4D_SORT_T_BLOCKS 16     // ... and the 64-thread tier's
#endif
constexpr int sort_min_blocks(int threads) {
    return threads == 1024 ? 1 : threads == 512 ? DM4D_SORT_M_BLOCKS : threads == 256 ? DM4D_SORT_H_BLOCKS
         : threads == 128 ? DM4D_SORT_S_BLOCKS : DM4D_SORT_T_BLOCKS;
}
template <int THREADS, int E, int R4>
__global__ void __launch_bounds__(THREADS, sort_min_blocks(THREADS)) sort_pack_kernel(RasterLayout L, int n_lo, int n_hi) {
    constexpr int SORT_CHUNK = THREADS * E;
    extern __shared__ __align__(16) unsigned long long sk[];   // [SORT_CHUNK]
    if (L.hdr->overflow) return;
    // tile_order is sorted by floor(log2(count)) descending: a CTA walks the list grid-stride and stops at the first
    // tile below its tier's range (the heavy tiers run as a few persistent CTAs over the head of the list)
    const int n_all = L.n_views * L.tiles;
    for (int it = blockIdx.x; it < n_all; it += gridDim.x) {
    const int tile = (int)L.tile_order[it];           // global (view, tile) index, heaviest first
    const unsigned int beg = L.tile_offset[tile];
    const int n = (int)(L.tile_offset[tile + 1] - beg);
    if (__clz(n) > __clz(n_lo + 1)) break;
    if (n <= n_lo || n > n_hi) continue;
    __syncthreads();                                  // the previous tile's pack still reads sk
    const int v = tile / L.tiles;
    unsigned long long* gk = L.keys + beg;
    const float4* grec = reinterpret_cast<const float4*>(L.g_rec + (size_t)v * L.P * L.rec);
    float4* srec = reinterpret_cast<float4*>(L.stream + (size_t)beg * L.rec);
    const float tile_y0 = (float)(((tile - v * L.tiles) / L.gx) * DM4D_TILE);
    const float tile_x0 = (float)(((tile - v * L.tiles) % L.gx) * DM4D_TILE);

    if (n <= SORT_CHUNK) {
        const int nfill = (n + E - 1) / E * E;                       // the merge sort pads to a multiple of E
        for (int i = threadIdx.x; i < nfill; i += blockDim.x) sk[i] = i < n ? gk[i] : KEY_INF;
        __syncthreads();
        smem_merge_sort<E>(sk, n);
        __syncthreads();
        pack_tile<R4>(sk, n, grec, srec, tile_x0, tile_y0);
        continue;
    }

    // Very large segment: sort SORT_CHUNK-sized chunks in shared memory, then merge with global
    // compare-exchange steps for strides >= SORT_CHUNK and shared-memory steps below.
    int npow2 = 2;
    while (npow2 < n) npow2 <<= 1;
    const int nchunks = npow2 / SORT_CHUNK;
    for (int c = 0; c < nchunks; ++c) {
        const int base = c * SORT_CHUNK;
        if (base >= n) break;
        for (int i = threadIdx.x; i < SORT_CHUNK; i += blockDim.x) sk[i] = base + i < n ? gk[base + i] : KEY_INF;
        __syncthreads();
        smem_network(sk, SORT_CHUNK, 2, SORT_CHUNK);
        for (int i = threadIdx.x; i < SORT_CHUNK; i += blockDim.x)
            if (base + i < n) gk[base + i] = sk[i];
        __syncthreads();
    }
    for (int k = 2 * SORT_CHUNK; k <= npow2; k <<= 1) {
        // flip step (span up to k) in global memory
        for (int t = threadIdx.x; t < npow2 / 2; t += blockDim.x) {
            const int hm = (k >> 1) - 1;
            const int off = t & hm, base = (t & ~hm) << 1;
            const int i = base + off, j = base + (k - 1 - off);
            if (j < n) cmpx(gk, i, j);
        }
        __syncthreads();
        for (int j = k >> 2; j >= SORT_CHUNK; j >>= 1) {
            for (int t = threadIdx.x; t < npow2 / 2; t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                if (i + j < n) cmpx(gk, i, i + j);
            }
            __syncthreads();
        }
        for (int c = 0; c < nchunks; ++c) {
            const int base = c * SORT_CHUNK;
            if (base >= n) break;
            for (int i = threadIdx.x; i < SORT_CHUNK; i += blockDim.x) sk[i] = base + i < n ? gk[base + i] : KEY_INF;
            __syncthreads();
            smem_network(sk, SORT_CHUNK, k, k);
            for (int i = threadIdx.x; i < SORT_CHUNK; i += blockDim.x)
                if (base + i < n) gk[base + i] = sk[i];
            __syncthreads();
        }
    }
    __syncthreads();
    pack_tile<R4>(gk, n, grec, srec, tile_x0, tile_y0);
    }
}

__global__ void export_state_kernel(RasterLayout L, int view, unsigned int* ranges, unsigned int* point_list,
                                    long long cap, unsigned int* n_contrib) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int vbeg = L.tile_offset[(size_t)view * L.tiles];
    const unsigned int vend = L.tile_offset[(size_t)(view + 1) * L.tiles];
    if (ranges && i < L.tiles) {
        const unsigned int b = L.tile_offset[(size_t)view * L.tiles + i], e = L.tile_offset[(size_t)view * L.tiles + i + 1];
        // empty tiles read (0,0) in the replaced rasterizer (zero-initialised ranges)
        ranges[2 * i] = b == e ? 0u : b - vbeg;
        ranges[2 * i + 1] = b == e ? 0u : e - vbeg;
    }
    if (point_list && i < (long long)(vend - vbeg) && i < cap)
        point_list[i] = __float_as_uint(L.stream[(size_t)(vbeg + i) * L.rec + 7]);
    if (n_contrib && i < (long long)L.H * L.W) n_contrib[i] = L.n_contrib[(size_t)view * L.H * L.W + i];
}

}  // namespace

template <int R4>
int launch_sort_pack_t(const RasterLayout& L, cudaStream_t s) {
    static bool configured = false;
    constexpr int K_S = 128 * 8, K_M = 512 * 8, K_L = 1024 * 8, K_X = 1024 * 16;
    if (!configured) {
        DM4D_CUDA_CHECK(cudaFuncSetAttribute(sort_pack_kernel<512, 8, R4>, cudaFuncAttributeMaxDynamicSharedMemorySize, K_M * 8));
        DM4D_CUDA_CHECK(cudaFuncSetAttribute(sort_pack_kernel<1024, 8, R4>, cudaFuncAttributeMaxDynamicSharedMemorySize, K_L * 8));
        DM4D_CUDA_CHECK(cudaFuncSetAttribute(sort_pack_kernel<1024, 16, R4>, cudaFuncAttributeMaxDynamicSharedMemorySize, K_X * 8));
        configured = true;
    }
    const int n_all = L.n_views * L.tiles;
    auto grid = [&](int per_sm) { return (unsigned)std::max(1, std::min(n_all, 148 * per_sm)); };
    // The tiers are independent: fork them onto side streams (works inside a stream capture too — the side
    // streams join the capture through the events), so the few CTAs of the heavy tiers (one 10^4-key tile takes a
    // 1024-thread CTA ~0.1 ms) run under the bulk of the small tiles instead of in front of them.
    // one set of side streams / events per device (one process per GPU is the design, but a process may own several)
    constexpr int MAX_DEV = 16, NSIDE = 3 + DM4D_SORT_TIER256 + DM4D_SORT_TIER64;
    static cudaStream_t side_all[MAX_DEV][NSIDE] = {};
    static cudaEvent_t fork_all[MAX_DEV] = {}, join_all[MAX_DEV][NSIDE] = {};
    int dev = 0;
    DM4D_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEV) { dm4d_set_error("sort_pack: device ordinal %d out of range", dev); return DM4D_EINVAL; }
    cudaStream_t* side = side_all[dev];
    cudaEvent_t& fork_ev = fork_all[dev];
    cudaEvent_t* join_ev = join_all[dev];
    if (!fork_ev) {
        for (int i = 0; i < NSIDE; ++i) {
            DM4D_CUDA_CHECK(cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking));
            DM4D_CUDA_CHECK(cudaEventCreateWithFlags(&join_ev[i], cudaEventDisableTiming));
        }
        DM4D_CUDA_CHECK(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming));
    }
    DM4D_CUDA_CHECK(cudaEventRecord(fork_ev, s));
    for (int i = 0; i < NSIDE; ++i) DM4D_CUDA_CHECK(cudaStreamWaitEvent(side[i], fork_ev, 0));
    constexpr int K_H = 256 * 8, K_T = 64 * 8;
    sort_pack_kernel<1024, 16, R4><<<grid(1), 1024, K_X * 8, side[0]>>>(L, K_L, 0x7fffffff);
    sort_pack_kernel<1024, 8, R4><<<grid(1), 1024, K_L * 8, side[1]>>>(L, K_M, K_L);
    sort_pack_kernel<512, 8, R4><<<grid(4), 512, K_M * 8, side[2]>>>(L, DM4D_SORT_TIER256 ? K_H : K_S, K_M);
#if DM4D_SORT_TIER256
    sort_pack_kernel<256, 8, R4><<<grid(8), 256, K_H * 8, side[3]>>>(L, K_S, K_H);
#endif
#if DM4D_SORT_TIER64
    // tile_order is heaviest first: the light tier's CTAs skip the head of the list with a few loads per tile
    sort_pack_kernel<64, 8, R4><<<grid(32), 64, K_T * 8, side[NSIDE - 1]>>>(L, 0, K_T);
    sort_pack_kernel<128, 8, R4><<<(unsigned)n_all, 128, K_S * 8, s>>>(L, K_T, K_S);
#else
    sort_pack_kernel<128, 8, R4><<<(unsigned)n_all, 128, K_S * 8, s>>>(L, 0, K_S);
#endif
    for (int i = 0; i < NSIDE; ++i) {
        DM4D_CUDA_CHECK(cudaEventRecord(join_ev[i], side[i]));
        DM4D_CUDA_CHECK(cudaStreamWaitEvent(s, join_ev[i], 0));
    }
    return DM4D_OK;
}

int launch_scan(const RasterLayout& L, cudaStream_t s) {
    { KernelTimer kt(DM4D_K_SCAN, s); scan_tiles_kernel<<<1, SCAN_THREADS, 0, s>>>(L); }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

int launch_scatter_sort_pack(const RasterLayout& L, cudaStream_t s) {
    const long long n = (long long)L.n_views * L.P;
    if (n == 0) return DM4D_OK;
    { KernelTimer kt(DM4D_K_SCATTER, s); scatter_kernel<<<(unsigned)((n + DM4D_BLOCK - 1) / DM4D_BLOCK), DM4D_BLOCK, 0, s>>>(L); }
    DM4D_CUDA_CHECK(cudaGetLastError());
    {
        KernelTimer kt(DM4D_K_SORT_PACK, s);
        const int rc = L.rec == 12 ? launch_sort_pack_t<3>(L, s) : launch_sort_pack_t<4>(L, s);
        if (rc) return rc;
    }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

// ---- feature re-binding (second rasterizer call of a view: same geometry, other per-Gaussian features) ----------------
namespace {
struct RebindArgs {
    int P, rec, channels, n_sets;
    const float* colors;  long long colors_stride;
    const float* colors2; long long colors2_stride;
    const float* view_params;
    const unsigned int* view_first;   // tile_offset of the source plan: view v owns instances [view_first[v*tiles], view_first[(v+1)*tiles])
    int n_views, tiles;
};
// One thread per sorted instance: copy the geometric half of the record (position, conic, opacity, cell mask, id, depth)
// and gather the new features by Gaussian id.  Reads 48/64 B + 12/24 B, writes 48/64 B per instance.
__global__ void __launch_bounds__(256) rebind_features_kernel(RebindArgs a, const BinHeader* __restrict__ hdr,
                                                              const float4* __restrict__ src, float4* __restrict__ dst) {
    const unsigned long long total = hdr->overflow ? 0ull : hdr->total;
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int r4 = a.rec / 4;
    const float4 r0 = src[i * r4 + 0], r1 = src[i * r4 + 1];
    const int id = __float_as_int(r1.w);
    // view of this instance: binary search over the per-view first offsets (n_views is small)
    int v = 0;
    {
        int lo = 0, hi = a.n_views - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((unsigned long long)a.view_first[(size_t)mid * a.tiles] <= i) lo = mid; else hi = mid - 1;
        }
        v = lo;
    }
    const float* vp = a.view_params + (size_t)v * DM4D_VIEW_STRIDE;
    const long long set = min(max((long long)vp[DM4D_VIEW_SET], 0ll), (long long)a.n_sets - 1);
    const float* c0 = a.colors + set * a.colors_stride + (size_t)id * 3;
    dst[i * r4 + 0] = r0;
    dst[i * r4 + 1] = r1;
    if (a.channels <= 3) {
        const float depth = src[i * r4 + 2].w;
        dst[i * r4 + 2] = make_float4(c0[0], c0[1], c0[2], depth);
    } else {
        const float depth = src[i * r4 + 3].z;
        const float* c1 = a.colors2 + set * a.colors2_stride + (size_t)id * 3;
        dst[i * r4 + 2] = make_float4(c0[0], c0[1], c0[2], c1[0]);
        dst[i * r4 + 3] = make_float4(c1[1], c1[2], depth, 0.f);
    }
}
}  // namespace

int launch_rebind_features(const dm4d_raster_desc* d, const RasterLayout& src, const RasterLayout& dst, cudaStream_t s) {
    // header + tile tables (count, offset, cursor, order) are contiguous at the start of the bin workspace
    const size_t head = (size_t)((const char*)src.keys - (const char*)src.hdr);
    DM4D_CUDA_CHECK(cudaMemcpyAsync(dst.hdr, src.hdr, head, cudaMemcpyDeviceToDevice, s));
    if (src.capacity == 0) return DM4D_OK;
    RebindArgs a;
    a.P = src.P; a.rec = src.rec; a.channels = src.channels; a.n_sets = d->n_sets;
    a.colors = d->colors; a.colors_stride = d->colors_stride;
    a.colors2 = d->colors2; a.colors2_stride = d->colors2_stride;
    a.view_params = d->view_params; a.view_first = src.tile_offset; a.n_views = src.n_views; a.tiles = src.tiles;
    const unsigned blocks = (unsigned)((src.capacity + 255) / 256);
    rebind_features_kernel<<<blocks, 256, 0, s>>>(a, src.hdr, reinterpret_cast<const float4*>(src.stream),
                                                  reinterpret_cast<float4*>(dst.stream));
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

int launch_export_state(const RasterLayout& L, int view, unsigned int* ranges, unsigned int* point_list,
                        long long cap, unsigned int* n_contrib, cudaStream_t s) {
    long long n = L.tiles;
    if ((long long)L.H * L.W > n) n = (long long)L.H * L.W;
    if (point_list && L.capacity > n) n = L.capacity;
    if (n == 0) return DM4D_OK;
    export_state_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(L, view, ranges, point_list, cap, n_contrib);
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}
