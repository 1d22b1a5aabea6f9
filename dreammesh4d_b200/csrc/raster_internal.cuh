// Internal layout shared by the rasterizer translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dm4d.h"

#define DM4D_BLOCK 256

// Height of a culling cell (cells are 4 pixels wide): 4x4 cells, one instance queue per HALF-warp, 16-bit masks.
// A 4x2 variant (quarter-warp queues, 32-bit masks) was built and measured in round 2: parity green, step time
// 2.064 -> 2.006 ms (-3 %, profiles/r2a_tune_cells4x2_fastexp.txt) — not worth the second code path, whose render side
// was removed; the mask generator in raster_binning.cu still takes the cell height as a parameter.
#ifndef DM4D_CELL_ROWS
#define DM4D_CELL_ROWS 4
#endif

struct BinHeader {
    unsigned long long total;   // R = number of (Gaussian, tile) instances over all views
    unsigned int overflow;      // 1 if R > capacity (nothing rendered)
    unsigned int total_segs;    // number of (tile, segment) work items of the backward (see DM4D_SEG)
};

// The backward splits every tile's depth-sorted instance list into SEGMENTS of DM4D_SEG instances that are processed
// independently (front to back, starting from the compositing state the forward checkpointed at the segment's first
// instance), so the serial chain of a heavy tile is bounded by one segment instead of the whole list.
#ifndef DM4D_SEG
#define DM4D_SEG 1024
#endif
__host__ __device__ inline long long seg_capacity(long long capacity, long long n_tiles) { return capacity / DM4D_SEG + n_tiles + 1; }
// checkpoint of one (segment, pixel): C accumulated features, accumulated depth, transmittance
__host__ __device__ inline int ckpt_floats(int channels) { return channels + 2; }

// Projected per-(view, Gaussian) record; also the element of the sorted instance stream.
//   f[0..3]  = x, y, conic.x, conic.y
//   f[4..7]  = conic.z, opacity, W, id (int bits)
//              W = per-Gaussian contribution threshold on the conic quadratic form (geom records), replaced
//              per (instance, tile) by the 16-bit CELL MASK (instance stream): bit 4 cy + cx set <=> some pixel of
//              the 4x4 cell (cx, cy) of the tile can reach alpha >= 1/255
//   3 ch: f[8..11]  = r, g, b, depth            6 ch: f[8..15] = r, g, b, n0, n1, n2, depth, 0
__host__ __device__ inline int rec_floats(int channels) { return channels <= 3 ? 12 : 16; }
__host__ __device__ inline int rec_depth_index(int channels) { return channels <= 3 ? 11 : 14; }
// Backward accumulator row per (view, Gaussian):
//   [0..1] dL/dmean2D, [2..4] dL/dconic (x, y(half), z), [5] dL/dopacity, [6] dL/ddepth, [7] pad,
//   [8..8+C) dL/dfeature, padded to a multiple of 4
__host__ __device__ inline int acc_floats(int channels) { return channels <= 3 ? 12 : 16; }

struct RasterLayout {
    int P, H, W, n_views, channels, gx, gy, tiles, rec, acc;
    long long capacity;
    // geom
    float* g_rec;               // [n_views*P*rec]
    unsigned int* g_rect;       // [n_views*P] packed minx | miny<<8 | maxx<<16 | maxy<<24 (0 = culled)
    // bin
    BinHeader* hdr;
    unsigned int* tile_count;   // [n_views*tiles]
    unsigned int* tile_offset;  // [n_views*tiles + 1]
    unsigned int* tile_cursor;  // [n_views*tiles]
    unsigned int* tile_order;   // [n_views*tiles] (view,tile) indices, heaviest tiles first (CTA launch order)
    unsigned int* seg_offset;   // [n_views*tiles + 1] first segment index of each (view,tile): exclusive scan of ceil(count / DM4D_SEG)
    unsigned int* seg_tile;     // [seg_cap] (view,tile) index of every segment
    long long seg_cap;          // capacity / DM4D_SEG + n_views*tiles
    float* ckpt;                // [seg_cap][channels+2][256] forward state at the first instance of every segment > 0 of its tile
    unsigned long long* keys;   // [capacity]  (depth bits << 32) | gaussian id
    float* stream;              // [capacity*rec] sorted instance records
    // img
    unsigned int* n_contrib;    // [n_views*H*W]
    // bwd
    float* accum;               // [n_views*P*acc]
};

static inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// Computes section sizes; if base pointers are given also fills the layout. Returns 0 or DM4D_E*.
int raster_make_layout(const dm4d_raster_desc* d, RasterLayout* L);
void raster_sizes(int P, int H, int W, int n_views, int channels, long long capacity, uint64_t* geom, uint64_t* bin,
                  uint64_t* img, uint64_t* bwd);

void dm4d_set_error(const char* fmt, ...);
#define DM4D_CUDA_CHECK(expr)                                                                   \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            dm4d_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return DM4D_ECUDA;                                                                  \
        }                                                                                       \
    } while (0)

// launchers implemented in the kernel translation units
int launch_preprocess(const dm4d_raster_desc* d, const RasterLayout& L, int32_t* radii, cudaStream_t s);
int launch_preprocess_backward(const dm4d_raster_desc* d, const RasterLayout& L, float* dL_dmeans3D,
                               float* dL_dmeans2D, float* dL_dcolors, float* dL_dcolors2, float* dL_dopacities,
                               float* dL_dscales, float* dL_drotations, cudaStream_t s);
int launch_scan(const RasterLayout& L, cudaStream_t s);
int launch_scatter_sort_pack(const RasterLayout& L, cudaStream_t s);
int launch_render_forward(const dm4d_raster_desc* d, const RasterLayout& L, float* out_color, float* out_depth,
                          float* out_alpha, cudaStream_t s);
int launch_render_backward(const dm4d_raster_desc* d, const RasterLayout& L, const float* out_color, const float* out_depth,
                           const float* out_alpha, const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                           cudaStream_t s);
int launch_rebind_features(const dm4d_raster_desc* d, const RasterLayout& src, const RasterLayout& dst, cudaStream_t s);
int launch_export_state(const RasterLayout& L, int view, unsigned int* ranges, unsigned int* point_list,
                        long long cap, unsigned int* n_contrib, cudaStream_t s);

// Optional per-kernel event timing (dm4d_profile_*). Usage: { KernelTimer t(DM4D_K_X, stream); kernel<<<...>>>(); }
struct KernelTimer {
    int id; cudaStream_t s; bool on; cudaEvent_t e0;
    KernelTimer(int id_, cudaStream_t s_);
    ~KernelTimer();
};
