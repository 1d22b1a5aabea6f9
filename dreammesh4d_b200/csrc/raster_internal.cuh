// Internal layout shared by the rasterizer translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dm4d.h"

#define DM4D_BLOCK 256

// Height of a culling cell (cells are 4 pixels wide).  4 (default, validated): 4x4 cells, one instance queue per
// HALF-warp, 16-bit masks.  2 (EXPERIMENTAL, prepared at the end of round 1 and not yet run on a GPU): 4x2 cells, one
// queue per QUARTER-warp, 32-bit masks — 1.07 instead of 1.21 warp iterations per instance in the CPU simulation
// (scripts/sim_cell_queues.py).  raster_binning.cu (mask) and raster_render.cu (queues, reductions) must agree.
#ifndef DM4D_CELL_ROWS
#define DM4D_CELL_ROWS 4
#endif

struct BinHeader {
    unsigned long long total;   // R = number of (Gaussian, tile) instances over all views
    unsigned int overflow;      // 1 if R > capacity (nothing rendered)
    unsigned int pad;
};

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
int launch_render_backward(const dm4d_raster_desc* d, const RasterLayout& L, const float* out_alpha,
                           const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha, cudaStream_t s);
int launch_rebind_features(const dm4d_raster_desc* d, const RasterLayout& src, const RasterLayout& dst, cudaStream_t s);
int launch_export_state(const RasterLayout& L, int view, unsigned int* ranges, unsigned int* point_list,
                        long long cap, unsigned int* n_contrib, cudaStream_t s);

// Optional per-kernel event timing (dm4d_profile_*). Usage: { KernelTimer t(DM4D_K_X, stream); kernel<<<...>>>(); }
struct KernelTimer {
    int id; cudaStream_t s; bool on; cudaEvent_t e0;
    KernelTimer(int id_, cudaStream_t s_);
    ~KernelTimer();
};
