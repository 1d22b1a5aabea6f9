// extern "C" entry points of libdm4d.so (declared in include/dm4d.h).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>
#include "raster_internal.cuh"

static thread_local char g_err[512] = "";

void dm4d_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* dm4d_last_error(void) { return g_err; }
extern "C" int dm4d_version(void) { return 140; }   // 1.4: node-centric skinning backward (dm4d_skin_node_incidence, two desc fields)

void raster_sizes(int P, int H, int W, int n_views, int channels, long long capacity, uint64_t* geom, uint64_t* bin,
                  uint64_t* img, uint64_t* bwd) {
    const int gx = (W + DM4D_TILE - 1) / DM4D_TILE, gy = (H + DM4D_TILE - 1) / DM4D_TILE;
    const uint64_t nt = (uint64_t)n_views * gx * gy, np = (uint64_t)n_views * P;
    const int rec = rec_floats(channels), acc = acc_floats(channels);
    if (geom) *geom = 256 + align_up(np * rec * 4, 256) + align_up(np * 4, 256);   // never zero-sized
    const uint64_t segs = (uint64_t)seg_capacity(capacity, (long long)nt);
    if (bin)
        *bin = 256 + align_up(nt * 4, 256) + align_up((nt + 1) * 4, 256) + align_up(nt * 4, 256) + align_up(nt * 4, 256) +
               align_up((nt + 1) * 4, 256) + align_up(segs * 4, 256) +
               align_up((uint64_t)capacity * 8, 256) + align_up((uint64_t)capacity * rec * 4, 256) +
               align_up(segs * 256 * ckpt_floats(channels) * 4, 256);
    if (img) *img = align_up((uint64_t)n_views * H * W * 4, 256);
    if (bwd) *bwd = 256 + align_up(np * acc * 4, 256);
}

int raster_make_layout(const dm4d_raster_desc* d, RasterLayout* L) {
    if (!d) { dm4d_set_error("desc is NULL"); return DM4D_EINVAL; }
    if (d->P < 0 || d->H <= 0 || d->W <= 0 || d->n_views <= 0 || d->n_sets <= 0) {
        dm4d_set_error("bad sizes P=%d H=%d W=%d n_views=%d n_sets=%d", d->P, d->H, d->W, d->n_views, d->n_sets);
        return DM4D_EINVAL;
    }
    if (d->channels != 3 && d->channels != 6) { dm4d_set_error("channels must be 3 or 6, got %d", d->channels); return DM4D_EINVAL; }
    const int gx = (d->W + DM4D_TILE - 1) / DM4D_TILE, gy = (d->H + DM4D_TILE - 1) / DM4D_TILE;
    if (gx > 255 || gy > 255) { dm4d_set_error("image too large: %dx%d tiles (max 255 per side)", gx, gy); return DM4D_EINVAL; }
    if (d->bin_capacity < 0 || d->bin_capacity >= (1ll << 31)) { dm4d_set_error("bin_capacity out of range"); return DM4D_EINVAL; }
    if ((long long)d->n_views * d->P >= (1ll << 31)) { dm4d_set_error("n_views*P too large"); return DM4D_EINVAL; }
    if (d->P > 0 && (!d->means3D || (!d->cov3D && (!d->scales || !d->rotations)) || !d->opacities || !d->colors || !d->view_params ||
                     (d->channels == 6 && !d->colors2))) {
        dm4d_set_error("NULL input pointer");
        return DM4D_EINVAL;
    }
    uint64_t geom, bin, img, bwd;
    raster_sizes(d->P, d->H, d->W, d->n_views, d->channels, d->bin_capacity, &geom, &bin, &img, &bwd);
    if (!d->geom || !d->bin || !d->img || d->geom_bytes < geom || d->bin_bytes < bin || d->img_bytes < img) {
        dm4d_set_error("workspace too small: need geom=%llu bin=%llu img=%llu, got %llu %llu %llu",
                       (unsigned long long)geom, (unsigned long long)bin, (unsigned long long)img,
                       (unsigned long long)d->geom_bytes, (unsigned long long)d->bin_bytes, (unsigned long long)d->img_bytes);
        return DM4D_ENOSPC;
    }
    if (((uintptr_t)d->geom | (uintptr_t)d->bin | (uintptr_t)d->img | (uintptr_t)d->bwd) & 255) {
        dm4d_set_error("workspaces must be 256-byte aligned");
        return DM4D_EINVAL;
    }
    L->P = d->P; L->H = d->H; L->W = d->W; L->n_views = d->n_views; L->channels = d->channels;
    L->gx = gx; L->gy = gy; L->tiles = gx * gy;
    L->rec = rec_floats(d->channels); L->acc = acc_floats(d->channels);
    L->capacity = d->bin_capacity;
    const uint64_t nt = (uint64_t)d->n_views * L->tiles, np = (uint64_t)d->n_views * d->P;
    char* p = (char*)d->geom;
    L->g_rec = (float*)p; p += align_up(np * L->rec * 4, 256);
    L->g_rect = (unsigned int*)p;
    p = (char*)d->bin;
    L->hdr = (BinHeader*)p; p += 256;
    L->tile_count = (unsigned int*)p; p += align_up(nt * 4, 256);
    L->tile_offset = (unsigned int*)p; p += align_up((nt + 1) * 4, 256);
    L->tile_cursor = (unsigned int*)p; p += align_up(nt * 4, 256);
    L->tile_order = (unsigned int*)p; p += align_up(nt * 4, 256);
    L->seg_cap = seg_capacity(L->capacity, (long long)nt);
    L->seg_offset = (unsigned int*)p; p += align_up((nt + 1) * 4, 256);
    L->seg_tile = (unsigned int*)p; p += align_up((uint64_t)L->seg_cap * 4, 256);
    L->keys = (unsigned long long*)p; p += align_up((uint64_t)L->capacity * 8, 256);
    L->stream = (float*)p; p += align_up((uint64_t)L->capacity * L->rec * 4, 256);
    L->ckpt = (float*)p;
    L->n_contrib = (unsigned int*)d->img;
    L->accum = (float*)d->bwd;
    return DM4D_OK;
}

extern "C" int dm4d_raster_workspace_bytes(int32_t P, int32_t H, int32_t W, int32_t n_views, int32_t channels,
                                           int64_t bin_capacity, uint64_t* geom_bytes, uint64_t* bin_bytes,
                                           uint64_t* img_bytes, uint64_t* bwd_bytes) {
    if (P < 0 || H <= 0 || W <= 0 || n_views <= 0 || (channels != 3 && channels != 6) || bin_capacity < 0) {
        dm4d_set_error("dm4d_raster_workspace_bytes: bad argument");
        return DM4D_EINVAL;
    }
    raster_sizes(P, H, W, n_views, channels, bin_capacity, geom_bytes, bin_bytes, img_bytes, bwd_bytes);
    return DM4D_OK;
}

extern "C" int dm4d_raster_plan(const dm4d_raster_desc* d, int32_t* radii, int64_t* num_rendered_host, void* stream) {
    RasterLayout L;
    int rc = raster_make_layout(d, &L);
    if (rc) return rc;
    if (!radii && d->P > 0) { dm4d_set_error("radii is NULL"); return DM4D_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    DM4D_CUDA_CHECK(cudaMemsetAsync(L.tile_count, 0, (size_t)L.n_views * L.tiles * sizeof(unsigned int), s));
    if ((rc = launch_preprocess(d, L, radii, s))) return rc;
    if ((rc = launch_scan(L, s))) return rc;
    if (num_rendered_host) {
        BinHeader h;
        DM4D_CUDA_CHECK(cudaMemcpyAsync(&h, L.hdr, sizeof(h), cudaMemcpyDeviceToHost, s));
        DM4D_CUDA_CHECK(cudaStreamSynchronize(s));
        *num_rendered_host = (int64_t)h.total;
    }
    return DM4D_OK;
}

extern "C" int dm4d_raster_render(const dm4d_raster_desc* d, float* out_color, float* out_depth, float* out_alpha,
                                  void* stream) {
    RasterLayout L;
    int rc = raster_make_layout(d, &L);
    if (rc) return rc;
    if (!out_color || !out_depth || !out_alpha) { dm4d_set_error("NULL output pointer"); return DM4D_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = launch_scatter_sort_pack(L, s))) return rc;
    return launch_render_forward(d, L, out_color, out_depth, out_alpha, s);
}

extern "C" int dm4d_raster_forward(const dm4d_raster_desc* d, float* out_color, float* out_depth, float* out_alpha,
                                   int32_t* radii, void* stream) {
    int rc = dm4d_raster_plan(d, radii, nullptr, stream);
    if (rc) return rc;
    return dm4d_raster_render(d, out_color, out_depth, out_alpha, stream);
}

extern "C" int dm4d_raster_render_features(const dm4d_raster_desc* planned, const dm4d_raster_desc* d, float* out_color,
                                           float* out_depth, float* out_alpha, void* stream) {
    RasterLayout Ls, Ld;
    int rc = raster_make_layout(planned, &Ls);
    if (rc) return rc;
    if ((rc = raster_make_layout(d, &Ld))) return rc;
    if (!out_color || !out_depth || !out_alpha) { dm4d_set_error("NULL output pointer"); return DM4D_EINVAL; }
    if (d->P != planned->P || d->H != planned->H || d->W != planned->W || d->n_views != planned->n_views ||
        d->channels != planned->channels || d->bin_capacity != planned->bin_capacity || d->geom != planned->geom ||
        d->bin == planned->bin || d->img == planned->img) {
        dm4d_set_error("dm4d_raster_render_features: desc must share sizes and `geom` with the planned desc and bring its own bin / img");
        return DM4D_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = launch_rebind_features(d, Ls, Ld, s))) return rc;
    return launch_render_forward(d, Ld, out_color, out_depth, out_alpha, s);
}

extern "C" int dm4d_raster_status(const dm4d_raster_desc* d, int64_t* num_rendered_host, int32_t* overflow_host,
                                  void* stream) {
    RasterLayout L;
    int rc = raster_make_layout(d, &L);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    BinHeader h;
    DM4D_CUDA_CHECK(cudaMemcpyAsync(&h, L.hdr, sizeof(h), cudaMemcpyDeviceToHost, s));
    DM4D_CUDA_CHECK(cudaStreamSynchronize(s));
    if (num_rendered_host) *num_rendered_host = (int64_t)h.total;
    if (overflow_host) *overflow_host = (int32_t)h.overflow;
    return DM4D_OK;
}

extern "C" int dm4d_raster_backward(const dm4d_raster_desc* d, const float* out_color, const float* out_depth,
                                    const float* out_alpha, const float* dL_dcolor,
                                    const float* dL_ddepth, const float* dL_dalpha, float* dL_dmeans3D,
                                    float* dL_dmeans2D, float* dL_dcolors, float* dL_dcolors2, float* dL_dopacities,
                                    float* dL_dscales, float* dL_drotations, void* stream) {
    RasterLayout L;
    int rc = raster_make_layout(d, &L);
    if (rc) return rc;
    uint64_t bwd;
    raster_sizes(d->P, d->H, d->W, d->n_views, d->channels, d->bin_capacity, nullptr, nullptr, nullptr, &bwd);
    if (!d->bwd || d->bwd_bytes < bwd) { dm4d_set_error("bwd workspace too small: need %llu", (unsigned long long)bwd); return DM4D_ENOSPC; }
    if (!out_color || !out_depth || !out_alpha || !dL_dcolor) { dm4d_set_error("out_color / out_depth / out_alpha / dL_dcolor is NULL"); return DM4D_EINVAL; }
    if (d->P == 0) return DM4D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = launch_render_backward(d, L, out_color, out_depth, out_alpha, dL_dcolor, dL_ddepth, dL_dalpha, s))) return rc;
    return launch_preprocess_backward(d, L, dL_dmeans3D, dL_dmeans2D, dL_dcolors, dL_dcolors2, dL_dopacities,
                                      dL_dscales, dL_drotations, s);
}

extern "C" int dm4d_raster_export_state(const dm4d_raster_desc* d, int32_t view, uint32_t* ranges,
                                        uint32_t* point_list, int64_t point_list_capacity, uint32_t* n_contrib,
                                        void* stream) {
    RasterLayout L;
    int rc = raster_make_layout(d, &L);
    if (rc) return rc;
    if (view < 0 || view >= d->n_views) { dm4d_set_error("view out of range"); return DM4D_EINVAL; }
    return launch_export_state(L, view, ranges, point_list, point_list_capacity, n_contrib, (cudaStream_t)stream);
}

// ---- per-kernel event timing ---------------------------------------------------------------------
namespace {
struct Rec { int id; cudaEvent_t e0, e1; };
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<Rec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
const char* const kNames[DM4D_K_COUNT] = {
    "preprocess_kernel", "scan_tiles_kernel", "scatter_kernel", "sort_pack_kernel", "render_forward_kernel",
    "render_backward_kernel", "preprocess_backward_kernel", "skin_vertex_forward_kernel",
    "skin_gaussian_forward_kernel", "skin_gaussian_backward_kernel", "skin_vertex_backward_kernel",
    "sugar_rest_frames_kernel", "arap_energy_kernel", "normal_consistency_kernel", "postops_forward_kernel",
    "postops_backward_kernels", "hexplane_forward_kernel", "hexplane_backward_kernel",
    "knn_kernel", "groupnorm_nhwc_forward_kernels", "groupnorm_nhwc_backward_kernels"};
}  // namespace

KernelTimer::KernelTimer(int id_, cudaStream_t s_) : id(id_), s(s_), on(g_prof_on), e0(nullptr) {
    if (!on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    e0 = prof_event();
    cudaEventRecord(e0, s);
}
KernelTimer::~KernelTimer() {
    if (!on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t e1 = prof_event();
    cudaEventRecord(e1, s);
    g_prof_recs.push_back({id, e0, e1});
}

extern "C" int dm4d_profile_enable(int on) { g_prof_on = on != 0; return DM4D_OK; }
extern "C" const char* dm4d_kernel_name(int id) { return (id >= 0 && id < DM4D_K_COUNT) ? kNames[id] : ""; }
extern "C" int dm4d_profile_collect(double* ms_host, int64_t* launches_host) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof_recs) {
        DM4D_CUDA_CHECK(cudaEventSynchronize(r.e1));
        float ms = 0.f;
        DM4D_CUDA_CHECK(cudaEventElapsedTime(&ms, r.e0, r.e1));
        if (ms_host) ms_host[r.id] += ms;
        if (launches_host) launches_host[r.id] += 1;
        g_prof_pool.push_back(r.e0);
        g_prof_pool.push_back(r.e1);
    }
    g_prof_recs.clear();
    return DM4D_OK;
}
