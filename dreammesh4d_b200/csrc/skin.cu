// placeholder until the fused skinning kernels land (next commit)
#include "raster_internal.cuh"
extern "C" int dm4d_skin_forward(const dm4d_skin_desc*, float*, float*, float*, float*, float*, void*) {
    dm4d_set_error("dm4d_skin_forward: not built yet"); return DM4D_EINVAL; }
extern "C" int dm4d_skin_backward(const dm4d_skin_desc*, const float*, const float*, const float*, const float*,
    const float*, const float*, const float*, float*, float*, float*, float*, float*, float*, void*) {
    dm4d_set_error("dm4d_skin_backward: not built yet"); return DM4D_EINVAL; }
extern "C" int dm4d_sugar_rest_frames(const float*, const int32_t*, const float*, int32_t, int32_t, int32_t, float*, float*, void*) {
    dm4d_set_error("dm4d_sugar_rest_frames: not built yet"); return DM4D_EINVAL; }
