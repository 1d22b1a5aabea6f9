/*
 * dm4d.h — C ABI of libdm4d.so: the B200 (sm_100a) hot path of DreamMesh4D's dynamic stage.
 *
 * Plain pointers and sizes only; every data pointer is a DEVICE pointer unless its name ends
 * in `_host`; `stream` is a cudaStream_t passed as void*.  The caller owns all memory
 * (inputs, outputs and the scratch workspaces, sized by the *_workspace_bytes queries), the
 * library owns none — the same ownership rule as the interface it replaces, where the
 * rasterizer's scratch is resize_()d caller-side torch tensors.
 * Every entry point returns 0 on success or a negative DM4D_E* code and never throws;
 * dm4d_last_error() gives the message (thread-local).
 *
 * Reference interfaces replaced (paths under /root/reference/):
 *   dm4d_raster_*  <- diff_gaussian_rasterization._C.rasterize_gaussians /
 *                     rasterize_gaussians_backward (un-vendored dependency, README.md:35),
 *                     as bound by GaussianRasterizer.forward at
 *                     custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_temporal.py:144,169-178,202-211
 *                     and .../diff_sugar_rasterizer_normal.py:132,161-195
 *   dm4d_skin_*    <- DynamicSuGaRModel._get_timed_vertex_attributes_from_dg
 *                     (custom/threestudio-dreammesh4d/geometry/dynamic_sugar.py:487-613),
 *                     get_timed_gs_attributes (:657-706), _get_gs_xyz_from_vertex (:726-743),
 *                     fuse_rotations (:877-889), get_timed_gs_normals (:357-364)
 *   dm4d_sugar_rest_frames <- SuGaRModel.quaternions / get_gs_normals
 *                     (custom/threestudio-dreammesh4d/geometry/sugar.py:490-526)
 *   dm4d_graph_knn <- DynamicSuGaRModel.build_deformation_graph, mode "eucdisc"
 *                     (custom/threestudio-dreammesh4d/geometry/dynamic_sugar.py:745-790,853-861)
 *   dm4d_hexplane_* <- interpolate_ms_features / HexPlaneField.forward
 *                     (custom/threestudio-dreammesh4d/geometry/deformation.py:141-174,242-248)
 *   dm4d_postops_* <- the image post-ops of DiffGaussian.forward
 *                     (custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_temporal.py:180-239)
 */
#ifndef DM4D_H
#define DM4D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DM4D_OK 0
#define DM4D_EINVAL (-1)   /* bad argument (shape, null pointer, unsupported option) */
#define DM4D_ECUDA (-2)    /* a CUDA runtime call failed */
#define DM4D_ENOSPC (-3)   /* a caller-provided workspace is too small */

#define DM4D_TILE 16            /* tile edge in pixels (BLOCK_X = BLOCK_Y of the replaced rasterizer) */
#define DM4D_MAX_CHANNELS 6     /* 3 (one pass) or 6 (RGB + normal pass fused, binning shared) */

/* Per-view parameter block: DM4D_VIEW_STRIDE floats per view, device memory.
 *   [ 0..15] viewmatrix   — world_view_transform, TRANSPOSED (row-vector) layout exactly as
 *                           the plugin passes it (threestudio/utils/ops.py:400-402)
 *   [16..31] projmatrix   — full_proj_transform, same layout (ops.py:408-410)
 *   [32..34] campos       — camera_center (ops.py:411); used by the SH path only
 *   [35] tanfovx [36] tanfovy [37] scale_modifier
 *   [38] set index (integer stored as float): which attribute set this view renders
 *   [39] reserved
 *   [40..45] bg colour, `channels` entries                                           */
#define DM4D_VIEW_STRIDE 48
#define DM4D_VIEW_TANFOVX 35
#define DM4D_VIEW_TANFOVY 36
#define DM4D_VIEW_SCALE_MOD 37
#define DM4D_VIEW_SET 38
#define DM4D_VIEW_BG 40

/* desc.flags: the caller promises n_sets == n_views and that the view -> set map is a bijection, so
 * per-set gradients can be stored instead of accumulated with atomics. */
#define DM4D_RASTER_VIEWS_DISTINCT_SETS 1

/* A batch of views rasterized in one launch sequence.  Gaussian attributes are organised in
 * "sets" (one per distinct timestamp in the dynamic stage, one in the static stage); each
 * attribute has its own set stride in floats, 0 meaning "shared by every set"
 * (opacity / colour / scale are time-invariant in the dynamic stage:
 * dynamic_sugar.py:720-723).  All arrays are contiguous fp32, row-major. */
typedef struct dm4d_raster_desc {
    int32_t P;                 /* Gaussians per set */
    int32_t H, W;              /* image size (pixels) */
    int32_t n_views;
    int32_t n_sets;
    int32_t channels;          /* 3 or 6 */
    int32_t flags;             /* DM4D_RASTER_* bits */
    int32_t reserved1;
    const float* means3D;   int64_t means3D_stride;     /* [., P, 3] */
    const float* scales;    int64_t scales_stride;      /* [., P, 3] */
    const float* rotations; int64_t rotations_stride;   /* [., P, 4] wxyz, NOT re-normalised */
    const float* opacities; int64_t opacities_stride;   /* [., P, 1] */
    const float* colors;    int64_t colors_stride;      /* [., P, 3] colors_precomp */
    const float* colors2;   int64_t colors2_stride;     /* [., P, 3] channels 3..5 (NULL if channels==3) */
    const float* view_params;                           /* [n_views, DM4D_VIEW_STRIDE] */
    void* geom; uint64_t geom_bytes;                    /* per-(view,Gaussian) projected state */
    void* bin;  uint64_t bin_bytes;                     /* tile / segment tables, keys, sorted instance stream, forward checkpoints */
    void* img;  uint64_t img_bytes;                     /* per-pixel n_contrib */
    void* bwd;  uint64_t bwd_bytes;                     /* backward accumulators (may be NULL for forward) */
    int64_t bin_capacity;                               /* instance capacity the bin workspace was sized for */
    /* cov3D_precomp of the replaced module (optional): the 3D covariances [., P, 6] (xx, xy, xz, yy, yz, zz) instead of
     * scales / rotations, which are then ignored and may be NULL; the scale modifier does not apply to them.  The backward
     * writes dL_dcov3D (same shape, the gradient w.r.t. the six unique entries: off-diagonal terms counted twice, as the
     * replaced rasterizer does) when the pointer is non-NULL. */
    const float* cov3D;     int64_t cov3D_stride;
    float* dL_dcov3D;
} dm4d_raster_desc;

/* Workspace sizes for a batch. `bin_capacity` = max number of (Gaussian, tile) instances. */
int dm4d_raster_workspace_bytes(int32_t P, int32_t H, int32_t W, int32_t n_views, int32_t channels,
                                int64_t bin_capacity, uint64_t* geom_bytes, uint64_t* bin_bytes,
                                uint64_t* img_bytes, uint64_t* bwd_bytes);

/* Phase A of the forward: project every (view, Gaussian), write radii [n_views, P] (int32), count
 * instances per tile and prefix-sum them.  If `num_rendered_host` is non-NULL the stream is
 * synchronised and the total instance count R is returned (the replaced rasterizer always does
 * this read-back); pass NULL to stay asynchronous and size `bin` from a capacity bound instead. */
int dm4d_raster_plan(const dm4d_raster_desc* d, int32_t* radii, int64_t* num_rendered_host, void* stream);

/* Phase B: scatter instances to their tiles, depth-sort every tile, pack the sorted instance
 * stream and alpha-composite.  Outputs: color [n_views, channels, H, W], depth [n_views,1,H,W],
 * alpha [n_views,1,H,W].  If R exceeded bin_capacity nothing is rendered and the device-side
 * overflow flag is set (see dm4d_raster_status). */
int dm4d_raster_render(const dm4d_raster_desc* d, float* out_color, float* out_depth, float* out_alpha,
                       void* stream);

/* plan (asynchronous) + render. */
int dm4d_raster_forward(const dm4d_raster_desc* d, float* out_color, float* out_depth, float* out_alpha,
                        int32_t* radii, void* stream);

/* Second pass over an already planned batch with OTHER per-Gaussian features (the reference's renderer calls the
 * rasterizer twice per view with identical means / scales / rotations / opacities — RGB, then normals as colours,
 * diff_sugar_rasterizer_temporal.py:169-178,202-211): re-uses projection, binning and the depth sort of `planned`
 * (a desc on which dm4d_raster_forward ran), re-binds `d->colors` (/colors2) onto the sorted instance stream and
 * composites.  `d` must equal `planned` in sizes, channels, bin_capacity and share its `geom` workspace; `d->bin`
 * and `d->img` are its OWN workspaces (same sizes), so both passes keep what their backward needs.
 * dm4d_raster_backward(d, ...) then works as after dm4d_raster_forward. */
int dm4d_raster_render_features(const dm4d_raster_desc* planned, const dm4d_raster_desc* d, float* out_color,
                                float* out_depth, float* out_alpha, void* stream);

/* Status header: the first 16 bytes of the `bin` workspace are
 *     struct { uint64_t num_rendered; uint32_t overflow; uint32_t n_segments; }
 * written by the plan phase on the device.  A caller that must not synchronise (CUDA-graph replay, a training loop)
 * reads or accumulates the flag with its own device-side ops and polls it every N steps; dm4d_raster_status is the
 * synchronous convenience form.  When `overflow` is 1 the render and backward kernels treat every tile as empty
 * (background images, zero gradients) — the caller must re-run with a larger capacity. */
/* Synchronises `stream` and reports the instance count and overflow flag of the last plan. */
int dm4d_raster_status(const dm4d_raster_desc* d, int64_t* num_rendered_host, int32_t* overflow_host,
                       void* stream);

/* Backward of dm4d_raster_forward (or dm4d_raster_render_features) for the same desc/workspaces.  `out_color`,
 * `out_depth`, `out_alpha` are the forward's output images (the replaced rasterizer saves alpha the same way; the
 * colour and depth images let each pixel start from the total of its composited sum, so the backward runs front to
 * back in independent segments — see csrc/raster_render.cu).  dL_ddepth / dL_dalpha may be NULL (zero).  Gradient
 * outputs have the shape of the matching input ([n_sets or 1, P, k]) and are fully overwritten (summed over the views
 * that used each set entry); dL_dmeans2D is [n_views, P, 3] (z = 0, NDC-scaled as in the replaced rasterizer).
 * Any output pointer may be NULL to skip it. */
int dm4d_raster_backward(const dm4d_raster_desc* d, const float* out_color, const float* out_depth, const float* out_alpha,
                         const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha, float* dL_dmeans3D,
                         float* dL_dmeans2D, float* dL_dcolors, float* dL_dcolors2, float* dL_dopacities,
                         float* dL_dscales, float* dL_drotations, void* stream);

/* Test/inspection helper: copies the integer binning state of one view to device arrays:
 * ranges [tiles, 2] (uint32, relative to the view's first instance), point_list [R_view] (uint32
 * Gaussian ids in sorted order), n_contrib [H, W] (uint32). Any pointer may be NULL. */
int dm4d_raster_export_state(const dm4d_raster_desc* d, int32_t view, uint32_t* ranges, uint32_t* point_list,
                             int64_t point_list_capacity, uint32_t* n_contrib, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sparse-control-point skinning + per-face surface-bound Gaussian update (fused).
 * n_t timestamps are deformed in one launch sequence.
 * ---------------------------------------------------------------------------------------------- */
typedef struct dm4d_skin_desc {
    int32_t n_t;               /* timestamps in this batch */
    int32_t V, F, M, K;        /* vertices, faces, control nodes, neighbours per vertex */
    int32_t g;                 /* Gaussians per face (1, 3, 4 or 6); P = F * g */
    int32_t method;            /* 0 = lbs, 1 = dqs, 2 = hybrid (dynamic_sugar.py:72) */
    int32_t reserved0;
    const float* rest_verts;   /* [V, 3]  SuGaRModel._points */
    const int32_t* faces;      /* [F, 3]  _surface_mesh_faces (int32) */
    const int32_t* nbr_idx;    /* [V, K]  _xyz_neighbor_node_idx (int32) */
    const float* nbr_w;        /* [V, K]  _xyz_neighbor_nodes_weights (row-normalised) */
    const float* bary;         /* [g, 3]  surface_triangle_bary_coords */
    const float* rest_quat;    /* [P, 4]  wxyz rest-pose quaternion (SuGaRModel.quaternions) */
    const float* node_trans;   /* [n_t, M, 3] */
    const float* node_rot;     /* [n_t, M, 4] xyzw, unit */
    const float* node_scale;   /* [n_t, M, 9] row-major 3x3 (I + strain) */
    const float* node_opacity; /* [n_t, M]    sigmoid'ed lbs weight */
    /* Optional (backward only; all three or none): the control nodes' incidence lists from dm4d_skin_node_incidence
     * and a scratch buffer.  With them the vertex backward runs node-centric: upstream gradients once per vertex into
     * vert_scratch, then one CTA per (node, timestamp) gathers its incidences (no atomics on the node tables,
     * bit-reproducible when M * n_t >= 1184; 55 us at C5).  Without them every warp sums the lanes that hit the same
     * node with shuffles and issues one floating-point reduction per (warp, distinct node, component) (66 us at C5). */
    const int32_t* node_inc_ptr; /* [M+1] */
    const int32_t* node_inc;     /* [V*K] flat (vertex, slot) indices e = v*K + k, grouped by node, ascending */
    float* vert_scratch;         /* [n_t, V, 16] caller-owned scratch, 16-byte aligned (overwritten) */
    /* Optional (backward only; all three or none): the vertices' incidence lists over the face corners
     * (dm4d_skin_node_incidence(faces, F, 3, V, ...): flat corner indices e = f*3 + c grouped by vertex, ascending) and a
     * scratch buffer.  With them the Gaussian stage writes one 32-byte gradient record per (timestamp, face corner) and the
     * vertex stage gathers them: no floating-point reductions into the per-vertex rows, reproducible summation order
     * (with the node lists above the whole backward is bit-reproducible; ~10 % slower at C5).  Without them: 16-byte
     * vector reductions (the default of the Python host side). */
    const int32_t* vert_inc_ptr; /* [V+1] */
    const int32_t* vert_inc;     /* [F*3] */
    float* corner_scratch;       /* [n_t, F*3, 8] caller-owned scratch, 16-byte aligned (overwritten) */
    /* Optional (forward and backward): [n_t, M, 12] scratch, 16-byte aligned (overwritten by every call).  With it the
     * per-node quantities every (vertex, neighbour) pair needs (rotation log, normalised quaternion, dual part) are
     * evaluated once per (timestamp, node) in a pre-pass instead of once per pair. */
    float* node_scratch;
} dm4d_skin_desc;

/* Incidence lists (start-up; the deformation graph of dynamic_sugar.py:745-861 and the mesh are fixed afterwards): for an
 * index table nbr_idx [V, K] with values in [0, M) — the control nodes of every vertex, or the three vertices of every
 * face (V := F, K := 3, M := number of vertices) — the flat positions e = v*K + k grouped by value, ascending:
 * inc_ptr [M+1], inc [V*K].  scratch: [M+1] int32; scratch[M] != 0 afterwards means that nbr_idx
 * held an index outside [0, M) (the lists are then incomplete). */
int dm4d_skin_node_incidence(const int32_t* nbr_idx, int32_t V, int32_t K, int32_t M, int32_t* inc_ptr, int32_t* inc,
                             int32_t* scratch, void* stream);

/* Outputs: verts [n_t,V,3], vert_rot [n_t,V,4] xyzw, means3D [n_t,P,3], rotations [n_t,P,4] wxyz
 * (normalised), normals [n_t,P,3] (deformed unit face normal repeated g times; may be NULL). */
int dm4d_skin_forward(const dm4d_skin_desc* d, float* verts, float* vert_rot, float* means3D,
                      float* rotations, float* normals, void* stream);

/* Backward: incoming dL_dmeans3D [n_t,P,3], dL_drotations [n_t,P,4], dL_dnormals [n_t,P,3] (any may
 * be NULL), plus optional direct gradients on the deformed vertices dL_dverts_in [n_t,V,3] and vertex
 * rotations dL_dvert_rot_in [n_t,V,4] (ARAP / mesh regularisers).  `verts`/`vert_rot` are the forward
 * outputs.  Scratch: dverts [n_t,V,4] (xyz + pad) and dvert_rot [n_t,V,4] (caller-owned, 16-byte aligned, overwritten).
 * Outputs (overwritten): dL_dnode_trans [n_t,M,3], dL_dnode_rot [n_t,M,4], dL_dnode_scale [n_t,M,9],
 * dL_dnode_opacity [n_t,M].  Exact Euclidean gradients of the forward formulas. */
int dm4d_skin_backward(const dm4d_skin_desc* d, const float* verts, const float* vert_rot,
                       const float* dL_dmeans3D, const float* dL_drotations, const float* dL_dnormals,
                       const float* dL_dverts_in, const float* dL_dvert_rot_in, float* dverts, float* dvert_rot,
                       float* dL_dnode_trans, float* dL_dnode_rot, float* dL_dnode_scale,
                       float* dL_dnode_opacity, void* stream);

/* Rest-pose frames of the surface-bound Gaussians: quaternions [P,4] wxyz (normalised) and unit
 * face normals repeated g times [P,3] (either may be NULL). complex_rot is SuGaRModel._quaternions [P,2]. */
int dm4d_sugar_rest_frames(const float* verts, const int32_t* faces, const float* complex_rot, int32_t V,
                           int32_t F, int32_t g, float* quaternions, float* normals, void* stream);

/* Backward of dm4d_sugar_rest_frames (static stage: vertices and in-plane rotations are learnable,
 * sugar.py:333-376).  dL_dquaternions [P,4] / dL_dnormals [P,3] may be NULL.  Outputs (overwritten):
 * dL_dverts [V,3], dL_dcomplex_rot [P,2] (required when dL_dquaternions is given). */
int dm4d_sugar_rest_frames_backward(const float* verts, const int32_t* faces, const float* complex_rot, int32_t V,
                                    int32_t F, int32_t g, const float* dL_dquaternions, const float* dL_dnormals,
                                    float* dL_dverts, float* dL_dcomplex_rot, void* stream);

/* ARAP energy of the deformed mesh with the rotations supplied by the deformation, per timestamp:
 *   E_t = sum_i sum_{j in N(i)} w_ij || (x'_i - x'_j) - R(q_i) (x_i - x_j) ||^2
 * Replaces ARAPCoach.compute_arap_energy(xyz_prime, vert_rotations)
 * (custom/threestudio-dreammesh4d/utils/arap_utils.py:183-224; caller system/sugar_4dgen.py:372-385).
 * One-ring as CSR: row_ptr [V+1], col [E] (int32), weights [E] (the cotangent weights of arap_utils.py:100-175).
 * verts [n_t,V,3] / vert_rot [n_t,V,4] xyzw are dm4d_skin_forward's outputs.  Outputs (overwritten): energy [n_t],
 * and, if non-NULL, the gradients dE/dverts [n_t,V,3] and dE/dvert_rot [n_t,V,4], which are exactly
 * dm4d_skin_backward's dL_dverts_in / dL_dvert_rot_in (after scaling by the loss weight). */
int dm4d_arap_energy(const float* rest_verts, const int32_t* row_ptr, const int32_t* col, const float* weights,
                     int32_t n_t, int32_t V, const float* verts, const float* vert_rot, float* energy,
                     float* dE_dverts, float* dE_dvert_rot, void* stream);

/* Mesh normal consistency of the deformed meshes, per timestamp (pytorch3d.loss.mesh_normal_consistency as called at
 * custom/threestudio-dreammesh4d/system/sugar_4dgen.py:214-225): pairs [n_pairs,4] int32 = (v0, v1, a, b) for every
 * pair of faces sharing edge (v0,v1) with opposite vertices a and b.  Outputs (overwritten): loss [n_t] (mean over
 * pairs) and, if non-NULL, d loss_t / d verts [n_t,V,3]. */
int dm4d_mesh_normal_consistency(const int32_t* pairs, int32_t n_pairs, int32_t n_t, int32_t V, const float* verts,
                                 float* loss, float* dL_dverts, void* stream);

/* ---- fused per-view image post-ops --------------------------------------------------
 * Replaces the element-wise / convolution / boolean-index tail of DiffGaussian.forward
 * (custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_temporal.py:180-193,212-218,229; Depth2Normal
 * :25-54; static twin diff_sugar_rasterizer_normal.py:172-206) and the stack/permute of
 * GaussianBatchRenderer.batch_forward (renderer/gaussian_batch_renderer.py:78-122).
 * Inputs are the rasterizer's planar outputs of a 6-channel pass (rgb + rendered normals) plus the batch's rays;
 * outputs are [n_views,H,W,C] (channel-last), exactly the comp_* tensors of the renderer's return dict. */
#define DM4D_POSTOPS_NORMAL_FROM_DIST 1   /* compute comp_normal_from_dist (needs rays_o / rays_d) */
#define DM4D_POSTOPS_STATIC 2             /* static renderer: the depth is detached outside the mask AFTER the position
                                             map was built, so the stencil gradient reaches unmasked pixels */
typedef struct dm4d_postops_desc {
    int32_t n_views, H, W, flags;
    const float* color6;   /* [n_views,6,H,W] */
    const float* depth;    /* [n_views,1,H,W] */
    const float* alpha;    /* [n_views,1,H,W] */
    const float* rays_o;   /* [n_views,H,W,3] or NULL */
    const float* rays_d;   /* [n_views,H,W,3] or NULL */
} dm4d_postops_desc;

/* comp_rgb, comp_normal, comp_normal_from_dist (NULL iff the flag is clear): [n_views,H,W,3];
 * comp_depth, comp_mask: [n_views,H,W,1]. */
int dm4d_postops_forward(const dm4d_postops_desc* d, float* comp_rgb, float* comp_normal,
                         float* comp_normal_from_dist, float* comp_depth, float* comp_mask, void* stream);
/* g_*: gradients w.r.t. the five outputs (any may be NULL = zero).  scratch: [n_views,H,W,6] floats, required with
 * DM4D_POSTOPS_NORMAL_FROM_DIST.  Outputs (overwritten): d_color6 [n_views,6,H,W], d_depth, d_alpha [n_views,1,H,W]
 * — the dL_dcolor / dL_ddepth / dL_dalpha inputs of dm4d_raster_backward.  No atomics: bit-reproducible. */
int dm4d_postops_backward(const dm4d_postops_desc* d, const float* g_rgb, const float* g_normal,
                          const float* g_normal_from_dist, const float* g_depth, const float* g_mask,
                          float* scratch, float* d_color6, float* d_depth, float* d_alpha, void* stream);

/* ---- fused HexPlane multi-scale feature lookup ------------------------------------
 * Replaces interpolate_ms_features (custom/threestudio-dreammesh4d/geometry/deformation.py:141-174; grid_sample_wrapper
 * :84-111 = bilinear, padding_mode='border', align_corners=True) as called by HexPlaneField.forward (:242-248):
 *   features[n, s*feat + c] = prod over the 6 planes (i,j) in combinations(range(4),2) of
 *                             bilinear(planes[s][p][c], coords[n,i], coords[n,j])
 * coords [n_points,4]: normalised (x,y,z,t) (after normalize_aabb, :80-81).  planes[s][p]: the reference's parameter
 * tensor of scale s, plane p, layout [1,feat,res[s][j],res[s][i]] (checkpoint-compatible).  res[s] = (x,y,z,t) sizes. */
#define DM4D_HEX_MAX_SCALES 8
typedef struct dm4d_hexplane_desc {
    int32_t n_points, n_scales, feat, reserved;
    const float* coords;
    const float* planes[DM4D_HEX_MAX_SCALES][6];
    int32_t res[DM4D_HEX_MAX_SCALES][4];
} dm4d_hexplane_desc;
/* features [n_points, n_scales*feat] (overwritten). */
int dm4d_hexplane_forward(const dm4d_hexplane_desc* d, float* features, void* stream);
/* dL_dplanes_host: HOST array of n_scales*6 DEVICE pointers (index s*6+p; NULL = skip that plane) to dense gradient
 * tensors of the planes' shapes, ZEROED by the caller; the kernel adds into the touched texels (RED.ADD.F32).
 * Coordinates receive no gradient (the control nodes and the timestamps are not trainable in the reference). */
int dm4d_hexplane_backward(const dm4d_hexplane_desc* d, const float* dL_dfeatures, float* const* dL_dplanes_host,
                           void* stream);

/* ---- deformation-graph construction ---------------------------------------------------
 * K nearest `nodes` [n_nodes,3] of every query point [n_queries,3], Euclidean: replaces the per-vertex Open3D KD-tree
 * queries of DynamicSuGaRModel.build_deformation_graph, mode "eucdisc"
 * (custom/threestudio-dreammesh4d/geometry/dynamic_sugar.py:765-790).  Outputs (overwritten): idx [n_queries,k] node
 * indices ordered by ascending (squared distance, index), sqdist [n_queries,k] (may be NULL) the SQUARED distances —
 * exactly the [1] and [2] results of KDTreeFlann.search_knn_vector_3d.  1 <= k <= min(17, n_nodes). */
int dm4d_graph_knn(const float* queries, int32_t n_queries, const float* nodes, int32_t n_nodes, int32_t k,
                   int32_t* idx, float* sqdist, void* stream);

/* One Jacobi sweep of the geodesic K-nearest-node propagation (deformation-graph mode "geodisc", replaces the per-vertex
 * heat-method solves of custom/threestudio-dreammesh4d/geometry/dynamic_sugar.py:791-849).  Mesh as CSR over vertices
 * (row_ptr [V+1], col [E], edge_len [E]); labels dist/node [V,k] ascending by (distance, node), node = -1 / dist = +inf
 * for empty slots.  Initialise the labels of every node's nearest vertex with (0, node); call with swapped in/out
 * buffers until `*changed` (device int32, zeroed by the caller before each sweep) stays 0.  1 <= k <= 17. */
int dm4d_graph_geodesic_sweep(int32_t V, int32_t k, const int32_t* row_ptr, const int32_t* col, const float* edge_len,
                              const float* dist_in, const int32_t* node_in, float* dist_out, int32_t* node_out,
                              int32_t* changed, void* stream);

/* ---- channels-last GroupNorm (+ channel bias in front, + SiLU behind) for the Zero123 networks of the SDS step --------
 * (SURVEY.md §8 row A9).  Replaces, between the tensor-core convolutions that stay library calls, the
 * ``normalization(ch) -> SiLU`` pairs of ResBlock (extern/ldm_zero123/modules/diffusionmodules/openaimodel.py:210-214,
 * 243-250,269-289; the time-embedding add ``h + emb_out`` :286 is the channel bias) and of ResnetBlock / Encoder
 * (.../diffusionmodules/model.py:118-138,489-492), and the plain GroupNorm of SpatialTransformer / AttnBlock
 * (modules/attention.py:275, model.py:169):
 *     y[n,p,c] = act(((x[n,p,c] + chan_bias[n,c]) - mean[n,g]) * rstd[n,g] * gamma[c] + beta[c]),  g = c / (C/G)
 * x, y: [N, HW, C] = the memory of a torch channels_last [N,C,H,W] tensor, `dtype` DM4D_F16 or DM4D_F32.
 * chan_bias [N,C] fp32 or NULL; gamma, beta [C] fp32; stats [N,G,2] fp32 (mean, rstd; output of the forward, input of
 * the backward); scratch [N*G*2 + N + 1] 4-byte words (group sums + per-sample arrival counters).  C % G == 0, C % 4 == 0,
 * C <= 4096, G <= 64.  Two streamed launches per call (statistics, apply), or one
 * register-resident launch for small fp16 activations. */
#define DM4D_F32 0
#define DM4D_F16 1
int dm4d_groupnorm_nhwc_forward(const void* x, const float* chan_bias, const float* gamma, const float* beta,
                                int32_t N, int32_t HW, int32_t C, int32_t G, float eps, int32_t silu, int32_t dtype,
                                float* stats, float* scratch, void* y, void* stream);
/* dx (same layout / dtype as x) from dy; gamma / beta / chan_bias receive no gradient (frozen in the SDS step). */
int dm4d_groupnorm_nhwc_backward(const void* x, const float* chan_bias, const void* dy, const float* gamma,
                                 const float* beta, int32_t N, int32_t HW, int32_t C, int32_t G, float eps, int32_t silu,
                                 int32_t dtype, const float* stats, float* scratch, void* dx, void* stream);

/* Epilogues the library convolutions / matmuls of the Zero123 networks do not fuse (channels-last, DM4D_F16 / DM4D_F32):
 *   dm4d_bias_residual_add_nhwc: out[m,c] = h[m,c] + bias[c] (+ residual[m,c]) — the convolution bias folded into the
 *     ResBlock's residual add (openaimodel.py:289, model.py:138); h, residual (or NULL), out: [M, C], bias [C] fp32, C % 4 == 0;
 *   dm4d_geglu: out[m,d] = proj[m,d] * gelu(proj[m,D+d]) with the exact GELU — the gated feed-forward of the transformer
 *     blocks (modules/attention.py:37-65); proj [M, 2D], out [M, D], D % 4 == 0. */
int dm4d_bias_residual_add_nhwc(const void* h, const void* residual, const float* bias, int64_t M, int32_t C,
                                int32_t dtype, void* out, void* stream);
int dm4d_geglu(const void* proj, int64_t M, int32_t D, int32_t dtype, void* out, void* stream);
/* Residual add + LayerNorm of the transformer blocks (modules/attention.py:199-246): x_out[m,:] = x[m,:] + delta[r,:] with
 * r = m (delta_bcast_rows == 0) or m / delta_bcast_rows (one delta row shared by that many consecutive rows: the
 * single-token cross-attention output); y[m,:] = LayerNorm(x_out[m,:]) * gamma + beta.  delta / x_out may be NULL (plain
 * LayerNorm of x).  x, delta, x_out, y: [M, C] in `dtype`; gamma, beta [C] fp32; C % 4 == 0, C <= 1536.  Forward only
 * (the UNet of the SDS step runs without gradient). */
int dm4d_add_layernorm(const void* x, const void* delta, int32_t delta_bcast_rows, const float* gamma, const float* beta,
                       int64_t M, int32_t C, float eps, int32_t dtype, void* x_out, void* y, void* stream);

/* Per-kernel device timing (CUDA events recorded on the launch stream around every kernel launch).
 * Kernel ids: see DM4D_K_* below.  dm4d_profile_collect synchronises the recorded events, ADDS the
 * elapsed milliseconds / launch counts since the last collect into ms[DM4D_K_COUNT] /
 * launches[DM4D_K_COUNT] (host arrays) and clears the record. */
enum {
    DM4D_K_PREPROCESS = 0, DM4D_K_SCAN, DM4D_K_SCATTER, DM4D_K_SORT_PACK, DM4D_K_RENDER_FWD,
    DM4D_K_RENDER_BWD, DM4D_K_PREPROCESS_BWD, DM4D_K_SKIN_VERT_FWD, DM4D_K_SKIN_GAUSS_FWD,
    DM4D_K_SKIN_GAUSS_BWD, DM4D_K_SKIN_VERT_BWD, DM4D_K_REST_FRAMES, DM4D_K_ARAP, DM4D_K_NORMAL_CONS,
    DM4D_K_POSTOPS_FWD, DM4D_K_POSTOPS_BWD, DM4D_K_HEXPLANE_FWD, DM4D_K_HEXPLANE_BWD,
    DM4D_K_GRAPH_KNN, DM4D_K_GROUPNORM_FWD, DM4D_K_GROUPNORM_BWD, DM4D_K_COUNT
};
int dm4d_profile_enable(int on);
int dm4d_profile_collect(double* ms_host, int64_t* launches_host);
const char* dm4d_kernel_name(int id);

const char* dm4d_last_error(void);
int dm4d_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DM4D_H */
