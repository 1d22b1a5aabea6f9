/*
 * raster_oracle.c — CPU ORACLE for the tile-based differentiable Gaussian rasterizer.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under dreammesh4d_b200/ may include, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, and only as the checker / reported CPU baseline.
 *
 * PARITY UNPINNED.  The algorithm lives in an un-vendored, un-pinned third-party
 * dependency of the reference: `diff-gaussian-rasterization` (ashawkey fork, returns
 * colour, radii, depth, alpha; /root/reference/README.md:35, requirements.txt:49).
 * Its sources are not under /root/reference and the reference holds no tests, golden
 * vectors or fixtures for this path.  This file restates the
 * published algorithm as transcribed in SURVEY.md Appendix A and is anchored on the
 * reference's own call sites:
 *   custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_temporal.py:129-144
 *       (GaussianRasterizationSettings fields), :169-178 and :202-211 (the two calls,
 *       4-tuple return colour/radii/depth/alpha)
 *   custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_normal.py:117-132,161-195
 *   threestudio/utils/ops.py:398-413 (matrix conventions: transposed / row-vector)
 *
 * Arithmetic spec (shared with the CUDA path so that the integer state — radii, tile
 * rectangles, sorted instance lists, ranges — is bit-exact): IEEE-754 binary32,
 * round-to-nearest, NO fused multiply-add contraction (compile with
 * -ffp-contract=off), sums evaluated left to right exactly as written below.
 * Build with -DORACLE_FP64 for a binary64 master copy of the same formulas.
 *
 * Section references "A.2 step n" / "A.3" are to SURVEY.md Appendix A.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORACLE_FP64
typedef double real;
#define R_SQRT sqrt
#define R_EXP exp
#define R_CEIL ceil
#else
typedef float real;
#define R_SQRT sqrtf
#define R_EXP expf
#define R_CEIL ceilf
#endif

#define BLOCK_X 16
#define BLOCK_Y 16
#define MAXC 8

typedef struct {
    uint32_t tile;
    real depth;
    uint32_t id;
} inst_t;

typedef struct oracle_raster {
    int P, H, W, C, gx, gy;
    real tanfovx, tanfovy, focal_x, focal_y, scale_mod;
    real view[16], proj[16], bg[MAXC];
    /* inputs (copied) */
    real *means, *scales, *rots, *opac, *colors;
    /* per-Gaussian state (A.2 step 9) */
    real *depth, *xy, *conic_op, *cov3d, *cov2d;
    int32_t *radii, *rect;        /* rect = (min.x, min.y, max.x, max.y) in tiles */
    uint32_t *tiles_touched;
    uint8_t *clampxy;             /* bit0: x clamp active, bit1: y clamp active (A.2 step 4) */
    /* binning */
    int64_t R;
    inst_t *list;                 /* sorted instances */
    uint32_t *ranges;             /* [T][2] */
    /* image state */
    real *out_color, *out_depth, *out_alpha;
    uint32_t *n_contrib;
    uint8_t *ambig;               /* per-pixel threshold-ambiguity flag */
    /* backward accumulators: double, so the oracle's own summation order is irrelevant */
    double *g_mean2d, *g_conic, *g_opac, *g_color, *g_depth;
    real amb_rel;                 /* relative decision margin for the ambiguity flag */
} oracle_raster;

static void *xcalloc(size_t n, size_t s) { void *p = calloc(n ? n : 1, s); if (!p) abort(); return p; }

oracle_raster *oracle_raster_create(int P, int H, int W, int C) {
    if (C < 1 || C > MAXC) return NULL;
    oracle_raster *o = (oracle_raster *)xcalloc(1, sizeof(*o));
    o->P = P; o->H = H; o->W = W; o->C = C;
    o->gx = (W + BLOCK_X - 1) / BLOCK_X;
    o->gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    o->means = xcalloc((size_t)P * 3, sizeof(real));
    o->scales = xcalloc((size_t)P * 3, sizeof(real));
    o->rots = xcalloc((size_t)P * 4, sizeof(real));
    o->opac = xcalloc((size_t)P, sizeof(real));
    o->colors = xcalloc((size_t)P * C, sizeof(real));
    o->depth = xcalloc((size_t)P, sizeof(real));
    o->xy = xcalloc((size_t)P * 2, sizeof(real));
    o->conic_op = xcalloc((size_t)P * 4, sizeof(real));
    o->cov3d = xcalloc((size_t)P * 6, sizeof(real));
    o->cov2d = xcalloc((size_t)P * 3, sizeof(real));
    o->radii = xcalloc((size_t)P, sizeof(int32_t));
    o->rect = xcalloc((size_t)P * 4, sizeof(int32_t));
    o->tiles_touched = xcalloc((size_t)P, sizeof(uint32_t));
    o->clampxy = xcalloc((size_t)P, 1);
    o->ranges = xcalloc((size_t)o->gx * o->gy * 2, sizeof(uint32_t));
    o->out_color = xcalloc((size_t)C * H * W, sizeof(real));
    o->out_depth = xcalloc((size_t)H * W, sizeof(real));
    o->out_alpha = xcalloc((size_t)H * W, sizeof(real));
    o->n_contrib = xcalloc((size_t)H * W, sizeof(uint32_t));
    o->ambig = xcalloc((size_t)H * W, 1);
    o->g_mean2d = xcalloc((size_t)P * 2, sizeof(double));
    o->g_conic = xcalloc((size_t)P * 3, sizeof(double));
    o->g_opac = xcalloc((size_t)P, sizeof(double));
    o->g_color = xcalloc((size_t)P * C, sizeof(double));
    o->g_depth = xcalloc((size_t)P, sizeof(double));
    o->amb_rel = (real)2e-5;
    return o;
}

void oracle_raster_destroy(oracle_raster *o) {
    if (!o) return;
    free(o->means); free(o->scales); free(o->rots); free(o->opac); free(o->colors);
    free(o->depth); free(o->xy); free(o->conic_op); free(o->cov3d); free(o->cov2d);
    free(o->radii); free(o->rect); free(o->tiles_touched); free(o->clampxy);
    free(o->list); free(o->ranges);
    free(o->out_color); free(o->out_depth); free(o->out_alpha); free(o->n_contrib); free(o->ambig);
    free(o->g_mean2d); free(o->g_conic); free(o->g_opac); free(o->g_color); free(o->g_depth);
    free(o);
}

void oracle_raster_set_ambiguity_margin(oracle_raster *o, double rel) { o->amb_rel = (real)rel; }

/* ------------------------------------------------------------------------------------------ */
/* A.2 preprocess                                                                               */
/* ------------------------------------------------------------------------------------------ */

static inline real ndc2pix(real v, int S) { return ((v + (real)1.0) * (real)S - (real)1.0) * (real)0.5; }

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline real rmin(real a, real b) { return a < b ? a : b; }
static inline real rmax(real a, real b) { return a > b ? a : b; }

/* rotation matrix of the UN-normalised quaternion (r,x,y,z), row-major (A.2 step 3) */
static inline void quat_to_R(const real *q, real Rm[3][3]) {
    const real r = q[0], x = q[1], y = q[2], z = q[3];
    Rm[0][0] = (real)1 - (real)2 * (y * y + z * z); Rm[0][1] = (real)2 * (x * y - r * z); Rm[0][2] = (real)2 * (x * z + r * y);
    Rm[1][0] = (real)2 * (x * y + r * z); Rm[1][1] = (real)1 - (real)2 * (x * x + z * z); Rm[1][2] = (real)2 * (y * z - r * x);
    Rm[2][0] = (real)2 * (x * z - r * y); Rm[2][1] = (real)2 * (y * z + r * x); Rm[2][2] = (real)1 - (real)2 * (x * x + y * y);
}

static void preprocess_one(oracle_raster *o, int i) {
    const real *V = o->view, *PV = o->proj;
    const real px = o->means[3 * i], py = o->means[3 * i + 1], pz = o->means[3 * i + 2];
    o->radii[i] = 0; o->tiles_touched[i] = 0; o->clampxy[i] = 0;
    o->rect[4 * i] = o->rect[4 * i + 1] = o->rect[4 * i + 2] = o->rect[4 * i + 3] = 0;

    /* step 1: view-space point; near cull */
    real tx = V[0] * px + V[4] * py + V[8] * pz + V[12];
    real ty = V[1] * px + V[5] * py + V[9] * pz + V[13];
    real tz = V[2] * px + V[6] * py + V[10] * pz + V[14];
    if (tz <= (real)0.2) return;

    /* step 2: clip space, perspective divide */
    real hx = PV[0] * px + PV[4] * py + PV[8] * pz + PV[12];
    real hy = PV[1] * px + PV[5] * py + PV[9] * pz + PV[13];
    real hw = PV[3] * px + PV[7] * py + PV[11] * pz + PV[15];
    real p_w = (real)1.0 / (hw + (real)0.0000001);
    real projx = hx * p_w, projy = hy * p_w;

    /* step 3: 3D covariance  Sigma = Rm diag(s)^2 Rm^T  via L = Rm diag(s) */
    real Rm[3][3], L[3][3];
    quat_to_R(&o->rots[4 * i], Rm);
    real s[3] = { o->scale_mod * o->scales[3 * i], o->scale_mod * o->scales[3 * i + 1], o->scale_mod * o->scales[3 * i + 2] };
    for (int a = 0; a < 3; ++a) for (int k = 0; k < 3; ++k) L[a][k] = Rm[a][k] * s[k];
    real S[3][3];
    for (int a = 0; a < 3; ++a) for (int b = a; b < 3; ++b) {
        S[a][b] = L[a][0] * L[b][0] + L[a][1] * L[b][1] + L[a][2] * L[b][2];
        S[b][a] = S[a][b];
    }
    real *c3 = &o->cov3d[6 * i];
    c3[0] = S[0][0]; c3[1] = S[0][1]; c3[2] = S[0][2]; c3[3] = S[1][1]; c3[4] = S[1][2]; c3[5] = S[2][2];

    /* step 4: EWA 2D covariance */
    const real limx = (real)1.3 * o->tanfovx, limy = (real)1.3 * o->tanfovy;
    const real txtz = tx / tz, tytz = ty / tz;
    uint8_t cl = 0;
    if (txtz < -limx || txtz > limx) cl |= 1;
    if (tytz < -limy || tytz > limy) cl |= 2;
    o->clampxy[i] = cl;
    real cx = rmin(limx, rmax(-limx, txtz)) * tz;
    real cy = rmin(limy, rmax(-limy, tytz)) * tz;
    real J00 = o->focal_x / tz, J02 = -(o->focal_x * cx) / (tz * tz);
    real J11 = o->focal_y / tz, J12 = -(o->focal_y * cy) / (tz * tz);
    /* W3[i][j] = V[i + 4 j];  A = J W3 (2x3) */
    real A[2][3], B[2][3];
    for (int j = 0; j < 3; ++j) {
        A[0][j] = J00 * V[0 + 4 * j] + J02 * V[2 + 4 * j];
        A[1][j] = J11 * V[1 + 4 * j] + J12 * V[2 + 4 * j];
    }
    for (int r = 0; r < 2; ++r) for (int j = 0; j < 3; ++j)
        B[r][j] = A[r][0] * S[0][j] + A[r][1] * S[1][j] + A[r][2] * S[2][j];
    real a = B[0][0] * A[0][0] + B[0][1] * A[0][1] + B[0][2] * A[0][2];
    real b = B[0][0] * A[1][0] + B[0][1] * A[1][1] + B[0][2] * A[1][2];
    real c = B[1][0] * A[1][0] + B[1][1] * A[1][1] + B[1][2] * A[1][2];
    a += (real)0.3; c += (real)0.3;
    o->cov2d[3 * i] = a; o->cov2d[3 * i + 1] = b; o->cov2d[3 * i + 2] = c;

    /* step 5: conic */
    real det = a * c - b * b;
    if (det == (real)0.0) return;
    real det_inv = (real)1.0 / det;
    real conx = c * det_inv, cony = -b * det_inv, conz = a * det_inv;

    /* step 6: radius, pixel centre */
    real mid = (real)0.5 * (a + c);
    real sq = R_SQRT(rmax((real)0.1, mid * mid - det));
    real lambda1 = mid + sq, lambda2 = mid - sq;
    real my_radius = R_CEIL((real)3.0 * R_SQRT(rmax(lambda1, lambda2)));
    real ix = ndc2pix(projx, o->W), iy = ndc2pix(projy, o->H);

    /* step 7: tile rectangle (C truncation toward zero) */
    int rminx = imin(o->gx, imax(0, (int)((ix - my_radius) / (real)BLOCK_X)));
    int rminy = imin(o->gy, imax(0, (int)((iy - my_radius) / (real)BLOCK_Y)));
    int rmaxx = imin(o->gx, imax(0, (int)((ix + my_radius + (real)(BLOCK_X - 1)) / (real)BLOCK_X)));
    int rmaxy = imin(o->gy, imax(0, (int)((iy + my_radius + (real)(BLOCK_Y - 1)) / (real)BLOCK_Y)));
    if ((rmaxx - rminx) * (rmaxy - rminy) == 0) return;

    /* step 9 */
    o->depth[i] = tz;
    o->radii[i] = (int32_t)my_radius;
    o->xy[2 * i] = ix; o->xy[2 * i + 1] = iy;
    o->conic_op[4 * i] = conx; o->conic_op[4 * i + 1] = cony; o->conic_op[4 * i + 2] = conz; o->conic_op[4 * i + 3] = o->opac[i];
    o->rect[4 * i] = rminx; o->rect[4 * i + 1] = rminy; o->rect[4 * i + 2] = rmaxx; o->rect[4 * i + 3] = rmaxy;
    o->tiles_touched[i] = (uint32_t)((rmaxx - rminx) * (rmaxy - rminy));
}

/* stable-radix-sort order of upstream == ascending (tile, depth bits, emission order);
 * emission is by ascending Gaussian id and a Gaussian emits each tile once, so
 * (tile, depth, id) is a total order that reproduces it. Depths are > 0.2,
 * so comparing the float value equals comparing its bit pattern. */
static int inst_cmp(const void *pa, const void *pb) {
    const inst_t *a = (const inst_t *)pa, *b = (const inst_t *)pb;
    if (a->tile != b->tile) return a->tile < b->tile ? -1 : 1;
    if (a->depth != b->depth) return a->depth < b->depth ? -1 : 1;
    if (a->id != b->id) return a->id < b->id ? -1 : 1;
    return 0;
}

static void binning(oracle_raster *o) {
    int64_t R = 0;
    for (int i = 0; i < o->P; ++i) R += o->tiles_touched[i];
    free(o->list);
    o->list = (inst_t *)xcalloc((size_t)R, sizeof(inst_t));
    o->R = R;
    int64_t k = 0;
    for (int i = 0; i < o->P; ++i) {
        if (o->radii[i] <= 0) continue;
        const int32_t *r = &o->rect[4 * i];
        for (int y = r[1]; y < r[3]; ++y) for (int x = r[0]; x < r[2]; ++x) {
            o->list[k].tile = (uint32_t)(y * o->gx + x);
            o->list[k].depth = o->depth[i];
            o->list[k].id = (uint32_t)i;
            ++k;
        }
    }
    qsort(o->list, (size_t)R, sizeof(inst_t), inst_cmp);
    memset(o->ranges, 0, (size_t)o->gx * o->gy * 2 * sizeof(uint32_t));
    for (int64_t j = 0; j < R; ++j) {
        uint32_t t = o->list[j].tile;
        if (j == 0 || o->list[j - 1].tile != t) o->ranges[2 * t] = (uint32_t)j;
        if (j == R - 1 || o->list[j + 1].tile != t) o->ranges[2 * t + 1] = (uint32_t)(j + 1);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* A.2 render                                                                                   */
/* ------------------------------------------------------------------------------------------ */

static inline int near_rel(real v, real thr, real rel) {
    real d = v - thr; if (d < 0) d = -d;
    return d <= rel * thr;
}

static void render_tile(oracle_raster *o, int tile) {
    const int C = o->C, W = o->W, H = o->H;
    const int tx0 = (tile % o->gx) * BLOCK_X, ty0 = (tile / o->gx) * BLOCK_Y;
    const uint32_t beg = o->ranges[2 * tile], end = o->ranges[2 * tile + 1];
    for (int ly = 0; ly < BLOCK_Y; ++ly) for (int lx = 0; lx < BLOCK_X; ++lx) {
        const int pxi = tx0 + lx, pyi = ty0 + ly;
        if (pxi >= W || pyi >= H) continue;
        const real pfx = (real)pxi, pfy = (real)pyi;
        real T = (real)1.0, Cacc[MAXC] = {0}, D = 0, Wgt = 0;
        uint32_t contributor = 0, last = 0;
        uint8_t amb = 0;
        for (uint32_t j = beg; j < end; ++j) {
            const uint32_t g = o->list[j].id;
            contributor++;
            const real dx = o->xy[2 * g] - pfx, dy = o->xy[2 * g + 1] - pfy;
            const real *co = &o->conic_op[4 * g];
            const real power = (real)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
            if (power > (real)-1e-6 && power != (real)0.0) amb |= 1;   /* power>0 skip is rounding-only */
            if (power > (real)0.0) continue;
            const real ea = co[3] * R_EXP(power);
            const real alpha = rmin((real)0.99, ea);
            if (near_rel(ea, (real)(1.0 / 255.0), o->amb_rel)) amb |= 1;
            if (alpha < (real)(1.0 / 255.0)) continue;
            const real test_T = T * ((real)1.0 - alpha);
            if (near_rel(test_T, (real)0.0001, (real)10 * o->amb_rel)) amb |= 1;
            if (test_T < (real)0.0001) break;     /* done = true */
            const real w = alpha * T;
            for (int ch = 0; ch < C; ++ch) Cacc[ch] += o->colors[(size_t)g * C + ch] * w;
            D += o->depth[g] * w;
            Wgt += w;
            T = test_T;
            last = contributor;
        }
        const size_t pix = (size_t)pyi * W + pxi;
        o->n_contrib[pix] = last;
        for (int ch = 0; ch < C; ++ch) o->out_color[(size_t)ch * H * W + pix] = Cacc[ch] + T * o->bg[ch];
        o->out_depth[pix] = D;
        o->out_alpha[pix] = Wgt;
        o->ambig[pix] = amb;
    }
}

/* Forward: inputs are row-major arrays of `real`. view/proj are the TRANSPOSED (row-vector)
 * 4x4 matrices exactly as the plugin passes them (threestudio/utils/ops.py:402-410). */
int oracle_raster_forward(oracle_raster *o, const real *means, const real *scales, const real *rots,
                          const real *opac, const real *colors, const real *view, const real *proj,
                          double tanfovx, double tanfovy, const real *bg, double scale_mod) {
    const int P = o->P;
    memcpy(o->means, means, sizeof(real) * 3 * P);
    memcpy(o->scales, scales, sizeof(real) * 3 * P);
    memcpy(o->rots, rots, sizeof(real) * 4 * P);
    memcpy(o->opac, opac, sizeof(real) * P);
    memcpy(o->colors, colors, sizeof(real) * o->C * P);
    memcpy(o->view, view, sizeof(real) * 16);
    memcpy(o->proj, proj, sizeof(real) * 16);
    memset(o->bg, 0, sizeof(o->bg));
    memcpy(o->bg, bg, sizeof(real) * o->C);
    o->tanfovx = (real)tanfovx; o->tanfovy = (real)tanfovy; o->scale_mod = (real)scale_mod;
    o->focal_x = (real)o->W / ((real)2.0 * o->tanfovx);
    o->focal_y = (real)o->H / ((real)2.0 * o->tanfovy);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) preprocess_one(o, i);
    binning(o);
    const int T = o->gx * o->gy;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < T; ++t) render_tile(o, t);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* A.3 backward                                                                                 */
/* ------------------------------------------------------------------------------------------ */

static inline void atomic_add_d(double *p, double v) {
#pragma omp atomic
    *p += v;
}

static void render_bwd_tile(oracle_raster *o, int tile, const real *dL_dC, const real *dL_dD, const real *dL_dA) {
    const int C = o->C, W = o->W, H = o->H;
    const int tx0 = (tile % o->gx) * BLOCK_X, ty0 = (tile / o->gx) * BLOCK_Y;
    const uint32_t beg = o->ranges[2 * tile], end = o->ranges[2 * tile + 1];
    const real ddelx_dx = (real)0.5 * (real)W, ddely_dy = (real)0.5 * (real)H;
    (void)end;
    for (int ly = 0; ly < BLOCK_Y; ++ly) for (int lx = 0; lx < BLOCK_X; ++lx) {
        const int pxi = tx0 + lx, pyi = ty0 + ly;
        if (pxi >= W || pyi >= H) continue;
        const size_t pix = (size_t)pyi * W + pxi;
        const real pfx = (real)pxi, pfy = (real)pyi;
        const uint32_t last = o->n_contrib[pix];
        const real T_final = (real)1.0 - o->out_alpha[pix];
        real T = T_final;
        real gC[MAXC];
        for (int ch = 0; ch < C; ++ch) gC[ch] = dL_dC[(size_t)ch * H * W + pix];
        const real gD = dL_dD ? dL_dD[pix] : (real)0, gA = dL_dA ? dL_dA[pix] : (real)0;
        real accum_rec[MAXC] = {0}, last_color[MAXC] = {0};
        real accum_d = 0, last_depth = 0, accum_a = 0, last_alpha = 0;
        real bg_dot = 0;
        for (int ch = 0; ch < C; ++ch) bg_dot += o->bg[ch] * gC[ch];
        for (uint32_t k = last; k-- > 0;) {   /* contributor index k+1 .. 1, back to front */
            const uint32_t j = beg + k;
            const uint32_t g = o->list[j].id;
            const real dx = o->xy[2 * g] - pfx, dy = o->xy[2 * g + 1] - pfy;
            const real *co = &o->conic_op[4 * g];
            const real power = (real)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
            if (power > (real)0.0) continue;
            const real G = R_EXP(power);
            const real alpha = rmin((real)0.99, co[3] * G);
            if (alpha < (real)(1.0 / 255.0)) continue;
            T = T / ((real)1.0 - alpha);
            const real w = alpha * T;
            real dL_dalpha = 0;
            for (int ch = 0; ch < C; ++ch) {
                const real c = o->colors[(size_t)g * C + ch];
                accum_rec[ch] = last_alpha * last_color[ch] + ((real)1.0 - last_alpha) * accum_rec[ch];
                last_color[ch] = c;
                dL_dalpha += (c - accum_rec[ch]) * gC[ch];
                atomic_add_d(&o->g_color[(size_t)g * C + ch], (double)(w * gC[ch]));
            }
            accum_d = last_alpha * last_depth + ((real)1.0 - last_alpha) * accum_d;
            last_depth = o->depth[g];
            dL_dalpha += (last_depth - accum_d) * gD;
            atomic_add_d(&o->g_depth[g], (double)(w * gD));
            accum_a = last_alpha + ((real)1.0 - last_alpha) * accum_a;
            dL_dalpha += ((real)1.0 - accum_a) * gA;
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / ((real)1.0 - alpha)) * bg_dot;
            const real dL_dG = co[3] * dL_dalpha;
            const real gdx = G * dx, gdy = G * dy;
            const real dG_ddelx = -gdx * co[0] - gdy * co[1];
            const real dG_ddely = -gdy * co[2] - gdx * co[1];
            atomic_add_d(&o->g_mean2d[2 * g], (double)(dL_dG * dG_ddelx * ddelx_dx));
            atomic_add_d(&o->g_mean2d[2 * g + 1], (double)(dL_dG * dG_ddely * ddely_dy));
            atomic_add_d(&o->g_conic[3 * g], (double)((real)-0.5 * gdx * dx * dL_dG));
            atomic_add_d(&o->g_conic[3 * g + 1], (double)((real)-0.5 * gdx * dy * dL_dG));
            atomic_add_d(&o->g_conic[3 * g + 2], (double)((real)-0.5 * gdy * dy * dL_dG));
            atomic_add_d(&o->g_opac[g], (double)(G * dL_dalpha));
        }
    }
}

/* preprocess backward for one Gaussian (A.3 "preprocess-bwd").  Outputs are `real`. */
static void preprocess_bwd_one(oracle_raster *o, int i, real *dmeans, real *dscales, real *drots) {
    for (int k = 0; k < 3; ++k) { dmeans[3 * i + k] = 0; dscales[3 * i + k] = 0; }
    for (int k = 0; k < 4; ++k) drots[4 * i + k] = 0;
    if (o->radii[i] <= 0) return;
    const real *V = o->view, *PV = o->proj;
    const real px = o->means[3 * i], py = o->means[3 * i + 1], pz = o->means[3 * i + 2];

    /* ---- (1) conic -> cov2D -> (cov3D, t) ---- */
    const real a = o->cov2d[3 * i], b = o->cov2d[3 * i + 1], c = o->cov2d[3 * i + 2];
    const real gx = (real)o->g_conic[3 * i], gy = (real)o->g_conic[3 * i + 1], gz = (real)o->g_conic[3 * i + 2];
    const real denom = a * c - b * b;
    const real d2 = (real)1.0 / (denom * denom + (real)0.0000001);
    real dL_da = 0, dL_db = 0, dL_dc = 0;
    /* recompute forward quantities */
    real tx = V[0] * px + V[4] * py + V[8] * pz + V[12];
    real ty = V[1] * px + V[5] * py + V[9] * pz + V[13];
    real tz = V[2] * px + V[6] * py + V[10] * pz + V[14];
    const real limx = (real)1.3 * o->tanfovx, limy = (real)1.3 * o->tanfovy;
    const real txtz = tx / tz, tytz = ty / tz;
    const real xmul = (txtz < -limx || txtz > limx) ? (real)0 : (real)1;
    const real ymul = (tytz < -limy || tytz > limy) ? (real)0 : (real)1;
    const real cx = rmin(limx, rmax(-limx, txtz)) * tz;
    const real cy = rmin(limy, rmax(-limy, tytz)) * tz;
    const real fx = o->focal_x, fy = o->focal_y;
    real J00 = fx / tz, J02 = -(fx * cx) / (tz * tz), J11 = fy / tz, J12 = -(fy * cy) / (tz * tz);
    real A[2][3], B[2][3], S[3][3];
    const real *c3 = &o->cov3d[6 * i];
    S[0][0] = c3[0]; S[0][1] = S[1][0] = c3[1]; S[0][2] = S[2][0] = c3[2];
    S[1][1] = c3[3]; S[1][2] = S[2][1] = c3[4]; S[2][2] = c3[5];
    for (int j = 0; j < 3; ++j) {
        A[0][j] = J00 * V[0 + 4 * j] + J02 * V[2 + 4 * j];
        A[1][j] = J11 * V[1 + 4 * j] + J12 * V[2 + 4 * j];
    }
    for (int r = 0; r < 2; ++r) for (int j = 0; j < 3; ++j)
        B[r][j] = A[r][0] * S[0][j] + A[r][1] * S[1][j] + A[r][2] * S[2][j];

    real GS[3][3] = {{0}};   /* dL/dSigma as a full symmetric matrix */
    real dA[2][3] = {{0}};
    if (denom * denom + (real)0.0000001 != (real)0) {
        dL_da = d2 * (-c * c * gx + (real)2 * b * c * gy + (denom - a * c) * gz);
        dL_dc = d2 * (-a * a * gz + (real)2 * a * b * gy + (denom - a * c) * gx);
        dL_db = d2 * (real)2 * (b * c * gx - (denom + (real)2 * b * b) * gy + a * b * gz);
        const real hb = (real)0.5 * dL_db;
        for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k)
            GS[j][k] = dL_da * A[0][j] * A[0][k] + hb * (A[0][j] * A[1][k] + A[1][j] * A[0][k]) + dL_dc * A[1][j] * A[1][k];
        for (int j = 0; j < 3; ++j) {
            dA[0][j] = (real)2 * (dL_da * B[0][j] + hb * B[1][j]);
            dA[1][j] = (real)2 * (hb * B[0][j] + dL_dc * B[1][j]);
        }
    }
    /* A = J W3 : dJ_im = sum_j dA_ij W3[m][j],  W3[m][j] = V[m + 4 j] */
    real dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
    for (int j = 0; j < 3; ++j) {
        dJ00 += dA[0][j] * V[0 + 4 * j];
        dJ02 += dA[0][j] * V[2 + 4 * j];
        dJ11 += dA[1][j] * V[1 + 4 * j];
        dJ12 += dA[1][j] * V[2 + 4 * j];
    }
    const real tz1 = (real)1 / tz, tz2 = tz1 * tz1, tz3 = tz2 * tz1;
    const real dL_dtx = xmul * -fx * tz2 * dJ02;
    const real dL_dty = ymul * -fy * tz2 * dJ12;
    const real dL_dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + ((real)2 * fx * cx) * tz3 * dJ02 + ((real)2 * fy * cy) * tz3 * dJ12;
    real dm[3];
    for (int j = 0; j < 3; ++j) dm[j] = V[0 + 4 * j] * dL_dtx + V[1 + 4 * j] * dL_dty + V[2 + 4 * j] * dL_dtz;

    /* ---- (2) 2D mean -> 3D mean through the perspective divide ---- */
    {
        const real hx = PV[0] * px + PV[4] * py + PV[8] * pz + PV[12];
        const real hy = PV[1] * px + PV[5] * py + PV[9] * pz + PV[13];
        const real hw = PV[3] * px + PV[7] * py + PV[11] * pz + PV[15];
        const real m_w = (real)1.0 / (hw + (real)0.0000001);
        const real mul1 = hx * m_w * m_w, mul2 = hy * m_w * m_w;
        const real g2x = (real)o->g_mean2d[2 * i], g2y = (real)o->g_mean2d[2 * i + 1];
        for (int k = 0; k < 3; ++k)
            dm[k] += (PV[0 + 4 * k] * m_w - PV[3 + 4 * k] * mul1) * g2x + (PV[1 + 4 * k] * m_w - PV[3 + 4 * k] * mul2) * g2y;
    }
    /* ---- (3) depth -> 3D mean (ashawkey) ---- */
    {
        const real gd = (real)o->g_depth[i];
        for (int k = 0; k < 3; ++k) dm[k] += (V[2 + 4 * k] - V[3 + 4 * k] * tz) * gd;
    }
    for (int k = 0; k < 3; ++k) dmeans[3 * i + k] = dm[k];

    /* ---- (5) cov3D -> (scale, quaternion): Sigma = L L^T, L = Rm diag(s) ---- */
    real Rm[3][3], L[3][3], dLm[3][3];
    const real *q = &o->rots[4 * i];
    quat_to_R(q, Rm);
    real s[3] = { o->scale_mod * o->scales[3 * i], o->scale_mod * o->scales[3 * i + 1], o->scale_mod * o->scales[3 * i + 2] };
    for (int a2 = 0; a2 < 3; ++a2) for (int k = 0; k < 3; ++k) L[a2][k] = Rm[a2][k] * s[k];
    for (int a2 = 0; a2 < 3; ++a2) for (int k = 0; k < 3; ++k)
        dLm[a2][k] = (real)2 * (GS[a2][0] * L[0][k] + GS[a2][1] * L[1][k] + GS[a2][2] * L[2][k]);
    real dR[3][3];
    for (int k = 0; k < 3; ++k) {
        real ds = 0;
        for (int a2 = 0; a2 < 3; ++a2) { ds += dLm[a2][k] * Rm[a2][k]; dR[a2][k] = dLm[a2][k] * s[k]; }
        dscales[3 * i + k] = o->scale_mod * ds;
    }
    const real r = q[0], x = q[1], y = q[2], z = q[3];
    drots[4 * i + 0] = (real)2 * (z * (dR[1][0] - dR[0][1]) + y * (dR[0][2] - dR[2][0]) + x * (dR[2][1] - dR[1][2]));
    drots[4 * i + 1] = (real)2 * (y * (dR[0][1] + dR[1][0]) + z * (dR[0][2] + dR[2][0]) + r * (dR[2][1] - dR[1][2])) - (real)4 * x * (dR[1][1] + dR[2][2]);
    drots[4 * i + 2] = (real)2 * (x * (dR[0][1] + dR[1][0]) + r * (dR[0][2] - dR[2][0]) + z * (dR[1][2] + dR[2][1])) - (real)4 * y * (dR[0][0] + dR[2][2]);
    drots[4 * i + 3] = (real)2 * (r * (dR[1][0] - dR[0][1]) + x * (dR[0][2] + dR[2][0]) + y * (dR[1][2] + dR[2][1])) - (real)4 * z * (dR[0][0] + dR[1][1]);
}

/* Backward. dL_dD / dL_dA may be NULL (treated as zero). Outputs (all `real`):
 * dmeans3D [P,3], dmeans2D [P,3] (z = 0; NDC-scaled, A.3), dcolors [P,C], dopac [P],
 * dscales [P,3], drots [P,4]. */
int oracle_raster_backward(oracle_raster *o, const real *dL_dC, const real *dL_dD, const real *dL_dA,
                           real *dmeans3D, real *dmeans2D, real *dcolors, real *dopac, real *dscales, real *drots) {
    const int P = o->P, C = o->C;
    memset(o->g_mean2d, 0, sizeof(double) * 2 * P);
    memset(o->g_conic, 0, sizeof(double) * 3 * P);
    memset(o->g_opac, 0, sizeof(double) * P);
    memset(o->g_color, 0, sizeof(double) * C * P);
    memset(o->g_depth, 0, sizeof(double) * P);
    const int T = o->gx * o->gy;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < T; ++t) render_bwd_tile(o, t, dL_dC, dL_dD, dL_dA);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        preprocess_bwd_one(o, i, dmeans3D, dscales, drots);
        dmeans2D[3 * i] = (real)o->g_mean2d[2 * i];
        dmeans2D[3 * i + 1] = (real)o->g_mean2d[2 * i + 1];
        dmeans2D[3 * i + 2] = 0;
        for (int ch = 0; ch < C; ++ch) dcolors[(size_t)i * C + ch] = (real)o->g_color[(size_t)i * C + ch];
        dopac[i] = (real)o->g_opac[i];
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* accessors                                                                                    */
/* ------------------------------------------------------------------------------------------ */
int64_t oracle_raster_num_rendered(const oracle_raster *o) { return o->R; }
const real *oracle_raster_color(const oracle_raster *o) { return o->out_color; }
const real *oracle_raster_depth(const oracle_raster *o) { return o->out_depth; }
const real *oracle_raster_alpha(const oracle_raster *o) { return o->out_alpha; }
const int32_t *oracle_raster_radii(const oracle_raster *o) { return o->radii; }
const int32_t *oracle_raster_rect(const oracle_raster *o) { return o->rect; }
const uint32_t *oracle_raster_tiles_touched(const oracle_raster *o) { return o->tiles_touched; }
const uint32_t *oracle_raster_ranges(const oracle_raster *o) { return o->ranges; }
const uint32_t *oracle_raster_n_contrib(const oracle_raster *o) { return o->n_contrib; }
const uint8_t *oracle_raster_ambiguous(const oracle_raster *o) { return o->ambig; }
const real *oracle_raster_xy(const oracle_raster *o) { return o->xy; }
const real *oracle_raster_gdepth(const oracle_raster *o) { return o->depth; }
const real *oracle_raster_conic_opacity(const oracle_raster *o) { return o->conic_op; }
void oracle_raster_point_list(const oracle_raster *o, uint32_t *ids, uint32_t *tiles) {
    for (int64_t j = 0; j < o->R; ++j) { ids[j] = o->list[j].id; if (tiles) tiles[j] = o->list[j].tile; }
}
int oracle_real_bytes(void) { return (int)sizeof(real); }
