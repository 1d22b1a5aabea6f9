// Fused sparse-control-point skinning (LBS / dual-quaternion / hybrid) + per-face surface-bound
// Gaussian update, forward and backward.
//
// Replaces the ~150 tiny PyTorch/pypose kernels per view of
//   custom/threestudio-dreammesh4d/geometry/dynamic_sugar.py:487-613 (_get_timed_vertex_attributes_from_dg),
//   :657-676 (get_timed_gs_attributes), :726-743 (_get_gs_xyz_from_vertex), :877-889 (fuse_rotations),
//   :347-364 (get_timed_gs_normals) and custom/threestudio-dreammesh4d/utils/dual_quaternions.py:94-131,184-231
// with two kernels per direction for ALL timestamps of a step:
//   skin_vertex_*   : one thread per (timestamp, vertex)   — K-neighbour gather, LBS + DQS + hybrid blend,
//                     log-space rotation blend
//   skin_gaussian_* : one thread per (timestamp, face)     — the g Gaussians of a face share the three
//                     vertex gathers and quaternion logs; writes rasterizer-ready means / wxyz rotations /
//                     normals
// Quaternion algebra follows pypose's SO3 semantics (xyzw).  Backward = exact
// Euclidean gradients of the forward formulas.
#include <algorithm>
#include "raster_internal.cuh"

// register budgets (resident CTAs per SM) of the kernels whose occupancy is register-bound; tuning: scripts/tune_c5.sh
#ifndef DM4D_SKIN_GF_BLOCKS
#define DM4D_SKIN_GF_BLOCKS 4      // measured at C5 x 8 timestamps (us): unbounded 114, 3: 114, 4: 103
#endif
#ifndef DM4D_SKIN_GB_BLOCKS
#define DM4D_SKIN_GB_BLOCKS 4      // 2: 237, 3: 196, 4: 187
#endif
#ifndef DM4D_SKIN_NB_BLOCKS
#define DM4D_SKIN_NB_BLOCKS 4      // vertex backward (upstream + node kernels): unbounded 222, 3: 222, 4: 210
#endif

namespace {

constexpr float EPS_LIE = 1e-6f;

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ f3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, f3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
__device__ __forceinline__ f3 vec(float4 q) { return mk3(q.x, q.y, q.z); }
__device__ __forceinline__ float4 mkq(f3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator*(float s, float4 a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// Hamilton product, xyzw
__device__ __forceinline__ float4 qmul(float4 a, float4 b) {
    const f3 av = vec(a), bv = vec(b);
    const f3 v = a.w * bv + b.w * av + cross(av, bv);
    return mkq(v, a.w * b.w - dot(av, bv));
}
__device__ __forceinline__ float4 qconj(float4 q) { return make_float4(-q.x, -q.y, -q.z, q.w); }
// pypose Act: p + 2 w (v x p) + 2 v x (v x p)
__device__ __forceinline__ f3 qact(float4 q, f3 p) {
    const f3 v = vec(q);
    const f3 uv = cross(v, p);
    return p + 2.f * (q.w * uv + cross(v, uv));
}
// gradient of <g, qact(q,p)> w.r.t. q (exact for non-unit q as well)
__device__ __forceinline__ void qact_bwd(float4 q, f3 p, f3 g, float4& dq) {
    const f3 v = vec(q);
    const f3 pxg = cross(p, g);
    const float vp = dot(v, p), gv = dot(g, v), gp = dot(g, p);
    const f3 dv = 2.f * (q.w * pxg) + 2.f * (vp * g + gv * p - (2.f * gp) * v);
    dq = mkq(dv, 2.f * dot(g, cross(v, p)));
}

__device__ __forceinline__ f3 so3_log(float4 q) {
    const f3 v = vec(q);
    const float n = sqrtf(dot(v, v));
    float f;
    if (n > EPS_LIE) f = 2.f * atanf(n / q.w) / n;
    else f = 2.f / q.w - (2.f / 3.f) * n * n / (q.w * q.w * q.w);
    return f * v;
}
// dL/dq given g = dL/d(log q):  f g + (g.v) c1 v  for the vector part, (g.v) c2 for w.  The three coefficients depend
// on q only (hoisted out of the incidence loop of the node-centric backward).
struct LogBwdCoef { float f, c1, c2; };
__device__ __forceinline__ LogBwdCoef so3_log_bwd_coef(float4 q) {
    const f3 v = vec(q);
    const float n2 = dot(v, v), n = sqrtf(n2), w = q.w;
    LogBwdCoef c;
    if (n > EPS_LIE) {
        c.f = 2.f * atanf(n / w) / n;
        const float s = w * w + n2;
        c.c1 = (2.f * w / s - c.f) / n2;
        c.c2 = -2.f / s;
    } else {
        const float w2 = w * w;
        c.f = 2.f / w - (2.f / 3.f) * n2 / (w2 * w);
        c.c1 = -(4.f / 3.f) / (w2 * w);
        c.c2 = -2.f / w2 + 2.f * n2 / (w2 * w2);
    }
    return c;
}
__device__ __forceinline__ float4 so3_log_bwd_apply(LogBwdCoef c, float4 q, f3 g) {
    const f3 v = vec(q);
    const float gv = dot(g, v);
    return mkq(c.f * g + (gv * c.c1) * v, gv * c.c2);
}
__device__ __forceinline__ float4 so3_log_bwd(float4 q, f3 g) { return so3_log_bwd_apply(so3_log_bwd_coef(q), q, g); }
__device__ __forceinline__ float4 so3_exp(f3 x) {
    const float th2 = dot(x, x), th = sqrtf(th2);
    float a, w;
    if (th > EPS_LIE) { a = sinf(0.5f * th) / th; w = cosf(0.5f * th); }
    else { a = 0.5f - th2 / 48.f + th2 * th2 / 3840.f; w = 1.f - th2 / 8.f + th2 * th2 / 384.f; }
    return mkq(a * x, w);
}
// exp and the three coefficients of its derivative from ONE sincos evaluation (the backward of the Gaussian stage needs
// both for every Gaussian: dq = exp(xi) to rebuild the rotation, d exp / d xi to differentiate it)
struct ExpBwdCoef { float a, da_over_th, dw_coef; };
__device__ __forceinline__ float4 so3_exp_coef(f3 x, ExpBwdCoef& c) {
    const float th2 = dot(x, x), th = sqrtf(th2);
    float w;
    if (th > EPS_LIE) {
        float sn, cs;
        sincosf(0.5f * th, &sn, &cs);
        c.a = sn / th; w = cs;
        c.da_over_th = (0.5f * cs * th - sn) / (th2 * th);
        c.dw_coef = -0.5f * c.a;
    } else {
        c.a = 0.5f - th2 / 48.f + th2 * th2 / 3840.f; w = 1.f - th2 / 8.f + th2 * th2 / 384.f;
        c.da_over_th = -1.f / 24.f + th2 / 960.f;
        c.dw_coef = -0.25f + th2 / 96.f;
    }
    return mkq(c.a * x, w);
}
__device__ __forceinline__ f3 so3_exp_bwd_apply(ExpBwdCoef c, f3 x, float4 g) {
    const f3 gv = vec(g);
    return c.a * gv + (dot(gv, x) * c.da_over_th + g.w * c.dw_coef) * x;
}
// dL/dx given g = dL/d(exp x) (xyzw)
__device__ __forceinline__ f3 so3_exp_bwd(f3 x, float4 g) {
    const float th2 = dot(x, x), th = sqrtf(th2);
    float a, da_over_th, dw_coef;   // dw/dx_j = dw_coef * x_j
    if (th > EPS_LIE) {
        const float s = sinf(0.5f * th), c = cosf(0.5f * th);
        a = s / th;
        da_over_th = (0.5f * c * th - s) / (th2 * th);
        dw_coef = -0.5f * a;
    } else {
        a = 0.5f - th2 / 48.f + th2 * th2 / 3840.f;
        da_over_th = -1.f / 24.f + th2 / 960.f;
        dw_coef = -0.25f + th2 / 96.f;
    }
    const f3 gv = vec(g);
    return a * gv + (dot(gv, x) * da_over_th + g.w * dw_coef) * x;
}

__device__ __forceinline__ float4 ldq(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct SkinK {
    dm4d_skin_desc d;
    int P;
    float* verts; float* vert_rot; float* means; float* rots; float* normals;
};

// ------------------------------------------------------------------------------------------------
// vertex stage
// ------------------------------------------------------------------------------------------------
struct VertSums {
    f3 x_l; float4 sq_r, sq_d; float lam_raw; f3 xi;
};

// Per-(timestamp, node) quantities every (vertex, neighbour) pair needs: log of the node rotation, its normalised
// quaternion and the dual part — an atan, a reciprocal square root and a quaternion product per PAIR when evaluated
// inline.  With desc.node_scratch a pre-pass of n_t * M threads evaluates them once: 3 float4 per (t, node).
__device__ __forceinline__ void node_pre_eval(int method, f3 tr, float4 q, f3& nlog, float4& qn, float4& qd) {
    nlog = so3_log(q);
    qn = make_float4(0, 0, 0, 0); qd = qn;
    if (method != 0) {
        const float inv = 1.f / sqrtf(dot4(q, q));
        qn = inv * q;
        qd = qmul(mkq(0.5f * tr, 0.f), qn);
    }
}
__global__ void __launch_bounds__(DM4D_BLOCK) skin_node_pre_kernel(dm4d_skin_desc d, float4* __restrict__ pre) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n_t * d.M) return;
    f3 nlog; float4 qn, qd;
    node_pre_eval(d.method, ld3(d.node_trans + (size_t)i * 3), ldq(d.node_rot + (size_t)i * 4), nlog, qn, qd);
    pre[(size_t)i * 3] = mkq(nlog, 0.f);
    pre[(size_t)i * 3 + 1] = qn;
    pre[(size_t)i * 3 + 2] = qd;
}

__device__ __forceinline__ void vertex_accumulate(const dm4d_skin_desc& d, int t, int v, f3 x, VertSums& s) {
    s.x_l = mk3(0, 0, 0); s.sq_r = make_float4(0, 0, 0, 0); s.sq_d = make_float4(0, 0, 0, 0); s.lam_raw = 0.f; s.xi = mk3(0, 0, 0);
    const float4* pre = reinterpret_cast<const float4*>(d.node_scratch);
    for (int k = 0; k < d.K; ++k) {
        const int n = d.nbr_idx[(size_t)v * d.K + k];
        const float w = d.nbr_w[(size_t)v * d.K + k];
        const size_t base = (size_t)t * d.M + n;
        const f3 tr = ld3(d.node_trans + base * 3);
        const float4 q = ldq(d.node_rot + base * 4);
        if (d.method != 1) {
            const float* S = d.node_scale + base * 9;
            const f3 y = mk3(S[0] * x.x + S[1] * x.y + S[2] * x.z, S[3] * x.x + S[4] * x.y + S[5] * x.z, S[6] * x.x + S[7] * x.y + S[8] * x.z);
            s.x_l = s.x_l + w * (qact(q, y) + tr);
        }
        f3 nlog; float4 qn, qd;
        if (pre) {
            nlog = vec(pre[base * 3]);
            if (d.method != 0) { qn = pre[base * 3 + 1]; qd = pre[base * 3 + 2]; }
        } else {
            node_pre_eval(d.method, tr, q, nlog, qn, qd);
        }
        if (d.method != 0) {
            s.sq_r = s.sq_r + w * qn;
            s.sq_d = s.sq_d + w * qd;
        }
        if (d.method == 2) s.lam_raw += w * d.node_opacity[base];
        s.xi = s.xi + w * nlog;
    }
}

__global__ void __launch_bounds__(DM4D_BLOCK) skin_vertex_forward_kernel(SkinK a) {
    const dm4d_skin_desc& d = a.d;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)d.n_t * d.V) return;
    const int t = (int)(idx / d.V), v = (int)(idx - (long long)t * d.V);
    const f3 x = ld3(d.rest_verts + (size_t)v * 3);
    VertSums s;
    vertex_accumulate(d, t, v, x, s);
    f3 out;
    if (d.method == 0) out = s.x_l;
    else {
        const float inv = 1.f / sqrtf(dot4(s.sq_r, s.sq_r));
        const float4 qn = inv * s.sq_r, dn = inv * s.sq_d;
        const f3 trans = vec(qmul(2.f * dn, qconj(qn)));
        const f3 x_d = qact(qn, x) + trans;
        if (d.method == 1) out = x_d;
        else {
            const float lam = fminf(s.lam_raw + 0.4f, 1.0f);
            out = lam * s.x_l + (1.f - lam) * x_d;
        }
    }
    st3(a.verts + (size_t)idx * 3, out);
    *reinterpret_cast<float4*>(a.vert_rot + (size_t)idx * 4) = so3_exp(s.xi);
}

struct SkinBwdK {
    dm4d_skin_desc d;
    int P;
    const float* verts; const float* vert_rot;
    const float* g_means; const float* g_rots; const float* g_normals;
    float* dverts; float* dvert_rot;                       // scratch [n_t,V,4] each: gradients arriving from the faces
    const float* dverts_in; const float* dvert_rot_in;     // optional direct gradients [n_t,V,3] / [n_t,V,4]
    // vertex-centric mode (all three or none): the Gaussian stage writes one 32-byte record per (t, face corner) and the
    // vertex stage gathers them through the vertex -> corner lists: no reductions, no memset, reproducible order
    const int32_t* vinc_ptr; const int32_t* vinc; float* corner;
    float* dn_trans; float* dn_rot; float* dn_scale; float* dn_opac;
};

// Upstream gradients of one (timestamp, vertex): everything of its backward that does not depend on which of the K
// neighbours is being differentiated.
struct VertUp {
    f3 x, dxi, dx_l; float dlam_raw; float4 dsq_r, dsq_d;
};

__device__ __forceinline__ void vertex_upstream(const SkinBwdK& a, int t, int v, VertUp& u) {
    const dm4d_skin_desc& d = a.d;
    const long long idx = (long long)t * d.V + v;
    const f3 x = ld3(d.rest_verts + (size_t)v * 3);
    f3 gx = mk3(0, 0, 0);
    float4 gr = make_float4(0, 0, 0, 0);
    if (a.corner) {
        const float* ct = a.corner + (size_t)t * d.F * 3 * 8;
        for (int i = a.vinc_ptr[v]; i < a.vinc_ptr[v + 1]; ++i) {
            const float* rec = ct + (size_t)a.vinc[i] * 8;
            gx = gx + vec(ldq(rec));
            gr = gr + ldq(rec + 4);
        }
    } else {
        gx = vec(ldq(a.dverts + (size_t)idx * 4));
        gr = ldq(a.dvert_rot + (size_t)idx * 4);
    }
    if (a.dverts_in) gx = gx + ld3(a.dverts_in + (size_t)idx * 3);
    if (a.dvert_rot_in) gr = gr + ldq(a.dvert_rot_in + (size_t)idx * 4);
    VertSums s;
    vertex_accumulate(d, t, v, x, s);
    u.x = x;
    u.dxi = so3_exp_bwd(s.xi, gr);                       // rotation: r = exp(xi)
    u.dx_l = mk3(0, 0, 0);
    u.dlam_raw = 0.f;
    u.dsq_r = make_float4(0, 0, 0, 0); u.dsq_d = make_float4(0, 0, 0, 0);
    if (d.method == 0) { u.dx_l = gx; return; }
    const float N2 = dot4(s.sq_r, s.sq_r), N = sqrtf(N2), inv = 1.f / N;
    const float4 qn = inv * s.sq_r, dn = inv * s.sq_d;
    f3 dx_d;
    if (d.method == 1) dx_d = gx;
    else {
        const float lam_in = s.lam_raw + 0.4f;
        const float lam = fminf(lam_in, 1.0f);
        const f3 trans = vec(qmul(2.f * dn, qconj(qn)));
        const f3 x_d = qact(qn, x) + trans;
        u.dx_l = lam * gx;
        dx_d = (1.f - lam) * gx;
        if (lam_in <= 1.0f) u.dlam_raw = dot(gx, s.x_l - x_d);
    }
    // x_d = act(qn, x) + xyz(2 dn (x) conj(qn))
    float4 dqn;
    qact_bwd(qn, x, dx_d, dqn);
    const float4 gm = mkq(dx_d, 0.f);                       // gradient of m = (2 dn) (x) conj(qn)
    const float4 ddn = 2.f * qmul(gm, qn);                  // d/d(a) <g, a (x) b> = g (x) conj(b), b = conj(qn)
    const float4 dconj = qmul(qconj(2.f * dn), gm);         // d/d(b) = conj(a) (x) g
    dqn = dqn + make_float4(-dconj.x, -dconj.y, -dconj.z, dconj.w);
    // qn = sq_r / N, dn = sq_d / N
    const float c = (dot4(qn, dqn) + dot4(dn, ddn)) * inv;
    u.dsq_r = inv * dqn + (-c) * qn;
    u.dsq_d = inv * ddn;
}

// Gradient of one (vertex, neighbour) incidence w.r.t. the node's attributes: g[0..2] translation, g[3..6] rotation
// (xyzw), g[7..15] scale (row-major 3x3), g[16] opacity (lbs weight).  tr / q / S are the node's attributes at t.
__device__ __forceinline__ void incidence_gradient(int method, float w, const VertUp& u, f3 tr, float4 q, LogBwdCoef lc,
                                                   const float* S, float (&g)[17]) {
    f3 dtr = mk3(0, 0, 0);
    float4 dq = so3_log_bwd_apply(lc, q, w * u.dxi);
#pragma unroll
    for (int i = 7; i < 17; ++i) g[i] = 0.f;
    if (method != 1) {
        const f3 x = u.x;
        const f3 y = mk3(S[0] * x.x + S[1] * x.y + S[2] * x.z, S[3] * x.x + S[4] * x.y + S[5] * x.z, S[6] * x.x + S[7] * x.y + S[8] * x.z);
        const f3 gl = w * u.dx_l;
        dtr = dtr + gl;
        float4 dq_act;
        qact_bwd(q, y, gl, dq_act);
        dq = dq + dq_act;
        // dy = R(q)^T g ; for the (generally unit) q the transpose action is act(conj(q), g) only when |q|=1,
        // so use the exact adjoint of p -> p + 2w(v x p) + 2 v x (v x p): g + 2w (g x v) + 2 (v x (v x g))
        const f3 vq = vec(q);
        const f3 dy = gl + 2.f * (q.w * cross(gl, vq) + cross(vq, cross(vq, gl)));
        g[7] = dy.x * x.x; g[8] = dy.x * x.y; g[9] = dy.x * x.z;
        g[10] = dy.y * x.x; g[11] = dy.y * x.y; g[12] = dy.y * x.z;
        g[13] = dy.z * x.x; g[14] = dy.z * x.y; g[15] = dy.z * x.z;
    }
    if (method != 0) {
        const float qq = dot4(q, q), inv = 1.f / sqrtf(qq);
        const float4 qn = inv * q;
        const float4 th = mkq(0.5f * tr, 0.f);
        // qd = th (x) qn
        const float4 gqd = w * u.dsq_d;
        float4 dqn = w * u.dsq_r + qmul(qconj(th), gqd);
        const float4 dth = qmul(gqd, qconj(qn));
        dtr = dtr + 0.5f * vec(dth);
        dq = dq + inv * (dqn + (-dot4(qn, dqn)) * qn);
    }
    if (method == 2) g[16] = w * u.dlam_raw;
    g[0] = dtr.x; g[1] = dtr.y; g[2] = dtr.z;
    g[3] = dq.x; g[4] = dq.y; g[5] = dq.z; g[6] = dq.w;
}

// Node-centric vertex backward (the default), two kernels:
//   skin_vertex_upstream_kernel : one thread per (timestamp, vertex) computes the vertex's upstream gradients ONCE and
//                                 stores them as 16 floats (64 B, four coalesced float4 stores);
//   skin_node_backward_kernel   : one CTA per (node, timestamp, split) walks the node's incidence list
//                                 (dm4d_skin_node_incidence: the (vertex, slot) pairs that reference the node, ascending),
//                                 gathers the 64 B records, applies the node-dependent ~100 flops per incidence, sums the
//                                 17 node gradients in registers and reduces them once per CTA.
// With one split the result is a plain store: no atomics, no memset, bit-reproducible.
__global__ void __launch_bounds__(DM4D_BLOCK) skin_vertex_upstream_kernel(SkinBwdK a, float4* __restrict__ up) {
    const dm4d_skin_desc& d = a.d;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)d.n_t * d.V) return;
    const int t = (int)(idx / d.V), v = (int)(idx - (long long)t * d.V);
    VertUp u;
    vertex_upstream(a, t, v, u);
    float4* o = up + (size_t)idx * 4;
    o[0] = make_float4(u.dxi.x, u.dxi.y, u.dxi.z, u.dlam_raw);
    o[1] = make_float4(u.dx_l.x, u.dx_l.y, u.dx_l.z, 0.f);
    o[2] = u.dsq_r;
    o[3] = u.dsq_d;
}

__global__ void __launch_bounds__(DM4D_BLOCK, DM4D_SKIN_NB_BLOCKS) skin_node_backward_kernel(SkinBwdK a, const float4* __restrict__ up,
                                                                        const int32_t* __restrict__ inc_ptr,
                                                                        const int32_t* __restrict__ inc) {
    __shared__ float part[DM4D_BLOCK / 32][17];
    const dm4d_skin_desc& d = a.d;
    const int n = blockIdx.x, t = blockIdx.y, K = d.K;
    const size_t base = (size_t)t * d.M + n;
    const f3 tr = ld3(d.node_trans + base * 3);
    const float4 q = ldq(d.node_rot + base * 4);
    const LogBwdCoef lc = so3_log_bwd_coef(q);
    float S[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) S[i] = d.method != 1 ? d.node_scale[base * 9 + i] : 0.f;
    float acc[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) acc[i] = 0.f;
    const int lo = inc_ptr[n], hi = inc_ptr[n + 1];
    const float4* upt = up + (size_t)t * d.V * 4;
    for (int i = lo + blockIdx.z * blockDim.x + threadIdx.x; i < hi; i += gridDim.z * blockDim.x) {
        const int e = inc[i];
        const int v = e / K;
        const float4 r0 = upt[(size_t)v * 4];
        VertUp u;
        u.dxi = mk3(r0.x, r0.y, r0.z); u.dlam_raw = r0.w;
        u.x = mk3(0, 0, 0); u.dx_l = u.x;
        u.dsq_r = make_float4(0, 0, 0, 0); u.dsq_d = u.dsq_r;
        if (d.method != 1) {                               // the records' halves a method does not use are not read
            const float4 r1 = upt[(size_t)v * 4 + 1];
            u.dx_l = mk3(r1.x, r1.y, r1.z);
            u.x = ld3(d.rest_verts + (size_t)v * 3);
        }
        if (d.method != 0) { u.dsq_r = upt[(size_t)v * 4 + 2]; u.dsq_d = upt[(size_t)v * 4 + 3]; }
        float g[17];
        incidence_gradient(d.method, d.nbr_w[e], u, tr, q, lc, S, g);
#pragma unroll
        for (int j = 0; j < 17; ++j) acc[j] += g[j];
    }
#pragma unroll
    for (int j = 0; j < 17; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int j = 0; j < 17; ++j) part[threadIdx.x >> 5][j] = acc[j];
    }
    __syncthreads();
    if (threadIdx.x < 17) {
        const int j = threadIdx.x;
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < DM4D_BLOCK / 32; ++w) tot += part[w][j];
        float* dst;
        if (j < 3) dst = a.dn_trans + base * 3 + j;
        else if (j < 7) dst = a.dn_rot + base * 4 + (j - 3);
        else if (j < 16) dst = a.dn_scale + base * 9 + (j - 7);
        else dst = a.dn_opac + base;
        if (gridDim.z == 1) *dst = tot;
        else if (tot != 0.f) atomicAdd(dst, tot);
    }
}

// List-free vertex backward (callers that pass no incidence lists): one thread per (timestamp, vertex) computes its upstream gradients and the 17
// node gradients of each of its K incidences; the warp then sums the lanes that hit the same (timestamp, node) with
// shuffles and issues ONE reduction per (warp, distinct node, component).  Vertices that are neighbours in memory are
// neighbours on the mesh and share their control nodes (1-3 distinct nodes per warp and slot), so the 17 atomics per
// incidence of a naive scatter (13.6 M on 8.7 k addresses at C5) become ~1 per 20 incidences, and nothing is staged in
// memory: no per-vertex records, no gather.  Measured at C5: 66 us against 55 us for the node-centric kernels above
// (245 us for the shared-memory tables of round 1); at 8 timestamps the distinct-node loop costs more than the gather
// (428 vs 226 us), so the host side builds the lists once and uses the node-centric path.  Not bit-reproducible.
__global__ void __launch_bounds__(DM4D_BLOCK) skin_vertex_backward_kernel(SkinBwdK a) {
    const dm4d_skin_desc& d = a.d;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < (long long)d.n_t * d.V;
    const int lane = threadIdx.x & 31;
    const int t = live ? (int)(idx / d.V) : 0, v = live ? (int)(idx - (long long)t * d.V) : 0;
    VertUp u;
    if (live) vertex_upstream(a, t, v, u);
    for (int k = 0; k < d.K; ++k) {
        float g[17];
        int key = -1;                                            // (timestamp, node) row of the gradient tables
        if (live) {
            const int n = d.nbr_idx[(size_t)v * d.K + k];
            const float w = d.nbr_w[(size_t)v * d.K + k];
            key = t * d.M + n;
            const float4 q = ldq(d.node_rot + (size_t)key * 4);
            incidence_gradient(d.method, w, u, ld3(d.node_trans + (size_t)key * 3), q, so3_log_bwd_coef(q),
                               d.method != 1 ? d.node_scale + (size_t)key * 9 : nullptr, g);
        } else {
#pragma unroll
            for (int j = 0; j < 17; ++j) g[j] = 0.f;
        }
        unsigned int remaining = __ballot_sync(0xffffffffu, key >= 0);
        while (remaining) {
            const int kk = __shfl_sync(0xffffffffu, key, __ffs((int)remaining) - 1);
            const bool mine = key == kk;
            remaining &= ~__ballot_sync(0xffffffffu, mine);
            float tot = 0.f;                                     // lane j keeps the warp total of component j
#pragma unroll
            for (int j = 0; j < 17; ++j) {
                float x = mine ? g[j] : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                if (lane == j) tot = x;
            }
            if (lane < 17 && tot != 0.f) {
                float* dst;
                if (lane < 3) dst = a.dn_trans + (size_t)kk * 3 + lane;
                else if (lane < 7) dst = a.dn_rot + (size_t)kk * 4 + (lane - 3);
                else if (lane < 16) dst = a.dn_scale + (size_t)kk * 9 + (lane - 7);
                else dst = a.dn_opac + kk;
                atomicAdd(dst, tot);
            }
        }
    }
}

// Incidence lists of the control nodes: inc_ptr [M+1], inc [V*K] = the flat indices e = v*K + k with nbr_idx[e] == n,
// ascending inside every node (deterministic summation order).  Start-up only (the graph is fixed after
// dynamic_sugar.py:745-861): a histogram, a one-CTA scan and one warp per node compacting in order.
__global__ void __launch_bounds__(DM4D_BLOCK) incidence_count_kernel(const int32_t* nbr_idx, int n, int M, int32_t* count, int32_t* bad) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int node = nbr_idx[e];
    if (node < 0 || node >= M) { *bad = 1; return; }
    atomicAdd(count + node, 1);
}
__global__ void __launch_bounds__(1024) incidence_scan_kernel(const int32_t* count, int M, int32_t* inc_ptr) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b = 0; b < M; b += 1024) {
        const int i = b + threadIdx.x;
        const int32_t c = i < M ? count[i] : 0;
        int32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += y; }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int32_t w = warp_tot[threadIdx.x], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, wi, o); if (threadIdx.x >= o) wi += y; }
            warp_tot[threadIdx.x] = wi - w;
        }
        __syncthreads();
        const int32_t excl = carry + warp_tot[threadIdx.x >> 5] + incl - c;
        if (i < M) inc_ptr[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) inc_ptr[M] = carry;
}
// Many short lists (vertex -> face corners: ~6 entries each, V lists): scatter with per-list cursors, then every list is
// sorted by one thread (insertion sort) so the order — and with it every floating-point sum over a list — is reproducible.
__global__ void __launch_bounds__(DM4D_BLOCK) incidence_scatter_kernel(const int32_t* idx, int n, int M, const int32_t* inc_ptr,
                                                                       int32_t* cursor, int32_t* inc) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int node = idx[e];
    if (node < 0 || node >= M) return;
    inc[inc_ptr[node] + atomicAdd(cursor + node, 1)] = e;
}
__global__ void __launch_bounds__(DM4D_BLOCK) incidence_sort_lists_kernel(int M, const int32_t* inc_ptr, int32_t* inc) {
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= M) return;
    const int lo = inc_ptr[node], hi = inc_ptr[node + 1];
    for (int i = lo + 1; i < hi; ++i) {
        const int32_t key = inc[i];
        int j = i - 1;
        while (j >= lo && inc[j] > key) { inc[j + 1] = inc[j]; --j; }
        inc[j + 1] = key;
    }
}

__global__ void __launch_bounds__(DM4D_BLOCK) incidence_fill_kernel(const int32_t* nbr_idx, int n, int M, const int32_t* inc_ptr, int32_t* inc) {
    const int node = blockIdx.x * (DM4D_BLOCK / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (node >= M) return;
    int pos = inc_ptr[node];
    for (int b = 0; b < n; b += 32) {
        const int e = b + lane;
        const bool hit = e < n && nbr_idx[e] == node;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) inc[pos + __popc(m & ((1u << lane) - 1u))] = e;
        pos += __popc(m);
    }
}

// ------------------------------------------------------------------------------------------------
// Gaussian stage: one thread per (timestamp, face)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 wxyz_to_xyzw(float4 q) { return make_float4(q.y, q.z, q.w, q.x); }
__device__ __forceinline__ float4 xyzw_to_wxyz(float4 q) { return make_float4(q.w, q.x, q.y, q.z); }

// Per-Gaussian arrays are [n_t*F*g, k] with k = 3 or 4: a thread owns the g*k consecutive floats of its face, so direct
// accesses are stride-(g*k) scalars (32 sectors per request).  Instead every warp moves its 32 faces' block (32*g*k
// contiguous floats) between global memory and a shared staging buffer with coalesced (16-byte when aligned) accesses.
constexpr int STAGE_FLOATS = 32 * 6 * 4;

__device__ __forceinline__ void warp_store_block(float* __restrict__ dst, const float* st, int n, int lane) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (n & 3) == 0) {
        float4* d4 = reinterpret_cast<float4*>(dst);
        const float4* s4 = reinterpret_cast<const float4*>(st);
        for (int i = lane; i < n / 4; i += 32) d4[i] = s4[i];
    } else {
        for (int i = lane; i < n; i += 32) dst[i] = st[i];
    }
}
__device__ __forceinline__ void warp_load_block(float* st, const float* __restrict__ src, int n, int lane) {
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (n & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(st);
        for (int i = lane; i < n / 4; i += 32) d4[i] = s4[i];
    } else {
        for (int i = lane; i < n; i += 32) st[i] = src[i];
    }
}

__global__ void __launch_bounds__(DM4D_BLOCK, DM4D_SKIN_GF_BLOCKS) skin_gaussian_forward_kernel(SkinK a) {
    __shared__ __align__(16) float stage_all[DM4D_BLOCK / 32][STAGE_FLOATS];
    const dm4d_skin_desc& d = a.d;
    const int lane = threadIdx.x & 31, g = d.g;
    float* st = stage_all[threadIdx.x >> 5];
    const long long total = (long long)d.n_t * d.F;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long wfirst = idx - lane;
    if (wfirst >= total) return;
    const int cnt = (int)min(32ll, total - wfirst);
    const bool live = idx < total;
    f3 x0 = mk3(0, 0, 0), x1 = x0, x2 = x0, L0 = x0, L1 = x0, L2 = x0, nrm = x0;
    int f = 0;
    if (live) {
        const int t = (int)(idx / d.F);
        f = (int)(idx - (long long)t * d.F);
        const int i0 = d.faces[(size_t)f * 3], i1 = d.faces[(size_t)f * 3 + 1], i2 = d.faces[(size_t)f * 3 + 2];
        const size_t vb = (size_t)t * d.V;
        x0 = ld3(a.verts + (vb + i0) * 3); x1 = ld3(a.verts + (vb + i1) * 3); x2 = ld3(a.verts + (vb + i2) * 3);
        L0 = so3_log(ldq(a.vert_rot + (vb + i0) * 4));
        L1 = so3_log(ldq(a.vert_rot + (vb + i1) * 4));
        L2 = so3_log(ldq(a.vert_rot + (vb + i2) * 4));
        nrm = cross(x1 - x0, x2 - x0);
        nrm = (1.f / fmaxf(sqrtf(dot(nrm, nrm)), 1e-6f)) * nrm;
        nrm = (1.f / fmaxf(sqrtf(dot(nrm, nrm)), 1e-12f)) * nrm;
    }
    // rest quaternions of the warp's faces: [F*g,4], contiguous per warp unless the warp straddles two timestamps
    const bool one_t = (wfirst / d.F) == ((wfirst + cnt - 1) / d.F);
    if (one_t) {
        warp_load_block(st, d.rest_quat + (size_t)(wfirst % d.F) * g * 4, cnt * g * 4, lane);
        __syncwarp();
    }
    float4 rest[6];
    if (live) {
#pragma unroll
        for (int j = 0; j < 6; ++j)
            if (j < g) rest[j] = wxyz_to_xyzw(one_t ? *reinterpret_cast<const float4*>(st + (lane * g + j) * 4)
                                                     : ldq(d.rest_quat + ((size_t)f * g + j) * 4));
    }
    __syncwarp();
    // rotations
    if (live) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (j >= g) break;
            const float b0 = d.bary[j * 3], b1 = d.bary[j * 3 + 1], b2 = d.bary[j * 3 + 2];
            const float4 dq = so3_exp(b0 * L0 + b1 * L1 + b2 * L2);
            float4 u = xyzw_to_wxyz(qmul(dq, rest[j]));
            const float inv = 1.f / fmaxf(sqrtf(dot4(u, u)), 1e-12f);
            *reinterpret_cast<float4*>(st + (lane * g + j) * 4) = inv * u;
        }
    }
    __syncwarp();
    warp_store_block(a.rots + (size_t)wfirst * g * 4, st, cnt * g * 4, lane);
    __syncwarp();
    // means
    if (live) {
        for (int j = 0; j < g; ++j) {
            const float b0 = d.bary[j * 3], b1 = d.bary[j * 3 + 1], b2 = d.bary[j * 3 + 2];
            st3(st + (lane * g + j) * 3, b0 * x0 + b1 * x1 + b2 * x2);
        }
    }
    __syncwarp();
    warp_store_block(a.means + (size_t)wfirst * g * 3, st, cnt * g * 3, lane);
    if (a.normals) {
        __syncwarp();
        if (live)
            for (int j = 0; j < g; ++j) st3(st + (lane * g + j) * 3, nrm);
        __syncwarp();
        warp_store_block(a.normals + (size_t)wfirst * g * 3, st, cnt * g * 3, lane);
    }
}

__global__ void __launch_bounds__(DM4D_BLOCK, DM4D_SKIN_GB_BLOCKS) skin_gaussian_backward_kernel(SkinBwdK a) {
    __shared__ __align__(16) float stage_all[DM4D_BLOCK / 32][STAGE_FLOATS];
    const dm4d_skin_desc& d = a.d;
    const int lane = threadIdx.x & 31, g = d.g;
    float* st = stage_all[threadIdx.x >> 5];
    const long long total = (long long)d.n_t * d.F;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long wfirst = idx - lane;
    if (wfirst >= total) return;
    const int cnt = (int)min(32ll, total - wfirst);
    const bool live = idx < total;
    const int t = live ? (int)(idx / d.F) : 0, f = live ? (int)(idx - (long long)t * d.F) : 0;
    int vi[3] = {0, 0, 0};
    const size_t vb = (size_t)t * d.V;
    f3 x[3], L[3];
    float4 r[3];
    LogBwdCoef lc[3];                 // log and its derivative share the atan
    if (live) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            vi[k] = d.faces[(size_t)f * 3 + k];
            x[k] = ld3(a.verts + (vb + vi[k]) * 3);
            r[k] = ldq(a.vert_rot + (vb + vi[k]) * 4);
            lc[k] = so3_log_bwd_coef(r[k]);
            L[k] = lc[k].f * vec(r[k]);
        }
    }
    f3 dx[3] = {mk3(0, 0, 0), mk3(0, 0, 0), mk3(0, 0, 0)};
    f3 dL[3] = {mk3(0, 0, 0), mk3(0, 0, 0), mk3(0, 0, 0)};
    f3 dn_sum = mk3(0, 0, 0);
    if (a.g_means) {
        warp_load_block(st, a.g_means + (size_t)wfirst * g * 3, cnt * g * 3, lane);
        __syncwarp();
        if (live)
            for (int j = 0; j < g; ++j) {
                const f3 gm = ld3(st + (lane * g + j) * 3);
#pragma unroll
                for (int k = 0; k < 3; ++k) dx[k] = dx[k] + d.bary[j * 3 + k] * gm;
            }
        __syncwarp();
    }
    if (a.g_normals) {
        warp_load_block(st, a.g_normals + (size_t)wfirst * g * 3, cnt * g * 3, lane);
        __syncwarp();
        if (live)
            for (int j = 0; j < g; ++j) dn_sum = dn_sum + ld3(st + (lane * g + j) * 3);
        __syncwarp();
    }
    if (a.g_rots) {
        warp_load_block(st, a.g_rots + (size_t)wfirst * g * 4, cnt * g * 4, lane);
        __syncwarp();
        if (live)
            for (int j = 0; j < g; ++j) {
                const float b[3] = {d.bary[j * 3], d.bary[j * 3 + 1], d.bary[j * 3 + 2]};
                const f3 xi = b[0] * L[0] + b[1] * L[1] + b[2] * L[2];
                ExpBwdCoef ec;
                const float4 dq = so3_exp_coef(xi, ec);
                const float4 rest = wxyz_to_xyzw(ldq(d.rest_quat + ((size_t)f * g + j) * 4));
                const float4 u = qmul(dq, rest);                         // xyzw
                const float nu = fmaxf(sqrtf(dot4(u, u)), 1e-12f), inv = 1.f / nu;
                const float4 qn = inv * u;
                const float4 gq = wxyz_to_xyzw(*reinterpret_cast<const float4*>(st + (lane * g + j) * 4));  // incoming gradient is wxyz
                const float4 du = inv * (gq + (-dot4(qn, gq)) * qn);
                const float4 ddq = qmul(du, qconj(rest));
                const f3 dxi = so3_exp_bwd_apply(ec, xi, ddq);
#pragma unroll
                for (int k = 0; k < 3; ++k) dL[k] = dL[k] + b[k] * dxi;
            }
    }
    if (live && a.g_normals) {
        const f3 e1 = x[1] - x[0], e2 = x[2] - x[0];
        const f3 c = cross(e1, e2);
        const float len = sqrtf(dot(c, c));
        // regular branch: n = c/|c|; clamped branch: n = c / 1e-6 (then re-normalised) -> gradient of the unnormalised direction
        f3 dc;
        if (len > 1e-6f) {
            const f3 n = (1.f / len) * c;
            dc = (1.f / len) * (dn_sum - dot(n, dn_sum) * n);
        } else {
            dc = 1e6f * dn_sum;
        }
        const f3 de1 = cross(e2, dc), de2 = cross(dc, e1);
        dx[0] = dx[0] - (de1 + de2);
        dx[1] = dx[1] + de1;
        dx[2] = dx[2] + de2;
    }
    if (a.corner) {
        // vertex-centric mode: the warp's 32 x 3 corner records (8 floats each = exactly the staging buffer) leave as one
        // contiguous block; the vertex stage gathers them
        __syncwarp();
        if (live) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 dr = a.g_rots ? so3_log_bwd_apply(lc[k], r[k], dL[k]) : make_float4(0, 0, 0, 0);
                float4* rec = reinterpret_cast<float4*>(st + (lane * 3 + k) * 8);
                rec[0] = make_float4(dx[k].x, dx[k].y, dx[k].z, 0.f);
                rec[1] = dr;
            }
        }
        __syncwarp();
        warp_store_block(a.corner + (size_t)wfirst * 24, st, cnt * 24, lane);
        return;
    }
    if (!live) return;
    // one 16-byte vector reduction per corner and quantity (the scratch rows are padded to float4)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        red_add_v4(a.dverts + (vb + vi[k]) * 4, dx[k].x, dx[k].y, dx[k].z, 0.f);
        if (a.g_rots) {
            const float4 dr = so3_log_bwd_apply(lc[k], r[k], dL[k]);
            red_add_v4(a.dvert_rot + (vb + vi[k]) * 4, dr.x, dr.y, dr.z, dr.w);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// rest-pose frames (sugar.py:490-526)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ f3 normalize12(f3 v) { return (1.f / fmaxf(sqrtf(dot(v, v)), 1e-12f)) * v; }

__global__ void __launch_bounds__(DM4D_BLOCK) sugar_rest_frames_kernel(const float* verts, const int32_t* faces,
                                                                       const float* complex_rot, int F, int g,
                                                                       float* quats, float* normals) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const f3 x0 = ld3(verts + (size_t)faces[f * 3] * 3), x1 = ld3(verts + (size_t)faces[f * 3 + 1] * 3),
             x2 = ld3(verts + (size_t)faces[f * 3 + 2] * 3);
    f3 R0 = cross(x1 - x0, x2 - x0);
    R0 = (1.f / fmaxf(sqrtf(dot(R0, R0)), 1e-6f)) * R0;
    R0 = normalize12(R0);
    const f3 b1 = normalize12(x0 - x1);
    const f3 b2 = normalize12(cross(R0, b1));
    for (int j = 0; j < g; ++j) {
        const size_t gi = (size_t)f * g + j;
        if (normals) st3(normals + gi * 3, R0);
        if (!quats) continue;
        float cr = complex_rot[gi * 2], ci = complex_rot[gi * 2 + 1];
        const float inv = 1.f / fmaxf(sqrtf(cr * cr + ci * ci), 1e-12f);
        cr *= inv; ci *= inv;
        const f3 R1 = cr * b1 + ci * b2;
        const f3 R2 = (-ci) * b1 + cr * b2;
        // matrix with COLUMNS (R0, R1, R2): m[r][c]
        const float m00 = R0.x, m01 = R1.x, m02 = R2.x, m10 = R0.y, m11 = R1.y, m12 = R2.y, m20 = R0.z, m21 = R1.z, m22 = R2.z;
        // pytorch3d matrix_to_quaternion (wxyz): candidate table, argmax of q_abs (first max wins)
        const float qa[4] = {sqrtf(fmaxf(0.f, 1.f + m00 + m11 + m22)), sqrtf(fmaxf(0.f, 1.f + m00 - m11 - m22)),
                             sqrtf(fmaxf(0.f, 1.f - m00 + m11 - m22)), sqrtf(fmaxf(0.f, 1.f - m00 - m11 + m22))};
        int best = 0;
        for (int k = 1; k < 4; ++k) if (qa[k] > qa[best]) best = k;
        float4 q;   // (w, x, y, z)
        if (best == 0) q = make_float4(qa[0] * qa[0], m21 - m12, m02 - m20, m10 - m01);
        else if (best == 1) q = make_float4(m21 - m12, qa[1] * qa[1], m10 + m01, m02 + m20);
        else if (best == 2) q = make_float4(m02 - m20, m10 + m01, qa[2] * qa[2], m12 + m21);
        else q = make_float4(m10 - m01, m20 + m02, m21 + m12, qa[3] * qa[3]);
        q = (1.f / (2.f * fmaxf(qa[best], 0.1f))) * q;
        q = (1.f / fmaxf(sqrtf(dot4(q, q)), 1e-12f)) * q;
        *reinterpret_cast<float4*>(quats + gi * 4) = q;
    }
}

// Backward of sugar_rest_frames_kernel: exact Euclidean gradient of (quaternions, normals) w.r.t. the mesh
// vertices and the per-Gaussian complex in-plane rotation (needed in the static stage, where both are learnable:
// sugar.py:333-376).  The quaternion branch chosen by matrix_to_quaternion is treated as locally constant.
__global__ void __launch_bounds__(DM4D_BLOCK) sugar_rest_frames_backward_kernel(const float* verts, const int32_t* faces,
                                                                                const float* complex_rot, int F, int g,
                                                                                const float* g_quats, const float* g_normals,
                                                                                float* dverts, float* dcomplex) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int i0 = faces[f * 3], i1 = faces[f * 3 + 1], i2 = faces[f * 3 + 2];
    const f3 x0 = ld3(verts + (size_t)i0 * 3), x1 = ld3(verts + (size_t)i1 * 3), x2 = ld3(verts + (size_t)i2 * 3);
    const f3 e1 = x1 - x0, e2 = x2 - x0;
    const f3 c = cross(e1, e2);
    const float len = sqrtf(dot(c, c));
    const float lenc = fmaxf(len, 1e-6f);
    f3 R0 = (1.f / lenc) * c;
    R0 = normalize12(R0);
    const f3 d = x0 - x1;
    const float dl = fmaxf(sqrtf(dot(d, d)), 1e-12f);
    const f3 b1 = (1.f / dl) * d;
    const f3 p = cross(R0, b1);
    const float pl = fmaxf(sqrtf(dot(p, p)), 1e-12f);
    const f3 b2 = (1.f / pl) * p;

    f3 dR0 = mk3(0, 0, 0), db1 = mk3(0, 0, 0), db2 = mk3(0, 0, 0);
    for (int j = 0; j < g; ++j) {
        const size_t gi = (size_t)f * g + j;
        if (g_normals) dR0 = dR0 + ld3(g_normals + gi * 3);
        if (!g_quats) { if (dcomplex) { dcomplex[gi * 2] = 0.f; dcomplex[gi * 2 + 1] = 0.f; } continue; }
        const float cr0 = complex_rot[gi * 2], ci0 = complex_rot[gi * 2 + 1];
        const float sn = fmaxf(sqrtf(cr0 * cr0 + ci0 * ci0), 1e-12f);
        const float cr = cr0 / sn, ci = ci0 / sn;
        const f3 R1 = cr * b1 + ci * b2;
        const f3 R2 = (-ci) * b1 + cr * b2;
        const float m[3][3] = {{R0.x, R1.x, R2.x}, {R0.y, R1.y, R2.y}, {R0.z, R1.z, R2.z}};
        const float lin[4] = {1.f + m[0][0] + m[1][1] + m[2][2], 1.f + m[0][0] - m[1][1] - m[2][2],
                              1.f - m[0][0] + m[1][1] - m[2][2], 1.f - m[0][0] - m[1][1] + m[2][2]};
        float qa[4];
        int best = 0;
        for (int k = 0; k < 4; ++k) { qa[k] = sqrtf(fmaxf(0.f, lin[k])); if (qa[k] > qa[best]) best = k; }
        float cand[4];
        if (best == 0) { cand[0] = lin[0]; cand[1] = m[2][1] - m[1][2]; cand[2] = m[0][2] - m[2][0]; cand[3] = m[1][0] - m[0][1]; }
        else if (best == 1) { cand[0] = m[2][1] - m[1][2]; cand[1] = lin[1]; cand[2] = m[1][0] + m[0][1]; cand[3] = m[0][2] + m[2][0]; }
        else if (best == 2) { cand[0] = m[0][2] - m[2][0]; cand[1] = m[1][0] + m[0][1]; cand[2] = lin[2]; cand[3] = m[1][2] + m[2][1]; }
        else { cand[0] = m[1][0] - m[0][1]; cand[1] = m[2][0] + m[0][2]; cand[2] = m[2][1] + m[1][2]; cand[3] = lin[3]; }
        const float D = 2.f * fmaxf(qa[best], 0.1f);
        float u[4], un = 0.f;
        for (int k = 0; k < 4; ++k) { u[k] = cand[k] / D; un += u[k] * u[k]; }
        un = fmaxf(sqrtf(un), 1e-12f);
        const float4 gq = ldq(g_quats + gi * 4);
        const float gqa[4] = {gq.x, gq.y, gq.z, gq.w};
        float q[4], qg = 0.f;
        for (int k = 0; k < 4; ++k) { q[k] = u[k] / un; qg += q[k] * gqa[k]; }
        float du[4], dudotu = 0.f;
        for (int k = 0; k < 4; ++k) { du[k] = (gqa[k] - q[k] * qg) / un; dudotu += du[k] * u[k]; }
        float dcand[4];
        for (int k = 0; k < 4; ++k) dcand[k] = du[k] / D;
        float dlin = (qa[best] > 0.1f) ? (-(dudotu / D) * 2.f) / (2.f * qa[best]) : 0.f;   // through D = 2 sqrt(lin)
        float dm[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        if (best == 0) { dlin += dcand[0]; dm[2][1] += dcand[1]; dm[1][2] -= dcand[1]; dm[0][2] += dcand[2]; dm[2][0] -= dcand[2]; dm[1][0] += dcand[3]; dm[0][1] -= dcand[3]; }
        else if (best == 1) { dm[2][1] += dcand[0]; dm[1][2] -= dcand[0]; dlin += dcand[1]; dm[1][0] += dcand[2]; dm[0][1] += dcand[2]; dm[0][2] += dcand[3]; dm[2][0] += dcand[3]; }
        else if (best == 2) { dm[0][2] += dcand[0]; dm[2][0] -= dcand[0]; dm[1][0] += dcand[1]; dm[0][1] += dcand[1]; dlin += dcand[2]; dm[1][2] += dcand[3]; dm[2][1] += dcand[3]; }
        else { dm[1][0] += dcand[0]; dm[0][1] -= dcand[0]; dm[2][0] += dcand[1]; dm[0][2] += dcand[1]; dm[2][1] += dcand[2]; dm[1][2] += dcand[2]; dlin += dcand[3]; }
        const float sg[4][3] = {{1, 1, 1}, {1, -1, -1}, {-1, 1, -1}, {-1, -1, 1}};
        for (int k = 0; k < 3; ++k) dm[k][k] += sg[best][k] * dlin;
        const f3 gR0 = mk3(dm[0][0], dm[1][0], dm[2][0]), gR1 = mk3(dm[0][1], dm[1][1], dm[2][1]), gR2 = mk3(dm[0][2], dm[1][2], dm[2][2]);
        dR0 = dR0 + gR0;
        const float dcr = dot(gR1, b1) + dot(gR2, b2), dci = dot(gR1, b2) - dot(gR2, b1);
        db1 = db1 + cr * gR1 - ci * gR2;
        db2 = db2 + ci * gR1 + cr * gR2;
        const float cd = cr * dcr + ci * dci;
        dcomplex[gi * 2] = (dcr - cr * cd) / sn;
        dcomplex[gi * 2 + 1] = (dci - ci * cd) / sn;
    }
    // b2 = p/|p|, p = R0 x b1
    const f3 dp = (1.f / pl) * (db2 - dot(b2, db2) * b2);
    dR0 = dR0 + cross(b1, dp);
    db1 = db1 + cross(dp, R0);
    // b1 = d/|d|, d = x0 - x1
    const f3 dd = (1.f / dl) * (db1 - dot(b1, db1) * b1);
    f3 dx0 = dd, dx1 = mk3(0, 0, 0) - dd, dx2 = mk3(0, 0, 0);
    // R0 = c/|c|, c = e1 x e2
    const f3 dc = len > 1e-6f ? (1.f / len) * (dR0 - dot(R0, dR0) * R0) : 1e6f * dR0;
    const f3 de1 = cross(e2, dc), de2 = cross(dc, e1);
    dx0 = dx0 - (de1 + de2);
    dx1 = dx1 + de1;
    dx2 = dx2 + de2;
    atomicAdd(dverts + (size_t)i0 * 3 + 0, dx0.x); atomicAdd(dverts + (size_t)i0 * 3 + 1, dx0.y); atomicAdd(dverts + (size_t)i0 * 3 + 2, dx0.z);
    atomicAdd(dverts + (size_t)i1 * 3 + 0, dx1.x); atomicAdd(dverts + (size_t)i1 * 3 + 1, dx1.y); atomicAdd(dverts + (size_t)i1 * 3 + 2, dx1.z);
    atomicAdd(dverts + (size_t)i2 * 3 + 0, dx2.x); atomicAdd(dverts + (size_t)i2 * 3 + 1, dx2.y); atomicAdd(dverts + (size_t)i2 * 3 + 2, dx2.z);
}

int check_desc(const dm4d_skin_desc* d) {
    if (!d) { dm4d_set_error("skin desc is NULL"); return DM4D_EINVAL; }
    if (d->n_t <= 0 || d->V <= 0 || d->F <= 0 || d->M <= 0 || d->K <= 0 || d->g <= 0 || d->g > 6) {
        dm4d_set_error("skin: bad sizes n_t=%d V=%d F=%d M=%d K=%d g=%d", d->n_t, d->V, d->F, d->M, d->K, d->g);
        return DM4D_EINVAL;
    }
    if (d->method < 0 || d->method > 2) { dm4d_set_error("skin: method must be 0 (lbs), 1 (dqs) or 2 (hybrid)"); return DM4D_EINVAL; }
    if (!d->rest_verts || !d->faces || !d->nbr_idx || !d->nbr_w || !d->bary || !d->rest_quat || !d->node_trans ||
        !d->node_rot || (d->method != 1 && !d->node_scale) || (d->method == 2 && !d->node_opacity)) {
        dm4d_set_error("skin: NULL input pointer");
        return DM4D_EINVAL;
    }
    return DM4D_OK;
}

}  // namespace

extern "C" int dm4d_skin_forward(const dm4d_skin_desc* d, float* verts, float* vert_rot, float* means3D,
                                 float* rotations, float* normals, void* stream) {
    int rc = check_desc(d);
    if (rc) return rc;
    if (!verts || !vert_rot || !means3D || !rotations) { dm4d_set_error("skin: NULL output pointer"); return DM4D_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    SkinK a;
    a.d = *d; a.P = d->F * d->g;
    a.verts = verts; a.vert_rot = vert_rot; a.means = means3D; a.rots = rotations; a.normals = normals;
    const long long nv = (long long)d->n_t * d->V, nf = (long long)d->n_t * d->F;
    if (d->node_scratch) {
        if (reinterpret_cast<uintptr_t>(d->node_scratch) & 15) { dm4d_set_error("skin: node_scratch must be 16-byte aligned"); return DM4D_EINVAL; }
        skin_node_pre_kernel<<<(unsigned)((d->n_t * d->M + DM4D_BLOCK - 1) / DM4D_BLOCK), DM4D_BLOCK, 0, s>>>(*d, reinterpret_cast<float4*>(d->node_scratch));
    }
    { KernelTimer kt(DM4D_K_SKIN_VERT_FWD, s); skin_vertex_forward_kernel<<<(unsigned)((nv + DM4D_BLOCK - 1) / DM4D_BLOCK), DM4D_BLOCK, 0, s>>>(a); }
    DM4D_CUDA_CHECK(cudaGetLastError());
    { KernelTimer kt(DM4D_K_SKIN_GAUSS_FWD, s); skin_gaussian_forward_kernel<<<(unsigned)((nf + DM4D_BLOCK - 1) / DM4D_BLOCK), DM4D_BLOCK, 0, s>>>(a); }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

extern "C" int dm4d_skin_backward(const dm4d_skin_desc* d, const float* verts, const float* vert_rot,
                                  const float* dL_dmeans3D, const float* dL_drotations, const float* dL_dnormals,
                                  const float* dL_dverts_in, const float* dL_dvert_rot_in, float* dverts,
                                  float* dvert_rot, float* dL_dnode_trans, float* dL_dnode_rot,
                                  float* dL_dnode_scale, float* dL_dnode_opacity, void* stream) {
    int rc = check_desc(d);
    if (rc) return rc;
    if (!verts || !vert_rot || !dverts || !dvert_rot || !dL_dnode_trans || !dL_dnode_rot || !dL_dnode_scale ||
        !dL_dnode_opacity) {
        dm4d_set_error("skin backward: NULL pointer");
        return DM4D_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t nv = (size_t)d->n_t * d->V, nf = (size_t)d->n_t * d->F, nm = (size_t)d->n_t * d->M;
    if ((reinterpret_cast<uintptr_t>(dverts) | reinterpret_cast<uintptr_t>(dvert_rot)) & 15) {
        dm4d_set_error("skin backward: dverts / dvert_rot must be 16-byte aligned");
        return DM4D_EINVAL;
    }
    const int n_vc = (d->vert_inc_ptr != nullptr) + (d->vert_inc != nullptr) + (d->corner_scratch != nullptr);
    if (n_vc != 0 && n_vc != 3) { dm4d_set_error("skin backward: vert_inc_ptr, vert_inc and corner_scratch go together"); return DM4D_EINVAL; }
    if (n_vc == 3 && (reinterpret_cast<uintptr_t>(d->corner_scratch) & 15)) { dm4d_set_error("skin backward: corner_scratch must be 16-byte aligned"); return DM4D_EINVAL; }
    const bool any_gauss_grad = dL_dmeans3D || dL_drotations || dL_dnormals;
    if (n_vc == 0) {
        DM4D_CUDA_CHECK(cudaMemsetAsync(dverts, 0, nv * 4 * sizeof(float), s));
        DM4D_CUDA_CHECK(cudaMemsetAsync(dvert_rot, 0, nv * 4 * sizeof(float), s));
    } else if (!any_gauss_grad) {
        DM4D_CUDA_CHECK(cudaMemsetAsync(d->corner_scratch, 0, nf * 3 * 8 * sizeof(float), s));   // the gather must read zeros
    }
    SkinBwdK a;
    a.d = *d; a.P = d->F * d->g;
    a.verts = verts; a.vert_rot = vert_rot;
    a.g_means = dL_dmeans3D; a.g_rots = dL_drotations; a.g_normals = dL_dnormals;
    a.dverts = dverts; a.dvert_rot = dvert_rot;
    a.dverts_in = dL_dverts_in; a.dvert_rot_in = dL_dvert_rot_in;
    a.vinc_ptr = d->vert_inc_ptr; a.vinc = d->vert_inc; a.corner = d->corner_scratch;
    a.dn_trans = dL_dnode_trans; a.dn_rot = dL_dnode_rot; a.dn_scale = dL_dnode_scale; a.dn_opac = dL_dnode_opacity;
    if (d->node_scratch) {           // same pre-pass as the forward (the caller may have reused the scratch in between)
        if (reinterpret_cast<uintptr_t>(d->node_scratch) & 15) { dm4d_set_error("skin: node_scratch must be 16-byte aligned"); return DM4D_EINVAL; }
        skin_node_pre_kernel<<<(unsigned)((d->n_t * d->M + DM4D_BLOCK - 1) / DM4D_BLOCK), DM4D_BLOCK, 0, s>>>(*d, reinterpret_cast<float4*>(d->node_scratch));
    }
    if (dL_dmeans3D || dL_drotations || dL_dnormals) {
        KernelTimer kt(DM4D_K_SKIN_GAUSS_BWD, s);
        skin_gaussian_backward_kernel<<<(unsigned)((nf + DM4D_BLOCK - 1) / DM4D_BLOCK), DM4D_BLOCK, 0, s>>>(a);
    }
    DM4D_CUDA_CHECK(cudaGetLastError());
    // enough CTAs to fill the GPU: the node lists are split when nodes x timestamps alone would not
    if ((d->node_inc_ptr != nullptr) != (d->node_inc != nullptr) || (d->node_inc && !d->vert_scratch)) {
        dm4d_set_error("skin backward: node_inc_ptr, node_inc and vert_scratch go together");
        return DM4D_EINVAL;
    }
    const int splits = d->node_inc ? std::max(1, std::min(16, (148 * 8 + d->M * d->n_t - 1) / (d->M * d->n_t))) : 0;
    if (splits != 1) {
        DM4D_CUDA_CHECK(cudaMemsetAsync(dL_dnode_trans, 0, nm * 3 * sizeof(float), s));
        DM4D_CUDA_CHECK(cudaMemsetAsync(dL_dnode_rot, 0, nm * 4 * sizeof(float), s));
        DM4D_CUDA_CHECK(cudaMemsetAsync(dL_dnode_scale, 0, nm * 9 * sizeof(float), s));
        DM4D_CUDA_CHECK(cudaMemsetAsync(dL_dnode_opacity, 0, nm * sizeof(float), s));
    }
    {
        KernelTimer kt(DM4D_K_SKIN_VERT_BWD, s);
        if (splits) {
            if (d->n_t > 65535) { dm4d_set_error("skin backward: n_t > 65535"); return DM4D_EINVAL; }
            float4* up = reinterpret_cast<float4*>(d->vert_scratch);
            skin_vertex_upstream_kernel<<<(unsigned)((nv + DM4D_BLOCK - 1) / DM4D_BLOCK), DM4D_BLOCK, 0, s>>>(a, up);
            skin_node_backward_kernel<<<dim3((unsigned)d->M, (unsigned)d->n_t, (unsigned)splits), DM4D_BLOCK, 0, s>>>(a, up, d->node_inc_ptr, d->node_inc);
        } else {
            skin_vertex_backward_kernel<<<(unsigned)((nv + DM4D_BLOCK - 1) / DM4D_BLOCK), DM4D_BLOCK, 0, s>>>(a);
        }
    }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

extern "C" int dm4d_skin_node_incidence(const int32_t* nbr_idx, int32_t V, int32_t K, int32_t M, int32_t* inc_ptr,
                                        int32_t* inc, int32_t* scratch, void* stream) {
    if (!nbr_idx || !inc_ptr || !inc || !scratch || V <= 0 || K <= 0 || M <= 0 || (long long)V * K > 0x7fffffffLL) {
        dm4d_set_error("dm4d_skin_node_incidence: bad argument");
        return DM4D_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int n = V * K;
    DM4D_CUDA_CHECK(cudaMemsetAsync(scratch, 0, ((size_t)M + 1) * sizeof(int32_t), s));
    incidence_count_kernel<<<(n + DM4D_BLOCK - 1) / DM4D_BLOCK, DM4D_BLOCK, 0, s>>>(nbr_idx, n, M, scratch, scratch + M);
    incidence_scan_kernel<<<1, 1024, 0, s>>>(scratch, M, inc_ptr);
    if ((long long)n <= 64ll * M) {
        // many short lists (e.g. vertex -> face corners): scatter + per-list sort
        DM4D_CUDA_CHECK(cudaMemsetAsync(scratch, 0, (size_t)M * sizeof(int32_t), s));
        incidence_scatter_kernel<<<(n + DM4D_BLOCK - 1) / DM4D_BLOCK, DM4D_BLOCK, 0, s>>>(nbr_idx, n, M, inc_ptr, scratch, inc);
        incidence_sort_lists_kernel<<<(M + DM4D_BLOCK - 1) / DM4D_BLOCK, DM4D_BLOCK, 0, s>>>(M, inc_ptr, inc);
    } else {
        // few long lists (control nodes): one warp per list compacts the entries in order
        const int per = DM4D_BLOCK / 32;
        incidence_fill_kernel<<<(M + per - 1) / per, DM4D_BLOCK, 0, s>>>(nbr_idx, n, M, inc_ptr, inc);
    }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

extern "C" int dm4d_sugar_rest_frames(const float* verts, const int32_t* faces, const float* complex_rot, int32_t V,
                                      int32_t F, int32_t g, float* quaternions, float* normals, void* stream) {
    if (!verts || !faces || V <= 0 || F <= 0 || g <= 0 || g > 6 || (quaternions && !complex_rot)) {
        dm4d_set_error("dm4d_sugar_rest_frames: bad argument");
        return DM4D_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    { KernelTimer kt(DM4D_K_REST_FRAMES, s); sugar_rest_frames_kernel<<<(unsigned)((F + DM4D_BLOCK - 1) / DM4D_BLOCK), DM4D_BLOCK, 0, s>>>(verts, faces, complex_rot, F, g, quaternions, normals); }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

extern "C" int dm4d_sugar_rest_frames_backward(const float* verts, const int32_t* faces, const float* complex_rot,
                                               int32_t V, int32_t F, int32_t g, const float* dL_dquaternions,
                                               const float* dL_dnormals, float* dL_dverts, float* dL_dcomplex_rot,
                                               void* stream) {
    if (!verts || !faces || V <= 0 || F <= 0 || g <= 0 || g > 6 || !dL_dverts || (dL_dquaternions && (!complex_rot || !dL_dcomplex_rot))) {
        dm4d_set_error("dm4d_sugar_rest_frames_backward: bad argument");
        return DM4D_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    DM4D_CUDA_CHECK(cudaMemsetAsync(dL_dverts, 0, (size_t)V * 3 * sizeof(float), s));
    { KernelTimer kt(DM4D_K_REST_FRAMES, s); sugar_rest_frames_backward_kernel<<<(unsigned)((F + DM4D_BLOCK - 1) / DM4D_BLOCK), DM4D_BLOCK, 0, s>>>(verts, faces, complex_rot, F, g, dL_dquaternions, dL_dnormals, dL_dverts, dL_dcomplex_rot); }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

// ------------------------------------------------------------------------------------------------
// ARAP energy on the deformed vertices
// ------------------------------------------------------------------------------------------------
// E_t = sum_i sum_{j in N(i)} w_ij || (x'_i - x'_j) - R(q_i) (x_i - x_j) ||^2   with the per-vertex rotations
// SUPPLIED by the deformation (ARAPCoach.compute_arap_energy with vert_rotations,
// custom/threestudio-dreammesh4d/utils/arap_utils.py:183-224; called per timestamp from
// system/sugar_4dgen.py:372-385).  One thread per (timestamp, vertex) walks its one-ring (CSR), and the same
// pass writes dE/dx' and dE/dq (xyzw, Euclidean), which feed dm4d_skin_backward's dL_dverts_in / dL_dvert_rot_in.
namespace {
__global__ void __launch_bounds__(DM4D_BLOCK) arap_energy_kernel(const float* rest_verts, const int32_t* row_ptr,
                                                                 const int32_t* col, const float* w, int n_t, int V,
                                                                 const float* verts, const float* vert_rot,
                                                                 float* energy, float* dverts, float* dvert_rot) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float e_local = 0.f;
    if (idx < (long long)n_t * V) {
        const int t = (int)(idx / V), i = (int)(idx - (long long)t * V);
        const size_t vb = (size_t)t * V;
        const f3 xi = ld3(rest_verts + (size_t)i * 3), xpi = ld3(verts + (vb + i) * 3);
        const float4 q = ldq(vert_rot + (vb + i) * 4);
        f3 gxi = mk3(0, 0, 0);
        float4 gq = make_float4(0, 0, 0, 0);
        for (int k = row_ptr[i]; k < row_ptr[i + 1]; ++k) {
            const int j = col[k];
            const float wij = w[k];
            const f3 e = xi - ld3(rest_verts + (size_t)j * 3);
            const f3 s = (xpi - ld3(verts + (vb + j) * 3)) - qact(q, e);
            e_local += wij * dot(s, s);
            const f3 g = (2.f * wij) * s;
            gxi = gxi + g;
            if (dverts) {
                float* pj = dverts + (vb + j) * 3;
                atomicAdd(pj + 0, -g.x); atomicAdd(pj + 1, -g.y); atomicAdd(pj + 2, -g.z);
            }
            float4 dq;
            qact_bwd(q, e, mk3(0, 0, 0) - g, dq);
            gq = gq + dq;
        }
        if (dverts) {
            float* pi = dverts + (vb + i) * 3;
            atomicAdd(pi + 0, gxi.x); atomicAdd(pi + 1, gxi.y); atomicAdd(pi + 2, gxi.z);
        }
        if (dvert_rot) *reinterpret_cast<float4*>(dvert_rot + (vb + i) * 4) = gq;
        // per-timestamp energy: warp-level pre-reduction when the whole warp is in range and in one timestamp
        const long long w_first = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31), w_last = w_first + 31;
        if (w_last < (long long)n_t * V && w_first / V == w_last / V) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e_local += __shfl_xor_sync(0xffffffffu, e_local, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(energy + t, e_local);
        } else {
            atomicAdd(energy + t, e_local);
        }
    }
}
}  // namespace

extern "C" int dm4d_arap_energy(const float* rest_verts, const int32_t* row_ptr, const int32_t* col, const float* weights,
                                int32_t n_t, int32_t V, const float* verts, const float* vert_rot, float* energy,
                                float* dE_dverts, float* dE_dvert_rot, void* stream) {
    if (!rest_verts || !row_ptr || !col || !weights || !verts || !vert_rot || !energy || n_t <= 0 || V <= 0) {
        dm4d_set_error("dm4d_arap_energy: bad argument");
        return DM4D_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t nv = (size_t)n_t * V;
    DM4D_CUDA_CHECK(cudaMemsetAsync(energy, 0, (size_t)n_t * sizeof(float), s));
    if (dE_dverts) DM4D_CUDA_CHECK(cudaMemsetAsync(dE_dverts, 0, nv * 3 * sizeof(float), s));
    { KernelTimer kt(DM4D_K_ARAP, s); arap_energy_kernel<<<(unsigned)((nv + DM4D_BLOCK - 1) / DM4D_BLOCK), DM4D_BLOCK, 0, s>>>(rest_verts, row_ptr, col, weights, n_t, V, verts, vert_rot, energy, dE_dverts, dE_dvert_rot); }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

// ------------------------------------------------------------------------------------------------
// mesh normal consistency on the deformed meshes
// ------------------------------------------------------------------------------------------------
// pytorch3d.loss.mesh_normal_consistency as called at custom/threestudio-dreammesh4d/system/sugar_4dgen.py:214-225
// (lambda_normal_consistency = 100): for every pair of faces sharing an edge (v0,v1) with opposite vertices a, b:
//   n0 = (v1-v0) x (a-v0),  n1 = (v1-v0) x (b-v0),  term = 1 - cos(n0, -n1);   loss = mean over pairs, mean over meshes.
// One thread per (timestamp, pair); the same pass writes d loss_t / d verts (scaled by 1/pairs).
namespace {
__global__ void __launch_bounds__(DM4D_BLOCK) normal_consistency_kernel(const int32_t* pairs, int n_pairs, int n_t, int V,
                                                                        const float* verts, float* loss, float* dverts) {
    // grid = (pair blocks, timestamps): a block belongs to ONE timestamp, so its loss terms are reduced in the
    // block and leave as one atomic (150k same-address atomics per timestamp serialised the old kernel: 1.85 ms)
    __shared__ float warp_part[DM4D_BLOCK / 32];
    const int t = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < n_pairs;
    float term = 0.f;
    if (live) {
    const int i0 = pairs[p * 4], i1 = pairs[p * 4 + 1], ia = pairs[p * 4 + 2], ib = pairs[p * 4 + 3];
    const size_t vb = (size_t)t * V;
    const f3 v0 = ld3(verts + (vb + i0) * 3), v1 = ld3(verts + (vb + i1) * 3), a = ld3(verts + (vb + ia) * 3),
             b = ld3(verts + (vb + ib) * 3);
    const f3 e = v1 - v0, ea = a - v0, eb = b - v0;
    const f3 n0 = cross(e, ea);
    const f3 m1 = mk3(0, 0, 0) - cross(e, eb);                 // -n1
    // torch.cosine_similarity: x.y / (max(|x|, eps) * max(|y|, eps)), eps = 1e-8
    const float l0 = sqrtf(dot(n0, n0)), l1 = sqrtf(dot(m1, m1));
    const float c0 = fmaxf(l0, 1e-8f), c1 = fmaxf(l1, 1e-8f);
    const float cosv = dot(n0, m1) / (c0 * c1);
    const float wgt = 1.0f / (float)n_pairs;
    term = (1.0f - cosv) * wgt;
    if (dverts) {
    // d(-cos)/dn0 = -(m1/(c0 c1) - cos * n0 / c0^2)   (norm clamp treated as inactive when l > eps)
    const f3 g0 = (-wgt) * ((1.f / (c0 * c1)) * m1 - ((l0 > 1e-8f ? cosv / (c0 * c0) : 0.f)) * n0);
    const f3 g1 = (-wgt) * ((1.f / (c0 * c1)) * n0 - ((l1 > 1e-8f ? cosv / (c1 * c1) : 0.f)) * m1);   // w.r.t. m1 = -n1
    // n0 = e x ea : de += ea x g0, dea += g0 x e ;  m1 = -(e x eb) : de -= eb x g1, deb -= g1 x e
    const f3 de = cross(ea, g0) - cross(eb, g1);
    const f3 dea = cross(g0, e);
    const f3 deb = mk3(0, 0, 0) - cross(g1, e);
    const f3 d0 = mk3(0, 0, 0) - (de + dea + deb);
    float* q0 = dverts + (vb + i0) * 3; float* q1 = dverts + (vb + i1) * 3;
    float* qa = dverts + (vb + ia) * 3; float* qb = dverts + (vb + ib) * 3;
    atomicAdd(q0, d0.x); atomicAdd(q0 + 1, d0.y); atomicAdd(q0 + 2, d0.z);
    atomicAdd(q1, de.x); atomicAdd(q1 + 1, de.y); atomicAdd(q1 + 2, de.z);
    atomicAdd(qa, dea.x); atomicAdd(qa + 1, dea.y); atomicAdd(qa + 2, dea.z);
    atomicAdd(qb, deb.x); atomicAdd(qb + 1, deb.y); atomicAdd(qb + 2, deb.z);
    }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = term;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < DM4D_BLOCK / 32; ++w) tot += warp_part[w];
        atomicAdd(loss + t, tot);
    }
}
}  // namespace

extern "C" int dm4d_mesh_normal_consistency(const int32_t* pairs, int32_t n_pairs, int32_t n_t, int32_t V,
                                            const float* verts, float* loss, float* dL_dverts, void* stream) {
    if (!pairs || !verts || !loss || n_pairs <= 0 || n_t <= 0 || V <= 0) {
        dm4d_set_error("dm4d_mesh_normal_consistency: bad argument");
        return DM4D_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    DM4D_CUDA_CHECK(cudaMemsetAsync(loss, 0, (size_t)n_t * sizeof(float), s));
    if (dL_dverts) DM4D_CUDA_CHECK(cudaMemsetAsync(dL_dverts, 0, (size_t)n_t * V * 3 * sizeof(float), s));
    if (n_t > 65535) { dm4d_set_error("dm4d_mesh_normal_consistency: n_t > 65535"); return DM4D_EINVAL; }
    const dim3 grid((unsigned)((n_pairs + DM4D_BLOCK - 1) / DM4D_BLOCK), (unsigned)n_t);
    { KernelTimer kt(DM4D_K_NORMAL_CONS, s); normal_consistency_kernel<<<grid, DM4D_BLOCK, 0, s>>>(pairs, n_pairs, n_t, V, verts, loss, dL_dverts); }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}
