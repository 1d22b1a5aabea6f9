// Per-(view, Gaussian) projection backward (SURVEY.md Appendix A.3 "preprocess-bwd"): conic -> 2D covariance ->
// (Sigma, view-space point) -> mean / scale / quaternion gradients.  Own translation unit with default FMA
// contraction (the forward's -fmad=false arithmetic spec only protects integer state; gradients carry none).
#include "raster_project.cuh"

namespace {

struct PreBwdArgs {
    PreArgs f;
    const float* accum;
    float* dmeans3D; int dmeans3D_atomic;
    float* dmeans2D;
    float* dcolors;  int dcolors_atomic;
    float* dcolors2; int dcolors2_atomic;
    float* dopac;    int dopac_atomic;
    float* dscales;  int dscales_atomic;
    float* drots;    int drots_atomic;
    float* dcov;     int dcov_atomic;
};

__device__ __forceinline__ void emit(float* p, float v, int atomic) {
    if (atomic) atomicAdd(p, v);
    else *p = v;
}

#ifndef DM4D_PREBWD_MIN_BLOCKS
#define DM4D_PREBWD_MIN_BLOCKS 4
#endif
template <bool COV>
__global__ void __launch_bounds__(DM4D_BLOCK, DM4D_PREBWD_MIN_BLOCKS) preprocess_backward_kernel(PreBwdArgs b) {
    __shared__ ViewCache vcache;
    ViewRows vc;
    const PreArgs& a = b.f;
    const long long first = (long long)blockIdx.x * blockDim.x;
    const long long idx = first + threadIdx.x;
    vc.fill(&vcache, a.view_params, (int)(first / a.P), a.n_views);
    __syncthreads();
    if (idx >= (long long)a.n_views * a.P) return;
    int v, g;
    split_index(first, threadIdx.x, a.P, v, g);
    const float* __restrict__ vp = vc.row(v);
    const long long set = min(max((long long)vp[DM4D_VIEW_SET], 0ll), (long long)a.n_sets - 1);
    const bool live = a.g_rect[idx] != 0u;
    const float* acc = b.accum + (size_t)idx * a.acc;

    float dm[3] = {0.f, 0.f, 0.f}, ds[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float g2x = 0.f, g2y = 0.f, gop = 0.f;
    float gcol[DM4D_MAX_CHANNELS] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

    if (live) {
        const float4 a0 = reinterpret_cast<const float4*>(acc)[0];
        const float4 a1 = reinterpret_cast<const float4*>(acc)[1];
        g2x = a0.x; g2y = a0.y;
        const float gx = a0.z, gy = a0.w, gz = a1.x;
        gop = a1.y;
        const float gd = a1.z;
        for (int ch = 0; ch < a.channels; ++ch) gcol[ch] = acc[8 + ch];

        const float* m = a.means3D + set * a.means3D_stride + (size_t)g * 3;
        const float mod = vp[DM4D_VIEW_SCALE_MOD];
        const float fx = (float)a.W / (2.0f * vp[DM4D_VIEW_TANFOVX]);
        const float fy = (float)a.H / (2.0f * vp[DM4D_VIEW_TANFOVY]);
        const float px = m[0], py = m[1], pz = m[2];
        float q[4] = {1.f, 0.f, 0.f, 0.f}, s[3] = {0.f, 0.f, 0.f};
        Proj pr;
        if (COV) {
            load_sigma(a.cov3D + set * a.cov3D_stride + (size_t)g * 6, pr.S);
        } else {
            const float* sc = a.scales + set * a.scales_stride + (size_t)g * 3;
            const float4 q4 = load_quat(a.rotations + set * a.rotations_stride + (size_t)g * 4);
            q[0] = q4.x; q[1] = q4.y; q[2] = q4.z; q[3] = q4.w;
            s[0] = mod * sc[0]; s[1] = mod * sc[1]; s[2] = mod * sc[2];
            gaussian_sigma(s[0], s[1], s[2], q[0], q[1], q[2], q[3], pr.S);
        }
        project_with_sigma(vp, px, py, pz, fx, fy, pr);
        const float* V = vp;
        const float* PV = vp + 16;

        // (1) conic -> cov2D -> (Sigma, t)
        const float A_ = pr.a, B_ = pr.b, C_ = pr.c;
        const float denom = A_ * C_ - B_ * B_;
        const float d2 = 1.0f / (denom * denom + 0.0000001f);
        const float dL_da = d2 * (-C_ * C_ * gx + 2.f * B_ * C_ * gy + (denom - A_ * C_) * gz);
        const float dL_dc = d2 * (-A_ * A_ * gz + 2.f * A_ * B_ * gy + (denom - A_ * C_) * gx);
        const float dL_db = d2 * 2.f * (B_ * C_ * gx - (denom + 2.f * B_ * B_) * gy + A_ * B_ * gz);
        const float hb = 0.5f * dL_db;
        float GS[3][3], dA[2][3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                GS[j][k] = dL_da * pr.A[0][j] * pr.A[0][k] + hb * (pr.A[0][j] * pr.A[1][k] + pr.A[1][j] * pr.A[0][k]) +
                           dL_dc * pr.A[1][j] * pr.A[1][k];
            dA[0][j] = 2.f * (dL_da * pr.B[0][j] + hb * pr.B[1][j]);
            dA[1][j] = 2.f * (hb * pr.B[0][j] + dL_dc * pr.B[1][j]);
        }
        float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            dJ00 += dA[0][j] * V[0 + 4 * j];
            dJ02 += dA[0][j] * V[2 + 4 * j];
            dJ11 += dA[1][j] * V[1 + 4 * j];
            dJ12 += dA[1][j] * V[2 + 4 * j];
        }
        const float tz1 = 1.f / pr.tz, tz2 = tz1 * tz1, tz3 = tz2 * tz1;
        const float dL_dtx = pr.xmul * -fx * tz2 * dJ02;
        const float dL_dty = pr.ymul * -fy * tz2 * dJ12;
        const float dL_dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.f * fx * pr.cx) * tz3 * dJ02 +
                             (2.f * fy * pr.cy) * tz3 * dJ12;
#pragma unroll
        for (int j = 0; j < 3; ++j) dm[j] = V[0 + 4 * j] * dL_dtx + V[1 + 4 * j] * dL_dty + V[2 + 4 * j] * dL_dtz;

        // (2) 2D mean -> 3D mean, (3) depth -> 3D mean
        const float m_w = pr.p_w;
        const float mul1 = pr.hx * m_w * m_w, mul2 = pr.hy * m_w * m_w;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            dm[k] += (PV[0 + 4 * k] * m_w - PV[3 + 4 * k] * mul1) * g2x + (PV[1 + 4 * k] * m_w - PV[3 + 4 * k] * mul2) * g2y;
            dm[k] += (V[2 + 4 * k] - V[3 + 4 * k] * pr.tz) * gd;
        }

        if (COV) {
            // gradient w.r.t. the six unique entries of the precomputed covariance (off-diagonal entries occur twice)
            dcov[0] = GS[0][0]; dcov[1] = 2.f * GS[0][1]; dcov[2] = 2.f * GS[0][2];
            dcov[3] = GS[1][1]; dcov[4] = 2.f * GS[1][2]; dcov[5] = GS[2][2];
        } else {
        // (5) Sigma = L L^T, L = Rm diag(s)
        float Rm[3][3], L[3][3], dLm[3][3], dR[3][3];
        quat_to_R(q[0], q[1], q[2], q[3], Rm);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) L[i][k] = Rm[i][k] * s[k];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) dLm[i][k] = 2.f * (GS[i][0] * L[0][k] + GS[i][1] * L[1][k] + GS[i][2] * L[2][k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < 3; ++i) { d += dLm[i][k] * Rm[i][k]; dR[i][k] = dLm[i][k] * s[k]; }
            ds[k] = mod * d;
        }
        const float r = q[0], x = q[1], y = q[2], z = q[3];
        dq[0] = 2.f * (z * (dR[1][0] - dR[0][1]) + y * (dR[0][2] - dR[2][0]) + x * (dR[2][1] - dR[1][2]));
        dq[1] = 2.f * (y * (dR[0][1] + dR[1][0]) + z * (dR[0][2] + dR[2][0]) + r * (dR[2][1] - dR[1][2])) - 4.f * x * (dR[1][1] + dR[2][2]);
        dq[2] = 2.f * (x * (dR[0][1] + dR[1][0]) + r * (dR[0][2] - dR[2][0]) + z * (dR[1][2] + dR[2][1])) - 4.f * y * (dR[0][0] + dR[2][2]);
        dq[3] = 2.f * (r * (dR[1][0] - dR[0][1]) + x * (dR[0][2] + dR[2][0]) + y * (dR[1][2] + dR[2][1])) - 4.f * z * (dR[0][0] + dR[1][1]);
        }
    }

    if (b.dmeans2D) {
        float* o = b.dmeans2D + (size_t)idx * 3;
        o[0] = g2x; o[1] = g2y; o[2] = 0.f;
    }
    // With atomics the outputs were zero-filled by the launcher; dead Gaussians add nothing.
    if (b.dmeans3D && (live || !b.dmeans3D_atomic)) {
        float* o = b.dmeans3D + set * a.means3D_stride + (size_t)g * 3;
        for (int k = 0; k < 3; ++k) emit(o + k, dm[k], b.dmeans3D_atomic);
    }
    if (b.dscales && (live || !b.dscales_atomic)) {
        float* o = b.dscales + set * a.scales_stride + (size_t)g * 3;
        for (int k = 0; k < 3; ++k) emit(o + k, ds[k], b.dscales_atomic);
    }
    if (b.drots && (live || !b.drots_atomic)) {
        float* o = b.drots + set * a.rotations_stride + (size_t)g * 4;
        for (int k = 0; k < 4; ++k) emit(o + k, dq[k], b.drots_atomic);
    }
    if (COV && b.dcov && (live || !b.dcov_atomic)) {
        float* o = b.dcov + set * a.cov3D_stride + (size_t)g * 6;
        for (int k = 0; k < 6; ++k) emit(o + k, dcov[k], b.dcov_atomic);
    }
    if (b.dopac && (live || !b.dopac_atomic)) emit(b.dopac + set * a.opacities_stride + g, gop, b.dopac_atomic);
    if (b.dcolors && (live || !b.dcolors_atomic)) {
        float* o = b.dcolors + set * a.colors_stride + (size_t)g * 3;
        for (int k = 0; k < 3; ++k) emit(o + k, gcol[k], b.dcolors_atomic);
    }
    if (b.dcolors2 && a.channels > 3 && (live || !b.dcolors2_atomic)) {
        float* o = b.dcolors2 + set * a.colors2_stride + (size_t)g * 3;
        for (int k = 0; k < 3; ++k) emit(o + k, gcol[3 + k], b.dcolors2_atomic);
    }
}

}  // namespace

int launch_preprocess_backward(const dm4d_raster_desc* d, const RasterLayout& L, float* dL_dmeans3D,
                               float* dL_dmeans2D, float* dL_dcolors, float* dL_dcolors2, float* dL_dopacities,
                               float* dL_dscales, float* dL_drotations, cudaStream_t s) {
    const long long n = (long long)L.n_views * L.P;
    if (n == 0) return DM4D_OK;
    PreBwdArgs b;
    b.f = make_pre_args(d, L, nullptr);
    b.accum = L.accum;
    // Outputs are fully overwritten. A set-strided attribute can be stored directly when the
    // caller promises a view<->set bijection (or there is a single view and a single set);
    // otherwise the output is zero-filled here and accumulated with atomics.
    const bool bijection = (d->flags & DM4D_RASTER_VIEWS_DISTINCT_SETS) && d->n_sets == d->n_views;
    auto mode = [&](float* p, long long stride, size_t elems_per_set) -> int {
        if (!p) return 0;
        const bool direct = (stride != 0 && bijection) || (d->n_views == 1 && (stride == 0 || d->n_sets == 1));
        if (!direct) {
            const size_t sets = stride == 0 ? 1 : (size_t)d->n_sets;
            cudaMemsetAsync(p, 0, sets * elems_per_set * sizeof(float), s);
        }
        return direct ? 0 : 1;
    };
    const size_t P = (size_t)L.P;
    b.dmeans3D = dL_dmeans3D; b.dmeans3D_atomic = mode(dL_dmeans3D, d->means3D_stride, P * 3);
    b.dmeans2D = dL_dmeans2D;
    b.dcolors = dL_dcolors;   b.dcolors_atomic = mode(dL_dcolors, d->colors_stride, P * 3);
    b.dcolors2 = dL_dcolors2; b.dcolors2_atomic = mode(dL_dcolors2, d->colors2_stride, P * 3);
    b.dopac = dL_dopacities;  b.dopac_atomic = mode(dL_dopacities, d->opacities_stride, P);
    b.dscales = dL_dscales;   b.dscales_atomic = mode(dL_dscales, d->scales_stride, P * 3);
    b.drots = dL_drotations;  b.drots_atomic = mode(dL_drotations, d->rotations_stride, P * 4);
    b.dcov = d->cov3D ? d->dL_dcov3D : nullptr; b.dcov_atomic = mode(b.dcov, d->cov3D_stride, P * 6);
    if (d->cov3D) { b.dscales = nullptr; b.drots = nullptr; }        // scales / rotations are not inputs of this call
    const unsigned blocks = (unsigned)((n + DM4D_BLOCK - 1) / DM4D_BLOCK);
    {
        KernelTimer kt(DM4D_K_PREPROCESS_BWD, s);
        if (d->cov3D) preprocess_backward_kernel<true><<<blocks, DM4D_BLOCK, 0, s>>>(b);
        else preprocess_backward_kernel<false><<<blocks, DM4D_BLOCK, 0, s>>>(b);
    }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}
