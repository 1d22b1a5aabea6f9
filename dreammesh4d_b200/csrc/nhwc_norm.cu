// Channels-last GroupNorm (+ per-sample channel bias in front, + SiLU behind) for the Zero123 networks of the SDS step
// (SURVEY.md §8 row A9).  The convolutions / matmuls of those networks stay PyTorch tensor-core library calls; this is
// the normalisation / activation glue between them, which in eager PyTorch costs more device time than the GEMMs:
// torch's CUDA GroupNorm only knows NCHW, so every ResBlock of the reference's UNetModel / Encoder
// (extern/ldm_zero123/modules/diffusionmodules/openaimodel.py:258-289, model.py:118-138) pays NHWC<->NCHW transposes
// around each cuDNN convolution plus separate passes for the time-embedding add, the norm and the SiLU.
//
//   y[n,p,c] = act( ((x[n,p,c] + e[n,c]) - mean[n,g]) * rstd[n,g] * gamma[c] + beta[c] ),   g = c / (C / G)
//
// Layout: x, y are [N, HW, C] (the memory of a torch channels_last [N,C,H,W] tensor), fp16 or fp32; statistics in fp32.
// Two launches each way: (1) per-(sample, pixel-slab) partial sums -> [N,G,2] with one atomic per (block, group);
// (2) element-wise apply with 8- / 16-byte vector accesses.  HBM-bound: forward reads x twice and writes y once.
// Backward (weights are frozen in the SDS step: no dgamma / dbeta) needs the two group sums of dz*gamma and dz*gamma*xhat.
#include <algorithm>
#include <cuda_fp16.h>
#include "raster_internal.cuh"

namespace {

constexpr int VEC = 4;                 // channels per thread per access (8 B for fp16, 16 B for fp32)
constexpr int MAX_GROUPS = 64;

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec4<__half> {
    static __device__ __forceinline__ void load(const __half* p, float (&v)[4]) {
        const uint2 t = *reinterpret_cast<const uint2*>(p);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    static __device__ __forceinline__ void store(__half* p, const float (&v)[4]) {
        uint2 t;
        *reinterpret_cast<__half2*>(&t.x) = __floats2half2_rn(v[0], v[1]);
        *reinterpret_cast<__half2*>(&t.y) = __floats2half2_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = t;
    }
};

struct NormArgs {
    int N, HW, C, G, cpg, cvec;          // cvec = C / VEC
    int rows_per_block;                  // pixels handled by one block of the statistics kernels
    float eps;
    int silu;
};

__device__ __forceinline__ float silu_f(float z) { return z / (1.0f + __expf(-z)); }
__device__ __forceinline__ float silu_grad(float z) {
    const float s = 1.0f / (1.0f + __expf(-z));
    return s * (1.0f + z * (1.0f - s));
}

// Statistics: block = (C/VEC) x k threads; thread (cv, r) owns channels [cv*4, cv*4+4) and walks the pixels r, r+k, ...
// of its slab.  Per-channel partial sums live in registers; at the end every thread folds its 4 channels into the
// block's group bins in shared memory, and the block issues one global atomic per (group, moment).
template <typename T, bool BWD>
__global__ void gn_stats_kernel(NormArgs a, const T* __restrict__ x, const float* __restrict__ chan_bias,
                                const T* __restrict__ dy, const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ stats, float* __restrict__ sums) {
    __shared__ float bins[2 * MAX_GROUPS];
    const int n = blockIdx.y;
    const int cv = threadIdx.x % a.cvec, r = threadIdx.x / a.cvec, k = blockDim.x / a.cvec;
    for (int i = threadIdx.x; i < 2 * a.G; i += blockDim.x) bins[i] = 0.f;
    __syncthreads();
    const int p0 = blockIdx.x * a.rows_per_block, p1 = min(a.HW, p0 + a.rows_per_block);
    const int c0 = cv * VEC;
    float s0[VEC], s1[VEC], e[VEC], gam[VEC], bet[VEC], mu[VEC], rs[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        s0[j] = s1[j] = 0.f;
        e[j] = chan_bias ? chan_bias[(size_t)n * a.C + c0 + j] : 0.f;
        if (BWD) {
            const int g = (c0 + j) / a.cpg;
            gam[j] = gamma[c0 + j]; bet[j] = beta[c0 + j];
            mu[j] = stats[((size_t)n * a.G + g) * 2]; rs[j] = stats[((size_t)n * a.G + g) * 2 + 1];
        }
    }
    if (r < k && cv < a.cvec) {
        for (int p = p0 + r; p < p1; p += k) {
            const size_t off = ((size_t)n * a.HW + p) * a.C + c0;
            float v[VEC];
            Vec4<T>::load(x + off, v);
            if (!BWD) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) { const float t = v[j] + e[j]; s0[j] += t; s1[j] += t * t; }
            } else {
                float d[VEC];
                Vec4<T>::load(dy + off, d);
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const float xh = (v[j] + e[j] - mu[j]) * rs[j];
                    float dz = d[j];
                    if (a.silu) dz *= silu_grad(xh * gam[j] + bet[j]);
                    const float t = dz * gam[j];
                    s0[j] += t; s1[j] += t * xh;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int g = (c0 + j) / a.cpg;
        atomicAdd(&bins[2 * g], s0[j]);
        atomicAdd(&bins[2 * g + 1], s1[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * a.G; i += blockDim.x) atomicAdd(&sums[(size_t)n * 2 * a.G + i], bins[i]);
}

// sums [N,G,2] (sum, sum of squares) -> stats [N,G,2] (mean, rstd)
__global__ void gn_finalize_kernel(int total, float count, float eps, const float* __restrict__ sums, float* __restrict__ stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float mean = sums[2 * i] / count;
    const float var = fmaxf(sums[2 * i + 1] / count - mean * mean, 0.f);
    stats[2 * i] = mean;
    stats[2 * i + 1] = rsqrtf(var + eps);
}

template <typename T>
__global__ void gn_apply_kernel(NormArgs a, const T* __restrict__ x, const float* __restrict__ chan_bias,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ stats, T* __restrict__ y) {
    const size_t total = (size_t)a.N * a.HW * a.cvec;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % a.cvec);
        const size_t np = i / a.cvec;
        const int n = (int)(np / a.HW);
        const int c0 = cv * VEC;
        float v[VEC], o[VEC];
        Vec4<T>::load(x + np * a.C + c0, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const int c = c0 + j, g = c / a.cpg;
            const float mean = stats[((size_t)n * a.G + g) * 2], rstd = stats[((size_t)n * a.G + g) * 2 + 1];
            const float e = chan_bias ? chan_bias[(size_t)n * a.C + c] : 0.f;
            const float z = (v[j] + e - mean) * rstd * gamma[c] + beta[c];
            o[j] = a.silu ? silu_f(z) : z;
        }
        Vec4<T>::store(y + np * a.C + c0, o);
    }
}

// dx = rstd * (dz*gamma - mean_g(dz*gamma) - xhat * mean_g(dz*gamma*xhat)),  dz = dy * act'(z)
// (d chan_bias = sum over pixels of dx is not produced: the time embedding carries no gradient in the SDS step)
template <typename T>
__global__ void gn_backward_apply_kernel(NormArgs a, const T* __restrict__ x, const float* __restrict__ chan_bias,
                                         const T* __restrict__ dy, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, const float* __restrict__ stats,
                                         const float* __restrict__ bsums, T* __restrict__ dx) {
    const size_t total = (size_t)a.N * a.HW * a.cvec;
    const float inv_count = 1.0f / ((float)a.HW * (float)a.cpg);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % a.cvec);
        const size_t np = i / a.cvec;
        const int n = (int)(np / a.HW);
        const int c0 = cv * VEC;
        float v[VEC], d[VEC], o[VEC];
        Vec4<T>::load(x + np * a.C + c0, v);
        Vec4<T>::load(dy + np * a.C + c0, d);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const int c = c0 + j, g = c / a.cpg;
            const size_t sg = ((size_t)n * a.G + g) * 2;
            const float mean = stats[sg], rstd = stats[sg + 1];
            const float e = chan_bias ? chan_bias[(size_t)n * a.C + c] : 0.f;
            const float xh = (v[j] + e - mean) * rstd;
            float dz = d[j];
            if (a.silu) dz *= silu_grad(xh * gamma[c] + beta[c]);
            o[j] = rstd * (dz * gamma[c] - bsums[sg] * inv_count - xh * bsums[sg + 1] * inv_count);
        }
        Vec4<T>::store(dx + np * a.C + c0, o);
    }
}

int make_args(int N, int HW, int C, int G, float eps, int silu, NormArgs* a, int* threads, dim3* grid) {
    if (N <= 0 || HW <= 0 || C <= 0 || G <= 0 || G > MAX_GROUPS || C % G || C % VEC || C / VEC > 1024) {
        dm4d_set_error("groupnorm_nhwc: need C %% G == 0, C %% 4 == 0, C <= 4096, G <= 64 (N=%d HW=%d C=%d G=%d)", N, HW, C, G);
        return DM4D_EINVAL;
    }
    a->N = N; a->HW = HW; a->C = C; a->G = G; a->cpg = C / G; a->cvec = C / VEC; a->eps = eps; a->silu = silu;
    const int k = max(1, min(256 / a->cvec, HW));
    *threads = a->cvec * k;
    // enough blocks per sample to fill the machine (148 SMs x a few CTAs), at least 8 pixels per thread row
    int slabs = max(1, min((HW + 8 * k - 1) / (8 * k), max(1, (148 * 8 + N - 1) / N)));
    a->rows_per_block = (HW + slabs - 1) / slabs;
    *grid = dim3((unsigned)((HW + a->rows_per_block - 1) / a->rows_per_block), (unsigned)N);
    return DM4D_OK;
}

unsigned apply_blocks(const NormArgs& a) {
    const size_t total = (size_t)a.N * a.HW * a.cvec;
    return (unsigned)min((size_t)148 * 16, (total + 255) / 256);
}

template <typename T>
int forward_t(const NormArgs& a, int threads, dim3 grid, const void* x, const float* cb, const float* gamma, const float* beta,
              float* stats, float* scratch, void* y, cudaStream_t s) {
    DM4D_CUDA_CHECK(cudaMemsetAsync(scratch, 0, (size_t)a.N * a.G * 2 * sizeof(float), s));
    {
        KernelTimer kt(DM4D_K_GROUPNORM_FWD, s);
        gn_stats_kernel<T, false><<<grid, threads, 0, s>>>(a, (const T*)x, cb, nullptr, nullptr, nullptr, nullptr, scratch);
        gn_finalize_kernel<<<(a.N * a.G + 127) / 128, 128, 0, s>>>(a.N * a.G, (float)a.HW * (float)a.cpg, a.eps, scratch, stats);
        gn_apply_kernel<T><<<apply_blocks(a), 256, 0, s>>>(a, (const T*)x, cb, gamma, beta, stats, (T*)y);
    }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

template <typename T>
int backward_t(const NormArgs& a, int threads, dim3 grid, const void* x, const float* cb, const void* dy, const float* gamma,
               const float* beta, const float* stats, float* scratch, void* dx, cudaStream_t s) {
    DM4D_CUDA_CHECK(cudaMemsetAsync(scratch, 0, (size_t)a.N * a.G * 2 * sizeof(float), s));
    {
        KernelTimer kt(DM4D_K_GROUPNORM_BWD, s);
        gn_stats_kernel<T, true><<<grid, threads, 0, s>>>(a, (const T*)x, cb, (const T*)dy, gamma, beta, stats, scratch);
        gn_backward_apply_kernel<T><<<apply_blocks(a), 256, 0, s>>>(a, (const T*)x, cb, (const T*)dy, gamma, beta, stats, scratch, (T*)dx);
    }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

// ---- convolution epilogues the library calls do not fuse ------------------------------------------------------------
// out[m, c] = h[m, c] + bias[c] (+ res[m, c]): the bias of a cuDNN convolution (PyTorch adds it in a separate broadcast
// pass over the output) folded into the residual add of the ResBlock that follows it (openaimodel.py:289,
// model.py:138) — one pass instead of two.  Grid-stride over 4-channel vectors.
template <typename T>
__global__ void __launch_bounds__(256) bias_residual_kernel(long long nvec, int cvec, const T* __restrict__ h,
                                                            const T* __restrict__ res, const float* __restrict__ bias,
                                                            T* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % cvec) * VEC;
        float v[VEC], r[VEC];
        Vec4<T>::load(h + i * VEC, v);
        const float4 b = *reinterpret_cast<const float4*>(bias + c0);
        v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
        if (res) {
            Vec4<T>::load(res + i * VEC, r);
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[j] += r[j];
        }
        Vec4<T>::store(out + i * VEC, v);
    }
}

// GEGLU of the transformer feed-forward (attention.py:37-65): proj [M, 2D] -> out[m, d] = proj[m, d] * gelu(proj[m, D + d])
// with the exact (erf) GELU; eager PyTorch runs chunk -> gelu -> mul as two passes over strided halves.
template <typename T>
__global__ void __launch_bounds__(256) geglu_kernel(long long nvec, int dvec, const T* __restrict__ proj, T* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / dvec;
        const int d0 = (int)(i - m * dvec) * VEC;
        const T* row = proj + m * (2ll * dvec * VEC);
        float a[VEC], g[VEC];
        Vec4<T>::load(row + d0, a);
        Vec4<T>::load(row + (size_t)dvec * VEC + d0, g);
#pragma unroll
        for (int j = 0; j < VEC; ++j) a[j] *= 0.5f * g[j] * (1.0f + erff(g[j] * 0.70710678118654752f));
        Vec4<T>::store(out + i * VEC, a);
    }
}

template <typename T>
int launch_bias_residual(long long M, int C, const void* h, const void* res, const float* bias, void* out, cudaStream_t s) {
    const long long nvec = M * (C / VEC);
    const unsigned blocks = (unsigned)std::min<long long>((nvec + 255) / 256, 148ll * 16);
    bias_residual_kernel<T><<<blocks, 256, 0, s>>>(nvec, C / VEC, (const T*)h, (const T*)res, bias, (T*)out);
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}
template <typename T>
int launch_geglu(long long M, int D, const void* proj, void* out, cudaStream_t s) {
    const long long nvec = M * (D / VEC);
    const unsigned blocks = (unsigned)std::min<long long>((nvec + 255) / 256, 148ll * 16);
    geglu_kernel<T><<<blocks, 256, 0, s>>>(nvec, D / VEC, (const T*)proj, (T*)out);
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

}  // namespace

extern "C" int dm4d_bias_residual_add_nhwc(const void* h, const void* residual, const float* bias, int64_t M, int32_t C,
                                           int32_t dtype, void* out, void* stream) {
    if (!h || !bias || !out || M <= 0 || C <= 0 || C % VEC) { dm4d_set_error("dm4d_bias_residual_add_nhwc: bad argument (C must be a multiple of 4)"); return DM4D_EINVAL; }
    if (dtype == DM4D_F16) return launch_bias_residual<__half>(M, C, h, residual, bias, out, (cudaStream_t)stream);
    if (dtype == DM4D_F32) return launch_bias_residual<float>(M, C, h, residual, bias, out, (cudaStream_t)stream);
    dm4d_set_error("dm4d_bias_residual_add_nhwc: dtype must be DM4D_F16 or DM4D_F32");
    return DM4D_EINVAL;
}

extern "C" int dm4d_geglu(const void* proj, int64_t M, int32_t D, int32_t dtype, void* out, void* stream) {
    if (!proj || !out || M <= 0 || D <= 0 || D % VEC) { dm4d_set_error("dm4d_geglu: bad argument (D must be a multiple of 4)"); return DM4D_EINVAL; }
    if (dtype == DM4D_F16) return launch_geglu<__half>(M, D, proj, out, (cudaStream_t)stream);
    if (dtype == DM4D_F32) return launch_geglu<float>(M, D, proj, out, (cudaStream_t)stream);
    dm4d_set_error("dm4d_geglu: dtype must be DM4D_F16 or DM4D_F32");
    return DM4D_EINVAL;
}

extern "C" int dm4d_groupnorm_nhwc_forward(const void* x, const float* chan_bias, const float* gamma, const float* beta,
                                           int32_t N, int32_t HW, int32_t C, int32_t G, float eps, int32_t silu,
                                           int32_t dtype, float* stats, float* scratch, void* y, void* stream) {
    NormArgs a; int threads; dim3 grid;
    int rc = make_args(N, HW, C, G, eps, silu, &a, &threads, &grid);
    if (rc) return rc;
    if (!x || !gamma || !beta || !stats || !scratch || !y) { dm4d_set_error("groupnorm_nhwc: NULL pointer"); return DM4D_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == DM4D_F16) return forward_t<__half>(a, threads, grid, x, chan_bias, gamma, beta, stats, scratch, y, s);
    if (dtype == DM4D_F32) return forward_t<float>(a, threads, grid, x, chan_bias, gamma, beta, stats, scratch, y, s);
    dm4d_set_error("groupnorm_nhwc: dtype must be DM4D_F16 or DM4D_F32");
    return DM4D_EINVAL;
}

extern "C" int dm4d_groupnorm_nhwc_backward(const void* x, const float* chan_bias, const void* dy, const float* gamma,
                                            const float* beta, int32_t N, int32_t HW, int32_t C, int32_t G, float eps,
                                            int32_t silu, int32_t dtype, const float* stats, float* scratch, void* dx,
                                            void* stream) {
    NormArgs a; int threads; dim3 grid;
    int rc = make_args(N, HW, C, G, eps, silu, &a, &threads, &grid);
    if (rc) return rc;
    if (!x || !dy || !gamma || !beta || !stats || !scratch || !dx) { dm4d_set_error("groupnorm_nhwc: NULL pointer"); return DM4D_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == DM4D_F16) return backward_t<__half>(a, threads, grid, x, chan_bias, dy, gamma, beta, stats, scratch, dx, s);
    if (dtype == DM4D_F32) return backward_t<float>(a, threads, grid, x, chan_bias, dy, gamma, beta, stats, scratch, dx, s);
    dm4d_set_error("groupnorm_nhwc: dtype must be DM4D_F16 or DM4D_F32");
    return DM4D_EINVAL;
}
