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
// Two streamed launches each way (statistics, apply) with 16-byte accesses, or — for the UNet's small fp16 activations —
// one launch that keeps its slab in registers between the two phases.
// Backward (weights are frozen in the SDS step: no dgamma / dbeta) needs the two group sums of dz*gamma and dz*gamma*xhat.
#include <algorithm>
#include <cmath>
#include <cuda_fp16.h>
#include "raster_internal.cuh"


namespace {

constexpr int VEC = 4;                 // channels per thread per access (8 B for fp16, 16 B for fp32)

// SMs of the current device (148 on a B200).  The persistent kernels below bound their grids by the RESIDENT capacity
// (SMs x occupancy): statically assigned work items must never sit behind CTAs that spin on them.
int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    return n;
}
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
    static __device__ __forceinline__ void round(float (&)[4]) {}
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
    static __device__ __forceinline__ void round(float (&v)[4]) {      // what a store + load in this dtype would give
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __half2float(__float2half_rn(v[j]));
    }
};

struct NormArgs {
    int N, HW, C, G, cpg, cvec;          // cvec = C / VEC
    int rows_per_block;                  // pixels handled by one block of the statistics kernels
    float eps;
    int silu;
};

// sigmoid through MUFU.EX2 + MUFU.RCP (1-2 ulp): the IEEE division cost ~10 instructions per element and made the
// backward passes instruction-bound (two evaluations per element and pass)
__device__ __forceinline__ float approx_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float silu_f(float z) { return z * approx_rcp(1.0f + __expf(-z)); }
__device__ __forceinline__ float silu_grad(float z) {
    const float s = approx_rcp(1.0f + __expf(-z));
    return s * (1.0f + z * (1.0f - s));
}

// Thread layout of a block: (C/VEC) x k threads; thread (cv, r) owns channels [cv*4, cv*4+4) and walks the pixels
// r, r+k, ... of a slab (coalesced 8- / 16-byte accesses along the channel axis).  Statistics: per-channel partial sums
// in registers, folded into the block's group bins in shared memory, one global atomic per (block, group, moment).
// Backward: dx = rstd * (dz*gamma - mean_g(dz*gamma) - xhat * mean_g(dz*gamma*xhat)),  dz = dy * act'(z)
// (d chan_bias is not produced: the time embedding carries no gradient in the SDS step).

// ---- streamed GroupNorm (every activation that does not take the register-resident path below) ------------------------
// Two launches each way over (sample, pixel slab) blocks with 16-byte accesses and four pixels in flight per thread:
//   statistics: per-channel partial sums in registers -> group bins in shared memory -> one global atomic per (block, group,
//               moment);   apply: the element-wise pass, group statistics finalised inline.
// A single persistent launch with an ordered work queue (statistics a bounded distance ahead of the apply pass so that the
// second read hits L2) was measured first: ncu showed the re-read still coming from DRAM (499 of 536 MB in the backward of
// an [8,128,256,256] activation) and the per-item synchronisation costing more than the saved launches — 157 / 353 us
// forward / backward against a 41 / 62 us HBM bound.
struct SumsArgs {
    float* sums;          // [N,G,2] zeroed: forward (sum, sum sq) / backward (sum dz*gamma, sum dz*gamma*xhat)
};
struct FusedArgs {        // arguments of the register-resident kernel
    NormArgs a;
    int slabs;
    int ahead;
    float* sums;
    int* counters;        // [1 + N] zeroed: (unused), per-sample arrivals
};

template <typename T> struct Wide;       // 16-byte accesses: 4 fp32 / 8 fp16 channels
template <> struct Wide<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) { Vec4<float>::load(p, v); }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) { Vec4<float>::store(p, v); }
};
template <> struct Wide<__half> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        const unsigned int w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__half* p, const float (&v)[8]) {
        unsigned int w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<__half2*>(&w[i]) = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

template <typename T, int VW> struct IO;
template <typename T> struct IO<T, 4> : Vec4<T> {};
template <> struct IO<__half, 8> : Wide<__half> {};

// a.cvec = C / VW here.  grid = (slabs, N); thread (cv, r) owns channels [cv*VW, cv*VW+VW) and walks the pixels
// p0 + r, p0 + r + k, ... of its slab.
template <typename T, bool BWD, int VW>
__global__ void __launch_bounds__(VW == 8 ? 512 : 1024) gn_stream_stats_kernel(NormArgs a, const T* __restrict__ x, const float* __restrict__ chan_bias,
                                                               const T* __restrict__ dy, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, const float* __restrict__ stats,
                                                               float* __restrict__ sums) {
    __shared__ float bins[2 * MAX_GROUPS];
    const int n = blockIdx.y;
    const int cv = threadIdx.x % a.cvec, r = threadIdx.x / a.cvec, k = blockDim.x / a.cvec;
    const int c0 = cv * VW;
    for (int i = threadIdx.x; i < 2 * a.G; i += blockDim.x) bins[i] = 0.f;
    __syncthreads();
    const int p0 = blockIdx.x * a.rows_per_block, p1 = min(a.HW, p0 + a.rows_per_block);
    float s0[VW], s1[VW], e[VW], gam[VW], bet[VW], mu[VW], rs[VW];
#pragma unroll
    for (int j = 0; j < VW; ++j) {
        s0[j] = s1[j] = 0.f;
        e[j] = chan_bias ? chan_bias[(size_t)n * a.C + c0 + j] : 0.f;
        if (BWD) {
            const size_t sg = ((size_t)n * a.G + (c0 + j) / a.cpg) * 2;
            gam[j] = gamma[c0 + j]; bet[j] = beta[c0 + j];
            mu[j] = stats[sg]; rs[j] = stats[sg + 1];
        }
    }
    const T* xn = x + (size_t)n * a.HW * a.C + c0;
    const T* dn = BWD ? dy + (size_t)n * a.HW * a.C + c0 : nullptr;
#pragma unroll 4
    for (int p = p0 + r; p < p1; p += k) {
        float v[VW];
        IO<T, VW>::load(xn + (size_t)p * a.C, v);
        if (!BWD) {
#pragma unroll
            for (int j = 0; j < VW; ++j) { const float t = v[j] + e[j]; s0[j] += t; s1[j] += t * t; }
        } else {
            float d[VW];
            IO<T, VW>::load(dn + (size_t)p * a.C, d);
#pragma unroll
            for (int j = 0; j < VW; ++j) {
                const float xh = (v[j] + e[j] - mu[j]) * rs[j];
                float dz = d[j];
                if (a.silu) dz *= silu_grad(xh * gam[j] + bet[j]);
                const float t = dz * gam[j];
                s0[j] += t; s1[j] += t * xh;
            }
        }
    }
    // channels of one thread that share a group are summed in registers before they touch the bins
#pragma unroll
    for (int j = 0; j < VW; ++j) {
        const int g = (c0 + j) / a.cpg;
        if (j + 1 < VW && (c0 + j + 1) / a.cpg == g) { s0[j + 1] += s0[j]; s1[j + 1] += s1[j]; continue; }
        atomicAdd(&bins[2 * g], s0[j]);
        atomicAdd(&bins[2 * g + 1], s1[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * a.G; i += blockDim.x) atomicAdd(&sums[(size_t)n * 2 * a.G + i], bins[i]);
}

template <typename T, bool BWD, int VW>
__global__ void __launch_bounds__(VW == 8 ? 512 : 1024) gn_stream_apply_kernel(NormArgs a, const T* __restrict__ x, const float* __restrict__ chan_bias,
                                                               const T* __restrict__ dy, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float* __restrict__ stats,
                                                               const float* __restrict__ sums, T* __restrict__ out) {
    const int n = blockIdx.y;
    const int cv = threadIdx.x % a.cvec, r = threadIdx.x / a.cvec, k = blockDim.x / a.cvec;
    const int c0 = cv * VW;
    const int p0 = blockIdx.x * a.rows_per_block, p1 = min(a.HW, p0 + a.rows_per_block);
    const float inv_count = 1.0f / ((float)a.HW * (float)a.cpg);
    // forward: y = act(v * sc + sh);  backward: dx = rs * (dz * gam - b0 - xh * b1), xh = v * rs + xo
    float sc[VW], sh[VW], gam[VW], bet[VW], rs[VW], xo[VW], b0[VW], b1[VW];
#pragma unroll
    for (int j = 0; j < VW; ++j) {
        const size_t sg = ((size_t)n * a.G + (c0 + j) / a.cpg) * 2;
        const float e = chan_bias ? chan_bias[(size_t)n * a.C + c0 + j] : 0.f;
        gam[j] = gamma[c0 + j]; bet[j] = beta[c0 + j];
        if (!BWD) {
            const float mean = sums[sg] * inv_count;
            const float rstd = rsqrtf(fmaxf(sums[sg + 1] * inv_count - mean * mean, 0.f) + a.eps);
            if (blockIdx.x == 0 && r == 0 && (c0 + j) % a.cpg == 0) { stats[sg] = mean; stats[sg + 1] = rstd; }
            sc[j] = rstd * gam[j];
            sh[j] = (e - mean) * sc[j] + bet[j];
        } else {
            rs[j] = stats[sg + 1];
            xo[j] = (e - stats[sg]) * rs[j];
            b0[j] = sums[sg] * inv_count; b1[j] = sums[sg + 1] * inv_count;
        }
    }
    const T* xn = x + (size_t)n * a.HW * a.C + c0;
    const T* dn = BWD ? dy + (size_t)n * a.HW * a.C + c0 : nullptr;
    T* on = out + (size_t)n * a.HW * a.C + c0;
#pragma unroll 4
    for (int p = p0 + r; p < p1; p += k) {
        float v[VW], o[VW];
        IO<T, VW>::load(xn + (size_t)p * a.C, v);
        if (!BWD) {
#pragma unroll
            for (int j = 0; j < VW; ++j) {
                const float z = v[j] * sc[j] + sh[j];
                o[j] = a.silu ? silu_f(z) : z;
            }
        } else {
            float d[VW];
            IO<T, VW>::load(dn + (size_t)p * a.C, d);
#pragma unroll
            for (int j = 0; j < VW; ++j) {
                const float xh = v[j] * rs[j] + xo[j];
                float dz = d[j];
                if (a.silu) dz *= silu_grad(xh * gam[j] + bet[j]);
                o[j] = rs[j] * (dz * gam[j] - b0[j] - xh * b1[j]);
            }
        }
        IO<T, VW>::store(on + (size_t)p * a.C, o);
    }
}

// Forward of SMALL activations (the UNet's: at most a few MB, one CTA per (sample, slab) with every CTA resident at once):
// the slab stays in registers between the statistics and the apply phase — x is read once, and the whole norm is one
// launch whose critical path is load -> group atomics -> per-sample arrival counter -> store.  fp16 only.
// Forward progress: a CTA spins on its sample's counter, so every CTA of the launch must eventually be resident — the
// launcher only takes this path when N x slabs <= SMs x occupancy.  Kernels of OTHER streams merely delay that (they
// drain); two of THESE launches running concurrently on one device could starve each other, so callers keep the norms of
// one device on one stream (the Zero123 networks and the step graph do).
constexpr int RES_IT = 12;                 // pixels per thread kept in registers (packed fp16: 2 registers each)

__global__ void __launch_bounds__(1024, 1) gn_resident_kernel(FusedArgs f, const __half* __restrict__ x, const float* __restrict__ chan_bias,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ stats, __half* __restrict__ out) {
    __shared__ float bins[2 * MAX_GROUPS];
    const NormArgs& a = f.a;
    const int n = blockIdx.x / f.slabs, slab = blockIdx.x - n * f.slabs;
    const int cv = threadIdx.x % a.cvec, r = threadIdx.x / a.cvec, k = blockDim.x / a.cvec;
    const int c0 = cv * VEC;
    const int p0 = slab * a.rows_per_block, p1 = min(a.HW, p0 + a.rows_per_block);
    const float inv_count = 1.0f / ((float)a.HW * (float)a.cpg);
    for (int i = threadIdx.x; i < 2 * a.G; i += blockDim.x) bins[i] = 0.f;
    __syncthreads();
    float e[VEC], s0[VEC], s1[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { e[j] = chan_bias ? chan_bias[(size_t)n * a.C + c0 + j] : 0.f; s0[j] = s1[j] = 0.f; }
    uint2 raw[RES_IT];
#pragma unroll
    for (int it = 0; it < RES_IT; ++it) {
        const int p = p0 + r + it * k;
        if (p < p1) {
            raw[it] = *reinterpret_cast<const uint2*>(x + ((size_t)n * a.HW + p) * a.C + c0);
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw[it].x));
            const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw[it].y));
            const float v[VEC] = {lo.x + e[0], lo.y + e[1], hi.x + e[2], hi.y + e[3]};
#pragma unroll
            for (int j = 0; j < VEC; ++j) { s0[j] += v[j]; s1[j] += v[j] * v[j]; }
        }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int g = (c0 + j) / a.cpg;
        atomicAdd(&bins[2 * g], s0[j]);
        atomicAdd(&bins[2 * g + 1], s1[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * a.G; i += blockDim.x) atomicAdd(&f.sums[(size_t)n * 2 * a.G + i], bins[i]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(&f.counters[1 + n], 1);
        while (atomicAdd(&f.counters[1 + n], 0) < f.slabs) __nanosleep(32);
        __threadfence();
    }
    __syncthreads();
    float sc[VEC], sh[VEC];                 // y = act(v * sc + sh)
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const size_t sg = ((size_t)n * a.G + (c0 + j) / a.cpg) * 2;
        const float mean = __ldcg(&f.sums[sg]) * inv_count;
        const float rstd = rsqrtf(fmaxf(__ldcg(&f.sums[sg + 1]) * inv_count - mean * mean, 0.f) + a.eps);
        if (slab == 0 && r == 0 && (c0 + j) % a.cpg == 0) { stats[sg] = mean; stats[sg + 1] = rstd; }
        sc[j] = rstd * gamma[c0 + j];
        sh[j] = (e[j] - mean) * sc[j] + beta[c0 + j];
    }
#pragma unroll
    for (int it = 0; it < RES_IT; ++it) {
        const int p = p0 + r + it * k;
        if (p < p1) {
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw[it].x));
            const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw[it].y));
            float o[VEC] = {lo.x * sc[0] + sh[0], lo.y * sc[1] + sh[1], hi.x * sc[2] + sh[2], hi.y * sc[3] + sh[3]};
            if (a.silu) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) o[j] = silu_f(o[j]);
            }
            Vec4<__half>::store(out + ((size_t)n * a.HW + p) * a.C + c0, o);
        }
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
    // enough work items per sample to fill the machine (148 SMs x a few CTAs), at least 8 pixels per thread row
    // (coarser items were measured slower even for the UNet's 10 MB activations: fewer CTAs in flight per pass)
    int slabs = max(1, min((HW + 8 * k - 1) / (8 * k), max(1, (148 * 8 + N - 1) / N)));
    a->rows_per_block = (HW + slabs - 1) / slabs;
    *grid = dim3((unsigned)((HW + a->rows_per_block - 1) / a->rows_per_block), (unsigned)N);
    return DM4D_OK;
}

#ifndef DM4D_GN_FWD_VW
#define DM4D_GN_FWD_VW 4      // fp16 channels per access in the forward kernels ([8,128,256,256]: 8: 118 us, 4: 105 us)
#endif
#ifndef DM4D_GN_BWD_VW
#define DM4D_GN_BWD_VW 4      // ... in the backward kernels (8: 16-byte accesses but ~100 registers: 304 us against 204 us)
#endif
template <typename T, bool BWD>
int launch_stream(const NormArgs& a4, const void* x, const float* cb, const void* dy, const float* gamma, const float* beta,
                  float* stats, float* scratch, void* out, cudaStream_t s) {
    constexpr int VW = sizeof(T) == 2 ? (BWD ? DM4D_GN_BWD_VW : DM4D_GN_FWD_VW) : 4;
    NormArgs a = a4;
    if (a.C % VW) { dm4d_set_error("groupnorm_nhwc: C must be a multiple of %d for this dtype (C=%d)", VW, a.C); return DM4D_EINVAL; }
    a.cvec = a.C / VW;
    const int k = std::max(1, std::min(256 / a.cvec, a.HW));
    const int threads = a.cvec * k;
    // enough blocks per sample to fill the machine (SMs x a few CTAs), at least 8 pixels per thread row
    const int slabs = std::max(1, std::min((a.HW + 8 * k - 1) / (8 * k), std::max(1, (sm_count() * 8 + a.N - 1) / a.N)));
    a.rows_per_block = (a.HW + slabs - 1) / slabs;
    const dim3 grid((unsigned)((a.HW + a.rows_per_block - 1) / a.rows_per_block), (unsigned)a.N);
    DM4D_CUDA_CHECK(cudaMemsetAsync(scratch, 0, (size_t)a.N * a.G * 2 * sizeof(float), s));
    {
        KernelTimer kt(BWD ? DM4D_K_GROUPNORM_BWD : DM4D_K_GROUPNORM_FWD, s);
        gn_stream_stats_kernel<T, BWD, VW><<<grid, threads, 0, s>>>(a, (const T*)x, cb, (const T*)dy, gamma, beta, stats, scratch);
        gn_stream_apply_kernel<T, BWD, VW><<<grid, threads, 0, s>>>(a, (const T*)x, cb, (const T*)dy, gamma, beta, stats, scratch, (T*)out);
    }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}

template <typename T>
int forward_t(const NormArgs& a, int threads, dim3 grid, const void* x, const float* cb, const float* gamma, const float* beta,
              float* stats, float* scratch, void* y, cudaStream_t s) {
    if constexpr (sizeof(T) == 2) {
        // resident single-pass variant when one CTA per (sample, slab) fits on the machine at once with <= RES_IT pixels per thread
        static int per_sm_cache[33] = {0};
        int& per_sm = per_sm_cache[threads / 32];
        if (per_sm == 0) DM4D_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_resident_kernel, threads, 0));
        const int k = threads / a.cvec;
        const int sms = sm_count();
        const int slabs = std::min((a.HW + k - 1) / k, std::max(1, sms * per_sm / a.N));
        const int rows = (a.HW + slabs - 1) / slabs;
        if (a.N * slabs <= sms * per_sm && (rows + k - 1) / k <= RES_IT) {
            FusedArgs f;
            f.a = a;
            f.a.rows_per_block = rows;
            f.slabs = (a.HW + rows - 1) / rows;
            f.ahead = a.N;
            f.sums = scratch;
            f.counters = reinterpret_cast<int*>(scratch + (size_t)a.N * a.G * 2);
            DM4D_CUDA_CHECK(cudaMemsetAsync(scratch, 0, ((size_t)a.N * a.G * 2 + a.N + 1) * sizeof(float), s));
            {
                KernelTimer kt(DM4D_K_GROUPNORM_FWD, s);
                gn_resident_kernel<<<(unsigned)(a.N * f.slabs), threads, 0, s>>>(f, (const __half*)x, cb, gamma, beta, stats, (__half*)y);
            }
            DM4D_CUDA_CHECK(cudaGetLastError());
            return DM4D_OK;
        }
    }
    return launch_stream<T, false>(a, x, cb, nullptr, gamma, beta, stats, scratch, y, s);
}

template <typename T>
int backward_t(const NormArgs& a, int threads, dim3 grid, const void* x, const float* cb, const void* dy, const float* gamma,
               const float* beta, const float* stats, float* scratch, void* dx, cudaStream_t s) {
    return launch_stream<T, true>(a, x, cb, dy, gamma, beta, const_cast<float*>(stats), scratch, dx, s);
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

// Residual add + LayerNorm of the transformer blocks (attention.py:199-246: x = x + attn(norm(x)) three times per block):
// x_out[m,:] = x[m,:] + delta[m or m / bcast_rows,:],  y[m,:] = LayerNorm(x_out[m,:]) * gamma + beta.  One warp per row,
// the row stays in registers (C <= 1536); eager PyTorch runs the add and a slower LayerNorm kernel as two passes.
constexpr int LN_MAXV = 12;                // 4-channel vectors per lane
template <typename T>
__global__ void __launch_bounds__(256) add_layernorm_kernel(long long M, int C, const T* __restrict__ x, const T* __restrict__ delta,
                                                            int bcast_rows, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, T* __restrict__ x_out,
                                                            T* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nvec = C / VEC;
    const T* xr = x + row * C;
    const T* dr = delta ? delta + (bcast_rows > 0 ? row / bcast_rows : row) * C : nullptr;
    float v[LN_MAXV][VEC];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) {
            Vec4<T>::load(xr + vi * VEC, v[i]);
            if (dr) {
                float d[VEC];
                Vec4<T>::load(dr + vi * VEC, d);
#pragma unroll
                for (int j = 0; j < VEC; ++j) v[i][j] += d[j];
                if (x_out) {
                    // the residual stream continues in the activation dtype: normalise what was stored
                    Vec4<T>::store(x_out + row * C + vi * VEC, v[i]);
                    Vec4<T>::round(v[i]);
                }
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) sum += v[i][j];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        if (lane + 32 * i < nvec) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) { const float t = v[i][j] - mean; sq += t * t; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)C + eps);
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int vi = lane + 32 * i;
        if (vi < nvec) {
            const float4 g = *reinterpret_cast<const float4*>(gamma + vi * VEC), b = *reinterpret_cast<const float4*>(beta + vi * VEC);
            const float o[VEC] = {(v[i][0] - mean) * rstd * g.x + b.x, (v[i][1] - mean) * rstd * g.y + b.y,
                                  (v[i][2] - mean) * rstd * g.z + b.z, (v[i][3] - mean) * rstd * g.w + b.w};
            Vec4<T>::store(y + row * C + vi * VEC, o);
        }
    }
}

template <typename T>
int launch_add_layernorm(long long M, int C, const void* x, const void* delta, int bcast_rows, const float* gamma, const float* beta,
                         float eps, void* x_out, void* y, cudaStream_t s) {
    const unsigned blocks = (unsigned)((M + 7) / 8);
    add_layernorm_kernel<T><<<blocks, 256, 0, s>>>(M, C, (const T*)x, (const T*)delta, bcast_rows, gamma, beta, eps, (T*)x_out, (T*)y);
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
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

extern "C" int dm4d_add_layernorm(const void* x, const void* delta, int32_t delta_bcast_rows, const float* gamma, const float* beta,
                                  int64_t M, int32_t C, float eps, int32_t dtype, void* x_out, void* y, void* stream) {
    if (!x || !gamma || !beta || !y || M <= 0 || C <= 0 || C % VEC || C > 32 * VEC * LN_MAXV || (delta && !x_out) || delta_bcast_rows < 0) {
        dm4d_set_error("dm4d_add_layernorm: bad argument (C %% 4 == 0, C <= %d, x_out required with delta)", 32 * VEC * LN_MAXV);
        return DM4D_EINVAL;
    }
    if (dtype == DM4D_F16) return launch_add_layernorm<__half>(M, C, x, delta, delta_bcast_rows, gamma, beta, eps, x_out, y, (cudaStream_t)stream);
    if (dtype == DM4D_F32) return launch_add_layernorm<float>(M, C, x, delta, delta_bcast_rows, gamma, beta, eps, x_out, y, (cudaStream_t)stream);
    dm4d_set_error("dm4d_add_layernorm: dtype must be DM4D_F16 or DM4D_F32");
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
