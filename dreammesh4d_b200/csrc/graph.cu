// Deformation-graph construction on the GPU: K nearest control nodes of every mesh vertex
// and of every node.
//
// Replaces the per-vertex Open3D KD-tree queries (one Python call + H2D copy per vertex) of
// DynamicSuGaRModel.build_deformation_graph, mode "eucdisc"
// (custom/threestudio-dreammesh4d/geometry/dynamic_sugar.py:765-790: search_knn_vector_3d(vertex, K) -> node indices
// [1] and SQUARED distances [2]; node-node connectivity = K+1 nearest nodes of a node minus itself, :768-776).
// Brute force: one thread per query, the nodes staged through shared memory in tiles, a sorted K-list per thread
// in registers.  V x M squared distances (5e7 at C3) — HBM traffic is 12 V + 12 M in, 8 K V out, the kernel is
// compute/issue bound on the insertion test and takes tens of microseconds; the reference takes seconds.
//
// Arithmetic spec (bit-exact neighbour sets): d2 = (dx*dx + dy*dy) + dz*dz with every product and sum a separately
// rounded binary32 operation (no FMA contraction); neighbours ordered by ascending (d2, node index).
#include "raster_internal.cuh"

namespace {

constexpr int KNN_TILE = 1024;

template <int KK>
__global__ void __launch_bounds__(DM4D_BLOCK) knn_kernel(const float* __restrict__ queries, int n_queries,
                                                         const float* __restrict__ nodes, int n_nodes, int k,
                                                         int32_t* __restrict__ idx, float* __restrict__ sqdist) {
    __shared__ float sn[KNN_TILE * 3];
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (q < n_queries) { qx = queries[3 * q]; qy = queries[3 * q + 1]; qz = queries[3 * q + 2]; }
    float bd[KK];
    int bi[KK];
#pragma unroll
    for (int i = 0; i < KK; ++i) { bd[i] = __int_as_float(0x7f800000); bi[i] = -1; }
    for (int base = 0; base < n_nodes; base += KNN_TILE) {
        const int cnt = min(KNN_TILE, n_nodes - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 3; i += blockDim.x) sn[i] = nodes[(size_t)base * 3 + i];
        __syncthreads();
        if (q >= n_queries) continue;
        for (int j = 0; j < cnt; ++j) {
            const float dx = __fsub_rn(qx, sn[3 * j]), dy = __fsub_rn(qy, sn[3 * j + 1]), dz = __fsub_rn(qz, sn[3 * j + 2]);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (d < bd[KK - 1]) {            // strict: of equal distances the lower node index (seen first) stays
                // insert into the ascending list (unrolled bubble from the tail)
                bd[KK - 1] = d; bi[KK - 1] = base + j;
#pragma unroll
                for (int i = KK - 1; i > 0; --i) {
                    if (bd[i] < bd[i - 1]) {
                        const float td = bd[i]; bd[i] = bd[i - 1]; bd[i - 1] = td;
                        const int ti = bi[i]; bi[i] = bi[i - 1]; bi[i - 1] = ti;
                    }
                }
            }
        }
    }
    if (q >= n_queries) return;
#pragma unroll
    for (int i = 0; i < KK; ++i)
        if (i < k) {
            idx[(size_t)q * k + i] = bi[i];
            if (sqdist) sqdist[(size_t)q * k + i] = bd[i];
        }
}

}  // namespace

extern "C" int dm4d_graph_knn(const float* queries, int32_t n_queries, const float* nodes, int32_t n_nodes, int32_t k,
                              int32_t* idx, float* sqdist, void* stream) {
    if (!queries || !nodes || !idx || n_queries < 0 || n_nodes <= 0 || k <= 0 || k > 17 || k > n_nodes) {
        dm4d_set_error("dm4d_graph_knn: bad argument (n_queries=%d n_nodes=%d k=%d; 1 <= k <= min(17, n_nodes))",
                       n_queries, n_nodes, k);
        return DM4D_EINVAL;
    }
    if (n_queries == 0) return DM4D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)((n_queries + DM4D_BLOCK - 1) / DM4D_BLOCK);
    {
        KernelTimer kt(DM4D_K_GRAPH_KNN, s);
        if (k <= 5) knn_kernel<5><<<blocks, DM4D_BLOCK, 0, s>>>(queries, n_queries, nodes, n_nodes, k, idx, sqdist);
        else if (k <= 9) knn_kernel<9><<<blocks, DM4D_BLOCK, 0, s>>>(queries, n_queries, nodes, n_nodes, k, idx, sqdist);
        else knn_kernel<17><<<blocks, DM4D_BLOCK, 0, s>>>(queries, n_queries, nodes, n_nodes, k, idx, sqdist);
    }
    DM4D_CUDA_CHECK(cudaGetLastError());
    return DM4D_OK;
}
