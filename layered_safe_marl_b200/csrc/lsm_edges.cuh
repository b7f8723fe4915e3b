// lsm_edges.cuh - compacted COO edge list of the step's adjacency output, in the order the policy's GNN expects
// (SURVEY.md section 8f, row N2).
//
// Replaces `TransformerConvNet.process_adj` (reference onpolicy/algorithms/utils/gnn.py:376-407), which runs on every
// policy forward: adj (B, E, E) -> adj.nonzero() in row-major (graph, row, col) order ->
//     edge_index = [graph * E + row, graph * E + col]   (2, nnz) int64
//     edge_attr  = adj[graph, row, col]                  (nnz, 1) float32
// with B = num_envs * N graphs (the runner concatenates the agent axis into the batch, graph_mpe_runner.py:398-410).
//
// Three launches, all HBM / L2 streaming integer work (no tensor cores):
//   lsm_edge_count_kernel   one warp per graph: popcount of the non-zero entries -> counts[g]
//   lsm_edge_scan_kernel    one block: exclusive prefix sum of counts -> offsets[g], offsets[B] = nnz
//   lsm_edge_fill_kernel    one warp per graph: row-major walk in chunks of 32 entries, ballot + prefix popcount
//                           compaction, writes starting at offsets[g]
// The dense matrices are read twice; they were just written by the emit kernel and mostly sit in the 126 MB L2.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lsm {

__global__ void __launch_bounds__(256) lsm_edge_count_kernel(const float* __restrict__ adj, int32_t* __restrict__ counts,
                                                             long long num_graphs, int EE) {
    const long long g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= num_graphs) return;
    const int lane = threadIdx.x & 31;
    const float* a = adj + g * EE;
    int c = 0;
    if ((EE & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) & 15u) == 0)) {
        const float4* a4 = reinterpret_cast<const float4*>(a);
        for (int k = lane; k < EE / 4; k += 32) {
            const float4 v = a4[k];
            c += (v.x != 0.0f) + (v.y != 0.0f) + (v.z != 0.0f) + (v.w != 0.0f);
        }
    } else {
        for (int k = lane; k < EE; k += 32) c += a[k] != 0.0f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) counts[g] = c;
}

// single block: every thread owns a contiguous run, block-level scan of the run totals in shared memory
__global__ void __launch_bounds__(1024) lsm_edge_scan_kernel(const int32_t* __restrict__ counts, long long* __restrict__ offsets,
                                                             long long num_graphs) {
    __shared__ long long part[1024];
    const int t = threadIdx.x, T = blockDim.x;
    const long long per = (num_graphs + T - 1) / T;
    const long long lo = (long long)t * per, hi = (lo + per) < num_graphs ? (lo + per) : num_graphs;
    long long s = 0;
    for (long long k = lo; k < hi; ++k) s += counts[k];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < T; o <<= 1) {       // Hillis-Steele inclusive scan over the T run totals
        const long long v = (t >= o) ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    long long run = part[t] - s;            // exclusive prefix of this thread's run
    for (long long k = lo; k < hi; ++k) { offsets[k] = run; run += counts[k]; }
    if (t == T - 1) offsets[num_graphs] = part[T - 1];
}

__global__ void __launch_bounds__(256) lsm_edge_fill_kernel(const float* __restrict__ adj, const long long* __restrict__ offsets,
                                                            long long* __restrict__ edge_index, float* __restrict__ edge_attr,
                                                            long long num_graphs, int E, long long capacity) {
    const long long g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= num_graphs) return;
    const int lane = threadIdx.x & 31;
    const int EE = E * E;
    const float* a = adj + g * EE;
    long long pos = offsets[g];
    const long long node0 = g * E;
    long long* src = edge_index;
    long long* dst = edge_index + capacity;
    for (int k0 = 0; k0 < EE; k0 += 32) {
        const int k = k0 + lane;
        const float v = k < EE ? a[k] : 0.0f;
        const bool nz = v != 0.0f;
        const unsigned m = __ballot_sync(0xffffffffu, nz);
        if (nz) {
            const long long p = pos + __popc(m & ((1u << lane) - 1u));
            if (p < capacity) {
                const int row = k / E, col = k - row * E;
                src[p] = node0 + row; dst[p] = node0 + col; edge_attr[p] = v;
            }
        }
        pos += __popc(m);
    }
}

}  // namespace lsm
