// lsm_host.h - host-visible launch helpers implemented next to the kernels (lsm_kernels.cu).
#pragma once
#include "lsm_device.cuh"

namespace lsm {
struct SpecGeometry {
    int rec_bytes;        // sizeof(EmitRec): per-env record between the agent and the emit kernel
    int scratch_bytes;    // sizeof(AgentScratch): per-env physics scratch of the agent kernel
    int agent_block, pair_block;
    int emit_smem, emit_threads;
};
bool spec_available(int dynamics, int N, int L, int O, SpecGeometry* g);
// agent kernel (specialised) or the fused generic kernel
cudaError_t kernel_prepare(int dynamics, int N, int L, int O, bool spec, int smem_bytes, int block_threads, int* regs,
                           int* blocks_per_sm);
cudaError_t spec_prepare_aux(int dynamics, int N, int L, int O, int* emit_regs, int* emit_blocks_per_sm, int* pair_regs);
cudaError_t upload_magnetic_tables(const double* cos_tab, const double* sin_tab);
cudaError_t kernel_launch(const KParams& kp, bool spec, int grid_blocks, int block_threads, int smem_bytes,
                          cudaStream_t stream, const void* persist_ptr, size_t persist_bytes);
cudaError_t spec_launch_pair(const KParams& kp, cudaStream_t stream, const void* persist_ptr, size_t persist_bytes);
cudaError_t spec_launch_emit(const KParams& kp, cudaStream_t stream, const void* persist_ptr, size_t persist_bytes, bool reserve_pair, bool pie);
cudaError_t spec_emit_blocks_per_sm(int dynamics, int N, int L, int O, bool reserve_pair, bool pie, int* out, int* regs);
int edge_count_envs_per_block(long long envs);
cudaError_t spec_launch_edge_count(const KParams& kp, cudaStream_t stream);
cudaError_t world_graph_launch(const KParams& kp, int32_t* counts, long long* offsets, long long* edge_index, double* edge_weight,
                               long long capacity, cudaStream_t stream);
cudaError_t pack_grid_launch(const GridDev& g, float* packed, long long cells);
cudaError_t episode_stats_launch(const double* ep_info, long long n, double* out, cudaStream_t stream);
cudaError_t rollout_insert_launch(const float* obs, const uint8_t* done, float* share_obs, float* masks, float* active_masks,
                                  long long n, int N, int D, cudaStream_t stream);
cudaError_t math_eval_launch(int op, const double* a, const double* b, double* out, long long n, cudaStream_t stream);
cudaError_t pad_grads_launch(const float* grads, float* grads8, long long cells);
cudaError_t edge_list_launch(const float* adj, int32_t* counts, long long* offsets, long long* edge_index, float* edge_attr,
                             long long num_graphs, int E, long long capacity, cudaStream_t stream);
}  // namespace lsm
