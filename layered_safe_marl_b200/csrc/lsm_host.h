// lsm_host.h - host-visible launch helpers implemented next to the kernels (lsm_kernels.cu).
#pragma once
#include "lsm_device.cuh"

namespace lsm {
bool spec_available(int dynamics, int N, int L, int* bytes_per_env, int* block_threads, int* stage_bytes);
cudaError_t kernel_prepare(int dynamics, int N, int L, bool spec, int smem_bytes, int block_threads, int* regs,
                           int* blocks_per_sm);
cudaError_t upload_magnetic_tables(const double* cos_tab, const double* sin_tab);
cudaError_t kernel_launch(const KParams& kp, bool spec, int grid_blocks, int block_threads, int smem_bytes,
                          cudaStream_t stream, const void* persist_ptr, size_t persist_bytes);
}  // namespace lsm
