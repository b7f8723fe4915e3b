// lsm_host.h - host-visible launch helpers implemented next to the kernel (lsm_kernels.cu).
#pragma once
#include "lsm_device.cuh"

namespace lsm {
cudaError_t fused_kernel_prepare(int dynamics, int smem_bytes, int block_threads, int* regs, int* blocks_per_sm);
cudaError_t upload_magnetic_tables(const double* cos_tab, const double* sin_tab);
cudaError_t fused_kernel_launch(const KParams& kp, int grid_blocks, int block_threads, int smem_bytes,
                                cudaStream_t stream, const void* persist_ptr, size_t persist_bytes);
}  // namespace lsm
