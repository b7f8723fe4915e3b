// lsm_kernels.cu - translation unit of the step kernels for sm_100a.
//
//   lsm_kernel_spec.cuh     kernels specialised at compile time on (dynamics, N, L) - the fast path
//   lsm_kernel_generic.cuh  run-time N / L fallback for every other configuration
//   lsm_step_common.cuh     dynamics, HJ filter resolution, grid interpolation shared by both
//
// Work decomposition ("warp per env group"): a warp owns EPW consecutive environments, G = next power
// of two >= N lanes per environment. Per-agent phases run one lane per agent; the graph observation
// (pairwise distances, radius-limited adjacency, node features) is built in the warp's shared-memory
// slice and written with all 32 lanes. There is no block-level barrier: warps only __syncwarp(), so
// one warp's stores overlap another warp's HJ-grid gathers.
//
// The reference mutates goal counters / done flags / velocities agent by agent WHILE it emits
// observations (multiagent/environment.py:979-1029). Each agent's own update depends only on its
// own state, so every lane computes its agent's pre- and post-update state in parallel and an
// observer i selects "post" for agents <= i and "pre" for agents > i.
#include "lsm_kernel_generic.cuh"
#include "lsm_kernel_spec.cuh"
#include "lsm_host.h"

#include <cstdlib>

namespace lsm {

// ---------------------------------------------------------------------------------------------
// host-side launch helpers (used by lsm_capi.cu)
// ---------------------------------------------------------------------------------------------
#define LSM_SPEC_LIST(X)                         \
    X(LSM_DYN_DOUBLE_INTEGRATOR, 8, 2)           \
    X(LSM_DYN_DOUBLE_INTEGRATOR, 3, 2)           \
    X(LSM_DYN_DOUBLE_INTEGRATOR, 32, 2)          \
    X(LSM_DYN_AIRTAXI, 10, 2)

constexpr int kSpecBlock = 128;

// register budgets: MINB blocks of 128 threads per SM -> 65536 / (128 * MINB) registers per thread
static int spec_minb() {
    const char* e = std::getenv("LSM_MINB");
    const int v = e ? std::atoi(e) : 2;
    return (v == 3 || v == 4) ? v : 2;
}

static const void* generic_ptr(int dynamics) {
    return dynamics == LSM_DYN_DOUBLE_INTEGRATOR ? (const void*)lsm_generic_kernel<LSM_DYN_DOUBLE_INTEGRATOR>
                                                 : (const void*)lsm_generic_kernel<LSM_DYN_AIRTAXI>;
}

static const void* spec_ptr(int dynamics, int N, int L, int* bytes_per_env, int* stage_bytes = nullptr) {
    const int minb = spec_minb();
#define X(DYN_, N_, L_)                                                                           \
    if (dynamics == DYN_ && N == N_ && L == L_) {                                                 \
        *bytes_per_env = (int)sizeof(EnvShared<DYN_, N_, L_>);                                    \
        if (stage_bytes) *stage_bytes = 2 * 32 * (DYN_ == LSM_DYN_DOUBLE_INTEGRATOR ? 10 : 11) * 4;   \
        if (minb == 2) return (const void*)lsm_spec_kernel<DYN_, N_, L_, kSpecBlock, 2>;          \
        if (minb == 3) return (const void*)lsm_spec_kernel<DYN_, N_, L_, kSpecBlock, 3>;          \
        return (const void*)lsm_spec_kernel<DYN_, N_, L_, kSpecBlock, 4>;                         \
    }
    LSM_SPEC_LIST(X)
#undef X
    return nullptr;
}

bool spec_available(int dynamics, int N, int L, int* bytes_per_env, int* block_threads, int* stage_bytes) {
    *block_threads = kSpecBlock;
    return spec_ptr(dynamics, N, L, bytes_per_env, stage_bytes) != nullptr;
}

cudaError_t kernel_prepare(int dynamics, int N, int L, bool spec, int smem_bytes, int block_threads, int* regs,
                           int* blocks_per_sm) {
    int dummy = 0;
    const void* fn = spec ? spec_ptr(dynamics, N, L, &dummy) : generic_ptr(dynamics);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, fn);
    if (e != cudaSuccess) return e;
    *regs = fa.numRegs;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fn, block_threads, smem_bytes);
}

cudaError_t upload_magnetic_tables(const double* cos_tab, const double* sin_tab) {
    cudaError_t e = cudaMemcpyToSymbol(c_mag_cos, cos_tab, sizeof(double) * kMagSegments);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_mag_sin, sin_tab, sizeof(double) * kMagSegments);
}

cudaError_t kernel_launch(const KParams& kp, bool spec, int grid_blocks, int block_threads, int smem_bytes,
                          cudaStream_t stream, const void* persist_ptr, size_t persist_bytes) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid_blocks, 1, 1);
    cfg.blockDim = dim3((unsigned)block_threads, 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    int nattr = 0;
    if (persist_ptr != nullptr && persist_bytes > 0) {
        // keep the HJ value grid resident in L2 while the observation stream flows through it
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = const_cast<void*>(persist_ptr);
        attr[0].val.accessPolicyWindow.num_bytes = persist_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        nattr = 1;
    }
    cfg.attrs = attr;
    cfg.numAttrs = nattr;
    void* args[] = { (void*)&kp };
    int dummy = 0;
    const void* fn = spec ? spec_ptr(kp.c.dynamics, kp.N, kp.L, &dummy) : generic_ptr(kp.c.dynamics);
    return cudaLaunchKernelExC(&cfg, fn, args);
}

}  // namespace lsm
