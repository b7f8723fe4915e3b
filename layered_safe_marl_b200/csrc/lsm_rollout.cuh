// lsm_rollout.cuh - the per-step bookkeeping of the rollout buffer in ONE launch.
//
// Replaces what `GMPERunner.insert` + `GraphReplayBuffer.insert` do on the host between two env.step calls
// (reference onpolicy/runner/shared/graph_mpe_runner.py:437-487, onpolicy/utils/graph_buffer.py:168-250):
//     masks[dones] = 0;  active_masks[dones] = 0;  active_masks[all agents of the env done] = 1
//     share_obs[env, agent] = concatenation of every agent's obs of that env (use_centralized_V)
// The observations / graphs / rewards themselves are written in place by the step kernels (zero-copy slots); the
// agent-id slots are constant and filled once. Pure streaming elementwise work: one thread per share_obs element.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lsm {

__global__ void __launch_bounds__(256) lsm_rollout_insert_kernel(const float* __restrict__ obs, const uint8_t* __restrict__ done,
                                                                 float* __restrict__ share_obs, float* __restrict__ masks,
                                                                 float* __restrict__ active_masks, long long n, int N, int D) {
    const int ND = N * D;
    const long long total = n * (long long)N * ND;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long ea = t / ND;                 // (env, agent)
        const int k = (int)(t - ea * ND);
        const long long env = ea / N;
        if (share_obs != nullptr) share_obs[t] = obs[env * ND + k];
        if (k == 0) {
            const int agent = (int)(ea - env * N);
            bool all_done = true;
            for (int a = 0; a < N; ++a) all_done = all_done && done[env * N + a] != 0;
            const bool d = done[env * N + agent] != 0;
            masks[ea] = d ? 0.0f : 1.0f;
            active_masks[ea] = all_done ? 1.0f : (d ? 0.0f : 1.0f);
        }
    }
}

// Sums of the 8 episode-summary columns over the shard's environments + the environment count: the 9 doubles the
// runner's log-time all-reduce exchanges. ONE block, fixed reduction order (deterministic).
__global__ void __launch_bounds__(256) lsm_episode_stats_kernel(const double* __restrict__ ep_info, long long n, int cols, double* __restrict__ out) {
    __shared__ double s_part[8][16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = 0; c < cols; ++c) {
        double acc = 0.0;
        for (long long e = threadIdx.x; e < n; e += blockDim.x) acc += ep_info[e * cols + c];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) s_part[warp][c] = acc;
    }
    __syncthreads();
    if (threadIdx.x < cols) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w][threadIdx.x];
        out[threadIdx.x] = t;
    }
    if (threadIdx.x == 0) out[cols] = (double)n;
}

}  // namespace lsm
