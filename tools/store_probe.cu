// store_probe.cu - how fast can 107 MB (cfg2) / 852 MB (cfg3) of output be WRITTEN on a B200, depending on the address
// pattern? Background (DESIGN.md, round-2 experiments): the emit kernel's bulk copies alone take 15 % longer than the same
// bytes written as one grid-strided compact window. This probe separates the candidates: region size per block, the
// alternation between two output arrays, the block -> environment mapping.
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/store_probe tools/store_probe.cu
//   run:   tools/store_probe <envs> <bytesA per env> <bytesB per env>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// mode 0: compact grid-strided window over A, then over B
// mode 1: block b takes envs b, b + G, ...: A[e] then B[e]                     (the emit kernel's pattern)
// mode 2: like 1, but A only, in regions of `region` bytes (B's bytes appended to A)
// mode 3: like 1, but block b takes a RUN of consecutive envs
// mode 4: like 1, A[e] for all envs of the block first, then B[e] for all of them
__global__ void __launch_bounds__(128) k(float4* A, float4* B, long long envs, long long qa, long long qb, int mode, long long region_q) {
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    const long long T = blockDim.x, G = gridDim.x, tid = threadIdx.x, b = blockIdx.x;
    if (mode == 0) {
        for (long long i = b * T + tid; i < envs * qa; i += G * T) A[i] = v;
        for (long long i = b * T + tid; i < envs * qb; i += G * T) B[i] = v;
    } else if (mode == 1) {
        for (long long e = b; e < envs; e += G) {
            for (long long i = tid; i < qa; i += T) A[e * qa + i] = v;
            for (long long i = tid; i < qb; i += T) B[e * qb + i] = v;
        }
    } else if (mode == 2) {
        const long long total = envs * (qa + qb), regions = (total + region_q - 1) / region_q;
        for (long long r = b; r < regions; r += G) {
            const long long base = r * region_q, lim = (base + region_q < total) ? region_q : total - base;
            for (long long i = tid; i < lim; i += T) A[base + i] = v;
        }
    } else if (mode == 3) {
        const long long per = (envs + G - 1) / G;
        for (long long e = b * per; e < (b + 1) * per && e < envs; ++e) {
            for (long long i = tid; i < qa; i += T) A[e * qa + i] = v;
            for (long long i = tid; i < qb; i += T) B[e * qb + i] = v;
        }
    } else if (mode == 5) {
        // like mode 1, but every warp-wide store covers whole 128-byte lines: the head of the region (up to the next line
        // boundary) is written by a few lanes first
        for (long long e = b; e < envs; e += G) {
            for (int arr = 0; arr < 2; ++arr) {
                float4* base = arr == 0 ? A + e * qa : B + e * qb;
                const long long q = arr == 0 ? qa : qb;
                const long long head = ((8 - (((unsigned long long)base >> 4) & 7)) & 7);      // float4s up to the next 128-B line
                if (tid < head) base[tid] = v;
                for (long long i = head + tid; i < q; i += T) base[i] = v;
            }
        }
    } else {
        for (long long e = b; e < envs; e += G) for (long long i = tid; i < qa; i += T) A[e * qa + i] = v;
        for (long long e = b; e < envs; e += G) for (long long i = tid; i < qb; i += T) B[e * qb + i] = v;
    }
}

int main(int argc, char** argv) {
    const long long envs = argc > 1 ? atoll(argv[1]) : 4096, ba = argc > 2 ? atoll(argv[2]) : 18432, bb = argc > 3 ? atoll(argv[3]) : 7680;
    const long long qa = ba / 16, qb = bb / 16;
    float4 *A, *B; float* F;
    CK(cudaMalloc(&A, envs * (ba + bb))); CK(cudaMalloc(&B, envs * bb)); CK(cudaMalloc(&F, 256ll << 20));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    auto run = [&](const char* name, int mode, int bps, long long region) {
        std::vector<float> t;
        for (int rep = 0; rep < 12; ++rep) {
            CK(cudaMemsetAsync(F, 0, 256ll << 20));
            CK(cudaEventRecord(e0));
            k<<<sms * bps, 128>>>(A, B, envs, qa, qb, mode, region / 16);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep >= 2) t.push_back(ms);
        }
        std::sort(t.begin(), t.end());
        const double med = t[t.size() / 2], gb = envs * (ba + bb) / 1e9;
        printf("%-46s blocks/SM %d  %8.1f us  %6.2f TB/s\n", name, bps, med * 1000.0, gb / med);
    };
    printf("envs %lld, A %lld B/env, B %lld B/env, total %.1f MB, flushed (256 MiB memset) before every launch\n", envs, ba, bb, envs * (ba + bb) / 1e6);
    for (int bps : {5, 10, 16}) run("0 compact window", 0, bps, 0);
    for (int bps : {2, 5, 10, 16}) run("1 per env, strided envs, A then B", 1, bps, 0);
    run("3 per env, runs of consecutive envs", 3, 5, 0);
    run("4 strided envs, all A first, then all B", 4, 5, 0);
    run("5 per env, strided, 128-B aligned warp segments", 5, 5, 0);
    { char nm[64]; snprintf(nm, sizeof nm, "2 one array, regions of %lld B (= A)", ba); run(nm, 2, 5, ba); }
    for (long long r : {2048ll, 8192ll, 16384ll, 32768ll, 65536ll, 262144ll, 1048576ll}) {
        char nm[64]; snprintf(nm, sizeof nm, "2 one array, regions of %lld B", r); run(nm, 2, 5, r);
    }
    return 0;
}
