// host_probe.c - what the host side of the e2e path can do on this box: multi-threaded streaming-store bandwidth into
// one buffer (what a host-side adjacency expander is bound by). usage: host_probe <MB> <max_threads>
#define _GNU_SOURCE
#include <emmintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { char* p; size_t n; int nt; } job;
static void* worker(void* a) {
    job* j = (job*)a;
    if (j->nt) {
        __m128i v = _mm_set1_epi32(0x3f800000);
        for (size_t k = 0; k + 64 <= j->n; k += 64) {
            _mm_stream_si128((__m128i*)(j->p + k), v); _mm_stream_si128((__m128i*)(j->p + k + 16), v);
            _mm_stream_si128((__m128i*)(j->p + k + 32), v); _mm_stream_si128((__m128i*)(j->p + k + 48), v);
        }
        _mm_sfence();
    } else memset(j->p, 1, j->n);
    return 0;
}
static double now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(int argc, char** argv) {
    size_t mb = argc > 1 ? atol(argv[1]) : 108; int maxt = argc > 2 ? atoi(argv[2]) : 32;
    size_t n = mb << 20; char* buf = aligned_alloc(4096, n); memset(buf, 0, n);
    for (int nt = 0; nt < 2; ++nt)
        for (int T = 1; T <= maxt; T *= 2) {
            pthread_t th[256]; job jb[256]; double best = 1e9;
            for (int rep = 0; rep < 5; ++rep) {
                double t0 = now();
                for (int t = 0; t < T; ++t) { size_t per = (n / T) & ~(size_t)63; jb[t] = (job){buf + t * per, per, nt}; pthread_create(&th[t], 0, worker, &jb[t]); }
                for (int t = 0; t < T; ++t) pthread_join(th[t], 0);
                double dt = now() - t0; if (dt < best) best = dt;
            }
            printf("%s threads=%d  %.2f ms  %.1f GB/s\n", nt ? "stream" : "memset", T, best * 1e3, n / best / 1e9);
        }
    return 0;
}
