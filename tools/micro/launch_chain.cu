// How fast can T host threads, one CUDA stream each, push chains of small dependent
// kernels through one GPU?  (diagnostics for the many-instance encoder: every picture
// is a chain of ~60-110 stream-ordered operations per instance)
//   launch_chain T N [busy_us] [ctas] [param_bytes: 0 | 1024 | 3800] [memset_every]
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
struct P1k { char b[1024]; };
struct P4k { char b[3800]; };
__global__ void k_busy(long long cycles) { long long t0 = clock64(); while (clock64() - t0 < cycles) { } }
__global__ void k_busy1k(long long cycles, P1k p) { long long t0 = clock64(); while (clock64() - t0 < cycles) { } if (p.b[5] == 77 && cycles < 0) printf("x"); }
__global__ void k_busy4k(long long cycles, P4k p) { long long t0 = clock64(); while (clock64() - t0 < cycles) { } if (p.b[5] == 77 && cycles < 0) printf("x"); }
struct Arg { int n, ctas, pb, ms; long long cyc; double secs; };
static void *worker(void *p) {
    Arg *a = (Arg *) p; cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    static P1k p1; static P4k p4; char *buf; cudaMalloc(&buf, 1 << 16);
    for (int rep = 0; rep < 2; rep++) {
        auto t0 = std::chrono::steady_clock::now();
        int n = rep ? a->n : 50;
        for (int i = 0; i < n; i++) {
            if (a->pb >= 3000) k_busy4k<<<a->ctas, 128, 0, s>>>(a->cyc, p4);
            else if (a->pb >= 1000) k_busy1k<<<a->ctas, 128, 0, s>>>(a->cyc, p1);
            else k_busy<<<a->ctas, 128, 0, s>>>(a->cyc);
            if (a->ms && (i % a->ms) == 0) cudaMemsetAsync(buf, 0, 4096, s);
        }
        cudaStreamSynchronize(s);
        a->secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return NULL;
}
int main(int argc, char **argv) {
    int T = atoi(argv[1]), N = atoi(argv[2]); double us = argc > 3 ? atof(argv[3]) : 0; int ctas = argc > 4 ? atoi(argv[4]) : 1;
    int pb = argc > 5 ? atoi(argv[5]) : 0, ms = argc > 6 ? atoi(argv[6]) : 0;
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    cudaFree(0);
    pthread_t th[256]; Arg a[256];
    for (int i = 0; i < T; i++) { a[i].n = N; a[i].ctas = ctas; a[i].pb = pb; a[i].ms = ms; a[i].cyc = (long long) (us * 1965.0); pthread_create(&th[i], NULL, worker, &a[i]); }
    double mx = 0; for (int i = 0; i < T; i++) { pthread_join(th[i], NULL); if (a[i].secs > mx) mx = a[i].secs; }
    printf("threads %3d  kernel %5.1f us x %d CTAs  params %4d B  memset every %d: %7.2f us per kernel per stream, %8.0f kernels/s total\n", T, us, ctas, pb, ms, 1e6 * mx / N, T * N / mx);
    return 0;
}
