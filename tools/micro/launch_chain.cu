// How fast can T host threads, one CUDA stream each, push chains of small dependent
// kernels through one GPU?  (diagnostics for the many-instance encoder: every picture
// is a chain of ~110 stream-ordered operations per instance)
//   launch_chain T N [busy_us] [ctas] [hog_ctas]
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
__global__ void k_busy(long long cycles) { long long t0 = clock64(); while (clock64() - t0 < cycles) { } }
__global__ void k_hog(volatile int *stop) { while (!*stop) __nanosleep(500); }
struct Arg { int n, ctas; long long cyc; double secs; };
static void *worker(void *p) {
    Arg *a = (Arg *) p; cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    for (int i = 0; i < 50; i++) k_busy<<<a->ctas, 128, 0, s>>>(a->cyc);
    cudaStreamSynchronize(s);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < a->n; i++) k_busy<<<a->ctas, 128, 0, s>>>(a->cyc);
    cudaStreamSynchronize(s);
    a->secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return NULL;
}
int main(int argc, char **argv) {
    int T = atoi(argv[1]), N = atoi(argv[2]); double us = argc > 3 ? atof(argv[3]) : 0; int ctas = argc > 4 ? atoi(argv[4]) : 1;
    int hog = argc > 5 ? atoi(argv[5]) : 0;
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    cudaFree(0);
    int *stop; cudaMallocManaged(&stop, 4); *stop = 0; cudaStream_t hs; cudaStreamCreateWithFlags(&hs, cudaStreamNonBlocking);
    int *dstop; cudaMalloc(&dstop, 4); cudaMemset(dstop, 0, 4);
    if (hog) k_hog<<<hog, 256, 0, hs>>>(dstop);
    pthread_t th[256]; Arg a[256];
    for (int i = 0; i < T; i++) { a[i].n = N; a[i].ctas = ctas; a[i].cyc = (long long) (us * 1965.0); pthread_create(&th[i], NULL, worker, &a[i]); }
    double mx = 0; for (int i = 0; i < T; i++) { pthread_join(th[i], NULL); if (a[i].secs > mx) mx = a[i].secs; }
    int one = 1; cudaMemcpyAsync(dstop, &one, 4, cudaMemcpyHostToDevice, 0); cudaDeviceSynchronize();
    printf("threads %3d  kernel %5.1f us x %d CTAs  hog %4d CTAs: %7.2f us per kernel per stream, %8.0f kernels/s total\n", T, us, ctas, hog, 1e6 * mx / N, T * N / mx);
    return 0;
}
