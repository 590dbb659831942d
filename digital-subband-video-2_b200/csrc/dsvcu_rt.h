/*
 * dsvcu_rt.h -- runtime shim for the DSV2 sm_100a kernels.
 *
 * Product build (nvcc, sm_100a): thin wrappers over the CUDA runtime.
 *
 * Test-only build (-DDSVCU_EMU, g++): the SAME kernel sources are compiled as
 * plain C++ and every launch is executed as a sequential loop over blocks with
 * blockDim = (1,1,1).  This exists so the kernel arithmetic/indexing can be
 * exercised by the CPU-only test suite in a container that has no GPU.  It is
 * never linked into libdsv2cuda.so and is not a fallback: the product library
 * is CUDA-only and fails loudly when no device is present.
 *
 * Kernels therefore follow two rules:
 *   - work is distributed with strided loops over (threadIdx.x, blockDim.x);
 *   - phases that communicate through shared memory are separated by
 *     __syncthreads(), and are race-free inside a phase.
 */
#ifndef DSVCU_RT_H
#define DSVCU_RT_H

#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>

#ifndef DSVCU_EMU
/* ------------------------------------------------------------------ CUDA */
#include <cuda_runtime.h>

typedef cudaStream_t dsvcu_stream_t;
typedef cudaEvent_t dsvcu_event_t;

#define DSVCU_KERNEL __global__ static
#define DSVCU_DEV __device__ __forceinline__
#define DSVCU_HD __host__ __device__ __forceinline__
#define DSVCU_SHARED __shared__
#define DSVCU_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; type *name = (type *) name##_raw_
#define DSVCU_LAUNCH(kern, grid, block, smem, stream, ...) \
    kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define DSVCU_SYNC() __syncthreads()
#define DSVCU_SYNCWARP() __syncwarp()
#define DSVCU_FENCE() __threadfence()

#define dsvcu_malloc(pp, n) cudaMalloc((void **) (pp), (n))
#define dsvcu_free_dev(p) cudaFree(p)
#define dsvcu_malloc_host(pp, n) cudaMallocHost((void **) (pp), (n))
#define dsvcu_free_host(p) cudaFreeHost(p)
#define dsvcu_memset_async(p, v, n, s) cudaMemsetAsync((p), (v), (n), (s))
#define dsvcu_h2d_async(d, h, n, s) cudaMemcpyAsync((d), (h), (n), cudaMemcpyHostToDevice, (s))
#define dsvcu_d2h_async(h, d, n, s) cudaMemcpyAsync((h), (d), (n), cudaMemcpyDeviceToHost, (s))
#define dsvcu_d2d_async(d, s0, n, s) cudaMemcpyAsync((d), (s0), (n), cudaMemcpyDeviceToDevice, (s))
/* frame upload / download: the far side may be host OR device memory (unified addressing) */
#define dsvcu_h2d_2d_async(d, dp, h, hp, w, ht, s) cudaMemcpy2DAsync((d), (dp), (h), (hp), (w), (ht), cudaMemcpyDefault, (s))
#define dsvcu_d2h_2d_async(h, hp, d, dp, w, ht, s) cudaMemcpy2DAsync((h), (hp), (d), (dp), (w), (ht), cudaMemcpyDefault, (s))
#define dsvcu_d2d_2d_async(d, dp, s0, sp, w, ht, s) cudaMemcpy2DAsync((d), (dp), (s0), (sp), (w), (ht), cudaMemcpyDeviceToDevice, (s))
#define dsvcu_stream_sync(s) cudaStreamSynchronize(s)
#define dsvcu_memset_2d_async(p, pitch, v, w, h, s) cudaMemset2DAsync((p), (pitch), (v), (w), (h), (s))

#else
/* ------------------------------------------------------- host emulation */
#include <algorithm>
#include <cstdlib>

struct dsvcu_dim3 {
    unsigned x, y, z;
    dsvcu_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef dsvcu_dim3 dim3;
typedef int dsvcu_stream_t;
typedef int dsvcu_event_t;
typedef int cudaError_t;
#define cudaSuccess 0

extern dsvcu_dim3 threadIdx, blockIdx, blockDim, gridDim;
extern unsigned char *dsvcu_emu_smem;
extern size_t dsvcu_emu_smem_size;

#define DSVCU_KERNEL static
#define DSVCU_DEV static inline
#define DSVCU_HD static inline
#define DSVCU_SHARED static
#define DSVCU_DYN_SMEM(type, name) type *name = (type *) dsvcu_emu_smem
#define DSVCU_SYNC() ((void) 0)
#define DSVCU_SYNCWARP() ((void) 0)
#define DSVCU_FENCE() ((void) 0)
#define __restrict__
#define __align__(n) alignas(n)
#define __launch_bounds__(...)

static inline void
dsvcu_emu_need_smem(size_t n)
{
    if (n > dsvcu_emu_smem_size) {
        free(dsvcu_emu_smem);
        dsvcu_emu_smem = (unsigned char *) calloc(1, n + 64);
        dsvcu_emu_smem_size = n;
    }
}

#define DSVCU_LAUNCH(kern, grid, block, smem, stream, ...)                     \
    do {                                                                       \
        dsvcu_dim3 g_ = dsvcu_dim3(grid);                                      \
        dsvcu_emu_need_smem(smem);                                             \
        gridDim = g_;                                                          \
        blockDim = dsvcu_dim3(1, 1, 1);                                        \
        threadIdx = dsvcu_dim3(0, 0, 0);                                       \
        for (unsigned bz_ = 0; bz_ < g_.z; bz_++)                              \
            for (unsigned by_ = 0; by_ < g_.y; by_++)                          \
                for (unsigned bx_ = 0; bx_ < g_.x; bx_++) {                    \
                    blockIdx = dsvcu_dim3(bx_, by_, bz_);                      \
                    kern(__VA_ARGS__);                                         \
                }                                                              \
    } while (0)

using std::max;
using std::min;
struct uint4 { unsigned x, y, z, w; };

static inline int dsvcu_malloc_(void **pp, size_t n) { *pp = calloc(1, n ? n : 1); return *pp ? 0 : 2; }
#define dsvcu_malloc(pp, n) dsvcu_malloc_((void **) (pp), (n))
#define dsvcu_free_dev(p) (free(p), 0)
#define dsvcu_malloc_host(pp, n) dsvcu_malloc_((void **) (pp), (n))
#define dsvcu_free_host(p) (free(p), 0)
#define dsvcu_memset_async(p, v, n, s) (memset((p), (v), (n)), 0)
#define dsvcu_h2d_async(d, h, n, s) (memcpy((d), (h), (n)), 0)
#define dsvcu_d2h_async(h, d, n, s) (memcpy((h), (d), (n)), 0)
#define dsvcu_d2d_async(d, s0, n, s) (memmove((d), (s0), (n)), 0)
static inline int
dsvcu_copy2d_(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h)
{
    size_t y;
    for (y = 0; y < h; y++) {
        memmove((char *) d + y * dp, (const char *) s + y * sp, w);
    }
    return 0;
}
#define dsvcu_h2d_2d_async(d, dp, h, hp, w, ht, s) dsvcu_copy2d_((d), (dp), (h), (hp), (w), (ht))
#define dsvcu_d2h_2d_async(h, hp, d, dp, w, ht, s) dsvcu_copy2d_((h), (hp), (d), (dp), (w), (ht))
#define dsvcu_d2d_2d_async(d, dp, s0, sp, w, ht, s) dsvcu_copy2d_((d), (dp), (s0), (sp), (w), (ht))
#define dsvcu_stream_sync(s) (0)
static inline int
dsvcu_memset2d_(void *d, size_t dp, int v, size_t w, size_t h)
{
    size_t y;
    for (y = 0; y < h; y++) {
        memset((char *) d + y * dp, v, w);
    }
    return 0;
}
#define dsvcu_memset_2d_async(p, pitch, v, w, h, s) dsvcu_memset2d_((p), (pitch), (v), (w), (h))

static inline int atomicAdd(int *p, int v) { int o = *p; *p += v; return o; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline int __ldg(const int *p) { return *p; }
static inline unsigned char __ldg(const unsigned char *p) { return *p; }
#endif /* DSVCU_EMU */

/* strided loop over the threads of a block */
#define DSVCU_TID ((int) threadIdx.x)
#define DSVCU_NTH ((int) blockDim.x)
#define PAR_FOR(i, n) for (int i = DSVCU_TID; i < (int) (n); i += DSVCU_NTH)

#endif /* DSVCU_RT_H */
