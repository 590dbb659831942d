/*
 * dsvcu_api.cu -- implementation of the C ABI declared in include/dsv_cuda.h.
 *
 * Owns device memory (frames, coefficient planes, scratch), builds the per-level
 * launch plans and launches the kernels in k_*.cuh.  Host-side scalar math that
 * the reference evaluates once per band/frame (lfquant, hfquant,
 * dsv_spatial_psy_factor, compute_filter_q, level/filter selection) lives here.
 */
#include "dsvcu_rt.h"
#include "k_sbt.cuh"
#include "k_quant.cuh"
#include "k_hzcc.cuh"
#include "k_bmc.cuh"
#include "k_filter.cuh"
#include "k_frame.cuh"
#include "k_hme.cuh"
#include "../../include/dsv_cuda.h"

#ifdef DSVCU_EMU
dsvcu_dim3 threadIdx, blockIdx, blockDim, gridDim;
unsigned char *dsvcu_emu_smem = 0;
size_t dsvcu_emu_smem_size = 0;
#endif

#define RSHIFT_UP(x, s) (((x) + (1 << (s)) - 1) >> (s))
#define FMT_HSHIFT(f) (((f) >> 2) & 3)
#define FMT_VSHIFT(f) ((f) & 3)
#define PSY_I_VISUAL_MASKING 4
#define PSY_P_VISUAL_MASKING 8

#ifndef DSVCU_EMU
/* Every encoder / decoder instance owns a stream and they are meant to overlap.
 * Streams are multiplexed onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues
 * (default 8): with more instances than queues, one instance's short kernels
 * line up behind another's long wavefront kernel.  Ask for the maximum unless
 * the user chose a value; only effective if no CUDA context exists yet. */
__attribute__((constructor)) static void
dsvcu_more_connections(void)
{
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
}
#endif

static char g_err[256] = "";
static int uniform_carveout(int device);
#ifdef DSVCU_DIAG
/* diagnostics build only (-DDSVCU_DIAG): experiment knobs, never in the product library */
static int g_pre_cap = getenv("DSVCU_PRE_GRID") ? atoi(getenv("DSVCU_PRE_GRID")) : 0; /* cap the prepass grid */
/* DSVCU_SKIP (bit mask): leave out whole kernel families to see what caps the many-instance
 * throughput (results are garbage, timing only): 1 transforms, 2 quantiser, 4 predict, 8 reconstruct,
 * 16 loop / intra filter, 32 border + pyramid, 64 search prepass, 128 search wavefront */
static int g_skip = getenv("DSVCU_SKIP") ? atoi(getenv("DSVCU_SKIP")) : 0;
static int g_sbt_cap = getenv("DSVCU_SBT_CTAS") ? atoi(getenv("DSVCU_SBT_CTAS")) : 0; /* CTAs per plane and transform launch */
static int g_grid_cap = getenv("DSVCU_GRID_CAP") ? atoi(getenv("DSVCU_GRID_CAP")) : 0; /* grid_for() ceiling */
#define DIAG_SKIP(bit) (g_skip & (bit))
#else
static const int g_pre_cap = 0, g_sbt_cap = 0, g_grid_cap = 0;
#define DIAG_SKIP(bit) 0
#endif
static long long g_launches = 0; /* kernels launched by every context of this process */

static int
fail(const char *what, int code)
{
#ifndef DSVCU_EMU
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString((cudaError_t) code));
#else
    snprintf(g_err, sizeof(g_err), "%s: error %d", what, code);
#endif
    return -1;
}

static int
fail_msg(const char *what)
{
    snprintf(g_err, sizeof(g_err), "%s", what);
    return -1;
}

#define CK(call)                                  \
    do {                                          \
        int e_ = (int) (call);                    \
        if (e_ != 0) return fail(#call, e_);      \
    } while (0)

#ifndef DSVCU_EMU
#define CK_LAUNCH(ctx)                                             \
    do {                                                           \
        (ctx)->launches++;                                         \
        __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);      \
        cudaError_t e_ = cudaGetLastError();                       \
        if (e_ != cudaSuccess) return fail("kernel launch", e_);   \
    } while (0)
#else
#define CK_LAUNCH(ctx) ((ctx)->launches++)
#endif

struct dsvcu_plane_t {
    uint8_t *base; /* first byte of the bordered plane */
    uint8_t *data; /* pixel (0,0) */
    int w, h, stride;
};

struct dsvcu_frame {
    uint8_t *alloc;
    size_t bytes;
    int nplanes;
    dsvcu_plane_t p[3];
};

struct dsvcu_coefs {
    int32_t *alloc;
    int32_t *data[3];
    int w[3], h[3];
};

struct dsvcu_ctx {
    int device;
    dsvcu_stream_t stream;
    int width, height, subsamp;
    int cw[3], ch[3]; /* coefficient plane dims */
    long long launches;
    /* side information */
    /* two blocks [vector field | blockdata]: the current picture's side information and
     * the previous picture's field (dsvcu_mvs_swap_prev exchanges them) */
    uint8_t *d_side[2];
    uint8_t *h_side;      /* pinned staging for one block (= h_side_set[stage_cur]) */
    uint8_t *h_side_set[2];
    size_t side_mv_bytes, side_bytes;
    int side_cur;
    uint8_t *d_blockdata; /* = d_side[side_cur] + side_mv_bytes */
    dsvcu_mv *d_mvs;      /* = d_side[side_cur] */
    int nblk_cap;
    /* transform scratch: two LL ping-pong planes per coefficient plane (the
     * three planes of a picture are transformed by the same launches) */
    int32_t *scratch[3][2];
    /* quantiser outputs */
    int32_t *d_qv[3];     /* dense scan-order quantiser output, per plane */
    int *d_chunk[3];
    int *d_meta;          /* [0..2] nsyms, [3..5] dc */
    int *h_meta;          /* pinned mirror */
    dsvcu_sym *d_syms[3];
    dsvcu_sym *h_syms[3]; /* pinned (= h_syms_set[stage_cur]) */
    /* Host staging comes in two sets so that a caller may fill the next picture's symbols and
     * side information while the copies of the previous picture are still in flight
     * (dsvcu_staging_flip); callers that wait for every picture only ever use set 0 */
    dsvcu_sym *h_syms_set[2][3];
    int stage_cur;
#ifndef DSVCU_EMU
    cudaEvent_t ev_stage[2]; /* last host-to-device copy out of the set has completed */
#endif
    int sym_cap[3];
    /* device-side entropy decode of batches of pictures (dsvcu_parse_begin / _end): two sets,
     * so that the planes of the next batch are parsed (on `pstream`) while the pictures of
     * the current one are reconstructed */
    struct ParseSet {
        uint8_t *h_bits, *d_bits; /* gathered plane bytes: pinned staging and device copy */
        size_t bits_cap;
        dsvcu_sym *d_syms;        /* one slot per (run, value) pair the plane headers announce */
        size_t syms_cap;
        HzSpan *h_spans, *d_spans;
        int *h_meta, *d_meta;     /* HZ_META_WORDS per plane */
        int spans_cap, n;
        int n_early;              /* planes [0, n_early) are part 0 of the batch, the rest part 1 */
        /* side information of the batch's inter pictures (k_hzcc.cuh, hz_parse_side) */
        HzSide *h_sides, *d_sides;
        int *h_side_ok, *d_side_ok;
        uint8_t *d_side_out;      /* one block [vector field | block flags] per picture */
        size_t side_out_cap, side_stride;
        int sides_cap, nsd, nsd_early;
        int pending[2];           /* part begun, result not collected yet */
#ifndef DSVCU_EMU
        cudaEvent_t ev_parsed[2]; /* meta words of the part are in pinned memory */
        cudaEvent_t ev_bits;      /* plane bytes and descriptors are on the device */
        cudaEvent_t ev_consumed;  /* last de-quantiser launch that read the set's symbols */
#endif
    } pset[2];
    int pset_last;
#ifndef DSVCU_EMU
    cudaStream_t pstream, pstream2; /* part 0 / part 1 of a batch */
#endif
    int *d_progress;
    int progress_cap;
    int me_smem_set;
    SbtJob sbt_job; /* launch descriptor under construction (3.7 KB: kept off the stack) */
    int filt_big_smem; /* opted in to > 48 KB dynamic shared memory on this device */
#ifndef DSVCU_EMU
    cudaEvent_t marks[DSVCU_MARKS];
#endif
    /* motion estimation */
    dsvcu_mv *d_mvf[ME_MAXLVL + 1]; /* [0] aliases d_mvs; [1..] live in d_me_zero */
    uint8_t *d_me_zero;             /* everything the search wants zeroed per picture: one memset */
    size_t me_zero_bytes;
    int *d_me_prog[ME_MAXLVL + 1];  /* per-level wavefront progress words */
    int me_prog_cap;
    dsvcu_mv *d_prev_mvf;
    int mvf_cap;
    MePre *d_pre; /* per-block prepass records of the level being searched */
    int *d_me;  /* [0..1] global motion, [2..5] accumulators, [6] luma avg */
    int *h_me;  /* pinned mirror */
    int me_nblk;
    dsvcu_mv *h_mvs; /* pinned mirror of a block array */
    int h_mvs_cap;
    int *d_lavg, *h_lavg; /* top-of-pyramid luma average */
#ifndef DSVCU_EMU
    cudaEvent_t ev0, ev1;
    cudaEvent_t ev_sym[3]; /* symbols of plane i are in pinned memory */
    cudaEvent_t ev_wait;
#endif
};

/* Host wait for everything queued on the context's stream.  Encoder / decoder
 * instances run on many host threads per GPU: the wait blocks in the OS
 * (cudaEventBlockingSync) instead of spinning, so waiting threads do not take
 * cores away from the ones that are packing bits or queueing kernels. */
static int
ctx_wait(dsvcu_ctx *c)
{
#ifndef DSVCU_EMU
    cudaError_t e = cudaEventRecord(c->ev_wait, c->stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(c->ev_wait);
    return (int) e;
#else
    (void) c;
    return 0;
#endif
}

static int
ilb2(unsigned n) /* dsv_lb2, dsv.c:449-459: ceil(log2(n)) */
{
    unsigned i = 1;
    int l = 0;
    while (i < n) {
        i <<= 1;
        l++;
    }
    return l;
}

static int
grid_for(int total, int threads)
{
    int g = (total + threads - 1) / threads;
    if (g < 1) g = 1;
    /* four CTAs per SM and a grid-stride loop inside: per-element kernels with grids of thousands of
     * short CTAs spend their time starting CTAs (measured: +2.5 % encoder throughput against 16 per SM) */
    if (g > 148 * 4) g = 148 * 4;
    if (g_grid_cap > 0 && g > g_grid_cap) g = g_grid_cap;
    return g;
}

/* ------------------------------------------------------------------ context */

extern "C" const char *
dsvcu_last_error(void)
{
    return g_err;
}

extern "C" int
dsvcu_device_count(void)
{
#ifndef DSVCU_EMU
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
#else
    return 1;
#endif
}

static void
coef_dims(int subsamp, int w, int h, int cw[3], int ch[3])
{
    /* dsv_mk_coefs, frame.c:29-60: chroma dims rounded up to even */
    int cwid = RSHIFT_UP(w, FMT_HSHIFT(subsamp));
    int chei = RSHIFT_UP(h, FMT_VSHIFT(subsamp));
    cwid = (cwid + 1) & ~1;
    chei = (chei + 1) & ~1;
    cw[0] = w;
    ch[0] = h;
    cw[1] = cw[2] = cwid;
    ch[1] = ch[2] = chei;
}

static int ctx_init(dsvcu_ctx *c, int device, int width, int height, int subsamp);

extern "C" int
dsvcu_ctx_create(dsvcu_ctx **out, int device, int width, int height, int subsamp)
{
    dsvcu_ctx *c;
    *out = NULL;
#ifndef DSVCU_EMU
    {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n <= 0) {
            snprintf(g_err, sizeof(g_err), "no CUDA device available (%s); this library has no CPU path",
                     e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
            return -1;
        }
        if (device < 0 || device >= n) {
            snprintf(g_err, sizeof(g_err), "device %d out of range (have %d)", device, n);
            return -1;
        }
        CK(cudaSetDevice(device));
    }
#endif
    c = (dsvcu_ctx *) calloc(1, sizeof(*c));
    if (!c) return -1;
    if (ctx_init(c, device, width, height, subsamp)) {
        dsvcu_ctx_destroy(c); /* releases whatever was created before the failure */
#ifndef DSVCU_EMU
        cudaGetLastError(); /* destroying handles that were never created leaves a sticky-looking error behind */
#endif
        return -1;
    }
    *out = c;
    return 0;
}

static int
ctx_init(dsvcu_ctx *c, int device, int width, int height, int subsamp)
{
    int i;
    size_t maxplane;
    c->device = device;
    c->width = width;
    c->height = height;
    c->subsamp = subsamp;
    coef_dims(subsamp, width, height, c->cw, c->ch);
#ifndef DSVCU_EMU
    if (uniform_carveout(device)) return -1;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    for (i = 0; i < 3; i++) {
        CK(cudaEventCreateWithFlags(&c->ev_sym[i], cudaEventDisableTiming | cudaEventBlockingSync));
    }
    CK(cudaEventCreateWithFlags(&c->ev_wait, cudaEventDisableTiming | cudaEventBlockingSync));
#endif
    maxplane = (size_t) c->cw[0] * c->ch[0];
    if ((size_t) c->cw[1] * c->ch[1] > maxplane) maxplane = (size_t) c->cw[1] * c->ch[1];
    for (i = 0; i < 3; i++) {
        CK(dsvcu_malloc(&c->scratch[i][0], (size_t) c->cw[i] * c->ch[i] * sizeof(int32_t)));
        CK(dsvcu_malloc(&c->scratch[i][1], (size_t) c->cw[i] * c->ch[i] * sizeof(int32_t)));
    }
    for (i = 0; i < 3; i++) {
        size_t pn = (size_t) c->cw[i] * c->ch[i];
        CK(dsvcu_malloc(&c->d_qv[i], (pn + CMP_CHUNK) * sizeof(int32_t)));
        CK(dsvcu_malloc(&c->d_chunk[i], (pn / CMP_CHUNK + 2) * sizeof(int)));
    }
    CK(dsvcu_malloc(&c->d_meta, 8 * sizeof(int)));
    CK(dsvcu_malloc_host(&c->h_meta, 8 * sizeof(int)));
    for (i = 0; i < 3; i++) {
        c->sym_cap[i] = c->cw[i] * c->ch[i] + 8;
        CK(dsvcu_malloc(&c->d_syms[i], (size_t) c->sym_cap[i] * sizeof(dsvcu_sym)));
        CK(dsvcu_malloc_host(&c->h_syms_set[0][i], (size_t) c->sym_cap[i] * sizeof(dsvcu_sym)));
        c->h_syms[i] = c->h_syms_set[0][i];
    }
    c->progress_cap = height / 4 + 64;
    CK(dsvcu_malloc(&c->d_progress, (size_t) c->progress_cap * sizeof(int)));
    CK(dsvcu_malloc(&c->d_me_zero, 256)); /* d_me until the search sizes its block (ensure_mvf) */
    c->d_me = (int *) c->d_me_zero;
    c->me_zero_bytes = 256;
    CK(dsvcu_malloc_host(&c->h_me, 16 * sizeof(int)));
    CK(dsvcu_malloc(&c->d_lavg, 4 * sizeof(int)));
    CK(dsvcu_malloc_host(&c->h_lavg, 4 * sizeof(int)));
    *c->h_lavg = 255;
    return 0;
}

static void
parse_set_free(dsvcu_ctx *c, int i)
{
    dsvcu_ctx::ParseSet *S = &c->pset[i];
    if (S->h_bits) dsvcu_free_host(S->h_bits);
    if (S->d_bits) dsvcu_free_dev(S->d_bits);
    if (S->d_syms) dsvcu_free_dev(S->d_syms);
    if (S->h_spans) dsvcu_free_host(S->h_spans);
    if (S->d_spans) dsvcu_free_dev(S->d_spans);
    if (S->h_meta) dsvcu_free_host(S->h_meta);
    if (S->d_meta) dsvcu_free_dev(S->d_meta);
    if (S->h_sides) dsvcu_free_host(S->h_sides);
    if (S->d_sides) dsvcu_free_dev(S->d_sides);
    if (S->h_side_ok) dsvcu_free_host(S->h_side_ok);
    if (S->d_side_ok) dsvcu_free_dev(S->d_side_ok);
    if (S->d_side_out) dsvcu_free_dev(S->d_side_out);
#ifndef DSVCU_EMU
    if (S->ev_parsed[0]) cudaEventDestroy(S->ev_parsed[0]);
    if (S->ev_parsed[1]) cudaEventDestroy(S->ev_parsed[1]);
    if (S->ev_bits) cudaEventDestroy(S->ev_bits);
    if (S->ev_consumed) cudaEventDestroy(S->ev_consumed);
#endif
    memset(S, 0, sizeof(*S));
}

extern "C" void
dsvcu_ctx_destroy(dsvcu_ctx *c)
{
    int i;
    if (!c) return;
#ifndef DSVCU_EMU
    cudaSetDevice(c->device);
    if (c->pstream) cudaStreamSynchronize(c->pstream);
    if (c->pstream2) cudaStreamSynchronize(c->pstream2);
    cudaStreamSynchronize(c->stream);
#endif
    for (i = 0; i < 3; i++) {
        dsvcu_free_dev(c->scratch[i][0]);
        dsvcu_free_dev(c->scratch[i][1]);
    }
    for (i = 0; i < 3; i++) {
        dsvcu_free_dev(c->d_qv[i]);
        dsvcu_free_dev(c->d_chunk[i]);
    }
    dsvcu_free_dev(c->d_meta);
    dsvcu_free_host(c->h_meta);
    for (i = 0; i < 3; i++) {
        dsvcu_free_dev(c->d_syms[i]);
        dsvcu_free_host(c->h_syms_set[0][i]);
        if (c->h_syms_set[1][i]) dsvcu_free_host(c->h_syms_set[1][i]);
    }
    for (i = 0; i < 2; i++) parse_set_free(c, i);
#ifndef DSVCU_EMU
    if (c->pstream) cudaStreamDestroy(c->pstream);
    if (c->pstream2) cudaStreamDestroy(c->pstream2);
#endif
    dsvcu_free_dev(c->d_progress);
    if (c->d_side[0]) dsvcu_free_dev(c->d_side[0]);
    if (c->d_side[1]) dsvcu_free_dev(c->d_side[1]);
    if (c->h_side) dsvcu_free_host(c->h_side);
    if (c->d_pre) dsvcu_free_dev(c->d_pre);
    if (c->d_me_zero) dsvcu_free_dev(c->d_me_zero);
    dsvcu_free_host(c->h_me);
    dsvcu_free_dev(c->d_lavg);
    dsvcu_free_host(c->h_lavg);
    if (c->h_mvs) dsvcu_free_host(c->h_mvs);
#ifndef DSVCU_EMU
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    for (i = 0; i < DSVCU_MARKS; i++) {
        if (c->marks[i]) cudaEventDestroy(c->marks[i]);
    }
    for (i = 0; i < 3; i++) cudaEventDestroy(c->ev_sym[i]);
    cudaEventDestroy(c->ev_wait);
    cudaStreamDestroy(c->stream);
#endif
    free(c);
}

extern "C" void *
dsvcu_host_alloc(size_t bytes)
{
    void *p = NULL;
    if (dsvcu_malloc_host(&p, bytes ? bytes : 1)) {
#ifndef DSVCU_EMU
        cudaGetLastError();
#endif
        return NULL;
    }
    return p;
}

extern "C" void
dsvcu_host_free(void *p)
{
    if (p) dsvcu_free_host(p);
}

extern "C" void *
dsvcu_ctx_stream(dsvcu_ctx *c)
{
#ifndef DSVCU_EMU
    return (void *) c->stream;
#else
    (void) c;
    return NULL;
#endif
}

extern "C" int
dsvcu_sync(dsvcu_ctx *c)
{
    (void) c;
    CK(ctx_wait(c));
    return 0;
}

extern "C" long long
dsvcu_launch_count(dsvcu_ctx *c)
{
    return c->launches;
}

/* host evaluation of the integer square root the motion search uses (same
 * source as the device function); lets the test-suite check it against the
 * reference's digit-by-digit form without a GPU */
extern "C" unsigned
dsvcu_isqrt(unsigned n)
{
    return me_isqrt(n);
}

extern "C" long long
dsvcu_total_launches(void)
{
    return __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
}

extern "C" int
dsvcu_timer_start(dsvcu_ctx *c)
{
#ifndef DSVCU_EMU
    CK(cudaEventRecord(c->ev0, c->stream));
#else
    (void) c;
#endif
    return 0;
}

extern "C" int
dsvcu_timer_stop_ms(dsvcu_ctx *c, float *ms)
{
#ifndef DSVCU_EMU
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
#else
    (void) c;
    *ms = 0.f;
#endif
    return 0;
}

#if defined(DSVCU_DIAG) && defined(ME_TIMING) && !defined(DSVCU_EMU)
/* Diagnostics build only: what does ONE extra warp per SM see while the encoder
 * instances run?  mode 0: a dependent integer chain in a hot loop (issue
 * arbitration only); mode 1: a dependent chain of L2 loads (ld.cg pointer chase
 * over 4 MB); mode 2: a dependent chain of shared-memory loads.  Returns the
 * average cycles per step over all probe warps. */
__global__ static void
k_probe(int mode, int iters, const unsigned *chase, unsigned long long *out)
{
    __shared__ unsigned sm[1024];
    unsigned v = threadIdx.x + blockIdx.x;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 197 + 31) & 1023;
    __syncthreads();
    long long t0 = clock64();
    if (mode == 0) {
        for (int i = 0; i < iters; i++) v = v * 1664525u + 1013904223u;
    } else if (mode == 1) {
        v = (blockIdx.x * 7919u) & ((1u << 20) - 1);
        for (int i = 0; i < iters; i++) v = __ldcg(chase + v);
    } else {
        v &= 1023;
        for (int i = 0; i < iters; i++) v = sm[v];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) {
        atomicAdd(&out[0], (unsigned long long) (t1 - t0));
        atomicAdd(&out[1], 1ull);
        if (v == 0xffffffffu) out[2] = v;
    }
}

extern "C" int
dsvcu_debug_probe(int mode, int iters, double *cycles_per_step)
{
    static unsigned *chase = NULL;
    static unsigned long long *out = NULL;
    static cudaStream_t st;
    unsigned long long h[3];
    if (!chase) {
        unsigned *hc = (unsigned *) malloc(4u << 20);
        unsigned n = 1u << 20, i;
        for (i = 0; i < n; i++) hc[i] = (unsigned) (((unsigned long long) i * 1000003ull + 12345ull) & (n - 1));
        cudaMalloc((void **) &chase, 4u << 20);
        cudaMemcpy(chase, hc, 4u << 20, cudaMemcpyHostToDevice);
        free(hc);
        cudaMalloc((void **) &out, 32);
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    }
    cudaMemsetAsync(out, 0, 32, st);
    k_probe<<<148, 32, 0, st>>>(mode, iters, chase, out);
    cudaMemcpyAsync(h, out, 24, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    *cycles_per_step = h[1] ? (double) h[0] / (double) h[1] / iters : 0.0;
    return 0;
}

extern "C" int
dsvcu_debug_me_counters(unsigned long long out[4], int reset)
{
    static const unsigned long long z[4] = { 0, 0, 0, 0 };
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_me_dbg, sizeof(z));
    if (reset) cudaMemcpyToSymbol(g_me_dbg, z, sizeof(z));
    return 0;
}
#endif

/* Device-side phase stamps for the encoder's profiler (DSV_PROFILE=2): record
 * mark k on the stream; read the time between two marks once both completed. */
extern "C" int
dsvcu_mark(dsvcu_ctx *c, int k)
{
#ifndef DSVCU_EMU
    if (k < 0 || k >= DSVCU_MARKS) return -1;
    if (!c->marks[k]) CK(cudaEventCreate(&c->marks[k]));
    CK(cudaEventRecord(c->marks[k], c->stream));
#else
    (void) c;
    (void) k;
#endif
    return 0;
}

extern "C" int
dsvcu_mark_elapsed_ms(dsvcu_ctx *c, int a, int b, float *ms)
{
    *ms = 0.f;
#ifndef DSVCU_EMU
    if (a < 0 || b < 0 || a >= DSVCU_MARKS || b >= DSVCU_MARKS || !c->marks[a] || !c->marks[b]) return -1;
    CK(cudaEventSynchronize(c->marks[b]));
    CK(cudaEventElapsedTime(ms, c->marks[a], c->marks[b]));
#else
    (void) c;
    (void) a;
    (void) b;
#endif
    return 0;
}

/* ------------------------------------------------------------------- frames */

static int
frame_alloc(dsvcu_frame **out, int nplanes, const int w[3], const int h[3])
{
    /* geometry of dsv_mk_frame with border, frame.c:62-113: stride =
     * round16(w + 64), 32 rows above and below; 8 spare rows keep the clamped
     * +3 sub-pel window inside the allocation */
    dsvcu_frame *f = (dsvcu_frame *) calloc(1, sizeof(*f));
    size_t off[3], total = 0;
    int i;
    if (!f) return -1;
    f->nplanes = nplanes;
    for (i = 0; i < nplanes; i++) {
        f->p[i].w = w[i];
        f->p[i].h = h[i];
        f->p[i].stride = (w[i] + 2 * FR_BORDER + 15) & ~15;
        off[i] = total;
        total += (size_t) f->p[i].stride * (h[i] + 2 * FR_BORDER + 8);
        total = (total + 255) & ~(size_t) 255;
    }
    f->bytes = total;
    CK(dsvcu_malloc(&f->alloc, total));
    for (i = 0; i < nplanes; i++) {
        f->p[i].base = f->alloc + off[i];
        f->p[i].data = f->p[i].base + (size_t) f->p[i].stride * FR_BORDER + FR_BORDER;
    }
    *out = f;
    return 0;
}

extern "C" int
dsvcu_frame_create(dsvcu_ctx *c, dsvcu_frame **out)
{
    int w[3], h[3], r;
    w[0] = c->width;
    h[0] = c->height;
    w[1] = w[2] = RSHIFT_UP(c->width, FMT_HSHIFT(c->subsamp));
    h[1] = h[2] = RSHIFT_UP(c->height, FMT_VSHIFT(c->subsamp));
    r = frame_alloc(out, 3, w, h);
    if (r == 0) {
        CK(dsvcu_memset_async((*out)->alloc, 0, (*out)->bytes, c->stream));
    }
    return r;
}

extern "C" int
dsvcu_frame_create_luma(dsvcu_ctx *c, dsvcu_frame **out, int w, int h)
{
    int ww[3] = { w, 0, 0 }, hh[3] = { h, 0, 0 }, r;
    r = frame_alloc(out, 1, ww, hh);
    if (r == 0) {
        CK(dsvcu_memset_async((*out)->alloc, 0, (*out)->bytes, c->stream));
    }
    return r;
}

extern "C" void
dsvcu_frame_destroy(dsvcu_ctx *c, dsvcu_frame *f)
{
    (void) c;
    if (!f) return;
    dsvcu_free_dev(f->alloc);
    free(f);
}

extern "C" int
dsvcu_frame_plane_dims(dsvcu_frame *f, int plane, int *w, int *h, int *stride)
{
    if (plane < 0 || plane >= f->nplanes) return -1;
    if (w) *w = f->p[plane].w;
    if (h) *h = f->p[plane].h;
    if (stride) *stride = f->p[plane].stride;
    return 0;
}

extern "C" int
dsvcu_frame_upload(dsvcu_ctx *c, dsvcu_frame *f, int plane, const uint8_t *src, int hstride)
{
    dsvcu_plane_t *p = &f->p[plane];
    CK(dsvcu_h2d_2d_async(p->data, p->stride, src, hstride, p->w, p->h, c->stream));
    return 0;
}

extern "C" int
dsvcu_frame_download(dsvcu_ctx *c, dsvcu_frame *f, int plane, uint8_t *dst, int hstride)
{
    dsvcu_plane_t *p = &f->p[plane];
    CK(dsvcu_d2h_2d_async(dst, hstride, p->data, p->stride, p->w, p->h, c->stream));
    return 0;
}

extern "C" int
dsvcu_frame_clear_plane(dsvcu_ctx *c, dsvcu_frame *f, int plane, int value)
{
    dsvcu_plane_t *p = &f->p[plane];
    CK(dsvcu_memset_2d_async(p->data, p->stride, value, p->w, p->h, c->stream));
    return 0;
}

extern "C" int
dsvcu_frame_upload_bordered(dsvcu_ctx *c, dsvcu_frame *f, int plane, const uint8_t *src)
{
    dsvcu_plane_t *p = &f->p[plane];
    CK(dsvcu_h2d_async(p->base, src, (size_t) p->stride * (p->h + 2 * FR_BORDER), c->stream));
    return 0;
}

extern "C" int
dsvcu_frame_download_bordered(dsvcu_ctx *c, dsvcu_frame *f, int plane, uint8_t *dst)
{
    dsvcu_plane_t *p = &f->p[plane];
    CK(dsvcu_d2h_async(dst, p->base, (size_t) p->stride * (p->h + 2 * FR_BORDER), c->stream));
    return 0;
}

/* ------------------------------------------------------------- coefficients */

extern "C" int
dsvcu_coefs_create(dsvcu_ctx *c, dsvcu_coefs **out)
{
    dsvcu_coefs *k = (dsvcu_coefs *) calloc(1, sizeof(*k));
    size_t n = 0;
    int i;
    if (!k) return -1;
    for (i = 0; i < 3; i++) {
        k->w[i] = c->cw[i];
        k->h[i] = c->ch[i];
        n += (size_t) k->w[i] * k->h[i];
    }
    CK(dsvcu_malloc(&k->alloc, n * sizeof(int32_t)));
    CK(dsvcu_memset_async(k->alloc, 0, n * sizeof(int32_t), c->stream));
    k->data[0] = k->alloc;
    k->data[1] = k->data[0] + (size_t) k->w[0] * k->h[0];
    k->data[2] = k->data[1] + (size_t) k->w[1] * k->h[1];
    *out = k;
    return 0;
}

extern "C" void
dsvcu_coefs_destroy(dsvcu_ctx *c, dsvcu_coefs *k)
{
    (void) c;
    if (!k) return;
    dsvcu_free_dev(k->alloc);
    free(k);
}

extern "C" int
dsvcu_coefs_plane_dims(dsvcu_coefs *k, int plane, int *w, int *h)
{
    if (w) *w = k->w[plane];
    if (h) *h = k->h[plane];
    return 0;
}

extern "C" int
dsvcu_coefs_upload(dsvcu_ctx *c, dsvcu_coefs *k, int plane, const int32_t *src)
{
    CK(dsvcu_h2d_async(k->data[plane], src, (size_t) k->w[plane] * k->h[plane] * sizeof(int32_t), c->stream));
    return 0;
}

extern "C" int
dsvcu_coefs_download(dsvcu_ctx *c, dsvcu_coefs *k, int plane, int32_t *dst)
{
    CK(dsvcu_d2h_async(dst, k->data[plane], (size_t) k->w[plane] * k->h[plane] * sizeof(int32_t), c->stream));
    return 0;
}

/* --------------------------------------------------------- side information */

static void
side_point(dsvcu_ctx *c)
{
    c->d_mvs = (dsvcu_mv *) c->d_side[c->side_cur];
    c->d_blockdata = c->d_side[c->side_cur] + c->side_mv_bytes;
    c->d_prev_mvf = (dsvcu_mv *) c->d_side[c->side_cur ^ 1];
}

static int
ensure_blocks(dsvcu_ctx *c, int n)
{
    int i;
    if (n <= c->nblk_cap) return 0;
    for (i = 0; i < 2; i++) {
        if (c->d_side[i]) dsvcu_free_dev(c->d_side[i]);
        c->d_side[i] = NULL;
    }
    for (i = 0; i < 2; i++) {
        if (c->h_side_set[i]) dsvcu_free_host(c->h_side_set[i]);
        c->h_side_set[i] = NULL;
    }
    c->h_side = NULL;
    c->side_mv_bytes = ((((size_t) n + 4) * sizeof(dsvcu_mv)) + 255) & ~(size_t) 255;
    c->side_bytes = c->side_mv_bytes + (((size_t) n + 64 + 255) & ~(size_t) 255);
    for (i = 0; i < 2; i++) {
        CK(dsvcu_malloc(&c->d_side[i], c->side_bytes));
        CK(dsvcu_memset_async(c->d_side[i], 0, c->side_bytes, c->stream));
    }
    CK(dsvcu_malloc_host(&c->h_side_set[0], c->side_bytes));
    CK(dsvcu_malloc_host(&c->h_side_set[1], c->side_bytes));
    c->h_side = c->h_side_set[c->stage_cur];
    c->side_cur = 0;
    side_point(c);
    c->nblk_cap = n;
    return 0;
}

extern "C" int
dsvcu_set_blockdata(dsvcu_ctx *c, const uint8_t *bd, int n)
{
    if (ensure_blocks(c, n)) return -1;
    CK(dsvcu_h2d_async(c->d_blockdata, bd, (size_t) n, c->stream));
    return 0;
}

extern "C" int
dsvcu_set_mvs(dsvcu_ctx *c, const void *mvs, int n)
{
    if (ensure_blocks(c, n)) return -1;
    CK(dsvcu_h2d_async(c->d_mvs, mvs, (size_t) n * sizeof(dsvcu_mv), c->stream));
    return 0;
}

/* dsvcu_set_blockdata + dsvcu_set_mvs (mvs may be NULL) as ONE copy from pinned
 * staging.  The staging block is reused: the caller must not call this again
 * before the previous copy has left the host, i.e. before its next wait on the
 * context (every picture waits for its symbols / its output). */
extern "C" int
dsvcu_set_side(dsvcu_ctx *c, const uint8_t *bd, const void *mvs, int n)
{
    if (ensure_blocks(c, n)) return -1;
    memcpy(c->h_side + c->side_mv_bytes, bd, (size_t) n);
    if (mvs) {
        memcpy(c->h_side, mvs, (size_t) n * sizeof(dsvcu_mv));
        CK(dsvcu_h2d_async(c->d_side[c->side_cur], c->h_side, c->side_mv_bytes + (size_t) n, c->stream));
    } else {
        CK(dsvcu_h2d_async(c->d_blockdata, c->h_side + c->side_mv_bytes, (size_t) n, c->stream));
    }
#ifndef DSVCU_EMU
    if (c->ev_stage[c->stage_cur]) CK(cudaEventRecord(c->ev_stage[c->stage_cur], c->stream));
#endif
    return 0;
}

/* The current vector field becomes "the previous picture's field" (DSV_HME.ref_mvf)
 * by exchanging the two side-information blocks: no copy.  The current field and
 * blockdata are undefined afterwards -- call it after the picture's last operator
 * that reads them has been queued. */
extern "C" int
dsvcu_mvs_swap_prev(dsvcu_ctx *c, int nblocks)
{
    if (ensure_blocks(c, nblocks)) return -1;
    c->side_cur ^= 1;
    side_point(c);
    return 0;
}

/* ---------------------------------------------------------------- transforms */

static int
sbt_nlevels(int w, int h) /* sbt.c:834-845 */
{
    return ilb2((unsigned) (w > h ? w : h));
}

/* which 1-D filter family a level uses (sbt.c:19-29, :862-930) */
static int
sbt_pick(int l, int lvls, int luma, int isP, int lossless)
{
    if (lossless) {
        return (l >= 1 && l <= lvls - 2) ? SBT_F_LOSSLESS : SBT_F_HAAR_SIMPLE;
    }
    if (luma && !isP && l == 4) return SBT_F_LLI;
    if (luma && isP && l == 4) return SBT_F_LLP;
    if (!luma && !isP && l >= 1 && l <= lvls - 2) return SBT_F_CC;
    if (luma && !isP && l == 2) return SBT_F_L2A;
    if (luma && !isP && l == 1) return SBT_F_L1;
    return (luma || !isP) ? SBT_F_HAAR : SBT_F_HAAR_SIMPLE;
}

static void
sbt_level_geom(SbtLevel *L, int w, int h, int l, const dsvcu_fmeta *fm, const uint8_t *bd)
{
    memset(L, 0, sizeof(*L));
    L->fw = w;
    L->sw = RSHIFT_UP(w, l - 1);
    L->sh = RSHIFT_UP(h, l - 1);
    L->cw = RSHIFT_UP(w, l);
    L->ch = RSHIFT_UP(h, l);
    L->blockdata = bd;
    L->nbh = fm->nblocks_h;
    L->dbx = (fm->nblocks_h << BLOCK_INTERP_P) / L->sw;
    L->dby = (fm->nblocks_v << BLOCK_INTERP_P) / L->sh;
}

/* level l of a plane, inverse direction */
static void
sbt_inv_level(dsvcu_ctx *c, SbtLevel *L, dsvcu_frame *dst, int plane, dsvcu_coefs *src, int q, const dsvcu_fmeta *fm, int l,
              int lvls)
{
    const int luma = (plane == 0);
    sbt_level_geom(L, src->w[plane], src->h[plane], l, fm, c->d_blockdata);
    L->hqp = luma ? (q / (fm->isP ? 14 : (l > 4 ? 2 : 8))) : (q / 2);
    L->ovf = (l >= 6 && l >= lvls - 3 && !fm->lossless);
    L->ll = (l == lvls) ? src->data[plane] : c->scratch[plane][(l + 1) & 1];
    L->bands = src->data[plane];
    if (l == 1) {
        L->px = dst->p[plane].data;
        L->px_stride = dst->p[plane].stride;
        L->px_w = dst->p[plane].w;
        L->px_h = dst->p[plane].h;
    } else {
        L->dst = c->scratch[plane][l & 1];
    }
}

static void
sbt_fwd_level(dsvcu_ctx *c, SbtLevel *L, dsvcu_frame *src, int plane, dsvcu_coefs *dst, const dsvcu_fmeta *fm, int l, int lvls)
{
    sbt_level_geom(L, dst->w[plane], dst->h[plane], l, fm, c->d_blockdata);
    L->ovf = (l >= 6 && l >= lvls - 3 && !fm->lossless);
    if (l == 1) {
        L->px = src->p[plane].data;
        L->px_stride = src->p[plane].stride;
        L->px_w = src->p[plane].w;
        L->px_h = src->p[plane].h;
    } else {
        L->src = c->scratch[plane][(l - 1) & 1];
    }
    L->out_ll = (l == lvls) ? dst->data[plane] : c->scratch[plane][l & 1];
    L->out_bands = dst->data[plane];
}

static int
sbt_is_lift(int f)
{
    return f != SBT_F_HAAR && f != SBT_F_HAAR_SIMPLE;
}

/* first level from which a plane's sub-images are at most two tiles: that
 * level and everything above it is transformed by one CTA in one launch */
static int
sbt_tail_start(int w, int h, int lvls)
{
    int l;
    for (l = 1; l <= lvls; l++) {
        int sw = RSHIFT_UP(w, l - 1), sh = RSHIFT_UP(h, l - 1);
        if (((sw + SBT_TW - 1) / SBT_TW) * ((sh + SBT_TH - 1) / SBT_TH) <= 2) break;
    }
    if (l > lvls) l = lvls;
    if (lvls - l + 1 > SBT_MAX_FUSED) l = lvls - SBT_MAX_FUSED + 1;
    return l;
}

static int
sbt_level_ctas(const SbtLevel *L, int f)
{
    if (sbt_is_lift(f)) return sbt_level_tiles(*L);
    return grid_for(L->cw * L->ch, SBT_THREADS);
}

/* Transform the planes selected by `mask` (bit p = plane p): one launch per
 * pyramid level covering every selected plane that still has a large level
 * there, and one launch for the fused small levels of all of them. */
static int
sbt_run(dsvcu_ctx *c, dsvcu_frame *fr, dsvcu_coefs *k, int q, const dsvcu_fmeta *fm, int mask, int fwd)
{
    int lvls[3], tail[3], p, l, maxbig = 0, stage;
    SbtJob *J = &c->sbt_job;
    for (p = 0; p < 3; p++) {
        if (!(mask & (1 << p))) continue;
        lvls[p] = sbt_nlevels(k->w[p], k->h[p]);
        tail[p] = sbt_tail_start(k->w[p], k->h[p], lvls[p]);
        if (tail[p] - 1 > maxbig) maxbig = tail[p] - 1;
    }
    /* stages: forward = levels 1..maxbig then the tails; inverse = the tails then maxbig..1 */
    if (!mask || DIAG_SKIP(1)) return 0;
    for (stage = 0; stage <= maxbig; stage++) {
        const int is_tail = fwd ? (stage == maxbig) : (stage == 0);
        int ctas = 0;
        l = fwd ? stage + 1 : maxbig - stage + 1;
        J->nplanes = 0;
        for (p = 0; p < 3; p++) {
            SbtPlaneJob *P;
            int n = 0, ll;
            if (!(mask & (1 << p))) continue;
            if (!is_tail && l >= tail[p]) continue;
            P = &J->p[J->nplanes];
            if (is_tail) {
                for (ll = fwd ? tail[p] : lvls[p]; fwd ? ll <= lvls[p] : ll >= tail[p]; ll += fwd ? 1 : -1) {
                    P->f[n] = sbt_pick(ll, lvls[p], p == 0, fm->isP, fm->lossless);
                    if (fwd) {
                        sbt_fwd_level(c, &P->L[n], fr, p, k, fm, ll, lvls[p]);
                    } else {
                        sbt_inv_level(c, &P->L[n], fr, p, k, q, fm, ll, lvls[p]);
                    }
                    n++;
                }
                P->ncta = 1;
            } else {
                P->f[0] = sbt_pick(l, lvls[p], p == 0, fm->isP, fm->lossless);
                if (fwd) {
                    sbt_fwd_level(c, &P->L[0], fr, p, k, fm, l, lvls[p]);
                } else {
                    sbt_inv_level(c, &P->L[0], fr, p, k, q, fm, l, lvls[p]);
                }
                n = 1;
                P->ncta = sbt_level_ctas(&P->L[0], P->f[0]);
                if (g_sbt_cap > 0 && P->ncta > g_sbt_cap) P->ncta = g_sbt_cap;
            }
            P->nlev = n;
            P->first_cta = ctas;
            ctas += P->ncta;
            J->nplanes++;
        }
        if (!J->nplanes) continue;
        if (fwd) {
            DSVCU_LAUNCH(k_sbt_fwd, ctas, SBT_THREADS, 0, c->stream, *J);
        } else {
            DSVCU_LAUNCH(k_sbt_inv, ctas, SBT_THREADS, 0, c->stream, *J);
        }
        CK_LAUNCH(c);
    }
    return 0;
}

extern "C" int
dsvcu_inv_sbt(dsvcu_ctx *c, dsvcu_frame *dst, int plane, dsvcu_coefs *src, int q, const dsvcu_fmeta *fm)
{
    return sbt_run(c, dst, src, q, fm, 1 << plane, 0);
}

extern "C" int
dsvcu_fwd_sbt(dsvcu_ctx *c, dsvcu_frame *src, int plane, dsvcu_coefs *dst, const dsvcu_fmeta *fm)
{
    return sbt_run(c, src, dst, 0, fm, 1 << plane, 1);
}

/* all three planes of a picture through the same launches */
extern "C" int
dsvcu_inv_sbt_frame(dsvcu_ctx *c, dsvcu_frame *dst, dsvcu_coefs *src, int q, const dsvcu_fmeta *fm, int plane_mask)
{
    return sbt_run(c, dst, src, q, fm, plane_mask & 7, 0);
}

extern "C" int
dsvcu_fwd_sbt_frame(dsvcu_ctx *c, dsvcu_frame *src, dsvcu_coefs *dst, const dsvcu_fmeta *fm, int plane_mask)
{
    return sbt_run(c, src, dst, 0, fm, plane_mask & 7, 1);
}

/* ------------------------------------------------------------- quantisation */

static int
psy_factor(const dsvcu_fmeta *fm, int subband) /* dsv_spatial_psy_factor, hzcc.c:66-86 */
{
    int scale, lo, hi;
    int bw = fm->blk_w, bh = fm->blk_h;
    if (subband == 1) {
        lo = (352 + bw - 1) / bw;
        hi = (1920 + bw - 1) / bw;
        scale = fm->nblocks_h;
    } else if (subband == 2) {
        lo = (288 + bh - 1) / bh;
        hi = (1080 + bh - 1) / bh;
        scale = fm->nblocks_v;
    } else {
        lo = ((352 + bw - 1) / bw) * ((288 + bh - 1) / bh);
        hi = ((1920 + bw - 1) / bw) * ((1080 + bh - 1) / bh);
        scale = fm->nblocks_h * fm->nblocks_v;
    }
    scale = scale - lo > 0 ? scale - lo : 0;
    return (scale << 7) / (hi - lo);
}

static int
lfquant(int q, int plane, const dsvcu_fmeta *fm) /* hzcc.c:88-105 */
{
    int psyfac = psy_factor(fm, 3);
    q -= (q * psyfac >> (7 + 3));
    if (q < 8) q = 8;
    if (plane) {
        if (q > 256) q = 256 + q / 4;
        return q < 768 ? q : 768;
    }
    return q < 3072 ? q : 3072;
}

static int
hfquant(const dsvcu_fmeta *fm, int plane, int subsamp, int q, int s, int l) /* hzcc.c:107-162 */
{
    int chroma = (plane != 0);
    int psyfac = psy_factor(fm, s);
    q /= 2;
    psyfac = q * psyfac >> (7 + (fm->isP ? 0 : 1));
    if (chroma) {
        int tl = l - 2;
        if (s == 1) {
            tl += FMT_HSHIFT(subsamp);
        } else if (s == 2) {
            tl += FMT_VSHIFT(subsamp);
        }
        q = (q * 6) / (4 - tl);
    } else {
        if (l == 1) {
            q += psyfac / 2;
        } else if (l == 2) {
            q += psyfac;
        }
    }
    if (fm->isP) {
        if (l != 2) {
            if (l == 0) {
                q *= 2;
                q -= psyfac;
            } else {
                q -= psyfac / 2;
            }
        }
        return (q / 4) > 8 ? (q / 4) : 8;
    }
    q = q * (15 + 3 * l) / 16;
    if (!chroma) {
        if (l == 0) {
            q = (q * 3) / 8;
        } else if (s == 3) {
            q *= 2;
        }
    } else {
        q /= 4;
        if (s == 3) q *= 2;
    }
    return q > 8 ? q : 8;
}

extern "C" int
dsvcu_scan_layout(int w, int h, int part_start[5])
{
    int l, pos;
    pos = RSHIFT_UP(w, 3) * RSHIFT_UP(h, 3);
    part_start[0] = 0;
    for (l = 0; l < 3; l++) {
        part_start[1 + l] = pos;
        pos += 3 * RSHIFT_UP(w, 3 - l) * RSHIFT_UP(h, 3 - l);
    }
    part_start[4] = pos;
    return pos;
}

static void
quant_level_geom(QuantLevel *Q, dsvcu_ctx *c, dsvcu_coefs *k, int plane, int q, const dsvcu_fmeta *fm, int l,
                 const int part_start[5])
{
    const int w = k->w[plane], h = k->h[plane];
    int s;
    memset(Q, 0, sizeof(*Q));
    Q->coefs = k->data[plane];
    Q->fw = w;
    Q->l = l;
    Q->w = RSHIFT_UP(w, 3 - (l < 0 ? 0 : l));
    Q->h = RSHIFT_UP(h, 3 - (l < 0 ? 0 : l));
    Q->isP = fm->isP;
    Q->luma = (plane == 0);
    Q->lossless = fm->lossless;
    Q->psy = (plane == 0) && (fm->do_psy & (fm->isP ? PSY_P_VISUAL_MASKING : PSY_I_VISUAL_MASKING));
    Q->nbh = fm->nblocks_h;
    Q->dbx = (fm->nblocks_h << 14) / Q->w;
    Q->dby = (fm->nblocks_v << 14) / Q->h;
    Q->blockdata = c->d_blockdata;
    Q->mvs = c->d_mvs;
    if (l < 0) return;
    for (s = 1; s <= 3; s++) {
        QuantBand *B = &Q->band[s - 1];
        B->ox = (s & 1) ? Q->w : 0;
        B->oy = (s & 2) ? Q->h : 0;
        B->pox = (s & 1) ? RSHIFT_UP(w, 4 - l) : 0;
        B->poy = (s & 2) ? RSHIFT_UP(h, 4 - l) : 0;
        B->gox = (s & 1) ? RSHIFT_UP(w, 5 - l) : 0;
        B->goy = (s & 2) ? RSHIFT_UP(h, 5 - l) : 0;
        B->qp = fm->lossless ? 1 : hfquant(fm, plane, c->subsamp, q, s, l);
        B->scan_base = part_start[1 + l] + (s - 1) * Q->w * Q->h;
    }
}

/* Quantise the planes in `mask` with shared launches (bit p = plane p): the LL
 * part, three levels (+ their aliased edges), then the ordered compaction of the
 * non-zero symbols into pinned host memory -- 10 launches whatever the number of
 * planes.  Level order inside a plane is a data dependency (a child reads its
 * quantised parent); planes are independent. */
static int
quant_run(dsvcu_ctx *c, dsvcu_coefs *k, int q, const dsvcu_fmeta *fm, int mask)
{
    QuantJob J;
    CompactJob CJ;
    int part[3][5], total[3], pl[3], n = 0, l, i, gmax, emax, cmax = 0;
    int qf = q * 3 / 2; /* fix_quant, hzcc.c:59-63 */

    for (i = 0; i < 3; i++) {
        if (mask & (1 << i)) pl[n++] = i;
    }
    if (!n) return 0;
    if (DIAG_SKIP(2)) {
        for (i = 0; i < n; i++) c->h_meta[pl[i]] = 0;
#ifndef DSVCU_EMU
        for (i = 0; i < n; i++) CK(cudaEventRecord(c->ev_sym[pl[i]], c->stream));
#endif
        return 0;
    }
    memset(&CJ, 0, sizeof(CJ));
    gmax = 1;
    for (i = 0; i < n; i++) {
        const int p = pl[i];
        total[i] = dsvcu_scan_layout(k->w[p], k->h[p], part[i]);
        quant_level_geom(&J.Q[i], c, k, p, qf, fm, -1, part[i]);
        J.Q[i].qv = c->d_qv[p];
        J.lfq[i] = fm->lossless ? 1 : lfquant(qf, p, fm);
        gmax = max(gmax, grid_for(J.Q[i].w * J.Q[i].h, 256));
    }
    DSVCU_LAUNCH(k_quant_ll, dim3(gmax, 1, n), 256, 0, c->stream, J);
    CK_LAUNCH(c);
    for (l = 0; l < 3; l++) {
        gmax = emax = 1;
        for (i = 0; i < n; i++) {
            quant_level_geom(&J.Q[i], c, k, pl[i], qf, fm, l, part[i]);
            J.Q[i].qv = c->d_qv[pl[i]];
            gmax = max(gmax, grid_for(J.Q[i].w * J.Q[i].h, 256));
            emax = max(emax, grid_for(J.Q[i].w + J.Q[i].h, 256));
        }
        DSVCU_LAUNCH(k_quant_hf, dim3(gmax, 3, n), 256, 0, c->stream, J);
        CK_LAUNCH(c);
        if (!fm->lossless) {
            DSVCU_LAUNCH(k_quant_hf_edge, dim3(emax, 3, n), 256, 0, c->stream, J);
            CK_LAUNCH(c);
        }
    }
    /* the symbol count, the DC coefficient and the ordered (position, value)
     * list are written by the kernels straight into pinned host memory
     * (unified addressing): the host entropy coder needs nothing else, so one
     * event is all it waits for while the GPU carries on with the inverse
     * transform */
    for (i = 0; i < n; i++) {
        const int p = pl[i];
        CJ.qv[i] = c->d_qv[p];
        CJ.n[i] = total[i];
        CJ.chunk[i] = c->d_chunk[p];
        CJ.nchunks[i] = (total[i] + CMP_CHUNK - 1) / CMP_CHUNK;
        CJ.out[i] = c->h_syms[p];
        CJ.out_n[i] = c->h_meta + p;
        CJ.dc_src[i] = k->data[p];
        CJ.dc_dst[i] = c->h_meta + 3 + p;
        cmax = max(cmax, CJ.nchunks[i]);
    }
    DSVCU_LAUNCH(k_compact_count, dim3(cmax, n, 1), CMP_THREADS, 0, c->stream, CJ);
    CK_LAUNCH(c);
    DSVCU_LAUNCH(k_compact_scan, n, 1024, 0, c->stream, CJ);
    CK_LAUNCH(c);
    DSVCU_LAUNCH(k_compact_scatter, dim3(cmax, n, 1), CMP_THREADS, 0, c->stream, CJ);
    CK_LAUNCH(c);
#ifndef DSVCU_EMU
    for (i = 0; i < n; i++) CK(cudaEventRecord(c->ev_sym[pl[i]], c->stream));
#endif
    return 0;
}

extern "C" int
dsvcu_quant_plane(dsvcu_ctx *c, dsvcu_coefs *k, int plane, int q, const dsvcu_fmeta *fm)
{
    return quant_run(c, k, q, fm, 1 << plane);
}

extern "C" int
dsvcu_quant_frame(dsvcu_ctx *c, dsvcu_coefs *k, int q, const dsvcu_fmeta *fm, int plane_mask)
{
    return quant_run(c, k, q, fm, plane_mask & 7);
}

extern "C" int
dsvcu_fetch_symbols(dsvcu_ctx *c, int plane, const dsvcu_symbol **syms, int *nsyms, int *dc)
{
    int n;
#ifndef DSVCU_EMU
    CK(cudaEventSynchronize(c->ev_sym[plane]));
#endif
    n = c->h_meta[plane];
    *syms = (const dsvcu_symbol *) c->h_syms[plane];
    *nsyms = n;
    *dc = c->h_meta[3 + plane];
    return 0;
}

/* Switch to the other set of host staging buffers (symbols, side information) and wait until
 * the device has finished copying out of it.  A decoder that calls this at the start of
 * every picture may parse picture n + 1 while picture n is still being copied / decoded. */
extern "C" int
dsvcu_staging_flip(dsvcu_ctx *c)
{
    int i;
    c->stage_cur ^= 1;
    for (i = 0; i < 3; i++) {
        if (!c->h_syms_set[c->stage_cur][i]) {
            CK(dsvcu_malloc_host(&c->h_syms_set[c->stage_cur][i], (size_t) c->sym_cap[i] * sizeof(dsvcu_sym)));
        }
        c->h_syms[i] = c->h_syms_set[c->stage_cur][i];
    }
    if (c->h_side_set[c->stage_cur]) c->h_side = c->h_side_set[c->stage_cur];
#ifndef DSVCU_EMU
    for (i = 0; i < 2; i++) {
        if (!c->ev_stage[i]) CK(cudaEventCreateWithFlags(&c->ev_stage[i], cudaEventDisableTiming | cudaEventBlockingSync));
    }
    CK(cudaEventSynchronize(c->ev_stage[c->stage_cur])); /* never recorded = complete */
#endif
    return 0;
}

extern "C" dsvcu_symbol *
dsvcu_symbol_staging(dsvcu_ctx *c, int plane, int *capacity)
{
    if (capacity) *capacity = c->sym_cap[plane];
    return (dsvcu_symbol *) c->h_syms[plane];
}

extern "C" int
dsvcu_dequant_plane(dsvcu_ctx *c, dsvcu_coefs *k, int plane, int q, const dsvcu_fmeta *fm,
                    int nsyms, const int level_start[5], int dc)
{
    const int w = k->w[plane], h = k->h[plane];
    int part[5], l;
    QuantLevel Q;
    int qf = q * 3 / 2;

    dsvcu_scan_layout(w, h, part);
    CK(dsvcu_memset_async(k->data[plane], 0, (size_t) w * h * sizeof(int32_t), c->stream));
    /* the raw DC rides behind the last symbol of the staging buffer */
    c->h_syms[plane][nsyms].pos = 0;
    c->h_syms[plane][nsyms].v = dc;
    CK(dsvcu_h2d_async(c->d_syms[plane], c->h_syms[plane], (size_t) (nsyms + 1) * sizeof(dsvcu_sym), c->stream));
#ifndef DSVCU_EMU
    if (c->ev_stage[c->stage_cur]) CK(cudaEventRecord(c->ev_stage[c->stage_cur], c->stream));
#endif
    if (nsyms > 0) {
        quant_level_geom(&Q, c, k, plane, qf, fm, -1, part);
        Q.syms = c->d_syms[plane];
        Q.sym_begin = level_start[0];
        Q.sym_end = level_start[1];
        if (Q.sym_end > Q.sym_begin) {
            DSVCU_LAUNCH(k_dequant_ll, grid_for(Q.sym_end - Q.sym_begin, 256), 256, 0, c->stream, Q,
                         fm->lossless ? 1 : lfquant(qf, plane, fm));
            CK_LAUNCH(c);
        }
        for (l = 0; l < 3; l++) {
            int wave;
            quant_level_geom(&Q, c, k, plane, qf, fm, l, part);
            Q.syms = c->d_syms[plane];
            Q.sym_begin = level_start[1 + l];
            Q.sym_end = level_start[2 + l];
            if (Q.sym_end <= Q.sym_begin) continue;
            for (wave = 0; wave < (fm->lossless ? 1 : 2); wave++) {
                Q.wave = wave;
                DSVCU_LAUNCH(k_dequant_hf, grid_for(Q.sym_end - Q.sym_begin, 256), 256, 0, c->stream, Q);
                CK_LAUNCH(c);
            }
        }
    }
    /* dst->data[0] = LL (hzcc.c:634) */
    CK(dsvcu_d2d_async(k->data[plane], &c->d_syms[plane][nsyms].v, sizeof(int), c->stream));
    return 0;
}


/* ---- entropy decode on the device (k_hzcc.cuh) ----
 *
 * dsvcu_parse_begin: `n` serialised coefficient planes (each starting at its 32-bit length
 * word) are gathered into pinned memory, copied to the device and parsed, one warp per plane,
 * on the context's parse streams -- beside whatever the context's main stream is doing.  The
 * batch comes in two parts with a launch, a stream and a completion event each: planes
 * [0, n_early) and the rest.  A part is ready when its LONGEST chain is, so the caller puts
 * the planes it needs first and that are short into part 0 and the long ones (pictures behind
 * a scene cut, needed late) into part 1.  Returns the set (0 / 1) the batch lives in.
 * dsvcu_parse_end waits for one part of that set and reports its planes:
 * ok[i] = 1: plane i was well-formed and its symbols are resident (dsvcu_dequant_parsed);
 * ok[i] = 0: the caller must parse that plane on the host, which reproduces the reference's
 * handling of damaged planes.  A set's symbols stay valid until the set is begun again, i.e.
 * for two batches. */

/* the number of (run, value) pairs a plane announces: [len32][SEG dc][align][count24] */
static uint32_t
plane_pair_count(const uint8_t *p, uint32_t len)
{
    uint32_t bit = 32, v = 1;
    const uint32_t nbits = len * 8u;
#define PBIT(b) ((b) < nbits ? (p[(b) >> 3] >> (7 - ((b) & 7))) & 1u : 1u)
    while (!PBIT(bit)) { /* (0 d)* 1 */
        v = (v << 1) | PBIT(bit + 1);
        bit += 2;
        if (bit > 32 + 66) return 0;
    }
    bit++;
    if (v != 1) bit++; /* sign of a nonzero DC */
#undef PBIT
    bit = (bit + 7) & ~7u;
    if (bit + 24 > nbits) return 0;
    return ((uint32_t) p[bit >> 3] << 16) | ((uint32_t) p[(bit >> 3) + 1] << 8) | p[(bit >> 3) + 2];
}

static int
parse_set_pending(const dsvcu_ctx *c, int set)
{
    return c->pset[set].pending[0] || c->pset[set].pending[1];
}

extern "C" int
dsvcu_parse_begin(dsvcu_ctx *c, const dsvcu_plane_bits *pl, int n, int n_early, const dsvcu_side_bits *sd, int nsd,
                  int nsd_early)
{
    size_t total = 0, nsyms = 0, at = 0, sat = 0;
    int i, set = 0, max_blk = 0;
    HzJob J;
    dsvcu_ctx::ParseSet *S;

    if (n <= 0) return fail_msg("dsvcu_parse_begin: no planes");
    if (!sd) nsd = nsd_early = 0;
    if (nsd < 0 || nsd_early < 0 || nsd_early > nsd) return fail_msg("dsvcu_parse_begin: bad split");
    for (i = 0; i < nsd; i++) {
        int k;
        if (!sd[i].base || sd[i].base_len > 0x0fffffffu || sd[i].nbh <= 0 || sd[i].nbv <= 0 ||
            (size_t) sd[i].nbh * sd[i].nbv > 0x00ffffffu) {
            return fail_msg("dsvcu_parse_begin: bad side information");
        }
        for (k = 0; k < HZ_SIDE_NSUB; k++) {
            if ((size_t) sd[i].off[k] + sd[i].len[k] > sd[i].base_len) return fail_msg("dsvcu_parse_begin: bad side information");
        }
        if (sd[i].nbh * sd[i].nbv > max_blk) max_blk = sd[i].nbh * sd[i].nbv;
        total += (((size_t) sd[i].base_len + 7) & ~(size_t) 7) + 16;
    }
    if (max_blk && ensure_blocks(c, max_blk)) return -1;
    /* the set that was begun longer ago; never one whose result has not been collected */
    if (n_early < 0 || n_early > n) return fail_msg("dsvcu_parse_begin: bad split");
    if (parse_set_pending(c, 0) && parse_set_pending(c, 1)) return fail_msg("dsvcu_parse_begin: two batches already in flight");
    set = parse_set_pending(c, 0) ? 1 : (parse_set_pending(c, 1) ? 0 : (c->pset_last ^ 1));
    S = &c->pset[set];
    for (i = 0; i < n; i++) {
        if (pl[i].len < 8 || pl[i].len > 0x3fffffffu || !pl[i].bits) return fail_msg("dsvcu_parse_begin: bad plane");
        total += ((size_t) pl[i].len + 7) & ~(size_t) 7;
    }
    if (total + 16 > 0xffffffffu / 8) return fail_msg("dsvcu_parse_begin: batch too large");
#ifndef DSVCU_EMU
    if (!c->pstream) CK(cudaStreamCreateWithFlags(&c->pstream, cudaStreamNonBlocking));
    if (!c->pstream2) CK(cudaStreamCreateWithFlags(&c->pstream2, cudaStreamNonBlocking));
    if (!S->ev_consumed) {
        CK(cudaEventCreateWithFlags(&S->ev_parsed[0], cudaEventDisableTiming | cudaEventBlockingSync));
        CK(cudaEventCreateWithFlags(&S->ev_parsed[1], cudaEventDisableTiming | cudaEventBlockingSync));
        CK(cudaEventCreateWithFlags(&S->ev_bits, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&S->ev_consumed, cudaEventDisableTiming));
    }
    /* de-quantiser launches of the batch that used this set last may still be queued */
    CK(cudaStreamWaitEvent(c->pstream, S->ev_consumed, 0));
#endif
    if (total + 16 > S->bits_cap || n > S->spans_cap) {
        /* (growing frees memory that queued work may still read) */
        CK(ctx_wait(c));
#ifndef DSVCU_EMU
        CK(cudaStreamSynchronize(c->pstream));
        CK(cudaStreamSynchronize(c->pstream2));
#endif
        if (total + 16 > S->bits_cap) {
            const size_t cap = total + total / 4 + 4096;
            if (S->h_bits) dsvcu_free_host(S->h_bits);
            if (S->d_bits) dsvcu_free_dev(S->d_bits);
            S->h_bits = S->d_bits = NULL;
            S->bits_cap = 0;
            CK(dsvcu_malloc_host(&S->h_bits, cap));
            CK(dsvcu_malloc(&S->d_bits, cap));
            S->bits_cap = cap;
        }
        if (n > S->spans_cap) {
            const int cap = n + 16;
            if (S->h_spans) dsvcu_free_host(S->h_spans);
            if (S->d_spans) dsvcu_free_dev(S->d_spans);
            if (S->h_meta) dsvcu_free_host(S->h_meta);
            if (S->d_meta) dsvcu_free_dev(S->d_meta);
            S->h_spans = S->d_spans = NULL;
            S->h_meta = S->d_meta = NULL;
            S->spans_cap = 0;
            CK(dsvcu_malloc_host(&S->h_spans, (size_t) cap * sizeof(HzSpan)));
            CK(dsvcu_malloc(&S->d_spans, (size_t) cap * sizeof(HzSpan)));
            CK(dsvcu_malloc_host(&S->h_meta, (size_t) cap * HZ_META_WORDS * sizeof(int)));
            CK(dsvcu_malloc(&S->d_meta, (size_t) cap * HZ_META_WORDS * sizeof(int)));
            S->spans_cap = cap;
        }
    }
    if (nsd > S->sides_cap || (size_t) nsd * c->side_bytes > S->side_out_cap) {
        CK(ctx_wait(c));
#ifndef DSVCU_EMU
        CK(cudaStreamSynchronize(c->pstream));
        CK(cudaStreamSynchronize(c->pstream2));
#endif
        if (nsd > S->sides_cap) {
            const int cap = nsd + 16;
            if (S->h_sides) dsvcu_free_host(S->h_sides);
            if (S->d_sides) dsvcu_free_dev(S->d_sides);
            if (S->h_side_ok) dsvcu_free_host(S->h_side_ok);
            if (S->d_side_ok) dsvcu_free_dev(S->d_side_ok);
            S->h_sides = S->d_sides = NULL;
            S->h_side_ok = S->d_side_ok = NULL;
            S->sides_cap = 0;
            CK(dsvcu_malloc_host(&S->h_sides, (size_t) cap * sizeof(HzSide)));
            CK(dsvcu_malloc(&S->d_sides, (size_t) cap * sizeof(HzSide)));
            CK(dsvcu_malloc_host(&S->h_side_ok, (size_t) cap * sizeof(int)));
            CK(dsvcu_malloc(&S->d_side_ok, (size_t) cap * sizeof(int)));
            S->sides_cap = cap;
        }
        if ((size_t) nsd * c->side_bytes > S->side_out_cap) {
            const size_t cap = (size_t) (nsd + 8) * c->side_bytes;
            if (S->d_side_out) dsvcu_free_dev(S->d_side_out);
            S->d_side_out = NULL;
            S->side_out_cap = 0;
            CK(dsvcu_malloc(&S->d_side_out, cap));
            S->side_out_cap = cap;
        }
    }
    S->side_stride = c->side_bytes;
    for (i = 0; i < n; i++) {
        HzSpan *sp = &S->h_spans[i];
        const size_t padded = ((size_t) pl[i].len + 7) & ~(size_t) 7;
        int part[5];
        uint32_t pairs = plane_pair_count(pl[i].bits, pl[i].len);
        /* a pair takes two bits or more and lands on a scan position of its own: anything
         * beyond that is a damaged plane, which the kernel will refuse */
        const uint32_t most = (uint32_t) dsvcu_scan_layout(pl[i].w, pl[i].h, part);
        if (pairs > most || (size_t) pairs > (size_t) pl[i].len * 4) pairs = 0;
        memcpy(S->h_bits + at, pl[i].bits, pl[i].len);
        memset(S->h_bits + at + pl[i].len, 0, padded - pl[i].len);
        sp->off = (uint32_t) at;
        sp->len = pl[i].len;
        sp->w = pl[i].w;
        sp->h = pl[i].h;
        sp->sym_base = (uint32_t) sat;
        sp->sym_cap = pairs;
        at += padded;
        sat += pairs;
    }
    for (i = 0; i < nsd; i++) {
        HzSide *D = &S->h_sides[i];
        const size_t padded = (((size_t) sd[i].base_len + 7) & ~(size_t) 7) + 16;
        int k;
        memcpy(S->h_bits + at, sd[i].base, sd[i].base_len);
        memset(S->h_bits + at + sd[i].base_len, 0, padded - sd[i].base_len);
        for (k = 0; k < HZ_SIDE_NSUB; k++) {
            D->off[k] = (uint32_t) at + sd[i].off[k];
            D->len[k] = sd[i].len[k];
        }
        D->nbh = sd[i].nbh;
        D->nbv = sd[i].nbv;
        D->flips = sd[i].flips;
        D->out_off = (uint32_t) ((size_t) i * c->side_bytes);
        D->mv_bytes = (uint32_t) c->side_mv_bytes;
        at += padded;
    }
    nsyms = sat;
    if (nsyms + 1 > S->syms_cap) {
        const size_t cap = nsyms + nsyms / 4 + 4096;
        CK(ctx_wait(c));
#ifndef DSVCU_EMU
        CK(cudaStreamSynchronize(c->pstream));
        CK(cudaStreamSynchronize(c->pstream2));
#endif
        if (S->d_syms) dsvcu_free_dev(S->d_syms);
        S->d_syms = NULL;
        S->syms_cap = 0;
        CK(dsvcu_malloc(&S->d_syms, cap * sizeof(dsvcu_sym)));
        S->syms_cap = cap;
    }
    memset(S->h_bits + at, 0, 16);
#ifndef DSVCU_EMU
    const cudaStream_t ps = c->pstream;
#else
    const dsvcu_stream_t ps = 0;
#endif
    CK(dsvcu_h2d_async(S->d_bits, S->h_bits, at + 16, ps));
    CK(dsvcu_h2d_async(S->d_spans, S->h_spans, (size_t) n * sizeof(HzSpan), ps));
    if (nsd) CK(dsvcu_h2d_async(S->d_sides, S->h_sides, (size_t) nsd * sizeof(HzSide), ps));
    J.bits = S->d_bits;
    J.syms = S->d_syms;
    J.side_out = S->d_side_out;
#ifndef DSVCU_EMU
    CK(cudaEventRecord(S->ev_bits, ps));
    CK(cudaStreamWaitEvent(c->pstream2, S->ev_bits, 0));
#endif
    for (int part = 0; part < 2; part++) {
        const int first = part ? n_early : 0, cnt = part ? n - n_early : n_early;
        const int sfirst = part ? nsd_early : 0, scnt = part ? nsd - nsd_early : nsd_early;
#ifndef DSVCU_EMU
        const cudaStream_t st = part ? c->pstream2 : c->pstream;
#else
        const dsvcu_stream_t st = 0;
#endif
        S->pending[part] = 0;
        if (!cnt && !scnt) continue;
        J.spans = S->d_spans + first;
        J.nspans = cnt;
        J.meta = S->d_meta + first * HZ_META_WORDS;
        J.sides = S->d_sides + sfirst;
        J.nsides = scnt;
        J.side_ok = S->d_side_ok + sfirst;
#ifndef DSVCU_EMU
        DSVCU_LAUNCH(k_hzcc_parse, (cnt + scnt + HZ_WARPS - 1) / HZ_WARPS, HZ_WARPS * 32, 0, st, J);
#else
        DSVCU_LAUNCH(k_hzcc_parse, cnt + scnt, 32, 0, st, J);
#endif
        CK_LAUNCH(c);
        if (cnt) {
            CK(dsvcu_d2h_async(S->h_meta + first * HZ_META_WORDS, S->d_meta + first * HZ_META_WORDS,
                               (size_t) cnt * HZ_META_WORDS * sizeof(int), st));
        }
        if (scnt) CK(dsvcu_d2h_async(S->h_side_ok + sfirst, S->d_side_ok + sfirst, (size_t) scnt * sizeof(int), st));
#ifndef DSVCU_EMU
        CK(cudaEventRecord(S->ev_parsed[part], st));
#endif
        S->pending[part] = 1;
    }
    S->n = n;
    S->n_early = n_early;
    S->nsd = nsd;
    S->nsd_early = nsd_early;
    c->pset_last = set;
    return set;
}

extern "C" int
dsvcu_parse_end(dsvcu_ctx *c, int set, int part, int *ok, int *side_ok)
{
    dsvcu_ctx::ParseSet *S;
    int i;
    if (set < 0 || set > 1 || part < 0 || part > 1 || !c->pset[set].n) return fail_msg("dsvcu_parse_end: no such batch");
    S = &c->pset[set];
    if (S->pending[part]) {
#ifndef DSVCU_EMU
        CK(cudaEventSynchronize(S->ev_parsed[part]));
#endif
        S->pending[part] = 0;
    }
    for (i = part ? S->n_early : 0; i < (part ? S->n : S->n_early); i++) {
        ok[i] = S->h_meta[i * HZ_META_WORDS + HZ_META_OK] == 1;
    }
    for (i = part ? S->nsd_early : 0; side_ok && i < (part ? S->nsd : S->nsd_early); i++) {
        side_ok[i] = S->h_side_ok[i] == 1;
    }
    return 0;
}

#ifdef DSVCU_EMU
static int g_emu_not_ready_every, g_emu_ready_calls;
extern "C" void
dsvcu_emu_parse_not_ready(int every)
{
    g_emu_not_ready_every = every;
    g_emu_ready_calls = 0;
}
#endif

/* has part `part` of the batch in `set` been parsed?  1: yes (dsvcu_parse_end will not wait),
 * 0: still running, -1: no such batch */
extern "C" int
dsvcu_parse_ready(dsvcu_ctx *c, int set, int part)
{
    dsvcu_ctx::ParseSet *S;
    if (set < 0 || set > 1 || part < 0 || part > 1 || !c->pset[set].n) return fail_msg("dsvcu_parse_ready: no such batch");
    S = &c->pset[set];
    if (!S->pending[part]) return 1;
#ifdef DSVCU_EMU
    /* test hook of the emulation build: every n-th question is answered "not yet"; -1: every
     * question about the second part of a batch */
    if (g_emu_not_ready_every > 0 && ++g_emu_ready_calls % g_emu_not_ready_every == 0) return 0;
    if (g_emu_not_ready_every < 0 && part == 1) return 0;
#else
    {
        const cudaError_t e = cudaEventQuery(S->ev_parsed[part]);
        if (e == cudaErrorNotReady) {
            (void) cudaGetLastError(); /* "not ready" is an answer, not an error: keep it away from CK_LAUNCH */
            return 0;
        }
        if (e != cudaSuccess) return fail("dsvcu_parse_ready", (int) e);
    }
#endif
    return 1;
}

/* a batch of one part without side information, begun and collected in one call: the batch is
 * set 0 or 1 (returned) */
extern "C" int
dsvcu_parse_planes(dsvcu_ctx *c, const dsvcu_plane_bits *pl, int n, int *ok)
{
    const int set = dsvcu_parse_begin(c, pl, n, n, NULL, 0, 0);
    if (set < 0) return -1;
    if (dsvcu_parse_end(c, set, 0, ok, NULL)) return -1;
    return set;
}

/* the vector field and block flags of picture `side` of a collected batch become the
 * context's current side information: the counterpart of dsvcu_set_side, device to device */
extern "C" int
dsvcu_set_side_parsed(dsvcu_ctx *c, int set, int side, int nblocks)
{
    const dsvcu_ctx::ParseSet *S;
    if (set < 0 || set > 1) return fail_msg("dsvcu_set_side_parsed: no such batch");
    S = &c->pset[set];
    if (side < 0 || side >= S->nsd || S->pending[side >= S->nsd_early] || S->h_side_ok[side] != 1 ||
        nblocks != S->h_sides[side].nbh * S->h_sides[side].nbv || S->side_stride != c->side_bytes) {
        return fail_msg("dsvcu_set_side_parsed: no such picture");
    }
    if (ensure_blocks(c, nblocks)) return -1;
    CK(dsvcu_d2d_async(c->d_side[c->side_cur], S->d_side_out + (size_t) side * S->side_stride,
                       c->side_mv_bytes + (size_t) nblocks, c->stream));
    return 0;
}

/* number of symbols the device parser found in plane `span` of a collected batch (tests) */
extern "C" int
dsvcu_parsed_count(dsvcu_ctx *c, int set, int span)
{
    const dsvcu_ctx::ParseSet *S;
    if (set < 0 || set > 1) return -1;
    S = &c->pset[set];
    if (S->pending[span >= S->n_early] || span < 0 || span >= S->n || S->h_meta[span * HZ_META_WORDS + HZ_META_OK] != 1) return -1;
    return S->h_meta[span * HZ_META_WORDS + HZ_META_NSYM];
}

/* de-quantise the three planes of one picture from spans first_span .. first_span + 2 of a
 * collected batch (all three must have been ok): the counterpart of three
 * dsvcu_dequant_plane calls, 7 launches, nothing crosses the bus */
extern "C" int
dsvcu_dequant_parsed(dsvcu_ctx *c, dsvcu_coefs *k, int q, const dsvcu_fmeta *fm, int set, int first_span)
{
    DequantJob J;
    const dsvcu_ctx::ParseSet *S;
    int part[3][5], p, l, wave, most = 0;
    const int qf = q * 3 / 2;

    if (set < 0 || set > 1) return fail_msg("dsvcu_dequant_parsed: no such batch");
    S = &c->pset[set];
    if (first_span < 0 || first_span + 3 > S->n || S->pending[first_span >= S->n_early] || S->pending[first_span + 2 >= S->n_early]) {
        return fail_msg("dsvcu_dequant_parsed: no such planes");
    }
    for (p = 0; p < 3; p++) {
        const HzSpan *sp = &S->h_spans[first_span + p];
        const int *m = S->h_meta + (first_span + p) * HZ_META_WORDS;
        if (m[HZ_META_OK] != 1 || sp->w != k->w[p] || sp->h != k->h[p]) {
            return fail_msg("dsvcu_dequant_parsed: plane was not parsed for this geometry");
        }
        dsvcu_scan_layout(k->w[p], k->h[p], part[p]);
        J.meta[p] = S->d_meta + (first_span + p) * HZ_META_WORDS;
        J.lfq[p] = fm->lossless ? 1 : lfquant(qf, p, fm);
        if (m[HZ_META_NSYM] > most) most = m[HZ_META_NSYM];
    }
    /* the three planes are one allocation (dsvcu_coefs_create): one zero-fill */
    CK(dsvcu_memset_async(k->alloc, 0,
                          ((size_t) k->w[0] * k->h[0] + (size_t) k->w[1] * k->h[1] + (size_t) k->w[2] * k->h[2]) * sizeof(int32_t),
                          c->stream));
    /* grid for the fullest level of the fullest plane (grid-stride loops inside) */
    const int grid = grid_for(most > 0 ? most : 1, 256 * 4);
    for (l = -1; l < 3; l++) {
        for (p = 0; p < 3; p++) {
            quant_level_geom(&J.Q[p], c, k, p, qf, fm, l, part[p]);
            J.Q[p].syms = S->d_syms + S->h_spans[first_span + p].sym_base;
        }
        if (l < 0) {
            DSVCU_LAUNCH(k_dequant_ll_m, dim3(grid, 3, 1), 256, 0, c->stream, J);
            CK_LAUNCH(c);
            continue;
        }
        for (wave = 0; wave < (fm->lossless ? 1 : 2); wave++) {
            DSVCU_LAUNCH(k_dequant_hf_m, dim3(grid, 3, 1), 256, 0, c->stream, J, wave);
            CK_LAUNCH(c);
        }
    }
#ifndef DSVCU_EMU
    CK(cudaEventRecord(S->ev_consumed, c->stream));
#endif
    return 0;
}

/* ----------------------------------------------- motion compensation, filters */

static void
bmc_fill(BmcArgs *A, dsvcu_ctx *c, const dsvcu_fmeta *fm, dsvcu_frame *ref, dsvcu_frame *pred, dsvcu_frame *res,
         dsvcu_frame *out, int mode)
{
    int i;
    memset(A, 0, sizeof(*A));
    for (i = 0; i < 3; i++) {
        BmcPlane *P = &A->pl[i];
        if (ref) {
            P->ref = ref->p[i].data;
            P->ref_stride = ref->p[i].stride;
        }
        if (pred) {
            P->pred = pred->p[i].data;
            P->pred_stride = pred->p[i].stride;
        }
        if (res) {
            P->res = res->p[i].data;
            P->res_stride = res->p[i].stride;
            P->src = P->res;
            P->src_stride = P->res_stride;
            P->w = res->p[i].w;
            P->h = res->p[i].h;
        }
        if (out) {
            P->out = out->p[i].data;
            P->out_stride = out->p[i].stride;
        }
        P->sh = i ? FMT_HSHIFT(c->subsamp) : 0;
        P->sv = i ? FMT_VSHIFT(c->subsamp) : 0;
    }
    A->mvs = c->d_mvs;
    A->nbh = fm->nblocks_h;
    A->nbv = fm->nblocks_v;
    A->blk_w = fm->blk_w;
    A->blk_h = fm->blk_h;
    A->tmc = fm->temporal_mc;
    A->lossless = fm->lossless;
    A->mode = mode;
}

static int
filter_q(const dsvcu_fmeta *fm, int q) /* compute_filter_q, bmc.c:376-388 */
{
    int psyf = psy_factor(fm, -1);
    if (q > 1536) q = 1536;
    q += q * psyf >> (7 + 3);
    if (q < 1024) q = 512 + q / 2;
    return q;
}

/* k_filter.cuh, skewed lockstep: FILT_CELLS cell rows per CTA, the planes of the
 * job back to back in one grid */
static int
run_filter_job(dsvcu_ctx *c, FiltJob *J, int nplanes)
{
    int i, ctas = 0, bands = 0, smem = 0;
    for (i = 0; i < nplanes; i++) {
        FiltArgs *F = &J->p[i];
        int b = filt_smem_bytes(*F);
        if (F->nrows <= 0 || F->ncols <= 0) F->nrows = 0;
        if (b > smem) smem = b;
        F->first_cta = ctas;
        ctas += filt_ctas(*F);
        bands += (F->nrows + FILT_G - 1) / FILT_G;
    }
    J->nplanes = nplanes;
    if (!ctas || DIAG_SKIP(16)) return 0;
    if (smem > 200 * 1024) {
        snprintf(g_err, sizeof(g_err), "picture too wide for the filter schedule");
        return -1;
    }
    if (bands + 1 > c->progress_cap) {
        dsvcu_free_dev(c->d_progress);
        c->progress_cap = bands + 64;
        CK(dsvcu_malloc(&c->d_progress, (size_t) c->progress_cap * sizeof(int)));
    }
    CK(dsvcu_memset_async(c->d_progress, 0, (size_t) (bands + 1) * sizeof(int), c->stream));
    J->ticket = c->d_progress + bands;
    bands = 0;
    for (i = 0; i < nplanes; i++) {
        J->p[i].progress = c->d_progress + bands;
        bands += (J->p[i].nrows + FILT_G - 1) / FILT_G;
    }
#ifndef DSVCU_EMU
    if (smem > 48 * 1024 && !c->filt_big_smem) {
        CK(cudaFuncSetAttribute(k_filter_skew, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        c->filt_big_smem = 1;
    }
#endif
    DSVCU_LAUNCH(k_filter_skew, ctas, FILT_THREADS, smem, c->stream, *J);
    CK_LAUNCH(c);
    return 0;
}

static void
filt_common(FiltArgs *F, dsvcu_ctx *c, const dsvcu_fmeta *fm)
{
    memset(F, 0, sizeof(*F));
    F->mvs = c->d_mvs;
    F->blockdata = c->d_blockdata;
    F->nbh = fm->nblocks_h;
    F->nbv = fm->nblocks_v;
    F->blk_w = fm->blk_w;
    F->blk_h = fm->blk_h;
}

static int
loop_filters(dsvcu_ctx *c, const dsvcu_fmeta *fm, int q, dsvcu_frame *f, int do_filter)
{
    /* luma_filter + chroma_filter, bmc.c:459-659: the three planes are
     * independent and run as three CTAs of one launch */
    FiltJob J;
    int i;
    if (fm->lossless) return 0;
    for (i = 0; i < 3; i++) {
        FiltArgs *F = &J.p[i];
        filt_common(F, c, fm);
        F->data = f->p[i].data;
        F->stride = f->p[i].stride;
        F->w = f->p[i].w;
        F->h = f->p[i].h;
        F->do_filter = do_filter;
        if (i == 0) {
            F->q = filter_q(fm, q);
            F->fthresh = 32 * (14 - ilb2((unsigned) F->q));
            F->sharpen = fm->inter_sharpen ? fm->temporal_mc : 0;
            F->ncols = F->w / 4;
            F->nrows = F->h / 4;
            F->mode = FILT_MODE_LUMA;
        } else {
            F->q = q;
            F->bw = fm->blk_w >> FMT_HSHIFT(c->subsamp);
            F->bh = fm->blk_h >> FMT_VSHIFT(c->subsamp);
            F->ncols = fm->nblocks_h;
            F->nrows = fm->nblocks_v;
            F->mode = FILT_MODE_CHROMA;
        }
    }
    return run_filter_job(c, &J, 3);
}

extern "C" int
dsvcu_sub_pred(dsvcu_ctx *c, const dsvcu_fmeta *fm, dsvcu_frame *pred, dsvcu_frame *resd, dsvcu_frame *ref)
{
    BmcArgs A;
    bmc_fill(&A, c, fm, ref, pred, resd, NULL, 0);
    DSVCU_LAUNCH(k_predict, dim3(fm->nblocks_h * fm->nblocks_v, 1, 1), BMC_THREADS, 0, c->stream, A);
    CK_LAUNCH(c);
    return 0;
}

/* dsv_sub_pred with the source picture read from `src` and the residual written
 * to `resd`: no clone of the source into the residual frame beforehand
 * (dsv_encoder.c:1292).  Like the reference's subtract(), the kernel works on whole
 * blocks, i.e. also on the border columns / rows the block grid covers past the
 * picture edge -- source pixels there come from the padded source's border, and
 * the residual written there is what the forward transform reads for the one
 * extra column of an odd-width chroma plane (sbt.c:807). */
extern "C" int
dsvcu_sub_pred_from(dsvcu_ctx *c, const dsvcu_fmeta *fm, dsvcu_frame *pred, dsvcu_frame *resd, dsvcu_frame *ref,
                    dsvcu_frame *src)
{
    BmcArgs A;
    int i;
    if (DIAG_SKIP(4)) return 0;
    bmc_fill(&A, c, fm, ref, pred, resd, NULL, 0);
    for (i = 0; i < 3; i++) {
        A.pl[i].src = src->p[i].data;
        A.pl[i].src_stride = src->p[i].stride;
    }
    DSVCU_LAUNCH(k_predict, dim3(fm->nblocks_h * fm->nblocks_v, 1, 1), BMC_THREADS, 0, c->stream, A);
    CK_LAUNCH(c);
    return 0;
}

extern "C" int
dsvcu_add_pred(dsvcu_ctx *c, const dsvcu_fmeta *fm, int q, dsvcu_frame *resd, dsvcu_frame *out, dsvcu_frame *ref,
               int do_filter)
{
    BmcArgs A;
    bmc_fill(&A, c, fm, ref, NULL, resd, out, 1);
    DSVCU_LAUNCH(k_predict, dim3(fm->nblocks_h * fm->nblocks_v, 1, 1), BMC_THREADS, 0, c->stream, A);
    CK_LAUNCH(c);
    return loop_filters(c, fm, q, out, do_filter);
}

extern "C" int
dsvcu_add_res(dsvcu_ctx *c, const dsvcu_fmeta *fm, int q, dsvcu_frame *resd, dsvcu_frame *pred, int do_filter)
{
    BmcArgs A;
    bmc_fill(&A, c, fm, NULL, pred, resd, NULL, 0);
    if (!DIAG_SKIP(8))
    DSVCU_LAUNCH(k_reconstruct, dim3(grid_for(fm->nblocks_h * fm->nblocks_v * fm->blk_h * (fm->blk_w / 16 > 0 ? fm->blk_w / 16 : 1), REC_THREADS), 3, 1),
                 REC_THREADS, 0, c->stream, A);
    CK_LAUNCH(c);
    return loop_filters(c, fm, q, resd, do_filter);
}

extern "C" int
dsvcu_intra_filter(dsvcu_ctx *c, int q, const dsvcu_fmeta *fm, int plane, dsvcu_frame *f, int do_filter)
{
    FiltJob J;
    FiltArgs *F = &J.p[0];
    if (fm->lossless || plane != 0 || !do_filter) return 0;
    memset(&J, 0, sizeof(J));
    filt_common(F, c, fm);
    F->data = f->p[0].data;
    F->stride = f->p[0].stride;
    F->w = f->p[0].w;
    F->h = f->p[0].h;
    F->q = filter_q(fm, q);
    F->fthresh = 32 * (14 - ilb2((unsigned) F->q));
    F->do_filter = 1;
    F->ncols = F->w / 4;
    F->nrows = F->h / 4;
    F->mode = FILT_MODE_INTRA;
    return run_filter_job(c, &J, 1);
}

extern "C" int
dsvcu_post_process(dsvcu_ctx *c, dsvcu_frame *f)
{
    dsvcu_plane_t *p = &f->p[0];
    DSVCU_LAUNCH(k_post_sharpen, grid_for((p->w / 4) * (p->h / 4), 256), 256, 0, c->stream, p->data, p->stride, p->w,
                 p->h);
    CK_LAUNCH(c);
    return 0;
}

/* ------------------------------------------------------------ frame helpers */

extern "C" int
dsvcu_extend_frame(dsvcu_ctx *c, dsvcu_frame *f, int luma_only)
{
    ExtArgs A;
    int i, n = luma_only ? 1 : f->nplanes, maxitems = 0;
    memset(&A, 0, sizeof(A));
    for (i = 0; i < n; i++) {
        int items = f->p[i].h + (f->p[i].w + 3) / 4 + 4;
        A.pl[i].data = f->p[i].data;
        A.pl[i].stride = f->p[i].stride;
        A.pl[i].w = f->p[i].w;
        A.pl[i].h = f->p[i].h;
        if (items > maxitems) maxitems = items;
    }
    DSVCU_LAUNCH(k_extend, dim3(grid_for(maxitems, 64), n, 1), 64, 0, c->stream, A);
    CK_LAUNCH(c);
    return 0;
}

extern "C" int
dsvcu_ds2x_luma(dsvcu_ctx *c, dsvcu_frame *dst, dsvcu_frame *src)
{
    dsvcu_plane_t *d = &dst->p[0], *s = &src->p[0];
    DSVCU_LAUNCH(k_ds2x, grid_for(d->w * d->h, 256), 256, 0, c->stream, d->data, d->stride, d->w, d->h, s->data,
                 s->stride);
    CK_LAUNCH(c);
    return 0;
}

extern "C" int
dsvcu_frame_copy(dsvcu_ctx *c, dsvcu_frame *dst, dsvcu_frame *src)
{
    int i;
    for (i = 0; i < dst->nplanes; i++) {
        CK(dsvcu_d2d_2d_async(dst->p[i].data, dst->p[i].stride, src->p[i].data, src->p[i].stride, src->p[i].w,
                              dst->p[i].h, c->stream));
    }
    return dsvcu_extend_frame(c, dst, 0);
}

/* ------------------------------------------------- motion estimation (encoder) */

struct dsvcu_pyramid {
    int levels;
    dsvcu_frame *f[ME_MAXLVL + 1]; /* [1..levels] */
};

extern "C" int
dsvcu_pyramid_create(dsvcu_ctx *c, dsvcu_pyramid **out, int levels)
{
    dsvcu_pyramid *p = (dsvcu_pyramid *) calloc(1, sizeof(*p));
    int i;
    if (!p || levels > ME_MAXLVL) return -1;
    p->levels = levels;
    for (i = 1; i <= levels; i++) {
        if (dsvcu_frame_create_luma(c, &p->f[i], RSHIFT_UP(c->width, i), RSHIFT_UP(c->height, i))) return -1;
    }
    *out = p;
    return 0;
}

extern "C" void
dsvcu_pyramid_destroy(dsvcu_ctx *c, dsvcu_pyramid *p)
{
    int i;
    if (!p) return;
    for (i = 1; i <= p->levels; i++) dsvcu_frame_destroy(c, p->f[i]);
    free(p);
}

extern "C" dsvcu_frame *
dsvcu_pyramid_level(dsvcu_pyramid *p, int level)
{
    return (level >= 1 && level <= p->levels) ? p->f[level] : NULL;
}

/* border of `base` (all its planes when extend_base, else taken as already
 * extended) + every level of the pyramid, in two launches (k_frame.cuh) */
static int
pyramid_run(dsvcu_ctx *c, dsvcu_pyramid *p, dsvcu_frame *base, int extend_base)
{
    PyrArgs A;
    int i, tiles;
    if (DIAG_SKIP(32)) return 0;
    memset(&A, 0, sizeof(A));
    for (i = 0; i < base->nplanes; i++) {
        A.base[i].data = base->p[i].data;
        A.base[i].stride = base->p[i].stride;
        A.base[i].w = base->p[i].w;
        A.base[i].h = base->p[i].h;
    }
    A.nbase = extend_base ? base->nplanes : 0;
    A.levels = p ? p->levels : 0;
    for (i = 1; i <= A.levels; i++) {
        A.lv[i].data = p->f[i]->p[0].data;
        A.lv[i].stride = p->f[i]->p[0].stride;
        A.lv[i].w = p->f[i]->p[0].w;
        A.lv[i].h = p->f[i]->p[0].h;
    }
    tiles = A.levels > 0 ? ((base->p[0].w + PYR_TILE - 1) / PYR_TILE) * ((base->p[0].h + PYR_TILE - 1) / PYR_TILE) : 0;
    A.ntiles = tiles;
    if (tiles > 0 || A.nbase > 0) {
        /* tiles of the base picture, then (when asked to) 16 CTAs for the base picture's border */
        DSVCU_LAUNCH(k_pyr_interior, tiles + (A.nbase > 0 ? 16 : 0), 256, 0, c->stream, A);
        CK_LAUNCH(c);
    }
    if (A.levels > 0) {
        DSVCU_LAUNCH(k_pyr_borders, 1, PYR_BORDER_THREADS, 0, c->stream, A);
        CK_LAUNCH(c);
    }
    return 0;
}

extern "C" int
dsvcu_pyramid_build(dsvcu_ctx *c, dsvcu_pyramid *p, dsvcu_frame *base)
{
    return pyramid_run(c, p, base, 0);
}

/* dsv_extend_frame(base) followed by mk_pyramid(base) */
extern "C" int
dsvcu_extend_pyramid(dsvcu_ctx *c, dsvcu_frame *base, dsvcu_pyramid *p)
{
    return pyramid_run(c, p, base, 1);
}

static int
ensure_mvf(dsvcu_ctx *c, int nblk)
{
    int i;
    if (ensure_blocks(c, nblk)) return -1;
    if (nblk <= c->mvf_cap) return 0;
    /* one block: [d_me (16 ints)] [progress words of every level] [vector fields of levels 1..5] */
    {
        const size_t mvf_bytes = (((size_t) nblk + 4) * sizeof(dsvcu_mv) + 255) & ~(size_t) 255;
        const int prog_cap = c->height / 8 + 64; /* rows of 16-px blocks at level 0, with room */
        const size_t prog_bytes = (((size_t) prog_cap * sizeof(int)) + 255) & ~(size_t) 255;
        size_t off = 256;
        if (c->d_me_zero) dsvcu_free_dev(c->d_me_zero);
        c->d_me_zero = NULL;
        c->me_zero_bytes = 256 + (ME_MAXLVL + 1) * prog_bytes + ME_MAXLVL * mvf_bytes;
        CK(dsvcu_malloc(&c->d_me_zero, c->me_zero_bytes));
        c->d_me = (int *) c->d_me_zero;
        for (i = 0; i <= ME_MAXLVL; i++) {
            c->d_me_prog[i] = (int *) (c->d_me_zero + off);
            off += prog_bytes;
        }
        c->me_prog_cap = prog_cap;
        for (i = 1; i <= ME_MAXLVL; i++) {
            c->d_mvf[i] = (dsvcu_mv *) (c->d_me_zero + off);
            off += mvf_bytes;
        }
    }
    if (c->d_pre) dsvcu_free_dev(c->d_pre);
    c->d_pre = NULL;
    CK(dsvcu_malloc(&c->d_pre, ((size_t) nblk + 4) * sizeof(MePre)));
    c->mvf_cap = nblk;
    return 0;
}

extern "C" int
dsvcu_set_prev_mvs(dsvcu_ctx *c, const void *mvs, int n)
{
    if (ensure_mvf(c, n)) return -1;
    CK(dsvcu_h2d_async(c->d_prev_mvf, mvs, (size_t) n * sizeof(dsvcu_mv), c->stream));
    return 0;
}

static void
me_plane(MePlane *m, dsvcu_frame *f, int plane)
{
    m->data = f->p[plane].data;
    m->stride = f->p[plane].stride;
    m->w = f->p[plane].w;
    m->h = f->p[plane].h;
}

extern "C" int
dsvcu_hme(dsvcu_ctx *c, const dsvcu_fmeta *fm, const dsvcu_hme_params *hp, dsvcu_frame *src, dsvcu_pyramid *src_pyr,
          dsvcu_frame *ref, dsvcu_pyramid *ref_pyr, dsvcu_frame *ogr, dsvcu_pyramid *ogr_pyr)
{
    const int nblk = fm->nblocks_h * fm->nblocks_v;
    int lvl;
    if (ensure_mvf(c, nblk)) return -1;
    c->me_nblk = nblk;
    c->d_mvf[0] = c->d_mvs;
    /* accumulators, progress words and the fields of levels 1..n in one go, level 0's field apart */
    CK(dsvcu_memset_async(c->d_me_zero, 0, c->me_zero_bytes, c->stream));
    CK(dsvcu_memset_async(c->d_mvs, 0, (size_t) nblk * sizeof(dsvcu_mv), c->stream));
    for (lvl = hp->pyramid_levels; lvl >= 0; lvl--) {
        MeArgs A;
        int step = 1 << lvl, rows, ctas;
        dsvcu_frame *fs = lvl ? src_pyr->f[lvl] : src;
        dsvcu_frame *fr = lvl ? ref_pyr->f[lvl] : ref;
        dsvcu_frame *fo = lvl ? ogr_pyr->f[lvl] : ogr;
        memset(&A, 0, sizeof(A));
        me_plane(&A.src[0], fs, 0);
        me_plane(&A.ref[0], fr, 0);
        me_plane(&A.ogr, fo, 0);
        if (lvl == 0) {
            me_plane(&A.src[1], fs, 1);
            me_plane(&A.src[2], fs, 2);
            me_plane(&A.ref[1], fr, 1);
            me_plane(&A.ref[2], fr, 2);
        }
        A.mvf = c->d_mvf[lvl];
        A.parent = (lvl < hp->pyramid_levels) ? c->d_mvf[lvl + 1] : NULL;
        A.ref_mvf = hp->use_prev_mvs ? c->d_prev_mvf : NULL;
        A.nxb = fm->nblocks_h;
        A.nyb = fm->nblocks_v;
        A.y_w = fm->blk_w;
        A.y_h = fm->blk_h;
        A.level = lvl;
        A.quant = hp->quant;
        A.effort = fm->effort;
        A.lossless = fm->lossless;
        A.skip_thresh = hp->skip_block_thresh;
        A.hs = FMT_HSHIFT(c->subsamp);
        A.vs = FMT_VSHIFT(c->subsamp);
        A.vid_w = c->width;
        A.vid_h = c->height;
        A.psyscale = psy_factor(fm, -1);
        A.gxy = c->d_me;
        A.acc = c->d_me + 2;
        rows = (fm->nblocks_v + step - 1) / step;
        A.nrows = rows;
        if (rows >= c->me_prog_cap) {
            snprintf(g_err, sizeof(g_err), "motion search: %d block rows exceed the progress table", rows);
            return -1;
        }
        A.progress = c->d_me_prog[lvl];
        A.ticket = c->d_me_prog[lvl] + c->me_prog_cap - 1; /* last word of the level's (zeroed) progress region */
        A.pre = c->d_pre;
        A.b2sr = (256 * (hp->quant * hp->quant >> 12) * fm->blk_w * fm->blk_h) / (c->width * c->height);
        {
            /* everything a block needs that does not depend on its same-level
             * neighbours, for all blocks at once */
            int cols = (fm->nblocks_h + step - 1) / step;
            int pctas = (cols * rows + ME_PRE_GROUPS - 1) / ME_PRE_GROUPS;
#ifdef DSVCU_EMU
            pctas = 1;
#endif
            if (pctas > 148 * 8) pctas = 148 * 8;
            if (g_pre_cap > 0 && pctas > g_pre_cap) pctas = g_pre_cap;
            if (!DIAG_SKIP(64))
            DSVCU_LAUNCH(k_me_prepass, pctas, ME_PRE_THREADS, 0, c->stream, A);
            CK_LAUNCH(c);
            if (lvl == 0 && fm->effort >= 4) {
                /* the two sub-pel measurements of every block, one warp each */
                int sctas = (2 * nblk + ME_SP_WARPS - 1) / ME_SP_WARPS;
#ifdef DSVCU_EMU
                sctas = 1;
#endif
                if (sctas > 148 * 8) sctas = 148 * 8;
                if (!DIAG_SKIP(64))
                DSVCU_LAUNCH(k_me_subpel, sctas, ME_SP_WARPS * 32, 0, c->stream, A);
                CK_LAUNCH(c);
            }
        }
        /* four block rows per warp; CTA k only waits for CTA k-1, dispatched first */
        ctas = (rows + ME_LVL_ROWS - 1) / ME_LVL_ROWS;
#ifndef DSVCU_EMU
        if (!c->me_smem_set) {
            CK(cudaFuncSetAttribute(k_me_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(MeLvlShared)));
            c->me_smem_set = 1;
        }
#endif
        if (!DIAG_SKIP(128))
        DSVCU_LAUNCH(k_me_level, ctas, ME_LVL_WARPS * 32, sizeof(MeLvlShared), c->stream, A);
        CK_LAUNCH(c);
        if (lvl != 0) {
            DSVCU_LAUNCH(k_me_global, 1, 256, 0, c->stream, c->d_mvf[lvl], fm->nblocks_h, fm->nblocks_v, lvl, c->d_me);
            CK_LAUNCH(c);
        }
    }
    return 0;
}

extern "C" int
dsvcu_mvs_to_prev(dsvcu_ctx *c, int nblocks)
{
    if (ensure_mvf(c, nblocks)) return -1;
    CK(dsvcu_d2d_async(c->d_prev_mvf, c->d_mvs, (size_t) nblocks * sizeof(dsvcu_mv), c->stream));
    return 0;
}

/* pinned mirror for block arrays read back by the host */
static int
ensure_hmvs(dsvcu_ctx *c, int nblk)
{
    if (nblk <= c->h_mvs_cap) return 0;
    if (c->h_mvs) dsvcu_free_host(c->h_mvs);
    c->h_mvs = NULL;
    CK(dsvcu_malloc_host(&c->h_mvs, ((size_t) nblk + 4) * sizeof(dsvcu_mv)));
    c->h_mvs_cap = nblk;
    return 0;
}

extern "C" int
dsvcu_hme_fetch(dsvcu_ctx *c, void *mvs_out, int nblocks, int *intra_pct, int *scene_change_blocks, int *avg_err)
{
    int elig;
    if (ensure_hmvs(c, nblocks)) return -1;
    CK(dsvcu_d2h_async(c->h_me, c->d_me, 16 * sizeof(int), c->stream));
    CK(dsvcu_d2h_async(c->h_mvs, c->d_mvs, (size_t) nblocks * sizeof(dsvcu_mv), c->stream));
    CK(ctx_wait(c));
    memcpy(mvs_out, c->h_mvs, (size_t) nblocks * sizeof(dsvcu_mv));
#if defined(ME_COUNT) && defined(DSVCU_EMU)
    {
        static const char *nm[8] = { "blocks", "decision landed on S", "metric look-ups", "  measured on demand", "sub-pel passes on demand",
                                     "reference statistics on demand", "max-sub-block metrics on demand", "intra error sums on demand" };
        for (int k = 0; k < 8; k++) fprintf(stderr, "  [me level 0 wavefront] %-34s %8ld\n", nm[k], g_me_cnt[k]);
        memset(g_me_cnt, 0, sizeof(g_me_cnt));
    }
#endif
    elig = c->h_me[4] ? c->h_me[4] : 1;
    *intra_pct = (c->h_me[2] * 100) / nblocks;
    *scene_change_blocks = c->h_me[3] * 100 / elig;
    *avg_err = (int) ((unsigned) c->h_me[5] / (unsigned) nblocks);
    return 0;
}

/* work counters of the last dsvcu_hme (valid after dsvcu_hme_fetch): [0] full-block metric
 * evaluations (SSE at levels > 1, the psycho-visual metric below), [1] sub-pel position metrics */
extern "C" int
dsvcu_hme_counters(dsvcu_ctx *c, long long out[2])
{
    out[0] = c->h_me[8];
    out[1] = c->h_me[9];
    return 0;
}

extern "C" int
dsvcu_intra_analysis_async(dsvcu_ctx *c, const dsvcu_fmeta *fm, dsvcu_frame *src, int nblocks)
{
    IaArgs A;
    int ctas;
    if (ensure_mvf(c, nblocks) || ensure_hmvs(c, nblocks)) return -1;
    memset(&A, 0, sizeof(A));
    me_plane(&A.src[0], src, 0);
    me_plane(&A.src[1], src, 1);
    me_plane(&A.src[2], src, 2);
    A.out = c->d_mvf[1]; /* scratch field; intra pictures run no motion search */
    A.nxb = fm->nblocks_h;
    A.nyb = fm->nblocks_v;
    A.y_w = fm->blk_w;
    A.y_h = fm->blk_h;
    A.hs = FMT_HSHIFT(c->subsamp);
    A.vs = FMT_VSHIFT(c->subsamp);
    A.do_psy = fm->do_psy;
    A.scale = 2 * psy_factor(fm, -1);
    ctas = (nblocks + ME_WARPS_PER_CTA - 1) / ME_WARPS_PER_CTA;
    if (ctas > 148 * 8) ctas = 148 * 8;
    DSVCU_LAUNCH(k_intra_analysis, ctas, ME_WARPS_PER_CTA * 32, 0, c->stream, A);
    CK_LAUNCH(c);
    CK(dsvcu_d2h_async(c->h_mvs, A.out, (size_t) nblocks * sizeof(dsvcu_mv), c->stream));
    return 0;
}

extern "C" int
dsvcu_intra_analysis_fetch(dsvcu_ctx *c, void *mvs_out, int nblocks)
{
    CK(ctx_wait(c));
    memcpy(mvs_out, c->h_mvs, (size_t) nblocks * sizeof(dsvcu_mv));
    return 0;
}

extern "C" int
dsvcu_intra_analysis(dsvcu_ctx *c, const dsvcu_fmeta *fm, dsvcu_frame *src, void *mvs_out, int nblocks)
{
    if (dsvcu_intra_analysis_async(c, fm, src, nblocks)) return -1;
    return dsvcu_intra_analysis_fetch(c, mvs_out, nblocks);
}

/* frame_luma_avg (dsv_encoder.c:108-127): sum over rows of (row sum / w), / h */
DSVCU_KERNEL void __launch_bounds__(256)
k_luma_avg(const uint8_t *data, int stride, int w, int h, int *out)
{
    DSVCU_SHARED unsigned tot;
    unsigned acc = 0;
    if (DSVCU_TID == 0) tot = 0;
    DSVCU_SYNC();
    PAR_FOR(j, h) {
        unsigned rav = 0;
        for (int i = 0; i < w; i++) rav += data[(size_t) j * stride + i];
        acc += rav / (unsigned) w;
    }
    atomicAdd(&tot, acc);
    DSVCU_SYNC();
    if (DSVCU_TID == 0) *out = (int) (tot / (unsigned) h);
}

extern "C" int
dsvcu_frame_luma_avg_async(dsvcu_ctx *c, dsvcu_frame *f)
{
    DSVCU_LAUNCH(k_luma_avg, 1, 256, 0, c->stream, f->p[0].data, f->p[0].stride, f->p[0].w, f->p[0].h, c->d_lavg);
    CK_LAUNCH(c);
    CK(dsvcu_d2h_async(c->h_lavg, c->d_lavg, sizeof(int), c->stream));
    return 0;
}

extern "C" unsigned
dsvcu_frame_luma_avg_result(dsvcu_ctx *c)
{
    return (unsigned) *c->h_lavg;
}

extern "C" int
dsvcu_frame_luma_avg(dsvcu_ctx *c, dsvcu_frame *f, unsigned *avg)
{
    if (dsvcu_frame_luma_avg_async(c, f)) return -1;
    CK(ctx_wait(c));
    *avg = dsvcu_frame_luma_avg_result(c);
    return 0;
}

/* ------------------------------------------------------ shared-memory carve-out
 *
 * An SM switches its L1 / shared-memory split only when it is empty.  With tens
 * of encoder instances the CTAs of twenty different kernels want to share SMs;
 * if every kernel asked for its own split they could only follow each other.
 * All kernels of the library therefore ask for the same carve-out (enough for
 * four search CTAs or the filter's largest job), once per device. */
static int
uniform_carveout(int device)
{
#ifndef DSVCU_EMU
    static int done[64];
#ifndef DSVCU_CARVEOUT_PCT
#define DSVCU_CARVEOUT_PCT 58
#endif
    static const int pct = DSVCU_CARVEOUT_PCT;
    if (device < 0 || device >= 64 || done[device] || pct < 0) return 0;
#define CARVE(k) CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct))
    CARVE(k_compact_count);
    CARVE(k_compact_scan);
    CARVE(k_compact_scatter);
    CARVE(k_dequant_hf);
    CARVE(k_dequant_ll);
    CARVE(k_ds2x);
    CARVE(k_extend);
    CARVE(k_filter_skew);
    CARVE(k_intra_analysis);
    CARVE(k_me_global);
    CARVE(k_me_level);
    CARVE(k_me_prepass);
    CARVE(k_me_subpel);
    CARVE(k_post_sharpen);
    CARVE(k_predict);
    CARVE(k_quant_hf);
    CARVE(k_quant_hf_edge);
    CARVE(k_quant_ll);
    CARVE(k_reconstruct);
    CARVE(k_sbt_fwd);
    CARVE(k_sbt_inv);
    CARVE(k_luma_avg);
    CARVE(k_pyr_interior);
    CARVE(k_pyr_borders);
#undef CARVE
    done[device] = 1;
#else
    (void) device;
#endif
    return 0;
}
