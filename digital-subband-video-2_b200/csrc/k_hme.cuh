/*
 * k_hme.cuh -- hierarchical motion estimation, sub-pel refinement, per-block
 * mode decision (skip / no-transmit / intra sub-blocks / EPRM) and the I-frame
 * block analysis.
 *
 * Replaces reference src/hme.c: dsv_hme (:2001-2016), refine_level (:1372-1833)
 * and everything it calls (metrics :97-352, block statistics :492-775, half /
 * quarter-pel interpolation :777-837, intra tests :839-1049, subpixel_ME
 * :1051-1164, candidate lists :1166-1298, refine_best_fpel_cand :1300-1370),
 * global_motion (:1973-1999) and dsv_intra_analysis (:1835-1971); plus the MV
 * helpers of src/dsv.c:324-447 it needs.
 *
 * Parallel structure: pyramid levels are sequential (one launch each).  Inside a
 * level, block (i,j) needs the FINAL vectors of its left, top and top-left
 * neighbours (spatial candidates, the MV predictor inside every rate term, the
 * neighbour difference used by the mode decision), so blocks are scheduled as a
 * wavefront: one warp owns one block row and walks it left to right once the
 * row above has published progress > i.  Inside a block
 * the 32 lanes split every pixel loop (SSE, the 2x2-cell psy metric, block
 * statistics, histograms, interpolation) and combine with warp reductions, so
 * all lanes take the same decisions and the reference's first-wins tie-break
 * order is kept by evaluating candidates in list order.
 */
#ifndef K_HME_CUH
#define K_HME_CUH

#include "dsvcu_rt.h"
#include "k_quant.cuh"

#ifndef ME_WARPS_PER_CTA
#define ME_WARPS_PER_CTA 8
#endif
/* residency targets (CTAs per SM) the register allocation is held to: with many
 * encoder instances on one GPU the register file is what runs out first */
#ifndef ME_MIN_CTAS
#define ME_MIN_CTAS 4
#endif
/* the prepass runs four blocks per warp (groups of ME_PRE_G lanes, k_hme_body.cuh) */
#ifndef ME_PRE_G
#define ME_PRE_G 8
#endif
#ifndef ME_PRE_THREADS
#define ME_PRE_THREADS 128
#endif
#define ME_PRE_GROUPS (ME_PRE_THREADS / ME_PRE_G)
#ifndef ME_PRE_MIN_CTAS
#define ME_PRE_MIN_CTAS 8
#endif
#ifndef ME_POLL_NS_CTA
#define ME_POLL_NS_CTA 128 /* same, for rows whose predecessor is in the same CTA */
#endif
#ifndef ME_POLL_NS
#define ME_POLL_NS 256 /* back-off between polls of the row above (a block takes ~30 us) */
#endif
/* ME_COUNT (host emulation only, diagnostics): how much pixel work is left in the
 * dependent wavefront pass at level 0, i.e. how often the prepass' speculation hit */
#if defined(ME_COUNT) && defined(DSVCU_EMU)
static long g_me_cnt[16];
static int g_me_in_wave = 0;
#define ME_CNT(k) (g_me_in_wave ? (void) g_me_cnt[k]++ : (void) 0)
#else
#define ME_CNT(k) ((void) 0)
#endif
enum { MEC_BLOCKS, MEC_AT_S, MEC_EVAL, MEC_EVAL_MISS, MEC_SUBPEL, MEC_REFSTATS, MEC_MAXSUB, MEC_ERR_INTRA, MEC_BLOCK_AVG };
#define ME_BORDER 32
#define ME_MAXLVL 5
#define SP_SZ 16
#define SP_DIM (SP_SZ + 1)
#define HP_STRIDE (SP_DIM * 2)
#define QP_STRIDE (SP_DIM * 4)
#define MEQ_OFF(fx, fy) (4 * (fx) + (4 * (fy)) * QP_STRIDE)

struct MePlane {
    const uint8_t *data;
    int stride, w, h;
};

struct MeArgs {
    MePlane src[3], ref[3], ogr; /* this level; chroma only used at level 0 */
    dsvcu_mv *mvf;               /* this level's field (zero-initialised) */
    const dsvcu_mv *parent;      /* level + 1 field or NULL */
    const dsvcu_mv *ref_mvf;     /* previous picture's final field or NULL */
    int nxb, nyb, y_w, y_h;
    int level, quant, effort, lossless, skip_thresh;
    int hs, vs, vid_w, vid_h, psyscale;
    const int *gxy; /* global motion from the level above */
    int *acc;       /* [0] nintra [1] ndiff [2] eligible [3] total_err; [6] full-block metric evaluations [7] sub-pel position metrics (all levels) */
    int *progress;
    int *ticket;       /* zeroed per launch: wavefront CTAs draw their logical index from it */
    int nrows;
    int b2sr;          /* (256 * (q*q >> 12) * blk_w * blk_h) / (width * height), dsv.c:370 */
    struct MePre *pre; /* per-block results of k_me_prepass for this level */
};

DSVCU_HD int me_abs(int v) { return v < 0 ? -v : v; }
DSVCU_HD int me_sqr(int v) { return v * v; }
DSVCU_HD int me_avg2(int a, int b) { return (a + b + 1) >> 1; }
DSVCU_HD unsigned me_uavg4(int a, int b, int c, int d) { return (unsigned) (a + b + c + d + 2) >> 2; }
DSVCU_HD int me_u8(int v) { return v > 255 ? 255 : (v < 0 ? 0 : v); }
DSVCU_HD int me_sar_r2(int v) { return (v + 2) >> 2; } /* DSV_SAR_R(v, 2) */

/* ---- packed-byte helpers: four horizontally adjacent pixels per 32-bit word.
 * On the device these map to one instruction each (byte-SIMD absolute
 * difference, 4-way dot product, byte permute); the host emulation spells them
 * out.  Blocks whose width is 4, 8, 16 or 32 take the packed paths below, any
 * other width the generic per-pixel loops. ---- */
#ifndef DSVCU_EMU
DSVCU_DEV uint32_t me_ld4(const uint8_t *p)
{
    /* unaligned 4-byte read from two aligned words; may touch up to 3 bytes
     * past p+3, which always lie inside the frame allocation */
    uintptr_t a = (uintptr_t) p;
    const uint32_t *q = (const uint32_t *) (a & ~(uintptr_t) 3);
    return __funnelshift_r(q[0], q[1], (unsigned) (a & 3) * 8);
}
/* the same when p is known to be 4-byte aligned */
DSVCU_DEV uint32_t me_ld4a(const uint8_t *p) { return *(const uint32_t *) p; }
DSVCU_DEV uint32_t me_absdiff4(uint32_t a, uint32_t b) { return __vabsdiffu4(a, b); }
DSVCU_DEV unsigned me_dot4(uint32_t a, uint32_t b, unsigned acc) { return __dp4a(a, b, acc); }
DSVCU_DEV uint32_t me_perm(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
#else
DSVCU_DEV uint32_t me_ld4(const uint8_t *p)
{
    return (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24);
}
DSVCU_DEV uint32_t me_ld4a(const uint8_t *p) { return me_ld4(p); }
DSVCU_DEV uint32_t me_absdiff4(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) {
        int x = (int) ((a >> (8 * k)) & 255) - (int) ((b >> (8 * k)) & 255);
        r |= (uint32_t) (x < 0 ? -x : x) << (8 * k);
    }
    return r;
}
DSVCU_DEV unsigned me_dot4(uint32_t a, uint32_t b, unsigned acc)
{
    for (int k = 0; k < 4; k++) acc += ((a >> (8 * k)) & 255) * ((b >> (8 * k)) & 255);
    return acc;
}
DSVCU_DEV uint32_t me_perm(uint32_t a, uint32_t b, uint32_t sel)
{
    uint64_t v = ((uint64_t) b << 32) | a;
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) r |= (uint32_t) ((v >> (8 * ((sel >> (4 * k)) & 7))) & 255) << (8 * k);
    return r;
}
#endif
#define ME_ONES 0x01010101u

/* log2(w / 4) for w in {4, 8, 16, 32}, else -1 (generic path) */
DSVCU_DEV int me_gshift(int w)
{
    return w == 16 ? 2 : (w == 8 ? 1 : (w == 32 ? 3 : (w == 4 ? 0 : -1)));
}

/* index of the lowest set bit (x != 0) */
DSVCU_DEV int
me_ctz(unsigned x)
{
#ifndef DSVCU_EMU
    return __ffs((int) x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

/* floor(sqrt(n)): what the reference's digit-by-digit iisqrt (hme.c:99-124)
 * returns; here from the hardware square root plus an exact integer fix-up
 * (checked against the digit-by-digit form over the full 32-bit range at
 * every perfect square +-2 and a dense sample, see tests) */
DSVCU_HD unsigned
me_isqrt(unsigned n)
{
    unsigned r = (unsigned) sqrtf((float) n);
    while ((unsigned long long) r * r > n) r--;
    while ((unsigned long long) (r + 1) * (r + 1) <= n) r++;
    return r;
}

#define ME_HPF(a, b, c, d) ((5 * ((b) + (c))) - ((a) + (d)))
#define ME_WIN (SP_DIM + 3) /* full-pel window rows/cols -1 .. SP_DIM+1 */
#define ME_MAXSP 7
#define ME_MEMO 32
#define ME_PRE_WORDS 256

/* ---- neighbour-independent part of a block, computed for every block of a
 * level in parallel by k_me_prepass and consumed by the wavefront.
 *
 * Besides what is neighbour-independent by construction (source statistics,
 * the non-spatial candidates and their metrics) the record carries a
 * SPECULATION: the prepass guesses the block's final full-pel vector S (the
 * best measured candidate by raw metric, followed through the reference's
 * descent with an assumed predictor) and computes everything the mode decision
 * needs AT S -- second sub-pel pass, reference-side statistics, the sub-block
 * metrics of the no-transmit test, the reference-side half of the intra test.
 * The wavefront uses those numbers when its real decision lands on S and falls
 * back to computing them on demand when it does not, so the speculation only
 * moves work out of the dependency chain; results are identical either way. ---- */
#define ME_PRE_NB 20 /* temporal (<= 9) + global + parent inliers (<= 9) */
#define ME_PRE_NM 32
#define ME_SV_RS 1    /* reference-side statistics at S */
#define ME_SV_BSUB 2  /* sub-block metrics at S (no-transmit test) */
#define ME_SV_ZSUB 4  /* sub-block metrics at the zero vector (skip test) */
#define ME_SV_INTRA 8 /* reference quadrant means at S (+ per-quadrant error sums where the gate passes) */
struct __align__(16) MePre {
    unsigned var_src, avg_src;
    int motion_bias, psy_pack; /* err_w | tex_w << 8 | avg_w << 16 */
    int lax, lay, has_list, nb;
    unsigned zoscore;
    int uavg, vavg, nm;
    short bx[ME_PRE_NB], by[ME_PRE_NB]; /* non-spatial candidates after (lax, lay), raw units */
    short mx[ME_PRE_NM], my[ME_PRE_NM]; /* positions already measured ... */
    unsigned mv[ME_PRE_NM];             /* ... and their raw metric */
    /* level 0 only from here on */
    int sx, sy, s_valid;                /* speculated full-pel vector, ME_SV_* bits */
    unsigned rs_ogrerr, rs_var, rs_avg; /* statistics against the reference at S */
    int rs_u, rs_v, rs_eprm;            /* eprm: bit0 i, bit1 d, bit2 r */
    unsigned bsub[3], zsub[3];          /* raw max-sub-block metrics at S / at zero */
    unsigned q_detail[4], q_avg[4];     /* source-side quadrant statistics (intra test) */
    unsigned qa_sub[4];                 /* mean of the reference quadrants at S */
    int qi_mask;                        /* quadrants whose error sums below are valid (ratio == 32) */
    unsigned qi_sub[4], qi_src[4], qi_inter[4];
    /* the inter error of a quadrant depends on the sub-pel gain `ratio` cell by cell
     * (hme.c:839-889): ratio-independent part + the 16 per-cell mean absolute errors
     * of each 8x8 quadrant let the wavefront rebuild it for any ratio */
    int qi_cells;                       /* 1: qi_rest / qi_ae are filled (16x16 blocks) */
    unsigned qi_rest[4];
    uint8_t qi_ae[64];
    int utex, vtex;
    /* sub-pel measurements: [0] around the parent average (lax, lay), [1] around S */
    int sp_valid[2], sp_nv[2];
    signed char sp_tx[2][ME_MAXSP + 1], sp_ty[2][ME_MAXSP + 1];
    unsigned sp_sc[2][ME_MAXSP];
    int pad_[3]; /* sizeof(MePre) is a multiple of 16: the wavefront fetches a record in 16-byte words */
};

static_assert(sizeof(MePre) % 16 == 0 && sizeof(MePre) <= ME_PRE_WORDS * 4, "MePre: whole 16-byte words, must fit the staging area");

#define ME_MAXCAND 40

/* ---- the per-block functions, for one warp per block and for four blocks per warp ---- */
#define ME_G 32
namespace meg32 {
#include "k_hme_body.cuh"
}
#undef ME_G
#define ME_G ME_PRE_G
namespace meg8 {
#include "k_hme_body.cuh"
}
#undef ME_G

/* lane / group bookkeeping of the kernels below (the host emulation runs one "lane" per CTA) */
#ifdef DSVCU_EMU
#define ME_KLANE 0
#define ME_WIC 0
#define ME_WARP ((int) blockIdx.x)
#define ME_NWARPS ((int) gridDim.x)
#define ME_GRP(g) ((int) blockIdx.x)
#define ME_NGRPS(g) ((int) gridDim.x)
#define ME_GIC(g) 0
#else
#define ME_KLANE ((int) (threadIdx.x & 31))
#define ME_WIC ((int) (threadIdx.x >> 5))
#define ME_WARP ((int) ((blockIdx.x * blockDim.x + threadIdx.x) >> 5))
#define ME_NWARPS ((int) ((gridDim.x * blockDim.x) >> 5))
#define ME_GRP(g) ((int) ((blockIdx.x * blockDim.x + threadIdx.x) / (g)))
#define ME_NGRPS(g) ((int) ((gridDim.x * blockDim.x) / (g)))
#define ME_GIC(g) ((int) (threadIdx.x / (g)))
#endif

/* neighbour-independent half of every block of a level: ME_PRE_G lanes per block,
 * four blocks (consecutive in a block row) per warp */
DSVCU_KERNEL void __launch_bounds__(ME_PRE_THREADS, ME_PRE_MIN_CTAS)
k_me_prepass(MeArgs A)
{
    DSVCU_SHARED meg8::MeScratch scratch[ME_PRE_GROUPS];
    meg8::MeScratch *S = &scratch[ME_GIC(ME_PRE_G)];
    const int step = 1 << A.level;
    if (ME_KLANE % ME_PRE_G == 0) S->ip = NULL; /* no sub-pel measurement in here: k_me_subpel */
#ifndef DSVCU_EMU
    __syncwarp();
#endif
    const int cols = (A.nxb + step - 1) / step, rows = (A.nyb + step - 1) / step;
    for (int b = ME_GRP(ME_PRE_G); b < cols * rows; b += ME_NGRPS(ME_PRE_G)) {
        int r = b / cols, c = b - r * cols;
        meg8::me_prepass_block(A, S, c * step, r * step);
    }
}

/* The sub-pel measurements of level 0 (half-pel image of a 17 x 17 window + the metric at
 * up to seven quarter-pel offsets, hme.c:1051-1164), two per block: around the parent
 * average and around the speculated winner.  One warp per measurement: 289 interpolated
 * samples and 7 x 32 metric work items keep all 32 lanes busy, which the 8-lane groups of
 * the prepass could not -- and without the interpolation scratch the prepass fits three
 * times as many blocks per SM. */
#define ME_SP_WARPS 8
DSVCU_KERNEL void __launch_bounds__(ME_SP_WARPS * 32)
k_me_subpel(MeArgs A)
{
    DSVCU_SHARED meg32::MeInterp interp[ME_SP_WARPS];
    DSVCU_SHARED meg32::MeScratch scratch[ME_SP_WARPS];
    meg32::MeScratch *S = &scratch[ME_WIC];
    const int ntask = 2 * A.nxb * A.nyb;
    if (ME_KLANE == 0) S->ip = &interp[ME_WIC];
#ifndef DSVCU_EMU
    __syncwarp();
#endif
    for (int t = ME_WARP; t < ntask; t += ME_NWARPS) {
        const int b = t >> 1;
        meg32::me_subpel_task(A, S, b % A.nxb, b / A.nxb, t & 1);
    }
}

/* Wavefront over the block rows of one pyramid level.  Block (i, j) needs the
 * final vectors of its left, top and top-left neighbours, so row
 * r may work on block c once row r-1 has finished block c.
 *
 * Since the prepass took the pixel work out of this pass, a block is mostly
 * scalar decisions, and what limits the whole encoder is how many warps (and
 * registers) the wavefronts of all the instances on the GPU keep resident while
 * they wait for each other.  So a warp owns ME_LVL_RPW = 4 consecutive rows, one
 * group of ME_LVL_G = 8 lanes each, in SKEWED LOCKSTEP: at warp-step t the group
 * of row g works on column t - g, which is exactly the dependency between the
 * rows of a warp -- no flags inside a warp, one __syncwarp per step, and the four
 * groups mostly walk the same instructions together.  Only the first row of a warp
 * waits for a progress word: shared memory between the warps of a CTA, global
 * memory (+ device-scope fence) between CTAs.  The vector field itself is read
 * with volatile loads (me_ldmv), i.e. from L2.  1080p level 0: 68 rows = 17 warps
 * in 5 CTAs, against 68 warps in 9 CTAs for one warp per row. */
#ifndef ME_LVL_G
#define ME_LVL_G 32 /* lanes per block row in the wavefront: 32 (one row per warp) or ME_PRE_G (four rows per warp) */
#endif
#if ME_LVL_G == 32
#define ME_LVL_NS meg32
#else
#define ME_LVL_NS meg8
#endif
#define ME_LVL_RPW (32 / ME_LVL_G)
#ifndef ME_LVL_WARPS
#define ME_LVL_WARPS (ME_LVL_G == 32 ? 8 : 4)
#endif
#define ME_LVL_ROWS (ME_LVL_WARPS * ME_LVL_RPW) /* rows per CTA */
struct MeLvlShared {
    ME_LVL_NS::MeInterp interp[ME_LVL_ROWS];
    ME_LVL_NS::MeScratch scratch[ME_LVL_ROWS];
    __align__(16) uint32_t pre_words[ME_LVL_ROWS][ME_PRE_WORDS];
    int sprog[ME_LVL_WARPS];
    int cta;
};

DSVCU_KERNEL void __launch_bounds__(ME_LVL_WARPS * 32, ME_LVL_G == 32 ? 2 : 4)
k_me_level(MeArgs A)
{
    DSVCU_DYN_SMEM(MeLvlShared, sh);
    const int step = 1 << A.level;
    int acc_local[4] = { 0, 0, 0, 0 };
#ifndef DSVCU_EMU
    /* CTA k waits for CTA k - 1; k is a ticket drawn at entry, not blockIdx, so a waiting CTA only
     * waits for CTAs that have already started, whatever the dispatch order */
    if (threadIdx.x == 0) sh->cta = atomicAdd(A.ticket, 1);
    __syncthreads();
    const int cta = sh->cta;
    const int wic = ME_WIC, g = ME_KLANE / ME_LVL_G;   /* warp in CTA, row group in warp */
    const int lr = wic * ME_LVL_RPW + g;                /* row in CTA */
    const int row = cta * ME_LVL_ROWS + lr;
    const int cols = (A.nxb + step - 1) / step;
    ME_LVL_NS::MeScratch *S = &sh->scratch[lr];
    if ((ME_KLANE % ME_LVL_G) == 0) S->ip = &sh->interp[lr];
    if (threadIdx.x < ME_LVL_WARPS) sh->sprog[threadIdx.x] = 0;
    __syncthreads();
    {
        /* the row above the warp's first row: previous warp of the CTA, or the last
         * warp of the previous CTA (its progress word is published globally) */
        const int wrow0 = cta * ME_LVL_ROWS + wic * ME_LVL_RPW;
        const bool above_global = (wic == 0), pub_global = (wic == ME_LVL_WARPS - 1);
        volatile const int *above = above_global ? (volatile const int *) (A.progress + cta - 1)
                                                 : (volatile const int *) (sh->sprog + wic - 1);
        int seen = (wrow0 == 0) ? 0x7fffffff : 0;
        if (wrow0 < A.nrows) {
            for (int t = 0; t < cols + ME_LVL_RPW - 1; t++) {
                const int col = t - g;
                const bool active = row < A.nrows && col >= 0 && col < cols;
                /* the warp's first row is the only one that depends on another warp */
                if (g == 0 && active && seen < col + 1) {
                    while ((seen = *above) < col + 1) {
                        __nanosleep(above_global ? ME_POLL_NS : ME_POLL_NS_CTA);
                    }
                    if (above_global) {
                        __threadfence();
                    } else {
                        __threadfence_block();
                    }
                }
                __syncwarp();
                if (active) ME_LVL_NS::me_block(A, S, sh->pre_words[lr], col * step, row * step, acc_local);
                /* vectors of this step must be visible before the next step's readers
                 * (the other groups of the warp, then the next warp / CTA) */
                if (pub_global) {
                    __threadfence();
                } else {
                    __threadfence_block();
                }
                __syncwarp();
                if (g == ME_LVL_RPW - 1 && active && (ME_KLANE % ME_LVL_G) == 0) {
                    *(volatile int *) (sh->sprog + wic) = col + 1;
                    if (pub_global) *(volatile int *) (A.progress + cta) = col + 1;
                }
            }
        }
    }
    if ((ME_KLANE % ME_LVL_G) == 0 && A.level == 0) {
        for (int k = 0; k < 4; k++) {
            if (acc_local[k]) atomicAdd(&A.acc[k], acc_local[k]);
        }
    }
#else
    ME_LVL_NS::MeScratch *S = &sh->scratch[0];
    S->ip = &sh->interp[0];
    for (int lr = 0; lr < ME_LVL_ROWS; lr++) {
        int row = (int) blockIdx.x * ME_LVL_ROWS + lr;
        if (row >= A.nrows) continue;
        for (int i = 0; i < A.nxb; i += step) {
#if defined(ME_COUNT)
            g_me_in_wave = (A.level == 0);
#endif
            ME_LVL_NS::me_block(A, S, sh->pre_words[0], i, row * step, acc_local);
#if defined(ME_COUNT)
            g_me_in_wave = 0;
#endif
        }
    }
    if (A.level == 0) {
        for (int k = 0; k < 4; k++) {
            if (acc_local[k]) atomicAdd(&A.acc[k], acc_local[k]);
        }
    }
#endif
}

/* global_motion (hme.c:1973-1999): average vector of a level, x2 */
DSVCU_KERNEL void __launch_bounds__(256)
k_me_global(const dsvcu_mv *vecs, int nxb, int nyb, int level, int *gxy)
{
    DSVCU_SHARED int sx, sy;
    int step = 1 << level;
    int cols = (nxb + step - 1) / step, rows = (nyb + step - 1) / step, n = cols * rows;
    int ax = 0, ay = 0;
    if (DSVCU_TID == 0) {
        sx = 0;
        sy = 0;
    }
    DSVCU_SYNC();
    PAR_FOR(k, n) {
        int r = k / cols, c = k - r * cols;
        const dsvcu_mv *m = vecs + c * step + r * step * nxb;
        ax += m->x;
        ay += m->y;
    }
    atomicAdd(&sx, ax);
    atomicAdd(&sy, ay);
    DSVCU_SYNC();
    if (DSVCU_TID == 0) {
        gxy[0] = n ? sx * 2 / n : 0;
        gxy[1] = n ? sy * 2 / n : 0;
    }
}

/* ---- dsv_intra_analysis (hme.c:1835-1971): one warp per block ---- */

struct IaArgs {
    MePlane src[3];
    dsvcu_mv *out;
    int nxb, nyb, y_w, y_h, hs, vs, do_psy, scale;
};

DSVCU_KERNEL void __launch_bounds__(ME_WARPS_PER_CTA * 32)
k_intra_analysis(IaArgs A)
{
    using namespace meg32;
    DSVCU_SHARED int hists[ME_WARPS_PER_CTA][16];
    int *hist = hists[ME_WIC];
    const int total = A.nxb * A.nyb;
    for (int b = ME_WARP; b < total; b += ME_NWARPS) {
        int j = b / A.nxb, i = b - j * A.nxb;
        int bx = i * A.y_w, by = j * A.y_h;
        const MePlane &sp = A.src[0];
        unsigned flags = 0;
        if (!(bx >= sp.w || by >= sp.h)) {
            const uint8_t *srcd = sp.data + by * sp.stride + bx;
            int bw = min(sp.w - bx, A.y_w), bh = min(sp.h - by, A.y_h);
            int cbx = i * (A.y_w >> A.hs), cby = j * (A.y_h >> A.vs), cbw = bw >> A.hs, cbh = bh >> A.vs;
            unsigned luma_detail, luma_avg, var_t;
            int maintain = 1, keep_hf = 1, npeaks = 0, foliage = 0, is_text = 0, ringing = 0;
            MeChroma cpsy;
            luma_detail = (unsigned) me_block_detail(srcd, sp.stride, bw, bh, &luma_avg);
            if (A.do_psy & (16 | 2)) {
                int skip_tones, uavg, vavg, tf = 0, tf2 = 0, qtex, hvar, luma_var, luma_tex;
                hvar = (int) me_block_hist_var(srcd, sp.stride, bw, bh, hist);
                qtex = me_quant_tex(srcd, sp.stride, bw, bh);
                luma_var = me_block_var(srcd, sp.stride, bw, bh, &luma_avg);
                luma_var /= (bw * bh);
                luma_tex = (int) me_block_tex(srcd, sp.stride, bw, bh);
                luma_tex /= (bw * bh);
                npeaks = me_block_peaks(srcd, sp.stride, bw, bh, hist, (int) luma_avg);
                is_text = (me_abs(npeaks - 2) <= 1);
                if (qtex == 1 || qtex == 2) tf2 = hvar <= 3 && (luma_tex >= 10 && luma_var >= luma_tex);
                if (qtex == 2 || qtex == 3) {
                    tf = luma_tex >= 8 && luma_var >= (2 * luma_tex);
                    tf &= (me_abs(hvar - 5) <= 3);
                }
                is_text &= (tf || tf2);
                me_c_average(A.src, cbx, cby, cbw, cbh, &uavg, &vavg);
                me_chroma_analysis(&cpsy, (int) luma_avg, uavg, vavg);
                foliage = (cpsy.nature && luma_avg < 160);
                foliage &= luma_detail > (unsigned) ((36 * bw * bh) / max(A.scale, 1));
                if (foliage) is_text = 0;
                skip_tones = cpsy.hifreq;
                if ((A.do_psy & 16) && !skip_tones && (foliage || (hvar <= (min(qtex - 3, 2) * 16) && qtex > 1))) {
                    ringing = 1;
                }
                var_t = 8;
                if (cpsy.nature || cpsy.greyish || cpsy.skinnish) {
                    var_t += 12;
                } else if (!cpsy.hifreq) {
                    var_t += 8;
                }
            } else {
                var_t = 16;
            }
            if (A.do_psy & (2 | 1)) {
                luma_detail /= (unsigned) (bw * bh);
                keep_hf &= luma_detail < 48;
                maintain = (luma_detail < var_t * 4);
            }
            if (A.do_psy & 2) {
                if (foliage) {
                    keep_hf = 0;
                    maintain = 1;
                } else if (is_text) {
                    keep_hf = 1;
                    maintain = 0;
                }
            }
            if ((A.do_psy & 16) && luma_avg < 24) ringing = 1;
            if (ringing) flags |= MVF_RINGING;
            if (maintain) flags |= MVF_MAINTAIN;
            if (keep_hf) flags |= MVF_SKIP;
        }
        if (ME_KLANE == 0) {
            dsvcu_mv *o = A.out + b;
            o->x = 0;
            o->y = 0;
            o->flags = flags;
            o->err = 0;
            o->dc = 0;
            o->submask = 0;
        }
    }
}

#endif /* K_HME_CUH */
