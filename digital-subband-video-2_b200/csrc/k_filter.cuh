/*
 * k_filter.cuh -- in-loop luma/chroma filters, intra deringing filter,
 * de-gradient sharpening.
 *
 * Replaces reference src/bmc.c: ihfilter4x4/ivfilter4x4 (:70-191), dsff4x4,
 * haar4x4, artf4x4 (:194-270), degrad4x4 (:276-337), dsv_post_process
 * (:340-361), curve_tex/compute_filter_q (:364-388), dsv_intra_filter
 * (:390-457), luma_filter (:459-602), chroma_filter (:604-659).
 *
 * The reference filters 4x4 cells in place in raster order; cell (p,q) reads
 * pixels already modified by (p-1,q) and (p+1,q-1), so the
 * legal parallel schedule is a wavefront with slope 2: cell p of row q may run
 * once row q-1 has completed cell p+1.
 *
 * Schedule ("skewed lockstep"): a group of FILT_LPC lanes owns one row of
 * cells; a warp therefore owns a band of FILT_G = 32 / FILT_LPC consecutive
 * rows, and at warp-step t group g works on column t - 2g.  Inside a band the
 * slope-2 dependency is satisfied by construction -- the warp executes in
 * lockstep, one __syncwarp per step, no flags, no polling.  Band b trails band
 * b-1 by 2*FILT_G steps; that single hand-off per step goes through a step
 * counter: in shared memory between the FILT_WPC bands of one CTA, in global
 * memory (with device-scope fences, and L2 loads for the two cell rows that
 * read what the neighbouring SM wrote) between CTAs.  CTAs are small (256
 * threads, 16 cell rows) so they fit beside whatever else is resident; the
 * three planes of a picture are filtered by one launch.  With eight or more
 * lanes per cell the two edges of every filtered line go to different lanes.
 */
#ifndef K_FILTER_CUH
#define K_FILTER_CUH

#include "dsvcu_rt.h"
#include "k_quant.cuh"

#define FILT_MODE_LUMA 0
#define FILT_MODE_INTRA 1
#define FILT_MODE_CHROMA 2

#ifndef FILT_LPC
#define FILT_LPC 16                       /* lanes per cell row (measured 1080p P picture, reconstruct + filters:
                                           * 4 lanes 1.89 ms, 8 lanes 1.46 ms, 16 lanes 1.17 ms -- fewer cells per
                                           * warp means less serialised divergence per step) */
#endif
#define FILT_G (32 / FILT_LPC)            /* cell rows per warp (band height) */
#ifndef FILT_WPC
#define FILT_WPC 8                        /* bands (warps) per CTA */
#endif
#define FILT_CELLS (FILT_WPC * FILT_G)    /* cell rows per CTA */
#define FILT_THREADS (FILT_WPC * 32)

struct FiltArgs {
    uint8_t *data;
    int stride, w, h;
    const dsvcu_mv *mvs;
    const uint8_t *blockdata;
    int nbh, nbv, blk_w, blk_h;
    int q;        /* luma/intra: compute_filter_q(); chroma: raw quant */
    int fthresh;
    int do_filter;
    int sharpen;
    int bw, bh;   /* chroma: block size in this plane */
    int ncols, nrows;
    int mode;
    int *progress; /* completed steps per band (global; zeroed before the launch) */
    int first_cta; /* this plane's first CTA in the grid */
};

struct FiltJob {
    FiltArgs p[3];
    int nplanes;
    int *ticket; /* zeroed before the launch: CTAs take their logical index from it (see k_filter_skew) */
};

#ifdef DSVCU_EMU
#define FILT_SUB 0
#define FILT_NSUB 1
#define F_GSYNC() ((void) 0)
#else
#define FILT_SUB ((int) (threadIdx.x & (FILT_LPC - 1)))
#define FILT_NSUB FILT_LPC
/* barrier + memory ordering among the lanes that share a cell */
#if FILT_LPC >= 32
#define F_GROUP_MASK 0xffffffffu
#else
#define F_GROUP_MASK ((((1u << FILT_LPC) - 1u)) << ((threadIdx.x & 31u) & ~(unsigned) (FILT_LPC - 1)))
#endif
#define F_GSYNC() __syncwarp(F_GROUP_MASK)
#endif
#define PXLD(p) (*(p))

DSVCU_HD int f_abs(int v) { return v < 0 ? -v : v; }
DSVCU_HD int f_clamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

#define F_LPF(e0, i0, e1, i1) ((5 * ((e0) + (i0)) + 3 * ((e1) + (i1)) + 8) >> 4)
#define F_TEST(t, avg, e0, e1, e2, i0, i1, i2)                                         \
    (f_abs((e0) - (avg)) < (t) && f_abs((i0) - (avg)) < (t) && f_abs((e1) - (avg)) < (t) && \
     f_abs((i1) - (avg)) < (t) && f_abs((e2) - (avg)) < (t) && f_abs((i2) - (avg)) < (t))

/* one line of the edge filter: 11 samples p[-3..7] with pitch d (bmc.c:86-127) */
template <bool VOL>
DSVCU_DEV int
f_px(const uint8_t *p)
{
#ifndef DSVCU_EMU
    if (VOL) return *(volatile const uint8_t *) p; /* written by another SM: read it from L2 */
#endif
    return *p;
}

template <bool VOL>
DSVCU_DEV void
f_edge_line_t(uint8_t *p, int d, int tE, int tM, int in_edge)
{
#undef PXLD
#define PXLD(q) f_px<VOL>(q)
    int e2 = PXLD(p - 3 * d), e1 = PXLD(p - 2 * d), e0 = PXLD(p - d);
    int i0 = PXLD(p), i1 = PXLD(p + d), i2 = PXLD(p + 2 * d);
    int n1 = 0, n0 = 0, m0 = 0, m1 = 0, m2 = 0;
    if (in_edge) {
        n1 = PXLD(p + 3 * d);
        n0 = PXLD(p + 4 * d);
        m0 = PXLD(p + 5 * d);
        m1 = PXLD(p + 6 * d);
        m2 = PXLD(p + 7 * d);
    }
    int avg = F_LPF(e0, i0, e1, i1);
    if (F_TEST(tE, avg, e0, e1, e2, i0, i1, i2)) {
        p[-2 * d] = (uint8_t) ((3 * (avg + e1) + 2 * e2 + 4) >> 3);
        p[0] = (uint8_t) avg;
        avg *= 5;
        p[-d] = (uint8_t) ((avg + 2 * e1 + e2 + 4) >> 3);
        p[d] = (uint8_t) ((avg + 2 * i1 + i2 + 4) >> 3);
    }
    if (in_edge) {
        /* second edge at +4: inner side is (n1, i2'), outer side m0..m2 where
         * i2 = p[2d] and n1 = p[3d] were not touched above */
        int ii2 = i2, ii1 = n1, ii0 = n0, ee0 = m0, ee1 = m1, ee2 = m2;
        avg = F_LPF(ee0, ii0, ee1, ii1);
        if (F_TEST(tM, avg, ee0, ee1, ee2, ii0, ii1, ii2)) {
            p[4 * d] = (uint8_t) avg;
            p[6 * d] = (uint8_t) ((3 * (avg + ee1) + 2 * ee2 + 4) >> 3);
            avg *= 5;
            p[3 * d] = (uint8_t) ((avg + 2 * ii1 + ii2 + 4) >> 3);
            p[5 * d] = (uint8_t) ((avg + 2 * ee1 + ee2 + 4) >> 3);
        }
    }
}

#undef PXLD
#define PXLD(p) (*(p))

DSVCU_DEV void
f_edge_line(uint8_t *p, int d, int tE, int tM, int in_edge)
{
    f_edge_line_t<false>(p, d, tE, tM, in_edge);
}

/* The two edges of a line are independent: the first reads p[-3d..2d] and
 * writes p[-2d..d], the second (at +4) reads p[2d..7d] and writes p[3d..6d].
 * With eight or more lanes per cell they go to different lanes. */
DSVCU_DEV void
f_edge_first(uint8_t *p, int d, int tE)
{
    int e2 = p[-3 * d], e1 = p[-2 * d], e0 = p[-d], i0 = p[0], i1 = p[d], i2 = p[2 * d];
    int avg = F_LPF(e0, i0, e1, i1);
    if (F_TEST(tE, avg, e0, e1, e2, i0, i1, i2)) {
        p[-2 * d] = (uint8_t) ((3 * (avg + e1) + 2 * e2 + 4) >> 3);
        p[0] = (uint8_t) avg;
        avg *= 5;
        p[-d] = (uint8_t) ((avg + 2 * e1 + e2 + 4) >> 3);
        p[d] = (uint8_t) ((avg + 2 * i1 + i2 + 4) >> 3);
    }
}

DSVCU_DEV void
f_edge_second(uint8_t *p, int d, int tM)
{
    int ii2 = p[2 * d], ii1 = p[3 * d], ii0 = p[4 * d], ee0 = p[5 * d], ee1 = p[6 * d], ee2 = p[7 * d];
    int avg = F_LPF(ee0, ii0, ee1, ii1);
    if (F_TEST(tM, avg, ee0, ee1, ee2, ii0, ii1, ii2)) {
        p[4 * d] = (uint8_t) avg;
        p[6 * d] = (uint8_t) ((3 * (avg + ee1) + 2 * ee2 + 4) >> 3);
        avg *= 5;
        p[3 * d] = (uint8_t) ((avg + 2 * ii1 + ii2 + 4) >> 3);
        p[5 * d] = (uint8_t) ((avg + 2 * ee1 + ee2 + 4) >> 3);
    }
}

/* line k (0..3) of a 4-line pass, both edges: one lane per (line, edge) when the
 * cell has at least eight lanes, one lane per line otherwise */
DSVCU_DEV void
f_edge_lanes(uint8_t *p0, ptrdiff_t line_pitch, int nlines, int d, int tE, int tM, int in_edge)
{
#if !defined(DSVCU_EMU) && FILT_LPC >= 8
    const int k = FILT_SUB & 3, half = FILT_SUB >> 2;
    if (k < nlines && half < 2) {
        uint8_t *p = p0 + k * line_pitch;
        if (half == 0) {
            f_edge_first(p, d, tE);
        } else if (in_edge) {
            f_edge_second(p, d, tM);
        }
    }
#else
    for (int k = FILT_SUB; k < nlines; k += FILT_NSUB) {
        f_edge_line(p0 + k * line_pitch, d, tE, tM, in_edge);
    }
#endif
}

/* ihfilter4x4 (bmc.c:70-128): lanes split the rows */
DSVCU_DEV void
f_hfilter(const FiltArgs &A, int x, int y, int edge, int tE, int tM)
{
    if (x < 4 || x > A.w - 4 || (edge && tE <= 0) || tM <= 0) return;
    int top = f_clamp(y, 0, A.h - 1), bot = f_clamp(y + 4, 0, A.h - 1);
    int in_edge = x < (A.w - 8);
    if (!edge) tE = tM;
    f_edge_lanes(A.data + (ptrdiff_t) top * A.stride + x, A.stride, bot - top, 1, tE, tM, in_edge);
}

/* ivfilter4x4 (bmc.c:130-191): lanes split the columns */
DSVCU_DEV void
f_vfilter(const FiltArgs &A, int x, int y, int edge, int tE, int tM)
{
    if (y < 4 || y > A.h - 4 || (edge && tE <= 0) || tM <= 0) return;
    int beg = f_clamp(x, 0, A.w - 1), end = f_clamp(x + 4, 0, A.w - 1);
    int in_edge = y < (A.h - 8);
    if (!edge) tE = tM;
    f_edge_lanes(A.data + (ptrdiff_t) y * A.stride + beg, 1, end - beg, A.stride, tE, tM, in_edge);
}

struct F4x4 {
    int p[16];
};

DSVCU_DEV void
f_load4x4(F4x4 &b, const uint8_t *a, int as)
{
    for (int r = 0; r < 4; r++) {
        for (int c = 0; c < 4; c++) {
            b.p[r * 4 + c] = PXLD(a + r * as + c);
        }
    }
}

/* artf4x4 + haar4x4 (bmc.c:227-270) on a loaded cell */
DSVCU_DEV void
f_artf(const F4x4 &b, int *psh, int *psv, int *pslh, int *pslv)
{
    int sh = 0, sv = 0;
    for (int y = 0; y < 4; y += 2) {
        for (int x = 0; x < 4; x += 2) {
            int x0 = b.p[y * 4 + x], x1 = b.p[y * 4 + x + 1];
            int x2 = b.p[(y + 1) * 4 + x], x3 = b.p[(y + 1) * 4 + x + 1];
            int HH = f_abs(x0 - x1 - x2 + x3) >> 1;
            sh += f_abs(x0 - x1 + x2 - x3) + HH;
            sv += f_abs(x0 + x1 - x2 - x3) + HH;
        }
    }
    int d0 = (b.p[0] + b.p[1] + b.p[4] + b.p[5] + 2) >> 2;
    int d1 = (b.p[2] + b.p[3] + b.p[6] + b.p[7] + 2) >> 2;
    int d2 = (b.p[8] + b.p[9] + b.p[12] + b.p[13] + 2) >> 2;
    int d3 = (b.p[10] + b.p[11] + b.p[14] + b.p[15] + 2) >> 2;
    int HH = f_abs(d0 - d1 - d2 + d3) >> 1;
    *psh = sh;
    *psv = sv;
    *pslh = f_abs(d0 - d1 + d2 - d3) + HH;
    *pslv = f_abs(d0 + d1 - d2 - d3) + HH;
}

/* dsff4x4 (bmc.c:194-225) */
DSVCU_DEV int
f_dsff(const F4x4 &b)
{
    int d0 = (b.p[0] + b.p[1] + b.p[4] + b.p[5] + 2) >> 2;
    int d1 = (b.p[2] + b.p[3] + b.p[6] + b.p[7] + 2) >> 2;
    int d2 = (b.p[8] + b.p[9] + b.p[12] + b.p[13] + 2) >> 2;
    int d3 = (b.p[10] + b.p[11] + b.p[14] + b.p[15] + 2) >> 2;
    int sh = f_abs((d0 + d1) - (d3 + d2));
    int sv = f_abs((d2 + d1) - (d3 + d0));
    if (max(sh, sv) < 8) return 0;
    d2 = 255 - d2;
    d3 = 255 - d3;
    sh = f_abs(d0 - d1 + d2 - d3);
    sv = f_abs(d0 + d1 - d2 - d3) >> 2;
    if (sh > sv) return (3 * sh + sv + 2) >> 2;
    return (3 * sv + sh + 2) >> 2;
}

/* degrad4x4 (bmc.c:276-337); executed by one lane */
DSVCU_DEV void
f_degrad(uint8_t *a, int as)
{
    int hist[16], avgs[16];
    F4x4 b;
    int lo = -1, hi = -1;
    for (int i = 0; i < 16; i++) {
        hist[i] = 0;
        avgs[i] = 0;
    }
    f_load4x4(b, a, as);
    for (int i = 0; i < 16; i++) {
        int t = b.p[i] >> 4;
        hist[t]++;
        avgs[t] += b.p[i];
    }
    for (int i = 0; i < 16; i++) {
        if (hist[i]) {
            if (lo == -1) lo = i;
            hi = i;
        }
    }
    if (lo >= hi) return;
    int alo = avgs[lo] / hist[lo], ahi = avgs[hi] / hist[hi];
    if (alo == 0) alo = 1;
    if (ahi == 0) ahi = 1;
    int flo = hist[lo], fhi = hist[hi];
    int t = (alo + ahi + 1) >> 1;
    for (int i = 0; i < 16; i++) {
        int os = b.p[i];
        if (os < t) {
            a[(i >> 2) * as + (i & 3)] = (uint8_t) (os + ((flo * (alo - os)) / 16));
        } else if (os > t) {
            a[(i >> 2) * as + (i & 3)] = (uint8_t) (os + ((fhi * (ahi - os)) / 16));
        }
    }
}

/* reductions over the lanes that share a cell */
#ifdef DSVCU_EMU
#define F_GRED(v, op) ((void) 0)
#else
#define F_GMASK F_GROUP_MASK
#define F_GRED(v, op)                                         \
    do {                                                      \
        for (int o_ = 1; o_ < FILT_LPC; o_ <<= 1) {           \
            int w_ = __shfl_xor_sync(F_GMASK, (v), o_);       \
            (v) = op((v), w_);                                \
        }                                                     \
    } while (0)
#endif
DSVCU_DEV int f_add(int a, int b) { return a + b; }

/* degrad4x4 on a staged cell, the cell's lanes taking its rows.  Only the lowest
 * and the highest occupied histogram bin matter (bmc.c:296-312), so the 16-bin
 * histogram reduces to a min/max and two (count, sum) pairs. */
DSVCU_DEV void
f_degrad_grp(uint8_t *a, int as)
{
    int lo = 16, hi = -1;
    for (int r = FILT_SUB; r < 4; r += FILT_NSUB) {
        for (int c = 0; c < 4; c++) {
            int t = a[r * as + c] >> 4;
            lo = min(lo, t);
            hi = max(hi, t);
        }
    }
    F_GRED(lo, min);
    F_GRED(hi, max);
    if (lo >= hi) return;
    int cnt = 0, sums = 0; /* lo in the low half, hi in the high half */
    for (int r = FILT_SUB; r < 4; r += FILT_NSUB) {
        for (int c = 0; c < 4; c++) {
            int p = a[r * as + c], t = p >> 4;
            if (t == lo) {
                cnt += 1;
                sums += p;
            }
            if (t == hi) {
                cnt += 1 << 16;
                sums += p << 16;
            }
        }
    }
    F_GRED(cnt, f_add);
    F_GRED(sums, f_add);
    const int flo = cnt & 0xffff, fhi = cnt >> 16;
    int alo = (sums & 0xffff) / flo, ahi = (sums >> 16) / fhi;
    if (alo == 0) alo = 1;
    if (ahi == 0) ahi = 1;
    const int t = (alo + ahi + 1) >> 1;
    for (int r = FILT_SUB; r < 4; r += FILT_NSUB) {
        for (int c = 0; c < 4; c++) {
            int os = a[r * as + c];
            if (os < t) {
                a[r * as + c] = (uint8_t) (os + ((flo * (alo - os)) / 16));
            } else if (os > t) {
                a[r * as + c] = (uint8_t) (os + ((fhi * (ahi - os)) / 16));
            }
        }
    }
}

DSVCU_DEV int
f_curve_tex(int tt)
{
    if (tt < 8) return (8 - tt) * 8;
    if (tt > 192) return 0;
    return tt - 7;
}

/* dsv_neighbordif2 (dsv.c:399-436) */
DSVCU_DEV void
f_neighbordif2(const dsvcu_mv *vecs, int nbh, int x, int y, int *dx, int *dy)
{
    const dsvcu_mv *cmv = vecs + x + y * nbh;
    int cmx = cmv->x, cmy = cmv->y;
    if (f_abs(cmx) < 2 && f_abs(cmy) < 2) {
        *dx = *dy = 0;
        return;
    }
    int vx0 = cmx, vx1 = cmx, vy0 = cmy, vy1 = cmy;
    if (x > 0) {
        const dsvcu_mv *mv = cmv - 1;
        if ((mv->x | mv->y) != 0 && !(mv->flags & MVF_SKIP)) {
            vx0 = mv->x;
            vy0 = mv->y;
        }
    }
    if (y > 0) {
        const dsvcu_mv *mv = cmv - nbh;
        if ((mv->x | mv->y) != 0 && !(mv->flags & MVF_SKIP)) {
            vx1 = mv->x;
            vy1 = mv->y;
        }
    }
    *dx = f_abs(vx0 - cmx) + f_abs(vy0 - cmy);
    *dy = f_abs(vx1 - cmx) + f_abs(vy1 - cmy);
}

/* Everything a cell needs that depends only on block data (vectors, flags):
 * evaluated by one lane per cell while the warp looks for active cells, then
 * broadcast to the warp when the cell is processed -- the scalar set-up
 * (divisions, vector loads, neighbour differences) leaves the serial chain. */
struct FPrep {
    int mvxy;  /* x | y << 16 */
    int bits;  /* flags (8) | submask << 8 | edgeh << 16 | edgehs << 17 | edgev << 18 | edgevs << 19 | blockdata << 24 */
    int nd;    /* ndx | ndy << 16 */
    int active;
};

DSVCU_DEV int
f_mod(int v, int m)
{
    return (m & (m - 1)) ? v % m : (v & (m - 1));
}

/* Block-level half of the preparation: what every cell of motion block
 * (fx, fy) shares.  Built for all blocks a CTA can touch before its first step
 * (table in shared memory), so the serial chain only adds the edge flags. */
struct FBlk {
    int mvxy; /* x | y << 16 */
    int nd;   /* ndx | ndy << 16 */
    int bits; /* flags (8) | submask << 8 | blockdata << 16 | active << 24 */
};

DSVCU_DEV FBlk
f_blk(const FiltArgs &A, int fx, int fy)
{
    FBlk B;
    B.mvxy = 0;
    B.nd = 0;
    B.bits = 0;
    if (A.mode == FILT_MODE_INTRA) {
        int bd = A.blockdata[fx + fy * A.nbh];
        B.bits = (bd << 16) | ((!(bd & BD_RING)) << 24);
        return B;
    }
    const dsvcu_mv mv = A.mvs[fx + fy * A.nbh];
    int ndx = 0, ndy = 0, active;
    B.mvxy = (mv.x & 0xffff) | ((int) mv.y << 16);
    B.bits = (int) (mv.flags & 255u) | ((int) mv.submask << 8);
    if (mv.flags & MVF_SKIP) return B;
    if (A.do_filter && !(mv.flags & MVF_INTRA)) {
        f_neighbordif2(A.mvs, A.nbh, fx, fy, &ndx, &ndy);
    }
    B.nd = (ndx & 0xffff) | (ndy << 16);
    active = (mv.flags & MVF_INTRA) || (A.do_filter && (ndx || ndy)) ||
             (A.sharpen && (mv.x & 3) && (mv.y & 3) && ((mv.x | mv.y) & 1) && f_abs(mv.x) < 8 && f_abs(mv.y) < 8);
    B.bits |= active << 24;
    return B;
}

/* cell (i, j) of a block whose shared part is B */
DSVCU_DEV FPrep
f_prep_cell(const FiltArgs &A, int i, int j, const FBlk &B)
{
    FPrep P;
    const int x = i * 4, y = j * 4;
    P.mvxy = B.mvxy;
    P.nd = B.nd;
    P.bits = 0;
    P.active = 0;
    if (y + 4 >= A.h || x + 4 >= A.w) return P;
    P.active = (B.bits >> 24) & 1;
    if (A.mode == FILT_MODE_INTRA) {
        P.bits = ((B.bits >> 16) & 255) << 24;
        return P;
    }
    P.bits = (B.bits & 0xffff) | ((f_mod(x, A.blk_w) == 0) << 16) | ((f_mod(x, A.blk_w / 2) == 0) << 17) |
             ((f_mod(y, A.blk_h) == 0) << 18) | ((f_mod(y, A.blk_h / 2) == 0) << 19);
    return P;
}

/* (fx, fy) = motion block of cell (i, j) as the reference maps it */
DSVCU_DEV FPrep
f_prep(const FiltArgs &A, int i, int j)
{
    const int nsbx = A.w / 4, nsby = A.h / 4;
    return f_prep_cell(A, i, j, f_blk(A, i * A.nbh / max(nsbx, 1), j * A.nbv / max(nsby, 1)));
}

/* chroma_filter thresholds of motion block (i, j) (bmc.c:620-640):
 * tx | ty << 15 | active << 30 */
DSVCU_DEV int
f_chroma_blk(const FiltArgs &A, int i, int j)
{
    const dsvcu_mv mv = A.mvs[i + j * A.nbh];
    if (mv.flags & MVF_SKIP) return 0;
    int it = f_clamp((64 * A.q) >> 12, 2, 32);
    int tx = it, ty = it;
    if (!(mv.flags & MVF_INTRA)) {
        int ndx, ndy;
        f_neighbordif2(A.mvs, A.nbh, i, j, &ndx, &ndy);
        int amx = f_abs(mv.x), amy = f_abs(mv.y);
        if (ndx < amy && ndy < amx) {
            tx = ty = 0;
        } else {
            tx = (min(ndy, 64) * A.q) >> 12;
            ty = (min(ndx, 64) * A.q) >> 12;
        }
    }
    tx = f_clamp(tx, 0, 32767);
    ty = f_clamp(ty, 0, 32767);
    return tx | (ty << 15) | ((tx > 0 || ty > 0) << 30);
}

/* ---- per-cell staging.  A cell reads and writes inside the 11 x 11 pixel
 * neighbourhood rows y-3..y+7, cols x-3..x+7.  Instead of three dependent
 * round trips to L2 (texture probe, horizontal pass, vertical pass) the warp
 * copies the neighbourhood (11 rows x 3 aligned words) into shared memory in
 * one batch, filters there, and writes back only the words its filters own
 * (rows y..y+3 for the horizontal pass, column word x..x+3 rows y-2..y+6 for the
 * vertical pass, the cell itself for the sharpener) -- the same pixels the
 * wavefront protocol already reserves for this cell. ---- */
#define FT_S 16                      /* tile pitch */
#define FT_ROWS 11
#define FT_BYTES (FT_ROWS * FT_S + 4)  /* 45 words: odd pitch between tiles, no bank conflicts */
#define FT_ORG (3 * FT_S + 4)        /* tile offset of pixel (x, y) */

/* 4-byte copy between word-aligned pixel addresses */
#ifdef DSVCU_EMU
#define F_CP4(dst, src) memcpy((dst), (src), 4)
#else
#define F_CP4(dst, src) (*(uint32_t *) (dst) = *(const uint32_t *) (src))
#endif

DSVCU_DEV void
f_tile_load(uint8_t *T, const FiltArgs &A, int x, int y, bool vol)
{
    /* Pixels a cell reads were last written by cell rows r-2 .. r.  When all of
     * them belong to this CTA, a plain (L1-cached) load observes their stores
     * once the step barrier has been passed; the first two rows of a CTA read
     * what another SM wrote and go to L2.  The cell's lanes take the tile rows
     * round-robin; all loads are issued before the first store so the batch
     * costs one memory round trip. */
    const int NIT = (FT_ROWS + FILT_NSUB - 1) / FILT_NSUB;
    uint32_t v[NIT][3];
    const uint8_t *g = A.data + (ptrdiff_t) (y - 3 + FILT_SUB) * A.stride + x - 4;
#pragma unroll
    for (int n = 0; n < NIT; n++) {
        if (FILT_SUB + n * FILT_NSUB < FT_ROWS) {
#pragma unroll
            for (int q = 0; q < 3; q++) {
#ifndef DSVCU_EMU
                v[n][q] = vol ? *(volatile const uint32_t *) (g + 4 * q) : *(const uint32_t *) (g + 4 * q);
#else
                (void) vol;
                memcpy(&v[n][q], g + 4 * q, 4);
#endif
            }
        }
        g += (ptrdiff_t) FILT_NSUB * A.stride;
    }
    uint8_t *t = T + FILT_SUB * FT_S;
#pragma unroll
    for (int n = 0; n < NIT; n++) {
        if (FILT_SUB + n * FILT_NSUB < FT_ROWS) {
#pragma unroll
            for (int q = 0; q < 3; q++) F_CP4(t + 4 * q, &v[n][q]);
        }
        t += FILT_NSUB * FT_S;
    }
    F_GSYNC();
}

/* what: 1 = horizontal pass region (rows y..y+3, all three words), 2 = vertical
 * pass region (the cell's column word, rows y-2..y+6), 4 = the cell */
DSVCU_DEV void
f_tile_store(const uint8_t *T, const FiltArgs &A, int x, int y, int what)
{
    F_GSYNC();
    for (int r = 1 + FILT_SUB; r < 10; r += FILT_NSUB) {
        const bool own = r >= 3 && r < 7; /* the cell's rows */
        uint8_t *g = A.data + (ptrdiff_t) (y - 3 + r) * A.stride + x - 4;
        const uint8_t *t = T + r * FT_S;
        if (own && (what & 1)) {
            F_CP4(g, t);
            F_CP4(g + 4, t + 4);
            F_CP4(g + 8, t + 8);
        } else if ((what & 2) || (own && (what & 4))) {
            F_CP4(g + 4, t + 4);
        }
    }
}

/* view of the tile with the plane's coordinates (pixel (x,y) at T + FT_ORG) */
DSVCU_DEV FiltArgs
f_tile_view(const FiltArgs &A, uint8_t *T, int x, int y)
{
    FiltArgs L = A;
    L.data = T + FT_ORG - ((ptrdiff_t) y * FT_S + x);
    L.stride = FT_S;
    return L;
}

/* one 4x4 cell of luma_filter (bmc.c:492-600); P = f_prep() of this cell */
DSVCU_DEV int
f_luma_cell(const FiltArgs &G, uint8_t *T, int i, int j, const FPrep &P, bool vol)
{
    const FiltArgs &A = G;
    const int x = i * 4, y = j * 4;
    int touched = 0;
    struct {
        int x, y;
        unsigned flags, submask;
    } mv;
    mv.x = (int16_t) (P.mvxy & 0xffff);
    mv.y = (int16_t) (P.mvxy >> 16);
    mv.flags = (unsigned) P.bits & 255u;
    mv.submask = ((unsigned) P.bits >> 8) & 255u;
    const int edgeh = (P.bits >> 16) & 1, edgehs = (P.bits >> 17) & 1;
    const int edgev = (P.bits >> 18) & 1, edgevs = (P.bits >> 19) & 1;
    const int amx = f_abs(mv.x), amy = f_abs(mv.y);
    const int q = A.q;
    uint8_t *dxy = T + FT_ORG;
    const FiltArgs L = f_tile_view(A, T, x, y);
    int what = 0;

    if (mv.flags & MVF_INTRA) {
        int tH = f_clamp((64 * q) >> 12, 2, 32), tL = f_clamp((32 * q) >> 12, 2, 32);
        int teh = edgeh, tev = edgev;
        if (mv.submask != 15) {
            teh |= edgehs;
            tev |= edgevs;
        }
        f_tile_load(T, A, x, y, vol);
        f_hfilter(L, x, y, teh, tH, tL);
        F_GSYNC();
        f_vfilter(L, x, y, tev, tH, tL);
        f_tile_store(T, A, x, y, 3);
        return 1;
    }
    int ndx = P.nd & 0xffff, ndy = (P.nd >> 16) & 0xffff;
    if (A.do_filter && (ndx || ndy)) {
        int tt, addx, addy, sh, sv, shl, svl;
        int eprm = (mv.flags & MVF_EPRM) != 0;
        int teh = edgeh || eprm, tev = edgev || eprm;
        int tndc = (ndx + ndy + 1) >> 1;
        F4x4 b;
        f_tile_load(T, A, x, y, vol);
        what |= 8; /* tile is loaded */
        f_load4x4(b, dxy, FT_S);
        f_artf(b, &sh, &sv, &shl, &svl);
        if (sh < 2 * sv && sv < 2 * sh) {
            if (ndx < amx) ndx >>= 1;
            if (ndy < amy) ndy >>= 1;
            shl = (shl > 128) ? 0 : (128 - shl);
            svl = (svl > 128) ? 0 : (128 - svl);
            int ix = min(amx, 32), iy = min(amy, 32);
            tt = ((sh * (32 - iy) + shl * iy) + 16) >> 5;
            tt += ((sv * (32 - ix) + svl * ix) + 16) >> 5;
            tt = (tt + 1) >> 1;
            if (ndx < amy && ndy < amx) tt = 0;
        } else {
            tt = (sh + sv + 1) >> 1;
        }
        tt = (tt * tndc + 4) >> 3;
        tt = (min(tt, A.fthresh) * q) >> 12;
        addx = (min(ndy, A.fthresh) * q) >> 12;
        addy = (min(ndx, A.fthresh) * q) >> 12;
        F_GSYNC();
        if (sh > 2 * sv || amy > 2 * amx) {
            f_vfilter(L, x, y, tev, tt + addy, tt);
            what |= 2;
        } else if (sv > 2 * sh || amx > 2 * amy) {
            f_hfilter(L, x, y, teh, tt + addx, tt);
            what |= 1;
        } else {
            f_hfilter(L, x, y, teh, tt + addx, tt);
            F_GSYNC();
            f_vfilter(L, x, y, tev, tt + addy, tt);
            what |= 3;
        }
        F_GSYNC();
        touched = 1;
    }
    if (A.sharpen && (mv.x & 3) && (mv.y & 3) && ((mv.x | mv.y) & 1) && amx < 8 && amy < 8) {
        if (!(what & 8)) f_tile_load(T, A, x, y, vol);
        f_degrad_grp(dxy, FT_S);
        what |= 4;
        touched = 1;
    }
    if (what & 7) f_tile_store(T, A, x, y, what & 7);
    return touched;
}

/* one 4x4 cell of dsv_intra_filter (bmc.c:411-455) */
DSVCU_DEV int
f_intra_cell(const FiltArgs &G, uint8_t *T, int i, int j, const FPrep &P, bool vol)
{
    const FiltArgs &A = G;
    const int x = i * 4, y = j * 4;
    const int flags = (P.bits >> 24) & 255;
    const int q = A.q;
    uint8_t *dxy = T + FT_ORG;
    const FiltArgs L = f_tile_view(A, T, x, y);
    int sh, sv, shl, svl, tt = 32;
    F4x4 b;
    f_tile_load(T, A, x, y, vol);
    f_load4x4(b, dxy, FT_S);
    f_artf(b, &sh, &sv, &shl, &svl);
    int mxs = max(sh, sv);
    if (!(mxs < 256 && mxs > 8)) return 0;
    if (flags & (BD_MAINTAIN | BD_STABLE)) {
        tt = f_dsff(b);
        if (flags & BD_STABLE) tt = tt * 5 >> 2;
    } else {
        tt >>= 2;
    }
    tt = tt * 2 / 3;
    tt = (tt * q) >> 12;
    tt = f_clamp(tt, 0, A.fthresh);
    F_GSYNC();
    f_hfilter(L, x, y, 0, tt, tt);
    F_GSYNC();
    f_vfilter(L, x, y, 0, tt, tt);
    F_GSYNC();
    tt = (sh > sv) ? (3 * sh + sv) : (3 * sv + sh);
    tt = f_curve_tex(tt);
    tt = 16 + ((tt + 2) >> 2);
    tt = (tt * q) >> 12;
    tt = f_clamp(tt, 0, A.fthresh);
    f_hfilter(L, x, y, 0, tt, tt);
    F_GSYNC();
    f_vfilter(L, x, y, 0, tt, tt);
    f_tile_store(T, A, x, y, 3);
    return 1;
}

/* one motion block of chroma_filter (bmc.c:620-657) */
template <bool VOL>
DSVCU_DEV int
f_chroma_cell(const FiltArgs &A, int i, int j, int blk)
{
    const int x = i * A.bw, y = j * A.bh;
    const int tx = blk & 32767, ty = (blk >> 15) & 32767;
    /* left side: the bh/4 row groups are independent, lanes take lines */
    if (!(x < 4 || x > A.w - 4 || tx <= 0)) {
        int in_edge = x < (A.w - 8);
        for (int k = FILT_SUB; k < A.bh; k += FILT_NSUB) {
            int z = k & ~3;
            if (y + z + 4 < A.h) {
                int top = f_clamp(y + z, 0, A.h - 1), bot = f_clamp(y + z + 4, 0, A.h - 1);
                int r = y + k;
                if (r >= top && r < bot) {
                    f_edge_line_t<VOL>(A.data + (size_t) r * A.stride + x, 1, tx, tx, in_edge);
                }
            }
        }
    }
    F_GSYNC();
    if (!(y < 4 || y > A.h - 4 || ty <= 0)) {
        int in_edge = y < (A.h - 8);
        for (int k = FILT_SUB; k < A.bw; k += FILT_NSUB) {
            int z = k & ~3;
            if (x + z + 4 < A.w) {
                int beg = f_clamp(x + z, 0, A.w - 1), end = f_clamp(x + z + 4, 0, A.w - 1);
                int c = x + k;
                if (c >= beg && c < end) {
                    f_edge_line_t<VOL>(A.data + (size_t) y * A.stride + c, A.stride, ty, ty, in_edge);
                }
            }
        }
    }
    F_GSYNC();
    return (tx > 0) || (ty > 0);
}

/* The skewed-lockstep schedule described in the file header.
 *
 * Step t of a band: group g handles cell t - 2g of row band*FILT_G + g.  The
 * first row of band b needs the last row of band b-1 to have COMPLETED cell
 * t+1, i.e. band b-1 to have completed step t + 2*FILT_G - 1: band b waits for
 * done[b-1] >= min(t + 2*FILT_G, nsteps).  A step in which no cell of the band
 * can touch pixels costs one look-up in the activity bitmap and a ballot.
 *
 * CTA c of a plane owns bands c*FILT_WPC .. +FILT_WPC-1 (one warp each); CTA c
 * only ever waits for CTA c-1, which the hardware dispatches first.
 * Dynamic shared memory: FILT_CELLS tiles | done[FILT_WPC] | activity bitmap of
 * the block rows this CTA's cells map to. */
DSVCU_HD int
filt_block_rows(const FiltArgs &A)
{
    /* upper bound on the motion-block rows one CTA's cell rows can map to */
    if (A.mode == FILT_MODE_CHROMA) return FILT_CELLS;
    return FILT_CELLS * 4 / max(A.blk_h, 4) + 2;
}

DSVCU_HD int
filt_smem_bytes(const FiltArgs &A)
{
    const int per_block = A.mode == FILT_MODE_CHROMA ? 4 : (int) sizeof(FBlk);
    return FILT_CELLS * FT_BYTES + 4 * FILT_WPC + per_block * filt_block_rows(A) * A.nbh + 16;
}

DSVCU_HD int
filt_ctas(const FiltArgs &A)
{
    return (A.nrows + FILT_CELLS - 1) / FILT_CELLS;
}

DSVCU_KERNEL void __launch_bounds__(FILT_THREADS)
k_filter_skew(FiltJob J)
{
    DSVCU_DYN_SMEM(uint8_t, smem);
    /* A CTA waits for the band above, which belongs to the CTA with the next lower index.  The
     * logical index is a ticket drawn at entry, not blockIdx: whoever holds ticket k started after
     * the holders of 0..k-1, so a waiting CTA only ever waits for CTAs that are already running
     * (or done), whatever order the hardware dispatches blocks in. */
#ifndef DSVCU_EMU
    __shared__ int s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(J.ticket, 1);
    __syncthreads();
    const int bid = s_ticket;
#else
    const int bid = (int) blockIdx.x;
#endif
    int pl = 0;
    while (pl + 1 < J.nplanes && bid >= J.p[pl + 1].first_cta) pl++;
    const FiltArgs &A = J.p[pl];
    const int ncols = A.ncols, nrows = A.nrows;
    int *done = (int *) (smem + FILT_CELLS * FT_BYTES);
    int *btab = done + FILT_WPC; /* chroma: one word per block; luma/intra: FBlk */
#ifndef DSVCU_EMU
    const int cta = bid - A.first_cta;
    const int lane = (int) (threadIdx.x & 31), warp = (int) (threadIdx.x >> 5);
    const int grp = lane / FILT_LPC;
    const int nbands = (nrows + FILT_G - 1) / FILT_G;
    const int nsteps = ncols + 2 * (FILT_G - 1);
    const int nsbx = max(A.w / 4, 1), nsby = max(A.h / 4, 1);
    const bool chroma = A.mode == FILT_MODE_CHROMA;
    const int band = cta * FILT_WPC + warp;
    const int row = band * FILT_G + grp;
    uint8_t *T = smem + (warp * FILT_G + grp) * FT_BYTES;
    /* block rows this CTA can touch -> block table */
    const int r0 = cta * FILT_CELLS, r1 = min(nrows, r0 + FILT_CELLS) - 1;
    const int fy0 = chroma ? r0 : min(r0 * A.nbv / nsby, A.nbv - 1);
    const int fy1 = chroma ? r1 : min(r1 * A.nbv / nsby, A.nbv - 1);
    const int nmap = (fy1 - fy0 + 1) * A.nbh;
    if (threadIdx.x < FILT_WPC) done[threadIdx.x] = 0;
    for (int b = (int) threadIdx.x; b < nmap; b += (int) blockDim.x) {
        const int fx = b % A.nbh, fy = fy0 + b / A.nbh;
        if (chroma) {
            btab[b] = f_chroma_blk(A, fx, fy);
        } else {
            ((FBlk *) btab)[b] = f_blk(A, fx, fy);
        }
    }
    __syncthreads();
    if (band >= nbands) return;
    const bool row_ok = row < nrows && (chroma || row * 4 + 4 < A.h);
    const int fy = chroma ? min(row, A.nbv - 1) : (row_ok ? min(row * A.nbv / nsby, A.nbv - 1) : fy0);
    const int abase = (fy - fy0) * A.nbh;
    /* hand-off words: from the band above / to the band below */
    const bool above_global = warp == 0, pub_global = warp == FILT_WPC - 1 && band + 1 < nbands;
    /* the CTA below reads pixels written by this CTA's last two cell rows */
    const bool dev_fence = band + 1 < nbands && warp * FILT_G + FILT_G - 1 >= FILT_CELLS - 2;
    volatile const int *above = above_global ? (volatile const int *) (A.progress + band - 1) : (volatile const int *) (done + warp - 1);
    /* cell rows whose footprint holds pixels written by the CTA above (luma: its
     * first two rows; chroma blocks: the first) read them from L2 */
    const bool vol = cta > 0 && warp * FILT_G + grp < (chroma ? 1 : 2);
    int seen = band ? 0 : 0x7fffffff;
    int fx = 0, rem = 0; /* luma: i * nbh = fx * nsbx + rem */
    for (int t = 0; t < nsteps; t++) {
        const int i = t - 2 * grp;
        bool act = false;
        int cblk = 0;
        FBlk B;
        if (i >= 0 && i < ncols && row_ok) {
            if (chroma) {
                cblk = btab[abase + i];
                act = (cblk >> 30) & 1;
            } else {
                B = ((const FBlk *) btab)[abase + min(fx, A.nbh - 1)];
                act = i * 4 + 4 < A.w && ((B.bits >> 24) & 1);
                rem += A.nbh;
                if (rem >= nsbx) {
                    rem -= nsbx;
                    fx++;
                }
            }
        }
        if (__ballot_sync(0xffffffffu, act)) {
            const int need = min(t + 2 * FILT_G, nsteps);
            if (seen < need) {
                unsigned ns = 32;
                while ((seen = *above) < need) {
                    __nanosleep(ns);
                    if (ns < 512) ns <<= 1;
                }
                if (above_global) {
                    __threadfence();
                } else {
                    __threadfence_block();
                }
            }
            if (act) {
                if (A.mode == FILT_MODE_LUMA) {
                    f_luma_cell(A, T, i, row, f_prep_cell(A, i, row, B), vol);
                } else if (A.mode == FILT_MODE_INTRA) {
                    f_intra_cell(A, T, i, row, f_prep_cell(A, i, row, B), vol);
                } else if (vol) {
                    f_chroma_cell<true>(A, i, row, cblk);
                } else {
                    f_chroma_cell<false>(A, i, row, cblk);
                }
            }
            if (dev_fence) {
                __threadfence();
            } else {
                __threadfence_block();
            }
            __syncwarp();
        }
        if (lane == 0) {
            done[warp] = t + 1;
            if (pub_global) *(volatile int *) (A.progress + band) = t + 1;
        }
    }
#else
    (void) done;
    (void) btab;
    if (bid != A.first_cta) return; /* emulation: the plane's first CTA walks it in raster order */
    for (int row = 0; row < nrows; row++) {
        for (int i = 0; i < ncols; i++) {
            if (A.mode == FILT_MODE_CHROMA) {
                int blk = f_chroma_blk(A, i, row);
                if ((blk >> 30) & 1) f_chroma_cell<false>(A, i, row, blk);
            } else {
                FPrep P = f_prep(A, i, row);
                if (!P.active) continue;
                if (A.mode == FILT_MODE_LUMA) {
                    f_luma_cell(A, smem, i, row, P, false);
                } else {
                    f_intra_cell(A, smem, i, row, P, false);
                }
            }
        }
    }
#endif
}

/* dsv_post_process (bmc.c:340-361): cells are disjoint -> fully parallel */
DSVCU_KERNEL void __launch_bounds__(256)
k_post_sharpen(uint8_t *data, int stride, int w, int h)
{
    const int nsbx = w / 4, nsby = h / 4;
    const int total = nsbx * nsby;
    for (int k = (int) blockIdx.x * DSVCU_NTH + DSVCU_TID; k < total; k += (int) gridDim.x * DSVCU_NTH) {
        int j = k / nsbx, i = k - j * nsbx;
        int x = i * 4, y = j * 4;
        if (y + 4 >= h || x + 4 >= w) continue;
        f_degrad(data + (size_t) y * stride + x, stride);
    }
}

#endif /* K_FILTER_CUH */
