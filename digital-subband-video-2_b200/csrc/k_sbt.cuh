/*
 * k_sbt.cuh -- integer subband transforms (forward + inverse) for sm_100a.
 *
 * Replaces reference src/sbt.c: dsv_fwd_sbt (:847-886), dsv_inv_sbt (:889-934),
 * the 1-D lifting filters (:278-447), the 2-D drivers (:449-544) and the three
 * Haar kernels (:546-795), with p2sbc/sbc2p (:798-831) fused into level 1.
 *
 * Design (not a port): the reference runs each level as whole-image row and
 * column passes over one in-place buffer plus a scratch copy.  Here one launch
 * per level does BOTH dimensions of a 64x32 output tile in shared memory
 * (phase-parallel lifting: all odd samples, then all even samples), reads the
 * four packed quadrants with halo once and writes each output once.  The LL
 * quadrant ping-pongs between two scratch planes, so nothing is copied back,
 * and level 1 converts from/to the u8 frame directly.
 *
 * Index conventions: a level works on the top-left sw x sh "sub-image" of a
 * plane whose row stride is fw.  Interleaved sample (X,Y) lives at packed
 * position (X even ? X/2 : cw + X/2, Y even ? Y/2 : ch + Y/2), cw=ceil(sw/2).
 */
#ifndef K_SBT_CUH
#define K_SBT_CUH

#include "dsvcu_rt.h"

#define SBT_F_LLI 0
#define SBT_F_LLP 1
#define SBT_F_CC 2
#define SBT_F_L2A 3
#define SBT_F_L1 4
#define SBT_F_LOSSLESS 5
#define SBT_F_HAAR 6        /* filtered inverse / plain forward */
#define SBT_F_HAAR_SIMPLE 7 /* simple inverse */

#define SBT_TW 64
#define SBT_TH 32
#define SBT_HALO 4
#define SBT_SW (SBT_TW + 2 * SBT_HALO)
#define SBT_SH (SBT_TH + 2 * SBT_HALO)
#define SBT_THREADS 256

#define BLOCK_INTERP_P 14
#define BD_RINGING (1 << 3) /* DSV_IS_RINGING, dsv_internal.h:106 */

struct SbtLevel {
    int fw;             /* row stride (ints) of every coefficient buffer */
    int sw, sh;         /* sub-image size at this level */
    int cw, ch;         /* low-band size = ceil(sw/2), ceil(sh/2) */
    const int32_t *ll;  /* inverse: where the LL quadrant comes from */
    const int32_t *bands; /* inverse: packed LH/HL/HH (the coefficient plane) */
    int32_t *dst;       /* inverse: interleaved output (l > 1) */
    uint8_t *px;        /* level 1: u8 plane (inverse output / forward input) */
    int px_stride, px_w, px_h;
    const int32_t *src; /* forward: interleaved input (l > 1) */
    int32_t *out_ll;    /* forward: where LL goes */
    int32_t *out_bands; /* forward: where LH/HL/HH go (the coefficient plane) */
    const uint8_t *blockdata;
    int nbh;            /* blocks per row */
    int dbx, dby;       /* 14-bit fixed-point block steps for this sub-image */
    int hqp;            /* filtered Haar clamp */
    int ovf;            /* overflow-safety flag */
};

/* ---------------------------------------------------------------- helpers */

DSVCU_HD int sbt_sar(int v, int s) { return v >> s; } /* arithmetic on int in CUDA and gcc */

DSVCU_HD int sbt_reflect(int i, int n) /* sbt.c:105-115 with n := len-1 */
{
    if (i < 0) i = -i;
    if (i >= n) i = n + n - i;
    return i;
}

DSVCU_HD int sbt_round2(int v) { return (v + (v < 0 ? -1 : 1)) / 2; }
DSVCU_HD int sbt_round4(int v) { return (v + (v < 0 ? -2 : 2)) / 4; }
DSVCU_HD int sbt_clamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <int F> DSVCU_HD int sbt_fwd_scale_l(int x)
{
    if (F == SBT_F_LLI || F == SBT_F_LLP) return x * 5 / 2;
    if (F == SBT_F_CC || F == SBT_F_L2A || F == SBT_F_L1) return x * 2;
    return x;
}
template <int F> DSVCU_HD int sbt_fwd_scale_h(int x)
{
    if (F == SBT_F_LLI || F == SBT_F_L1) return x * 4;
    if (F == SBT_F_LLP) return x * 2;
    if (F == SBT_F_L2A) {
        int th = x * 3;
        return th - sbt_sar(th, 3);
    }
    return x;
}
template <int F> DSVCU_HD int sbt_inv_scale_l(int x)
{
    if (F == SBT_F_LLI || F == SBT_F_LLP) return x * 2 / 5;
    if (F == SBT_F_CC || F == SBT_F_L2A || F == SBT_F_L1) return x / 2;
    return x;
}
template <int F> DSVCU_HD int sbt_inv_scale_h(int x)
{
    if (F == SBT_F_LLI || F == SBT_F_L1) return x / 4;
    if (F == SBT_F_LLP) return x / 2;
    if (F == SBT_F_L2A) {
        int th = x / 3;
        return th + sbt_sar(th, 3);
    }
    return x;
}

/* low-pass lifting term for even sample i of a line of length n held in
 * v[(i - base) * s]; `ring` selects the L2 ringing taps (sbt.c:119-146). */
template <int F> DSVCU_HD int
sbt_lo_term(const int32_t *v, int s, int base, int i, int n, int ring)
{
    if (i == 0) {
        return v[(1 - base) * s] >> 1;
    }
    if (F == SBT_F_CC || F == SBT_F_L2A) {
        int c0, ca, cs;
        if (F == SBT_F_CC) {
            c0 = 3; ca = 8; cs = 4;
        } else if (ring) {
            c0 = 3; ca = 4; cs = 3;
        } else {
            c0 = 9; ca = 16; cs = 5;
        }
        int a = v[(sbt_reflect(i - 3, n - 1) - base) * s];
        int d = v[(sbt_reflect(i + 3, n - 1) - base) * s];
        int b = v[(i - 1 - base) * s];
        int c = v[(i + 1 - base) * s];
        return (-a + c0 * (b + c) - d + ca) >> cs;
    }
    return (v[(i - 1 - base) * s] + v[(i + 1 - base) * s] + 2) >> 2;
}

/* high-pass lifting term for odd sample i (sbt.c:190-197) */
DSVCU_HD int
sbt_hi_term(const int32_t *v, int s, int base, int i, int n)
{
    if (i < n - 1) {
        return (v[(i - 1 - base) * s] + v[(i + 1 - base) * s] + 1) >> 1;
    }
    return v[(i - 1 - base) * s]; /* n even, last sample */
}

/* --------------------------------------------------- inverse lifting level */

template <int F>
DSVCU_DEV void
sbt_inv_lift_tile(const SbtLevel &L, int32_t *t, int tile_x, int tile_y)
{
    const int X0 = tile_x * SBT_TW - SBT_HALO;
    const int Y0 = tile_y * SBT_TH - SBT_HALO;
    const int sw = L.sw, sh = L.sh, cw = L.cw, ch = L.ch, fw = L.fw;
    const int even_w = sw & ~1, even_h = sh & ~1;

    /* gather the four quadrants (with halo) into interleaved order and apply
     * the vertical un-scale (sbt.c:160-168 applied along columns first) */
    PAR_FOR(k, SBT_SH * SBT_SW) {
        int ty = k / SBT_SW, tx = k - ty * SBT_SW;
        int X = X0 + tx, Y = Y0 + ty;
        int v = 0;
        if (X >= 0 && X < sw && Y >= 0 && Y < sh) {
            int pxk = (X & 1) ? cw + (X >> 1) : (X >> 1);
            int pyk = (Y & 1) ? ch + (Y >> 1) : (Y >> 1);
            const int32_t *p = (pxk < cw && pyk < ch) ? L.ll : L.bands;
            v = p[pyk * fw + pxk];
            v = (Y & 1) ? sbt_inv_scale_h<F>(v) : sbt_inv_scale_l<F>(v);
        }
        t[k] = v;
    }
    DSVCU_SYNC();
    /* vertical: even rows  v -= lo(odd neighbours) */
    PAR_FOR(k, (SBT_SH / 2) * SBT_SW) {
        int ty = (k / SBT_SW) * 2, tx = k % SBT_SW;
        int X = X0 + tx, Y = Y0 + ty;
        if (ty >= 3 && ty < SBT_SH - 3 && X >= 0 && X < sw && Y >= 0 && Y < even_h) {
            int ring = 0;
            if (F == SBT_F_L2A && Y >= 2) {
                /* column pass of inv_L2a_2d: block column from the PACKED column
                 * index, block row stepping 2*dby per even sample (sbt.c:520-535) */
                int pxk = (X & 1) ? cw + (X >> 1) : (X >> 1);
                int bc = (pxk * L.dbx) >> BLOCK_INTERP_P;
                int br = (((Y - 2) >> 1) * (2 * L.dby)) >> BLOCK_INTERP_P;
                ring = L.blockdata[bc + br * L.nbh] & BD_RINGING;
            }
            t[ty * SBT_SW + tx] -= sbt_lo_term<F>(t + tx, SBT_SW, Y0, Y, sh, ring);
        }
    }
    DSVCU_SYNC();
    /* vertical: odd rows  v += hi(even neighbours) */
    PAR_FOR(k, (SBT_SH / 2) * SBT_SW) {
        int ty = (k / SBT_SW) * 2 + 1, tx = k % SBT_SW;
        int X = X0 + tx, Y = Y0 + ty;
        if (ty >= 4 && ty < SBT_SH - 4 && X >= 0 && X < sw && Y >= 0 && Y < sh) {
            t[ty * SBT_SW + tx] += sbt_hi_term(t + tx, SBT_SW, Y0, Y, sh);
        }
    }
    DSVCU_SYNC();
    /* horizontal un-scale on rows that will be used */
    PAR_FOR(k, SBT_TH * SBT_SW) {
        int ty = k / SBT_SW + SBT_HALO, tx = k % SBT_SW;
        int X = X0 + tx;
        int v = t[ty * SBT_SW + tx];
        t[ty * SBT_SW + tx] = (X & 1) ? sbt_inv_scale_h<F>(v) : sbt_inv_scale_l<F>(v);
    }
    DSVCU_SYNC();
    PAR_FOR(k, SBT_TH * (SBT_SW / 2)) {
        int ty = k / (SBT_SW / 2) + SBT_HALO, tx = (k % (SBT_SW / 2)) * 2;
        int X = X0 + tx, Y = Y0 + ty;
        if (tx >= 3 && tx < SBT_SW - 3 && Y < sh && X >= 0 && X < even_w) {
            int ring = 0;
            if (F == SBT_F_L2A && X >= 2) {
                /* row pass: block row from the interleaved row (sbt.c:536-540) */
                int br = (Y * L.dby) >> BLOCK_INTERP_P;
                int bc = (((X - 2) >> 1) * (2 * L.dbx)) >> BLOCK_INTERP_P;
                ring = L.blockdata[bc + br * L.nbh] & BD_RINGING;
            }
            t[ty * SBT_SW + tx] -= sbt_lo_term<F>(t + ty * SBT_SW, 1, X0, X, sw, ring);
        }
    }
    if (sw == 1 && tile_x == 0) {
        /* A line of ONE sample: the reference still executes "v[0] -= v[s] >> 1" (sbt.c:199, :221,
         * :296) and, with the rows filtered in place in the coefficient plane (inv_2d, :462-473),
         * v[s] is the plane's coefficient at column 1 of that row -- a sample of a finer level's
         * HL band that no coarser level has touched yet.  (Pictures 8 or more times as tall as
         * wide get here.  The same thing in a COLUMN of one sample reads scratch memory of another
         * level: DESIGN.md, known divergences.) */
        PAR_FOR(k, SBT_TH) {
            int ty = k + SBT_HALO, Y = Y0 + ty;
            if (Y < sh) t[ty * SBT_SW + SBT_HALO] -= L.bands[Y * fw + 1] >> 1;
        }
    }
    DSVCU_SYNC();
    PAR_FOR(k, SBT_TH * (SBT_SW / 2)) {
        int ty = k / (SBT_SW / 2) + SBT_HALO, tx = (k % (SBT_SW / 2)) * 2 + 1;
        int X = X0 + tx, Y = Y0 + ty;
        if (tx >= 4 && tx < SBT_SW - 4 && Y < sh && X >= 0 && X < sw) {
            t[ty * SBT_SW + tx] += sbt_hi_term(t + ty * SBT_SW, 1, X0, X, sw);
        }
    }
    DSVCU_SYNC();
    /* store the interior */
    if (L.px) { /* level 1: +128, clamp, u8 (sbt.c:816-831) */
        PAR_FOR(k, SBT_TH * (SBT_TW / 4)) {
            int ty = k / (SBT_TW / 4), tx = (k % (SBT_TW / 4)) * 4;
            int X = X0 + SBT_HALO + tx, Y = Y0 + SBT_HALO + ty;
            if (Y < L.px_h && X < L.px_w) {
                const int32_t *r = t + (ty + SBT_HALO) * SBT_SW + tx + SBT_HALO;
                uint8_t *o = L.px + (size_t) Y * L.px_stride + X;
                if (X + 3 < L.px_w) {
                    uint32_t w = (uint32_t) sbt_clamp(r[0] + 128, 0, 255) |
                                 ((uint32_t) sbt_clamp(r[1] + 128, 0, 255) << 8) |
                                 ((uint32_t) sbt_clamp(r[2] + 128, 0, 255) << 16) |
                                 ((uint32_t) sbt_clamp(r[3] + 128, 0, 255) << 24);
                    *(uint32_t *) o = w;
                } else {
                    for (int j = 0; j < 4 && X + j < L.px_w; j++) {
                        o[j] = (uint8_t) sbt_clamp(r[j] + 128, 0, 255);
                    }
                }
            }
        }
    } else {
        PAR_FOR(k, SBT_TH * SBT_TW) {
            int ty = k / SBT_TW, tx = k % SBT_TW;
            int X = X0 + SBT_HALO + tx, Y = Y0 + SBT_HALO + ty;
            if (X < sw && Y < sh) {
                L.dst[Y * fw + X] = t[(ty + SBT_HALO) * SBT_SW + tx + SBT_HALO];
            }
        }
    }
}

/* --------------------------------------------------- forward lifting level */

DSVCU_HD int
sbt_fwd_fetch(const SbtLevel &L, int X, int Y)
{
    if (L.px) { /* level 1: u8 - 128; rows past the plane stay 0 (sbt.c:798-813) */
        if (Y >= L.px_h) return 0;
        return (int) L.px[(size_t) Y * L.px_stride + X] - 128;
    }
    return L.src[Y * L.fw + X];
}

template <int F>
DSVCU_DEV void
sbt_fwd_lift_tile(const SbtLevel &L, int32_t *t, int tile_x, int tile_y)
{
    const int X0 = tile_x * SBT_TW - SBT_HALO;
    const int Y0 = tile_y * SBT_TH - SBT_HALO;
    const int sw = L.sw, sh = L.sh, cw = L.cw, ch = L.ch, fw = L.fw;
    const int even_w = sw & ~1, even_h = sh & ~1;

    PAR_FOR(k, SBT_SH * SBT_SW) {
        int ty = k / SBT_SW, tx = k - ty * SBT_SW;
        int X = X0 + tx, Y = Y0 + ty;
        t[k] = (X >= 0 && X < sw && Y >= 0 && Y < sh) ? sbt_fwd_fetch(L, X, Y) : 0;
    }
    DSVCU_SYNC();
    /* rows first (fwd_2d, sbt.c:449-460): odd -= hi, even += lo, scale */
    PAR_FOR(k, SBT_SH * (SBT_SW / 2)) {
        int ty = k / (SBT_SW / 2), tx = (k % (SBT_SW / 2)) * 2 + 1;
        int X = X0 + tx, Y = Y0 + ty;
        if (tx >= 1 && tx < SBT_SW - 1 && Y >= 0 && Y < sh && X >= 0 && X < sw) {
            t[ty * SBT_SW + tx] -= sbt_hi_term(t + ty * SBT_SW, 1, X0, X, sw);
        }
    }
    DSVCU_SYNC();
    PAR_FOR(k, SBT_SH * (SBT_SW / 2)) {
        int ty = k / (SBT_SW / 2), tx = (k % (SBT_SW / 2)) * 2;
        int X = X0 + tx, Y = Y0 + ty;
        if (tx >= 4 && tx < SBT_SW - 4 && Y >= 0 && Y < sh && X >= 0 && X < even_w) {
            int ring = 0;
            if (F == SBT_F_L2A && X >= 2) {
                int br = (Y * L.dby) >> BLOCK_INTERP_P;
                int bc = (((X - 2) >> 1) * (2 * L.dbx)) >> BLOCK_INTERP_P;
                ring = L.blockdata[bc + br * L.nbh] & BD_RINGING;
            }
            t[ty * SBT_SW + tx] += sbt_lo_term<F>(t + ty * SBT_SW, 1, X0, X, sw, ring);
        }
    }
    if (sw == 1 && tile_x == 0) {
        /* a row of ONE sample: "v[0] += v[s] >> 1" (sbt.c:199, :221) with v = the coefficient plane
         * itself (fwd_2d, :449-460): v[s] is column 1 of the row, the HL band the level that went
         * from two columns to one has already written (see sbt_inv_lift_tile) */
        PAR_FOR(k, SBT_SH) {
            int Y = Y0 + k;
            if (Y >= 0 && Y < sh) t[k * SBT_SW + SBT_HALO] += L.out_bands[Y * fw + 1] >> 1;
        }
    }
    DSVCU_SYNC();
    PAR_FOR(k, SBT_SH * SBT_SW) {
        int tx = k % SBT_SW;
        int X = X0 + tx;
        t[k] = (X & 1) ? sbt_fwd_scale_h<F>(t[k]) : sbt_fwd_scale_l<F>(t[k]);
    }
    DSVCU_SYNC();
    /* columns */
    PAR_FOR(k, (SBT_SH / 2) * SBT_SW) {
        int ty = (k / SBT_SW) * 2 + 1, tx = k % SBT_SW;
        int X = X0 + tx, Y = Y0 + ty;
        if (ty >= 1 && ty < SBT_SH - 1 && X >= 0 && X < sw && Y >= 0 && Y < sh) {
            t[ty * SBT_SW + tx] -= sbt_hi_term(t + tx, SBT_SW, Y0, Y, sh);
        }
    }
    DSVCU_SYNC();
    PAR_FOR(k, (SBT_SH / 2) * SBT_SW) {
        int ty = (k / SBT_SW) * 2, tx = k % SBT_SW;
        int X = X0 + tx, Y = Y0 + ty;
        if (ty >= 4 && ty < SBT_SH - 4 && X >= 0 && X < sw && Y >= 0 && Y < even_h) {
            int ring = 0;
            if (F == SBT_F_L2A && Y >= 2) {
                int pxk = (X & 1) ? cw + (X >> 1) : (X >> 1);
                int bc = (pxk * L.dbx) >> BLOCK_INTERP_P;
                int br = (((Y - 2) >> 1) * (2 * L.dby)) >> BLOCK_INTERP_P;
                ring = L.blockdata[bc + br * L.nbh] & BD_RINGING;
            }
            t[ty * SBT_SW + tx] += sbt_lo_term<F>(t + tx, SBT_SW, Y0, Y, sh, ring);
        }
    }
    DSVCU_SYNC();
    /* scale vertically + scatter the interior into the packed quadrants */
    PAR_FOR(k, SBT_TH * SBT_TW) {
        /* order threads so that consecutive threads write consecutive packed
         * columns: first the even X of a row, then the odd X */
        int ty = k / SBT_TW, j = k % SBT_TW;
        int tx = (j < SBT_TW / 2) ? j * 2 : (j - SBT_TW / 2) * 2 + 1;
        int X = X0 + SBT_HALO + tx, Y = Y0 + SBT_HALO + ty;
        if (X < sw && Y < sh) {
            int v = t[(ty + SBT_HALO) * SBT_SW + tx + SBT_HALO];
            v = (Y & 1) ? sbt_fwd_scale_h<F>(v) : sbt_fwd_scale_l<F>(v);
            int pxk = (X & 1) ? cw + (X >> 1) : (X >> 1);
            int pyk = (Y & 1) ? ch + (Y >> 1) : (Y >> 1);
            int32_t *o = (pxk < cw && pyk < ch) ? L.out_ll : L.out_bands;
            o[pyk * fw + pxk] = v;
        }
    }
}

/* ASF93 analysis, luma I-frame level 1 (sbt.c:389-421, 2-D driver :474-494).
 * FIR on unmodified input, so each dimension is out-of-place: t -> u -> t. */
DSVCU_HD int
sbt_asf_lo(const int32_t *v, int s, int base, int e, int n, int ring)
{
#define AT(i) v[(sbt_reflect((i), n - 1) - base) * s]
    if (ring) {
        return 46 * AT(e) + 20 * (AT(e - 1) + AT(e + 1)) - 9 * (AT(e - 2) + AT(e + 2)) -
               4 * (AT(e - 3) + AT(e + 3)) + 2 * (AT(e - 4) + AT(e + 4));
    }
    return 46 * AT(e) + 19 * (AT(e - 1) + AT(e + 1)) - 8 * (AT(e - 2) + AT(e + 2)) -
           3 * (AT(e - 3) + AT(e + 3)) + 1 * (AT(e - 4) + AT(e + 4));
#undef AT
}

/* one output sample of filterL1 at interleaved index i (even -> L, odd -> H),
 * already scaled; n is even */
DSVCU_HD int
sbt_asf_sample(const int32_t *v, int s, int base, int i, int n, int ring)
{
#define V(j) v[((j) - base) * s]
    if (i < 2) { /* first pair: simple lifting x2 / x4 (sbt.c:406-416) */
        int a1 = V(1) - ((V(0) + V(2) + 1) >> 1);
        if (i == 1) return a1 * 4;
        return (V(0) + (a1 >> 1)) * 2;
    }
    if (i >= n - 2) { /* last pair */
        int b3 = V(n - 3) - ((V(n - 4) + V(n - 2) + 1) >> 1);
        int b1 = V(n - 1) - V(n - 2);
        if (i == n - 1) return b1 * 4;
        return (V(n - 2) + ((b3 + b1 + 2) >> 2)) * 2;
    }
    if (i & 1) {
        int H = 32 * V(i) - 16 * (V(i - 1) + V(i + 1));
        return (H + 4) >> 3;
    }
    return (sbt_asf_lo(v, s, base, i, n, ring) + 16) >> 5;
#undef V
}

DSVCU_DEV void
sbt_fwd_l1_tile(const SbtLevel &L, int32_t *t, int tile_x, int tile_y)
{
    int32_t *u = t + SBT_SH * SBT_SW;
    const int X0 = tile_x * SBT_TW - SBT_HALO;
    const int Y0 = tile_y * SBT_TH - SBT_HALO;
    const int sw = L.sw, sh = L.sh, cw = L.cw, ch = L.ch, fw = L.fw;

    /* the horizontal FIR needs +-4 columns around every column of the tile
     * INCLUDING the halo columns used by nothing: only interior columns are
     * produced horizontally, but for all SBT_SH rows (vertical halo). */
    PAR_FOR(k, SBT_SH * SBT_SW) {
        int ty = k / SBT_SW, tx = k - ty * SBT_SW;
        int X = X0 + tx, Y = Y0 + ty;
        t[k] = (X >= 0 && X < sw && Y >= 0 && Y < sh) ? sbt_fwd_fetch(L, X, Y) : 0;
    }
    DSVCU_SYNC();
    PAR_FOR(k, SBT_SH * SBT_TW) {
        int ty = k / SBT_TW, tx = k % SBT_TW + SBT_HALO;
        int X = X0 + tx, Y = Y0 + ty;
        int v = 0;
        if (X < sw && Y >= 0 && Y < sh) {
            /* row pass: block row from the row, block column stepping 2*dbx per
             * output pair (sbt.c:392-405, :485-490) */
            int br = (Y * L.dby) >> BLOCK_INTERP_P;
            int bc = ((X >> 1) * (2 * L.dbx)) >> BLOCK_INTERP_P;
            int ring = 0;
            if (X >= 2 && X < sw - 2) {
                ring = L.blockdata[bc + br * L.nbh] & BD_RINGING;
            }
            v = sbt_asf_sample(t + ty * SBT_SW, 1, X0, X, sw, ring);
        }
        u[ty * SBT_SW + tx] = v;
    }
    DSVCU_SYNC();
    PAR_FOR(k, SBT_TH * SBT_TW) {
        int ty = k / SBT_TW, j = k % SBT_TW;
        int tx = (j < SBT_TW / 2) ? j * 2 : (j - SBT_TW / 2) * 2 + 1;
        int X = X0 + SBT_HALO + tx, Y = Y0 + SBT_HALO + ty;
        if (X < sw && Y < sh) {
            int pxk = (X & 1) ? cw + (X >> 1) : (X >> 1);
            int pyk = (Y & 1) ? ch + (Y >> 1) : (Y >> 1);
            /* column pass: block column from the PACKED column (sbt.c:491-494) */
            int bc = (pxk * L.dbx) >> BLOCK_INTERP_P;
            int br = ((Y >> 1) * (2 * L.dby)) >> BLOCK_INTERP_P;
            int ring = 0;
            if (Y >= 2 && Y < sh - 2) {
                ring = L.blockdata[bc + br * L.nbh] & BD_RINGING;
            }
            int v = sbt_asf_sample(u + tx + SBT_HALO, SBT_SW, Y0, Y, sh, ring);
            int32_t *o = (pxk < cw && pyk < ch) ? L.out_ll : L.out_bands;
            o[pyk * fw + pxk] = v;
        }
    }
}

/* ------------------------------------------------------------ Haar levels */

/* Inverse Haar, one thread per 2x2 output cell (sbt.c:615-795).  FILTERED
 * selects the LL-gradient-guided nudge of LH/HL (C.3.1.2). */
template <bool FILTERED>
DSVCU_DEV void
sbt_inv_haar_cells(const SbtLevel &L, int cta, int ncta)
{
    const int sw = L.sw, sh = L.sh, cw = L.cw, ch = L.ch, fw = L.fw;
    const int oddw = sw & 1, oddh = sh & 1;
    const int ncx = cw, ncy = ch; /* cells incl. the odd remainder column/row */
    const int total = ncx * ncy;
    const int ovf = L.ovf;
    for (int k = cta * DSVCU_NTH + DSVCU_TID; k < total; k += ncta * DSVCU_NTH) {
        int cy = k / ncx, cx = k - cy * ncx;
        int x = cx * 2, y = cy * 2;
        int fullx = (x < sw - oddw), fully = (y < sh - oddh);
        /* reads: LL region may alias the first LH/HL coefficient (see below) */
#define SRC(px_, py_) (((px_) < cw && (py_) < ch) ? L.ll[(py_) * fw + (px_)] : L.bands[(py_) * fw + (px_)])
        int LLv = L.ll[cy * fw + cx] * (1 << ovf);
        int LH = fullx ? L.bands[cy * fw + cw + cx] : 0;
        int HL = fully ? L.bands[(ch + cy) * fw + cx] : 0;
        int HH = (fullx && fully) ? L.bands[(ch + cy) * fw + cw + cx] : 0;
        if (FILTERED && fullx && fully) {
            int inX = x > 0 && x < (sw - oddw - 1);
            int inY = y > 0 && y < (sh - oddh - 1);
            int nudge, tt, lp, ln, mn, mx;
            if (inX) {
                /* spLL[idx + 1] can be the first LH coefficient when sw is
                 * even -- literal in-place read (sbt.c:723-724) */
                lp = SRC(cx - 1, cy) * (1 << ovf);
                ln = SRC(cx + 1, cy) * (1 << ovf);
                mx = LLv - ln;
                mn = lp - LLv;
                if (mn > mx) { tt = mn; mn = mx; mx = tt; }
                mx = min(mx, 0);
                mn = max(mn, 0);
                if (mx != mn) {
                    tt = sbt_round4(lp - ln);
                    nudge = sbt_round2(sbt_clamp(tt, mx, mn) - (LH * 2));
                    LH += sbt_clamp(nudge, -L.hqp, L.hqp);
                }
            }
            if (inY) {
                lp = SRC(cx, cy - 1) * (1 << ovf);
                ln = SRC(cx, cy + 1) * (1 << ovf);
                mx = LLv - ln;
                mn = lp - LLv;
                if (mn > mx) { tt = mn; mn = mx; mx = tt; }
                mx = min(mx, 0);
                mn = max(mn, 0);
                if (mx != mn) {
                    tt = sbt_round4(lp - ln);
                    nudge = sbt_round2(sbt_clamp(tt, mx, mn) - (HL * 2));
                    HL += sbt_clamp(nudge, -L.hqp, L.hqp);
                }
            }
        }
#undef SRC
        int o00, o01 = 0, o10 = 0, o11 = 0;
        if (fullx && fully) {
            o00 = (LLv + LH + HL + HH) / 4;
            o01 = (LLv - LH + HL - HH) / 4;
            o10 = (LLv + LH - HL - HH) / 4;
            o11 = (LLv - LH - HL + HH) / 4;
        } else if (fully) { /* odd last column */
            o00 = (LLv + HL) / 4;
            o10 = (LLv - HL) / 4;
        } else if (fullx) { /* odd last row */
            o00 = (LLv + LH) / 4;
            o01 = (LLv - LH) / 4;
        } else {
            o00 = LLv / 4;
        }
        if (L.px) {
            uint8_t *o = L.px + (size_t) y * L.px_stride + x;
            if (y < L.px_h) {
                if (x < L.px_w) o[0] = (uint8_t) sbt_clamp(o00 + 128, 0, 255);
                if (fullx && x + 1 < L.px_w) o[1] = (uint8_t) sbt_clamp(o01 + 128, 0, 255);
            }
            if (fully && y + 1 < L.px_h) {
                o += L.px_stride;
                if (x < L.px_w) o[0] = (uint8_t) sbt_clamp(o10 + 128, 0, 255);
                if (fullx && x + 1 < L.px_w) o[1] = (uint8_t) sbt_clamp(o11 + 128, 0, 255);
            }
        } else {
            int32_t *o = L.dst + y * fw + x;
            o[0] = o00;
            if (fullx) o[1] = o01;
            if (fully) {
                o[fw] = o10;
                if (fullx) o[fw + 1] = o11;
            }
        }
    }
}

/* Forward Haar (sbt.c:546-612) */
DSVCU_DEV void
sbt_fwd_haar_cells(const SbtLevel &L, int cta, int ncta)
{
    const int sw = L.sw, sh = L.sh, cw = L.cw, ch = L.ch, fw = L.fw;
    const int oddw = sw & 1, oddh = sh & 1;
    const int total = cw * ch;
    const int dv = L.ovf ? 2 : 1;
    for (int k = cta * DSVCU_NTH + DSVCU_TID; k < total; k += ncta * DSVCU_NTH) {
        int cy = k / cw, cx = k - cy * cw;
        int x = cx * 2, y = cy * 2;
        int fullx = (x < sw - oddw), fully = (y < sh - oddh);
        int x0 = sbt_fwd_fetch(L, x, y);
        int x1 = fullx ? sbt_fwd_fetch(L, x + 1, y) : 0;
        int x2 = fully ? sbt_fwd_fetch(L, x, y + 1) : 0;
        int x3 = (fullx && fully) ? sbt_fwd_fetch(L, x + 1, y + 1) : 0;
        if (fullx && fully) {
            L.out_ll[cy * fw + cx] = (x0 + x1 + x2 + x3) / dv;
            L.out_bands[cy * fw + cw + cx] = x0 - x1 + x2 - x3;
            L.out_bands[(ch + cy) * fw + cx] = x0 + x1 - x2 - x3;
            L.out_bands[(ch + cy) * fw + cw + cx] = x0 - x1 - x2 + x3;
        } else if (fully) {
            L.out_ll[cy * fw + cx] = 2 * (x0 + x2) / dv;
            L.out_bands[(ch + cy) * fw + cx] = 2 * (x0 - x2);
        } else if (fullx) {
            L.out_ll[cy * fw + cx] = 2 * (x0 + x1) / dv;
            L.out_bands[cy * fw + cw + cx] = 2 * (x0 - x1);
        } else {
            L.out_ll[cy * fw + cx] = (x0 * 4) / dv;
        }
    }
}

/* ------------------------------------------------------------ fused launches
 *
 * One launch transforms one level of up to three planes (their CTAs back to
 * back in the grid), or -- for the small top of the pyramid, where a level is
 * one or two tiles -- ALL remaining levels of each plane with one CTA per
 * plane, the LL image ping-ponging between the plane's two scratch buffers
 * behind a CTA barrier.  1080p: 6 launches per picture and direction instead
 * of 31. */
#define SBT_MAX_FUSED 10

struct SbtPlaneJob {
    SbtLevel L[SBT_MAX_FUSED]; /* in execution order */
    int f[SBT_MAX_FUSED];
    int nlev;                  /* > 1 only with ncta == 1 */
    int first_cta, ncta;
};

struct SbtJob {
    SbtPlaneJob p[3];
    int nplanes;
};

DSVCU_HD int
sbt_level_tiles(const SbtLevel &L)
{
    return ((L.sw + SBT_TW - 1) / SBT_TW) * ((L.sh + SBT_TH - 1) / SBT_TH);
}

#define SBT_TILE_LOOP(call)                                              \
    do {                                                                 \
        const int ntx_ = (L.sw + SBT_TW - 1) / SBT_TW;                   \
        const int nt_ = sbt_level_tiles(L);                              \
        for (int tile_ = cta; tile_ < nt_; tile_ += P.ncta) {            \
            call(L, t, tile_ % ntx_, tile_ / ntx_);                      \
            DSVCU_SYNC(); /* the tile buffer is reused */                \
        }                                                                \
    } while (0)

DSVCU_KERNEL void __launch_bounds__(SBT_THREADS)
k_sbt_inv(SbtJob J)
{
    DSVCU_SHARED int32_t t[SBT_SH * SBT_SW];
    int pl = 0;
    while (pl + 1 < J.nplanes && (int) blockIdx.x >= J.p[pl + 1].first_cta) pl++;
    const SbtPlaneJob &P = J.p[pl];
    const int cta = (int) blockIdx.x - P.first_cta;
    for (int lv = 0; lv < P.nlev; lv++) {
        const SbtLevel &L = P.L[lv];
        switch (P.f[lv]) {
            case SBT_F_LLI: SBT_TILE_LOOP(sbt_inv_lift_tile<SBT_F_LLI>); break;
            case SBT_F_LLP: SBT_TILE_LOOP(sbt_inv_lift_tile<SBT_F_LLP>); break;
            case SBT_F_CC: SBT_TILE_LOOP(sbt_inv_lift_tile<SBT_F_CC>); break;
            case SBT_F_L2A: SBT_TILE_LOOP(sbt_inv_lift_tile<SBT_F_L2A>); break;
            case SBT_F_L1: SBT_TILE_LOOP(sbt_inv_lift_tile<SBT_F_L1>); break;
            case SBT_F_LOSSLESS: SBT_TILE_LOOP(sbt_inv_lift_tile<SBT_F_LOSSLESS>); break;
            case SBT_F_HAAR: sbt_inv_haar_cells<true>(L, cta, P.ncta); break;
            default: sbt_inv_haar_cells<false>(L, cta, P.ncta); break;
        }
        if (lv + 1 < P.nlev) {
#ifndef DSVCU_EMU
            __threadfence_block();
#endif
            DSVCU_SYNC();
        }
    }
}

DSVCU_KERNEL void __launch_bounds__(SBT_THREADS)
k_sbt_fwd(SbtJob J)
{
    DSVCU_SHARED int32_t t[2 * SBT_SH * SBT_SW];
    int pl = 0;
    while (pl + 1 < J.nplanes && (int) blockIdx.x >= J.p[pl + 1].first_cta) pl++;
    const SbtPlaneJob &P = J.p[pl];
    const int cta = (int) blockIdx.x - P.first_cta;
    for (int lv = 0; lv < P.nlev; lv++) {
        const SbtLevel &L = P.L[lv];
        switch (P.f[lv]) {
            case SBT_F_LLI: SBT_TILE_LOOP(sbt_fwd_lift_tile<SBT_F_LLI>); break;
            case SBT_F_LLP: SBT_TILE_LOOP(sbt_fwd_lift_tile<SBT_F_LLP>); break;
            case SBT_F_CC: SBT_TILE_LOOP(sbt_fwd_lift_tile<SBT_F_CC>); break;
            case SBT_F_L2A: SBT_TILE_LOOP(sbt_fwd_lift_tile<SBT_F_L2A>); break;
            case SBT_F_L1: SBT_TILE_LOOP(sbt_fwd_l1_tile); break;
            case SBT_F_LOSSLESS: SBT_TILE_LOOP(sbt_fwd_lift_tile<SBT_F_LOSSLESS>); break;
            default: sbt_fwd_haar_cells(L, cta, P.ncta); break;
        }
        if (lv + 1 < P.nlev) {
#ifndef DSVCU_EMU
            __threadfence_block();
#endif
            DSVCU_SYNC();
        }
    }
}

#endif /* K_SBT_CUH */
