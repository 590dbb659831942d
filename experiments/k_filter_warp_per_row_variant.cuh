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
 * pixels already modified by (p-1,q) and (p+1,q-1).  The
 * GPU schedule is a wavefront with slope 2: one warp owns one row of cells and
 * walks it left to right; before touching cell p of row q it waits until row
 * q-1 has published progress >= p+2.  Lanes split the four pixel lines of a
 * cell.  Progress counters live in global memory and are published with a
 * release fence; pixels are read with volatile loads so every read observes
 * what other SMs have published.
 */
#ifndef K_FILTER_CUH
#define K_FILTER_CUH

#include "dsvcu_rt.h"
#include "k_quant.cuh"

#define FILT_MODE_LUMA 0
#define FILT_MODE_INTRA 1
#define FILT_MODE_CHROMA 2
#define FILT_WARPS_PER_CTA 16

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
    int *progress;
    int mode;
    int cached; /* set per row by the kernel: tile loads may use L1 */
};

#ifdef DSVCU_EMU
#define PXLD(p) (*(p))
#define PROG_LD(p) (*(p))
#define FILT_LANE 0
#define FILT_NLANES 1
#define FILT_WARP ((int) blockIdx.x)
#define FILT_NWARPS ((int) gridDim.x)
#else
#define PXLD(p) (*(volatile const uint8_t *) (p))
#define PROG_LD(p) (*(volatile const int *) (p))
#define FILT_LANE ((int) (threadIdx.x & 31))
#define FILT_NLANES 32
#define FILT_WARP ((int) ((blockIdx.x * blockDim.x + threadIdx.x) >> 5))
#define FILT_NWARPS ((int) ((gridDim.x * blockDim.x) >> 5))
#endif

DSVCU_HD int f_abs(int v) { return v < 0 ? -v : v; }
DSVCU_HD int f_clamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

#define F_LPF(e0, i0, e1, i1) ((5 * ((e0) + (i0)) + 3 * ((e1) + (i1)) + 8) >> 4)
#define F_TEST(t, avg, e0, e1, e2, i0, i1, i2)                                         \
    (f_abs((e0) - (avg)) < (t) && f_abs((i0) - (avg)) < (t) && f_abs((e1) - (avg)) < (t) && \
     f_abs((i1) - (avg)) < (t) && f_abs((e2) - (avg)) < (t) && f_abs((i2) - (avg)) < (t))

/* one line of the edge filter: 11 samples p[-3..7] with pitch d (bmc.c:86-127) */
DSVCU_DEV void
f_edge_line(uint8_t *p, int d, int tE, int tM, int in_edge)
{
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

/* ihfilter4x4 (bmc.c:70-128): lanes split the rows */
DSVCU_DEV void
f_hfilter(const FiltArgs &A, int x, int y, int edge, int tE, int tM)
{
    if (x < 4 || x > A.w - 4 || (edge && tE <= 0) || tM <= 0) return;
    int top = f_clamp(y, 0, A.h - 1), bot = f_clamp(y + 4, 0, A.h - 1);
    int in_edge = x < (A.w - 8);
    if (!edge) tE = tM;
    for (int r = top + FILT_LANE; r < bot; r += FILT_NLANES) {
        f_edge_line(A.data + (size_t) r * A.stride + x, 1, tE, tM, in_edge);
    }
}

/* ivfilter4x4 (bmc.c:130-191): lanes split the columns */
DSVCU_DEV void
f_vfilter(const FiltArgs &A, int x, int y, int edge, int tE, int tM)
{
    if (y < 4 || y > A.h - 4 || (edge && tE <= 0) || tM <= 0) return;
    int beg = f_clamp(x, 0, A.w - 1), end = f_clamp(x + 4, 0, A.w - 1);
    int in_edge = y < (A.h - 8);
    if (!edge) tE = tM;
    for (int c = beg + FILT_LANE; c < end; c += FILT_NLANES) {
        f_edge_line(A.data + (size_t) y * A.stride + c, A.stride, tE, tM, in_edge);
    }
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

DSVCU_DEV FPrep
f_prep(const FiltArgs &A, int i, int j)
{
    FPrep P;
    P.mvxy = 0;
    P.bits = 0;
    P.nd = 0;
    P.active = 0;
    if (A.mode == FILT_MODE_CHROMA) {
        return P; /* chroma blocks keep the direct path */
    }
    const int nsbx = A.w / 4, nsby = A.h / 4;
    const int x = i * 4, y = j * 4;
    if (y + 4 >= A.h || x + 4 >= A.w) return P;
    const int fy = j * A.nbv / nsby, fx = i * A.nbh / nsbx;
    if (A.mode == FILT_MODE_INTRA) {
        int bd = A.blockdata[fx + fy * A.nbh];
        P.bits = bd << 24;
        P.active = !(bd & BD_RING);
        return P;
    }
    const dsvcu_mv mv = A.mvs[fx + fy * A.nbh];
    int ndx = 0, ndy = 0;
    P.mvxy = (mv.x & 0xffff) | ((int) mv.y << 16);
    P.bits = (int) (mv.flags & 255u) | ((int) mv.submask << 8) | (((x % A.blk_w) == 0) << 16) |
             (((x % (A.blk_w / 2)) == 0) << 17) | (((y % A.blk_h) == 0) << 18) | (((y % (A.blk_h / 2)) == 0) << 19);
    if (mv.flags & MVF_SKIP) return P;
    if (A.do_filter && !(mv.flags & MVF_INTRA)) {
        f_neighbordif2(A.mvs, A.nbh, fx, fy, &ndx, &ndy);
    }
    P.nd = (ndx & 0xffff) | (ndy << 16);
    P.active = (mv.flags & MVF_INTRA) || (A.do_filter && (ndx || ndy)) ||
               (A.sharpen && (mv.x & 3) && (mv.y & 3) && ((mv.x | mv.y) & 1) && f_abs(mv.x) < 8 && f_abs(mv.y) < 8);
    return P;
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
#define FT_BYTES (FT_ROWS * FT_S)
#define FT_ORG (3 * FT_S + 4)        /* tile offset of pixel (x, y) */

DSVCU_DEV void
f_tile_load(uint8_t *T, const FiltArgs &A, int x, int y)
{
    for (int k = FILT_LANE; k < FT_ROWS * 3; k += FILT_NLANES) {
        int r = k / 3, q = k - r * 3;
        const uint8_t *g = A.data + (ptrdiff_t) (y - 3 + r) * A.stride + x - 4 + 4 * q;
#ifndef DSVCU_EMU
        /* pixels a cell reads were last written by rows r-2 .. r of the same
         * picture.  For rows >= 2 inside a CTA all of them ran on this SM, so an
         * L1-cached load is coherent (after the block-scope fence of the
         * hand-off); the first two rows of a CTA read what another SM wrote and
         * go to L2. */
        *(uint32_t *) (T + r * FT_S + 4 * q) = A.cached ? *(const uint32_t *) g : *(volatile const uint32_t *) g;
#else
        memcpy(T + r * FT_S + 4 * q, g, 4);
#endif
    }
    DSVCU_SYNCWARP();
}

/* what: 1 = horizontal pass region, 2 = vertical pass region, 4 = the cell */
DSVCU_DEV void
f_tile_store(const uint8_t *T, const FiltArgs &A, int x, int y, int what)
{
    DSVCU_SYNCWARP();
    for (int k = FILT_LANE; k < 12 + 9; k += FILT_NLANES) {
        int r, q;
        if (k < 12) {
            if (!(what & 5)) continue;
            r = 3 + k / 3;
            q = k % 3;
            if (!(what & 1) && q != 1) continue; /* sharpener only: the cell's own word */
        } else {
            if (!(what & 2)) continue;
            r = 1 + (k - 12);
            q = 1;
            if ((what & 5) && r >= 3 && r < 7) continue; /* already written above */
        }
        uint8_t *g = A.data + (ptrdiff_t) (y - 3 + r) * A.stride + x - 4 + 4 * q;
#ifndef DSVCU_EMU
        *(uint32_t *) g = *(const uint32_t *) (T + r * FT_S + 4 * q);
#else
        memcpy(g, T + r * FT_S + 4 * q, 4);
#endif
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
f_luma_cell(const FiltArgs &G, uint8_t *T, int i, int j, const FPrep &P)
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
        f_tile_load(T, A, x, y);
        f_hfilter(L, x, y, teh, tH, tL);
        DSVCU_SYNCWARP();
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
        f_tile_load(T, A, x, y);
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
        DSVCU_SYNCWARP();
        if (sh > 2 * sv || amy > 2 * amx) {
            f_vfilter(L, x, y, tev, tt + addy, tt);
            what |= 2;
        } else if (sv > 2 * sh || amx > 2 * amy) {
            f_hfilter(L, x, y, teh, tt + addx, tt);
            what |= 1;
        } else {
            f_hfilter(L, x, y, teh, tt + addx, tt);
            DSVCU_SYNCWARP();
            f_vfilter(L, x, y, tev, tt + addy, tt);
            what |= 3;
        }
        DSVCU_SYNCWARP();
        touched = 1;
    }
    if (A.sharpen && (mv.x & 3) && (mv.y & 3) && ((mv.x | mv.y) & 1) && amx < 8 && amy < 8) {
        if (!(what & 8)) f_tile_load(T, A, x, y);
        if (FILT_LANE == 0) f_degrad(dxy, FT_S);
        what |= 4;
        touched = 1;
    }
    if (what & 7) f_tile_store(T, A, x, y, what & 7);
    return touched;
}

/* one 4x4 cell of dsv_intra_filter (bmc.c:411-455) */
DSVCU_DEV int
f_intra_cell(const FiltArgs &G, uint8_t *T, int i, int j, const FPrep &P)
{
    const FiltArgs &A = G;
    const int x = i * 4, y = j * 4;
    const int flags = (P.bits >> 24) & 255;
    const int q = A.q;
    uint8_t *dxy = T + FT_ORG;
    const FiltArgs L = f_tile_view(A, T, x, y);
    int sh, sv, shl, svl, tt = 32;
    F4x4 b;
    f_tile_load(T, A, x, y);
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
    DSVCU_SYNCWARP();
    f_hfilter(L, x, y, 0, tt, tt);
    DSVCU_SYNCWARP();
    f_vfilter(L, x, y, 0, tt, tt);
    DSVCU_SYNCWARP();
    tt = (sh > sv) ? (3 * sh + sv) : (3 * sv + sh);
    tt = f_curve_tex(tt);
    tt = 16 + ((tt + 2) >> 2);
    tt = (tt * q) >> 12;
    tt = f_clamp(tt, 0, A.fthresh);
    f_hfilter(L, x, y, 0, tt, tt);
    DSVCU_SYNCWARP();
    f_vfilter(L, x, y, 0, tt, tt);
    f_tile_store(T, A, x, y, 3);
    return 1;
}

/* one motion block of chroma_filter (bmc.c:620-657) */
DSVCU_DEV int
f_chroma_cell(const FiltArgs &A, int i, int j)
{
    const dsvcu_mv mv = A.mvs[i + j * A.nbh];
    if (mv.flags & MVF_SKIP) return 0;
    const int x = i * A.bw, y = j * A.bh;
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
    /* left side: the bh/4 row groups are independent, lanes take lines */
    if (!(x < 4 || x > A.w - 4 || tx <= 0)) {
        int in_edge = x < (A.w - 8);
        for (int k = FILT_LANE; k < A.bh; k += FILT_NLANES) {
            int z = k & ~3;
            if (y + z + 4 < A.h) {
                int top = f_clamp(y + z, 0, A.h - 1), bot = f_clamp(y + z + 4, 0, A.h - 1);
                int r = y + k;
                if (r >= top && r < bot) {
                    f_edge_line(A.data + (size_t) r * A.stride + x, 1, tx, tx, in_edge);
                }
            }
        }
    }
    DSVCU_SYNCWARP();
    if (!(y < 4 || y > A.h - 4 || ty <= 0)) {
        int in_edge = y < (A.h - 8);
        for (int k = FILT_LANE; k < A.bw; k += FILT_NLANES) {
            int z = k & ~3;
            if (x + z + 4 < A.w) {
                int beg = f_clamp(x + z, 0, A.w - 1), end = f_clamp(x + z + 4, 0, A.w - 1);
                int c = x + k;
                if (c >= beg && c < end) {
                    f_edge_line(A.data + (size_t) y * A.stride + c, A.stride, ty, ty, in_edge);
                }
            }
        }
    }
    DSVCU_SYNCWARP();
    return (tx > 0) || (ty > 0);
}

/* Can cell (i, row) touch pixels at all?  Decided from block data only (vectors,
 * flags), never from pixels, so it can be evaluated ahead of the wavefront.
 * Cells that cannot are skipped without waiting for their neighbours. */
DSVCU_DEV int
f_cell_active(const FiltArgs &A, int i, int j)
{
    if (A.mode == FILT_MODE_CHROMA) {
        const dsvcu_mv mv = A.mvs[i + j * A.nbh];
        if (mv.flags & MVF_SKIP) return 0;
        if (mv.flags & MVF_INTRA) return 1;
        int ndx, ndy;
        f_neighbordif2(A.mvs, A.nbh, i, j, &ndx, &ndy);
        if (ndx < f_abs(mv.y) && ndy < f_abs(mv.x)) return 0;
        return ((min(ndy, 64) * A.q) >> 12) > 0 || ((min(ndx, 64) * A.q) >> 12) > 0;
    }
    const int nsbx = A.w / 4, nsby = A.h / 4;
    const int x = i * 4, y = j * 4;
    if (y + 4 >= A.h || x + 4 >= A.w) return 0;
    const int fy = j * A.nbv / nsby, fx = i * A.nbh / nsbx;
    if (A.mode == FILT_MODE_INTRA) {
        return !(A.blockdata[fx + fy * A.nbh] & BD_RING);
    }
    const dsvcu_mv mv = A.mvs[fx + fy * A.nbh];
    if (mv.flags & MVF_SKIP) return 0;
    if (mv.flags & MVF_INTRA) return 1;
    if (A.do_filter) {
        int ndx, ndy;
        f_neighbordif2(A.mvs, A.nbh, fx, fy, &ndx, &ndy);
        if (ndx || ndy) return 1;
    }
    return A.sharpen && (mv.x & 3) && (mv.y & 3) && ((mv.x | mv.y) & 1) && f_abs(mv.x) < 8 && f_abs(mv.y) < 8;
}

/* Slope-2 wavefront over rows of cells (see file header).  Protocol: row r
 * publishes progress P = "cells < P of this row are complete, and row r-1 has
 * completed cells < P+1" (the second half covers the footprint overlap between
 * rows r-1 and r+1 in the same column).  An active cell i waits for row r-1 to
 * reach min(i+2, ncols).  Inactive cells are not visited one by one: the warp
 * finds the next active cell with a ballot and, while it waits for that cell's
 * dependency, keeps relaying the progress of the row above (minus one cell), so
 * a region without filtering costs one flag round trip per row instead of one
 * per cell.
 *
 * A CTA owns FILT_WARPS_PER_CTA consecutive rows, one warp each.  Hand-offs
 * between rows of the same CTA go through shared-memory progress words and
 * block-scope fences (tens of cycles); only the last row of a CTA also
 * publishes to global memory with a device-scope fence, for the first row of
 * the next CTA.  Pixels always travel through L2 (volatile accesses). */
#ifndef DSVCU_EMU
#define FILT_FENCE(dev)                  \
    do {                                 \
        if (dev) {                       \
            __threadfence();             \
        } else {                         \
            __threadfence_block();       \
        }                                \
    } while (0)
#endif

DSVCU_DEV void
f_row(const FiltArgs &A0, uint8_t *T, int row, volatile int *sprog, int lr)
{
    FiltArgs A = A0;
    const int ncols = A.ncols;
    A.cached = (lr >= 2);
#ifndef DSVCU_EMU
    const int lane = FILT_LANE;
    /* where the progress of the row above lives, and who needs ours */
    const bool above_global = (lr == 0);
    const bool pub_global = (lr == FILT_WARPS_PER_CTA - 1);
    const bool dev_fence = (lr >= FILT_WARPS_PER_CTA - 2); /* rows whose pixels the next CTA reads */
    volatile const int *above = above_global ? (volatile const int *) (A.progress + row - 1) : (sprog + lr - 1);
    int seen = (row == 0) ? 0x7fffffff : 0; /* progress of the row above, cached */
    int published = 0;
#define F_PUBLISH(v)                                                      \
    do {                                                                  \
        published = (v);                                                  \
        if (lane == 0) {                                                  \
            sprog[lr] = published;                                        \
            if (pub_global) *(volatile int *) (A.progress + row) = published; \
        }                                                                 \
    } while (0)
#else
    (void) sprog;
    (void) lr;
#endif
    for (int base = 0; base < ncols; base += FILT_NLANES) {
#ifndef DSVCU_EMU
        int cell = base + lane;
        FPrep mine;
        bool act;
        if (A.mode == FILT_MODE_CHROMA) {
            mine = f_prep(A, 0, 0);
            act = cell < ncols && f_cell_active(A, cell, row);
        } else {
            mine = f_prep(A, min(cell, ncols - 1), row);
            act = cell < ncols && mine.active;
        }
        unsigned mask = __ballot_sync(0xffffffffu, act);
#else
        FPrep P = f_prep(A, min(base, ncols - 1), row);
        unsigned mask = (base < ncols && (A.mode == FILT_MODE_CHROMA ? f_cell_active(A, base, row) : P.active)) ? 1u : 0u;
#endif
        while (mask) {
#ifndef DSVCU_EMU
            const int src = __ffs(mask) - 1;
            int i = base + src;
            int need = min(i + 2, ncols);
            FPrep P;
            P.mvxy = __shfl_sync(0xffffffffu, mine.mvxy, src);
            P.bits = __shfl_sync(0xffffffffu, mine.bits, src);
            P.nd = __shfl_sync(0xffffffffu, mine.nd, src);
            P.active = 1;
            mask &= mask - 1;
            while (seen < need) {
                seen = *above;
                int relay = min(i, seen >= ncols ? i : seen - 1);
                if (relay > published) F_PUBLISH(relay);
                if (seen < need && above_global) __nanosleep(32);
            }
            FILT_FENCE(above_global);
            if (i > published) F_PUBLISH(i);
#else
            int i = base;
            mask = 0;
#endif
            if (A.mode == FILT_MODE_LUMA) {
                f_luma_cell(A, T, i, row, P);
            } else if (A.mode == FILT_MODE_INTRA) {
                f_intra_cell(A, T, i, row, P);
            } else {
                f_chroma_cell(A, i, row);
            }
#ifndef DSVCU_EMU
            FILT_FENCE(dev_fence);
            __syncwarp();
            F_PUBLISH(i + 1);
#endif
        }
    }
#ifndef DSVCU_EMU
    /* tail without active cells: relay until the row above is done */
    while (seen < ncols) {
        seen = *above;
        int relay = seen >= ncols ? ncols : seen - 1;
        if (relay > published) F_PUBLISH(relay);
        if (seen < ncols && above_global) __nanosleep(32);
    }
    FILT_FENCE(dev_fence);
    F_PUBLISH(ncols);
#undef F_PUBLISH
#endif
}

DSVCU_KERNEL void __launch_bounds__(FILT_WARPS_PER_CTA * 32)
k_filter_wavefront(FiltArgs A)
{
    __align__(16) DSVCU_SHARED uint8_t tiles[FILT_WARPS_PER_CTA][FT_BYTES];
    DSVCU_SHARED int sprog[FILT_WARPS_PER_CTA];
#ifndef DSVCU_EMU
    const int lr = (int) (threadIdx.x >> 5);
    const int row = (int) blockIdx.x * FILT_WARPS_PER_CTA + lr;
    if (threadIdx.x < FILT_WARPS_PER_CTA) sprog[threadIdx.x] = 0;
    __syncthreads();
    if (row < A.nrows) f_row(A, tiles[lr], row, sprog, lr);
#else
    for (int lr = 0; lr < FILT_WARPS_PER_CTA; lr++) {
        int row = (int) blockIdx.x * FILT_WARPS_PER_CTA + lr;
        if (row < A.nrows) f_row(A, tiles[0], row, sprog, lr);
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
