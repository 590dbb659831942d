/*
 * k_frame.cuh -- frame helpers the pixel path needs on the device.
 *
 * Replaces reference src/frame.c: extend_plane + downsample_strip (:250-410,
 * the 32-px border whose pixels are 4-sample averages of the nearest edge),
 * dsv_ds2x_frame_luma (:210-234).  dsv_frame_copy (:185-207) is a 2-D device
 * copy followed by k_extend.
 */
#ifndef K_FRAME_CUH
#define K_FRAME_CUH

#include "dsvcu_rt.h"

#define FR_BORDER 32

struct ExtPlane {
    uint8_t *data; /* pixel (0,0) */
    int stride, w, h;
};

struct ExtArgs {
    ExtPlane pl[3];
};

/* average of group g (4 samples, or the remainder group) along an edge:
 * downsample_strip, frame.c:250-355.  p = first sample, d = pitch, n = length */
DSVCU_DEV int
fr_strip(const uint8_t *p, int d, int n, int g)
{
    int len = n & ~3, rem = n & 3;
    if (g * 4 < len) {
        const uint8_t *q = p + (size_t) (g * 4) * d;
        return (q[0] + q[d] + q[2 * d] + q[3 * d] + 2) >> 2;
    }
    int sum = 0;
    for (int i = 0; i < rem; i++) {
        sum += p[(size_t) (len + i) * d];
    }
    return rem ? sum / rem : 0;
}

/* `n` copies of byte v at p (n = 32 or 4 here), with the widest stores the address allows:
 * planes start 16-byte aligned and strides are multiples of 16, so the left border and
 * every 4-column group are aligned; the right border is when the width is */
DSVCU_DEV void
fr_fill(uint8_t *p, int v, int n)
{
    const uint32_t w4 = (uint32_t) v * 0x01010101u;
    const uintptr_t a = (uintptr_t) p;
    if (n == FR_BORDER && (a & 15) == 0) {
        uint4 q;
        q.x = q.y = q.z = q.w = w4;
        ((uint4 *) p)[0] = q;
        ((uint4 *) p)[1] = q;
    } else if ((a & 3) == 0 && (n & 3) == 0) {
        for (int i = 0; i < n / 4; i++) ((uint32_t *) p)[i] = w4;
    } else {
        for (int i = 0; i < n; i++) p[i] = (uint8_t) v;
    }
}

/* border of one plane; work items: h rows (left+right), ceil(w/4) column groups
 * (top+bottom), 4 corners; item k of `total` is done by caller-chosen threads */
DSVCU_DEV void
fr_extend_item(const ExtPlane &P, int k)
{
    const int w = P.w, h = P.h, s = P.stride;
    const int ngc = (w + 3) / 4;
    if (k < h) {
        int j = k;
        int l = fr_strip(P.data, s, h, j / 4);
        int r = fr_strip(P.data + (w - 1), s, h, j / 4);
        uint8_t *line = P.data + (size_t) j * s;
        fr_fill(line - FR_BORDER, l, FR_BORDER);
        fr_fill(line + w, r, FR_BORDER);
    } else if (k < h + ngc) {
        int g = k - h;
        int t = fr_strip(P.data, 1, w, g);
        int b = fr_strip(P.data + (size_t) (h - 1) * s, 1, w, g);
        int x0 = g * 4, x1 = min(w, x0 + 4);
        for (int j = 0; j < FR_BORDER; j++) {
            fr_fill(P.data - (size_t) (j + 1) * s + x0, t, x1 - x0);
            fr_fill(P.data + (size_t) (h + j) * s + x0, b, x1 - x0);
        }
    } else {
        /* corners average the two adjacent strip ends (frame.c:377-380);
         * the right/bottom ends use the last FULL group */
        int cidx = k - h - ngc;
        int ts0 = fr_strip(P.data, 1, w, 0), ts1 = fr_strip(P.data, 1, w, w / 4 - 1);
        int bs0 = fr_strip(P.data + (size_t) (h - 1) * s, 1, w, 0);
        int bs1 = fr_strip(P.data + (size_t) (h - 1) * s, 1, w, w / 4 - 1);
        int ls0 = fr_strip(P.data, s, h, 0), ls1 = fr_strip(P.data, s, h, h / 4 - 1);
        int rs0 = fr_strip(P.data + (w - 1), s, h, 0), rs1 = fr_strip(P.data + (w - 1), s, h, h / 4 - 1);
        int v, cx, cy;
        if (cidx == 0) {
            v = (ts0 + ls0 + 1) >> 1; cx = -FR_BORDER; cy = -FR_BORDER;
        } else if (cidx == 1) {
            v = (ts1 + rs0 + 1) >> 1; cx = w; cy = -FR_BORDER;
        } else if (cidx == 2) {
            v = (ls1 + bs0 + 1) >> 1; cx = -FR_BORDER; cy = h;
        } else {
            v = (bs1 + rs1 + 1) >> 1; cx = w; cy = h;
        }
        for (int j = 0; j < FR_BORDER; j++) {
            fr_fill(P.data + (ptrdiff_t) (cy + j) * s + cx, v, FR_BORDER);
        }
    }
}

DSVCU_DEV int
fr_extend_items(const ExtPlane &P)
{
    return P.h + (P.w + 3) / 4 + 4;
}

DSVCU_KERNEL void __launch_bounds__(256)
k_extend(ExtArgs A)
{
    const ExtPlane P = A.pl[blockIdx.y];
    const int total = fr_extend_items(P);
    for (int k = (int) blockIdx.x * DSVCU_NTH + DSVCU_TID; k < total; k += (int) gridDim.x * DSVCU_NTH) {
        fr_extend_item(P, k);
    }
}

/* 2x2 box downsample of luma (frame.c:210-234); reads the source border when
 * the source height/width is odd */
DSVCU_KERNEL void __launch_bounds__(256)
k_ds2x(uint8_t *dst, int ds, int dw, int dh, const uint8_t *src, int ss)
{
    const int total = dw * dh;
    for (int k = (int) blockIdx.x * DSVCU_NTH + DSVCU_TID; k < total; k += (int) gridDim.x * DSVCU_NTH) {
        int j = k / dw, i = k - j * dw;
        const uint8_t *sp = src + (size_t) (2 * j) * ss + 2 * i;
        dst[(size_t) j * ds + i] = (uint8_t) ((sp[0] + sp[1] + sp[ss] + sp[ss + 1] + 2) >> 2);
    }
}

/* ---- border extension + the whole 2x luma pyramid in two launches -------------
 *
 * mk_pyramid (dsv_encoder.c:493-516) is "downsample, extend, repeat": ten small
 * launches per pyramid when done literally, and every picture needs two pyramids
 * (source, reconstruction).  A level only needs the BORDER of the level below
 * for its last row / column when that level's height / width is odd
 * (frame.c:210-234), so the work splits into
 *   k_pyr_interior  every level's pixels that depend on interior pixels only,
 *                   hierarchically from 64x64 tiles of the base picture
 *                   (one CTA per tile, levels kept in shared memory);
 *   k_pyr_borders   one CTA walks the levels in order: the at most one remaining
 *                   row + column of a level (from the level below and its
 *                   border), then that level's border -- a few thousand items
 *                   per level -- starting with the border of the base picture
 *                   (all its planes) when asked to.
 * Same arithmetic as k_ds2x / k_extend, bit-identical planes and borders. */
#define PYR_MAXLVL 5
#define PYR_TILE 64
#ifndef PYR_BORDER_THREADS
#define PYR_BORDER_THREADS 1024
#endif
struct PyrArgs {
    ExtPlane base[3]; /* [0] = luma = pyramid level 0 */
    int nbase;        /* planes of the base picture to extend first (0: already extended) */
    int ntiles;       /* k_pyr_interior: CTAs [0, ntiles) are tiles, the rest extend the base planes */
    int levels;
    ExtPlane lv[PYR_MAXLVL + 1]; /* [1..levels] */
};

/* extent of the pixels of level k that depend on interior pixels only */
DSVCU_DEV void
pyr_clean(const PyrArgs &A, int k, int *cw, int *ch)
{
    int w = A.base[0].w, h = A.base[0].h;
    for (int i = 0; i < k; i++) {
        w >>= 1;
        h >>= 1;
    }
    *cw = w;
    *ch = h;
}

DSVCU_KERNEL void __launch_bounds__(256)
k_pyr_interior(PyrArgs A)
{
    DSVCU_SHARED uint32_t tw[2][PYR_TILE * PYR_TILE / 4]; /* word storage: 4-byte aligned tile rows */
    uint8_t (*t)[PYR_TILE * PYR_TILE] = (uint8_t (*)[PYR_TILE * PYR_TILE]) tw;
    const ExtPlane &B = A.base[0];
    const int tiles_x = (B.w + PYR_TILE - 1) / PYR_TILE;
    if ((int) blockIdx.x >= A.ntiles) {
        /* the CTAs behind the tiles: border of the base picture's planes (reads interior
         * pixels, writes border pixels: independent of the tiles) */
        const int e = (int) blockIdx.x - A.ntiles, ne = (int) gridDim.x - A.ntiles;
        for (int p = 0; p < A.nbase; p++) {
            const int total = fr_extend_items(A.base[p]);
            for (int k = e * DSVCU_NTH + DSVCU_TID; k < total; k += ne * DSVCU_NTH) fr_extend_item(A.base[p], k);
        }
        return;
    }
    const int tx = (int) blockIdx.x % tiles_x, ty = (int) blockIdx.x / tiles_x;
    const int x0 = tx * PYR_TILE, y0 = ty * PYR_TILE;
    /* level 0 tile, word-wise (tile rows are 4-byte aligned: stride and the 32-px border are multiples of 4) */
    for (int k = DSVCU_TID; k < PYR_TILE * PYR_TILE / 4; k += DSVCU_NTH) {
        int j = k / (PYR_TILE / 4), i = (k - j * (PYR_TILE / 4)) * 4;
        uint32_t v = 0;
        if (y0 + j < B.h && x0 + i < B.w) {
            const uint8_t *p = B.data + (size_t) (y0 + j) * B.stride + x0 + i;
            if (x0 + i + 3 < B.w) {
                v = *(const uint32_t *) p;
            } else {
                for (int b = 0; b < 4 && x0 + i + b < B.w; b++) v |= (uint32_t) p[b] << (8 * b);
            }
        }
        *(uint32_t *) &t[0][j * PYR_TILE + i] = v;
    }
    DSVCU_SYNC();
    int side = PYR_TILE;
    for (int l = 1; l <= A.levels; l++) {
        const uint8_t *src = t[(l - 1) & 1];
        uint8_t *dst = t[l & 1];
        const ExtPlane &D = A.lv[l];
        int cw, ch;
        const int ox = x0 >> l, oy = y0 >> l;
        pyr_clean(A, l, &cw, &ch);
        side >>= 1;
        for (int k = DSVCU_TID; k < side * side; k += DSVCU_NTH) {
            int j = k / side, i = k - j * side;
            const uint8_t *sp = src + (2 * j) * (2 * side) + 2 * i;
            uint8_t v = (uint8_t) ((sp[0] + sp[1] + sp[2 * side] + sp[2 * side + 1] + 2) >> 2);
            dst[j * side + i] = v;
            if (ox + i < cw && oy + j < ch) D.data[(size_t) (oy + j) * D.stride + ox + i] = v;
        }
        DSVCU_SYNC();
    }
}

DSVCU_KERNEL void __launch_bounds__(PYR_BORDER_THREADS)
k_pyr_borders(PyrArgs A)
{
    /* (the base picture's border was written by the extra CTAs of k_pyr_interior) */
    for (int l = 1; l <= A.levels; l++) {
        const ExtPlane &S = (l == 1) ? A.base[0] : A.lv[l - 1];
        const ExtPlane &D = A.lv[l];
        int cw, ch;
        pyr_clean(A, l, &cw, &ch);
        /* what k_pyr_interior left out: columns [cw, w) over all rows, rows [ch, h) over the first cw columns */
        const int ncol = D.w - cw, nrow = D.h - ch;
        const int total = ncol * D.h + nrow * cw;
        PAR_FOR(k, total) {
            int i, j;
            if (k < ncol * D.h) {
                j = k / ncol;
                i = cw + (k - j * ncol);
            } else {
                int q = k - ncol * D.h;
                j = ch + q / cw;
                i = q - (q / cw) * cw;
            }
            const uint8_t *sp = S.data + (size_t) (2 * j) * S.stride + 2 * i;
            D.data[(size_t) j * D.stride + i] = (uint8_t) ((sp[0] + sp[1] + sp[S.stride] + sp[S.stride + 1] + 2) >> 2);
        }
        DSVCU_SYNC();
        {
            const int items = fr_extend_items(D);
            PAR_FOR(k, items) fr_extend_item(D, k);
        }
        DSVCU_SYNC();
    }
}

#endif /* K_FRAME_CUH */
