/*
 * dsv_mvutil.c -- motion-vector field helpers used by the host-side motion
 * (de)coder: neighbour predictor, neighbour difference, rate estimate.
 * Same arithmetic as reference src/dsv.c:324-459 (spec B.2.3.4); the device
 * twins of these live in csrc/ for the kernels that need them.
 */
#include <stdlib.h>
#include "dsv_host.h"
#include "dsv_mvutil_inl.h"

int
dsv_lb2(unsigned n)
{
    int l = 0;
    unsigned i = 1;
    while (i < n) {
        i <<= 1;
        l++;
    }
    return l;
}

void
dsv_movec_pred(DSV_MV *vecs, DSV_PARAMS *p, int x, int y, int *px, int *py)
{
    dsv_movec_pred_inl(vecs, p, x, y, px, py);
}

void
dsv_neighbordif2(DSV_MV *vecs, DSV_PARAMS *p, int x, int y, int *dx, int *dy)
{
    dsv_neighbordif2_inl(vecs, p, x, y, dx, dy);
}

int
dsv_neighbordif(DSV_MV *vecs, DSV_PARAMS *p, int x, int y)
{
    return dsv_neighbordif_inl(vecs, p, x, y);
}

static int
seg_len(int v) /* bit length of the SEG code of v */
{
    unsigned x;
    int nb = -1;
    if (v < 0) {
        v = -v;
    }
    v++;
    for (x = (unsigned) v; x; x >>= 1) {
        nb++;
    }
    return nb * 2 + 1 + (v ? 1 : 0);
}

int
dsv_mv_cost(DSV_MV *vecs, DSV_PARAMS *p, int i, int j, int mx, int my, int q, int sqr)
{
    int px, py, bits, b2sr;
    dsv_movec_pred(vecs, p, i, j, &px, &py);
    bits = seg_len(mx - px) + seg_len(my - py);
    b2sr = (256 * (q * q >> DSV_MAX_QP_BITS) * p->blk_w * p->blk_h) / (p->vidmeta->width * p->vidmeta->height);
    bits += bits * b2sr >> 7;
    return sqr ? bits * bits : bits;
}

/* resolution-dependent perceptual scale, 0 at CIF .. 128 at 1080p
 * (reference dsv_spatial_psy_factor, hzcc.c:66-86) */
int
dsv_spatial_psy_factor(DSV_PARAMS *p, int subband)
{
    int lo, hi, cur;
    int cif_h = DSV_UDIV_ROUND_UP(352, p->blk_w), cif_v = DSV_UDIV_ROUND_UP(288, p->blk_h);
    int fhd_h = DSV_UDIV_ROUND_UP(1920, p->blk_w), fhd_v = DSV_UDIV_ROUND_UP(1080, p->blk_h);
    if (subband == 1) {
        lo = cif_h;
        hi = fhd_h;
        cur = p->nblocks_h;
    } else if (subband == 2) {
        lo = cif_v;
        hi = fhd_v;
        cur = p->nblocks_v;
    } else {
        lo = cif_h * cif_v;
        hi = fhd_h * fhd_v;
        cur = p->nblocks_h * p->nblocks_v;
    }
    cur = MAX(0, cur - lo);
    return (cur << 7) / (hi - lo);
}

void
dsv_fmeta_from_params(dsvcu_fmeta *fm, const DSV_PARAMS *p, int isP, unsigned fnum)
{
    fm->isP = isP;
    fm->lossless = p->lossless;
    fm->do_psy = p->do_psy;
    fm->blk_w = p->blk_w;
    fm->blk_h = p->blk_h;
    fm->nblocks_h = p->nblocks_h;
    fm->nblocks_v = p->nblocks_v;
    fm->temporal_mc = p->temporal_mc;
    fm->inter_sharpen = p->vidmeta ? p->vidmeta->inter_sharpen : 0;
    fm->effort = p->effort;
    fm->fnum = fnum;
}
