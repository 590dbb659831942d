/*
 * dsv_mvutil_inl.h -- the per-block helpers of the motion (de)coder as inline functions:
 * neighbour predictor and neighbour difference (reference src/dsv.c:324-459, spec B.2.3.4).
 * dsv_mvutil.c exports them under their reference names; the decoder's side-information
 * loop (8160 blocks per 1080p picture, on a host thread that has a GPU to feed) includes
 * this header instead of paying four calls per block.
 */
#ifndef DSV_MVUTIL_INL_H
#define DSV_MVUTIL_INL_H

#include <stdlib.h>
#include "dsv_host.h"

/* pick whichever of left/top is closer to the gradient left + top - topleft */
static inline int
dsv_grad_pick(int left, int top, int topleft)
{
    int g = left + top - topleft;
    return (abs(g - left) < abs(g - top)) ? left : top;
}

static inline void
dsv_movec_pred_inl(DSV_MV *vecs, DSV_PARAMS *p, int x, int y, int *px, int *py)
{
    int lx = 0, ly = 0, tx = 0, ty = 0, dx = 0, dy = 0;
    DSV_MV *row = vecs + y * p->nblocks_h;
    if (x > 0) {
        lx = row[x - 1].u.mv.x;
        ly = row[x - 1].u.mv.y;
    }
    if (y > 0) {
        tx = row[x - p->nblocks_h].u.mv.x;
        ty = row[x - p->nblocks_h].u.mv.y;
        if (x > 0) {
            dx = row[x - 1 - p->nblocks_h].u.mv.x;
            dy = row[x - 1 - p->nblocks_h].u.mv.y;
        }
    }
    *px = dsv_grad_pick(lx, tx, dx);
    *py = dsv_grad_pick(ly, ty, dy);
}

static inline void
dsv_neighbordif2_inl(DSV_MV *vecs, DSV_PARAMS *p, int x, int y, int *dx, int *dy)
{
    DSV_MV *c = vecs + x + y * p->nblocks_h, *n;
    int cx = c->u.mv.x, cy = c->u.mv.y;
    int lx = cx, ly = cy, tx = cx, ty = cy;

    if (abs(cx) < 2 && abs(cy) < 2) {
        *dx = *dy = 0;
        return;
    }
    if (x > 0) {
        n = c - 1;
        if (n->u.all && !DSV_MV_IS_SKIP(n)) {
            lx = n->u.mv.x;
            ly = n->u.mv.y;
        }
    }
    if (y > 0) {
        n = c - p->nblocks_h;
        if (n->u.all && !DSV_MV_IS_SKIP(n)) {
            tx = n->u.mv.x;
            ty = n->u.mv.y;
        }
    }
    *dx = abs(lx - cx) + abs(ly - cy);
    *dy = abs(tx - cx) + abs(ty - cy);
}

static inline int
dsv_neighbordif_inl(DSV_MV *vecs, DSV_PARAMS *p, int x, int y)
{
    int a, b;
    dsv_neighbordif2_inl(vecs, p, x, y, &a, &b);
    return (a + b) / 3;
}

#endif /* DSV_MVUTIL_INL_H */
