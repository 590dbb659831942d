/*
 * k_hme_body.cuh -- the per-block functions of the motion search (see k_hme.cuh),
 * written against a GROUP of ME_G lanes that work on one block together:
 *   ME_LANE   lane inside the group, ME_NL = ME_G lanes per group
 *   me_wsum.. reductions over the group, ME_SYNC() group barrier
 * k_hme.cuh includes this file twice, inside namespaces meg32 (one warp per block:
 * the dependent wavefront pass, where latency per block matters) and meg8 (four
 * blocks per warp: the neighbour-independent prepass, where only the instruction
 * count matters -- the scalar control code of four blocks issues once, and the
 * 8x8 / 4x4 sub-block loops fill their lanes).  The host emulation runs both with
 * a single lane.
 */
#undef ME_LANE
#undef ME_NL
#undef ME_GMASK
#undef ME_GBASE
#undef ME_SYNC
#ifdef DSVCU_EMU
#define ME_LANE 0
#define ME_NL 1
#define ME_GMASK 0xffffffffu
#define ME_GBASE 0
#define ME_SYNC() ((void) 0)
DSVCU_DEV int me_wsum(int v) { return v; }
DSVCU_DEV unsigned me_wsumu(unsigned v) { return v; }
DSVCU_DEV int me_wor(int v) { return v; }
DSVCU_DEV unsigned me_wmaxu(unsigned v) { return v; }
DSVCU_DEV unsigned me_wminu(unsigned v) { return v; }
DSVCU_DEV unsigned me_gballot(int p) { return p ? 1u : 0u; } /* bit l = lane l of the group */
DSVCU_DEV int me_gbcast(int v, int lane) { (void) lane; return v; }
#else
#define ME_LANE ((int) (threadIdx.x & (ME_G - 1)))
#define ME_NL ME_G
/* first lane of the group inside its warp, and the group's lane mask */
#define ME_GBASE ((int) (threadIdx.x & 31u & ~(unsigned) (ME_G - 1)))
#define ME_GMASK (ME_G == 32 ? 0xffffffffu : (((1u << (ME_G & 31)) - 1u) << ME_GBASE))
#define ME_SYNC() __syncwarp(ME_GMASK)
DSVCU_DEV int me_wsum(int v) { return __reduce_add_sync(ME_GMASK, v); }
DSVCU_DEV unsigned me_wsumu(unsigned v) { return __reduce_add_sync(ME_GMASK, v); }
DSVCU_DEV int me_wor(int v) { return (int) __reduce_or_sync(ME_GMASK, (unsigned) v); }
DSVCU_DEV unsigned me_wmaxu(unsigned v) { return __reduce_max_sync(ME_GMASK, v); }
DSVCU_DEV unsigned me_wminu(unsigned v) { return __reduce_min_sync(ME_GMASK, v); }
DSVCU_DEV unsigned me_gballot(int p) { return __ballot_sync(ME_GMASK, p) >> ME_GBASE; } /* bit l = lane l of the group */
DSVCU_DEV int me_gbcast(int v, int lane) { return __shfl_sync(ME_GMASK, v, lane, ME_G); }
#endif

struct MePsy {
    int err_w, tex_w, avg_w;
};

/* one 2x2 cell of the psycho-visual metric (METR_CALC, hme.c:126-134) */
DSVCU_DEV unsigned
me_cell(int a1, int a2, int a3, int a4, int b1, int b2, int b3, int b4, const MePsy &p)
{
    int s0 = (int) me_uavg4(a1, a2, a3, a4), s1 = (int) me_uavg4(b1, b2, b3, b4);
    int se = (int) me_uavg4(me_abs(a1 - b1), me_abs(a2 - b2), me_abs(a3 - b3), me_abs(a4 - b4));
    int ta = (int) me_uavg4(me_abs(a1 - a2), me_abs(a2 - a3), me_abs(a3 - a4), me_abs(a4 - a1));
    int tb = (int) me_uavg4(me_abs(b1 - b2), me_abs(b2 - b3), me_abs(b3 - b4), me_abs(b4 - b1));
    unsigned acc = 0;
    acc += (unsigned) (me_sqr(se) << p.err_w);
    acc += (unsigned) (me_sqr(ta - tb) << p.tex_w);
    acc += (unsigned) (me_sqr(s0 - s1) << p.avg_w);
    return acc;
}

/* the same on packed cells A = (a1,a2,a3,a4), B = (b1,b2,b3,b4) */
DSVCU_DEV unsigned
me_cell4(uint32_t A, uint32_t B, const MePsy &p)
{
    int s0 = (int) ((me_dot4(A, ME_ONES, 2)) >> 2), s1 = (int) ((me_dot4(B, ME_ONES, 2)) >> 2);
    int se = (int) ((me_dot4(me_absdiff4(A, B), ME_ONES, 2)) >> 2);
    int ta = (int) ((me_dot4(me_absdiff4(A, me_perm(A, A, 0x0321)), ME_ONES, 2)) >> 2);
    int tb = (int) ((me_dot4(me_absdiff4(B, me_perm(B, B, 0x0321)), ME_ONES, 2)) >> 2);
    unsigned acc = (unsigned) (me_sqr(se) << p.err_w);
    acc += (unsigned) (me_sqr(ta - tb) << p.tex_w);
    acc += (unsigned) (me_sqr(s0 - s1) << p.avg_w);
    return acc;
}

/* raw accumulator of the psy metric over w x h (umetr_wxh, hme.c:191-196) */
DSVCU_DEV unsigned
me_umetr(const uint8_t *a, int as, const uint8_t *b, int bs, int w, int h, const MePsy &p)
{
    int cw = w / 2, ch = h / 2, n = cw * ch;
    unsigned acc = 0;
    const int gs = me_gshift(w);
    if (gs >= 0) {
        /* one work item = 4 pixels x 2 rows = two 2x2 cells */
        const int ng = ch << gs, gm = (1 << gs) - 1;
        /* the source block is word-aligned at every level whose block origin is
         * a multiple of 4 (always at level 0): one load instead of two per word */
        const bool al = ((((uintptr_t) a) | (unsigned) as) & 3) == 0;
        for (int g = ME_LANE; g < ng; g += ME_NL) {
            int y = (g >> gs) * 2, x = (g & gm) * 4;
            uint32_t a0 = al ? me_ld4a(a + y * as + x) : me_ld4(a + y * as + x);
            uint32_t a1 = al ? me_ld4a(a + (y + 1) * as + x) : me_ld4(a + (y + 1) * as + x);
            uint32_t b0 = me_ld4(b + y * bs + x), b1 = me_ld4(b + (y + 1) * bs + x);
            acc += me_cell4(me_perm(a0, a1, 0x5410), me_perm(b0, b1, 0x5410), p);
            acc += me_cell4(me_perm(a0, a1, 0x7632), me_perm(b0, b1, 0x7632), p);
        }
        return me_wsumu(acc);
    }
    for (int k = ME_LANE; k < n; k += ME_NL) {
        int j = k / cw, i = k - j * cw;
        const uint8_t *pa = a + (2 * j) * as + 2 * i, *pb = b + (2 * j) * bs + 2 * i;
        acc += me_cell(pa[0], pa[1], pa[as], pa[as + 1], pb[0], pb[1], pb[bs], pb[bs + 1], p);
    }
    return me_wsumu(acc);
}

/* fastmetr (hme.c:271-306): sqrt-normalised */
DSVCU_DEV unsigned
me_metr(const uint8_t *a, int as, const uint8_t *b, int bs, int w, int h, const MePsy &p)
{
    if (w == 0 || h == 0) return 0x7fffffffu;
    unsigned acc = me_umetr(a, as, b, bs, w, h, p);
    return me_isqrt(acc) * (unsigned) w * (unsigned) h / (unsigned) me_avg2(w, h);
}

/* SSE over w x h (hme.c:198-242) */
DSVCU_DEV unsigned
me_sse(const uint8_t *a, int as, const uint8_t *b, int bs, int w, int h)
{
    if (w == 0 || h == 0) return 0x7fffffffu;
    unsigned acc = 0;
    int n = w * h;
    const int gs = me_gshift(w);
    if (gs >= 0) {
        const int ng = h << gs, gm = (1 << gs) - 1;
        const bool al = ((((uintptr_t) a) | (unsigned) as) & 3) == 0;
        for (int g = ME_LANE; g < ng; g += ME_NL) {
            int y = g >> gs, x = (g & gm) * 4;
            uint32_t sa = al ? me_ld4a(a + y * as + x) : me_ld4(a + y * as + x);
            uint32_t d = me_absdiff4(sa, me_ld4(b + y * bs + x));
            acc = me_dot4(d, d, acc);
        }
        return me_wsumu(acc);
    }
    for (int k = ME_LANE; k < n; k += ME_NL) {
        int j = k / w, i = k - j * w;
        int d = (int) a[j * as + i] - (int) b[j * bs + i];
        acc += (unsigned) (d * d);
    }
    return me_wsumu(acc);
}

DSVCU_DEV unsigned
me_hier_metr(int level, const uint8_t *a, int as, const uint8_t *b, int bs, int w, int h, const MePsy &p)
{
    if (level > 1) return me_sse(a, as, b, bs, w, h);
    return me_metr(a, as, b, bs, w, h, p);
}

/* ---- MV field helpers (dsv.c:324-447) ---- */

/* entries of the level being built are written by other warps: bypass L1 */
DSVCU_DEV void
me_ldmv(const dsvcu_mv *p, int *x, int *y, unsigned *fl)
{
#ifndef DSVCU_EMU
    int w = *(volatile const int *) p;
    *x = (int16_t) (w & 0xffff);
    *y = (int16_t) (w >> 16);
    if (fl) *fl = *((volatile const unsigned *) p + 1);
#else
    *x = p->x;
    *y = p->y;
    if (fl) *fl = p->flags;
#endif
}

DSVCU_DEV int
me_grad_pick(int left, int top, int topleft)
{
    int g = left + top - topleft;
    return (me_abs(g - left) < me_abs(g - top)) ? left : top;
}

DSVCU_DEV void
me_movec_pred(const dsvcu_mv *vecs, int nbh, int x, int y, int *px, int *py)
{
    int lx = 0, ly = 0, tx = 0, ty = 0, dx = 0, dy = 0;
    const dsvcu_mv *c = vecs + x + y * nbh;
    if (x > 0) {
        me_ldmv(c - 1, &lx, &ly, NULL);
    }
    if (y > 0) {
        me_ldmv(c - nbh, &tx, &ty, NULL);
        if (x > 0) {
            me_ldmv(c - nbh - 1, &dx, &dy, NULL);
        }
    }
    *px = me_grad_pick(lx, tx, dx);
    *py = me_grad_pick(ly, ty, dy);
}

DSVCU_DEV int
me_seg_len(int v)
{
    /* 2 * floor(log2(|v| + 1)) + 2 */
    if (v < 0) v = -v;
    v++;
#ifndef DSVCU_EMU
    return (31 - __clz(v)) * 2 + 2;
#else
    return (31 - __builtin_clz((unsigned) v)) * 2 + 2;
#endif
}

/* mv_cost (hme.c:354-366) on top of dsv_mv_cost (dsv.c:357-374) */
/* the predictor of a block depends only on its left / top / top-left
 * neighbours, which are final before the block starts: computed once per block
 * (MePred) instead of inside every rate term */
struct MePred {
    int x, y;
};

DSVCU_DEV int
me_mv_cost(const MeArgs &A, const MePred &pr, int mx, int my, int level)
{
    int px = pr.x, py = pr.y, bits, b2sr, q = A.quant;
    int sqr = level > 1;
    bits = me_seg_len(mx - px) + me_seg_len(my - py);
    b2sr = A.b2sr;
    bits += bits * b2sr >> 7;
    if (sqr) bits *= bits;
    bits = min(bits, 1 << 19);
    if (sqr) return bits * (q * q >> 12) >> (12 - 2);
    return 3 * bits * q >> 12;
}

DSVCU_DEV int
me_neighbordif(const dsvcu_mv *vecs, int nbh, int x, int y)
{
    const dsvcu_mv *c = vecs + x + y * nbh;
    int cx, cy, nx, ny;
    unsigned nf;
    me_ldmv(c, &cx, &cy, NULL);
    int lx = cx, ly = cy, tx = cx, ty = cy;
    if (me_abs(cx) < 2 && me_abs(cy) < 2) return 0;
    if (x > 0) {
        me_ldmv(c - 1, &nx, &ny, &nf);
        if ((nx | ny) != 0 && !(nf & MVF_SKIP)) {
            lx = nx;
            ly = ny;
        }
    }
    if (y > 0) {
        me_ldmv(c - nbh, &nx, &ny, &nf);
        if ((nx | ny) != 0 && !(nf & MVF_SKIP)) {
            tx = nx;
            ty = ny;
        }
    }
    return (me_abs(lx - cx) + me_abs(ly - cy) + me_abs(tx - cx) + me_abs(ty - cy)) / 3;
}

/* ---- block statistics (hme.c:492-775); every lane returns the same value ---- */

DSVCU_DEV int
me_block_avg(const uint8_t *a, int as, int w, int h)
{
    int s = 0, n = w * h;
    const int gs = me_gshift(w);
    if (gs >= 0) {
        const int ng = h << gs, gm = (1 << gs) - 1;
        unsigned u = 0;
        for (int g = ME_LANE; g < ng; g += ME_NL) {
            u = me_dot4(me_ld4(a + (g >> gs) * as + (g & gm) * 4), ME_ONES, u);
        }
        return me_wsum((int) u) / (w * h);
    }
    for (int k = ME_LANE; k < n; k += ME_NL) {
        int j = k / w, i = k - j * w;
        s += a[j * as + i];
    }
    return me_wsum(s) / (w * h);
}

/* sums of horizontal / vertical absolute gradients (block_tex core) */
DSVCU_DEV void
me_grad_sums(const uint8_t *a, int as, int w, int h, unsigned *psh, unsigned *psv, int *psum)
{
    unsigned sh = 0, sv = 0;
    int s = 0, n = w * h;
    const int gs = me_gshift(w);
    if (gs >= 0) {
        const int ng = h << gs, gm = (1 << gs) - 1;
        unsigned us = 0;
        for (int g = ME_LANE; g < ng; g += ME_NL) {
            int y = g >> gs, x = (g & gm) * 4;
            const uint8_t *p = a + y * as + x;
            uint32_t c = me_ld4(p);
            uint32_t dl = me_absdiff4(c, me_ld4(p - 1));
            us = me_dot4(c, ME_ONES, us);
            if (x == 0) dl &= 0xffffff00u; /* column 0 has no left neighbour */
            sh = me_dot4(dl, ME_ONES, sh);
            if (y > 0) sv = me_dot4(me_absdiff4(c, me_ld4(p - as)), ME_ONES, sv);
        }
        *psh = me_wsumu(sh);
        *psv = me_wsumu(sv);
        *psum = me_wsum((int) us);
        return;
    }
    for (int k = ME_LANE; k < n; k += ME_NL) {
        int j = k / w, i = k - j * w;
        int px = a[j * as + i];
        s += px;
        if (i > 0) sh += (unsigned) me_abs(px - a[j * as + i - 1]);
        if (j > 0) sv += (unsigned) me_abs(px - a[(j - 1) * as + i]);
    }
    *psh = me_wsumu(sh);
    *psv = me_wsumu(sv);
    *psum = me_wsum(s);
}

DSVCU_DEV unsigned
me_block_tex(const uint8_t *a, int as, int w, int h)
{
    unsigned sh, sv;
    int s;
    me_grad_sums(a, as, w, h, &sh, &sv, &s);
    return max(sh, sv);
}

DSVCU_DEV int
me_abs_dev(const uint8_t *a, int as, int w, int h, int mean)
{
    int var = 0, n = w * h;
    const int gs = me_gshift(w);
    if (gs >= 0 && mean >= 0 && mean <= 255) {
        const int ng = h << gs, gm = (1 << gs) - 1;
        const uint32_t m4 = (uint32_t) mean * ME_ONES;
        unsigned u = 0;
        for (int g = ME_LANE; g < ng; g += ME_NL) {
            u = me_dot4(me_absdiff4(me_ld4(a + (g >> gs) * as + (g & gm) * 4), m4), ME_ONES, u);
        }
        return me_wsum((int) u);
    }
    for (int k = ME_LANE; k < n; k += ME_NL) {
        int j = k / w, i = k - j * w;
        var += me_abs((int) a[j * as + i] - mean);
    }
    return me_wsum(var);
}

DSVCU_DEV int
me_block_var(const uint8_t *a, int as, int w, int h, unsigned *avg)
{
    int s = me_block_avg(a, as, w, h);
    *avg = (unsigned) s;
    return me_abs_dev(a, as, w, h, s);
}

DSVCU_DEV int
me_block_detail(const uint8_t *a, int as, int w, int h, unsigned *avg)
{
    unsigned sh, sv;
    int s, var, tex;
    me_grad_sums(a, as, w, h, &sh, &sv, &s);
    s /= (w * h);
    *avg = (unsigned) s;
    var = me_abs_dev(a, as, w, h, s) >> 1;
    tex = (int) max(sh, sv) - var;
    return var + max(tex, 0);
}

DSVCU_DEV int
me_quant_tex(const uint8_t *a, int as, int w, int h)
{
    unsigned sh = 0, sv = 0;
    int n = w * h;
    const int gs = me_gshift(w);
    if (gs >= 0) {
        const int ng = h << gs, gm = (1 << gs) - 1;
        for (int g = ME_LANE; g < ng; g += ME_NL) {
            int y = g >> gs, x = (g & gm) * 4;
            const uint8_t *p = a + y * as + x;
            uint32_t c = (me_ld4(p) >> 4) & 0x0f0f0f0fu;
            uint32_t dr = me_absdiff4(c, (me_ld4(p + 1) >> 4) & 0x0f0f0f0fu);
            if (x == w - 4) dr &= 0x00ffffffu; /* last column has no right neighbour */
            sh = me_dot4(dr, dr, sh);
            if (y > 0) {
                uint32_t du = me_absdiff4(c, (me_ld4(p - as) >> 4) & 0x0f0f0f0fu);
                sv = me_dot4(du, du, sv);
            }
        }
        sh = me_wsumu(sh);
        sv = me_wsumu(sv);
        return (int) (me_isqrt(max(sh, sv)) / (unsigned) me_avg2(w, h));
    }
    for (int k = ME_LANE; k < n; k += ME_NL) {
        int j = k / w, i = k - j * w;
        int px = a[j * as + i] >> 4;
        if (i < w - 1) {
            int d = px - (a[j * as + i + 1] >> 4);
            sh += (unsigned) (d * d);
        }
        if (j > 0) {
            int d = px - (a[(j - 1) * as + i] >> 4);
            sv += (unsigned) (d * d);
        }
    }
    sh = me_wsumu(sh);
    sv = me_wsumu(sv);
    return (int) (me_isqrt(max(sh, sv)) / (unsigned) me_avg2(w, h));
}

/* 16-bin histograms are built in per-warp shared memory */
DSVCU_DEV void
me_hist_clear(int *hist)
{
    for (int k = ME_LANE; k < 16; k += ME_NL) hist[k] = 0;
    ME_SYNC();
}

DSVCU_DEV unsigned
me_block_hist_var(const uint8_t *a, int as, int w, int h, int *hist)
{
    unsigned avg, quant16, var = 0;
    int n = w * h;
    const int gs = me_gshift(w);
    me_hist_clear(hist);
    avg = (unsigned) me_block_avg(a, as, w, h);
    if (avg == 0) avg = 1;
    quant16 = ((1u << 3) << 16) / avg;
    if (gs >= 0) {
        const int ng = h << gs, gm = (1 << gs) - 1;
        for (int g = ME_LANE; g < ng; g += ME_NL) {
            uint32_t c = me_ld4(a + (g >> gs) * as + (g & gm) * 4);
            for (int k = 0; k < 4; k++) {
                unsigned hi = ((c >> (8 * k)) & 255u) * quant16 >> 16;
                atomicAdd(&hist[hi > 15 ? 15 : hi], 1);
            }
        }
    } else {
        for (int k = ME_LANE; k < n; k += ME_NL) {
            int j = k / w, i = k - j * w;
            int hi = (int) (a[j * as + i] * quant16 >> 16);
            atomicAdd(&hist[hi < 0 ? 0 : (hi > 15 ? 15 : hi)], 1);
        }
    }
    ME_SYNC();
    avg = 0;
    for (int x = 0; x < 16; x++) avg += (unsigned) hist[x];
    avg /= 16;
    for (int x = 0; x < 16; x++) var += ((unsigned) hist[x] - avg) * ((unsigned) hist[x] - avg);
    ME_SYNC();
    return (var * 16 * 16) / (unsigned) (16 * w * h * w * h);
}

DSVCU_DEV int
me_block_peaks(const uint8_t *a, int as, int w, int h, int *hist, int bavg)
{
    int avg = bavg, maxv = 0, npeaks = 0, quant16, cw, ch, n;
    me_hist_clear(hist);
    if (avg == 0) avg = 1;
    quant16 = ((1 << 3) << 16) / avg;
    cw = w / 2;
    ch = h / 2;
    n = cw * ch;
    {
        const int gs = me_gshift(w);
        if (gs >= 0) {
            const int ng = ch << gs, gm = (1 << gs) - 1;
            for (int g = ME_LANE; g < ng; g += ME_NL) {
                int y = (g >> gs) * 2, x = (g & gm) * 4;
                uint32_t r0 = me_ld4(a + y * as + x), r1 = me_ld4(a + (y + 1) * as + x);
                int d0 = (int) (me_dot4(me_perm(r0, r1, 0x5410), ME_ONES, 2) >> 2);
                int d1 = (int) (me_dot4(me_perm(r0, r1, 0x7632), ME_ONES, 2) >> 2);
                atomicAdd(&hist[min(d0 * quant16 >> 16, 15)], 1);
                atomicAdd(&hist[min(d1 * quant16 >> 16, 15)], 1);
            }
        } else {
            for (int k = ME_LANE; k < n; k += ME_NL) {
                int j = k / cw, i = k - j * cw;
                const uint8_t *p = a + (2 * j) * as + 2 * i;
                int ds = (int) me_uavg4(p[0], p[1], p[as], p[as + 1]);
                int hi = ds * quant16 >> 16;
                atomicAdd(&hist[min(hi, 15)], 1);
            }
        }
    }
    ME_SYNC();
    avg = 0;
    for (int x = 0; x < 16; x++) {
        maxv = max(maxv, hist[x]);
        avg += hist[x];
    }
    avg /= 16;
    maxv >>= 2;
    for (int x = 0; x < 16; x++) {
        int c = hist[x], is_peak = 1;
        if (x > 0) is_peak &= (c > hist[x - 1]);
        if (x < 15) is_peak &= (c > hist[x + 1]);
        is_peak &= (c > maxv) || (c > avg);
        npeaks += is_peak;
    }
    ME_SYNC();
    return npeaks;
}

DSVCU_DEV void
me_c_average(const MePlane *pl, int x, int y, int w, int h, int *uavg, int *vavg)
{
    int su = 0, sv = 0, n = w * h;
    const int gs = me_gshift(w);
    if (gs >= 0) {
        const int ng = h << gs, gm = (1 << gs) - 1;
        unsigned uu = 0, uv = 0;
        for (int g = ME_LANE; g < ng; g += ME_NL) {
            int j = g >> gs, i = (g & gm) * 4;
            uu = me_dot4(me_ld4(pl[1].data + (y + j) * pl[1].stride + x + i), ME_ONES, uu);
            uv = me_dot4(me_ld4(pl[2].data + (y + j) * pl[2].stride + x + i), ME_ONES, uv);
        }
        su = (int) uu;
        sv = (int) uv;
    } else {
        for (int k = ME_LANE; k < n; k += ME_NL) {
            int j = k / w, i = k - j * w;
            su += pl[1].data[(y + j) * pl[1].stride + x + i];
            sv += pl[2].data[(y + j) * pl[2].stride + x + i];
        }
    }
    su = me_wsum(su);
    sv = me_wsum(sv);
    /* w*h == 0 divides by zero in the reference too; callers never pass it */
    *uavg = su / (w * h);
    *vavg = sv / (w * h);
}

struct MeChroma {
    int nature, hifreq, greyish, skinnish;
};

DSVCU_DEV void
me_chroma_analysis(MeChroma *c, int y, int u, int v)
{
    c->nature = u < 128 && v < 160;
    c->greyish = me_abs(u - 128) < 8 && me_abs(v - 128) < 8;
    c->skinnish = (y > 80) && (y < 230) && me_abs(u - 108) < 24 && me_abs(v - 148) < 24;
    c->hifreq = (u > 160) && !c->greyish && !c->skinnish;
}

DSVCU_DEV int
me_invalid_block(int fw, int fh, int bx, int by, int bw, int bh, int pad)
{
    return (bx - pad) < -ME_BORDER || (by - pad) < -ME_BORDER || (bx + bw + pad) >= (fw + ME_BORDER) ||
           (by + bh + pad) >= (fh + ME_BORDER);
}

/* max over the four quadrants of the raw psy metric, luma + both chroma planes
 * (yuv_max_subblock_err, hme.c:368-411) */
DSVCU_DEV void
me_yuv_max_sub(unsigned out[3], const MePlane *sp, const MePlane *rp, int bx, int by, int brx, int bry, int bw, int bh,
               int cbx, int cby, int cbrx, int cbry, int cbw, int cbh, const MePsy &psy)
{
    bw /= 2;
    bh /= 2;
    cbw /= 2;
    cbh /= 2;
    ME_CNT(MEC_MAXSUB);
    for (int z = 0; z < 3; z++) {
        unsigned sub[4] = { 0, 0, 0, 0 };
        int pos = 0;
        for (int g = 0; g <= bh; g += (bh + !bh)) {
            for (int f = 0; f <= bw; f += (bw + !bw)) {
                const uint8_t *s = sp[z].data + (by + g) * sp[z].stride + bx + f;
                const uint8_t *r = rp[z].data + (bry + g) * rp[z].stride + brx + f;
                if (pos < 4) sub[pos] = me_umetr(s, sp[z].stride, r, rp[z].stride, bw, bh, psy);
                pos++;
            }
        }
        bx = cbx;
        by = cby;
        brx = cbrx;
        bry = cbry;
        bw = cbw;
        bh = cbh;
        out[z] = max(max(sub[0], sub[1]), max(sub[2], sub[3]));
    }
}

/* calc_EPRM (hme.c:452-490): does MV / intra(ref avg) / intra(src avg)
 * prediction clip anywhere in the block?  OR over pixels == the early-out scan */
DSVCU_DEV void
me_calc_eprm(const uint8_t *src, int ss, const uint8_t *mvr, int rs, int avg_src, int avg_ref, int w, int h, int *eprmi,
             int *eprmd, int *eprmr)
{
    int ci = 0, cd = 0, cr = 0, n = w * h;
    avg_src -= 128;
    avg_ref -= 128;
    {
        const int gs = me_gshift(w);
        if (gs >= 0) {
            const int ng = h << gs, gm = (1 << gs) - 1;
            for (int g = ME_LANE; g < ng; g += ME_NL) {
                int j = g >> gs, i = (g & gm) * 4;
                uint32_t sw = me_ld4(src + j * ss + i), rw = me_ld4(mvr + j * rs + i);
                for (int k = 0; k < 4; k++) {
                    int s = (int) ((sw >> (8 * k)) & 255u);
                    cr |= ((s - (int) ((rw >> (8 * k)) & 255u)) + 128) & ~0xff;
                    ci |= (s - avg_ref) & ~0xff;
                    cd |= (s - avg_src) & ~0xff;
                }
            }
        } else {
            for (int k = ME_LANE; k < n; k += ME_NL) {
                int j = k / w, i = k - j * w;
                int s = src[j * ss + i];
                cr |= ((s - (int) mvr[j * rs + i]) + 128) & ~0xff;
                ci |= (s - avg_ref) & ~0xff;
                cd |= (s - avg_src) & ~0xff;
            }
        }
    }
    *eprmi = me_wor(ci != 0);
    *eprmd = me_wor(cd != 0);
    *eprmr = me_wor(cr != 0);
}

/* ---- sub-pel refinement (hme.c:777-837, :1051-1164) ---- */


/* Half-pel image (34 x 34, HP_STRIDE) of the 17 x 17 window at r, as the
 * reference's hpel() builds it (hme.c:787-813).  The reference then expands it
 * to a 68 x 68 quarter-pel image by bilinear averaging (qpel(), :815-837) of
 * which the search samples 7 x 256 points; here those points are averaged from
 * the half-pel image on the fly (me_qsample), same arithmetic. */
DSVCU_DEV void
me_interp(uint8_t *tmph, uint8_t *win, int16_t *hbuf, const uint8_t *r, int rs)
{
    /* stage the full-pel window once: ME_WIN = 20 bytes per row = five (unaligned) words */
    for (int k = ME_LANE; k < ME_WIN * (ME_WIN / 4); k += ME_NL) {
        int j = k / (ME_WIN / 4), i = (k - j * (ME_WIN / 4)) * 4;
        *(uint32_t *) (win + j * ME_WIN + i) = me_ld4(r + (j - 1) * rs + i - 1);
    }
    ME_SYNC();
    /* horizontal half-pel sums for rows -1 .. SP_DIM+1 */
    for (int k = ME_LANE; k < ME_WIN * SP_DIM; k += ME_NL) {
        int j = k / SP_DIM, i = k - j * SP_DIM;
        const uint8_t *p = win + j * ME_WIN + i + 1;
        hbuf[k] = (int16_t) ME_HPF(p[-1], p[0], p[1], p[2]);
    }
    ME_SYNC();
    for (int k = ME_LANE; k < SP_DIM * SP_DIM; k += ME_NL) {
        int j = k / SP_DIM, i = k - j * SP_DIM;
        const uint8_t *p = win + (j + 1) * ME_WIN + i + 1;
        uint8_t *d = tmph + (2 * j) * HP_STRIDE + 2 * i;
        int c = ME_HPF(hbuf[k], hbuf[k + SP_DIM], hbuf[k + 2 * SP_DIM], hbuf[k + 3 * SP_DIM]);
        d[0] = p[0];
        d[1] = (uint8_t) me_u8((ME_HPF(p[-1], p[0], p[1], p[2]) + 4) >> 3);
        d[HP_STRIDE] = (uint8_t) me_u8((ME_HPF(p[-ME_WIN], p[0], p[ME_WIN], p[2 * ME_WIN]) + 4) >> 3);
        d[HP_STRIDE + 1] = (uint8_t) me_u8((c + 32) >> 6);
    }
    ME_SYNC();
}

/* quarter-pel sample (qx, qy) of the image the reference's qpel() would build:
 * a, avg2(a,b), avg2(a,c) or avg4(a,b,c,e) by the parity of (qx, qy).  All four
 * cases are (a + h[ox] + h[oy*S] + h[ox + oy*S] + 2) >> 2 with ox, oy the parity
 * bits ((2a+2b+2)>>2 == (a+b+1)>>1), so the sample is branch-free. */
DSVCU_DEV int
me_qsample(const uint8_t *tmph, int qx, int qy)
{
    const uint8_t *h0 = tmph + (qy >> 1) * HP_STRIDE + (qx >> 1);
    int ox = qx & 1, oy = (qy & 1) * HP_STRIDE;
    return (h0[0] + h0[ox] + h0[oy] + h0[ox + oy] + 2) >> 2;
}

/* me_qsample at a position whose parity bits are known: ox in {0, 1}, oy in {0, HP_STRIDE}.
 * Both even = the half-pel sample itself ((4a + 2) >> 2 == a), one odd = the average of two
 * ((2a + 2b + 2) >> 2 == (a + b + 1) >> 1): one, two or four loads instead of always four. */
DSVCU_DEV int
me_qs(const uint8_t *h0, int ox, int oy)
{
    if (!(ox | oy)) return h0[0];
    if (!oy) return (h0[0] + h0[1] + 1) >> 1;
    if (!ox) return (h0[0] + h0[oy] + 1) >> 1;
    return (h0[0] + h0[1] + h0[oy] + h0[1 + oy] + 2) >> 2;
}

/* psy metric of the 16 x 16 source window against the quarter-pel image at
 * offset (tx, ty) quarter pels (qpsad, hme.c:244-269) */
DSVCU_DEV unsigned
me_qpsad(const uint8_t *a, int as, const uint8_t *tmph, int tx, int ty, const MePsy &psy)
{
    unsigned acc = 0;
    for (int g = ME_LANE; g < (SP_SZ / 2) * (SP_SZ / 4); g += ME_NL) {
        /* one work item = 4 source pixels x 2 rows = two cells */
        int y = (g >> 2) * 2, x = (g & 3) * 4;
        uint32_t a0 = me_ld4(a + y * as + x), a1 = me_ld4(a + (y + 1) * as + x);
        int qx = 4 + tx + 4 * x, qy = 4 + ty + 4 * y;
        uint32_t B0 = (uint32_t) me_qsample(tmph, qx, qy) | ((uint32_t) me_qsample(tmph, qx + 4, qy) << 8) |
                      ((uint32_t) me_qsample(tmph, qx, qy + 4) << 16) | ((uint32_t) me_qsample(tmph, qx + 4, qy + 4) << 24);
        uint32_t B1 = (uint32_t) me_qsample(tmph, qx + 8, qy) | ((uint32_t) me_qsample(tmph, qx + 12, qy) << 8) |
                      ((uint32_t) me_qsample(tmph, qx + 8, qy + 4) << 16) | ((uint32_t) me_qsample(tmph, qx + 12, qy + 4) << 24);
        acc += me_cell4(me_perm(a0, a1, 0x5410), B0, psy);
        acc += me_cell4(me_perm(a0, a1, 0x7632), B1, psy);
    }
    acc = me_wsumu(acc);
    return me_isqrt(acc) * (unsigned) SP_SZ * (unsigned) SP_SZ / (unsigned) SP_SZ;
}

/* cell metric with the source-side terms (mean s0, texture ta) precomputed */
DSVCU_DEV unsigned
me_cell4_pre(uint32_t A, int s0, int ta, uint32_t B, const MePsy &p)
{
    int s1 = (int) ((me_dot4(B, ME_ONES, 2)) >> 2);
    int se = (int) ((me_dot4(me_absdiff4(A, B), ME_ONES, 2)) >> 2);
    int tb = (int) ((me_dot4(me_absdiff4(B, me_perm(B, B, 0x0321)), ME_ONES, 2)) >> 2);
    unsigned acc = (unsigned) (me_sqr(se) << p.err_w);
    acc += (unsigned) (me_sqr(ta - tb) << p.tex_w);
    acc += (unsigned) (me_sqr(s0 - s1) << p.avg_w);
    return acc;
}

/* me_qpsad for up to 7 offsets in one pass over the source window: the source
 * cells are loaded (and their own terms computed) once, and the offsets give
 * independent accumulation chains */
DSVCU_DEV void
me_qpsad_multi(const uint8_t *a, int as, const uint8_t *tmph, int nv, const signed char *tx, const signed char *ty, const MePsy &psy,
               unsigned *out)
{
    unsigned acc[ME_MAXSP];
    for (int v = 0; v < ME_MAXSP; v++) acc[v] = 0;
    for (int g = ME_LANE; g < (SP_SZ / 2) * (SP_SZ / 4); g += ME_NL) {
        int y = (g >> 2) * 2, x = (g & 3) * 4;
        uint32_t a0 = me_ld4(a + y * as + x), a1 = me_ld4(a + (y + 1) * as + x);
        uint32_t A0 = me_perm(a0, a1, 0x5410), A1 = me_perm(a0, a1, 0x7632);
        int s00 = (int) (me_dot4(A0, ME_ONES, 2) >> 2), s01 = (int) (me_dot4(A1, ME_ONES, 2) >> 2);
        int ta0 = (int) (me_dot4(me_absdiff4(A0, me_perm(A0, A0, 0x0321)), ME_ONES, 2) >> 2);
        int ta1 = (int) (me_dot4(me_absdiff4(A1, me_perm(A1, A1, 0x0321)), ME_ONES, 2) >> 2);
#ifndef DSVCU_EMU
#pragma unroll
#endif
        for (int v = 0; v < ME_MAXSP; v++) {
            if (v < nv) {
                /* the eight samples of this work item: quarter-pel positions 4 apart = half-pel
                 * image positions 2 apart, all with the offset's parity (uniform over the lanes) */
                const int qx = 4 + tx[v] + 4 * x, qy = 4 + ty[v] + 4 * y;
                const int ox = qx & 1, oy = (qy & 1) * HP_STRIDE;
                const uint8_t *h0 = tmph + (qy >> 1) * HP_STRIDE + (qx >> 1), *h1 = h0 + 2 * HP_STRIDE;
                uint32_t B0 = (uint32_t) me_qs(h0, ox, oy) | ((uint32_t) me_qs(h0 + 2, ox, oy) << 8) |
                              ((uint32_t) me_qs(h1, ox, oy) << 16) | ((uint32_t) me_qs(h1 + 2, ox, oy) << 24);
                uint32_t B1 = (uint32_t) me_qs(h0 + 4, ox, oy) | ((uint32_t) me_qs(h0 + 6, ox, oy) << 8) |
                              ((uint32_t) me_qs(h1 + 4, ox, oy) << 16) | ((uint32_t) me_qs(h1 + 6, ox, oy) << 24);
                acc[v] += me_cell4_pre(A0, s00, ta0, B0, psy) + me_cell4_pre(A1, s01, ta1, B1, psy);
            }
        }
    }
    for (int v = 0; v < ME_MAXSP; v++) {
        if (v < nv) out[v] = me_isqrt(me_wsumu(acc[v])) * (unsigned) SP_SZ * (unsigned) SP_SZ / (unsigned) SP_SZ;
    }
}

/* scratch of one sub-pel measurement (half-pel image and its staging) */
struct __align__(16) MeInterp {
    uint8_t tmph[(2 + HP_STRIDE) * (2 + HP_STRIDE)];
    uint8_t win[ME_WIN * ME_WIN + 16];
    int16_t hbuf[(SP_DIM + 3) * SP_DIM + 4];
};

struct MeScratch {
    MeInterp *ip; /* NULL where no sub-pel measurement can happen (the prepass) */
    int hist[16];
    /* full-pel metric memo of the current block: position -> raw metric */
    short memo_x[ME_MEMO], memo_y[ME_MEMO];
    unsigned memo_v[ME_MEMO];
    /* work counters of the current block (lane 0): full-block metric evaluations, sub-pel position metrics */
    int n_evals, n_subpel;
};

/* Sub-pel refinement, split in two.  me_subpel_measure: everything that depends
 * only on the full-pel position -- the four neighbour SSEs that order the
 * search (hme.c:1084-1136), the half-pel image and the metric at the <= 7 test
 * offsets in the reference's order (:1137-1160).  me_subpel_decide: the scalar
 * part that needs the block's running best score and rate predictor. */
struct MeSubpel {
    int nv;
    signed char tx[ME_MAXSP + 1], ty[ME_MAXSP + 1];
    unsigned sc[ME_MAXSP];
};

DSVCU_DEV void
me_subpel_measure(const MeArgs &A, MeScratch *S, MeSubpel *M, int fpelx, int fpely, int bx, int by, int bw, int bh,
                  const MePsy &psy)
{
    const MePlane &sp = A.src[0], &rp = A.ref[0];
    unsigned quad[4], ms1, ms2;
    int pri[2], sec[2], diag[2], xx, yy, nv = 0;
    ME_CNT(MEC_SUBPEL);
    const int ddx[4] = { 1, -1, 0, 0 }, ddy[4] = { 0, 0, 1, -1 };
    {
        const uint8_t *s = sp.data + by * sp.stride + bx;
        for (int n = 0; n < 4; n++) {
            const uint8_t *r = rp.data + (by + fpely + ddy[n]) * rp.stride + bx + fpelx + ddx[n];
            quad[n] = me_sse(s, sp.stride, r, rp.stride, bw, bh);
        }
    }
    xx = bx + ((bw >> 1) - ((SP_SZ + 1) / 2));
    yy = by + ((bh >> 1) - ((SP_SZ + 1) / 2));
    me_interp(S->ip->tmph, S->ip->win, S->ip->hbuf, rp.data + (yy + fpely - 1) * rp.stride + xx + fpelx - 1, rp.stride);

    pri[0] = 0; pri[1] = -1;
    sec[0] = -1; sec[1] = 0;
    ms1 = quad[1];
    ms2 = quad[3];
    if (quad[3] >= quad[2]) {
        pri[0] = 0; pri[1] = 1;
        ms2 = quad[2];
    }
    if (quad[1] >= quad[0]) {
        sec[0] = 1; sec[1] = 0;
        ms1 = quad[0];
    }
    if (ms2 > ms1) {
        int t0 = sec[0], t1 = sec[1];
        sec[0] = pri[0]; sec[1] = pri[1];
        pri[0] = t0; pri[1] = t1;
    }
    diag[0] = pri[0] + sec[0];
    diag[1] = pri[1] + sec[1];
    /* test order of the reference: half then quarter steps along pri, sec,
     * diag, then pri + diag */
    for (int n = 0; n <= 6; n++) {
        int t[2];
        if (n == 6) {
            t[0] = pri[0] + diag[0];
            t[1] = pri[1] + diag[1];
        } else {
            int hp = !(n & 1);
            const int *tv = (n >> 1) == 0 ? pri : ((n >> 1) == 1 ? sec : diag);
            t[0] = tv[0] * (1 << hp);
            t[1] = tv[1] * (1 << hp);
        }
        if (((t[0] | t[1]) & 1) && A.effort < 8) continue;
        M->tx[nv] = (signed char) t[0];
        M->ty[nv] = (signed char) t[1];
        nv++;
    }
    M->nv = nv;
    if (ME_LANE == 0) {
        S->n_evals += 4;     /* the four neighbour SSEs that order the search */
        S->n_subpel += nv;
    }
    me_qpsad_multi(sp.data + yy * sp.stride + xx, sp.stride, S->ip->tmph, nv, M->tx, M->ty, psy, M->sc);
}

/* The reference walks the offsets in order and keeps a strictly better score,
 * i.e. it ends on the FIRST offset that attains the minimum, if that minimum
 * beats the running best: one offset per lane, a min reduction and a ballot. */
DSVCU_DEV unsigned
me_subpel_decide(const MeArgs &A, int nv, const signed char *tx, const signed char *ty, const unsigned *sc, int *outx,
                 int *outy, int fpelx, int fpely, const MePred &pr, unsigned best, int bw, int bh)
{
    int yarea = bw * bh, bestx = 0, besty = 0;
    int area_ratio = 8 * (SP_SZ * SP_SZ) / yarea, iarea_ratio = 8 * yarea / (SP_SZ * SP_SZ);
    best = best * (unsigned) area_ratio >> 3;
    for (int base = 0; base < nv; base += ME_NL) {
        const int n = base + ME_LANE;
        int mx = 0, my = 0;
        unsigned score = 0xffffffffu;
        if (n < nv) {
            mx = tx[n];
            my = ty[n];
            score = sc[n] + (unsigned) me_mv_cost(A, pr, fpelx * 4 + mx, fpely * 4 + my, 0);
        }
        const unsigned m = me_wminu(score);
        if (best > m) {
            const int w = (int) me_ctz(me_gballot(n < nv && score == m));
            best = m;
            bestx = me_gbcast(mx, w);
            besty = me_gbcast(my, w);
        }
    }
    *outx = bestx;
    *outy = besty;
    return best * (unsigned) iarea_ratio >> 3;
}

/* ---- intra sub-block tests (hme.c:839-1049) ---- */

/* ae_out / rest_out (optional, 8-wide sub-blocks only): the per-cell mean absolute
 * error and the ratio-independent part of the inter error, see MePre::qi_ae */
DSVCU_DEV void
me_err_intra(const uint8_t *a, int as, const uint8_t *b, int bs, int avg_sb, int avg_src, int w, int h, unsigned *intra_err,
             unsigned *intrasrc_err, unsigned *inter_err, const MePsy &psy, int ratio, uint8_t *ae_out = NULL,
             unsigned *rest_out = NULL)
{
    unsigned isb = 0, isrc = 0, inter = 0, rest = 0;
    int cw = w / 2, ch = h / 2, n = cw * ch;
    const int gs = me_gshift(w);
    ME_CNT(MEC_ERR_INTRA);
    if (gs >= 0 && avg_sb >= 0 && avg_sb <= 255 && avg_src >= 0 && avg_src <= 255) {
        const int ng = ch << gs, gm = (1 << gs) - 1;
        const uint32_t sb4 = (uint32_t) avg_sb * ME_ONES, sr4 = (uint32_t) avg_src * ME_ONES;
        for (int g = ME_LANE; g < ng; g += ME_NL) {
            int y = (g >> gs) * 2, x = (g & gm) * 4;
            uint32_t a0 = me_ld4(a + y * as + x), a1 = me_ld4(a + (y + 1) * as + x);
            uint32_t b0 = me_ld4(b + y * bs + x), b1 = me_ld4(b + (y + 1) * bs + x);
            for (int c = 0; c < 2; c++) {
                uint32_t A = me_perm(a0, a1, c ? 0x7632 : 0x5410), B = me_perm(b0, b1, c ? 0x7632 : 0x5410);
                int s0 = (int) (me_dot4(A, ME_ONES, 2) >> 2), s1 = (int) (me_dot4(B, ME_ONES, 2) >> 2);
                int ae = (int) (me_dot4(me_absdiff4(A, B), ME_ONES, 2) >> 2);
                int ta = (int) (me_dot4(me_absdiff4(A, me_perm(A, A, 0x0321)), ME_ONES, 2) >> 2);
                int tb = (int) (me_dot4(me_absdiff4(B, me_perm(B, B, 0x0321)), ME_ONES, 2) >> 2);
                unsigned r_ = (unsigned) (me_sqr(ta - tb) << psy.tex_w) + (unsigned) (me_sqr(s0 - s1) << psy.avg_w);
                if (ae_out) ae_out[(g >> gs) * cw + (g & gm) * 2 + c] = (uint8_t) ae;
                inter += (unsigned) (me_sqr(ae) * ratio >> (5 - psy.err_w));
                inter += r_;
                rest += r_;
                ae = (int) (me_dot4(me_absdiff4(A, sb4), ME_ONES, 2) >> 2);
                isb += (unsigned) (me_sqr(ae) << psy.err_w);
                isb += (unsigned) (me_sqr(ta) << psy.tex_w);
                isb += (unsigned) (me_sqr(s0 - avg_sb) << (psy.avg_w + 1));
                ae = (int) (me_dot4(me_absdiff4(A, sr4), ME_ONES, 2) >> 2);
                isrc += (unsigned) (me_sqr(ae) << psy.err_w);
                isrc += (unsigned) (me_sqr(ta) << psy.tex_w);
                isrc += (unsigned) (me_sqr(s0 - avg_src) << (psy.avg_w + 1));
            }
        }
        *intra_err = me_wsumu(isb);
        *intrasrc_err = me_wsumu(isrc);
        *inter_err = me_wsumu(inter) * (unsigned) ratio >> 5;
        if (rest_out) *rest_out = me_wsumu(rest);
        return;
    }
    for (int k = ME_LANE; k < n; k += ME_NL) {
        int j = k / cw, i = k - j * cw;
        const uint8_t *pa = a + (2 * j) * as + 2 * i, *pb = b + (2 * j) * bs + 2 * i;
        int a1 = pa[0], a2 = pa[1], a3 = pa[as], a4 = pa[as + 1];
        int b1 = pb[0], b2 = pb[1], b3 = pb[bs], b4 = pb[bs + 1];
        int s0 = (int) me_uavg4(a1, a2, a3, a4), s1 = (int) me_uavg4(b1, b2, b3, b4);
        int ae, ta, tb;
        unsigned r_;
        ae = (int) me_uavg4(me_abs(a1 - b1), me_abs(a2 - b2), me_abs(a3 - b3), me_abs(a4 - b4));
        ta = (int) me_uavg4(me_abs(a1 - a2), me_abs(a2 - a3), me_abs(a3 - a4), me_abs(a4 - a1));
        tb = (int) me_uavg4(me_abs(b1 - b2), me_abs(b2 - b3), me_abs(b3 - b4), me_abs(b4 - b1));
        if (ae_out) ae_out[k] = (uint8_t) ae;
        r_ = (unsigned) (me_sqr(ta - tb) << psy.tex_w) + (unsigned) (me_sqr(s0 - s1) << psy.avg_w);
        inter += (unsigned) (me_sqr(ae) * ratio >> (5 - psy.err_w));
        inter += r_;
        rest += r_;
        ae = (int) me_uavg4(me_abs(a1 - avg_sb), me_abs(a2 - avg_sb), me_abs(a3 - avg_sb), me_abs(a4 - avg_sb));
        isb += (unsigned) (me_sqr(ae) << psy.err_w);
        isb += (unsigned) (me_sqr(ta) << psy.tex_w);
        isb += (unsigned) (me_sqr(s0 - avg_sb) << (psy.avg_w + 1));
        ae = (int) me_uavg4(me_abs(a1 - avg_src), me_abs(a2 - avg_src), me_abs(a3 - avg_src), me_abs(a4 - avg_src));
        isrc += (unsigned) (me_sqr(ae) << psy.err_w);
        isrc += (unsigned) (me_sqr(ta) << psy.tex_w);
        isrc += (unsigned) (me_sqr(s0 - avg_src) << (psy.avg_w + 1));
    }
    *intra_err = me_wsumu(isb);
    *intrasrc_err = me_wsumu(isrc);
    *inter_err = me_wsumu(inter) * (unsigned) ratio >> 5;
    if (rest_out) *rest_out = me_wsumu(rest);
}

/* the inter error of an 8x8 quadrant for sub-pel gain `ratio` from its per-cell
 * mean absolute errors and the ratio-independent rest (same sums as me_err_intra
 * with err_w = 0) */
DSVCU_DEV unsigned
me_inter_from_cells(const uint8_t *ae, unsigned rest, unsigned ratio)
{
    unsigned acc = 0;
    for (int k = ME_LANE; k < 16; k += ME_NL) {
        acc += (unsigned) (me_sqr((int) ae[k]) * (int) ratio >> 5);
    }
    return (me_wsumu(acc) + rest) * ratio >> 5;
}

struct MeMv { /* working copy of the block's DSV_MV */
    int x, y;
    unsigned flags;
    unsigned err, dc, submask;
};

/* at_s: the block's vector is the prepass' speculated one, so the reference-side
 * quadrant means (and, for full-pel vectors, the error sums of the quadrants that
 * pass the gate) come from the record instead of the pixels */
DSVCU_DEV void
me_test_intra_y(const MeArgs &A, const MePre *P, int at_s, const dsvcu_mv *refmv, MeMv *mv, const uint8_t *srcd, int ss,
                const uint8_t *refd, int rs, int detail_src, int avg_src, int neidif, unsigned ratio, int bw, int bh)
{
    int sbw = bw / 2, sbh = bh / 2, bit_index = 0, nsub = 0;
    unsigned avg_tot = 0, err_sub = 0, err_src = 0;
    MePsy psy;
    int rx = refmv ? refmv->x : mv->x, ry = refmv ? refmv->y : mv->y;
    if ((mv->x | mv->y) != 0 && neidif < 3 && me_abs(rx - mv->x) < 3 && me_abs(ry - mv->y) < 3) return;
    if (sbw == 0 || sbh == 0) return;
    psy.err_w = 0;
    psy.tex_w = 1;
    psy.avg_w = 2;
    detail_src += detail_src / max(neidif, 1);
    for (int g = 0; g <= sbh; g += (sbh + !sbh)) {
        for (int f = 0; f <= sbw; f += (sbw + !sbw)) {
            const uint8_t *src_d = srcd + f + g * ss, *mvr_d = refd + f + g * rs;
            unsigned avg_local, avg_sub, local_detail, dcd, sub_err, src_err, intererr;
            int dc, lo, hi, lerp, sub_better, src_better;
            if (bit_index < 4 && !(mv->submask & (1u << bit_index))) {
                avg_sub = at_s ? P->qa_sub[bit_index] : (unsigned) me_block_avg(mvr_d, rs, sbw, sbh);
                local_detail = P->q_detail[bit_index]; /* source-side, from the prepass */
                avg_local = P->q_avg[bit_index];
                dcd = (unsigned) me_abs((int) avg_local - (int) avg_sub) + 2;
                if (!(local_detail > ((dcd * dcd * (unsigned) bw * (unsigned) bh * ratio) >> 5))) {
                    dc = (int) (avg_local + (unsigned) avg_src * 3 + 2) >> 2;
                    if (at_s && (P->qi_mask & (1 << bit_index)) && (ratio == (1u << 5) || P->qi_cells)) {
                        sub_err = P->qi_sub[bit_index];
                        src_err = P->qi_src[bit_index];
                        intererr = (ratio == (1u << 5)) ? P->qi_inter[bit_index]
                                                        : me_inter_from_cells(P->qi_ae + 16 * bit_index, P->qi_rest[bit_index], ratio);
                    } else {
                        me_err_intra(src_d, ss, mvr_d, rs, (int) avg_sub, dc, sbw, sbh, &sub_err, &src_err, &intererr, psy,
                                     (int) ratio);
                    }
                    lo = me_avg2(detail_src, (int) local_detail);
                    hi = detail_src;
                    lerp = (lo * (32 - A.psyscale) + hi * A.psyscale) >> 5;
                    local_detail = (unsigned) max(lerp, lo);
                    sub_better = (sub_err + local_detail) < intererr;
                    src_better = (src_err + local_detail) < intererr;
                    if (sub_better || src_better) {
                        mv->submask |= (1u << bit_index);
                        err_src += src_err;
                        err_sub += sub_err;
                        avg_tot += (sub_err < src_err) ? avg_sub : (unsigned) dc;
                        nsub++;
                        detail_src = detail_src * 4 / 5;
                    }
                }
            }
            bit_index++;
        }
    }
    if (mv->submask) {
        mv->flags |= MVF_INTRA;
        mv->dc = (err_src < err_sub) ? ((avg_tot / (unsigned) nsub) | 0x100u) : 0;
    }
}

DSVCU_DEV void
me_test_intra_c(const MeArgs &A, MeMv *mv, unsigned mad, unsigned detail_src, unsigned avg_src, int cbx, int cby, int cbmx,
                int cbmy, int cbw, int cbh)
{
    int sbw = cbw / 2, sbh = cbh / 2, bit_index = 0;
    unsigned thr, avg_ramp;
    if (A.effort < 6) return;
    thr = (mv->flags & MVF_INTRA) ? detail_src : detail_src * detail_src;
    if (sbw == 0 || sbh == 0 || mad <= thr || thr > 64 || (me_abs(mv->x) < 4 && me_abs(mv->y) < 4)) return;
    avg_ramp = avg_src * avg_src >> 8;
    for (int g = 0; g <= sbh; g += (sbh + !sbh)) {
        for (int f = 0; f <= sbw; f += (sbw + !sbw)) {
            if (bit_index < 4 && !(mv->submask & (1u << bit_index))) {
                int us, vs, um, vm;
                unsigned dif;
                me_c_average(A.src, cbx + f, cby + g, sbw, sbh, &us, &vs);
                me_c_average(A.ref, cbmx + f, cbmy + g, sbw, sbh, &um, &vm);
                dif = (unsigned) (me_sqr(us - um) + me_sqr(vs - vm)) * avg_ramp >> 8;
                if (dif > thr) mv->submask |= (1u << bit_index);
            }
            bit_index++;
        }
    }
    if (mv->submask) mv->flags |= MVF_INTRA;
}


/* Full-pel metric memo.  The candidate scan and the descent probe overlapping
 * positions, and k_me_prepass has already measured the neighbour-independent
 * candidates; the value is a pure function of the position (the reference
 * recomputes it).  Entries live in per-warp shared memory, one per lane, so a
 * lookup is one compare + ballot. */
DSVCU_DEV int
me_memo_find(const MeScratch *S, int n, int dx, int dy)
{
#ifndef DSVCU_EMU
    /* n <= ME_MEMO = 32 entries, ME_NL of them looked at per step */
    for (int base = 0; base < n; base += ME_NL) {
        const int l = base + ME_LANE;
        const unsigned hit = __ballot_sync(ME_GMASK, l < n && S->memo_x[l] == dx && S->memo_y[l] == dy) >> ME_GBASE;
        if (hit) return base + __ffs(hit) - 1;
    }
    return -1;
#else
    for (int k = 0; k < n; k++) {
        if (S->memo_x[k] == dx && S->memo_y[k] == dy) return k;
    }
    return -1;
#endif
}

DSVCU_DEV void
me_memo_add(MeScratch *S, int &n, int dx, int dy, unsigned v)
{
    if (n < ME_MEMO) {
        if (ME_LANE == 0) {
            S->memo_x[n] = (short) dx;
            S->memo_y[n] = (short) dy;
            S->memo_v[n] = v;
        }
        n++;
        ME_SYNC();
    }
}

DSVCU_DEV unsigned
me_eval(MeScratch *S, int &mn, int level, const uint8_t *srcd, int ss, const MePlane &rp, int bx, int by, int dx, int dy, int bw,
        int bh, const MePsy &psy)
{
    int k = me_memo_find(S, mn, dx, dy);
    ME_CNT(MEC_EVAL);
    if (k >= 0) return S->memo_v[k];
    ME_CNT(MEC_EVAL_MISS);
    unsigned sc = me_hier_metr(level, srcd, ss, rp.data + (by + dy) * rp.stride + bx + dx, rp.stride, bw, bh, psy);
    if (ME_LANE == 0) S->n_evals++;
    me_memo_add(S, mn, dx, dy, sc);
    return sc;
}

/* position with the smallest raw metric among the memo's entries (first entry
 * wins ties); returns 0 when the memo is empty.  Only steers the speculation. */
DSVCU_DEV int
me_memo_argmin(const MeScratch *S, int n, int *px, int *py, unsigned *pv)
{
    if (n <= 0) return 0;
    int k = 0;
    unsigned m = 0xffffffffu;
#ifndef DSVCU_EMU
    for (int base = 0; base < n; base += ME_NL) {
        const int l = base + ME_LANE;
        const unsigned v = l < n ? S->memo_v[l] : 0xffffffffu;
        const unsigned cm = __reduce_min_sync(ME_GMASK, v);
        const unsigned at = __ballot_sync(ME_GMASK, l < n && v == cm) >> ME_GBASE;
        if (cm < m && at) {
            m = cm;
            k = base + __ffs(at) - 1;
        }
    }
#else
    for (int a = 0; a < n; a++) {
        if (S->memo_v[a] < m) {
            m = S->memo_v[a];
            k = a;
        }
    }
#endif
    *px = S->memo_x[k];
    *py = S->memo_y[k];
    *pv = S->memo_v[k];
    return 1;
}

/* source-block statistics -> metric weights and motion bias (hme.c:1445-1481) */
DSVCU_DEV void
me_src_stats(const MeArgs &A, MeScratch *S, const uint8_t *srcd, int ss, int bw, int bh, int gx, int gy, unsigned *pvar,
             unsigned *pavg, int *pbias, MePsy *ppsy)
{
    MePsy psy;
    unsigned var_src = 0, avg_src = 0;
    int motion_bias = A.y_w * A.y_h;
    psy.err_w = 2;
    psy.tex_w = 1;
    psy.avg_w = 0;
    if (A.level <= 1) {
        int tvar;
        var_src = (unsigned) me_block_detail(srcd, ss, bw, bh, &avg_src);
        tvar = (int) (var_src + (var_src >> 10) * (var_src >> 10));
        tvar = ((int) (8u * (unsigned) tvar * (unsigned) A.quant) >> 9) / (bw * bh);
        if (tvar) {
            int hvar = (int) me_block_hist_var(srcd, ss, bw, bh, S->hist);
            int qtex = me_quant_tex(srcd, ss, bw, bh);
            int npeaks = me_block_peaks(srcd, ss, bw, bh, S->hist, (int) avg_src);
            motion_bias += tvar * (hvar - qtex) * npeaks;
        }
        motion_bias = max(motion_bias, 0) / (2 + (me_abs(gx) + me_abs(gy)));
        if (var_src <= (unsigned) (8 * bw * bh * A.quant >> 9)) {
            psy.err_w = 2;
            psy.tex_w = 1;
            psy.avg_w = 2;
            motion_bias = 0;
        } else {
            psy.err_w = 1;
            psy.tex_w = 2;
            psy.avg_w = 1;
        }
        if (var_src > (unsigned) (24 * bw * bh)) psy.avg_w = 0;
    }
    *pvar = var_src;
    *pavg = avg_src;
    *pbias = motion_bias;
    *ppsy = psy;
}

/* candidates that do not depend on same-level neighbours: parent average with
 * outlier rejection (find_inliers, hme.c:1258-1298), temporal neighbours of the
 * previous picture's field (:1229-1256), global motion, parent inliers.
 * Returns has_list (the reference only builds the list when the parent level
 * gave at least one vector); values are raw (before the >> level) */
DSVCU_DEV int
me_nonspatial(const MeArgs &A, int i, int j, int gx, int gy, int *plax, int *play, int *bx, int *by, int *pnb)
{
    const int step = 1 << A.level, nxb = A.nxb, nyb = A.nyb;
    int nb = 0;
    *plax = 0;
    *play = 0;
    *pnb = 0;
    if (!A.parent) return 0;
    const int pt[18] = { 0, 0, -2, 0, 2, 0, 0, -2, 0, 2, -2, -2, 2, 2, 2, -2, -2, 2 };
    int pmask = ~((step << 1) - 1);
    int pi = i & pmask, pj = j & pmask;
    int lx[9], ly[9], npar = 0, sumx = 0, sumy = 0;
    for (int m = 0; m < 9; m++) {
        int x = pi + pt[2 * m] * step, y = pj + pt[2 * m + 1] * step;
        if (x >= 0 && x < nxb && y >= 0 && y < nyb) {
            const dsvcu_mv *pmv = A.parent + x + y * nxb;
            lx[npar] = pmv->x;
            ly[npar] = pmv->y;
            sumx += pmv->x;
            sumy += pmv->y;
            npar++;
        }
    }
    if (!npar) return 0;
    int dist[9], keep[9], nl = 0, avgd = 0, ssd = 0, thresh, ax = 0, ay = 0;
    int lax = sumx / npar, lay = sumy / npar;
    for (int m = 0; m < npar; m++) {
        dist[m] = me_sqr(lx[m] - lax) + me_sqr(ly[m] - lay);
        avgd += dist[m];
    }
    avgd /= npar;
    for (int m = 0; m < npar; m++) ssd += me_sqr(dist[m] - avgd);
    thresh = avgd + (int) me_isqrt((unsigned) (ssd / npar));
    for (int m = 0; m < npar; m++) {
        if (dist[m] <= thresh) {
            ax += lx[m];
            ay += ly[m];
            keep[nl++] = m;
        }
    }
    if (nl) {
        lax = ax / nl;
        lay = ay / nl;
    }
    *plax = lax;
    *play = lay;
    if (A.ref_mvf) {
        const int rectx[9] = { 0, 1, -1, 0, 0, -1, 1, -1, 1 };
        const int recty[9] = { 0, 0, 0, 1, -1, -1, -1, 1, 1 };
        for (int k = 0; k < 9; k++) {
            int rx = i + rectx[k] * step, ry = j + recty[k] * step;
            if (rx < 0 || ry < 0 || rx >= nxb || ry >= nyb) continue;
            bx[nb] = me_sar_r2(A.ref_mvf[rx + ry * nxb].x);
            by[nb] = me_sar_r2(A.ref_mvf[rx + ry * nxb].y);
            nb++;
        }
    }
    bx[nb] = gx;
    by[nb] = gy;
    nb++;
    for (int m = 0; m < nl; m++) {
        bx[nb] = lx[keep[m]];
        by[nb] = ly[keep[m]];
        nb++;
    }
    *pnb = nb;
    return 1;
}

/* statistics of a block against the reference at full-pel offset (fx, fy):
 * metric against the ORIGINAL reference picture, detail + mean of the
 * prediction, chroma means, EPRM clipping tests (hme.c:1640-1690) */
struct MeRefStats {
    unsigned ogrerr, var_ref, avg_ref;
    int u, v, eprm;
};

DSVCU_DEV void
me_ref_stats(const MeArgs &A, MeRefStats *R, const uint8_t *srcd, int i, int j, int bx, int by, int bw, int bh, int fx, int fy,
             int avg_src, const MePsy &psy)
{
    const MePlane &sp = A.src[0], &rp = A.ref[0];
    const uint8_t *refd = rp.data + (by + fy) * rp.stride + bx + fx;
    const uint8_t *ogrd = A.ogr.data + (by + fy) * A.ogr.stride + bx + fx;
    int e0, e1, e2;
    ME_CNT(MEC_REFSTATS);
    R->ogrerr = me_metr(srcd, sp.stride, ogrd, A.ogr.stride, bw, bh, psy);
    R->var_ref = (unsigned) me_block_detail(refd, rp.stride, bw, bh, &R->avg_ref);
    me_c_average(A.ref, i * (A.y_w >> A.hs) + (fx >> A.hs), j * (A.y_h >> A.vs) + (fy >> A.vs), bw >> A.hs, bh >> A.vs, &R->u,
                 &R->v);
    me_calc_eprm(srcd, sp.stride, refd, rp.stride, avg_src, (int) R->avg_ref, bw, bh, &e0, &e1, &e2);
    R->eprm = (e0 ? 1 : 0) | (e1 ? 2 : 0) | (e2 ? 4 : 0);
}

/* neighbour-independent half of refine_level's block loop, all blocks of the
 * level in parallel (one warp per block) */
DSVCU_DEV void
me_prepass_block(const MeArgs &A, MeScratch *S, int i, int j)
{
    const int level = A.level, step = 1 << level;
    const MePlane &sp = A.src[0], &rp = A.ref[0];
    const int gx = A.gxy[0], gy = A.gxy[1];
    int bx = (i * A.y_w) >> level, by = (j * A.y_h) >> level;
    MePre *P = A.pre + i + j * A.nxb;
    if (bx >= sp.w || by >= sp.h) return;
    const uint8_t *srcd = sp.data + by * sp.stride + bx;
    int bw = min(sp.w - bx, A.y_w), bh = min(sp.h - by, A.y_h);
    unsigned var_src, avg_src, zoscore;
    int motion_bias, lax, lay, nb, cbx[ME_PRE_NB], cby[ME_PRE_NB], has, mn = 0, uavg = 0, vavg = 0;
    MePsy psy;
    if (ME_LANE == 0) S->n_evals = S->n_subpel = 0;
    me_src_stats(A, S, srcd, sp.stride, bw, bh, gx, gy, &var_src, &avg_src, &motion_bias, &psy);
    has = me_nonspatial(A, i, j, gx, gy, &lax, &lay, cbx, cby, &nb);
    /* measure zero, the parent average and the list (valid, distinct positions) */
    for (int k = -2; k < (has ? nb : 0); k++) {
        int dx, dy;
        if (k == -2) {
            dx = 0;
            dy = 0;
        } else if (k == -1) {
            if (!has) continue;
            dx = (int16_t) lax >> level;
            dy = (int16_t) lay >> level;
        } else {
            dx = (int16_t) cbx[k] >> level;
            dy = (int16_t) cby[k] >> level;
        }
        if (me_invalid_block(rp.w, rp.h, bx + dx, by + dy, bw, bh, 0)) continue;
        if (mn >= ME_PRE_NM) break;
        (void) me_eval(S, mn, level, srcd, sp.stride, rp, bx, by, dx, dy, bw, bh, psy);
    }
    /* ---- speculation: where will the search end?  Start at the best measured
     * candidate and walk the reference's cross / diagonal descent
     * (refine_best_fpel_cand, hme.c:1300-1370) with the rate term taken against an
     * assumed predictor (the start itself: vector fields are smooth, the real
     * predictor is most often the neighbours' common vector).  Every position
     * probed on the way lands in the memo, which is what the wavefront will look
     * up; (sx, sy) is where the level-0 statistics below are taken. ---- */
    int sx = 0, sy = 0;
    {
        unsigned sbest = 0;
        if (me_memo_argmin(S, mn, &sx, &sy, &sbest)) {
            const int rectx[5] = { 0, 1, -1, 0, 0 }, recty[5] = { 0, 0, 0, 1, -1 };
            MePred sp_pred;
            int again = 1, rounds = 0;
            sp_pred.x = sx * step * 4;
            sp_pred.y = sy * step * 4;
            sbest += (unsigned) me_mv_cost(A, sp_pred, sx * step * 4, sy * step * 4, level);
            while (again && rounds < 6 && mn <= ME_PRE_NM - 5) {
                unsigned metr[4] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu }, score;
                int tvx, tvy;
                again = 0;
                rounds++;
                for (int k = 1; k < 5; k++) {
                    tvx = sx + rectx[k];
                    tvy = sy + recty[k];
                    if (me_invalid_block(rp.w, rp.h, bx + tvx, by + tvy, bw, bh, 0)) continue;
                    score = me_eval(S, mn, level, srcd, sp.stride, rp, bx, by, tvx, tvy, bw, bh, psy);
                    metr[k - 1] = score;
                    score += (unsigned) me_mv_cost(A, sp_pred, tvx * step * 4, tvy * step * 4, level);
                    if (sbest > score) {
                        sbest = score;
                        sx = tvx;
                        sy = tvy;
                        again = 1;
                        break;
                    }
                }
                if (again) continue;
                tvx = sx + rectx[(metr[0] <= metr[1]) ? 1 : 2];
                tvy = sy + recty[(metr[2] <= metr[3]) ? 3 : 4];
                if (me_invalid_block(rp.w, rp.h, bx + tvx, by + tvy, bw, bh, 0)) break;
                score = me_eval(S, mn, level, srcd, sp.stride, rp, bx, by, tvx, tvy, bw, bh, psy);
                score += (unsigned) me_mv_cost(A, sp_pred, tvx * step * 4, tvy * step * 4, level);
                if (sbest > score) {
                    sbest = score;
                    sx = tvx;
                    sy = tvy;
                    again = 1;
                }
            }
        }
    }
    zoscore = me_metr(srcd, sp.stride, A.ogr.data + by * A.ogr.stride + bx, A.ogr.stride, bw, bh, psy);
    MeRefStats rs;
    int s_valid = 0, utex = 0, vtex = 0, qi_mask = 0;
    unsigned q_detail[4] = { 0, 0, 0, 0 }, q_avg[4] = { 0, 0, 0, 0 }, qa_sub[4] = { 0, 0, 0, 0 };
    unsigned qi_sub[4] = { 0, 0, 0, 0 }, qi_src[4] = { 0, 0, 0, 0 }, qi_inter[4] = { 0, 0, 0, 0 }, qi_rest[4] = { 0, 0, 0, 0 };
    const int qi_cells = (bw == 16 && bh == 16); /* 8x8 quadrants: 16 cells each */
    unsigned bsub[3] = { 0, 0, 0 }, zsub[3] = { 0, 0, 0 };
    if (level == 0) {
        const int sbw = bw / 2, sbh = bh / 2;
        const int cbw = bw >> A.hs, cbh = bh >> A.vs, ccx = i * (A.y_w >> A.hs), ccy = j * (A.y_h >> A.vs);
        const int s_ok = !me_invalid_block(rp.w, rp.h, bx + sx, by + sy, bw, bh, 0);
        int qn = 0;
        me_c_average(A.src, ccx, ccy, cbw, cbh, &uavg, &vavg);
        if (sbw && sbh) {
            for (int g = 0; g <= sbh; g += (sbh + !sbh)) {
                for (int f = 0; f <= sbw; f += (sbw + !sbw)) {
                    if (qn < 4) q_detail[qn] = (unsigned) me_block_detail(srcd + f + g * sp.stride, sp.stride, sbw, sbh, &q_avg[qn]);
                    qn++;
                }
            }
        }
        if (cbw > 0 && cbh > 0) {
            utex = (int) me_block_tex(A.src[1].data + ccy * A.src[1].stride + ccx, A.src[1].stride, cbw, cbh);
            vtex = (int) me_block_tex(A.src[2].data + ccy * A.src[2].stride + ccx, A.src[2].stride, cbw, cbh);
        }
            if (s_ok) {
            /* what the mode decision reads at the final position (hme.c:1640-1690) */
            me_ref_stats(A, &rs, srcd, i, j, bx, by, bw, bh, sx, sy, (int) avg_src, psy);
            s_valid |= ME_SV_RS;
            if (!A.lossless) {
                MeChroma cpsy;
                me_chroma_analysis(&cpsy, (int) avg_src, uavg, vavg);
                const int y_pre = me_abs((int) avg_src - (int) rs.avg_ref) <= 2;
                const int c_pre = !cpsy.greyish && me_avg2(me_abs(uavg - rs.u), me_abs(vavg - rs.v)) <= 2;
                if ((y_pre || c_pre) && cbw > 0 && cbh > 0) {
                    me_yuv_max_sub(bsub, A.src, A.ref, bx, by, bx + sx, by + sy, bw, bh, ccx, ccy, ccx + (sx >> A.hs),
                                   ccy + (sy >> A.vs), cbw, cbh, psy);
                    s_valid |= ME_SV_BSUB;
                }
            }
            if (sbw && sbh) {
                /* reference-side half of test_subblock_intra_y (hme.c:891-1001) at S, for a
                 * full-pel vector (ratio = 32); the wavefront redoes the scalar part */
                const uint8_t *refd = rp.data + (by + sy) * rp.stride + bx + sx;
                MePsy ipsy;
                int q = 0;
                ipsy.err_w = 0;
                ipsy.tex_w = 1;
                ipsy.avg_w = 2;
                for (int g = 0; g <= sbh; g += (sbh + !sbh)) {
                    for (int f = 0; f <= sbw; f += (sbw + !sbw)) {
                        if (q < 4) {
                            const uint8_t *src_d = srcd + f + g * sp.stride, *mvr_d = refd + f + g * rp.stride;
                            unsigned dcd;
                            qa_sub[q] = (unsigned) me_block_avg(mvr_d, rp.stride, sbw, sbh);
                            dcd = (unsigned) me_abs((int) q_avg[q] - (int) qa_sub[q]) + 2;
                            if (!(q_detail[q] > ((dcd * dcd * (unsigned) bw * (unsigned) bh * 32u) >> 5))) {
                                int dc = (int) (q_avg[q] + (unsigned) avg_src * 3 + 2) >> 2;
                                me_err_intra(src_d, sp.stride, mvr_d, rp.stride, (int) qa_sub[q], dc, sbw, sbh, &qi_sub[q],
                                             &qi_src[q], &qi_inter[q], ipsy, 32, qi_cells ? P->qi_ae + 16 * q : NULL,
                                             &qi_rest[q]);
                                qi_mask |= 1 << q;
                            }
                        }
                        q++;
                    }
                }
                s_valid |= ME_SV_INTRA;
            }
        }
            /* skip test (hme.c:1695-1729): only blocks that can end on the zero vector */
        if (A.skip_thresh >= 0 && !A.lossless && cbw > 0 && cbh > 0 &&
            ((sx | sy) == 0 || zoscore < 2u * (unsigned) (A.quant * bw * bh >> 11))) {
            me_yuv_max_sub(zsub, A.src, A.ref, bx, by, bx, by, bw, bh, ccx, ccy, ccx, ccy, cbw, cbh, psy);
            s_valid |= ME_SV_ZSUB;
        }
            /* the sub-pel measurements around (lax, lay) and around S are taken by k_me_subpel,
         * one warp per measurement, once this kernel has written (lax, lay) and S */
    }
    ME_SYNC();
    if (ME_LANE == 0) {
        if (S->n_evals) atomicAdd(&A.acc[6], S->n_evals);
        if (S->n_subpel) atomicAdd(&A.acc[7], S->n_subpel);
        P->sx = sx;
        P->sy = sy;
        P->s_valid = s_valid;
        if (s_valid & ME_SV_RS) {
            P->rs_ogrerr = rs.ogrerr;
            P->rs_var = rs.var_ref;
            P->rs_avg = rs.avg_ref;
            P->rs_u = rs.u;
            P->rs_v = rs.v;
            P->rs_eprm = rs.eprm;
        }
        for (int k = 0; k < 3; k++) {
            P->bsub[k] = bsub[k];
            P->zsub[k] = zsub[k];
        }
        for (int k = 0; k < 4; k++) {
            P->q_detail[k] = q_detail[k];
            P->q_avg[k] = q_avg[k];
            P->qa_sub[k] = qa_sub[k];
            P->qi_sub[k] = qi_sub[k];
            P->qi_src[k] = qi_src[k];
            P->qi_inter[k] = qi_inter[k];
            P->qi_rest[k] = qi_rest[k];
        }
        P->qi_mask = qi_mask;
        P->qi_cells = qi_cells;
        P->utex = utex;
        P->vtex = vtex;
        P->sp_valid[0] = P->sp_valid[1] = 0; /* set by k_me_subpel */
        P->sp_nv[0] = P->sp_nv[1] = 0;
        P->var_src = var_src;
        P->avg_src = avg_src;
        P->motion_bias = motion_bias;
        P->psy_pack = psy.err_w | (psy.tex_w << 8) | (psy.avg_w << 16);
        P->lax = lax;
        P->lay = lay;
        P->has_list = has;
        P->nb = nb;
        P->zoscore = zoscore;
        P->uavg = uavg;
        P->vavg = vavg;
        P->nm = mn;
        for (int k = 0; k < nb; k++) {
            P->bx[k] = (short) cbx[k];
            P->by[k] = (short) cby[k];
        }
    }
    for (int k = ME_LANE; k < mn; k += ME_NL) {
        P->mx[k] = S->memo_x[k];
        P->my[k] = S->memo_y[k];
        P->mv[k] = S->memo_v[k];
    }
    ME_SYNC();
}

/* One sub-pel measurement of the level-0 prepass record of block (i, j): slot 0 around
 * the parent average (the reference's first pass, always taken), slot 1 around the
 * speculated winner S when that is another position (its second pass). */
DSVCU_DEV void
me_subpel_task(const MeArgs &A, MeScratch *S, int i, int j, int slot)
{
    const MePlane &sp = A.src[0], &rp = A.ref[0];
    MePre *P = A.pre + i + j * A.nxb;
    const int bx = i * A.y_w, by = j * A.y_h;
    if (bx >= sp.w || by >= sp.h) return;
    const int bw = min(sp.w - bx, A.y_w), bh = min(sp.h - by, A.y_h);
    const int lax = P->lax, lay = P->lay, sx = P->sx, sy = P->sy;
    const int fx = slot ? sx : lax, fy = slot ? sy : lay;
    MePsy psy;
    MeSubpel M;
    if (slot && sx == lax && sy == lay) return;
    if (me_invalid_block(rp.w, rp.h, bx + fx, by + fy, bw, bh, 4)) return;
    psy.err_w = P->psy_pack & 255;
    psy.tex_w = (P->psy_pack >> 8) & 255;
    psy.avg_w = (P->psy_pack >> 16) & 255;
    if (ME_LANE == 0) S->n_evals = S->n_subpel = 0;
    ME_SYNC();
    me_subpel_measure(A, S, &M, fx, fy, bx, by, bw, bh, psy);
    if (ME_LANE == 0) {
        atomicAdd(&A.acc[6], S->n_evals);
        atomicAdd(&A.acc[7], S->n_subpel);
        P->sp_nv[slot] = M.nv;
        for (int k = 0; k < M.nv; k++) {
            P->sp_tx[slot][k] = M.tx[k];
            P->sp_ty[slot][k] = M.ty[k];
            P->sp_sc[slot][k] = M.sc[k];
        }
        P->sp_valid[slot] = 1;
    }
    ME_SYNC();
}

/* ---- one block of refine_level (hme.c:1413-1823) ---- */


DSVCU_DEV void
me_block(const MeArgs &A, MeScratch *S, uint32_t *pre_words, int i, int j, int *acc_local)
{
    const int level = A.level, step = 1 << level;
    const MePlane &sp = A.src[0], &rp = A.ref[0];
    const int nxb = A.nxb, nyb = A.nyb;
    int bx = (i * A.y_w) >> level, by = (j * A.y_h) >> level;
    dsvcu_mv *out = A.mvf + i + j * nxb;
    int cx[ME_MAXCAND], cy[ME_MAXCAND], n = 0;
    int bw, bh, dx, dy, lax = 0, lay = 0, motion_bias, good_enough = 0;
    unsigned best, score_zero, score, best_score, qthresh, var_src = 0, avg_src = 0;
    MePsy psy;
    const uint8_t *srcd;

    psy.err_w = 2;
    psy.tex_w = 1;
    psy.avg_w = 0;
    if (bx >= sp.w || by >= sp.h) {
        return; /* field is zero-initialised: inter, zero vector */
    }
    srcd = sp.data + by * sp.stride + bx;
    bw = min(sp.w - bx, A.y_w);
    bh = min(sp.h - by, A.y_h);
    MePred pred;
    int mn = 0; /* entries in the block's metric memo */
    if (ME_LANE == 0) S->n_evals = S->n_subpel = 0;
    /* one coalesced batch of loads (a single L2 round trip) instead of a miss
     * per touched line of the record */
    {
        /* 16 bytes per lane and step (sizeof(MePre) is a multiple of 16, records and the staging area are 16-byte aligned) */
        const uint4 *g = (const uint4 *) (A.pre + i + j * nxb);
        uint4 *d = (uint4 *) pre_words;
        for (int k = ME_LANE; k < (int) (sizeof(MePre) / 16); k += ME_NL) d[k] = g[k];
        ME_SYNC();
    }
    const MePre *P = (const MePre *) pre_words;
    me_movec_pred(A.mvf, nxb, i, j, &pred.x, &pred.y);
    /* neighbour-independent results of k_me_prepass: statistics, metric weights,
     * the non-spatial candidates and their metrics (memo seed) */
    var_src = P->var_src;
    avg_src = P->avg_src;
    motion_bias = P->motion_bias;
    psy.err_w = P->psy_pack & 255;
    psy.tex_w = (P->psy_pack >> 8) & 255;
    psy.avg_w = (P->psy_pack >> 16) & 255;
    mn = P->nm;
    for (int k = ME_LANE; k < mn; k += ME_NL) {
        S->memo_x[k] = P->mx[k];
        S->memo_y[k] = P->my[k];
        S->memo_v[k] = P->mv[k];
    }
    ME_SYNC();
    cx[n] = 0;
    cy[n] = 0;
    n++;
    if (P->has_list) {
        const int nb = P->nb;
        lax = P->lax;
        lay = P->lay;
        cx[n] = lax;
        cy[n] = lay;
        n++;
        /* spatial predictions (hme.c:1202-1227); vectors pass through the
         * qpel->fpel rounding whatever unit they are stored in */
        if (level == 0) {
            cx[n] = me_sar_r2(pred.x);
            cy[n] = me_sar_r2(pred.y);
            n++;
        }
        if (i > 0) {
            int mx_, my_;
            me_ldmv(A.mvf + (i - step) + j * nxb, &mx_, &my_, NULL);
            cx[n] = me_sar_r2(mx_);
            cy[n] = me_sar_r2(my_);
            n++;
        }
        if (j > 0) {
            int mx_, my_;
            me_ldmv(A.mvf + i + (j - step) * nxb, &mx_, &my_, NULL);
            cx[n] = me_sar_r2(mx_);
            cy[n] = me_sar_r2(my_);
            n++;
        }
        if (i > 0 && j > 0) {
            int mx_, my_;
            me_ldmv(A.mvf + (i - step) + (j - step) * nxb, &mx_, &my_, NULL);
            cx[n] = me_sar_r2(mx_);
            cy[n] = me_sar_r2(my_);
            n++;
        }
        /* temporal neighbours, global motion, parent inliers (from the prepass) */
        for (int k = 0; k < nb; k++) {
            cx[n] = P->bx[k];
            cy[n] = P->by[k];
            n++;
        }
    }
    /* candidates live in int16 fields in the reference */
    for (int k = 0; k < n; k++) {
        cx[k] = (int16_t) cx[k] >> level;
        cy[k] = (int16_t) cy[k] >> level;
    }
    {
        /* The reference removes duplicate positions (remove_dupes, hme.c:1166-1183)
         * and scores the survivors in list order, keeping a strictly better score.
         * A duplicate scores exactly like its first occurrence and so can never
         * replace it: the list is scored as it stands, one candidate per lane --
         * memo look-up (nearly every candidate was measured by the prepass), rate
         * term, bias -- and the winner is the FIRST candidate that attains the
         * minimum (min reduction + ballot).  Candidates the memo does not hold (the
         * spatial ones, now and then) are measured by the whole group, in list order. */
        int bestx = cx[0], besty = cy[0];
        best_score = score_zero = 0xffffffffu;
        for (int base = 0; base < n; base += ME_NL) {
            const int k = base + ME_LANE;
            int mx_ = 0, my_ = 0, live = 0, hit = 0;
            unsigned v = 0;
            if (k < n) {
                mx_ = cx[k];
                my_ = cy[k];
                live = !me_invalid_block(rp.w, rp.h, bx + mx_, by + my_, bw, bh, 0);
            }
            if (live) {
                for (int e = 0; e < mn; e++) {
                    if (S->memo_x[e] == mx_ && S->memo_y[e] == my_) {
                        v = S->memo_v[e];
                        hit = 1;
                        break;
                    }
                }
            }
            for (unsigned miss = me_gballot(live && !hit); miss; miss &= miss - 1) {
                const int l = me_ctz(miss);
                const unsigned sc = me_eval(S, mn, level, srcd, sp.stride, rp, bx, by, me_gbcast(mx_, l), me_gbcast(my_, l),
                                            bw, bh, psy);
                if (ME_LANE == l) v = sc;
            }
            /* candidate 0 is the zero vector (always a valid position) */
            if (base == 0) score_zero = (unsigned) me_gbcast((int) v, 0);
            score = 0xffffffffu;
            if (live) {
                score = v + (unsigned) me_mv_cost(A, pred, mx_ * step * 4, my_ * step * 4, level);
                if (mx_ == lax && my_ == lay) score = (unsigned) max((int) score - (motion_bias >> level), 0);
            }
            const unsigned m = me_wminu(score);
            if (best_score > m) {
                const int w = me_ctz(me_gballot(live && score == m));
                best_score = m;
                bestx = me_gbcast(mx_, w);
                besty = me_gbcast(my_, w);
            }
        }
        dx = bestx;
        dy = besty;
    }
    best = best_score;
    qthresh = (unsigned) (A.quant * bw * bh >> 11);
    {
        unsigned zoscore = P->zoscore;
        if (me_abs(dx) <= 1 && me_abs(dy) <= 1) qthresh *= 2;
        if (zoscore < qthresh) {
            best = (level == 0) ? score_zero : 0;
            dx = 0;
            dy = 0;
            good_enough = 1;
        }
    }
    if (!good_enough) {
        /* refine_best_fpel_cand (hme.c:1300-1370) */
        const int rectx[9] = { 0, 1, -1, 0, 0, -1, 1, -1, 1 };
        const int recty[9] = { 0, 0, 0, 1, -1, -1, -1, 1, 1 };
        unsigned metr[4] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu };
        int again = 1;
        while (again) {
            int tvx, tvy;
            again = 0;
            for (int k = 0; k < 5; k++) {
                tvx = dx + rectx[k];
                tvy = dy + recty[k];
                if (me_invalid_block(rp.w, rp.h, bx + tvx, by + tvy, bw, bh, 0)) continue;
                score = me_eval(S, mn, level, srcd, sp.stride, rp, bx, by, tvx, tvy, bw, bh, psy);
                if (k >= 1) metr[k - 1] = score;
                if (level == 0 && !tvx && !tvy && score <= qthresh) {
                    dx = tvx;
                    dy = tvy;
                    best = score;
                    good_enough = 1;
                    break;
                }
                score += (unsigned) me_mv_cost(A, pred, tvx * step * 4, tvy * step * 4, level);
                if (best > score) {
                    best = score;
                    dx = tvx;
                    dy = tvy;
                    again = 1;
                    break;
                }
            }
            if (again || good_enough) continue;
            tvx = dx + rectx[(metr[0] <= metr[1]) ? 1 : 2];
            tvy = dy + recty[(metr[2] <= metr[3]) ? 3 : 4];
            if (me_invalid_block(rp.w, rp.h, bx + tvx, by + tvy, bw, bh, 0)) break;
            score = me_eval(S, mn, level, srcd, sp.stride, rp, bx, by, tvx, tvy, bw, bh, psy);
            score += (unsigned) me_mv_cost(A, pred, tvx * step * 4, tvy * step * 4, level);
            if (best > score) {
                best = score;
                dx = tvx;
                dy = tvy;
                again = 1;
            }
        }
    }

    MeMv mv;
    mv.x = dx * step;
    mv.y = dy * step;
    mv.flags = 0;
    mv.err = 0;
    mv.dc = 0;
    mv.submask = 0;

    if (level == 0) {
        int fpelx = mv.x, fpely = mv.y, subx = 0, suby = 0;
        unsigned yarea = (unsigned) (bw * bh), best_fp;
        if (fpelx == lax && fpely == lay) best += (unsigned) motion_bias;
        best_fp = best;
        if (A.effort >= 4) {
            int tried_la = 0;
            if (!me_invalid_block(rp.w, rp.h, bx + lax, by + lay, bw, bh, 4)) {
                /* measured by the prepass; only the decision is left */
                if (best_fp != 0) {
                    best = me_subpel_decide(A, P->sp_nv[0], P->sp_tx[0], P->sp_ty[0], P->sp_sc[0], &subx, &suby, lax, lay, pred,
                                            best_fp, bw, bh);
                } else {
                    best = best_fp;
                }
                tried_la = 1;
                if (subx | suby) {
                    fpelx = lax;
                    fpely = lay;
                }
            }
            /* the reference repeats the refinement around the full-pel winner;
             * when that is the position just tried (and nothing was found) the
             * second pass would recompute the very same numbers */
            if (!(subx | suby) && !good_enough && !(tried_la && fpelx == lax && fpely == lay) &&
                !me_invalid_block(rp.w, rp.h, bx + fpelx, by + fpely, bw, bh, 4)) {
                if (best_fp != 0) {
                    if (P->sp_valid[1] && fpelx == P->sx && fpely == P->sy) {
                        /* the prepass guessed this winner and measured around it */
                        best = me_subpel_decide(A, P->sp_nv[1], P->sp_tx[1], P->sp_ty[1], P->sp_sc[1], &subx, &suby, fpelx,
                                                fpely, pred, best_fp, bw, bh);
                    } else {
                        MeSubpel M;
                        me_subpel_measure(A, S, &M, fpelx, fpely, bx, by, bw, bh, psy);
                        best = me_subpel_decide(A, M.nv, M.tx, M.ty, M.sc, &subx, &suby, fpelx, fpely, pred, best_fp, bw, bh);
                    }
                } else {
                    best = best_fp;
                }
            }
        }
        mv.x = fpelx * 4 + subx;
        mv.y = fpely * 4 + suby;
        /* publish the vector now: the neighbour difference below reads it */
        if (ME_LANE == 0) {
            out->x = (int16_t) mv.x;
            out->y = (int16_t) mv.y;
            out->flags = 0;
        }
        ME_SYNC();
        {
            const uint8_t *refd = rp.data + (by + fpely) * rp.stride + bx + fpelx;
            unsigned var_ref, avg_ref, mad, ogrerr, ogrmad, avg_y_dif, avg_c_dif;
            int uavg_src, vavg_src, uavg_ref, vavg_ref, cbx, cby, cbw, cbh, cbmx, cbmy;
            int eprmi, eprmd, eprmr, neidif, oob, ipolvar, dv, skipped = 0;
            unsigned skipt = ((unsigned) A.quant * (unsigned) A.quant) >> 19;
            unsigned ratio = 1 << 5, chroma_ratio;
            MeChroma cpsy;
            const dsvcu_mv *refmv = A.ref_mvf ? A.ref_mvf + i + j * nxb : NULL;

            if ((mv.x | mv.y) & 3) ratio = (best << 5) / (best_fp + !best_fp);
            /* is the block where the prepass expected it?  then everything the mode
             * decision needs from the pixels is in the record */
            const int at_s = (fpelx == P->sx && fpely == P->sy);
            ME_CNT(MEC_BLOCKS);
            if (at_s) ME_CNT(MEC_AT_S);
            {
                MeRefStats rs;
                if (at_s && (P->s_valid & ME_SV_RS)) {
                    rs.ogrerr = P->rs_ogrerr;
                    rs.var_ref = P->rs_var;
                    rs.avg_ref = P->rs_avg;
                    rs.u = P->rs_u;
                    rs.v = P->rs_v;
                    rs.eprm = P->rs_eprm;
                } else {
                    me_ref_stats(A, &rs, srcd, i, j, bx, by, bw, bh, fpelx, fpely, (int) avg_src, psy);
                }
                ogrerr = rs.ogrerr;
                var_ref = rs.var_ref;
                avg_ref = rs.avg_ref;
                uavg_ref = rs.u;
                vavg_ref = rs.v;
                eprmi = rs.eprm & 1;
                eprmd = (rs.eprm >> 1) & 1;
                eprmr = (rs.eprm >> 2) & 1;
            }
            ogrmad = (ogrerr + yarea / 2) / yarea;
            ogrmad = ogrmad * ratio >> 5;
            mad = (best + yarea / 2) / yarea;
            dv = (int) min(ratio, 32u);
            ipolvar = (int) ((var_src * (unsigned) dv + var_ref * (unsigned) (32 - dv)) >> 5);
            dv = me_abs((int) var_src - ipolvar);
            if ((var_src > 16 * yarea) && (var_src < 32 * yarea)) mv.flags |= MVF_MAINTAIN;

            cbx = i * (A.y_w >> A.hs);
            cby = j * (A.y_h >> A.vs);
            cbmx = cbx + (fpelx >> A.hs);
            cbmy = cby + (fpely >> A.vs);
            cbw = bw >> A.hs;
            cbh = bh >> A.vs;
            chroma_ratio = ((unsigned) (cbw * cbh) << 4) / yarea;
            uavg_src = P->uavg;
            vavg_src = P->vavg;
            me_chroma_analysis(&cpsy, (int) avg_src, uavg_src, vavg_src);
            avg_y_dif = (unsigned) me_abs((int) avg_src - (int) avg_ref);
            avg_c_dif = (unsigned) me_avg2(me_abs(uavg_src - uavg_ref), me_abs(vavg_src - vavg_ref));
            {   /* outofbounds (hme.c:413-424) */
                int limx = ((nxb - 1) * A.y_w) - 1, limy = ((nyb - 1) * A.y_h) - 1;
                int px = i * A.y_w + (mv.x >> 2), py = j * A.y_h + (mv.y >> 2);
                oob = (px < 0 || py < 0 || px >= limx || py >= limy);
            }
            neidif = me_neighbordif(A.mvf, nxb, i, j);

            if ((good_enough || (mv.x | mv.y) == 0) && A.skip_thresh >= 0 && !A.lossless) {
                unsigned cth, sth = skipt * yarea, zsub[3];
                sth += 4 * var_src;
                sth += yarea * (unsigned) A.skip_thresh;
                if (A.quant < (1 << 10)) sth = sth * (unsigned) A.quant >> 10;
                if (avg_y_dif <= 2) sth = max(sth, 3 * (yarea + var_src));
                sth = max(sth, yarea);
                if (good_enough) sth *= 2;
                if (P->s_valid & ME_SV_ZSUB) {
                    zsub[0] = P->zsub[0];
                    zsub[1] = P->zsub[1];
                    zsub[2] = P->zsub[2];
                } else {
                    me_yuv_max_sub(zsub, A.src, A.ref, bx, by, bx, by, bw, bh, cbx, cby, cbx, cby, cbw, cbh, psy);
                }
                cth = (chroma_ratio * sth * max(skipt, 1u) >> (4 + 1));
                zsub[0] = zsub[0] * ratio >> 5;
                zsub[1] = zsub[1] * ratio >> 5;
                zsub[2] = zsub[2] * ratio >> 5;
                zsub[0] += (unsigned) me_sqr((int) avg_src - (int) avg_ref) * yarea;
                if (zsub[0] <= sth && zsub[1] <= cth && zsub[2] <= cth) {
                    mv.flags |= MVF_SKIP;
                    mv.x = 0;
                    mv.y = 0;
                    mv.err = 0;
                    skipped = 1;
                }
            }
            if (!skipped) {
                if (!oob && !A.lossless) {
                    int y_pre = (avg_y_dif <= 2), c_pre = !cpsy.greyish && (avg_c_dif <= 2);
                    if (y_pre || c_pre) {
                        unsigned bsub[3], xth = skipt * yarea;
                        int utex, vtex, carea = 4 * cbw * cbh;
                        if (at_s && (P->s_valid & ME_SV_BSUB)) {
                            bsub[0] = P->bsub[0];
                            bsub[1] = P->bsub[1];
                            bsub[2] = P->bsub[2];
                        } else {
                            me_yuv_max_sub(bsub, A.src, A.ref, bx, by, bx + fpelx, by + fpely, bw, bh, cbx, cby, cbmx, cbmy,
                                           cbw, cbh, psy);
                        }
                        xth += (unsigned) ipolvar;
                        xth = (unsigned) max((int) xth - ((int) yarea * neidif * 2), 0);
                        xth = xth * (unsigned) A.quant >> 12;
                        xth = xth < 32 ? 32 : (xth > yarea * 4 ? yarea * 4 : xth);
                        bsub[0] = bsub[0] * ratio >> 5;
                        bsub[1] = bsub[1] * ratio >> 5;
                        bsub[2] = bsub[2] * ratio >> 5;
                        if (y_pre && bsub[0] < 4 * xth) mv.flags |= MVF_NOXMITY;
                        utex = P->utex;
                        vtex = P->vtex;
                        c_pre &= (utex > carea || vtex > carea);
                        xth = chroma_ratio * xth >> 4;
                        if (c_pre && bsub[1] < xth && bsub[2] < xth) mv.flags |= MVF_NOXMITC;
                    }
                    if ((unsigned) dv < (var_src / 4)) mv.flags |= MVF_SIMCMPLX;
                }
                me_test_intra_y(A, P, at_s && (P->s_valid & ME_SV_INTRA), refmv, &mv, srcd, sp.stride, refd, rp.stride, ipolvar,
                                (int) avg_src, neidif, ratio, bw, bh);
                me_test_intra_c(A, &mv, mad, (unsigned) (ipolvar / (bw * bh)), avg_src, cbx, cby, cbmx, cbmy, cbw, cbh);
                if (!(mv.flags & MVF_NOXMITY)) {
                    mv.err = mad & 0xffff;
                    acc_local[3] += (int) mad;
                }
                acc_local[1] += (ogrmad > 11) + (avg_c_dif >= 32);
            }
            if (best > 0) acc_local[2]++;
            if (mv.flags & MVF_INTRA) {
                int merged = (mv.dc & 0x100) ? eprmd : eprmi;
                if (mv.submask != 15) merged |= eprmr;
                if (merged) mv.flags |= MVF_EPRM;
                acc_local[0]++;
                mv.x = fpelx * 4;
                mv.y = fpely * 4;
            } else {
                int merged = eprmr;
                if (mv.submask) merged |= eprmi;
                if (merged) mv.flags |= MVF_EPRM;
            }
            if (mv.flags & (MVF_INTRA | MVF_EPRM)) mv.flags &= ~(unsigned) MVF_SIMCMPLX;
        }
    }
    if (ME_LANE == 0) {
        if (S->n_evals) atomicAdd(&A.acc[6], S->n_evals);
        if (S->n_subpel) atomicAdd(&A.acc[7], S->n_subpel);
        out->x = (int16_t) mv.x;
        out->y = (int16_t) mv.y;
        out->flags = mv.flags;
        out->err = (uint16_t) mv.err;
        out->dc = (uint16_t) mv.dc;
        out->submask = (uint8_t) mv.submask;
    }
    ME_SYNC();
}

