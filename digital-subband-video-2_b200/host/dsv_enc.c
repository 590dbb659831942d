/*
 * dsv_enc.c -- encoder control for the B200 build.
 *
 * Public behaviour is that of the reference encoder (src/dsv_encoder.c): the
 * same DSV_ENCODER fields, the same decisions (GOP logic :1247-1271, scene-change
 * detection :545-651, rate control :72-106 + :252-467, loop-filter switch
 * :518-543, per-block side information :796-932, :692-794) and therefore the
 * same bytes.  What is different is where the work happens:
 *
 *   host (this file)   rate control, GOP / scene-cut logic, block side
 *                      information, all bit packing
 *   device (dsv_cuda.h) every pixel operator: border extension, the luma
 *                      pyramids, motion search + mode decision, intra block
 *                      analysis, prediction / residual, forward + inverse SBT,
 *                      quantisation, reconstruction, loop / intra filters
 *
 * Source, residual, prediction and reconstructed pictures never leave the GPU.
 * Per picture the host uploads the source once, reads back one small block
 * array (vectors + flags, or intra flags) to take its decisions, uploads the
 * final block data, queues the whole pixel pipeline on one stream and then only
 * waits for the ordered non-zero symbol list of each plane, which it packs into
 * the HZCC bit format while the GPU finishes reconstruction and prepares the
 * reference pyramid of the next picture.
 */
#include <stdlib.h>
#include <string.h>
#include "dsv_host.h"
#include "../../include/dsv_encoder.h"

#include <time.h>

/* optional wall-clock phase accounting (DSV_PROFILE=1): where one encoder
 * instance spends its time, summed over pictures and printed by dsv_enc_free */
static int g_prof = -1;
enum { PH_UPLOAD, PH_ANALYSIS_WAIT, PH_DECIDE, PH_SIDEINFO, PH_QUEUE, PH_SYM_WAIT, PH_PACK, PH_N };
static const char *ph_name[PH_N] = { "upload+queue analysis", "wait analysis", "decisions", "side info bits",
                                     "queue pixel pipeline", "wait symbols", "pack planes" };
static double
now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
#define PROF_MARK(g, ph)                         \
    do {                                         \
        if (g_prof > 0) {                        \
            double t_ = now_ms();                \
            (g)->ph_ms[ph] += t_ - (g)->ph_t0;   \
            (g)->ph_t0 = t_;                     \
        }                                        \
    } while (0)

#define QPCT(pct) ((pct) * DSV_RC_QUAL_SCALE)
#define ISQ(x) ((x) * (x))

/* device-side state of one encoder instance; hangs off DSV_ENCODER.ref */
struct _DSV_ENCDATA {
    dsvcu_ctx *ctx;
    int w, h, subsamp;
    dsvcu_frame *src[2];      /* padded source pictures: current / reference's original */
    dsvcu_pyramid *src_pyr[2];
    dsvcu_frame *rec[2];      /* residual -> reconstruction: current / reference */
    dsvcu_pyramid *ref_pyr;   /* pyramid of the reference reconstruction */
    dsvcu_frame *pred;
    dsvcu_coefs *coefs;
    int cur;                  /* index of the current picture in src[] / rec[] */
    int have_ref;             /* rec[cur ^ 1] / src[cur ^ 1] hold a reference */
    int ref_has_mvs;          /* the reference picture went through the motion search */
    DSV_MV *mvs;              /* this picture's field (host copy) */
    DSV_MV *imv;              /* intra analysis flags (host copy) */
    int nblk;
    int pyr_levels;           /* depth the three pyramids were created with */
    double ph_ms[8], ph_t0;
    int ph_frames;
    double gpu_ms[8];         /* DSV_PROFILE=2: device time between phase stamps */
    int gpu_frames, marks_live;
};

int dsv_get_thread_device(void);

static void
gpu_state_free(DSV_ENCDATA *g)
{
    int i;
    if (!g) {
        return;
    }
    if (g_prof > 0 && g->ph_frames) {
        double tot = 0;
        for (i = 0; i < PH_N; i++) tot += g->ph_ms[i];
        printf("[dsv_enc profile] %d pictures, %.2f ms/picture:", g->ph_frames, tot / g->ph_frames);
        for (i = 0; i < PH_N; i++) printf(" %s %.2f;", ph_name[i], g->ph_ms[i] / g->ph_frames);
        printf("\n");
        if (g->gpu_frames) {
            static const char *gn[6] = { "upload+extend+pyramid", "motion search", "(host decisions)", "sub+fwd+quant+inv",
                                         "reconstruct+filters", "extend+ref pyramid" };
            printf("[dsv_enc device] %d P pictures:", g->gpu_frames);
            for (i = 0; i < 6; i++) printf(" %s %.2f;", gn[i], g->gpu_ms[i] / g->gpu_frames);
            printf("\n");
        }
    }
    if (g->ctx) {
        dsvcu_sync(g->ctx);
        for (i = 0; i < 2; i++) {
            if (g->src[i]) dsvcu_frame_destroy(g->ctx, g->src[i]);
            if (g->rec[i]) dsvcu_frame_destroy(g->ctx, g->rec[i]);
            if (g->src_pyr[i]) dsvcu_pyramid_destroy(g->ctx, g->src_pyr[i]);
        }
        if (g->ref_pyr) dsvcu_pyramid_destroy(g->ctx, g->ref_pyr);
        if (g->pred) dsvcu_frame_destroy(g->ctx, g->pred);
        if (g->coefs) dsvcu_coefs_destroy(g->ctx, g->coefs);
        dsvcu_ctx_destroy(g->ctx);
    }
    free(g->mvs);
    free(g->imv);
    free(g);
}

static DSV_ENCDATA *
gpu_state_get(DSV_ENCODER *enc, int nblk)
{
    DSV_ENCDATA *g = enc->ref;
    DSV_META *m = &enc->vidmeta;
    int i, ok = 1;
    if (g) {
        return g;
    }
    g = calloc(1, sizeof(*g));
    if (!g) {
        return NULL;
    }
    g->w = m->width;
    g->h = m->height;
    g->subsamp = m->subsamp;
    g->nblk = nblk;
    g->pyr_levels = enc->pyramid_levels;
    g->mvs = calloc((size_t) nblk, sizeof(DSV_MV));
    g->imv = calloc((size_t) nblk, sizeof(DSV_MV));
    ok = g->mvs && g->imv && !dsvcu_ctx_create(&g->ctx, dsv_get_thread_device(), m->width, m->height, m->subsamp);
    for (i = 0; ok && i < 2; i++) {
        ok = !dsvcu_frame_create(g->ctx, &g->src[i]) && !dsvcu_frame_create(g->ctx, &g->rec[i]) &&
             !dsvcu_pyramid_create(g->ctx, &g->src_pyr[i], enc->pyramid_levels);
    }
    ok = ok && !dsvcu_pyramid_create(g->ctx, &g->ref_pyr, enc->pyramid_levels) &&
         !dsvcu_frame_create(g->ctx, &g->pred) && !dsvcu_coefs_create(g->ctx, &g->coefs);
    if (!ok) {
        DSV_ERROR(("GPU encoder state: %s", dsvcu_last_error()));
        gpu_state_free(g);
        return NULL;
    }
    enc->ref = g;
    return g;
}

static void picture_geometry(DSV_ENCODER *enc, DSV_PARAMS *p);

/* hands the device buffers of a finished encoder to a fresh one (used by the
 * chunked drivers: a new encoder per closed-GOP chunk without re-allocating
 * device memory).  Only when everything the buffers were sized by agrees:
 * picture geometry, the number of blocks (block-size overrides change it) and
 * the resolved pyramid depth; otherwise the fresh encoder builds its own state */
void
dsv_enc_recycle(DSV_ENCODER *from, DSV_ENCODER *to)
{
    DSV_ENCDATA *g = from->ref;
    DSV_PARAMS p;
    if (!g || to->ref || g->w != to->vidmeta.width || g->h != to->vidmeta.height || g->subsamp != to->vidmeta.subsamp) {
        return;
    }
    memset(&p, 0, sizeof(p));
    picture_geometry(to, &p); /* resolves to->pyramid_levels exactly as the first picture would */
    if (p.nblocks_h * p.nblocks_v != g->nblk || to->pyramid_levels != g->pyr_levels) {
        return;
    }
    from->ref = NULL;
    g->cur = 0;
    g->have_ref = 0;
    g->ref_has_mvs = 0;
    memset(g->mvs, 0, (size_t) g->nblk * sizeof(DSV_MV));
    memset(g->imv, 0, (size_t) g->nblk * sizeof(DSV_MV));
    to->ref = g;
}

#define GPU(call)                                              \
    do {                                                       \
        if (call) {                                            \
            DSV_ERROR(("%s: %s", #call, dsvcu_last_error())); \
            return -1;                                         \
        }                                                      \
    } while (0)

/* the same once the picture's bit writer exists: it is released on the way out */
#define GPU_BW(call)                                           \
    do {                                                       \
        if (call) {                                            \
            DSV_ERROR(("%s: %s", #call, dsvcu_last_error())); \
            dsv_bw_free(&bw);                                  \
            return -1;                                         \
        }                                                      \
    } while (0)

/* ------------------------------------------------- quality -> quantiser */

/* piecewise-exponential curve sampled every 10 quality points
 * (reference dsv_encoder.c:72-88) */
static int
qp_curve_point(int v)
{
    const int unit = 10 * DSV_RC_QUAL_SCALE;
    int d = (100 * DSV_RC_QUAL_SCALE) - v;
    int oct = d / unit, frac = d % unit;
    int qp = (((unit - frac) * (1 << oct) + frac * (2 << oct)) / unit) - 1;
    return CLAMP(qp * 4, 0, DSV_MAX_QP);
}

/* reference dsv_encoder.c:90-106 */
static int
quality_to_qp(int v)
{
    int from_top = (100 * DSV_RC_QUAL_SCALE) - v;
    int third, frac;
    if (from_top < 60) {
        return from_top + 16; /* top of the range is linear */
    }
    third = (v * 2) / 3;
    frac = (v * 2) % 3;
    return (qp_curve_point(third) * (3 - frac) + frac * qp_curve_point(third + 1)) / 3;
}

/* ---------------------------------------------------- motion statistics */

/* reference avg_motion, dsv_encoder.c:129-176 */
static int
motion_field_summary(DSV_ENCODER *enc, DSV_MV *vecs, DSV_PARAMS *p)
{
    int nblk = p->nblocks_h * p->nblocks_v;
    int i, j, sx = 0, sy = 0, chaotic = 0, calm = 0, avg;

    for (j = 0; j < p->nblocks_v; j++) {
        for (i = 0; i < p->nblocks_h; i++) {
            DSV_MV *mv = &vecs[i + j * p->nblocks_h];
            int ndx, ndy;
            if (DSV_MV_IS_SKIP(mv)) {
                calm++;
                continue;
            }
            sx += mv->u.mv.x;
            sy += mv->u.mv.y;
            dsv_neighbordif2(vecs, p, i, j, &ndx, &ndy);
            if (ndx > 4 || ndy > 4) {
                chaotic++;
            } else {
                calm++;
            }
        }
    }
    avg = (abs(sx) + abs(sy)) / (nblk * 2);
    avg = MAX(avg, 1);
    enc->curr_avgmot = avg;
    enc->motion_static = calm * 100 / nblk;
    chaotic = chaotic * 100 / nblk;
    if (enc->prev_chaos < 0) {
        enc->prev_chaos = chaotic;
    } else {
        enc->prev_chaos = (enc->prev_chaos + enc->motion_chaos) / 2;
    }
    enc->motion_chaos = chaotic;
    return avg;
}

/* reference scene_complexity, dsv_encoder.c:179-250 */
static int
motion_field_complexity(DSV_ENCODER *enc, DSV_MV *vecs, DSV_PARAMS *p)
{
    int nblk = p->nblocks_h * p->nblocks_v;
    int k, cx = 0, ceiling;

    if (enc->rc_mode == DSV_RATE_CONTROL_ABR) {
        ceiling = dsv_mv_cost(vecs, p, 0, 0, 64, 64, enc->prev_quant, 0) + 12 + 64;
        ceiling = (ceiling * nblk + 1) >> 1;
        for (k = 0; k < nblk; k++) {
            DSV_MV *mv = &vecs[k];
            if (!DSV_MV_IS_SKIP(mv)) {
                cx += dsv_mv_cost(vecs, p, k % p->nblocks_h, k / p->nblocks_h, mv->u.mv.x, mv->u.mv.y,
                                  enc->prev_quant, 0);
                cx += (int) mv->err - (int) enc->avg_err;
            }
            if (DSV_MV_IS_INTRA(mv)) {
                cx += (mv->submask == DSV_MASK_ALL_INTRA) ? 16 : 4;
            }
        }
    } else if (enc->rc_mode == DSV_RATE_CONTROL_CRF) {
        ceiling = 70 * nblk;
        for (k = 0; k < nblk; k++) {
            DSV_MV *mv = &vecs[k];
            if (DSV_MV_IS_SKIP(mv)) {
                cx -= 100;
            } else {
                cx += dsv_mv_cost(vecs, p, k % p->nblocks_h, k / p->nblocks_h, mv->u.mv.x, mv->u.mv.y,
                                  enc->prev_quant, 0);
            }
            if (DSV_MV_IS_INTRA(mv)) {
                cx += (mv->submask == DSV_MASK_ALL_INTRA) ? 100 : 40;
            }
        }
    } else {
        return 0;
    }
    return cx <= 0 ? 0 : cx * 100 / ceiling;
}

/* decides whether a predicted picture must be coded as intra after all
 * (reference scene_change_detection, dsv_encoder.c:545-651).  Returns 1 and
 * clears p->has_ref when it must. */
static int
scene_cut_decision(DSV_ENCODER *enc, DSV_MV *vecs, DSV_PARAMS *p, DSV_FNUM fnum)
{
    int nblk = p->nblocks_h * p->nblocks_v;
    int intra_pct = enc->curr_intra_pct, scb = enc->curr_scblocks;
    int avgmot, chaos, dchaos, since_gop, cplx, close_fac, shift, sq_ipct, likely, score, cut, k;
    int cum_intra = 0, cum_skip = 0;

    avgmot = motion_field_summary(enc, vecs, p);
    chaos = enc->motion_chaos;
    dchaos = abs(chaos - enc->prev_chaos);
    since_gop = (int) fnum - (int) enc->prev_gop;
    cplx = motion_field_complexity(enc, vecs, p);
    close_fac = since_gop / MAX(abs(enc->gop) * 3 / 4, 1);
    if (cplx > 256 && chaos < 5) {
        shift = 9;
    } else if (cplx > chaos * 2) {
        shift = 8;
    } else if (cplx > chaos) {
        shift = 7;
    } else {
        shift = 6;
    }
    sq_ipct = ISQ(intra_pct) >> 5;
    likely = (intra_pct * 3 / 2 > scb) + (sq_ipct > scb);
    if (scb > enc->scene_change_pct && chaos < 34) {
        scb = ISQ(scb * 2) / MAX(enc->scene_change_pct, 1);
        likely++;
    } else {
        scb = ISQ(scb) / MAX(enc->scene_change_pct, 1);
    }
    shift = MAX(shift - likely, 5);
    score = MAX((dchaos / 16) + (enc->avg_err / 8), 1) * scb * MAX(cplx, 1) * MAX(close_fac, 1) >> (shift + 1);
    DSV_INFO(("frame %d: avg_err=%d avg_mot=%d chaos=%d%% complexity=%d intra=%d%% scene score=%d", (int) fnum,
              enc->avg_err, avgmot, chaos, cplx, intra_pct, score));

    cut = enc->do_scd && (score > 120 || (score > enc->scene_change_pct && avgmot < 20 &&
                                          enc->motion_chaos <= MAX(enc->prev_chaos - 10, 30)));
    if (cut || intra_pct > enc->intra_pct_thresh) {
        p->has_ref = 0;
        return 1;
    }
    enc->curr_complexity = cplx;

    /* blocks that have been intra at any time since the last intra picture */
    DSV_ASSERT(enc->intra_map);
    for (k = 0; k < nblk; k++) {
        DSV_MV *mv = &vecs[k];
        enc->intra_map[k] |= (uint8_t) !!DSV_MV_IS_INTRA(mv);
        if (enc->intra_map[k]) {
            if (DSV_MV_IS_SKIP(mv) || mv->u.all == 0) {
                int wgt = DSV_MV_IS_MAINTAIN(mv) ? 2 : 1;
                cum_intra += wgt * 2 - 1;
                cum_skip += wgt;
            } else if (DSV_MV_IS_NOXMITY(mv) && DSV_MV_IS_MAINTAIN(mv)) {
                cum_intra++;
            }
        }
        cum_intra += enc->intra_map[k];
    }
    cum_intra = cum_intra * 100 / nblk;
    cum_skip = cum_skip * 100 / nblk;
    if (cum_intra > enc->intra_pct_thresh && enc->curr_avgmot < 10 &&
        enc->motion_chaos <= CLAMP((enc->prev_chaos / 2) + cum_skip, 20, 40)) {
        DSV_INFO(("too much cumulative intra (%d%%): inserting an intra picture", cum_intra));
        p->has_ref = 0;
        return 1;
    }
    return 0;
}

/* ------------------------------------------------------- rate control */

/* constant-rate-factor quality for this picture (dsv_encoder.c:266-320) */
static int
rc_quality_crf(DSV_ENCODER *enc, DSV_PARAMS *p, int forced_intra, unsigned top_luma_avg)
{
    DSV_META *vf = p->vidmeta;
    const int bound = QPCT(25);
    int isP = p->has_ref;
    int lo = isP ? enc->min_quality : enc->min_I_frame_quality, hi = enc->max_quality;
    int anchor = CLAMP(enc->quality, lo, hi);
    int fps = (vf->fps_num << 5) / vf->fps_den;
    int gop = CLAMP(enc->gop, 1, (10 * fps >> 5));
    int calm = ISQ(enc->motion_static) / 75;
    int plex, target, q;

    if (calm < enc->motion_static) {
        calm = enc->motion_static;
    }
    if (!isP) {
        plex = (forced_intra ? 2 : 1) * calm - enc->motion_chaos;
    } else {
        plex = (ISQ(MIN(enc->avg_err, enc->motion_chaos / 3)) / 2) + calm - (3 * enc->motion_chaos);
    }
    plex = (plex * gop * vf->fps_den) / (vf->fps_num << 4);
    plex = CLAMP(plex, -bound / 4, bound / 4);
    target = (anchor + 3 * MAX(enc->rf_avg, enc->quality) + 2) >> 2;
    target = CLAMP(target, enc->quality - bound, enc->quality + bound);
    if (enc->do_dark_intra_boost && top_luma_avg < 80) {
        int step = (80 - (int) top_luma_avg) / 5;
        step = CLAMP(step, 5, 16) - 5;
        plex += ISQ(step) / 4;
    }
    q = target + plex;
    if (!isP) {
        int back = (DSV_RC_QUAL_MAX - q) / (1 + (enc->motion_chaos / 4));
        q += (back * gop * vf->fps_den) / (vf->fps_num << 4);
    }
    q = CLAMP(q, enc->quality - bound, enc->quality + bound);
    q = CLAMP(q, lo, hi);
    enc->rc_qual = (unsigned) MAX(q, 0);
    return q;
}

/* average-bitrate quality for this picture (dsv_encoder.c:321-454) */
static int
rc_quality_abr(DSV_ENCODER *enc, DSV_PARAMS *p, DSV_FNUM fnum, DSV_FNUM prev_I, unsigned top_luma_avg)
{
    DSV_META *vf = p->vidmeta;
    int isP = p->has_ref;
    int q = (int) enc->rc_qual;
    int fps = (vf->fps_num << 5) / vf->fps_den;
    int rf, want, dir, delta, floor_p, lo;

    if (fps == 0) {
        fps = 1;
    }
    if (enc->prev_complexity < 0) {
        enc->prev_complexity = enc->curr_complexity;
    }
    want = (int) (((enc->bitrate << 5) / (unsigned) fps) >> 3); /* bytes per picture */
    rf = enc->rf_avg ? enc->rf_avg : want;
    dir = (rf - want) > 0 ? -1 : 1;
    enc->min_q_step = CLAMP(enc->min_q_step, 1, DSV_RC_QUAL_MAX);
    enc->max_q_step = CLAMP(enc->max_q_step, 1, DSV_RC_QUAL_MAX);

    if (!isP) {
        unsigned miss = (unsigned) abs(rf - want);
        if (miss > 32768) {
            miss = 32768;
        }
        delta = (int) ((miss * miss) / (unsigned) ((dir > 0 ? 32 : 64) * want));
        if (delta > QPCT(12)) {
            delta -= QPCT(8);
        } else if (delta > QPCT(8)) {
            delta -= QPCT(4);
        } else if (delta > QPCT(4)) {
            delta -= QPCT(2);
        }
        delta = MIN(delta, QPCT(25));
        q = MAX(q, enc->avg_P_frame_q) + dir * delta;
        if (enc->prev_complexity < 15) {
            q += QPCT(2);
        } else if (enc->prev_complexity < 30) {
            q += QPCT(1);
        } else if (enc->prev_complexity > 40) {
            q -= QPCT(1);
        } else if (enc->prev_complexity > 60) {
            q -= QPCT(2);
        }
        enc->prev_I_frame_quality = q;
    } else {
        delta = (abs(rf - want) * QPCT(100)) / want;
        if (dir < 0 && delta < enc->min_q_step) {
            delta = 0;
        }
        delta = MIN(delta, enc->max_q_step * (dir > 0 ? 1 : 8));
        q += dir * delta;
    }
    floor_p = enc->avg_P_frame_q - QPCT(4);
    floor_p = CLAMP(floor_p, enc->min_quality, enc->max_quality);
    lo = isP ? floor_p : enc->min_I_frame_quality;
    if (enc->do_dark_intra_boost && !isP && top_luma_avg < 80) {
        int step = (80 - (int) top_luma_avg) / 5;
        q += CLAMP(step, 5, 16);
    }
    q = CLAMP(q, lo, enc->max_quality);
    q = CLAMP(q, 0, DSV_RC_QUAL_MAX);
    enc->rc_qual = (unsigned) q;
    enc->prev_complexity = enc->curr_complexity;

    if (enc->rc_pergop) {
        q = enc->prev_I_frame_quality;
        q = CLAMP(q, enc->min_quality, enc->max_quality);
    } else if (fnum > 0 && isP) {
        const int step = QPCT(8);
        int gop = CLAMP(enc->gop, 1, 60);
        int half = MAX(gop / 2, 1);
        int dist = abs((int) fnum - (int) prev_I), ramp, err_pen;
        if (dist >= enc->gop / 2) {
            dist = abs((int) fnum - ((int) prev_I + gop / 2));
            ramp = step - (step * dist / half);
        } else {
            ramp = step * dist / half;
        }
        q += CLAMP(ramp, 0, step) / 2;
        err_pen = CLAMP((enc->avg_err * enc->avg_err) >> 1, 0, QPCT(16));
        q -= err_pen;
        q = CLAMP(q, floor_p, enc->max_quality);
        if (enc->gop <= (2 * fps >> 5)) { /* short GOP: stay near the intra picture's quality */
            if (enc->prev_I_frame_quality < q) {
                q = enc->prev_I_frame_quality;
            } else {
                q = (3 * q + enc->prev_I_frame_quality) >> 2;
            }
            q = CLAMP(q, enc->min_quality, enc->max_quality);
        }
    }
    return q;
}

/* reference quality2quant, dsv_encoder.c:252-467 */
static int
pick_quantiser(DSV_ENCODER *enc, DSV_PARAMS *p, DSV_FNUM fnum, DSV_FNUM prev_I, int forced_intra, unsigned top_luma_avg)
{
    int q, quant;
    if (enc->rc_mode == DSV_RATE_CONTROL_CRF) {
        q = rc_quality_crf(enc, p, forced_intra, top_luma_avg);
    } else if (enc->rc_mode == DSV_RATE_CONTROL_ABR) {
        q = rc_quality_abr(enc, p, fnum, prev_I, top_luma_avg);
    } else {
        q = enc->quality;
        enc->rc_qual = (unsigned) q;
    }
    quant = p->lossless ? 1 : quality_to_qp(q);
    enc->prev_quant = quant;
    DSV_INFO(("frame quant = %d from quality (%d/%d)%%", quant, q, DSV_RC_QUAL_SCALE));
    return quant;
}

/* reference compute_auto_filter, dsv_encoder.c:518-543 */
static void
pick_loop_filter(DSV_ENCODER *enc, DSV_PARAMS *p, int quant)
{
    int chaos = enc->motion_chaos;
    int psy = dsv_spatial_psy_factor(p, -1);
    int norm = ISQ(quant) >> 15;
    int rel = (ISQ(enc->curr_intra_pct) + enc->curr_scblocks + enc->avg_err * chaos) / MAX(norm, 1);
    int avg_chaos = (enc->prev_chaos + chaos + 1) >> 1;
    int thresh = 8;
    rel += rel * psy >> 7;
    thresh += thresh * psy >> 5;
    thresh -= (MIN(avg_chaos, 48) * psy * MAX(enc->avg_err / 2, 1) / (128 * (thresh - 2)));
    enc->auto_filter = chaos <= 1 || rel > thresh;
}

/* ------------------------------------------------------ bit packing */

static void
put_packet_header(DSV_BITWR *bw, int type)
{
    dsv_bw_bits(bw, 8, DSV_FOURCC_0);
    dsv_bw_bits(bw, 8, DSV_FOURCC_1);
    dsv_bw_bits(bw, 8, DSV_FOURCC_2);
    dsv_bw_bits(bw, 8, DSV_FOURCC_3);
    dsv_bw_bits(bw, 8, DSV_VERSION_MINOR);
    dsv_bw_bits(bw, 8, (unsigned) type);
    dsv_bw_bits(bw, 32, 0); /* link to the previous packet, patched later */
    dsv_bw_bits(bw, 32, 0); /* link to the next packet */
}

static void
finish_packet(DSV_BITWR *bw, DSV_BUF *out)
{
    size_t n;
    dsv_bw_align(bw);
    n = dsv_bw_byte(bw);
    dsv_mk_buf(out, (int) n);
    memcpy(out->data, bw->buf, n);
    dsv_bw_free(bw);
}

static void
put_be32(uint8_t *p, unsigned v)
{
    p[0] = (uint8_t) (v >> 24);
    p[1] = (uint8_t) (v >> 16);
    p[2] = (uint8_t) (v >> 8);
    p[3] = (uint8_t) v;
}

/* B.1 link offsets (reference set_link_offsets, dsv_encoder.c:470-491) */
static void
link_packet(DSV_ENCODER *enc, DSV_BUF *buf, int is_eos)
{
    unsigned next = is_eos ? 0 : buf->len;
    put_be32(buf->data + DSV_PACKET_PREV_OFFSET, (unsigned) enc->prev_link);
    put_be32(buf->data + DSV_PACKET_NEXT_OFFSET, next);
    enc->prev_link = (int) next;
}

/* B.2.1 (reference encode_metadata, dsv_encoder.c:951-990) */
static void
make_metadata_packet(DSV_ENCODER *enc, DSV_BUF *out)
{
    DSV_BITWR bw;
    DSV_META *m = &enc->vidmeta;
    dsv_bw_init(&bw, 64);
    put_packet_header(&bw, DSV_PT_META);
    dsv_bw_ueg(&bw, (unsigned) m->width);
    dsv_bw_ueg(&bw, (unsigned) m->height);
    dsv_bw_ueg(&bw, (unsigned) m->subsamp);
    dsv_bw_ueg(&bw, (unsigned) m->fps_num);
    dsv_bw_ueg(&bw, (unsigned) m->fps_den);
    dsv_bw_ueg(&bw, (unsigned) m->aspect_num);
    dsv_bw_ueg(&bw, (unsigned) m->aspect_den);
    dsv_bw_ueg(&bw, (unsigned) m->inter_sharpen);
    dsv_bw_bit(&bw, 0); /* no reserved bits */
    finish_packet(&bw, out);
    put_be32(out->data + DSV_PACKET_NEXT_OFFSET, out->len);
}

/* appends a length-prefixed byte-aligned sub-stream */
static void
put_substream(DSV_BITWR *bw, const uint8_t *data, size_t bytes)
{
    dsv_bw_align(bw);
    dsv_bw_ueg(bw, (unsigned) bytes);
    dsv_bw_align(bw);
    dsv_bw_bytes(bw, data, bytes);
}

static void
put_rle_substream(DSV_BITWR *bw, DSV_RLEWR *r)
{
    size_t n = dsv_rle_wr_end(r);
    put_substream(bw, r->bw.buf, n);
    dsv_bw_free(&r->bw);
}

/* stability downscale from the frame rate (dsv_encoder.c:821-837) */
static int
stability_shift(const DSV_META *m)
{
    int fps = DSV_UDIV_ROUND(m->fps_num, m->fps_den);
    if (fps <= 24) return 6;
    if (fps <= 30) return 4;
    if (fps <= 60) return 2;
    return 0;
}

static int
block_is_still(DSV_ENCODER *enc, int i, int div)
{
    return (enc->stability[i].x / div) == 0 && (enc->stability[i].y / div) == 0;
}

/* signalling polarity of the run-length coded flag planes (gather_stats,
 * dsv_encoder.c:992-1037 + :1085-1095) */
static void
choose_markers(DSV_ENCODER *enc, DSV_PARAMS *p, DSV_FNUM fnum, DSV_MV *mvs, DSV_MV *imv, int *stats)
{
    int nblk = p->nblocks_h * p->nblocks_v;
    int i, div;
    for (i = 0; i < DSV_MAX_STAT; i++) {
        stats[i] = DSV_ONE_MARKER;
    }
    if (enc->effort < 7) {
        stats[DSV_MAINTAIN_STAT] = DSV_ZERO_MARKER;
        stats[DSV_RINGING_STAT] = DSV_ZERO_MARKER;
        return;
    }
    div = (enc->refresh_ctr >= enc->stable_refresh) ? 0 : (int) enc->refresh_ctr;
    if (div <= 0) {
        div = 1;
    }
    for (i = 0; i < nblk; i++) {
        int still;
        if (p->has_ref) {
            DSV_MV *mv = &mvs[i];
            still = !DSV_MV_IS_INTRA(mv) && DSV_MV_IS_SKIP(mv);
            if (!DSV_MV_IS_SKIP(mv)) {
                stats[DSV_MODE_STAT] += DSV_MV_IS_INTRA(mv) ? 1 : -1;
                stats[DSV_EPRM_STAT] += DSV_MV_IS_EPRM(mv) ? 1 : -1;
            }
        } else {
            DSV_MV *mv = &imv[i];
            if (fnum > 0 && enc->do_temporal_aq) {
                still = block_is_still(enc, i, div);
            } else {
                still = !!DSV_MV_IS_SKIP(mv);
            }
            stats[DSV_MAINTAIN_STAT] += DSV_MV_IS_MAINTAIN(mv) ? 1 : -1;
            stats[DSV_RINGING_STAT] += DSV_MV_IS_RINGING(mv) ? 1 : -1;
        }
        stats[DSV_STABLE_STAT] += still ? 1 : -1;
    }
    for (i = 0; i < DSV_MAX_STAT; i++) {
        stats[i] = stats[i] > 0 ? DSV_ZERO_MARKER : DSV_ONE_MARKER;
    }
}

/* B.2.3.1: skip (P) / stable (I) flags; also resets blockdata[] and clears the
 * vectors of skipped blocks (reference encode_stable_blocks, :796-883) */
static void
code_stability(DSV_ENCODER *enc, DSV_PARAMS *p, DSV_FNUM fnum, DSV_MV *mvs, DSV_MV *imv, const int *stats, DSV_BITWR *bw)
{
    int nblk = p->nblocks_h * p->nblocks_v;
    int i, div, shift = stability_shift(p->vidmeta);
    DSV_RLEWR rle;

    dsv_rle_wr_init(&rle, (size_t) nblk / 4 + 64);
    if (enc->refresh_ctr >= enc->stable_refresh) {
        enc->refresh_ctr = 0;
        memset(enc->stability, 0, sizeof(*enc->stability) * (size_t) nblk);
    }
    div = (int) enc->refresh_ctr;
    if (div <= 0) {
        div = 1;
    }
    for (i = 0; i < nblk; i++) {
        int still;
        if (p->has_ref) {
            DSV_MV *mv = &mvs[i];
            uint8_t bd = 0;
            if (DSV_MV_IS_SKIP(mv)) {
                mv->u.all = 0;
            }
            if (DSV_MV_IS_INTRA(mv)) {
                still = 0;
                bd |= DSV_IS_INTRA;
            } else {
                still = !!DSV_MV_IS_SKIP(mv);
                if (!still) {
                    enc->stability[i].x += abs(mv->u.mv.x) >> shift;
                    enc->stability[i].y += abs(mv->u.mv.y) >> shift;
                }
            }
            bd |= (uint8_t) (still << DSV_SKIP_BIT);
            if (DSV_MV_IS_SIMCMPLX(mv)) {
                bd |= DSV_IS_SIMCMPLX;
            }
            enc->blockdata[i] = bd;
        } else {
            still = (fnum > 0 && enc->do_temporal_aq) ? block_is_still(enc, i, div) : 0;
            still |= !!DSV_MV_IS_SKIP(&imv[i]);
            enc->blockdata[i] = (uint8_t) (still << DSV_STABLE_BIT);
        }
        dsv_rle_wr_put(&rle, stats[DSV_STABLE_STAT] == DSV_ONE_MARKER ? still : !still);
    }
    put_rle_substream(bw, &rle);
}

/* B.2.3.2 / B.2.3.3 (reference encode_intra_meta, :886-932) */
static void
code_intra_flags(DSV_ENCODER *enc, DSV_PARAMS *p, DSV_MV *imv, const int *stats, DSV_BITWR *bw)
{
    int nblk = p->nblocks_h * p->nblocks_v, i;
    DSV_RLEWR ring, keep;
    dsv_rle_wr_init(&ring, (size_t) nblk / 4 + 64);
    dsv_rle_wr_init(&keep, (size_t) nblk / 4 + 64);
    for (i = 0; i < nblk; i++) {
        int r = !!DSV_MV_IS_RINGING(&imv[i]), m = !!DSV_MV_IS_MAINTAIN(&imv[i]);
        enc->blockdata[i] |= (uint8_t) ((r << DSV_RINGING_BIT) | (m << DSV_MAINTAIN_BIT));
        dsv_rle_wr_put(&ring, stats[DSV_RINGING_STAT] == DSV_ONE_MARKER ? r : !r);
        dsv_rle_wr_put(&keep, stats[DSV_MAINTAIN_STAT] == DSV_ONE_MARKER ? m : !m);
    }
    put_rle_substream(bw, &ring);
    put_rle_substream(bw, &keep);
}

/* B.2.3.4 motion data: five sub-streams (reference encode_motion, :692-794) */
static void
code_motion(DSV_ENCODER *enc, DSV_PARAMS *p, DSV_MV *mvs, const int *stats, DSV_BITWR *bw)
{
    DSV_BITWR vx, vy, sbim;
    DSV_RLEWR mode, eprm;
    int i, j, nblk = p->nblocks_h * p->nblocks_v;

    dsv_bw_init(&vx, (size_t) nblk * 2 + 64);
    dsv_bw_init(&vy, (size_t) nblk * 2 + 64);
    dsv_bw_init(&sbim, (size_t) nblk + 64);
    dsv_rle_wr_init(&mode, (size_t) nblk / 4 + 64);
    dsv_rle_wr_init(&eprm, (size_t) nblk / 4 + 64);
    for (j = 0; j < p->nblocks_v; j++) {
        for (i = 0; i < p->nblocks_h; i++) {
            int idx = i + j * p->nblocks_h;
            DSV_MV *mv = &mvs[idx];
            int is_eprm = !!DSV_MV_IS_EPRM(mv), is_intra = !!DSV_MV_IS_INTRA(mv);
            int px, py, cx, cy;

            enc->blockdata[idx] |= (uint8_t) (is_eprm << DSV_EPRM_BIT);
            if (DSV_MV_IS_SKIP(mv)) {
                enc->blockdata[idx] |= DSV_IS_STABLE;
                continue;
            }
            dsv_movec_pred(mvs, p, i, j, &px, &py);
            if (is_intra) {
                /* intra blocks carry full-pel vectors */
                px = DSV_SAR_R(px, 2);
                py = DSV_SAR_R(py, 2);
                cx = DSV_SAR(mv->u.mv.x, 2);
                cy = DSV_SAR(mv->u.mv.y, 2);
                mv->u.mv.x = (int16_t) (cx * 4);
                mv->u.mv.y = (int16_t) (cy * 4);
                if (mv->submask == DSV_MASK_ALL_INTRA) {
                    dsv_bw_bit(&sbim, 1);
                } else {
                    dsv_bw_bit(&sbim, 0);
                    dsv_bw_bits(&sbim, 4, mv->submask);
                }
                if (mv->dc & DSV_SRC_DC_PRED) {
                    dsv_bw_bit(&sbim, 1);
                    dsv_bw_bits(&sbim, 8, mv->dc & 0xff);
                } else {
                    dsv_bw_bit(&sbim, 0);
                }
            } else {
                cx = mv->u.mv.x;
                cy = mv->u.mv.y;
            }
            dsv_bw_seg(&vx, cx - px);
            dsv_bw_seg(&vy, cy - py);
            if (dsv_neighbordif(mvs, p, i, j) > DSV_NDIF_THRESH) {
                enc->blockdata[idx] |= DSV_IS_STABLE;
            }
            dsv_rle_wr_put(&mode, stats[DSV_MODE_STAT] == DSV_ONE_MARKER ? is_intra : !is_intra);
            dsv_rle_wr_put(&eprm, stats[DSV_EPRM_STAT] == DSV_ONE_MARKER ? is_eprm : !is_eprm);
        }
    }
    /* order: DSV_SUB_MODE, MV_X, MV_Y, SBIM, EPRM */
    put_rle_substream(bw, &mode);
    dsv_bw_align(&vx);
    put_substream(bw, vx.buf, dsv_bw_byte(&vx));
    dsv_bw_align(&vy);
    put_substream(bw, vy.buf, dsv_bw_byte(&vy));
    dsv_bw_align(&sbim);
    put_substream(bw, sbim.buf, dsv_bw_byte(&sbim));
    put_rle_substream(bw, &eprm);
    dsv_bw_free(&vx);
    dsv_bw_free(&vy);
    dsv_bw_free(&sbim);
}

/* ---------------------------------------------------------- one picture */

static int
default_block_dim(int dim)
{
    return dim > 1280 ? DSV_MAX_BLOCK_SIZE : DSV_MIN_BLOCK_SIZE;
}

/* block geometry + pyramid depth (reference encode_one_frame, :1200-1241) */
static void
picture_geometry(DSV_ENCODER *enc, DSV_PARAMS *p)
{
    int w = enc->vidmeta.width, h = enc->vidmeta.height;
    p->blk_w = default_block_dim(w);
    p->blk_h = default_block_dim(h);
    if (abs(w - h) < MIN(w, h)) {
        p->blk_w = p->blk_h = MIN(p->blk_w, p->blk_h);
    }
    if (enc->block_size_override_x >= 0) {
        p->blk_w = 16 << enc->block_size_override_x;
    }
    if (enc->block_size_override_y >= 0) {
        p->blk_h = 16 << enc->block_size_override_y;
    }
    p->blk_w = CLAMP(p->blk_w, DSV_MIN_BLOCK_SIZE, DSV_MAX_BLOCK_SIZE);
    p->blk_h = CLAMP(p->blk_h, DSV_MIN_BLOCK_SIZE, DSV_MAX_BLOCK_SIZE);
    p->nblocks_h = DSV_UDIV_ROUND_UP(w, p->blk_w);
    p->nblocks_v = DSV_UDIV_ROUND_UP(h, p->blk_h);
    if (enc->pyramid_levels == 0) {
        int lv = dsv_lb2((unsigned) MIN(w, h));
        int most = MAX(p->nblocks_h, p->nblocks_v);
        while ((1 << lv) > most) {
            lv--;
        }
        enc->pyramid_levels = CLAMP(lv, 3, DSV_MAX_PYRAMID_LEVELS);
    }
}

/* returns 1 when a metadata packet must precede this picture, -1 on error */
static int
encode_picture(DSV_ENCODER *enc, DSV_FRAME *frame, DSV_FNUM fnum, DSV_BUF *out, DSV_PARAMS *p, DSV_ENCDATA **pg)
{
    DSV_ENCDATA *g;
    DSV_BITWR bw;
    dsvcu_fmeta fm;
    dsvcu_hme_params hp;
    dsvcu_frame *src, *rec, *ref_src, *ref_rec;
    DSV_FNUM prev_I = enc->prev_gop;
    unsigned top_avg = 0;
    int stats[DSV_MAX_STAT];
    int gop_start = 0, forced_intra = 0, tried_motion = 0, quant, nblk, i, inter_filter = 0;

    memset(p, 0, sizeof(*p));
    p->vidmeta = &enc->vidmeta;
    p->effort = enc->effort;
    p->do_psy = enc->do_psy;
    p->temporal_mc = (int) DSV_TEMPORAL_MC(fnum);
    p->lossless = (enc->quality == DSV_RC_QUAL_MAX);
    picture_geometry(enc, p);
    nblk = p->nblocks_h * p->nblocks_v;
    if (enc->stability == NULL) {
        enc->stability = dsv_alloc((int) sizeof(*enc->stability) * nblk);
        enc->blockdata = dsv_alloc(nblk);
    }
    g = gpu_state_get(enc, nblk);
    if (!g) {
        return -1;
    }
    *pg = g;
    if (g_prof < 0) {
        g_prof = getenv("DSV_PROFILE") ? atoi(getenv("DSV_PROFILE")) : 0;
    }
    g->ph_t0 = g_prof > 0 ? now_ms() : 0;
    g->ph_frames++;
    if (g_prof > 1) {
        if (g->marks_live) {
            for (i = 0; i < 6; i++) {
                float ms = 0.f;
                dsvcu_mark_elapsed_ms(g->ctx, i, i + 1, &ms);
                g->gpu_ms[i] += ms;
            }
            g->gpu_frames++;
            g->marks_live = 0;
        }
        dsvcu_mark(g->ctx, 0);
    }
    src = g->src[g->cur];
    rec = g->rec[g->cur];
    ref_src = g->src[g->cur ^ 1];
    ref_rec = g->rec[g->cur ^ 1];

    /* source picture -> device, padded; luma pyramid; top-level brightness */
    for (i = 0; i < 3; i++) {
        GPU(dsvcu_frame_upload(g->ctx, src, i, frame->planes[i].data, frame->planes[i].stride));
    }
    GPU(dsvcu_extend_pyramid(g->ctx, src, g->src_pyr[g->cur]));

    if (enc->force_metadata || ((enc->prev_gop + (DSV_FNUM) enc->gop) <= fnum)) {
        gop_start = 1;
        enc->prev_gop = fnum;
        enc->force_metadata = 0;
    }
    if (enc->gop == DSV_GOP_INTRA) {
        p->is_ref = 0;
        p->has_ref = 0;
    } else {
        p->is_ref = 1;
        p->has_ref = !gop_start;
    }
    enc->avg_err = 0;
    if (!enc->intra_map) {
        enc->intra_map = dsv_alloc(nblk);
    }

    dsv_fmeta_from_params(&fm, p, p->has_ref, fnum);
    if (p->has_ref) {
        if (!g->have_ref) {
            DSV_ERROR(("no reference picture for a predicted picture"));
            return -1;
        }
        hp.quant = enc->prev_quant;
        hp.skip_block_thresh = enc->skip_block_thresh;
        hp.pyramid_levels = enc->pyramid_levels;
        hp.use_prev_mvs = g->ref_has_mvs;
        if (g_prof > 1) dsvcu_mark(g->ctx, 1);
        GPU(dsvcu_hme(g->ctx, &fm, &hp, src, g->src_pyr[g->cur], ref_rec, g->ref_pyr, ref_src, g->src_pyr[g->cur ^ 1]));
        tried_motion = 1;
        if (g_prof > 1) dsvcu_mark(g->ctx, 2);
    } else {
        GPU(dsvcu_intra_analysis_async(g->ctx, &fm, src, nblk));
    }
    if (enc->do_dark_intra_boost && enc->rc_mode != DSV_RATE_CONTROL_CQP) {
        GPU(dsvcu_frame_luma_avg_async(g->ctx, dsvcu_pyramid_level(g->src_pyr[g->cur], enc->pyramid_levels)));
    }
    PROF_MARK(g, PH_UPLOAD);
    /* first (and for most pictures only) wait on analysis results */
    if (tried_motion) {
        GPU(dsvcu_hme_fetch(g->ctx, g->mvs, nblk, &enc->curr_intra_pct, &enc->curr_scblocks, &enc->avg_err));
        forced_intra = scene_cut_decision(enc, g->mvs, p, fnum);
        if (!p->has_ref) {
            fm.isP = 0;
            GPU(dsvcu_intra_analysis_async(g->ctx, &fm, src, nblk));
        }
    }
    if (!p->has_ref) {
        GPU(dsvcu_intra_analysis_fetch(g->ctx, g->imv, nblk));
    }
    PROF_MARK(g, PH_ANALYSIS_WAIT);
    top_avg = dsvcu_frame_luma_avg_result(g->ctx);
    if (enc->variable_i_interval && forced_intra) {
        enc->prev_gop = fnum;
    }
    if (!p->has_ref) {
        memset(enc->intra_map, 0, (size_t) nblk);
    }
    quant = pick_quantiser(enc, p, fnum, prev_I, forced_intra, top_avg);
    pick_loop_filter(enc, p, quant);
    dsv_fmeta_from_params(&fm, p, p->has_ref, fnum);

    PROF_MARK(g, PH_DECIDE);
    /* ---- picture packet: header and block side information (host) ---- */
    dsv_bw_init(&bw, (size_t) nblk * 8 + 4096);
    put_packet_header(&bw, DSV_MAKE_PT(p->is_ref, p->has_ref));
    dsv_bw_align(&bw);
    dsv_bw_bits(&bw, 32, fnum);
    choose_markers(enc, p, fnum, g->mvs, g->imv, stats);
    dsv_bw_align(&bw);
    dsv_bw_ueg(&bw, (unsigned) (dsv_lb2((unsigned) p->blk_w) - 4));
    dsv_bw_ueg(&bw, (unsigned) (dsv_lb2((unsigned) p->blk_h) - 4));
    dsv_bw_align(&bw);
    dsv_bw_bit(&bw, stats[DSV_STABLE_STAT]);
    if (p->has_ref) {
        dsv_bw_bit(&bw, stats[DSV_MODE_STAT]);
        dsv_bw_bit(&bw, stats[DSV_EPRM_STAT]);
        inter_filter = enc->do_inter_filter == 1 || (enc->do_inter_filter == -1 && enc->auto_filter);
        dsv_bw_bit(&bw, inter_filter);
    } else {
        dsv_bw_bit(&bw, stats[DSV_MAINTAIN_STAT]);
        dsv_bw_bit(&bw, stats[DSV_RINGING_STAT]);
        dsv_bw_bit(&bw, enc->do_intra_filter);
    }
    dsv_bw_bits(&bw, DSV_MAX_QP_BITS, (unsigned) quant);
    dsv_bw_bit(&bw, 0); /* no per-picture reserved bits */
    dsv_bw_align(&bw);
    code_stability(enc, p, fnum, g->mvs, g->imv, stats, &bw);
    if (p->has_ref) {
        dsv_bw_align(&bw);
        code_motion(enc, p, g->mvs, stats, &bw);
    } else {
        code_intra_flags(enc, p, g->imv, stats, &bw);
    }
    dsv_bw_align(&bw);

    PROF_MARK(g, PH_SIDEINFO);
    /* ---- pixel pipeline (device), queued in one go ---- */
    GPU_BW(dsvcu_set_side(g->ctx, enc->blockdata, p->has_ref ? g->mvs : NULL, nblk));
    if (tried_motion && !p->has_ref) {
        /* a picture that went through the search but is coded intra: the field as the
         * host left it is the next picture's temporal predictor
         * (DSV_HME.ref_mvf = reference picture's final_mvs) */
        GPU_BW(dsvcu_set_prev_mvs(g->ctx, g->mvs, nblk));
    }
    if (g_prof > 1 && p->has_ref) dsvcu_mark(g->ctx, 3);
    /* the reference clones the padded source into the residual frame and works in
     * place (dsv_encoder.c:1292); here an intra picture is transformed straight
     * from the source, a predicted one has its residual written into `rec` */
    if (p->has_ref) {
        GPU_BW(dsvcu_sub_pred_from(g->ctx, &fm, g->pred, rec, ref_rec, src));
        GPU_BW(dsvcu_fwd_sbt_frame(g->ctx, rec, g->coefs, &fm, 7));
    } else {
        GPU_BW(dsvcu_fwd_sbt_frame(g->ctx, src, g->coefs, &fm, 7));
    }
    GPU_BW(dsvcu_quant_frame(g->ctx, g->coefs, quant, &fm, 7));
    GPU_BW(dsvcu_inv_sbt_frame(g->ctx, rec, g->coefs, quant, &fm, 7));
    if (!p->has_ref) {
        GPU_BW(dsvcu_intra_filter(g->ctx, quant, &fm, 0, rec, enc->do_intra_filter));
    }
    if (p->has_ref) {
        if (g_prof > 1) dsvcu_mark(g->ctx, 4);
        GPU_BW(dsvcu_add_res(g->ctx, &fm, quant, rec, g->pred, inter_filter));
        if (g_prof > 1) dsvcu_mark(g->ctx, 5);
    }
    if (p->is_ref) {
        GPU_BW(dsvcu_extend_pyramid(g->ctx, rec, g->ref_pyr));
    }
    if (g_prof > 1 && p->has_ref) {
        dsvcu_mark(g->ctx, 6);
        g->marks_live = 1;
    }
    if (tried_motion && p->has_ref) {
        /* every reader of this picture's field is queued: it becomes the next
         * picture's temporal predictor by exchanging buffers, not by copying */
        GPU_BW(dsvcu_mvs_swap_prev(g->ctx, nblk));
    }

    PROF_MARK(g, PH_QUEUE);
    /* ---- coefficient planes: pack symbols as they arrive ---- */
    for (i = 0; i < 3; i++) {
        const dsvcu_symbol *syms;
        int nsym, dc, cw, ch;
        GPU_BW(dsvcu_fetch_symbols(g->ctx, i, &syms, &nsym, &dc));
        PROF_MARK(g, PH_SYM_WAIT);
        dsvcu_coefs_plane_dims(g->coefs, i, &cw, &ch);
        dsv_hzcc_write_plane(&bw, syms, nsym, dc, cw, ch);
        PROF_MARK(g, PH_PACK);
    }
    finish_packet(&bw, out);

    if (enc->frame_callback) {
        DSV_FRAME *hrec = dsv_mk_frame(g->subsamp, g->w, g->h, 0);
        for (i = 0; i < 3; i++) {
            GPU_BW(dsvcu_frame_download(g->ctx, rec, i, hrec->planes[i].data, hrec->planes[i].stride));
        }
        GPU_BW(dsvcu_sync(g->ctx));
        enc->frame_callback(&enc->vidmeta, frame, hrec);
        dsv_frame_ref_dec(hrec);
    }
    if (p->is_ref) {
        g->have_ref = 1;
        g->ref_has_mvs = tried_motion;
        g->cur ^= 1;
    }
    return gop_start;
}

/* ------------------------------------------------------------ public API */

void
dsv_enc_init(DSV_ENCODER *enc)
{
    memset(enc, 0, sizeof(*enc));
    enc->prev_gop = (DSV_FNUM) -1;
    enc->quality = DSV_QUALITY_PERCENT(80);
    enc->gop = 48;
    enc->effort = DSV_MAX_EFFORT;
    enc->rc_mode = DSV_RATE_CONTROL_CRF;
    enc->bitrate = INT_MAX;
    enc->min_q_step = 4;
    enc->max_q_step = 1;
    enc->min_quality = enc->quality - DSV_USER_QUAL_TO_RC_QUAL(5);
    enc->max_quality = DSV_RC_QUAL_MAX;
    enc->min_I_frame_quality = enc->quality - DSV_USER_QUAL_TO_RC_QUAL(2);
    enc->prev_chaos = -1;
    enc->prev_complexity = -1;
    enc->curr_complexity = -1;
    enc->intra_pct_thresh = 90;
    enc->stable_refresh = 24;
    enc->scene_change_pct = 85;
    enc->do_scd = 1;
    enc->variable_i_interval = 1;
    enc->block_size_override_x = -1;
    enc->block_size_override_y = -1;
    enc->do_temporal_aq = 1;
    enc->do_psy = DSV_PSY_ALL;
    enc->do_dark_intra_boost = 1;
    enc->do_intra_filter = 1;
    enc->do_inter_filter = -1;
}

void
dsv_enc_start(DSV_ENCODER *enc)
{
    enc->quality = CLAMP(enc->quality, 0, DSV_RC_QUAL_MAX);
    if (enc->rc_mode == DSV_RATE_CONTROL_CRF) {
        enc->rc_qual = (unsigned) CLAMP(enc->quality + QPCT(5), enc->min_I_frame_quality, enc->max_quality);
        enc->rf_avg = (int) enc->rc_qual;
        enc->avg_P_frame_q = enc->quality;
    } else if (enc->rc_mode == DSV_RATE_CONTROL_ABR) {
        enc->rc_qual = (unsigned) enc->quality;
        enc->avg_P_frame_q = enc->quality * 4 / 5;
    }
    enc->stats.iminq = enc->stats.pminq = INT_MAX;
    enc->stats.imins = enc->stats.pmins = INT_MAX;
    enc->force_metadata = 1;
}

void
dsv_enc_free(DSV_ENCODER *enc)
{
    if (enc->ref) {
        gpu_state_free(enc->ref);
        enc->ref = NULL;
    }
    if (enc->stability) {
        dsv_free(enc->stability);
        enc->stability = NULL;
    }
    if (enc->blockdata) {
        dsv_free(enc->blockdata);
        enc->blockdata = NULL;
    }
    if (enc->intra_map) {
        dsv_free(enc->intra_map);
        enc->intra_map = NULL;
    }
}

void
dsv_enc_set_metadata(DSV_ENCODER *enc, DSV_META *md)
{
    memcpy(&enc->vidmeta, md, sizeof(DSV_META));
}

void
dsv_enc_force_metadata(DSV_ENCODER *enc)
{
    enc->force_metadata = 1;
}

/* B.2.2 */
void
dsv_enc_end_of_stream(DSV_ENCODER *enc, DSV_BUF *bufs)
{
    DSV_BITWR bw;
    dsv_bw_init(&bw, DSV_PACKET_HDR_SIZE + 8);
    put_packet_header(&bw, DSV_PT_EOS);
    finish_packet(&bw, &bufs[0]);
    link_packet(enc, &bufs[0], 1);
}

/* bookkeeping after a picture (reference dsv_enc, dsv_encoder.c:1471-1570) */
static void
account_picture(DSV_ENCODER *enc, DSV_PARAMS *p, DSV_MV *mvs, unsigned bytes)
{
    struct DSV_STATS *st = &enc->stats;
    int k, nblk = p->nblocks_h * p->nblocks_v;
    if (p->has_ref) {
        st->pnum++;
        st->pfnum += !!enc->auto_filter;
        st->psize += bytes;
        st->pqual += enc->rc_qual;
        st->pmaxq = MAX(enc->rc_qual, st->pmaxq);
        st->pmaxs = MAX(bytes, st->pmaxs);
        st->pminq = MIN(enc->rc_qual, st->pminq);
        st->pmins = MIN(bytes, st->pmins);
        for (k = 0; k < nblk; k++) {
            DSV_MV *mv = &mvs[k];
            if (DSV_MV_IS_EPRM(mv)) {
                st->eprm++;
            }
            if (DSV_MV_IS_SKIP(mv)) {
                st->skip++;
            } else if (DSV_MV_IS_INTRA(mv)) {
                st->mbI++;
                st->mbdc += !!(mv->dc & DSV_SRC_DC_PRED);
                if (mv->submask != DSV_MASK_ALL_INTRA) {
                    int b;
                    st->mbsub++;
                    for (b = 0; b < 4; b++) {
                        st->mbsubs[b] += (mv->submask >> b) & 1;
                    }
                }
            } else {
                int x = mv->u.mv.x, y = mv->u.mv.y;
                st->mbP++;
                if (x & 1) st->qpx++; else if (x & 3) st->hpx++; else st->fpx++;
                if (y & 1) st->qpy++; else if (y & 3) st->hpy++; else st->fpy++;
            }
        }
        st->mb += (unsigned) nblk;
        enc->refresh_ctr++;
    } else {
        st->inum++;
        st->ifnum += !!enc->do_intra_filter;
        st->isize += bytes;
        st->iqual += enc->rc_qual;
        st->imaxq = MAX(enc->rc_qual, st->imaxq);
        st->imaxs = MAX(bytes, st->imaxs);
        st->iminq = MIN(enc->rc_qual, st->iminq);
        st->imins = MIN(bytes, st->imins);
    }
    if (enc->rc_mode != DSV_RATE_CONTROL_CQP) {
        enc->rf_total += (enc->rc_mode == DSV_RATE_CONTROL_CRF) ? enc->rc_qual : bytes;
        enc->rf_reset++;
        if (p->has_ref) {
            enc->total_P_frame_q += (int) enc->rc_qual;
            enc->avg_P_frame_q = enc->total_P_frame_q / (int) enc->rf_reset;
        }
        enc->rf_avg = (int) (enc->rf_total / enc->rf_reset);
        if (enc->rf_reset >= DSV_RF_RESET) {
            enc->rf_total = (unsigned) enc->rf_avg;
            enc->total_P_frame_q = enc->total_P_frame_q / (int) enc->rf_reset;
            enc->rf_reset = 1;
        }
    }
}

int
dsv_enc(DSV_ENCODER *enc, DSV_FRAME *frame, DSV_BUF *bufs)
{
    DSV_PARAMS prm;
    DSV_ENCDATA *g = NULL;
    DSV_BUF pic;
    DSV_FNUM fnum;
    int nbuf = 0, r;

    if (frame == NULL) {
        DSV_ERROR(("null frame passed to encoder!"));
        return 0;
    }
    if (bufs == NULL) {
        DSV_ERROR(("null buffer list passed to encoder!"));
        return 0;
    }
    fnum = enc->next_fnum++;
    memset(&pic, 0, sizeof(pic));
    r = encode_picture(enc, frame, fnum, &pic, &prm, &g);
    dsv_frame_ref_dec(frame);
    if (r < 0) {
        dsv_buf_free(&pic);
        return 0;
    }
    if (r) {
        make_metadata_packet(enc, &bufs[nbuf]);
        link_packet(enc, &bufs[nbuf], 0);
        nbuf++;
    }
    bufs[nbuf] = pic;
    link_packet(enc, &bufs[nbuf], 0);
    nbuf++;
    account_picture(enc, &prm, g->mvs, pic.len);
    return nbuf;
}
