/*
 * refops.c -- TEST INFRASTRUCTURE ONLY.  Flat-array entry points around the
 * UNMODIFIED reference operators, so the parity tests can feed one operator at
 * a time with the same inputs they give to the CUDA path.
 *
 * Built by oracle/Makefile against the reference headers where they lie
 * (-I/root/reference/src) and linked with oracle/_ref/libdsvref.so; the result
 * is oracle/_ref/librefops.so.  No reference source is copied.  Nothing here is
 * linked into or called by the product library.
 *
 * Frames cross this boundary as tightly packed planar YUV (visible area only,
 * Y then U then V); the harness builds bordered reference frames with the
 * reference's own dsv_mk_frame / dsv_extend_frame.
 *
 * Operators wrapped (reference file:line):
 *   dsv_hme              hme.c:2001-2016   (+ mk_pyramid, dsv_encoder.c:493-516)
 *   dsv_intra_analysis   hme.c:1835-1971
 *   dsv_fwd_sbt          sbt.c:847-886
 *   dsv_inv_sbt          sbt.c:889-934
 *   dsv_encode_plane     hzcc.c:585-613
 *   dsv_sub_pred         bmc.c:1057-1070
 *   dsv_add_res          bmc.c:1072-1090
 *   dsv_intra_filter     bmc.c:390-457
 */
#include <stdlib.h>
#include <string.h>
#include "dsv_encoder.h"
#include "dsv_internal.h"

typedef struct {
    int w, h, subsamp;
    int fps_num, fps_den;
    int effort, do_psy;
    int blk_w, blk_h;
    int temporal_mc, lossless, inter_sharpen;
    int skip_thresh, pyramid_levels;
    int isP;
    unsigned fnum;
} refop_cfg;

static DSV_META g_meta;

static void
mk_params(const refop_cfg *c, DSV_PARAMS *p)
{
    memset(&g_meta, 0, sizeof(g_meta));
    g_meta.width = c->w;
    g_meta.height = c->h;
    g_meta.subsamp = c->subsamp;
    g_meta.fps_num = c->fps_num;
    g_meta.fps_den = c->fps_den;
    g_meta.aspect_num = g_meta.aspect_den = 1;
    g_meta.inter_sharpen = c->inter_sharpen;
    memset(p, 0, sizeof(*p));
    p->vidmeta = &g_meta;
    p->effort = c->effort;
    p->do_psy = c->do_psy;
    p->is_ref = 1;
    p->has_ref = c->isP;
    p->blk_w = c->blk_w;
    p->blk_h = c->blk_h;
    p->nblocks_h = (c->w + c->blk_w - 1) / c->blk_w;
    p->nblocks_v = (c->h + c->blk_h - 1) / c->blk_h;
    p->temporal_mc = c->temporal_mc;
    p->lossless = c->lossless;
}

static DSV_FRAME *
frame_from_yuv(const refop_cfg *c, const uint8_t *yuv, int extend)
{
    DSV_FRAME *f = dsv_mk_frame(c->subsamp, c->w, c->h, 1);
    int i, y;
    for (i = 0; i < 3; i++) {
        DSV_PLANE *p = &f->planes[i];
        for (y = 0; y < p->h; y++) {
            memcpy(DSV_GET_LINE(p, y), yuv, p->w);
            yuv += p->w;
        }
    }
    if (extend) {
        dsv_extend_frame(f);
    }
    return f;
}

static void
frame_to_yuv(DSV_FRAME *f, uint8_t *yuv)
{
    int i, y;
    for (i = 0; i < 3; i++) {
        DSV_PLANE *p = &f->planes[i];
        for (y = 0; y < p->h; y++) {
            memcpy(yuv, DSV_GET_LINE(p, y), p->w);
            yuv += p->w;
        }
    }
}

static void
pyramid(DSV_FRAME *base, DSV_FRAME **pyr, int levels)
{
    DSV_FRAME *prev = base;
    int i;
    for (i = 0; i < levels; i++) {
        pyr[i] = dsv_mk_frame(base->format, DSV_ROUND_SHIFT(base->width, i + 1), DSV_ROUND_SHIFT(base->height, i + 1), 1);
        dsv_ds2x_frame_luma(pyr[i], prev);
        dsv_extend_frame_luma(pyr[i]);
        prev = pyr[i];
    }
}

int
refop_sizeof_mv(void)
{
    return (int) sizeof(DSV_MV);
}

/* out3 = { intra %, scene-change blocks %, average error } */
int
refop_hme(const refop_cfg *c, const uint8_t *src_yuv, const uint8_t *ref_yuv, const uint8_t *ogr_yuv,
          const DSV_MV *prev_mvs, int quant, DSV_MV *out_mvs, int *out3)
{
    DSV_PARAMS prm;
    DSV_ENCODER enc;
    DSV_HME hme;
    DSV_FRAME *src, *ref, *ogr, *ps[DSV_MAX_PYRAMID_LEVELS], *pr[DSV_MAX_PYRAMID_LEVELS], *po[DSV_MAX_PYRAMID_LEVELS];
    int i, nblk;

    mk_params(c, &prm);
    nblk = prm.nblocks_h * prm.nblocks_v;
    memset(&enc, 0, sizeof(enc));
    enc.pyramid_levels = c->pyramid_levels;
    enc.skip_block_thresh = c->skip_thresh;
    src = frame_from_yuv(c, src_yuv, 1);
    ref = frame_from_yuv(c, ref_yuv, 1);
    ogr = frame_from_yuv(c, ogr_yuv, 1);
    pyramid(src, ps, c->pyramid_levels);
    pyramid(ref, pr, c->pyramid_levels);
    pyramid(ogr, po, c->pyramid_levels);
    memset(&hme, 0, sizeof(hme));
    hme.enc = &enc;
    hme.params = &prm;
    hme.quant = quant;
    hme.src[0] = src;
    hme.ref[0] = ref;
    hme.ogr[0] = ogr;
    hme.ref_mvf = (DSV_MV *) prev_mvs;
    for (i = 0; i < c->pyramid_levels; i++) {
        hme.src[i + 1] = ps[i];
        hme.ref[i + 1] = pr[i];
        hme.ogr[i + 1] = po[i];
    }
    out3[1] = out3[2] = 0;
    out3[0] = dsv_hme(&hme, &out3[1], &out3[2]);
    memcpy(out_mvs, hme.mvf[0], (size_t) nblk * sizeof(DSV_MV));
    for (i = 0; i <= c->pyramid_levels; i++) {
        dsv_free(hme.mvf[i]);
    }
    for (i = 0; i < c->pyramid_levels; i++) {
        dsv_frame_ref_dec(ps[i]);
        dsv_frame_ref_dec(pr[i]);
        dsv_frame_ref_dec(po[i]);
    }
    dsv_frame_ref_dec(src);
    dsv_frame_ref_dec(ref);
    dsv_frame_ref_dec(ogr);
    return 0;
}

/* pyramid level `level` (1..n) of a frame, luma, WITH its 32-px border:
 * out must hold (h_l + 64) * stride bytes; returns the stride */
int
refop_pyramid_level(const refop_cfg *c, const uint8_t *yuv, int level, uint8_t *out, int *pw, int *ph)
{
    DSV_FRAME *f = frame_from_yuv(c, yuv, 1), *p[DSV_MAX_PYRAMID_LEVELS];
    DSV_PLANE *pl;
    int i, stride;
    pyramid(f, p, level);
    pl = &p[level - 1]->planes[0];
    stride = pl->stride;
    *pw = pl->w;
    *ph = pl->h;
    memcpy(out, pl->data - DSV_FRAME_BORDER * stride - DSV_FRAME_BORDER, (size_t) stride * (pl->h + 2 * DSV_FRAME_BORDER));
    for (i = 0; i < level; i++) {
        dsv_frame_ref_dec(p[i]);
    }
    dsv_frame_ref_dec(f);
    return stride;
}

int
refop_intra_analysis(const refop_cfg *c, const uint8_t *src_yuv, DSV_MV *out_mvs)
{
    DSV_PARAMS prm;
    DSV_FRAME *src;
    DSV_MV *mv;
    mk_params(c, &prm);
    src = frame_from_yuv(c, src_yuv, 1);
    mv = dsv_intra_analysis(src, &prm);
    memcpy(out_mvs, mv, (size_t) prm.nblocks_h * prm.nblocks_v * sizeof(DSV_MV));
    dsv_free(mv);
    dsv_frame_ref_dec(src);
    return 0;
}

static void
coef_dims(const refop_cfg *c, int plane, int *w, int *h)
{
    DSV_COEFS k[3];
    dsv_mk_coefs(k, c->subsamp, c->w, c->h);
    *w = k[plane].width;
    *h = k[plane].height;
    dsv_free(k[0].data);
}

int
refop_coef_dims(const refop_cfg *c, int plane, int *w, int *h)
{
    coef_dims(c, plane, w, h);
    return 0;
}

/* forward transform of one plane of `yuv`; coefs_out = width*height int32 */
int
refop_fwd_sbt(const refop_cfg *c, int plane, const uint8_t *yuv, const uint8_t *blockdata, int32_t *coefs_out)
{
    DSV_PARAMS prm;
    DSV_FMETA fm;
    DSV_COEFS k[3];
    DSV_FRAME *f;
    mk_params(c, &prm);
    f = frame_from_yuv(c, yuv, 1);
    dsv_mk_coefs(k, c->subsamp, c->w, c->h);
    memset(&fm, 0, sizeof(fm));
    fm.params = &prm;
    fm.blockdata = (uint8_t *) blockdata;
    fm.cur_plane = (uint8_t) plane;
    fm.isP = (uint8_t) c->isP;
    fm.fnum = c->fnum;
    dsv_fwd_sbt(&f->planes[plane], &k[plane], &fm);
    memcpy(coefs_out, k[plane].data, (size_t) k[plane].width * k[plane].height * sizeof(int32_t));
    dsv_free(k[0].data);
    dsv_frame_ref_dec(f);
    return 0;
}

/* quantise + entropy-code one plane.  coefs: in = transform output, out = the
 * de-quantised values the reference leaves in place.  bits_out/len: the plane's
 * bytes exactly as dsv_encode_plane appends them */
int
refop_encode_plane(const refop_cfg *c, int plane, int q, int32_t *coefs, const uint8_t *blockdata, const DSV_MV *mvs,
                   uint8_t *bits_out, int *len)
{
    DSV_PARAMS prm;
    DSV_FMETA fm;
    DSV_COEFS k;
    DSV_BS bs;
    int w, h;
    mk_params(c, &prm);
    coef_dims(c, plane, &w, &h);
    k.data = coefs;
    k.width = w;
    k.height = h;
    memset(&fm, 0, sizeof(fm));
    fm.params = &prm;
    fm.blockdata = (uint8_t *) blockdata;
    fm.mvs = (DSV_MV *) mvs;
    fm.cur_plane = (uint8_t) plane;
    fm.isP = (uint8_t) c->isP;
    fm.fnum = c->fnum;
    dsv_bs_init(&bs, bits_out);
    dsv_encode_plane(&bs, &k, q, &fm);
    dsv_bs_align(&bs);
    *len = (int) dsv_bs_ptr(&bs);
    return 0;
}

int
refop_inv_sbt(const refop_cfg *c, int plane, int q, const int32_t *coefs, const uint8_t *blockdata, uint8_t *plane_out)
{
    DSV_PARAMS prm;
    DSV_FMETA fm;
    DSV_COEFS k[3];
    DSV_FRAME *f;
    DSV_PLANE *p;
    int y;
    mk_params(c, &prm);
    f = dsv_mk_frame(c->subsamp, c->w, c->h, 1);
    dsv_mk_coefs(k, c->subsamp, c->w, c->h);
    memcpy(k[plane].data, coefs, (size_t) k[plane].width * k[plane].height * sizeof(int32_t));
    memset(&fm, 0, sizeof(fm));
    fm.params = &prm;
    fm.blockdata = (uint8_t *) blockdata;
    fm.cur_plane = (uint8_t) plane;
    fm.isP = (uint8_t) c->isP;
    fm.fnum = c->fnum;
    dsv_inv_sbt(&f->planes[plane], &k[plane], q, &fm);
    p = &f->planes[plane];
    for (y = 0; y < p->h; y++) {
        memcpy(plane_out + (size_t) y * p->w, DSV_GET_LINE(p, y), p->w);
    }
    dsv_free(k[0].data);
    dsv_frame_ref_dec(f);
    return 0;
}

/* prediction + residual of the source against the (extended) reference */
int
refop_sub_pred(const refop_cfg *c, const DSV_MV *mvs, const uint8_t *src_yuv, const uint8_t *ref_yuv, uint8_t *pred_yuv,
               uint8_t *resd_yuv)
{
    DSV_PARAMS prm;
    DSV_FRAME *resd, *ref, *pred;
    mk_params(c, &prm);
    resd = frame_from_yuv(c, src_yuv, 1);
    ref = frame_from_yuv(c, ref_yuv, 1);
    pred = dsv_mk_frame(c->subsamp, c->w, c->h, 1);
    dsv_sub_pred((DSV_MV *) mvs, &prm, pred, resd, ref);
    frame_to_yuv(pred, pred_yuv);
    frame_to_yuv(resd, resd_yuv);
    dsv_frame_ref_dec(resd);
    dsv_frame_ref_dec(ref);
    dsv_frame_ref_dec(pred);
    return 0;
}

/* reconstruction + loop filters; resd_yuv is updated in place */
int
refop_add_res(const refop_cfg *c, const DSV_MV *mvs, const uint8_t *blockdata, int q, uint8_t *resd_yuv,
              const uint8_t *pred_yuv, int do_filter)
{
    DSV_PARAMS prm;
    DSV_FMETA fm;
    DSV_FRAME *resd, *pred;
    mk_params(c, &prm);
    resd = frame_from_yuv(c, resd_yuv, 0);
    pred = frame_from_yuv(c, pred_yuv, 0);
    memset(&fm, 0, sizeof(fm));
    fm.params = &prm;
    fm.blockdata = (uint8_t *) blockdata;
    fm.mvs = (DSV_MV *) mvs;
    fm.isP = 1;
    fm.fnum = c->fnum;
    dsv_add_res((DSV_MV *) mvs, &fm, q, resd, pred, do_filter);
    frame_to_yuv(resd, resd_yuv);
    dsv_frame_ref_dec(resd);
    dsv_frame_ref_dec(pred);
    return 0;
}

int
refop_intra_filter(const refop_cfg *c, int q, const uint8_t *blockdata, uint8_t *yuv, int do_filter)
{
    DSV_PARAMS prm;
    DSV_FMETA fm;
    DSV_FRAME *f;
    int i;
    mk_params(c, &prm);
    f = frame_from_yuv(c, yuv, 0);
    memset(&fm, 0, sizeof(fm));
    fm.params = &prm;
    fm.blockdata = (uint8_t *) blockdata;
    fm.isP = 0;
    fm.fnum = c->fnum;
    for (i = 0; i < 3; i++) {
        fm.cur_plane = (uint8_t) i;
        dsv_intra_filter(q, &prm, &fm, i, &f->planes[i], do_filter);
    }
    frame_to_yuv(f, yuv);
    dsv_frame_ref_dec(f);
    return 0;
}

/* the decoder's half of a coefficient plane: `bits`/`len` as refop_encode_plane
 * produced them -> de-quantised coefficients (dsv_decode_plane, hzcc.c:616-649,
 * into a zeroed plane like dsv_decoder.c:506-519).  Returns dsv_decode_plane's
 * success flag. */
int
refop_decode_plane(const refop_cfg *c, int plane, int q, const uint8_t *bits, int len, const uint8_t *blockdata,
                   const DSV_MV *mvs, int32_t *coefs_out)
{
    DSV_PARAMS prm;
    DSV_FMETA fm;
    DSV_COEFS k[3];
    DSV_BS bs;
    uint8_t *copy;
    int ok;
    mk_params(c, &prm);
    dsv_mk_coefs(k, c->subsamp, c->w, c->h);
    memset(&fm, 0, sizeof(fm));
    fm.params = &prm;
    fm.blockdata = (uint8_t *) blockdata;
    fm.mvs = (DSV_MV *) mvs;
    fm.cur_plane = (uint8_t) plane;
    fm.isP = (uint8_t) c->isP;
    fm.fnum = c->fnum;
    copy = calloc((size_t) len + 64, 1);
    memcpy(copy, bits, (size_t) len);
    dsv_bs_init(&bs, copy);
    ok = dsv_decode_plane(&bs, &k[plane], q, &fm);
    memcpy(coefs_out, k[plane].data, (size_t) k[plane].width * k[plane].height * sizeof(int32_t));
    free(copy);
    dsv_free(k[0].data);
    return ok;
}

/* one plane of dsv_extend_frame's result WITH its 32-px border (frame.c:357-434):
 * out must hold (h + 64) * stride bytes; returns the stride */
int
refop_extend_plane(const refop_cfg *c, const uint8_t *yuv, int plane, uint8_t *out, int *pw, int *ph)
{
    DSV_FRAME *f = frame_from_yuv(c, yuv, 1);
    DSV_PLANE *pl = &f->planes[plane];
    int stride = pl->stride;
    *pw = pl->w;
    *ph = pl->h;
    memcpy(out, pl->data - DSV_FRAME_BORDER * stride - DSV_FRAME_BORDER, (size_t) stride * (pl->h + 2 * DSV_FRAME_BORDER));
    dsv_frame_ref_dec(f);
    return stride;
}
