/*
 * dsv_dec.c -- decoder control: packet / picture header parsing, block side
 * information (stability, intra meta, motion) and the per-picture GPU
 * schedule.
 *
 * Bitstream syntax follows the reference decoder (src/dsv_decoder.c:21-238,
 * :393-590; spec B.1-B.2.3).  What differs is where pixels live: coefficient
 * planes, the residual, the output picture and the reference picture are
 * device-resident (dsv_cuda.h).  Per picture the host parses bits into small
 * arrays (blockdata, motion vectors, ordered coefficient symbols), uploads
 * them, queues dequant -> inverse SBT -> (intra filter | predict + reconstruct
 * + loop filters) -> border extension on one stream, and copies the finished
 * picture back once.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#include "dsv_host.h"
#include "dsv_bits_inl.h"
#include "dsv_mvutil_inl.h"
#include "../../include/dsv_decoder.h"

/* optional thread-CPU phase accounting (DSV_PROFILE=1): where one decoder instance's host
 * thread spends its time, summed over pictures and printed by dsv_dec_free */
static int g_prof = -1;
enum { DP_SIDE, DP_PLANES, DP_QUEUE, DP_TAIL, DP_PREPARSE, DP_N };
static const char *dp_name[DP_N] = { "header + side information", "coefficient planes (parse)",
                                     "queue device work", "download + wait",
                                     "batch entropy decode on the device (gather + wait)" };
static double
cpu_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static double
wall_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
#define DPROF_START(s)                \
    do {                              \
        if (g_prof > 0) {             \
            (s)->ph_t0 = cpu_ms();    \
            (s)->ph_w0 = wall_ms();   \
        }                             \
    } while (0)
#define DPROF(s, ph)                              \
    do {                                          \
        if (g_prof > 0) {                         \
            double t_ = cpu_ms(), w_ = wall_ms(); \
            (s)->ph_ms[ph] += t_ - (s)->ph_t0;    \
            (s)->ph_wall[ph] += w_ - (s)->ph_w0;  \
            (s)->ph_t0 = t_;                      \
            (s)->ph_w0 = w_;                      \
        }                                         \
    } while (0)

typedef struct {
    DSV_IMAGE img; /* first member: DSV_DECODER.ref points at this object */
    dsvcu_ctx *ctx;
    dsvcu_coefs *coefs;
    dsvcu_frame *resd;
    dsvcu_frame *pic[2]; /* output / reference, swapped after every reference picture */
    int cur;
    int have_ref;
    int w, h, subsamp;
    uint8_t *blockdata;
    DSV_MV *mvs;
    int nblk_cap;
    double ph_ms[DP_N], ph_t0, ph_wall[DP_N], ph_w0;
    int ph_frames;
    int ph_where[3]; /* pictures parsed by the device (part 0, part 1) and by the host */
    int ph_side_dev; /* pictures whose side information was decoded on the device too */
    int ph_stolen;   /* pictures sent to the device parser that the host parsed because it got there first */
    /* batches of pictures whose coefficient planes are entropy-decoded on the device ahead of
     * time (dsv_dec_preparse), one per parse set of the context: first span of picture i in
     * the batch, or -1 where the host parses */
    struct {
        int *first;
        int *side;   /* index of the picture's side information in the batch, or -1 (host) */
        int nsides, nsides_early;
        int n, cap;
        int nspans;  /* planes handed to the device */
        int nearly;  /* of which part 0 (short planes, needed first); the rest is part 1 */
        int cset;    /* the context's parse set they went to */
        int pending[2]; /* part launched, result not looked at yet */
    } pre[2];
    int pre_last;
} DEC_STATE;

int dsv_get_thread_device(void);

/* set by the whole-stream drivers (dsv_pipe.c): the next decoded picture of
 * this thread is copied device -> dst directly (tightly packed planes) instead
 * of into a freshly allocated host frame */
static __thread uint8_t *tls_direct_out = NULL;

void
dsv_dec_direct_output(uint8_t *dst)
{
    tls_direct_out = dst;
}

/* set by the whole-stream drivers: pictures decoded by this thread are NOT waited for;
 * dsv_dec returns as soon as the picture's work is queued (its frame, written straight to
 * the caller's memory, is complete after dsv_dec_flush).  Host parsing of picture n + 1
 * then overlaps the device work of picture n. */
static __thread int tls_async = 0;

void
dsv_dec_set_async(int on)
{
    tls_async = on;
}

static int preparse_collect(DEC_STATE *s, int set, int part);

int
dsv_dec_flush(DSV_DECODER *d)
{
    DEC_STATE *s = (DEC_STATE *) d->ref;
    int set;
    /* a batch of planes that was sent ahead and is not wanted any more */
    for (set = 0; s && s->ctx && set < 2; set++) {
        (void) preparse_collect(s, set, 0);
        (void) preparse_collect(s, set, 1);
        s->pre[set].n = 0;
    }
    if (s && s->ctx && dsvcu_sync(s->ctx)) {
        DSV_ERROR(("dsv_dec_flush: %s", dsvcu_last_error()));
        return -1;
    }
    return 0;
}

/* set by the whole-stream drivers before dsv_dec: the packet is picture `idx` of the batch
 * that dsv_dec_preparse put into `set` (idx -1: parse the planes on the host) */
static __thread int tls_parsed = -1, tls_parsed_set = 0;

void
dsv_dec_use_parsed(int set, int idx)
{
    tls_parsed_set = set & 1;
    tls_parsed = idx;
}

static void
state_free(DEC_STATE *s)
{
    if (!s) {
        return;
    }
    if (g_prof > 0 && s->ph_frames) {
        int i;
        char line[1024];
        int at = snprintf(line, sizeof(line), "[dsv_dec profile] %d pictures, host thread CPU (wall) ms per picture:",
                          s->ph_frames);
        for (i = 0; i < DP_N && at < (int) sizeof(line) - 96; i++) {
            at += snprintf(line + at, sizeof(line) - (size_t) at, " %s %.3f (%.3f);", dp_name[i],
                           s->ph_ms[i] / s->ph_frames, s->ph_wall[i] / s->ph_frames);
        }
        fprintf(stderr, "%s planes parsed on the device for %d + %d pictures (part 0 + part 1; side information of %d), "
                "on the host for %d (%d of them because the device parser was not through yet)\n", line, s->ph_where[0],
                s->ph_where[1], s->ph_side_dev, s->ph_where[2], s->ph_stolen);
    }
    if (s->ctx) {
        dsvcu_sync(s->ctx);
        if (s->coefs) dsvcu_coefs_destroy(s->ctx, s->coefs);
        if (s->resd) dsvcu_frame_destroy(s->ctx, s->resd);
        if (s->pic[0]) dsvcu_frame_destroy(s->ctx, s->pic[0]);
        if (s->pic[1]) dsvcu_frame_destroy(s->ctx, s->pic[1]);
        dsvcu_ctx_destroy(s->ctx);
    }
    free(s->blockdata);
    free(s->mvs);
    free(s->pre[0].first);
    free(s->pre[1].first);
    free(s->pre[0].side);
    free(s->pre[1].side);
    free(s);
}

static DEC_STATE *
state_get(DSV_DECODER *d)
{
    DEC_STATE *s = (DEC_STATE *) d->ref;
    DSV_META *m = &d->vidmeta;
    if (s && (s->w != m->width || s->h != m->height || s->subsamp != m->subsamp)) {
        state_free(s);
        s = NULL;
        d->ref = NULL;
    }
    if (s) {
        return s;
    }
    s = calloc(1, sizeof(*s));
    if (!s) {
        return NULL;
    }
    s->img.refcount = 1;
    s->w = m->width;
    s->h = m->height;
    s->subsamp = m->subsamp;
    if (dsvcu_ctx_create(&s->ctx, dsv_get_thread_device(), m->width, m->height, m->subsamp) ||
        dsvcu_coefs_create(s->ctx, &s->coefs) || dsvcu_frame_create(s->ctx, &s->resd) ||
        dsvcu_frame_create(s->ctx, &s->pic[0]) || dsvcu_frame_create(s->ctx, &s->pic[1])) {
        DSV_ERROR(("GPU decoder state: %s", dsvcu_last_error()));
        state_free(s);
        return NULL;
    }
    d->ref = &s->img;
    return s;
}

/* B.1 packet header; returns the packet type or -1 */
static int
read_packet_hdr(DSV_BITRD *br)
{
    int c0 = (int) dsv_br_bits(br, 8), c1 = (int) dsv_br_bits(br, 8);
    int c2 = (int) dsv_br_bits(br, 8), c3 = (int) dsv_br_bits(br, 8);
    int type;
    if (c0 != DSV_FOURCC_0 || c1 != DSV_FOURCC_1 || c2 != DSV_FOURCC_2 || c3 != DSV_FOURCC_3) {
        DSV_ERROR(("bad 4cc (%c %c %c %c)\n", c0, c1, c2, c3));
        return -1;
    }
    (void) dsv_br_bits(br, 8); /* minor version */
    type = (int) dsv_br_bits(br, 8);
    (void) dsv_br_bits(br, 32); /* prev link */
    (void) dsv_br_bits(br, 32); /* next link */
    return type;
}

/* B.2.1 metadata packet.  Returns 0 when the fields describe a picture this
 * build can hold (the reference accepts anything and fails later in its
 * allocator; a device allocation sized by a corrupt field must not happen) */
#define DSV_DEC_MAX_DIM 16384
static int
read_meta(DSV_DECODER *d, DSV_BITRD *br)
{
    DSV_META m;
    memset(&m, 0, sizeof(m));
    m.width = (int) dsv_br_ueg(br);
    m.height = (int) dsv_br_ueg(br);
    m.subsamp = (int) dsv_br_ueg(br);
    m.fps_num = (int) dsv_br_ueg(br);
    m.fps_den = (int) dsv_br_ueg(br);
    m.aspect_num = (int) dsv_br_ueg(br);
    m.aspect_den = (int) dsv_br_ueg(br);
    m.inter_sharpen = (int) dsv_br_ueg(br);
    m.reserved = dsv_br_bit(br) ? (int) dsv_br_bits(br, 15) : 0;
    if (m.width <= 0 || m.height <= 0 || m.width > DSV_DEC_MAX_DIM || m.height > DSV_DEC_MAX_DIM) {
        DSV_ERROR(("metadata: unusable picture size %d x %d", m.width, m.height));
        return -1;
    }
    switch (m.subsamp) {
        case DSV_SUBSAMP_444:
        case DSV_SUBSAMP_422:
        case DSV_SUBSAMP_UYVY:
        case DSV_SUBSAMP_420:
        case DSV_SUBSAMP_411:
        case DSV_SUBSAMP_410:
            break;
        default:
            DSV_ERROR(("metadata: unknown subsampling code %d", m.subsamp));
            return -1;
    }
    d->vidmeta = m;
    return 0;
}

/* a length-prefixed, byte-aligned sub-stream inside the picture payload.  The
 * coded length is not trusted: a sub-stream that would start or end outside the
 * packet yields an empty one and puts the reader at the end of the packet (every
 * later read then sees the zero padding); returns -1 in that case */
static int
open_substream(DSV_BITRD *in, const uint8_t **start, size_t *len)
{
    size_t n = dsv_br_ueg(in), at;
    dsv_br_align(in);
    at = dsv_br_byte(in);
    if (at > in->len || n > in->len - at) {
        *start = in->buf + in->len;
        *len = 0;
        in->pos = in->len * 8;
        return -1;
    }
    *start = in->buf + at;
    *len = n;
    in->pos += n * 8;
    return 0;
}

/* B.2.3.1 stability (I) / skip (P) bits */
/* The side information of a picture is read with the inline readers of dsv_bits_inl.h (same
 * codes, same behaviour at and behind the end of a sub-stream as dsv_bits.c): 8160 blocks per
 * 1080p picture, on a host thread whose job is to keep a GPU fed. */
static void
frle_end(const DSV_FRLE *e)
{
    if (e->nz > 1) {
        DSV_ERROR(("%d remaining in run", e->nz));
    }
}

static int
read_stability(DSV_BITRD *in, uint8_t *blockdata, int nblk, int isP, const int *stats)
{
    DSV_FRLE rle;
    const uint8_t *p;
    size_t len;
    int i, shift = isP ? DSV_SKIP_BIT : DSV_STABLE_BIT;
    const int flip = stats[DSV_STABLE_STAT] == DSV_ZERO_MARKER;

    dsv_br_align(in);
    if (open_substream(in, &p, &len)) {
        return -1;
    }
    dsv_frle_init(&rle, p, len + 8);
    for (i = 0; i < nblk; i++) {
        const int bit = dsv_frle_get(&rle) ^ flip;
        blockdata[i] = (uint8_t) (bit << shift);
    }
    frle_end(&rle);
    return 0;
}

/* B.2.3.2 ringing + B.2.3.3 maintain bits of an intra picture */
static int
read_intra_meta(DSV_BITRD *in, uint8_t *blockdata, int nblk, const int *stats)
{
    DSV_FRLE rr, rm;
    const uint8_t *p;
    size_t len;
    int i;
    const int flip_r = stats[DSV_RINGING_STAT] == DSV_ZERO_MARKER, flip_m = stats[DSV_MAINTAIN_STAT] == DSV_ZERO_MARKER;

    dsv_br_align(in);
    if (open_substream(in, &p, &len)) {
        return -1;
    }
    dsv_frle_init(&rr, p, len + 8);
    dsv_br_align(in);
    if (open_substream(in, &p, &len)) {
        return -1;
    }
    dsv_frle_init(&rm, p, len + 8);
    for (i = 0; i < nblk; i++) {
        const int br = dsv_frle_get(&rr) ^ flip_r, bm = dsv_frle_get(&rm) ^ flip_m;
        blockdata[i] |= (uint8_t) ((bm << DSV_MAINTAIN_BIT) | (br << DSV_RINGING_BIT));
    }
    frle_end(&rr);
    frle_end(&rm);
    return 0;
}

/* B.2.3.4 motion data: five sub-streams, vectors coded against the
 * left/top/top-left predictor */
static int
read_motion(DSV_BITRD *in, DSV_PARAMS *prm, uint8_t *blockdata, DSV_MV *mvs, const int *stats)
{
    DSV_FR sub[DSV_SUB_NSUB];
    DSV_FRLE mode_rle, eprm_rle;
    int i, j;
    const int flip_mode = stats[DSV_MODE_STAT] == DSV_ZERO_MARKER, flip_eprm = stats[DSV_EPRM_STAT] == DSV_ZERO_MARKER;

    dsv_br_align(in);
    for (i = 0; i < DSV_SUB_NSUB; i++) {
        const uint8_t *p;
        size_t len;
        if (open_substream(in, &p, &len)) {
            return -1;
        }
        if (i == DSV_SUB_MODE) {
            dsv_frle_init(&mode_rle, p, len + 8);
        } else if (i == DSV_SUB_EPRM) {
            dsv_frle_init(&eprm_rle, p, len + 8);
        } else {
            sub[i].buf = p;
            sub[i].len = len + 8;
            sub[i].pos = 0;
        }
    }
    for (j = 0; j < prm->nblocks_v; j++) {
        for (i = 0; i < prm->nblocks_h; i++) {
            int idx = i + j * prm->nblocks_h;
            DSV_MV *mv = &mvs[idx];
            int mode, eprm, px, py;

            if (blockdata[idx] & DSV_IS_SKIP) {
                DSV_MV_SET_SKIP(mv, 1);
                mv->u.all = 0;
                blockdata[idx] |= DSV_IS_STABLE;
                continue;
            }
            DSV_MV_SET_SKIP(mv, 0);
            mode = dsv_frle_get(&mode_rle) ^ flip_mode;
            eprm = dsv_frle_get(&eprm_rle) ^ flip_eprm;
            DSV_MV_SET_INTRA(mv, mode);
            DSV_MV_SET_EPRM(mv, eprm);
            blockdata[idx] &= (uint8_t) ~DSV_IS_STABLE;
            blockdata[idx] |= (uint8_t) (eprm << DSV_EPRM_BIT);

            dsv_movec_pred_inl(mvs, prm, i, j, &px, &py);
            if (mode) {
                px = DSV_SAR_R(px, 2);
                py = DSV_SAR_R(py, 2);
            }
            mv->u.mv.x = (int16_t) (dsv_fr_seg(&sub[DSV_SUB_MV_X]) + px);
            mv->u.mv.y = (int16_t) (dsv_fr_seg(&sub[DSV_SUB_MV_Y]) + py);
            if (mode) {
                DSV_FR *sb = &sub[DSV_SUB_SBIM];
                mv->u.mv.x *= 4; /* intra vectors are full-pel */
                mv->u.mv.y *= 4;
                mv->submask = dsv_fr_bit(sb) ? DSV_MASK_ALL_INTRA : (uint8_t) dsv_fr_bits(sb, 4);
                mv->dc = dsv_fr_bit(sb) ? (uint16_t) (dsv_fr_bits(sb, 8) | DSV_SRC_DC_PRED) : 0;
                blockdata[idx] |= DSV_IS_INTRA;
            }
            if (dsv_neighbordif_inl(mvs, prm, i, j) > DSV_NDIF_THRESH) {
                blockdata[idx] |= DSV_IS_STABLE;
            }
        }
    }
    frle_end(&mode_rle);
    frle_end(&eprm_rle);
    return 0;
}

void
dsv_dec_free(DSV_DECODER *d)
{
    if (d->ref) {
        state_free((DEC_STATE *) d->ref);
        d->ref = NULL;
    }
}

DSV_META *
dsv_get_metadata(DSV_DECODER *d)
{
    DSV_META *m = dsv_alloc(sizeof(DSV_META));
    memcpy(m, &d->vidmeta, sizeof(DSV_META));
    return m;
}

#define GPU(call)                                              \
    do {                                                       \
        if (call) {                                            \
            DSV_ERROR(("%s: %s", #call, dsvcu_last_error())); \
            return DSV_DEC_ERROR;                              \
        }                                                      \
    } while (0)

/* ---- debug overlay on the OUTPUT copy of a picture (the reference picture stays clean),
 * reference dsv_decoder.c:243-350: block grid, markers for stable / skipped and
 * "maintain" blocks, motion vectors as lines, intra sub-block dots.  Pure host drawing. */
#define OVL_SHADE 255

static void
ovl_put(DSV_PLANE *lp, int x, int y, int v)
{
    if (x >= 0 && y >= 0 && x < lp->w && y < lp->h) {
        lp->data[(size_t) y * lp->stride + x] = (uint8_t) v;
    }
}

/* Bresenham line from the block centre along the (quarter-pel valued) vector */
static void
ovl_vector(DSV_PLANE *lp, int x0, int y0, int vx, int vy, int bw, int bh)
{
    int x1, y1, dx, dy, sx, sy, err;
    x0 += bw / 2;
    y0 += bh / 2;
    x1 = x0 + vx;
    y1 = y0 + vy;
    dx = abs(x1 - x0);
    dy = abs(y1 - y0);
    sx = x0 < x1 ? 1 : -1;
    sy = y0 < y1 ? 1 : -1;
    err = dx - dy;
    ovl_put(lp, x0, y0, OVL_SHADE);
    while (x0 != x1 || y0 != y1) {
        int e2;
        ovl_put(lp, x0, y0, OVL_SHADE);
        e2 = 2 * err;
        if (e2 > -dy) {
            err -= dy;
            x0 += sx;
        }
        if (e2 < dx) {
            err += dx;
            y0 += sy;
        }
    }
}

static void
overlay_info(DSV_FRAME *f, const DSV_PARAMS *p, const uint8_t *blockdata, const DSV_MV *mvs, int mode)
{
    DSV_PLANE *lp = &f->planes[0];
    const int bw = p->blk_w, bh = p->blk_h;
    int i, j, k;
    for (j = 0; j < p->nblocks_v; j++) {
        const int y = j * bh;
        /* the reference fills the whole line including the stride padding; only the
         * visible part exists in a tightly packed output frame */
        memset(lp->data + (size_t) y * lp->stride, OVL_SHADE, (size_t) MIN(lp->stride, lp->w));
        for (i = 0; i < p->nblocks_h; i++) {
            const int x = i * bw, idx = i + j * p->nblocks_h;
            const DSV_MV *mv = mvs ? &mvs[idx] : NULL;
            for (k = y; k < y + bh && k < lp->h; k++) {
                ovl_put(lp, x, k, OVL_SHADE);
            }
            if (mode & DSV_DRAW_STABHQ) {
                const int a = x + bw / 2, b = y + bh / 2;
                if (blockdata[idx] & (DSV_IS_SKIP | DSV_IS_STABLE)) {
                    for (k = -bw / 4; k <= bw / 4; k++) {
                        ovl_put(lp, a + k, b, (k & 1) * 255);
                    }
                }
                if (blockdata[idx] & DSV_IS_MAINTAIN) {
                    for (k = -bh / 4; k <= bh / 4; k++) {
                        ovl_put(lp, a, b + k, (k & 1) * 255);
                    }
                }
            }
            if (mv && (mode & DSV_DRAW_MOVECS) && !(blockdata[idx] & DSV_IS_SKIP)) {
                ovl_vector(lp, x, y, mv->u.mv.x, mv->u.mv.y, bw, bh);
            }
            if (mv && (mode & DSV_DRAW_IBLOCK)) {
                static const int qx[4] = { 1, 3, 1, 3 }, qy[4] = { 1, 1, 3, 3 };
                static const int bit[4] = { DSV_MASK_INTRA00, DSV_MASK_INTRA01, DSV_MASK_INTRA10, DSV_MASK_INTRA11 };
                for (k = 0; k < 4; k++) {
                    if (mv->submask & bit[k]) {
                        /* unlike the other marks the reference writes these unchecked; blocks that
                         * overhang the picture are clipped here */
                        ovl_put(lp, x + bw * qx[k] / 4, y + bh * qy[k] / 4, OVL_SHADE);
                    }
                }
            }
        }
    }
}

static int
decode_picture(DSV_DECODER *d, DEC_STATE *s, DSV_BITRD *br, int pkt_type, DSV_FRAME **out, DSV_FNUM *fn)
{
    DSV_META *meta = &d->vidmeta;
    DSV_PARAMS *p = &s->img.params;
    dsvcu_fmeta fm;
    dsvcu_frame *dst, *ref;
    DSV_FRAME *host;
    DSV_FNUM fno;
    int stats[DSV_MAX_STAT];
    int i, nblk, quant, is_ref, do_filter, isP, good_planes = 0, on_device = 0, planes_dev, side_dev, stolen = 0;

    if (g_prof < 0) {
        g_prof = getenv("DSV_PROFILE") ? atoi(getenv("DSV_PROFILE")) : 0;
    }
    DPROF_START(s);
    if (g_prof > 0) {
        s->ph_frames++;
    }
    memset(p, 0, sizeof(*p));
    p->vidmeta = meta;
    p->has_ref = DSV_PT_HAS_REF(pkt_type);
    isP = p->has_ref;
    is_ref = DSV_PT_IS_REF(pkt_type);

    dsv_br_align(br);
    fno = dsv_br_bits(br, 32);
    dsv_br_align(br);
    {
        /* exponents above 1 are outside the format (and an unchecked shift is undefined) */
        unsigned ex = dsv_br_ueg(br), ey = dsv_br_ueg(br);
        if (ex > 2 || ey > 2) {
            DSV_ERROR(("bad block size exponents %u, %u", ex, ey));
            return DSV_DEC_ERROR;
        }
        p->blk_w = 16 << ex;
        p->blk_h = 16 << ey;
    }
    if (p->blk_w < DSV_MIN_BLOCK_SIZE || p->blk_h < DSV_MIN_BLOCK_SIZE || p->blk_w > DSV_MAX_BLOCK_SIZE ||
        p->blk_h > DSV_MAX_BLOCK_SIZE) {
        return DSV_DEC_ERROR;
    }
    p->nblocks_h = DSV_UDIV_ROUND_UP(meta->width, p->blk_w);
    p->nblocks_v = DSV_UDIV_ROUND_UP(meta->height, p->blk_h);
    nblk = p->nblocks_h * p->nblocks_v;

    dsv_br_align(br);
    for (i = 0; i < DSV_MAX_STAT; i++) {
        stats[i] = DSV_ONE_MARKER;
    }
    stats[DSV_STABLE_STAT] = (int) dsv_br_bit(br);
    if (!isP) {
        stats[DSV_MAINTAIN_STAT] = (int) dsv_br_bit(br);
        stats[DSV_RINGING_STAT] = (int) dsv_br_bit(br);
    } else {
        stats[DSV_MODE_STAT] = (int) dsv_br_bit(br);
        stats[DSV_EPRM_STAT] = (int) dsv_br_bit(br);
    }
    do_filter = (int) dsv_br_bit(br);
    quant = (int) dsv_br_bits(br, DSV_MAX_QP_BITS);
    p->lossless = (quant == 1);
    p->reserved = dsv_br_bit(br) ? (int) dsv_br_bits(br, 15) : 0;
    dsv_br_align(br);

    /* planes (and, for an inter picture, side information) that went ahead to the device parser */
    if (tls_parsed >= 0 && tls_parsed < s->pre[tls_parsed_set].n && s->pre[tls_parsed_set].first[tls_parsed] >= 0) {
        const int part = s->pre[tls_parsed_set].first[tls_parsed] >= s->pre[tls_parsed_set].nearly;
        DPROF(s, DP_SIDE);
        /* Never wait for the device parser: if this picture's part of the batch is not through
         * yet, the picture is parsed here and now (the device's copy of it goes unused) -- the
         * host thread has nothing better to do, and the device, which runs behind the host,
         * gets its next picture sooner.  A host that is short of cores runs behind the parser
         * and finds its pictures ready. */
        if (s->pre[tls_parsed_set].pending[part] && dsvcu_parse_ready(s->ctx, s->pre[tls_parsed_set].cset, part) == 0) {
            stolen = 1;
            s->ph_stolen++;
        } else if (preparse_collect(s, tls_parsed_set, part)) {
            return DSV_DEC_ERROR;
        }
        DPROF(s, DP_PREPARSE);
    }
    planes_dev = !stolen && tls_parsed >= 0 && tls_parsed < s->pre[tls_parsed_set].n &&
                 s->pre[tls_parsed_set].first[tls_parsed] >= 0;
    side_dev = planes_dev && isP && !d->draw_info && s->pre[tls_parsed_set].side[tls_parsed] >= 0;

    if (!side_dev) {
        if (nblk > s->nblk_cap) {
            free(s->blockdata);
            free(s->mvs);
            s->blockdata = malloc((size_t) nblk);
            s->mvs = malloc((size_t) nblk * sizeof(DSV_MV));
            s->nblk_cap = nblk;
            if (!s->blockdata || !s->mvs) {
                free(s->blockdata);
                free(s->mvs);
                s->blockdata = NULL;
                s->mvs = NULL;
                s->nblk_cap = 0;
                return DSV_DEC_ERROR;
            }
        }
        memset(s->blockdata, 0, (size_t) nblk);
        memset(s->mvs, 0, (size_t) nblk * sizeof(DSV_MV));
        if (read_stability(br, s->blockdata, nblk, isP, stats) ||
            (isP ? read_motion(br, p, s->blockdata, s->mvs, stats) : read_intra_meta(br, s->blockdata, nblk, stats))) {
            DSV_ERROR(("side information runs past the end of the packet"));
            return DSV_DEC_ERROR;
        }
        dsv_br_align(br);
    }

    p->temporal_mc = isP ? (int) DSV_TEMPORAL_MC(fno) : 0;
    dsv_fmeta_from_params(&fm, p, isP, fno);
    if (tls_async && tls_direct_out && !(side_dev && planes_dev)) {
        /* the previous picture may still be reading the staging set just used */
        GPU(dsvcu_staging_flip(s->ctx));
    }
    DPROF(s, DP_SIDE);
    if (side_dev) {
        /* vector field and block flags were decoded on the device, with the planes */
        GPU(dsvcu_set_side_parsed(s->ctx, s->pre[tls_parsed_set].cset, s->pre[tls_parsed_set].side[tls_parsed], nblk));
    } else {
        GPU(dsvcu_set_side(s->ctx, s->blockdata, isP ? s->mvs : NULL, nblk));
    }
    DPROF(s, DP_QUEUE);

    /* intra pictures are reconstructed straight into the output picture;
     * inter pictures into the residual frame, then predicted + added */
    dst = s->pic[s->cur];
    ref = s->pic[s->cur ^ 1];
    if (planes_dev) {
        /* the symbols of all three planes are already on the device */
        GPU(dsvcu_dequant_parsed(s->ctx, s->coefs, quant, &fm, s->pre[tls_parsed_set].cset,
                                 s->pre[tls_parsed_set].first[tls_parsed]));
        DPROF(s, DP_QUEUE);
        good_planes = 7;
        on_device = 1;
        s->ph_where[s->pre[tls_parsed_set].first[tls_parsed] >= s->pre[tls_parsed_set].nearly]++;
        s->ph_side_dev += side_dev;
    } else {
        s->ph_where[2]++;
    }
    for (i = 0; i < 3 && !on_device; i++) {
        int cap, nsym, lstart[5], dc, cw, ch;
        dsvcu_symbol *st = dsvcu_symbol_staging(s->ctx, i, &cap);
        dsvcu_coefs_plane_dims(s->coefs, i, &cw, &ch);
        nsym = dsv_hzcc_read_plane(br, st, cap - 1, cw, ch, lstart, &dc);
        DPROF(s, DP_PLANES);
        if (nsym < 0) {
            DSV_ERROR(("decoding error in plane %d", i));
            /* the reference leaves a fresh (zeroed) residual plane here */
            GPU(dsvcu_frame_clear_plane(s->ctx, isP ? s->resd : dst, i, 0));
            continue;
        }
        GPU(dsvcu_dequant_plane(s->ctx, s->coefs, i, quant, &fm, nsym, lstart, dc));
        DPROF(s, DP_QUEUE);
        good_planes |= 1 << i;
    }
    /* the planes that decoded go through the inverse transform together */
    GPU(dsvcu_inv_sbt_frame(s->ctx, isP ? s->resd : dst, s->coefs, quant, &fm, good_planes));
    if (!isP && (good_planes & 1)) {
        GPU(dsvcu_intra_filter(s->ctx, quant, &fm, 0, dst, do_filter));
    }
    *fn = fno;
    if (isP) {
        if (!s->have_ref) {
            DSV_WARNING(("reference frame not found"));
            return DSV_DEC_ERROR;
        }
        GPU(dsvcu_add_pred(s->ctx, &fm, quant, s->resd, dst, ref, do_filter));
    }
    if (is_ref) {
        GPU(dsvcu_extend_frame(s->ctx, dst, 0));
    }

    DPROF(s, DP_QUEUE);
    if (tls_direct_out) {
        host = dsv_load_planar_frame(meta->subsamp, tls_direct_out, meta->width, meta->height);
    } else {
        host = dsv_mk_frame(meta->subsamp, meta->width, meta->height, 0);
    }
    for (i = 0; i < 3; i++) {
        GPU(dsvcu_frame_download(s->ctx, dst, i, host->planes[i].data, host->planes[i].stride));
    }
    if (!(tls_async && tls_direct_out) || d->draw_info) {
        GPU(dsvcu_sync(s->ctx)); /* (the overlay below is drawn into the finished host copy) */
    }
    DPROF(s, DP_TAIL);
    if (is_ref) {
        s->cur ^= 1;
        s->have_ref = 1;
    }
    if (d->draw_info) {
        overlay_info(host, p, s->blockdata, isP ? s->mvs : NULL, d->draw_info);
    }
    *out = host;
    return DSV_DEC_OK;
}

/* Where the three coefficient planes of a picture packet start, found without decoding
 * anything: every sub-stream in front of them is length-prefixed.  Mirrors the reading order
 * of decode_picture; returns -1 for anything unusual (the packet is then parsed the normal
 * way, which knows what to do with damaged input). */
static int
locate_planes(const DEC_STATE *s, const DSV_META *meta, const uint8_t *pkt, size_t len, dsvcu_plane_bits out[3],
              dsvcu_side_bits *side)
{
    DSV_BITRD br;
    const uint8_t *p;
    size_t n, at;
    unsigned ex, ey;
    int type, isP, i, nsub, stats;

    side->base = NULL;
    if (len < 64) {
        return -1;
    }
    /* the reader looks ahead 8 bytes; nothing in front of the planes may come that close to the end */
    dsv_br_init(&br, pkt, len - 8);
    type = read_packet_hdr(&br);
    if (type < 0 || !DSV_PT_IS_PIC(type)) {
        return -1;
    }
    isP = DSV_PT_HAS_REF(type);
    dsv_br_align(&br);
    (void) dsv_br_bits(&br, 32);
    dsv_br_align(&br);
    ex = dsv_br_ueg(&br);
    ey = dsv_br_ueg(&br);
    if (ex > 2 || ey > 2) {
        return -1;
    }
    dsv_br_align(&br);
    stats = (int) dsv_br_bits(&br, 3); /* stable + (maintain, ringing | mode, eprm), first bit on top */
    (void) dsv_br_bit(&br);            /* do_filter */
    (void) dsv_br_bits(&br, DSV_MAX_QP_BITS);
    if (dsv_br_bit(&br)) {
        (void) dsv_br_bits(&br, 15);
    }
    dsv_br_align(&br);
    /* stability; then motion (five sub-streams) or ringing + maintain */
    nsub = 1 + (isP ? DSV_SUB_NSUB : 2);
    for (i = 0; i < nsub; i++) {
        if (!isP || i < 2) {
            dsv_br_align(&br);
        }
        if (open_substream(&br, &p, &n)) {
            return -1;
        }
        if (isP) {
            if (i == 0) {
                side->base = p;
            }
            side->off[i] = (uint32_t) (p - side->base);
            side->len[i] = (uint32_t) n;
        }
    }
    dsv_br_align(&br);
    at = dsv_br_byte(&br);
    if (isP) {
        /* skip bits, then DSV_SUB_MODE, _MV_X, _MV_Y, _SBIM, _EPRM in the order they are stored;
         * the readers look up to 8 bytes past a sub-stream (into the planes that follow) */
        side->base_len = (uint32_t) ((size_t) (pkt + at - side->base) + 8);
        side->nbh = DSV_UDIV_ROUND_UP(meta->width, 16 << ex);
        side->nbv = DSV_UDIV_ROUND_UP(meta->height, 16 << ey);
        /* a statistic bit equal to DSV_ZERO_MARKER (1) says the bits of that stream are stored inverted */
        side->flips = ((stats >> 2) & 1) | (((stats >> 1) & 1) << 1) | ((stats & 1) << 2);
    }
    for (i = 0; i < 3; i++) {
        size_t plen;
        int cw, ch;
        if (at + 4 > len) {
            return -1;
        }
        plen = ((size_t) pkt[at] << 24) | ((size_t) pkt[at + 1] << 16) | ((size_t) pkt[at + 2] << 8) | pkt[at + 3];
        if (plen == 0 || plen > len - at - 4) {
            return -1;
        }
        dsvcu_coefs_plane_dims(s->coefs, i, &cw, &ch);
        out[i].bits = pkt + at;
        out[i].len = (uint32_t) (4 + plen);
        out[i].w = cw;
        out[i].h = ch;
        at += 4 + plen;
    }
    return 0;
}

/* Which pictures go to the device parser, and into which part of the batch.
 *
 * A plane is one serial chain; a device thread walks it at about 5 cycles per instruction,
 * ~0.2 us per (run, value) pair = ~0.25 us per byte of picture (measured,
 * profiles/r2_ncu_hzcc_parse.txt) -- ten times slower than a core.  What the device offers is
 * that all chains of a batch run side by side, beside the reconstruction of the pictures in
 * front of them, and cost the host nothing.  A part of the batch is ready when its LONGEST
 * chain is, so a picture belongs on the device if its chain is over by the time the decoder
 * gets to it -- which is a question of how much work lies in front of it in the batch:
 *
 *   the walk below keeps a clock t (ms from the start of the batch): a picture parsed on the
 *   host advances it by its parse time (~0.025 us per byte) or by the time the device needs
 *   to reconstruct a picture, whichever is longer; a picture parsed on the device by the
 *   reconstruction time.  The first picture whose chain is over at t (+ half a millisecond)
 *   opens part 0 and fixes its deadline; later pictures join part 0 if their chain meets that
 *   deadline, part 1 if it is over by the time they are reached (and not longer than the chain
 *   that opened part 1, whose waiters would otherwise wait for it), else they stay on the host.
 *
 * 1080p, qp 60, GOP of 48: the intra picture (550 KB: 137 ms on the device, 11 ms on a core)
 * stays on the host, the P pictures (50 KB: 10-12 ms) are part 0 and ready when the host is
 * done with the intra picture, the 100-190 KB pictures behind a scene cut at picture 40 are
 * part 1.  CIF: the 25 KB intra picture (6 ms) at the head of a stream stays on the host --
 * the 2 KB pictures behind it are ready in half a millisecond --, the one that leads the
 * second GOP goes to part 1.
 * dsv_set_device_entropy_limits replaces the model by fixed sizes. */
#define PREPARSE_DEV_MS_PER_BYTE 0.25e-3
#define PREPARSE_HOST_MS_PER_BYTE 0.025e-3
#define PREPARSE_SLACK_MS 0.5
static volatile size_t g_early_bytes = 0, g_late_bytes = 0;
static volatile int g_late_from = 0;

/* fixed limits instead of the model (include/dsv_session.h): packets up to early_bytes form
 * part 0, longer ones up to late_bytes part 1 unless they are among the first late_from
 * pictures of the batch; early_bytes <= 0 restores the model */
void
dsv_set_device_entropy_limits(long early_bytes, long late_bytes, int late_from)
{
    g_early_bytes = early_bytes > 0 ? (size_t) early_bytes : 0;
    g_late_bytes = late_bytes > 0 ? (size_t) late_bytes : 0;
    g_late_from = late_from > 0 ? late_from : 0;
}

/* cls[i] = 0 / 1: picture i goes into part 0 / 1 of the device batch, 2: the host parses it */
static void
preparse_classify(const size_t *len, int n, int width, int height, uint8_t *cls)
{
    const double t_pic = 0.3 + (double) width * height * 0.55e-6; /* reconstruction of one picture, ms */
    double t = 0, dl0 = -1, dl1 = -1;
    int i;
    for (i = 0; i < n; i++) {
        const double chain = (double) len[i] * PREPARSE_DEV_MS_PER_BYTE;
        if (g_early_bytes) {
            cls[i] = len[i] <= g_early_bytes ? 0 : ((len[i] <= g_late_bytes && i >= g_late_from) ? 1 : 2);
            continue;
        }
        if (dl0 < 0 && chain <= t + PREPARSE_SLACK_MS) {
            dl0 = t + PREPARSE_SLACK_MS;
        }
        if (dl0 >= 0 && chain <= dl0) {
            cls[i] = 0;
        } else if (chain + PREPARSE_SLACK_MS <= t && (dl1 < 0 || chain <= dl1)) {
            if (dl1 < 0) {
                dl1 = chain;
            }
            cls[i] = 1;
        } else {
            cls[i] = 2;
        }
        if (cls[i] == 2) {
            const double hp = (double) len[i] * PREPARSE_HOST_MS_PER_BYTE;
            t += hp > t_pic ? hp : t_pic;
        } else {
            t += t_pic;
        }
    }
}

/* Entropy-decode the coefficient planes of the next `n` picture packets on the device, in
 * one batch on the context's parse streams (they carry no coder state from one to the
 * next).  Does not wait: the result is collected when the first picture that needs it is
 * decoded.  The driver announces picture i of the batch with dsv_dec_use_parsed(set, i)
 * before handing its packet to dsv_dec; pictures whose planes were not located, are too
 * long or turn out not to be well-formed are parsed on the host as usual.  Needs the
 * metadata packet to have been seen.  Returns the set (0 / 1) the batch occupies,
 * -1 on a device error (nothing is pending then). */
int
dsv_dec_preparse(DSV_DECODER *d, const uint8_t *const *pkt, const size_t *len, int n)
{
    DEC_STATE *s;
    dsvcu_plane_bits *pl;
    dsvcu_side_bits *sd, one;
    uint8_t *cls;
    int i, m = 0, ms = 0, set, part;

    if (!d->got_metadata || n <= 0) {
        return -1;
    }
    s = state_get(d);
    if (!s) {
        return -1;
    }
    if (g_prof < 0) {
        g_prof = getenv("DSV_PROFILE") ? atoi(getenv("DSV_PROFILE")) : 0;
    }
    DPROF_START(s);
    pl = malloc((size_t) n * 3 * sizeof(*pl));
    sd = malloc((size_t) n * sizeof(*sd));
    if (!pl || !sd) {
        free(pl);
        free(sd);
        return -1;
    }
    /* The slots alternate: the one used last holds the batch the decoder is working on (or
     * about to), the other one the batch before it, which is done with.  A part of THAT batch
     * may never have been collected -- its pictures were all parsed on the host because they
     * were reached before the parser was through -- so it is collected now, which at worst
     * waits for a chain that has had two batches' time. */
    set = s->pre_last ^ 1;
    (void) preparse_collect(s, set, 0);
    (void) preparse_collect(s, set, 1);
    s->pre[set].n = 0;
    if (n > s->pre[set].cap) {
        free(s->pre[set].first);
        free(s->pre[set].side);
        s->pre[set].first = malloc((size_t) n * sizeof(int));
        s->pre[set].side = malloc((size_t) n * sizeof(int));
        s->pre[set].cap = (s->pre[set].first && s->pre[set].side) ? n : 0;
        if (!s->pre[set].cap) {
            free(s->pre[set].first);
            free(s->pre[set].side);
            s->pre[set].first = s->pre[set].side = NULL;
            free(pl);
            free(sd);
            return -1;
        }
    }
    for (i = 0; i < n; i++) {
        s->pre[set].first[i] = s->pre[set].side[i] = -1;
    }
    cls = malloc((size_t) n);
    if (!cls) {
        free(pl);
        free(sd);
        return -1;
    }
    preparse_classify(len, n, d->vidmeta.width, d->vidmeta.height, cls);
    for (part = 0; part < 2; part++) {
        for (i = 0; i < n; i++) {
            const int mine = cls[i] == part;
            if (mine && locate_planes(s, &d->vidmeta, pkt[i], len[i], pl + m, &one) == 0) {
                s->pre[set].first[i] = m;
                m += 3;
                /* the picture's side information rides along (not when the overlay wants the
                 * vectors on the host) */
                if (one.base && !d->draw_info) {
                    s->pre[set].side[i] = ms;
                    sd[ms++] = one;
                }
            }
        }
        if (!part) {
            s->pre[set].nearly = m;
            s->pre[set].nsides_early = ms;
        }
    }
    free(cls);
    s->pre[set].nspans = m;
    s->pre[set].nsides = ms;
    if (m) {
        const int got = dsvcu_parse_begin(s->ctx, pl, m, s->pre[set].nearly, sd, ms, s->pre[set].nsides_early);
        if (got < 0) {
            DSV_ERROR(("dsv_dec_preparse: %s", dsvcu_last_error()));
            free(pl);
            free(sd);
            return -1;
        }
        s->pre[set].cset = got;
        s->pre[set].pending[0] = s->pre[set].nearly > 0;
        s->pre[set].pending[1] = m > s->pre[set].nearly;
    }
    s->pre[set].n = n;
    s->pre_last = set;
    free(pl);
    free(sd);
    DPROF(s, DP_PREPARSE);
    return set;
}

/* is the first part of a batch still on its way (begun, not collected)? */
int
dsv_dec_preparse_pending(DSV_DECODER *d, int set)
{
    DEC_STATE *s = d->ref ? (DEC_STATE *) d->ref : NULL;
    return s ? s->pre[set & 1].pending[0] : 0;
}

/* wait for one part of the batch in `set` and strike the pictures that the device parser
 * refused */
static int
preparse_collect(DEC_STATE *s, int set, int part)
{
    int *ok, *sok, i;
    if (!s->pre[set].pending[part]) {
        return 0;
    }
    s->pre[set].pending[part] = 0;
    ok = malloc((size_t) (s->pre[set].nspans + s->pre[set].nsides + 1) * sizeof(int));
    sok = ok ? ok + s->pre[set].nspans : NULL;
    if (!ok || dsvcu_parse_end(s->ctx, s->pre[set].cset, part, ok, sok)) {
        DSV_ERROR(("dsv_dec_preparse: %s", ok ? dsvcu_last_error() : "out of memory"));
        free(ok);
        s->pre[set].n = 0;
        return -1;
    }
    for (i = 0; i < s->pre[set].n; i++) {
        const int f = s->pre[set].first[i], sd = s->pre[set].side[i];
        if (f >= 0 && (f >= s->pre[set].nearly) == part) {
            if (!(ok[f] && ok[f + 1] && ok[f + 2])) {
                s->pre[set].first[i] = -1;
            }
            /* (a picture's side information is in the part its planes are in) */
            if (sd >= 0 && !sok[sd]) {
                s->pre[set].side[i] = -1;
            }
        }
    }
    free(ok);
    return 0;
}

int
dsv_dec(DSV_DECODER *d, DSV_BUF *buffer, DSV_FRAME **out, DSV_FNUM *fn)
{
    DSV_BITRD br;
    uint8_t *pkt;
    int pkt_type, ret = DSV_DEC_ERROR;

    *fn = (DSV_FNUM) -1;
    *out = NULL;
    /* the bit reader looks ahead 8 bytes: parse from a zero-padded copy */
    pkt = malloc((size_t) buffer->len + 32);
    if (!pkt) {
        dsv_buf_free(buffer);
        return DSV_DEC_ERROR;
    }
    memcpy(pkt, buffer->data, buffer->len);
    memset(pkt + buffer->len, 0, 32);
    dsv_br_init(&br, pkt, buffer->len);
    pkt_type = read_packet_hdr(&br);

    if (pkt_type == -1) {
        ret = DSV_DEC_ERROR;
    } else if (!DSV_PT_IS_PIC(pkt_type)) {
        if (pkt_type == DSV_PT_META) {
            if (read_meta(d, &br) == 0) {
                d->got_metadata = 1;
                ret = DSV_DEC_GOT_META;
            }
        } else if (pkt_type == DSV_PT_EOS) {
            ret = DSV_DEC_EOS;
        }
    } else if (!d->got_metadata) {
        DSV_WARNING(("no metadata, skipping frame"));
        ret = DSV_DEC_OK;
    } else {
        DEC_STATE *s = state_get(d);
        ret = s ? decode_picture(d, s, &br, pkt_type, out, fn) : DSV_DEC_ERROR;
    }
    free(pkt);
    dsv_buf_free(buffer);
    return ret;
}

/* dsv_post_process (reference bmc.c:340-361, called by the CLI for -postsharp):
 * decoder-side sharpening of a HOST luma plane.  The plane makes a round trip
 * through the device; the context is kept per thread and per plane size. */
void
dsv_post_process(DSV_PLANE *dp)
{
    static __thread dsvcu_ctx *ctx = NULL;
    static __thread dsvcu_frame *fr = NULL;
    static __thread int cw = 0, ch = 0;
    if (!dp || !dp->data) {
        return;
    }
    if (ctx && (cw != dp->w || ch != dp->h)) {
        dsvcu_frame_destroy(ctx, fr);
        dsvcu_ctx_destroy(ctx);
        ctx = NULL;
        fr = NULL;
    }
    if (!ctx) {
        if (dsvcu_ctx_create(&ctx, dsv_get_thread_device(), dp->w, dp->h, DSV_SUBSAMP_420) ||
            dsvcu_frame_create_luma(ctx, &fr, dp->w, dp->h)) {
            DSV_ERROR(("dsv_post_process: %s", dsvcu_last_error()));
            ctx = NULL;
            return;
        }
        cw = dp->w;
        ch = dp->h;
    }
    if (dsvcu_frame_upload(ctx, fr, 0, dp->data, dp->stride) || dsvcu_post_process(ctx, fr) ||
        dsvcu_frame_download(ctx, fr, 0, dp->data, dp->stride) || dsvcu_sync(ctx)) {
        DSV_ERROR(("dsv_post_process: %s", dsvcu_last_error()));
    }
}
