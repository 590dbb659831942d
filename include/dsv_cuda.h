/*
 * dsv_cuda.h -- C ABI of the B200 (sm_100a) pixel path of DSV2.
 *
 * This is the one new boundary the B200 build adds to the codec: host C
 * (bitstream, rate control, I/O) on one side, device-resident frames and
 * hand-written kernels on the other.  Every entry point is plain C: opaque
 * handles, raw pointers and sizes, `int` status (0 = ok, negative = failure;
 * dsvcu_last_error() gives the text).  Nothing here ever falls back to a CPU
 * implementation: without a usable CUDA device dsvcu_ctx_create() fails.
 *
 * Each stage call replaces one operator of the reference's internal boundary
 * (reference src/dsv_internal.h:112-147, src/dsv_encoder.h:215, src/dsv.h:
 * 232-237); the reference file:line is given next to each prototype.  All stage
 * calls are asynchronous on the context's stream; dsvcu_sync() waits.
 */
#ifndef DSV2_B200_DSV_CUDA_H
#define DSV2_B200_DSV_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dsvcu_ctx dsvcu_ctx;
typedef struct dsvcu_frame dsvcu_frame; /* u8 planar frame, 32-px border, on device */
typedef struct dsvcu_coefs dsvcu_coefs; /* int32 subband planes, on device */

/* per-frame parameters the operators read (subset of DSV_PARAMS/DSV_FMETA,
 * reference dsv.h:242-268, dsv_internal.h:40-47) */
typedef struct {
    int isP;
    int lossless;
    int do_psy;
    int blk_w, blk_h;
    int nblocks_h, nblocks_v;
    int temporal_mc;
    int inter_sharpen;
    int effort;
    unsigned fnum;
} dsvcu_fmeta;

/* one coded coefficient: position in HZCC scan order + quantised value */
typedef struct {
    uint32_t pos;
    int32_t v;
} dsvcu_symbol;

/* ---- context ---- */
int dsvcu_device_count(void);
const char *dsvcu_last_error(void);
int dsvcu_ctx_create(dsvcu_ctx **out, int device, int width, int height, int subsamp);
void dsvcu_ctx_destroy(dsvcu_ctx *ctx);
/* the CUDA stream (cudaStream_t) all work of this context is issued on */
void *dsvcu_ctx_stream(dsvcu_ctx *ctx);
int dsvcu_sync(dsvcu_ctx *ctx);

/* page-locked host memory (frame / stream staging); NULL on failure */
void *dsvcu_host_alloc(size_t bytes);
void dsvcu_host_free(void *p);

/* ---- device objects ---- */
/* frame with the stream geometry (3 planes) or a luma-only frame of w x h */
int dsvcu_frame_create(dsvcu_ctx *ctx, dsvcu_frame **out);
int dsvcu_frame_create_luma(dsvcu_ctx *ctx, dsvcu_frame **out, int w, int h);
void dsvcu_frame_destroy(dsvcu_ctx *ctx, dsvcu_frame *f);
int dsvcu_frame_plane_dims(dsvcu_frame *f, int plane, int *w, int *h, int *stride);
/* host <-> device, visible w x h area of one plane (host rows `hstride` apart) */
int dsvcu_frame_upload(dsvcu_ctx *ctx, dsvcu_frame *f, int plane, const uint8_t *src, int hstride);
int dsvcu_frame_download(dsvcu_ctx *ctx, dsvcu_frame *f, int plane, uint8_t *dst, int hstride);
/* memset of the visible area of one plane */
int dsvcu_frame_clear_plane(dsvcu_ctx *ctx, dsvcu_frame *f, int plane, int value);
/* whole bordered plane, (h+64) rows of `stride` bytes; used by tests of the border */
int dsvcu_frame_upload_bordered(dsvcu_ctx *ctx, dsvcu_frame *f, int plane, const uint8_t *src);
int dsvcu_frame_download_bordered(dsvcu_ctx *ctx, dsvcu_frame *f, int plane, uint8_t *dst);

int dsvcu_coefs_create(dsvcu_ctx *ctx, dsvcu_coefs **out);
void dsvcu_coefs_destroy(dsvcu_ctx *ctx, dsvcu_coefs *c);
int dsvcu_coefs_plane_dims(dsvcu_coefs *c, int plane, int *w, int *h);
int dsvcu_coefs_upload(dsvcu_ctx *ctx, dsvcu_coefs *c, int plane, const int32_t *src);
int dsvcu_coefs_download(dsvcu_ctx *ctx, dsvcu_coefs *c, int plane, int32_t *dst);

/* per-frame block side information (reference DSV_FMETA.blockdata / .mvs) */
int dsvcu_set_blockdata(dsvcu_ctx *ctx, const uint8_t *blockdata, int nblocks);
int dsvcu_set_mvs(dsvcu_ctx *ctx, const void *dsv_mv_array, int nblocks);
/* both in one host-to-device copy through pinned staging (dsv_mv_array may be NULL);
 * not to be called again before the next wait on the context */
int dsvcu_set_side(dsvcu_ctx *ctx, const uint8_t *blockdata, const void *dsv_mv_array, int nblocks);

/* ---- subband transforms ---- */
/* dsv_fwd_sbt, reference sbt.c:847-886 (dsv_internal.h:112) */
int dsvcu_fwd_sbt(dsvcu_ctx *ctx, dsvcu_frame *src, int plane, dsvcu_coefs *dst, const dsvcu_fmeta *fm);
/* dsv_inv_sbt, reference sbt.c:889-934 (dsv_internal.h:113) */
int dsvcu_inv_sbt(dsvcu_ctx *ctx, dsvcu_frame *dst, int plane, dsvcu_coefs *src, int q, const dsvcu_fmeta *fm);
/* the planes selected by plane_mask (bit p = plane p) through shared launches: same results as
 * the per-plane calls, one launch per pyramid level for all of them */
int dsvcu_fwd_sbt_frame(dsvcu_ctx *ctx, dsvcu_frame *src, dsvcu_coefs *dst, const dsvcu_fmeta *fm, int plane_mask);
int dsvcu_inv_sbt_frame(dsvcu_ctx *ctx, dsvcu_frame *dst, dsvcu_coefs *src, int q, const dsvcu_fmeta *fm, int plane_mask);

/* ---- quantisation: arithmetic half of dsv_encode_plane / dsv_decode_plane ---- */
/* reference hzcc.c:254-448 (quantise in place, leave the de-quantised value,
 * produce the ordered symbol list on the device) */
int dsvcu_quant_plane(dsvcu_ctx *ctx, dsvcu_coefs *c, int plane, int q, const dsvcu_fmeta *fm);
/* the planes selected by plane_mask through shared launches (same results as the per-plane calls) */
int dsvcu_quant_frame(dsvcu_ctx *ctx, dsvcu_coefs *c, int q, const dsvcu_fmeta *fm, int plane_mask);
/* wait for the symbols of `plane`; pointers stay valid until the next quant of
 * that plane.  *dc receives coefficient 0 (sent raw, hzcc.c:599-602) */
int dsvcu_fetch_symbols(dsvcu_ctx *ctx, int plane, const dsvcu_symbol **syms, int *nsyms, int *dc);
/* reference hzcc.c:450-583.  `syms` must come from dsvcu_symbol_staging(); the
 * list is in scan order; level_start[0..4] = index of the first symbol of the
 * LL part, level 0, 1, 2 and the end.  Zero-fills the plane first. */
dsvcu_symbol *dsvcu_symbol_staging(dsvcu_ctx *ctx, int plane, int *capacity);
/* the staging buffers (symbols, dsvcu_set_side) exist twice: switch to the other set, waiting
 * until the device has copied everything out of it.  Lets a decoder parse the next picture
 * while the previous one is still in flight; pointers from dsvcu_symbol_staging are per set. */
int dsvcu_staging_flip(dsvcu_ctx *ctx);
int dsvcu_dequant_plane(dsvcu_ctx *ctx, dsvcu_coefs *c, int plane, int q, const dsvcu_fmeta *fm,
                        int nsyms, const int level_start[5], int dc);
/* Entropy decode on the device: the bit-parsing half of dsv_decode_plane (reference
 * hzcc.c:450-583, :585-649; codes of bs.c:96-137, :237-251) for a BATCH of planes -- typically
 * the 3 x N planes of the N pictures of a closed GOP, which carry no entropy-coder state from
 * one to the next.  pl[i].bits points at plane i's 32-bit length word, pl[i].len = 4 + that
 * length, (w, h) = dsvcu_coefs_plane_dims.
 *   dsvcu_parse_begin  gathers, uploads and launches on the context's parse streams (beside
 *                      the work queued on its main stream) and returns the set (0 or 1) the
 *                      batch occupies; at most two batches exist at a time, a set's symbols
 *                      live until the set is begun again.  The batch has two parts, planes
 *                      [0, n_early) and [n_early, n), each with a launch and a completion of
 *                      its own: a part is ready when its longest plane is, so short planes
 *                      that are needed first go into part 0, long ones needed late into part 1.
 *   dsvcu_parse_end    waits for one part (0 / 1) of the batch.  ok[i] = 1: symbols of plane i
 *                      are resident on the device; ok[i] = 0: not a well-formed plane, parse
 *                      it on the host (dsvcu_dequant_plane) -- the host parser reproduces the
 *                      reference's handling of damaged planes.  Only the entries of that part
 *                      are written.
 *   dsvcu_parse_planes one part, begun and collected in one call; returns the set. */
typedef struct {
    const uint8_t *bits;
    uint32_t len;
    int w, h;
} dsvcu_plane_bits;
/* Side information of an inter picture for the same batch (reference dsv_decoder.c:
 * decode_stability_blocks :77-101, decode_motion :137-222; spec B.2.3.1, B.2.3.4): `base` points
 * into the packet, at or in front of the first of the six sub-streams' data -- skip bits, mode
 * bits, vector x, vector y, intra sub-block masks + DC, EPRM bits -- and base_len bytes are
 * copied (they must reach 8 bytes past the last sub-stream: the readers look that far);
 * off[k] / len[k] = start (relative to base) and coded length of sub-stream k; flips: bit 0 / 1 /
 * 2 = the skip / mode / EPRM bits are stored inverted.  side_ok[i] = 0: a reader ran past the
 * end of its sub-stream -- decode that picture's side information on the host. */
typedef struct {
    const uint8_t *base;
    uint32_t base_len;
    uint32_t off[6], len[6];
    int nbh, nbv;
    int flips;
} dsvcu_side_bits;
int dsvcu_parse_begin(dsvcu_ctx *ctx, const dsvcu_plane_bits *pl, int n, int n_early, const dsvcu_side_bits *sd,
                      int nsd, int nsd_early);
int dsvcu_parse_end(dsvcu_ctx *ctx, int set, int part, int *ok, int *side_ok);
/* 1: that part of the batch has been parsed (dsvcu_parse_end will not wait), 0: still running */
int dsvcu_parse_ready(dsvcu_ctx *ctx, int set, int part);
/* picture `side` of a collected batch: its vector field and block flags become the context's
 * current side information (what dsvcu_set_side uploads), device to device */
int dsvcu_set_side_parsed(dsvcu_ctx *ctx, int set, int side, int nblocks);
int dsvcu_parse_planes(dsvcu_ctx *ctx, const dsvcu_plane_bits *pl, int n, int *ok);
/* symbols found in plane `span` of a collected batch, -1 if it was not ok */
int dsvcu_parsed_count(dsvcu_ctx *ctx, int set, int span);
/* the three planes of one picture, parsed as spans first_span .. first_span + 2 of a collected
 * batch: zero-fill + de-quantise (hzcc.c:450-583) without the symbols leaving the device */
int dsvcu_dequant_parsed(dsvcu_ctx *ctx, dsvcu_coefs *c, int q, const dsvcu_fmeta *fm, int set, int first_span);
/* number of scan positions of a plane and the first scan position of each
 * HZCC part (LL, level 0, 1, 2, end) -- the host coder needs them */
int dsvcu_scan_layout(int w, int h, int part_start[5]);

/* ---- motion compensation, reconstruction, filters ---- */
/* dsv_sub_pred, reference bmc.c:1057-1070 */
int dsvcu_sub_pred(dsvcu_ctx *ctx, const dsvcu_fmeta *fm, dsvcu_frame *pred, dsvcu_frame *resd, dsvcu_frame *ref);
/* the same with the source read from `src` and the residual written to `resd`
 * (saves cloning the source into the residual frame, dsv_encoder.c:1292) */
int dsvcu_sub_pred_from(dsvcu_ctx *ctx, const dsvcu_fmeta *fm, dsvcu_frame *pred, dsvcu_frame *resd, dsvcu_frame *ref,
                        dsvcu_frame *src);
/* dsv_add_pred, reference bmc.c:1093-1111 */
int dsvcu_add_pred(dsvcu_ctx *ctx, const dsvcu_fmeta *fm, int q, dsvcu_frame *resd, dsvcu_frame *out,
                   dsvcu_frame *ref, int do_filter);
/* dsv_add_res, reference bmc.c:1072-1090 */
int dsvcu_add_res(dsvcu_ctx *ctx, const dsvcu_fmeta *fm, int q, dsvcu_frame *resd, dsvcu_frame *pred, int do_filter);
/* dsv_intra_filter, reference bmc.c:390-457 */
int dsvcu_intra_filter(dsvcu_ctx *ctx, int q, const dsvcu_fmeta *fm, int plane, dsvcu_frame *f, int do_filter);
/* dsv_post_process, reference bmc.c:340-361 (luma) */
int dsvcu_post_process(dsvcu_ctx *ctx, dsvcu_frame *f);

/* ---- frame helpers ---- */
/* dsv_extend_frame / dsv_extend_frame_luma, reference frame.c:413-434 */
int dsvcu_extend_frame(dsvcu_ctx *ctx, dsvcu_frame *f, int luma_only);
/* dsv_ds2x_frame_luma, reference frame.c:210-234 */
int dsvcu_ds2x_luma(dsvcu_ctx *ctx, dsvcu_frame *dst, dsvcu_frame *src);
/* dsv_frame_copy, reference frame.c:185-207 (copies, then extends dst) */
int dsvcu_frame_copy(dsvcu_ctx *ctx, dsvcu_frame *dst, dsvcu_frame *src);

/* ---- motion estimation / block analysis (encoder) ---- */
typedef struct dsvcu_pyramid dsvcu_pyramid; /* luma-only 2x pyramid, levels 1..n */
int dsvcu_pyramid_create(dsvcu_ctx *ctx, dsvcu_pyramid **out, int levels);
void dsvcu_pyramid_destroy(dsvcu_ctx *ctx, dsvcu_pyramid *p);
/* mk_pyramid, reference dsv_encoder.c:493-516 (ds2x + luma border per level) */
int dsvcu_pyramid_build(dsvcu_ctx *ctx, dsvcu_pyramid *p, dsvcu_frame *base);
/* dsv_extend_frame(base) + mk_pyramid(base) fused: the border of every plane of
 * `base`, then all pyramid levels with their borders, in two launches */
int dsvcu_extend_pyramid(dsvcu_ctx *ctx, dsvcu_frame *base, dsvcu_pyramid *p);
dsvcu_frame *dsvcu_pyramid_level(dsvcu_pyramid *p, int level); /* 1..n */

typedef struct {
    int quant;             /* DSV_HME.quant = previous picture's quantiser */
    int skip_block_thresh; /* DSV_ENCODER.skip_block_thresh */
    int pyramid_levels;
    int use_prev_mvs;      /* DSV_HME.ref_mvf != NULL */
} dsvcu_hme_params;

/* the previous picture's final vector field (DSV_HME.ref_mvf), from the host ... */
int dsvcu_set_prev_mvs(dsvcu_ctx *ctx, const void *dsv_mv_array, int nblocks);
/* ... or from the device's current MV array (after dsvcu_hme / dsvcu_set_mvs) */
int dsvcu_mvs_to_prev(dsvcu_ctx *ctx, int nblocks);
/* ... or without a copy: the current array BECOMES the previous picture's field and the
 * current array / blockdata are undefined afterwards (call after the picture's last
 * operator that reads them has been queued) */
int dsvcu_mvs_swap_prev(dsvcu_ctx *ctx, int nblocks);
/* dsv_hme, reference hme.c:2001-2016 (struct DSV_HME, dsv_encoder.h:202-213).
 * Leaves the final field on the device as the current MV array (as if
 * dsvcu_set_mvs had been called with it). */
int dsvcu_hme(dsvcu_ctx *ctx, const dsvcu_fmeta *fm, const dsvcu_hme_params *hp, dsvcu_frame *src,
              dsvcu_pyramid *src_pyr, dsvcu_frame *ref, dsvcu_pyramid *ref_pyr, dsvcu_frame *ogr,
              dsvcu_pyramid *ogr_pyr);
/* waits; copies the field to `mvs_out` (nblocks DSV_MV) and the three scalars
 * dsv_hme returns (intra %, scene-change blocks %, average error) */
int dsvcu_hme_fetch(dsvcu_ctx *ctx, void *mvs_out, int nblocks, int *intra_pct, int *scene_change_blocks,
                    int *avg_err);
/* work counters of the last dsvcu_hme, valid after dsvcu_hme_fetch: [0] full-block metric
 * evaluations, [1] sub-pel position metrics (the unit SURVEY section 8d asks the search to be measured in) */
int dsvcu_hme_counters(dsvcu_ctx *ctx, long long out[2]);
/* dsv_intra_analysis, reference hme.c:1835-1971; result via dsvcu_hme_fetch-like copy */
int dsvcu_intra_analysis(dsvcu_ctx *ctx, const dsvcu_fmeta *fm, dsvcu_frame *src, void *mvs_out, int nblocks);
/* the same in two halves: queue the kernel + copy, then wait and read */
int dsvcu_intra_analysis_async(dsvcu_ctx *ctx, const dsvcu_fmeta *fm, dsvcu_frame *src, int nblocks);
int dsvcu_intra_analysis_fetch(dsvcu_ctx *ctx, void *mvs_out, int nblocks);
/* frame_luma_avg, reference dsv_encoder.c:108-127 (sum of per-row averages / h); waits */
int dsvcu_frame_luma_avg(dsvcu_ctx *ctx, dsvcu_frame *f, unsigned *avg);
/* queued form; the result is valid after the next wait on the context's stream */
int dsvcu_frame_luma_avg_async(dsvcu_ctx *ctx, dsvcu_frame *f);
unsigned dsvcu_frame_luma_avg_result(dsvcu_ctx *ctx);

/* floor(sqrt(n)) exactly as the motion search computes it (reference iisqrt, hme.c:99-124) */
unsigned dsvcu_isqrt(unsigned n);

/* ---- timing on the context's stream (CUDA events) ---- */
int dsvcu_timer_start(dsvcu_ctx *ctx);
int dsvcu_timer_stop_ms(dsvcu_ctx *ctx, float *ms); /* synchronises */
#define DSVCU_MARKS 8
int dsvcu_mark(dsvcu_ctx *ctx, int k);                                 /* stamp k on the context's stream */
int dsvcu_mark_elapsed_ms(dsvcu_ctx *ctx, int a, int b, float *ms);   /* waits for stamp b */
/* number of kernels this context has launched so far */
long long dsvcu_launch_count(dsvcu_ctx *ctx);
/* ... and every context of this process together */
long long dsvcu_total_launches(void);

#ifdef __cplusplus
}
#endif
#endif
