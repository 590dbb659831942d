/*
 * dsv_encoder.h -- public encoder API, B200 build.
 *
 * Declares the reference encoder interface (reference src/dsv_encoder.h:17-221)
 * unchanged for callers: the DSV_ENCODER configuration/state block with the same
 * field order, the rate-control and effort constants, and dsv_enc_init /
 * dsv_enc_set_metadata / dsv_enc_start / dsv_enc / dsv_enc_end_of_stream /
 * dsv_enc_free.  Rate control, scene-change detection, GOP logic and bit
 * packing run on the host exactly as in the reference; every pixel operator
 * (pyramid, motion search, prediction, transforms, quantisation,
 * reconstruction, filters) runs on the GPU through dsv_cuda.h.
 */
#ifndef DSV2_B200_DSV_ENCODER_H
#define DSV2_B200_DSV_ENCODER_H

#include <limits.h>
#include "dsv.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DSV_ENCODER_VERSION 14

#define DSV_GOP_INTRA 0
#define DSV_GOP_INF INT_MAX

#define DSV_ENC_NUM_BUFS 0x03
#define DSV_ENC_FINISHED 0x04

#define DSV_MIN_EFFORT 0
#define DSV_MAX_EFFORT 10

#define DSV_RATE_CONTROL_CRF 0
#define DSV_RATE_CONTROL_ABR 1
#define DSV_RATE_CONTROL_CQP 2

#define DSV_MAX_PYRAMID_LEVELS 5

#define DSV_RC_QUAL_SCALE 4
#define DSV_MAX_QUALITY (100)
#define DSV_RC_QUAL_MAX ((DSV_MAX_QUALITY * DSV_RC_QUAL_SCALE))
#define DSV_USER_QUAL_TO_RC_QUAL(user) ((user) * DSV_RC_QUAL_SCALE)
#define DSV_QUALITY_PERCENT(pct) (pct)

#define DSV_PSY_ADAPTIVE_QUANT (1 << 0)
#define DSV_PSY_CONTENT_ANALYSIS (1 << 1)
#define DSV_PSY_I_VISUAL_MASKING (1 << 2)
#define DSV_PSY_P_VISUAL_MASKING (1 << 3)
#define DSV_PSY_ADAPTIVE_RINGING (1 << 4)
#define DSV_PSY_ALL 0xff

#define DSV_RF_RESET 256

struct _DSV_ENCDATA; /* opaque in this build: device-side picture state */
typedef struct _DSV_ENCDATA DSV_ENCDATA;

typedef struct {
    int quality;
    int effort;
    int gop;
    int do_scd;
    int do_temporal_aq;
    int do_psy;
    int do_dark_intra_boost;
    int do_intra_filter;
    int do_inter_filter;
    int skip_block_thresh;
    int block_size_override_x;
    int block_size_override_y;
    int variable_i_interval;
    int rc_mode;
    unsigned bitrate;
    int rc_pergop;
    int min_q_step;
    int max_q_step;
    int min_quality;
    int max_quality;
    int min_I_frame_quality;
    int prev_I_frame_quality;
    int intra_pct_thresh;
    int scene_change_pct;
    unsigned stable_refresh;
    int pyramid_levels;

    struct DSV_STATS {
        unsigned inum, pnum;
        unsigned iqual, pqual;
        unsigned iminq, pminq;
        unsigned imaxq, pmaxq;
        unsigned isize, psize;
        unsigned imins, pmins;
        unsigned imaxs, pmaxs;
        unsigned mb, mbI, mbP, mbdc, mbsub;
        unsigned mbsubs[4];
        unsigned eprm, skip;
        unsigned fpx, hpx, qpx;
        unsigned fpy, hpy, qpy;
        unsigned ifnum, pfnum;
    } stats;

    /* internal state (same slots as the reference) */
    unsigned rc_qual;
    unsigned rf_total;
    unsigned rf_reset;
    int rf_avg;
    int total_P_frame_q;
    int avg_P_frame_q;
    int prev_complexity;
    int curr_complexity;
    int curr_avgmot;
    int curr_intra_pct;
    int curr_scblocks;
    int prev_chaos;
    int motion_chaos;
    int motion_static;
    int avg_err;
    int auto_filter;

    void (*frame_callback)(DSV_META *m, DSV_FRAME *orig, DSV_FRAME *recon);

    DSV_FNUM next_fnum;
    DSV_ENCDATA *ref; /* B200 build: owns the device-side encoder state */
    DSV_META vidmeta;
    int prev_link;
    int force_metadata;

    struct DSV_STAB_ACC {
        int32_t x, y;
    } *stability;
    unsigned refresh_ctr;
    uint8_t *blockdata;
    uint8_t *intra_map;

    DSV_FNUM prev_gop;
    int prev_quant;
} DSV_ENCODER;

extern void dsv_enc_init(DSV_ENCODER *enc);
extern void dsv_enc_free(DSV_ENCODER *enc);
extern void dsv_enc_set_metadata(DSV_ENCODER *enc, DSV_META *md);
extern void dsv_enc_force_metadata(DSV_ENCODER *enc);
extern void dsv_enc_start(DSV_ENCODER *enc);
/* encode one frame (consumed); returns how many buffers were produced (1 or 2:
 * optional metadata packet, then the picture packet) */
extern int dsv_enc(DSV_ENCODER *enc, DSV_FRAME *frame, DSV_BUF *bufs);
extern void dsv_enc_end_of_stream(DSV_ENCODER *enc, DSV_BUF *bufs);

#ifdef __cplusplus
}
#endif
#endif
