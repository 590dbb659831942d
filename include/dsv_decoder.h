/*
 * dsv_decoder.h -- public decoder API, B200 build.
 *
 * Same types, constants and calls as the reference decoder interface
 * (reference src/dsv_decoder.h:17-67): a zero-initialised DSV_DECODER, one
 * dsv_dec() call per packet, DSV_DEC_* status codes, dsv_get_metadata() and
 * dsv_dec_free().  Behind it every picture is reconstructed on the GPU
 * (dequantisation, inverse subband transform, motion compensation, in-loop
 * filters) and only the finished frame is copied back to a host DSV_FRAME.
 */
#ifndef DSV2_B200_DSV_DECODER_H
#define DSV2_B200_DSV_DECODER_H

#include "dsv.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DSV_DECODER_VERSION 2

typedef struct {
    DSV_PARAMS params;
    DSV_FRAME *out_frame;
    DSV_FRAME *ref_frame;
    uint8_t *blockdata;
    int refcount;
} DSV_IMAGE;

typedef struct {
    DSV_META vidmeta;
    DSV_IMAGE *ref; /* B200 build: owns the device-side decoder state */
#define DSV_DRAW_STABHQ 1
#define DSV_DRAW_MOVECS 2
#define DSV_DRAW_IBLOCK 4
    int draw_info;
    int got_metadata;
} DSV_DECODER;

#define DSV_DEC_OK 0
#define DSV_DEC_ERROR 1
#define DSV_DEC_EOS 2
#define DSV_DEC_GOT_META 3
#define DSV_DEC_NEED_NEXT 4

/* decode one packet; consumes `buf`; on DSV_DEC_OK *out holds a new reference
 * to the decoded frame (release with dsv_frame_ref_dec) and *fn its number */
extern int dsv_dec(DSV_DECODER *d, DSV_BUF *buf, DSV_FRAME **out, DSV_FNUM *fn);
extern DSV_META *dsv_get_metadata(DSV_DECODER *d);
extern void dsv_dec_free(DSV_DECODER *d);
/* decoder-side luma sharpening of a host plane (reference dsv_internal.h:147,
 * bmc.c:340-361; what the CLI's -postsharp calls) */
extern void dsv_post_process(DSV_PLANE *dp);

#ifdef __cplusplus
}
#endif
#endif
