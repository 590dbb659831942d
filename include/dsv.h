/*
 * dsv.h -- public types of the DSV2 (bitstream v2.8) codec API, B200 build.
 *
 * Drop-in declaration of the reference's public interface (reference
 * src/dsv.h:17-330): identical type names, struct layouts, constants and
 * function names so that a caller written against the reference (its CLI,
 * src/dsv_main.c) compiles and links against libdsv2cuda.so unchanged.  The
 * implementation behind it is new: pixel operators run as sm_100a kernels on
 * device-resident frames (see dsv_cuda.h); only bit packing, rate control and
 * I/O stay on the host.
 */
#ifndef DSV2_B200_DSV_H
#define DSV2_B200_DSV_H

#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- packet framing (reference dsv.h:29-50) ---- */
#define DSV_FOURCC_0 'D'
#define DSV_FOURCC_1 'S'
#define DSV_FOURCC_2 'V'
#define DSV_FOURCC_3 '2'
#define DSV_VERSION_MINOR 8

#define DSV_PT_META 0x00
#define DSV_PT_PIC 0x04
#define DSV_PT_EOS 0x10
#define DSV_MAKE_PT(is_ref, has_ref) (DSV_PT_PIC | ((is_ref) << 1) | (has_ref))
#define DSV_PT_IS_PIC(t) ((t) & DSV_PT_PIC)
#define DSV_PT_IS_REF(t) (((t) & 0x6) == 0x6)
#define DSV_PT_HAS_REF(t) ((t) & 0x1)

#define DSV_PACKET_HDR_SIZE 14 /* fourcc, minor, type, prev link, next link */
#define DSV_PACKET_TYPE_OFFSET 5
#define DSV_PACKET_PREV_OFFSET 6
#define DSV_PACKET_NEXT_OFFSET 10

#define DSV_MIN_BLOCK_SIZE 16
#define DSV_MAX_BLOCK_SIZE 32

/* ---- arithmetic helpers (reference dsv.h:56-81) ---- */
#ifndef MIN
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#endif
#ifndef MAX
#define MAX(a, b) ((a) > (b) ? (a) : (b))
#endif
#ifndef CLAMP
#define CLAMP(x, lo, hi) ((x) < (lo) ? (lo) : ((x) > (hi) ? (hi) : (x)))
#endif
#define DSV_ROUND_SHIFT(x, s) (((x) + (1 << (s)) - 1) >> (s))
#define DSV_ROUND_POW2(x, p) (((x) + (1 << (p)) - 1) & ((unsigned) (~0) << (p)))
#define DSV_UDIV_ROUND_UP(a, b) (((a) + (b) - 1) / (b))
#define DSV_UDIV_ROUND(a, b) (((a) + ((b) / 2)) / (b))
#define DSV_SAR(v, s) ((v) < 0 ? ~(~(v) >> (s)) : (v) >> (s))
#define DSV_SAR_R(v, s) DSV_SAR((v) + (1 << ((s) - 1)), (s))

/* ---- chroma formats (reference dsv.h:83-101) ---- */
#define DSV_FMT_FULL_V 0x0
#define DSV_FMT_DIV2_V 0x1
#define DSV_FMT_DIV4_V 0x2
#define DSV_FMT_FULL_H 0x0
#define DSV_FMT_DIV2_H 0x4
#define DSV_FMT_DIV4_H 0x8
#define DSV_SUBSAMP_444 (DSV_FMT_FULL_H | DSV_FMT_FULL_V)
#define DSV_SUBSAMP_422 (DSV_FMT_DIV2_H | DSV_FMT_FULL_V)
#define DSV_SUBSAMP_UYVY (0x10 | DSV_SUBSAMP_422)
#define DSV_SUBSAMP_420 (DSV_FMT_DIV2_H | DSV_FMT_DIV2_V)
#define DSV_SUBSAMP_411 (DSV_FMT_DIV4_H | DSV_FMT_FULL_V)
#define DSV_SUBSAMP_410 (DSV_FMT_DIV4_H | DSV_FMT_DIV4_V)
#define DSV_FORMAT_H_SHIFT(f) (((f) >> 2) & 0x3)
#define DSV_FORMAT_V_SHIFT(f) ((f) & 0x3)

typedef uint32_t DSV_FNUM;

typedef struct {
    int width, height, subsamp;
    int fps_num, fps_den;
    int aspect_num, aspect_den;
    int inter_sharpen;
    int reserved;
} DSV_META;

typedef struct {
    uint8_t *data;
    int len;
    int format;
    int stride;
    int w, h;
} DSV_PLANE;

typedef int32_t DSV_SBC;
typedef struct {
    DSV_SBC *data;
    int width, height;
} DSV_COEFS;

typedef struct {
    uint8_t *alloc;
    DSV_PLANE planes[3];
    int refcount;
    int format;
    int width, height;
    int border;
} DSV_FRAME;

#define DSV_NDIF_THRESH (2 * 4)

#define DSV_STABLE_STAT 0
#define DSV_MAINTAIN_STAT 1
#define DSV_RINGING_STAT 2
#define DSV_MODE_STAT 3
#define DSV_EPRM_STAT 4
#define DSV_MAX_STAT 5
#define DSV_ONE_MARKER 0
#define DSV_ZERO_MARKER 1

#define DSV_MODE_INTER 0
#define DSV_MODE_INTRA 1
#define DSV_MASK_INTRA00 1
#define DSV_MASK_INTRA01 2
#define DSV_MASK_INTRA10 4
#define DSV_MASK_INTRA11 8
#define DSV_MASK_ALL_INTRA 15

/* per-block motion record, 16 bytes (reference dsv.h:171-216) */
typedef struct {
    union {
        struct {
            int16_t x, y;
        } mv;
        int32_t all;
    } u;
    uint32_t flags;
    uint16_t err;
    uint16_t dc;
    uint8_t submask;
} DSV_MV;

#define DSV_SRC_DC_PRED 0x100
#define DSV_IS_SUBPEL(v) (((v)->u.mv.x | (v)->u.mv.y) & 3)
#define DSV_IS_QPEL(v) (((v)->u.mv.x | (v)->u.mv.y) & 1)
#define DSV_IS_DIAG(v) (((v)->u.mv.x & 3) && ((v)->u.mv.y & 3))
#define DSV_TEMPORAL_MC(fno) ((fno) % 2)

#define DSV_MV_BIT_INTRA 0
#define DSV_MV_BIT_EPRM 1
#define DSV_MV_BIT_MAINTAIN 2
#define DSV_MV_BIT_SKIP 3
#define DSV_MV_BIT_RINGING 4
#define DSV_MV_BIT_NOXMITY 5
#define DSV_MV_BIT_NOXMITC 6
#define DSV_MV_BIT_SIMCMPLX 7
#define DSV_MV_TEST(mv, bit) ((mv)->flags & (1u << (bit)))
#define DSV_MV_IS_INTRA(mv) DSV_MV_TEST(mv, DSV_MV_BIT_INTRA)
#define DSV_MV_IS_EPRM(mv) DSV_MV_TEST(mv, DSV_MV_BIT_EPRM)
#define DSV_MV_IS_MAINTAIN(mv) DSV_MV_TEST(mv, DSV_MV_BIT_MAINTAIN)
#define DSV_MV_IS_SKIP(mv) DSV_MV_TEST(mv, DSV_MV_BIT_SKIP)
#define DSV_MV_IS_RINGING(mv) DSV_MV_TEST(mv, DSV_MV_BIT_RINGING)
#define DSV_MV_IS_NOXMITY(mv) DSV_MV_TEST(mv, DSV_MV_BIT_NOXMITY)
#define DSV_MV_IS_NOXMITC(mv) DSV_MV_TEST(mv, DSV_MV_BIT_NOXMITC)
#define DSV_MV_IS_SIMCMPLX(mv) DSV_MV_TEST(mv, DSV_MV_BIT_SIMCMPLX)
#define DSV_BIT_SET(v, b, on) ((v) &= ~(1 << (b)), (v) |= ((on) << (b)))
#define DSV_MV_SET_INTRA(mv, b) DSV_BIT_SET((mv)->flags, DSV_MV_BIT_INTRA, b)
#define DSV_MV_SET_EPRM(mv, b) DSV_BIT_SET((mv)->flags, DSV_MV_BIT_EPRM, b)
#define DSV_MV_SET_MAINTAIN(mv, b) DSV_BIT_SET((mv)->flags, DSV_MV_BIT_MAINTAIN, b)
#define DSV_MV_SET_SKIP(mv, b) DSV_BIT_SET((mv)->flags, DSV_MV_BIT_SKIP, b)
#define DSV_MV_SET_RINGING(mv, b) DSV_BIT_SET((mv)->flags, DSV_MV_BIT_RINGING, b)
#define DSV_MV_SET_NOXMITY(mv, b) DSV_BIT_SET((mv)->flags, DSV_MV_BIT_NOXMITY, b)
#define DSV_MV_SET_NOXMITC(mv, b) DSV_BIT_SET((mv)->flags, DSV_MV_BIT_NOXMITC, b)
#define DSV_MV_SET_SIMCMPLX(mv, b) DSV_BIT_SET((mv)->flags, DSV_MV_BIT_SIMCMPLX, b)

#define DSV_GET_LINE(p, y) ((p)->data + (y) * (p)->stride)
#define DSV_GET_XY(p, x, y) ((p)->data + (x) + (y) * (p)->stride)

#define DSV_MAX_QP_BITS 12
#define DSV_MAX_QP ((1 << DSV_MAX_QP_BITS) - 1)

/* per-frame coding parameters (reference dsv.h:242-268) */
typedef struct {
    DSV_META *vidmeta;
    int effort;
    int do_psy;
    int is_ref;
    int has_ref;
    int blk_w, blk_h;
    int nblocks_h, nblocks_v;
    int temporal_mc;
    int lossless;
    int reserved;
} DSV_PARAMS;

typedef struct {
    uint8_t *data;
    unsigned len;
} DSV_BUF;

/* host frames / coefficient planes (reference frame.c) */
extern void dsv_mk_coefs(DSV_COEFS *c, int format, int width, int height);
extern DSV_FRAME *dsv_mk_frame(int format, int width, int height, int border);
extern DSV_FRAME *dsv_load_planar_frame(int format, void *data, int width, int height);
extern DSV_FRAME *dsv_frame_ref_inc(DSV_FRAME *frame);
extern void dsv_frame_ref_dec(DSV_FRAME *frame);
extern void dsv_frame_copy(DSV_FRAME *dst, DSV_FRAME *src);
extern DSV_FRAME *dsv_clone_frame(DSV_FRAME *f, int border);
extern void dsv_plane_xy(DSV_FRAME *f, DSV_PLANE *out, int c, int x, int y);

/* buffers, memory, raw yuv I/O, logging (reference dsv.c) */
extern void dsv_mk_buf(DSV_BUF *buf, int size);
extern void dsv_buf_free(DSV_BUF *buf);
extern int dsv_yuv_write(FILE *out, int fno, DSV_PLANE *p);
extern int dsv_yuv_write_seq(FILE *out, DSV_PLANE *p);
extern int dsv_yuv_read(FILE *in, int fno, uint8_t *o, int w, int h, int subsamp);
extern int dsv_yuv_read_seq(FILE *in, uint8_t *o, int w, int h, int subsamp);
extern void *dsv_alloc(int size);
extern void dsv_free(void *ptr);
extern void dsv_memory_report(void);

#define DSV_LEVEL_NONE 0
#define DSV_LEVEL_ERROR 1
#define DSV_LEVEL_WARNING 2
#define DSV_LEVEL_INFO 3
#define DSV_LEVEL_DEBUG 4
extern char *dsv_lvlname[DSV_LEVEL_DEBUG + 1];
extern void dsv_set_log_level(int level);
extern int dsv_get_log_level(void);

#define DSV_LOG_LVL(level, x)                                            \
    do {                                                                 \
        if ((level) <= dsv_get_log_level()) {                            \
            printf("[DSV][%s] %s(%d): ", dsv_lvlname[level], __FILE__, __LINE__); \
            printf x;                                                    \
            printf("\n");                                                \
        }                                                                \
    } while (0)
#define DSV_ERROR(x) DSV_LOG_LVL(DSV_LEVEL_ERROR, x)
#define DSV_WARNING(x) DSV_LOG_LVL(DSV_LEVEL_WARNING, x)
#define DSV_INFO(x) DSV_LOG_LVL(DSV_LEVEL_INFO, x)
#define DSV_DEBUG(x) DSV_LOG_LVL(DSV_LEVEL_DEBUG, x)
#define DSV_ASSERT(c)                          \
    do {                                       \
        if (!(c)) {                            \
            DSV_ERROR(("assert: " #c));        \
            exit(-1);                          \
        }                                      \
    } while (0)

#ifdef __cplusplus
}
#endif
#endif
