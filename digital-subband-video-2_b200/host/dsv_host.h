/*
 * dsv_host.h -- internal declarations of the host layer (not installed).
 *
 * The host layer is C89-style C: bitstream packing/parsing, the HZCC
 * coefficient coder's serialisation half, packet framing, encoder control.  It
 * talks to the GPU only through include/dsv_cuda.h.
 */
#ifndef DSV2_B200_HOST_H
#define DSV2_B200_HOST_H

#include <limits.h>
#include "../../include/dsv.h"
#include "../../include/dsv_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* blockdata bits (reference dsv_internal.h:96-110) */
#define DSV_STABLE_BIT 0
#define DSV_MAINTAIN_BIT 1
#define DSV_SKIP_BIT 2
#define DSV_RINGING_BIT 3
#define DSV_INTRA_BIT 4
#define DSV_EPRM_BIT 5
#define DSV_SIMCMPLX_BIT 6
#define DSV_IS_STABLE (1 << DSV_STABLE_BIT)
#define DSV_IS_MAINTAIN (1 << DSV_MAINTAIN_BIT)
#define DSV_IS_SKIP (1 << DSV_SKIP_BIT)
#define DSV_IS_RINGING (1 << DSV_RINGING_BIT)
#define DSV_IS_INTRA (1 << DSV_INTRA_BIT)
#define DSV_IS_EPRM (1 << DSV_EPRM_BIT)
#define DSV_IS_SIMCMPLX (1 << DSV_SIMCMPLX_BIT)

/* motion sub-streams (reference dsv_internal.h:30-36) */
#define DSV_SUB_MODE 0
#define DSV_SUB_MV_X 1
#define DSV_SUB_MV_Y 2
#define DSV_SUB_SBIM 3
#define DSV_SUB_EPRM 4
#define DSV_SUB_NSUB 5

#define DSV_FRAME_BORDER DSV_MAX_BLOCK_SIZE
#define DSV_EOP_SYMBOL 0x55

/* ---- bit reader: MSB-first, 64-bit window; the buffer must be readable for 8
 * bytes past `len` (callers pad) ---- */
typedef struct {
    const uint8_t *buf;
    size_t pos; /* in bits */
    size_t len; /* in bytes */
} DSV_BITRD;

void dsv_br_init(DSV_BITRD *br, const uint8_t *buf, size_t len);
void dsv_br_align(DSV_BITRD *br);
unsigned dsv_br_bit(DSV_BITRD *br);
unsigned dsv_br_bits(DSV_BITRD *br, unsigned n); /* n <= 32 */
unsigned dsv_br_ueg(DSV_BITRD *br);
int dsv_br_seg(DSV_BITRD *br);
int dsv_br_neg(DSV_BITRD *br);
int dsv_br_nrice(DSV_BITRD *br, int *rk, int damp);
#define dsv_br_byte(br) ((br)->pos >> 3)

/* zero-bit run-length reader (reference bs.c:277-330) */
typedef struct {
    DSV_BITRD br;
    int nz;
} DSV_RLERD;
void dsv_rle_rd_init(DSV_RLERD *r, const uint8_t *buf, size_t len);
int dsv_rle_rd_get(DSV_RLERD *r);
void dsv_rle_rd_end(DSV_RLERD *r);

/* ---- bit writer: appends to a growable byte buffer ---- */
typedef struct {
    uint8_t *buf;
    size_t cap;  /* bytes */
    size_t pos;  /* bits */
} DSV_BITWR;

void dsv_bw_init(DSV_BITWR *bw, size_t initial_bytes);
void dsv_bw_free(DSV_BITWR *bw);
void dsv_bw_align(DSV_BITWR *bw);
void dsv_bw_bit(DSV_BITWR *bw, int v);
void dsv_bw_bits(DSV_BITWR *bw, unsigned n, unsigned v);
void dsv_bw_ueg(DSV_BITWR *bw, unsigned v);
void dsv_bw_seg(DSV_BITWR *bw, int v);
void dsv_bw_neg(DSV_BITWR *bw, int v);
void dsv_bw_nrice(DSV_BITWR *bw, int v, int *rk, int damp);
void dsv_bw_bytes(DSV_BITWR *bw, const uint8_t *data, size_t n); /* aligned append */
void dsv_bw_patch32(DSV_BITWR *bw, size_t byte_off, unsigned v);
void dsv_bw_patch24(DSV_BITWR *bw, size_t byte_off, unsigned v);
#define dsv_bw_byte(bw) ((bw)->pos >> 3)

typedef struct {
    DSV_BITWR bw;
    int nz;
} DSV_RLEWR;
void dsv_rle_wr_init(DSV_RLEWR *r, size_t initial_bytes);
void dsv_rle_wr_put(DSV_RLEWR *r, int b);
size_t dsv_rle_wr_end(DSV_RLEWR *r); /* returns byte length */

/* ---- HZCC plane serialisation (entropy half of reference hzcc.c) ---- */
/* writes one plane: [32-bit length][SEG dc][24-bit count][(run,value)...][0x55] */
void dsv_hzcc_write_plane(DSV_BITWR *bw, const dsvcu_symbol *syms, int nsyms, int dc, int w, int h);
int dsv_hzcc_pack_plane(const dsvcu_symbol *syms, int nsyms, int dc, int w, int h, uint8_t *out, int cap);
int dsv_hzcc_unpack_plane(const uint8_t *bits, int len, dsvcu_symbol *syms, int cap, int w, int h, int level_start[5],
                          int *dc);
/* parses one plane into `syms` (capacity `cap`); fills level_start[5] and *dc.
 * returns the symbol count, or -1 when the plane is corrupt (bad length / EOP) */
int dsv_hzcc_read_plane(DSV_BITRD *br, dsvcu_symbol *syms, int cap, int w, int h, int level_start[5], int *dc);

/* ---- MV helpers (reference dsv.c:324-447) ---- */
void dsv_movec_pred(DSV_MV *vecs, DSV_PARAMS *p, int x, int y, int *px, int *py);
void dsv_neighbordif2(DSV_MV *vecs, DSV_PARAMS *p, int x, int y, int *dx, int *dy);
int dsv_neighbordif(DSV_MV *vecs, DSV_PARAMS *p, int x, int y);
int dsv_mv_cost(DSV_MV *vecs, DSV_PARAMS *p, int i, int j, int mx, int my, int q, int sqr);
int dsv_lb2(unsigned n);
int dsv_spatial_psy_factor(DSV_PARAMS *p, int subband);

/* ---- misc host helpers ---- */
void dsv_host_copy_planes(DSV_FRAME *dst, DSV_FRAME *src);
int dsv_y4m_read_hdr(FILE *in, int *w, int *h, int *subsamp, int *fpsn, int *fpsd, int *aspn, int *aspd);
int dsv_y4m_read_frame(FILE *in, uint8_t *o, int w, int h, int subsamp);
void dsv_y4m_write_hdr(FILE *out, int w, int h, int subsamp, int fpsn, int fpsd, int aspn, int aspd);
void dsv_y4m_write_frame_hdr(FILE *out);

void dsv_fmeta_from_params(dsvcu_fmeta *fm, const DSV_PARAMS *p, int isP, unsigned fnum);

#ifdef __cplusplus
}
#endif
#endif
