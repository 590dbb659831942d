/*
 * dsv_bits.c -- MSB-first bit reader / writer and the DSV2 variable-length
 * codes: interleaved exp-Golomb (UEG / SEG / NEG), adaptive Rice (URC / NRC)
 * and zero-bit run-length coding.
 *
 * Byte format is frozen by the bitstream (reference src/bs.c:17-330, spec
 * section B).  The implementation is new: the reader works on a 64-bit
 * big-endian window with count-leading-zeros instead of bit-at-a-time reads,
 * the writer ORs whole code words into a zeroed growable buffer.
 */
#include <stdlib.h>
#include <string.h>
#include "dsv_host.h"

static uint64_t
load_be64(const uint8_t *p)
{
    uint64_t v;
    memcpy(&v, p, 8);
#if defined(__BYTE_ORDER__) && (__BYTE_ORDER__ == __ORDER_BIG_ENDIAN__)
    return v;
#else
    return __builtin_bswap64(v);
#endif
}

static void
store_be64(uint8_t *p, uint64_t v)
{
#if !(defined(__BYTE_ORDER__) && (__BYTE_ORDER__ == __ORDER_BIG_ENDIAN__))
    v = __builtin_bswap64(v);
#endif
    memcpy(p, &v, 8);
}

/* ------------------------------------------------------------------ reader */

void
dsv_br_init(DSV_BITRD *br, const uint8_t *buf, size_t len)
{
    br->buf = buf;
    br->pos = 0;
    br->len = len;
}

void
dsv_br_align(DSV_BITRD *br)
{
    br->pos = (br->pos + 7) & ~(size_t) 7;
}

/* next >= 57 valid bits, left-aligned */
static uint64_t
br_peek(const DSV_BITRD *br)
{
    size_t byte = br->pos >> 3;
    if (byte >= br->len) {
        return 0; /* past the end: behave like zero padding */
    }
    return load_be64(br->buf + byte) << (br->pos & 7);
}

unsigned
dsv_br_bit(DSV_BITRD *br)
{
    unsigned b = (unsigned) (br_peek(br) >> 63);
    br->pos++;
    return b;
}

unsigned
dsv_br_bits(DSV_BITRD *br, unsigned n)
{
    uint64_t w;
    if (n == 0) {
        return 0;
    }
    w = br_peek(br);
    br->pos += n;
    return (unsigned) (w >> (64 - n));
}

/* UEG: pairs (0, data bit) ... terminated by a single 1 */
unsigned
dsv_br_ueg(DSV_BITRD *br)
{
    unsigned v = 1;
    for (;;) {
        uint64_t w = br_peek(br);
        /* stop bits sit at even offsets from the window start */
        uint64_t stops = w & 0xAAAAAAAAAAAAAA00ULL;
        int pairs, i;
        if (stops == 0) {
            /* 28 full pairs without a terminator */
            for (i = 0; i < 28; i++) {
                v = (v << 1) | (unsigned) ((w >> 62) & 1);
                w <<= 2;
            }
            br->pos += 56;
            if ((br->pos >> 3) >= br->len) {
                return v - 1;
            }
            continue;
        }
        pairs = __builtin_clzll(stops) >> 1;
        for (i = 0; i < pairs; i++) {
            v = (v << 1) | (unsigned) ((w >> 62) & 1);
            w <<= 2;
        }
        br->pos += (size_t) (2 * pairs + 1);
        return v - 1;
    }
}

int
dsv_br_seg(DSV_BITRD *br)
{
    int v = (int) dsv_br_ueg(br);
    if (v && dsv_br_bit(br)) {
        return -v;
    }
    return v;
}

int
dsv_br_neg(DSV_BITRD *br)
{
    int v = (int) dsv_br_ueg(br) + 1;
    if (v && dsv_br_bit(br)) {
        return -v;
    }
    return v;
}

int
dsv_br_nrice(DSV_BITRD *br, int *rk, int damp)
{
    int k = (*rk) >> damp;
    unsigned q = 0, uv;
    for (;;) {
        uint64_t w = br_peek(br);
        if ((w >> 8) == 0) {
            q += 56;
            br->pos += 56;
            if ((br->pos >> 3) >= br->len) {
                break;
            }
            continue;
        }
        {
            int n = __builtin_clzll(w);
            q += (unsigned) n;
            br->pos += (size_t) n + 1;
        }
        break;
    }
    if (q) {
        (*rk)++;
    } else if (*rk > 0) {
        (*rk)--;
    }
    uv = (q << k) | dsv_br_bits(br, (unsigned) k);
    uv += 1;
    return (int) (uv >> 1) ^ -(int) (uv & 1);
}

void
dsv_rle_rd_init(DSV_RLERD *r, const uint8_t *buf, size_t len)
{
    dsv_br_init(&r->br, buf, len);
    r->nz = 0;
}

int
dsv_rle_rd_get(DSV_RLERD *r)
{
    if (r->nz == 0) {
        r->nz = (int) dsv_br_ueg(&r->br);
        return r->nz == 0;
    }
    r->nz--;
    return r->nz == 0;
}

void
dsv_rle_rd_end(DSV_RLERD *r)
{
    if (r->nz > 1) {
        DSV_ERROR(("%d remaining in run", r->nz));
    }
}

/* ------------------------------------------------------------------ writer */

static void
bw_reserve(DSV_BITWR *bw, size_t bits_more)
{
    size_t need = ((bw->pos + bits_more) >> 3) + 24;
    if (need > bw->cap) {
        size_t ncap = bw->cap * 2;
        if (ncap < need) {
            ncap = need * 2;
        }
        bw->buf = realloc(bw->buf, ncap);
        memset(bw->buf + bw->cap, 0, ncap - bw->cap);
        bw->cap = ncap;
    }
}

/* for the register-resident writer of dsv_bits_inl.h */
void
dsv_bw_reserve(DSV_BITWR *bw, size_t bits_more)
{
    bw_reserve(bw, bits_more);
}

void
dsv_bw_init(DSV_BITWR *bw, size_t initial_bytes)
{
    if (initial_bytes < 64) {
        initial_bytes = 64;
    }
    bw->buf = calloc(1, initial_bytes);
    bw->cap = initial_bytes;
    bw->pos = 0;
}

void
dsv_bw_free(DSV_BITWR *bw)
{
    free(bw->buf);
    bw->buf = NULL;
    bw->cap = bw->pos = 0;
}

void
dsv_bw_align(DSV_BITWR *bw)
{
    bw->pos = (bw->pos + 7) & ~(size_t) 7;
    bw_reserve(bw, 0);
}

/* OR the low n (<= 32) bits of v at the current position */
static void
bw_or(DSV_BITWR *bw, unsigned n, unsigned v)
{
    uint8_t *p = bw->buf + (bw->pos >> 3);
    unsigned sh = (unsigned) (bw->pos & 7);
    uint64_t cur = load_be64(p);
    uint64_t val = (n == 32) ? (uint64_t) v : ((uint64_t) v & (((uint64_t) 1 << n) - 1));
    cur |= val << (64 - sh - n);
    store_be64(p, cur);
    bw->pos += n;
}

void
dsv_bw_bit(DSV_BITWR *bw, int v)
{
    bw_reserve(bw, 1);
    bw_or(bw, 1, v ? 1u : 0u);
}

void
dsv_bw_bits(DSV_BITWR *bw, unsigned n, unsigned v)
{
    if (n == 0) {
        return;
    }
    bw_reserve(bw, n);
    bw_or(bw, n, v);
}

void
dsv_bw_ueg(DSV_BITWR *bw, unsigned v)
{
    int nb, i;
    uint64_t code = 0;
    v++;
    nb = 31 - __builtin_clz(v); /* floor(log2(v)) */
    bw_reserve(bw, 2 * 32 + 1);
    if (nb <= 15) {
        /* interleave: 0 b(nb-1) 0 b(nb-2) ... 0 b0 1  -> 2*nb+1 <= 31 bits */
        for (i = nb - 1; i >= 0; i--) {
            code = (code << 2) | ((v >> i) & 1);
        }
        code = (code << 1) | 1;
        bw_or(bw, (unsigned) (2 * nb + 1), (unsigned) code);
        return;
    }
    for (i = nb - 1; i >= 0; i--) {
        bw_or(bw, 2, (v >> i) & 1);
    }
    bw_or(bw, 1, 1);
}

void
dsv_bw_seg(DSV_BITWR *bw, int v)
{
    int s = v < 0;
    unsigned a = s ? (unsigned) -v : (unsigned) v;
    dsv_bw_ueg(bw, a);
    if (a) {
        dsv_bw_bit(bw, s);
    }
}

void
dsv_bw_neg(DSV_BITWR *bw, int v)
{
    int s = v < 0;
    unsigned a = s ? (unsigned) -v : (unsigned) v;
    dsv_bw_ueg(bw, a - 1);
    if (a) {
        dsv_bw_bit(bw, s);
    }
}

void
dsv_bw_nrice(DSV_BITWR *bw, int v, int *rk, int damp)
{
    unsigned uv = ((unsigned) (2 * v) ^ (v < 0 ? ~0u : 0u)) - 1;
    unsigned k = (unsigned) ((*rk) >> damp);
    unsigned q = uv >> k;
    if (q) {
        (*rk)++;
    } else if (*rk > 0) {
        (*rk)--;
    }
    bw_reserve(bw, (size_t) q + 1 + 32);
    bw->pos += q; /* q zero bits: the buffer is already clear */
    bw_or(bw, 1, 1);
    if (k) {
        bw_or(bw, k, uv);
    }
}

void
dsv_bw_bytes(DSV_BITWR *bw, const uint8_t *data, size_t n)
{
    if (bw->pos & 7) {
        DSV_ERROR(("append to unaligned bit writer"));
    }
    if (n == 0) {
        return;
    }
    bw_reserve(bw, n * 8);
    memcpy(bw->buf + (bw->pos >> 3), data, n);
    bw->pos += n * 8;
}

void
dsv_bw_patch32(DSV_BITWR *bw, size_t off, unsigned v)
{
    bw->buf[off + 0] = (uint8_t) (v >> 24);
    bw->buf[off + 1] = (uint8_t) (v >> 16);
    bw->buf[off + 2] = (uint8_t) (v >> 8);
    bw->buf[off + 3] = (uint8_t) v;
}

void
dsv_bw_patch24(DSV_BITWR *bw, size_t off, unsigned v)
{
    bw->buf[off + 0] = (uint8_t) (v >> 16);
    bw->buf[off + 1] = (uint8_t) (v >> 8);
    bw->buf[off + 2] = (uint8_t) v;
}

void
dsv_rle_wr_init(DSV_RLEWR *r, size_t initial_bytes)
{
    dsv_bw_init(&r->bw, initial_bytes);
    r->nz = 0;
}

void
dsv_rle_wr_put(DSV_RLEWR *r, int b)
{
    if (b) {
        dsv_bw_ueg(&r->bw, (unsigned) r->nz);
        r->nz = 0;
        return;
    }
    r->nz++;
}

size_t
dsv_rle_wr_end(DSV_RLEWR *r)
{
    dsv_bw_ueg(&r->bw, (unsigned) r->nz);
    r->nz = 0;
    dsv_bw_align(&r->bw);
    return dsv_bw_byte(&r->bw);
}
