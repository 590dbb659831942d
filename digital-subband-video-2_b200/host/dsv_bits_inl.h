/*
 * dsv_bits_inl.h -- register-resident bit writer / reader for the one loop
 * that matters on the host: the (run, value) pairs of a coefficient plane
 * (dsv_hzcc.c).  Same codes and byte format as dsv_bits.c (reference
 * src/bs.c:17-330); the state lives in a local struct so the compiler keeps it
 * in registers, code words are built whole (bit-spread table for the
 * interleaved exp-Golomb code) and leave 32 bits at a time.
 */
#ifndef DSV_BITS_INL_H
#define DSV_BITS_INL_H

#include <stdint.h>
#include <string.h>
#include "dsv_host.h"

void dsv_bw_reserve(DSV_BITWR *bw, size_t bits_more); /* dsv_bits.c */

/* ------------------------------------------------------------------ writer */

typedef struct {
    uint8_t *p;   /* next byte to write */
    uint64_t acc; /* pending bits in the low `n` bits */
    int n;        /* < 32 between calls */
} DSV_FW;

/* bit i of the argument moved to bit 2i */
static inline uint32_t
dsv_spread16(uint32_t x)
{
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

static inline void
dsv_fw_begin(DSV_FW *f, DSV_BITWR *bw)
{
    int part = (int) (bw->pos & 7);
    f->p = bw->buf + (bw->pos >> 3);
    f->n = part;
    f->acc = part ? (uint64_t) (f->p[0] >> (8 - part)) : 0;
}

/* publish the position; the partial byte is written left-aligned (the buffer
 * behind it is still zero, as the OR-ing writer of dsv_bits.c expects) */
static inline void
dsv_fw_end(DSV_FW *f, DSV_BITWR *bw)
{
    int n = f->n;
    uint64_t acc = f->acc;
    uint8_t *p = f->p;
    while (n >= 8) {
        *p++ = (uint8_t) (acc >> (n - 8));
        n -= 8;
    }
    if (n) {
        *p = (uint8_t) (acc << (8 - n));
    }
    bw->pos = (size_t) (p - bw->buf) * 8 + (size_t) n;
}

/* append the low `len` (1..32) bits of v */
static inline void
dsv_fw_put(DSV_FW *f, uint32_t v, int len)
{
    f->acc = (f->acc << len) | (uint64_t) v;
    f->n += len;
    if (f->n >= 32) {
        uint32_t w = (uint32_t) (f->acc >> (f->n - 32));
        f->p[0] = (uint8_t) (w >> 24);
        f->p[1] = (uint8_t) (w >> 16);
        f->p[2] = (uint8_t) (w >> 8);
        f->p[3] = (uint8_t) w;
        f->p += 4;
        f->n -= 32;
    }
}

/* interleaved exp-Golomb: 0 b(nb-1) 0 b(nb-2) ... 0 b0 1 for v+1 = 1 b(nb-1)..b0 */
static inline void
dsv_fw_ueg(DSV_FW *f, uint32_t v)
{
    int nb;
    v++;
    nb = 31 - __builtin_clz(v);
    if (nb <= 15) {
        dsv_fw_put(f, (dsv_spread16(v & ((1u << nb) - 1)) << 1) | 1u, 2 * nb + 1);
    } else {
        /* high pairs first, then the low 15 pairs and the terminator */
        int hi = nb - 15;
        dsv_fw_put(f, dsv_spread16((v >> 15) & ((1u << hi) - 1)), 2 * hi);
        dsv_fw_put(f, (dsv_spread16(v & 0x7fffu) << 1) | 1u, 31);
    }
}

static inline void
dsv_fw_zeros(DSV_FW *f, unsigned q)
{
    while (q >= 32) {
        dsv_fw_put(f, 0, 32);
        q -= 32;
    }
    if (q) dsv_fw_put(f, 0, (int) q);
}

/* ------------------------------------------------------------------ reader */

typedef struct {
    const uint8_t *buf;
    size_t len;
    size_t pos; /* bits */
} DSV_FR;

static inline uint64_t
dsv_fr_peek(const DSV_FR *r)
{
    size_t byte = r->pos >> 3;
    uint64_t v;
    if (byte >= r->len) {
        return 0; /* past the end: behave like zero padding */
    }
    memcpy(&v, r->buf + byte, 8);
#if !(defined(__BYTE_ORDER__) && (__BYTE_ORDER__ == __ORDER_BIG_ENDIAN__))
    v = __builtin_bswap64(v);
#endif
    return v << (r->pos & 7);
}

static inline unsigned
dsv_fr_ueg(DSV_FR *r)
{
    unsigned v = 1;
    for (;;) {
        uint64_t w = dsv_fr_peek(r);
        uint64_t stops = w & 0xAAAAAAAAAAAAAA00ULL;
        int pairs, i;
        if (stops == 0) {
            for (i = 0; i < 28; i++) {
                v = (v << 1) | (unsigned) ((w >> 62) & 1);
                w <<= 2;
            }
            r->pos += 56;
            if ((r->pos >> 3) >= r->len) {
                return v - 1;
            }
            continue;
        }
        pairs = __builtin_clzll(stops) >> 1;
        if (pairs <= 16) {
            /* compress the data bits (even bit positions after the shift) of the leading pairs */
            uint64_t x = pairs ? (w >> (64 - 2 * pairs)) : 0; /* 2*pairs bits: 0 d 0 d ... */
            x &= 0x5555555555555555ULL;
            x = (x | (x >> 1)) & 0x3333333333333333ULL;
            x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0fULL;
            x = (x | (x >> 4)) & 0x00ff00ff00ff00ffULL;
            x = (x | (x >> 8)) & 0x0000ffff0000ffffULL;
            v = (v << pairs) | (unsigned) (x & 0xffff);
        } else {
            for (i = 0; i < pairs; i++) {
                v = (v << 1) | (unsigned) ((w >> 62) & 1);
                w <<= 2;
            }
        }
        r->pos += (size_t) (2 * pairs + 1);
        return v - 1;
    }
}

static inline unsigned
dsv_fr_bits(DSV_FR *r, unsigned n)
{
    uint64_t w;
    if (n == 0) {
        return 0;
    }
    w = dsv_fr_peek(r);
    r->pos += n;
    return (unsigned) (w >> (64 - n));
}

static inline int
dsv_fr_neg(DSV_FR *r)
{
    int v = (int) dsv_fr_ueg(r) + 1;
    if (v && dsv_fr_bits(r, 1)) {
        return -v;
    }
    return v;
}

static inline int
dsv_fr_nrice(DSV_FR *r, int *rk, int damp)
{
    int k = (*rk) >> damp;
    unsigned q = 0, uv;
    for (;;) {
        uint64_t w = dsv_fr_peek(r);
        if ((w >> 8) == 0) {
            q += 56;
            r->pos += 56;
            if ((r->pos >> 3) >= r->len) {
                break;
            }
            continue;
        }
        {
            int n = __builtin_clzll(w);
            q += (unsigned) n;
            r->pos += (size_t) n + 1;
            /* the remainder bits are usually inside the same window */
            if (n + 1 + k <= 56) {
                uv = (q << k) | (k ? (unsigned) ((w << (n + 1)) >> (64 - k)) : 0u);
                r->pos += (size_t) k;
                goto done;
            }
        }
        break;
    }
    uv = (q << k) | dsv_fr_bits(r, (unsigned) k);
done:
    if (q) {
        (*rk)++;
    } else if (*rk > 0) {
        (*rk)--;
    }
    uv += 1;
    return (int) (uv >> 1) ^ -(int) (uv & 1);
}

static inline unsigned
dsv_fr_bit(DSV_FR *r)
{
    unsigned b = (unsigned) (dsv_fr_peek(r) >> 63);
    r->pos++;
    return b;
}

static inline int
dsv_fr_seg(DSV_FR *r)
{
    int v = (int) dsv_fr_ueg(r);
    if (v && dsv_fr_bit(r)) {
        return -v;
    }
    return v;
}

/* zero-bit run-length reader (dsv_rle_rd_get of dsv_bits.c, reference bs.c:277-330) */
typedef struct {
    DSV_FR r;
    int nz;
} DSV_FRLE;

static inline void
dsv_frle_init(DSV_FRLE *e, const uint8_t *buf, size_t len)
{
    e->r.buf = buf;
    e->r.len = len;
    e->r.pos = 0;
    e->nz = 0;
}

static inline int
dsv_frle_get(DSV_FRLE *e)
{
    if (e->nz == 0) {
        e->nz = (int) dsv_fr_ueg(&e->r);
        return e->nz == 0;
    }
    e->nz--;
    return e->nz == 0;
}

#endif /* DSV_BITS_INL_H */
