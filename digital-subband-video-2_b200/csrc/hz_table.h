/*
 * hz_table.h -- joint prefix table for the (value, next run) pairs of a coefficient
 * plane, shared by the device parser (k_hzcc.cuh) and the host parser (host/dsv_hzcc.c).
 *
 * The serialised plane (reference src/hzcc.c:308-439 writes it, :450-583 reads it) is a
 * chain  run value run value ...: runs in interleaved exp-Golomb (src/bs.c:96-137), values
 * of the LL part as exp-Golomb magnitude + sign, values of the three levels in adaptive Rice
 * with parameter k (src/bs.c:237-251).  The reader always knows, before it looks at a value,
 * which code the value is in (the scan position is known) and which code follows it (a run).
 * So the next HZT_BITS bits of the stream, together with k, decide value, run and the length
 * of both: one table row per k < HZT_KMAX plus one row for the LL part.
 *
 * entry: bits 0-3 length of both codes (0: they do not fit into HZT_BITS bits, use the
 *        general reader), bits 4-9 the run, bit 10 "quotient of the Rice code was not zero"
 *        (drives the adaptation of k), bits 16-31 the value (signed).
 */
#ifndef HZ_TABLE_H
#define HZ_TABLE_H

#include <stdint.h>

#ifndef HZT_FN
#define HZT_FN static inline
#endif
#ifndef HZT_CLZ32
#define HZT_CLZ32(v) ((v) ? __builtin_clz(v) : 32)
#endif

#define HZT_BITS 11
#define HZT_SIZE (1 << HZT_BITS)
#define HZT_KMAX 4           /* rows 0 .. HZT_KMAX-1: Rice parameter k */
#define HZT_ROW_LL HZT_KMAX  /* row of the LL part */
#define HZT_ROWS (HZT_KMAX + 1)

#define HZT_LEN(e) ((int) ((e) & 15u))
#define HZT_RUN(e) (((e) >> 4) & 63u)
#define HZT_QNZ(e) (((e) >> 10) & 1u)
#define HZT_VAL(e) ((int) (int32_t) (e) >> 16)

/* even bits 0, 2, .. 30 of x packed into the low 16 bits */
HZT_FN uint32_t
hzt_even_bits(uint32_t x)
{
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    x = (x | (x >> 8)) & 0x0000ffffu;
    return x;
}

/* interleaved exp-Golomb code at the top of `top`, of which `avail` bits are real:
 * returns its length (0: not complete within avail) and the value */
HZT_FN int
hzt_ueg(uint32_t top, int avail, uint32_t *val)
{
    const uint32_t stops = top & 0xAAAAAAAAu;
    int pairs;
    if (!stops) return 0;
    pairs = HZT_CLZ32(stops) >> 1;
    if (2 * pairs + 1 > avail) return 0;
    *val = pairs ? ((1u << pairs) | hzt_even_bits(top >> (32 - 2 * pairs))) - 1u : 0u;
    return 2 * pairs + 1;
}

HZT_FN uint32_t
hzt_entry(uint32_t prefix, int row)
{
    uint32_t top = prefix << (32 - HZT_BITS), run = 0, qnz = 0;
    int avail = HZT_BITS, len, n;
    int v;
    if (row < HZT_KMAX) {
        /* Rice: z zeros, a one, k remainder bits; the value is the zig-zag of uv + 1 */
        const int k = row;
        uint32_t uv, rest;
        int z;
        if (!top) return 0;
        z = HZT_CLZ32(top);
        if (z + 1 + k > avail) return 0;
        rest = top << z << 1;
        uv = ((uint32_t) z << k) | ((rest >> 1) >> (31 - k));
        qnz = z != 0;
        uv += 1;
        v = (int) (uv >> 1) ^ -(int) (uv & 1);
        len = z + 1 + k;
    } else {
        /* LL: magnitude - 1 in exp-Golomb, then the sign */
        uint32_t m;
        n = hzt_ueg(top, avail, &m);
        if (!n || n + 1 > avail) return 0;
        v = (int) m + 1;
        if ((top << n) & 0x80000000u) v = -v;
        len = n + 1;
    }
    top <<= len;
    avail -= len;
    n = hzt_ueg(top, avail, &run);
    if (!n) return 0;
    len += n;
    return (uint32_t) len | (run << 4) | (qnz << 10) | ((uint32_t) (v & 0xffff) << 16);
}

#endif /* HZ_TABLE_H */
