/*
 * dsv_hzcc.c -- serialisation half of the Hierarchical Zero Coefficient Coder.
 *
 * The reference fuses quantisation and entropy coding in one raster walk over
 * every coefficient (src/hzcc.c:234-448 / :450-583).  Here the arithmetic runs
 * on the GPU (csrc/k_quant.cuh) and the host only sees the ordered list of
 * non-zero symbols: (scan position, value).  Writing a plane is a walk over
 * that list (run = gap between positions); reading a plane rebuilds the list
 * in O(non-zeros) instead of O(coefficients).
 *
 * Plane layout (reference hzcc.c:585-649): [32-bit byte length][SEG dc]
 * [align][24-bit pair count][align][(UEG run, value)...][align][0x55][align],
 * value = NEG in the LL part, adaptive Rice with damping 3+level elsewhere.
 */
#include <string.h>
#include <pthread.h>
#include "dsv_host.h"
#include <stdlib.h>
#include "dsv_bits_inl.h"
#include "../csrc/hz_table.h"

/* joint (value, next run) prefix table, see csrc/hz_table.h: 40 KB, built on first use */
static uint32_t hz_tab[HZT_ROWS * HZT_SIZE];
static pthread_once_t hz_tab_once = PTHREAD_ONCE_INIT;

static void
hz_tab_build(void)
{
    int i;
    for (i = 0; i < HZT_ROWS * HZT_SIZE; i++) {
        hz_tab[i] = hzt_entry((uint32_t) i & (HZT_SIZE - 1), i >> HZT_BITS);
    }
}

void
dsv_hzcc_write_plane(DSV_BITWR *bw, const dsvcu_symbol *syms, int nsyms, int dc, int w, int h)
{
    int part[5], i, l = -1, vk = 0;
    size_t start, cnt_at;
    unsigned prev = 0;

    dsvcu_scan_layout(w, h, part);
    dsv_bw_align(bw);
    start = dsv_bw_byte(bw);
    dsv_bw_bits(bw, 32, 0);
    dsv_bw_seg(bw, dc);

    dsv_bw_align(bw);
    cnt_at = dsv_bw_byte(bw);
    dsv_bw_bits(bw, 24, 0);
    dsv_bw_align(bw);
    {
        /* the hot loop of the host side: ~45 k pairs per 1080p picture.  The bit
         * accumulator lives in registers; whole code words leave 32 bits at a time. */
        DSV_FW f;
        dsv_bw_reserve(bw, (size_t) nsyms * 24 + 1024);
        dsv_fw_begin(&f, bw);
        for (i = 0; i < nsyms; i++) {
            unsigned pos = syms[i].pos;
            int v = syms[i].v;
            if ((size_t) (f.p - bw->buf) + 96 > bw->cap) {
                dsv_fw_end(&f, bw);
                dsv_bw_reserve(bw, (size_t) (nsyms - i) * 24 + 8192);
                dsv_fw_begin(&f, bw);
            }
            while (l < 2 && pos >= (unsigned) part[l + 2]) {
                l++;
            }
            dsv_fw_ueg(&f, pos - prev);
            if (l < 0) {
                /* NEG: |v| - 1 as UEG, then the sign */
                unsigned a = v < 0 ? (unsigned) -v : (unsigned) v;
                dsv_fw_ueg(&f, a - 1);
                if (a) dsv_fw_put(&f, v < 0, 1);
            } else {
                const int damp = 3 + l;
                unsigned uv = ((unsigned) (2 * v) ^ (v < 0 ? ~0u : 0u)) - 1;
                unsigned k = (unsigned) (vk >> damp);
                unsigned q = k < 32 ? uv >> k : 0;
                if (k > 24 || q > 256) {
                    /* out of the ordinary: let the general writer deal with it */
                    dsv_fw_end(&f, bw);
                    dsv_bw_nrice(bw, v, &vk, damp);
                    dsv_fw_begin(&f, bw);
                } else {
                    if (q) {
                        vk++;
                    } else if (vk > 0) {
                        vk--;
                    }
                    if (q + 1 + k <= 32) {
                        dsv_fw_put(&f, (1u << k) | (uv & ((1u << k) - 1)), (int) (q + 1 + k));
                    } else {
                        dsv_fw_zeros(&f, q);
                        dsv_fw_put(&f, 1, 1);
                        if (k) dsv_fw_put(&f, uv & ((1u << k) - 1), (int) k);
                    }
                }
            }
            prev = pos + 1;
        }
        dsv_fw_end(&f, bw);
    }
    dsv_bw_align(bw);
    dsv_bw_patch24(bw, cnt_at, (unsigned) nsyms);

    dsv_bw_bits(bw, 8, DSV_EOP_SYMBOL);
    dsv_bw_align(bw);
    dsv_bw_patch32(bw, start, (unsigned) (dsv_bw_byte(bw) - start - 4));
}

int
dsv_hzcc_read_plane(DSV_BITRD *br, dsvcu_symbol *syms, int cap, int w, int h, int level_start[5], int *dc)
{
    int part[5], total, n = 0, l = -1, vk = 0, i;
    unsigned plen;
    size_t start, limit;
    int runs, truncated = 0;
    unsigned cur = 0, run;

    pthread_once(&hz_tab_once, hz_tab_build);
    total = dsvcu_scan_layout(w, h, part);
    for (i = 0; i < 5; i++) {
        level_start[i] = 0;
    }
    *dc = 0;
    dsv_br_align(br);
    plen = dsv_br_bits(br, 32);
    dsv_br_align(br);
    if (!(plen > 0 && plen < (unsigned) w * (unsigned) h * sizeof(DSV_SBC) * 2)) {
        DSV_ERROR(("plane length was strange: %d", (int) plen));
        return -1;
    }
    start = dsv_br_byte(br);
    limit = start + plen;
    *dc = dsv_br_seg(br);

    dsv_br_align(br);
    runs = (int) dsv_br_bits(br, 24);
    dsv_br_align(br);
    /* (run, value) pairs.  As in the reference (hzcc.c:476-486) the next run
     * is fetched before the bounds test, and a pair whose bits end at or past
     * the declared plane length is dropped together with everything after it */
    {
        DSV_FR r;
        size_t lenbits;
        r.buf = br->buf;
        r.len = br->len;
        r.pos = br->pos;
        lenbits = r.len * 8;
        size_t safe_bytes;
        int idle = 0;
        /* bytes of the buffer a burst (below) may have loaded: every code it reads then starts
         * and ends well inside both the buffer and the declared plane */
        safe_bytes = r.len < limit ? r.len : limit;
        safe_bytes = safe_bytes > 16 ? safe_bytes - 16 : 0;
        run = (runs-- > 0) ? dsv_fr_ueg(&r) : UINT_MAX;
        while (run != UINT_MAX) {
            unsigned pos;
            int v;
            /* Burst: while the pairs stay inside one part of the scan, inside the table
             * (both codes within HZT_BITS bits, Rice parameter below HZT_KMAX) and away from
             * the end of the data, they are read from a 64-bit window held in a register --
             * one table look-up and one shift per pair, no memory access on the dependent
             * chain except the look-up.  Whatever ends the burst is handled by the general
             * code below, one pair at a time, exactly as before. */
            if (idle > 0) {
                idle--;
            } else if (runs > 0 && (r.pos >> 3) + 8 <= safe_bytes) {
                const unsigned bnd = (unsigned) (l < 2 ? part[l + 2] : total);
                const uint8_t *next = r.buf + (r.pos >> 3) + 8;
                const uint8_t *const stop = r.buf + safe_bytes;
                uint64_t win;
                int nbits = 64 - (int) (r.pos & 7), got = 0;
                {
                    uint64_t w8;
                    memcpy(&w8, next - 8, 8);
#if !(defined(__BYTE_ORDER__) && (__BYTE_ORDER__ == __ORDER_BIG_ENDIAN__))
                    w8 = __builtin_bswap64(w8);
#endif
                    win = w8 << (r.pos & 7);
                }
                for (;;) {
                    unsigned row;
                    uint32_t e;
                    pos = cur + run;
                    if (pos >= bnd || pos < cur || runs <= 0 || next > stop) {
                        break;
                    }
                    if (nbits < 32) {
                        uint32_t w4;
                        memcpy(&w4, next, 4);
#if !(defined(__BYTE_ORDER__) && (__BYTE_ORDER__ == __ORDER_BIG_ENDIAN__))
                        w4 = __builtin_bswap32(w4);
#endif
                        win |= (uint64_t) w4 << (32 - nbits);
                        next += 4;
                        nbits += 32;
                    }
                    row = l < 0 ? (unsigned) HZT_ROW_LL : (unsigned) (vk >> (3 + l));
                    if (row >= HZT_ROWS || (l >= 0 && row >= HZT_KMAX)) {
                        break;
                    }
                    e = hz_tab[row * HZT_SIZE + (uint32_t) (win >> (64 - HZT_BITS))];
                    if (!HZT_LEN(e)) {
                        break;
                    }
                    win <<= HZT_LEN(e);
                    nbits -= HZT_LEN(e);
                    runs--;
                    if (l >= 0) {
                        vk += HZT_QNZ(e) ? 1 : -(vk > 0);
                    }
                    if (n < cap && pos != 0) {
                        syms[n].pos = pos;
                        syms[n].v = HZT_VAL(e);
                        n++;
                    }
                    cur = pos + 1;
                    run = HZT_RUN(e);
                    got++;
                }
                r.pos = (size_t) (next - r.buf) * 8 - (size_t) nbits;
                if (got < 4) {
                    idle = 32; /* not table country (large values): do not keep setting up windows */
                }
            }
            pos = cur + run;
            if (pos >= (unsigned) total || pos < cur) {
                break;
            }
            while (l < 2 && pos >= (unsigned) part[l + 2]) {
                l++;
                level_start[l + 1] = n;
            }
            /* value of this pair + run of the next: one look-up where both codes fit into
             * HZT_BITS bits, the general readers otherwise (same bits, same result) */
            {
                const int row = l < 0 ? HZT_ROW_LL : vk >> (3 + l);
                uint32_t e = 0;
                /* (the general readers see zeros once a code STARTS behind the end of the buffer:
                 * the table is only asked while both codes start inside it) */
                if (runs > 0 && (l < 0 || row < HZT_KMAX) && r.pos + HZT_BITS < lenbits) {
                    e = hz_tab[row * HZT_SIZE + (uint32_t) (dsv_fr_peek(&r) >> (64 - HZT_BITS))];
                }
                if (HZT_LEN(e)) {
                    r.pos += (size_t) HZT_LEN(e);
                    v = HZT_VAL(e);
                    run = HZT_RUN(e);
                    runs--;
                    if (l >= 0) {
                        if (HZT_QNZ(e)) {
                            vk++;
                        } else if (vk > 0) {
                            vk--;
                        }
                    }
                } else {
                    v = (l < 0) ? dsv_fr_neg(&r) : dsv_fr_nrice(&r, &vk, 3 + l);
                    run = (runs-- > 0) ? dsv_fr_ueg(&r) : UINT_MAX;
                }
            }
            if ((r.pos >> 3) >= limit) {
                truncated = 1;
                break;
            }
            if (n < cap && pos != 0) {
                syms[n].pos = pos;
                syms[n].v = v;
                n++;
            }
            cur = pos + 1;
        }
        br->pos = r.pos;
    }
    while (l < 2) {
        l++;
        level_start[l + 1] = n;
    }
    level_start[4] = n;

    /* end-of-plane marker: checked where sequential parsing stopped */
    if (!truncated) {
        dsv_br_align(br);
    }
    if (dsv_br_bits(br, 8) != DSV_EOP_SYMBOL) {
        DSV_ERROR(("bad eop, frame data incomplete and/or corrupt"));
        br->pos = limit * 8;
        return -1;
    }
    br->pos = limit * 8;
    return n;
}

/* convenience for callers that want one plane as a byte string (tests, tools):
 * returns the byte length, or -1 when `cap` is too small */
int
dsv_hzcc_pack_plane(const dsvcu_symbol *syms, int nsyms, int dc, int w, int h, uint8_t *out, int cap)
{
    DSV_BITWR bw;
    int n;
    dsv_bw_init(&bw, (size_t) nsyms * 4 + 64);
    dsv_hzcc_write_plane(&bw, syms, nsyms, dc, w, h);
    n = (int) dsv_bw_byte(&bw);
    if (n > cap) {
        n = -1;
    } else {
        memcpy(out, bw.buf, (size_t) n);
    }
    dsv_bw_free(&bw);
    return n;
}

/* test hook, mirror of dsv_hzcc_pack_plane: parses one serialised plane (as the
 * reference's dsv_encode_plane wrote it) into the ordered symbol list */
int
dsv_hzcc_unpack_plane(const uint8_t *bits, int len, dsvcu_symbol *syms, int cap, int w, int h, int level_start[5], int *dc)
{
    DSV_BITRD br;
    uint8_t *copy = calloc((size_t) len + 32, 1);
    int n;
    if (!copy) {
        return -1;
    }
    memcpy(copy, bits, (size_t) len);
    dsv_br_init(&br, copy, (size_t) len);
    n = dsv_hzcc_read_plane(&br, syms, cap, w, h, level_start, dc);
    free(copy);
    return n;
}
