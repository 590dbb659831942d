/*
 * Differential test of the plane (run, value) writer / reader of host/dsv_hzcc.c
 * (register-resident bit I/O, dsv_bits_inl.h) against the same loops written
 * with the general bit writer / reader of host/dsv_bits.c, on random symbol
 * lists: dense, sparse, long runs, large values, unaligned starts, truncated
 * and bit-damaged input.  usage: hzcc_fuzz [planes]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include "dsv_host.h"
int dsvcu_scan_layout(int w, int h, int part_start[5])
{
    int l, pos;
#define RU(v, s) (((v) + (1 << (s)) - 1) >> (s))
    pos = RU(w, 3) * RU(h, 3);
    part_start[0] = 0;
    for (l = 0; l < 3; l++) { part_start[1 + l] = pos; pos += 3 * RU(w, 3 - l) * RU(h, 3 - l); }
    part_start[4] = pos;
    return pos;
}
/* the previous (generic, bit-writer based) loops */
static void old_write(DSV_BITWR *bw, const dsvcu_symbol *syms, int nsyms, int dc, int w, int h)
{
    int part[5], i, l = -1, vk = 0; size_t start, cnt_at; unsigned prev = 0;
    dsvcu_scan_layout(w, h, part);
    dsv_bw_align(bw); start = dsv_bw_byte(bw); dsv_bw_bits(bw, 32, 0); dsv_bw_seg(bw, dc);
    dsv_bw_align(bw); cnt_at = dsv_bw_byte(bw); dsv_bw_bits(bw, 24, 0); dsv_bw_align(bw);
    for (i = 0; i < nsyms; i++) {
        unsigned pos = syms[i].pos;
        while (l < 2 && pos >= (unsigned) part[l + 2]) l++;
        dsv_bw_ueg(bw, pos - prev);
        if (l < 0) dsv_bw_neg(bw, syms[i].v); else dsv_bw_nrice(bw, syms[i].v, &vk, 3 + l);
        prev = pos + 1;
    }
    dsv_bw_align(bw); dsv_bw_patch24(bw, cnt_at, (unsigned) nsyms);
    dsv_bw_bits(bw, 8, DSV_EOP_SYMBOL); dsv_bw_align(bw);
    dsv_bw_patch32(bw, start, (unsigned) (dsv_bw_byte(bw) - start - 4));
}
static int old_read(DSV_BITRD *br, dsvcu_symbol *syms, int cap, int w, int h, int level_start[5], int *dc)
{
    int part[5], total, n = 0, l = -1, vk = 0, i; unsigned plen; size_t start, limit; int runs, truncated = 0; unsigned cur = 0, run;
    total = dsvcu_scan_layout(w, h, part);
    for (i = 0; i < 5; i++) level_start[i] = 0;
    *dc = 0; dsv_br_align(br); plen = dsv_br_bits(br, 32); dsv_br_align(br);
    if (!(plen > 0 && plen < (unsigned) w * (unsigned) h * sizeof(DSV_SBC) * 2)) return -1;
    start = dsv_br_byte(br); limit = start + plen; *dc = dsv_br_seg(br);
    dsv_br_align(br); runs = (int) dsv_br_bits(br, 24); dsv_br_align(br);
    run = (runs-- > 0) ? dsv_br_ueg(br) : UINT_MAX;
    while (run != UINT_MAX) {
        unsigned pos = cur + run; int v;
        if (pos >= (unsigned) total || pos < cur) break;
        while (l < 2 && pos >= (unsigned) part[l + 2]) { l++; level_start[l + 1] = n; }
        v = (l < 0) ? dsv_br_neg(br) : dsv_br_nrice(br, &vk, 3 + l);
        run = (runs-- > 0) ? dsv_br_ueg(br) : UINT_MAX;
        if (dsv_br_byte(br) >= limit) { truncated = 1; break; }
        if (n < cap && pos != 0) { syms[n].pos = pos; syms[n].v = v; n++; }
        cur = pos + 1;
    }
    while (l < 2) { l++; level_start[l + 1] = n; }
    level_start[4] = n;
    if (!truncated) dsv_br_align(br);
    if (dsv_br_bits(br, 8) != DSV_EOP_SYMBOL) { br->pos = limit * 8; return -1; }
    br->pos = limit * 8;
    return n;
}
int main(int argc, char **argv)
{
    const int planes = argc > 1 ? atoi(argv[1]) : 3000;
    static const int dims[][2] = { {1920, 1080}, {352, 288}, {64, 48}, {17, 9}, {960, 540}, {2048, 256} };
    dsvcu_symbol *s = malloc(sizeof(*s) * 3000000), *o1 = malloc(sizeof(*s) * 3000000), *o2 = malloc(sizeof(*s) * 3000000);
    int it, bad = 0; long long tot = 0;
    srand(12345);
    for (it = 0; it < planes && !bad; it++) {
        int w = dims[it % 6][0], h = dims[it % 6][1], part[5], total = dsvcu_scan_layout(w, h, part), n = 0;
        int mode = rand() % 6, dc = (rand() % 4001) - 2000;
        unsigned pos = rand() % 3 ? 1 : 0;
        while (pos < (unsigned) total) {
            int gap, v;
            switch (mode) {
                case 0: gap = rand() % 4; v = (rand() % 5) - 2; break;
                case 1: gap = rand() % 3000; v = (rand() % 2001) - 1000; break;
                case 2: gap = (rand() % 50 == 0) ? rand() % 200000 : rand() % 8; v = (rand() % 9) - 4; break;
                case 3: gap = 0; v = (rand() % 65) - 32; break;
                case 4: gap = rand() % 30; v = (rand() % 300 == 0) ? (rand() % 200001) - 100000 : (rand() % 3) - 1; break;
                default: gap = rand() % 100; v = (rand() % 2) ? (1 << 17) - rand() % 5 : -(1 << 17) + rand() % 5; if (rand() % 200) v = (rand() % 7) - 3; break;
            }
            if (!v) v = (rand() & 1) ? 1 : -1;
            pos += gap; if (pos >= (unsigned) total) break;
            if (pos != 0) { s[n].pos = pos; s[n].v = v; n++; }
            pos++;
        }
        DSV_BITWR a, b; int pre = rand() % 3;
        dsv_bw_init(&a, 64); dsv_bw_init(&b, 64);
        if (pre) { dsv_bw_bits(&a, 5, 19); dsv_bw_bits(&b, 5, 19); } /* unaligned start is aligned by the writer */
        old_write(&a, s, n, dc, w, h);
        dsv_hzcc_write_plane(&b, s, n, dc, w, h);
        if (a.pos != b.pos || memcmp(a.buf, b.buf, dsv_bw_byte(&a) + 1)) { printf("WRITE MISMATCH it=%d mode=%d n=%d pos %zu vs %zu\n", it, mode, n, a.pos, b.pos); bad = 1; }
        /* trailing zero check: OR-writer invariants */
        { size_t k; for (k = dsv_bw_byte(&b) + 1; k < b.cap; k++) if (b.buf[k]) { printf("dirty tail it=%d at %zu\n", it, k); bad = 1; break; } }
        {
            DSV_BITRD r1, r2; int l1[5], l2[5], d1, d2, m1, m2; size_t len = dsv_bw_byte(&a);
            /* also exercise truncation: sometimes lie about the length */
            size_t use = (it % 7 == 3 && len > 40) ? len - 1 - rand() % 20 : len;
            uint8_t *copy = calloc(1, len + 32); memcpy(copy, a.buf, len);
            if (it % 11 == 5 && len > 60) copy[20 + rand() % (len - 40)] ^= 1 << (rand() % 8); /* bit error */
            dsv_br_init(&r1, copy, use); dsv_br_init(&r2, copy, use);
            if (pre) { r1.pos = 5; r2.pos = 5; }
            m1 = old_read(&r1, o1, 3000000, w, h, l1, &d1);
            m2 = dsv_hzcc_read_plane(&r2, o2, 3000000, w, h, l2, &d2);
            if (m1 != m2 || d1 != d2 || r1.pos != r2.pos || memcmp(l1, l2, sizeof(l1)) || (m1 > 0 && memcmp(o1, o2, sizeof(*s) * m1))) {
                printf("READ MISMATCH it=%d mode=%d n=%d m %d vs %d pos %zu vs %zu\n", it, mode, n, m1, m2, r1.pos, r2.pos); bad = 1;
            }
            free(copy);
        }
        tot += n;
        dsv_bw_free(&a); dsv_bw_free(&b);
    }
    printf("%s: %d planes, %lld symbols\n", bad ? "FAILED" : "all equal", it, tot);
    return bad;
}
