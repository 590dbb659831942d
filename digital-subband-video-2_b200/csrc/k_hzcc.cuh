/*
 * k_hzcc.cuh -- entropy decode on the device: the coefficient planes of a batch of
 * pictures and the side information (vector fields, block flags) of its inter pictures.
 *
 * Replaces the bit-parsing half of reference src/hzcc.c: dsv_decode_plane
 * (:585-649) / hzcc_dec (:450-583) with the codes of src/bs.c (interleaved
 * exp-Golomb :96-137, adaptive Rice :237-251): a serialised plane
 *   [32-bit byte length][SEG dc][align][24-bit pair count][align]
 *   [(UEG run, value)...][align][0x55]
 * becomes the ordered list of (scan position, value) symbols that the
 * de-quantiser kernels (k_quant.cuh) scatter -- the same list the host parser
 * (host/dsv_hzcc.c) produces.
 *
 * A plane is one serial chain (every code length depends on the previous
 * code, the Rice parameter adapts per symbol), but planes are independent and
 * pictures carry no entropy-coder state: one warp per plane, all planes of
 * a batch of pictures (a closed GOP) in one launch on a stream of its own, so
 * that the chains run beside the reconstruction of the pictures before them.
 * The side information of an inter picture (second half of this file) is one
 * more such chain per picture.  Off the host this removes ~80 % of the CPU
 * time of a decoded picture (1.7 -> 0.3 ms at 1080p).
 *
 * The kernel accepts only what a well-formed encoder writes.  Anything else
 * (truncated plane, bad end marker, position overflow, oversized Rice
 * parameter, ...) clears the plane's `ok` word and the host parser -- which
 * reproduces the reference's behaviour on damaged input -- decodes that picture
 * instead.
 */
#ifndef K_HZCC_CUH
#define K_HZCC_CUH

#include "dsvcu_rt.h"
#include "k_quant.cuh"
#ifndef DSVCU_EMU
#define HZT_FN __host__ __device__ __forceinline__ static
#define HZT_CLZ32(v) hzt_clz32_(v)
__host__ __device__ __forceinline__ static int
hzt_clz32_(uint32_t v)
{
#ifdef __CUDA_ARCH__
    return __clz((int) v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}
#endif
#include "hz_table.h"

#define HZ_EOP 0x55
#define HZ_META_WORDS 8 /* nsym, level_start[0..4], dc, ok */
#define HZ_META_NSYM 0
#define HZ_META_LS 1
#define HZ_META_DC 6
#define HZ_META_OK 7

struct HzSpan {
    uint32_t off;      /* first byte of the plane (its length word) in the batch buffer; multiple of 8 */
    uint32_t len;      /* bytes available to this plane: 4 + declared length */
    int w, h;          /* coefficient plane size */
    uint32_t sym_base; /* first symbol slot of this plane */
    uint32_t sym_cap;
};

/* The side information of one inter picture (reference dsv_decoder.c: decode_stability_blocks,
 * decode_motion; spec B.2.3.1, B.2.3.4): six length-prefixed sub-streams -- skip bits, then
 * mode bits, vector x, vector y, intra sub-block masks + DC, EPRM bits -- decoded into the
 * picture's vector field and block flags, in the layout dsvcu_set_side uploads. */
#define HZ_SIDE_NSUB 6
#define HZ_SIDE_SKIP 0 /* then the five of DSV_SUB_*: mode, x, y, sbim, eprm */

struct HzSide {
    uint32_t off[HZ_SIDE_NSUB]; /* first byte of each sub-stream in the batch buffer (multiples of 4 apart
                                 * from the base: see `shift`) */
    uint32_t len[HZ_SIDE_NSUB]; /* its coded length in bytes; the readers may look 8 bytes further */
    int nbh, nbv;               /* blocks per row / column */
    int flips;                  /* bit i: the bits of RLE stream i are stored inverted (0 skip, 1 mode, 2 eprm) */
    uint32_t out_off;           /* this picture's block of `side_out` */
    uint32_t mv_bytes;          /* size of the vector part of that block; the flags follow it */
};

struct HzJob {
    const uint8_t *bits; /* batch buffer, 16 zero bytes behind the last plane */
    const HzSpan *spans;
    int nspans;
    dsvcu_sym *syms;
    int *meta; /* HZ_META_WORDS per span */
    /* side-information chains of the batch part, handled by the warps behind the planes' */
    const HzSide *sides;
    int nsides;
    uint8_t *side_out;
    int *side_ok;
};

/* big-endian bit reader: `win` holds the next `n` bits left-aligned (zeros below them); the
 * bytes behind them start at buf[nb], and the word at buf[nb] is already in `nxt` (loaded one
 * refill ahead, so that its latency is not on the dependent chain) */
struct HzRd {
    const uint8_t *buf;
    uint32_t nb;  /* next byte to enter the window */
    uint32_t end; /* bytes in the plane (reads beyond see zeros) */
    uint64_t win;
    int n;
    uint32_t nxt;
};

DSVCU_DEV int
hz_clz32(uint32_t v)
{
#ifndef DSVCU_EMU
    return __clz((int) v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}

/* the raw (little-endian as loaded) word at byte `at`, a multiple of 4; the batch buffer is
 * padded, reads behind the plane see zeros */
DSVCU_DEV uint32_t
hz_raw(const HzRd &r, uint32_t at)
{
    return at < r.end ? *(const uint32_t *) (r.buf + at) : 0u;
}

DSVCU_DEV uint32_t
hz_swap(uint32_t w)
{
#ifndef DSVCU_EMU
    return __byte_perm(w, 0, 0x0123);
#else
    return __builtin_bswap32(w);
#endif
}

DSVCU_DEV void
hz_open(HzRd &r, const uint8_t *buf, uint32_t len)
{
    r.buf = buf;
    r.end = len;
    r.win = ((uint64_t) hz_swap(hz_raw(r, 0)) << 32) | hz_swap(hz_raw(r, 4));
    r.n = 64;
    r.nb = 8;
    r.nxt = hz_raw(r, 8);
}

/* at least 32 valid bits in the window.  The word that enters was loaded at the previous
 * refill and is not touched before this one (not even byte-swapped): its latency stays off
 * the dependent chain. */
DSVCU_DEV void
hz_fill(HzRd &r)
{
    if (r.n <= 32) {
        r.win |= (uint64_t) hz_swap(r.nxt) << (32 - r.n);
        r.nb += 4;
        r.n += 32;
        r.nxt = hz_raw(r, r.nb);
    }
}

DSVCU_DEV void
hz_skip(HzRd &r, int k)
{
    r.win <<= k;
    r.n -= k;
}

/* bits consumed so far, from the start of the plane */
DSVCU_DEV uint32_t
hz_pos(const HzRd &r)
{
    return r.nb * 8u - (uint32_t) r.n;
}

DSVCU_DEV uint32_t
hz_bits(HzRd &r, int k) /* 0 <= k <= 32 */
{
    hz_fill(r);
    if (k == 0) return 0;
    uint32_t v = (uint32_t) (r.win >> (64 - k));
    hz_skip(r, k);
    return v;
}

DSVCU_DEV void
hz_align(HzRd &r)
{
    hz_skip(r, r.n & 7); /* the window ends on a byte boundary */
}

#define HZ_BAD 0xffffffffu

/* interleaved exp-Golomb (bs.c:96-137): (0 d)* 1 -- any length; HZ_BAD when the code runs
 * off the plane */
DSVCU_DEV uint32_t
hz_ueg(HzRd &r)
{
    uint32_t v = 1;
    for (;;) {
        hz_fill(r);
        const uint32_t top = (uint32_t) (r.win >> 32);
        const uint32_t stops = top & 0xAAAAAAAAu;
        if (stops == 0) { /* 16 pairs, no terminator yet */
            v = (v << 16) | hzt_even_bits(top);
            hz_skip(r, 32);
            if (r.nb > r.end + 16) return HZ_BAD;
            continue;
        }
        const int pairs = hz_clz32(stops) >> 1;
        if (pairs) v = (v << pairs) | hzt_even_bits(top >> (32 - 2 * pairs));
        hz_skip(r, 2 * pairs + 1);
        return v - 1;
    }
}

DSVCU_DEV int
hz_seg(HzRd &r)
{
    int v = (int) hz_ueg(r);
    if (v && hz_bits(r, 1)) return -v;
    return v;
}

/* nonzero value of the LL part (bs.c: UEG of |v| - 1, then the sign) */
DSVCU_DEV int
hz_neg(HzRd &r)
{
    const int v = (int) hz_ueg(r) + 1;
    return hz_bits(r, 1) ? -v : v;
}

/* adaptive Rice (bs.c:237-251); k <= 24 is checked by the caller */
DSVCU_DEV int
hz_nrice(HzRd &r, int *rk, int k)
{
    uint32_t q, uv;
    hz_fill(r);
    const uint32_t top = (uint32_t) (r.win >> 32);
    const int z = hz_clz32(top);
    if (z + 1 + k <= 32) { /* the whole code is in the upper half of the window */
        const uint32_t rest = top << z << 1;
        q = (uint32_t) z;
        uv = (q << k) | ((rest >> 1) >> (31 - k));
        hz_skip(r, z + 1 + k);
    } else {
        q = 0;
        for (;;) {
            hz_fill(r);
            const uint32_t t = (uint32_t) (r.win >> 32);
            if (t == 0) {
                q += 32;
                hz_skip(r, 32);
                if (r.nb > r.end + 16) break;
                continue;
            }
            const int n = hz_clz32(t);
            q += (uint32_t) n;
            hz_skip(r, n + 1);
            break;
        }
        uv = (q << k) | hz_bits(r, k);
    }
    *rk += q ? 1 : -(*rk > 0);
    uv += 1;
    return (int) (uv >> 1) ^ -(int) (uv & 1);
}

/* scan-order partition of a plane (same as dsvcu_scan_layout) */
DSVCU_DEV int
hz_layout(int w, int h, int part[5])
{
    int pos = ((w + 7) >> 3) * ((h + 7) >> 3);
    part[0] = 0;
    for (int l = 0; l < 3; l++) {
        part[1 + l] = pos;
        pos += 3 * ((w + (1 << (3 - l)) - 1) >> (3 - l)) * ((h + (1 << (3 - l)) - 1) >> (3 - l));
    }
    part[4] = pos;
    return pos;
}

/* the prefix table lives in shared memory; on the device it is addressed through its 32-bit
 * shared-window address, held in a register for the whole walk (the compiler otherwise rebuilds
 * the window base from a special register in front of every look-up, on the dependent chain) */
#ifndef DSVCU_EMU
typedef uint32_t hz_tab_t;
DSVCU_DEV hz_tab_t
hz_tab_ref(const uint32_t *tab)
{
    uint32_t a = (uint32_t) __cvta_generic_to_shared(tab), kept;
    asm volatile("mov.u32 %0, %1;" : "=r"(kept) : "r"(a)); /* opaque: not rematerialised */
    return kept;
}
DSVCU_DEV uint32_t
hz_tab_at(hz_tab_t tab, uint32_t idx)
{
    uint32_t e;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(tab + idx * 4u));
    return e;
}
#else
typedef const uint32_t *hz_tab_t;
DSVCU_DEV hz_tab_t hz_tab_ref(const uint32_t *tab) { return tab; }
DSVCU_DEV uint32_t hz_tab_at(hz_tab_t tab, uint32_t idx) { return tab[idx]; }
#endif

/* one plane; returns 1 when it was well-formed.  `tab`: HZT_ROWS rows of HZT_SIZE entries */
DSVCU_DEV int
hz_parse_plane(const HzJob &J, const HzSpan &S, int *meta, const hz_tab_t tab)
{
    HzRd r;
    int part[5];
    const uint32_t total = (uint32_t) hz_layout(S.w, S.h, part);
    dsvcu_sym *out = J.syms + S.sym_base;
    hz_open(r, J.bits + S.off, S.len);
    const uint32_t plen = hz_bits(r, 32);
    if (!(plen > 0 && (uint64_t) plen < (uint64_t) S.w * (uint64_t) S.h * sizeof(int32_t) * 2) || plen + 4 != S.len) {
        return 0;
    }
    const int dc = hz_seg(r);
    hz_align(r);
    const uint32_t runs = hz_bits(r, 24);
    hz_align(r);
    if (runs > S.sym_cap) return 0; /* (the host sized the symbol slots from this very field) */
    /* the parts of the scan in turn: LL (l = -1), then levels 0..2; `bound` = first position
     * behind the part, `ls` = index of the first symbol of each part.  Every step of the
     * chain is  value of this pair, run of the next pair: one table look-up where both codes
     * fit into HZT_BITS bits (nearly always), the general readers otherwise. */
    int n = 0, l = -1, vk = 0, ls1 = 0, ls2 = 0, ls3 = 0;
    uint32_t cur = 0, bound = (uint32_t) part[1];
    uint32_t run = runs ? hz_ueg(r) : 0;
    if (run > 0x3fffffffu) return 0;
    for (uint32_t i = 0; i < runs; i++) {
        const uint32_t pos = cur + run;
        while (pos >= bound) {
            if (pos >= total) return 0;
            l++;
            if (l == 0) ls1 = n;
            if (l == 1) ls2 = n;
            if (l == 2) ls3 = n;
            bound = l < 2 ? (uint32_t) part[l + 2] : total;
        }
        const int k = l < 0 ? HZT_ROW_LL : vk >> (3 + l);
        int v;
        uint32_t e = 0;
        hz_fill(r);
        if ((l < 0 || k < HZT_KMAX) && i + 1 < runs) e = hz_tab_at(tab, (uint32_t) k * HZT_SIZE + (uint32_t) (r.win >> (64 - HZT_BITS)));
        if (HZT_LEN(e)) {
            hz_skip(r, HZT_LEN(e));
            v = HZT_VAL(e);
            run = HZT_RUN(e);
            if (l >= 0) vk += HZT_QNZ(e) ? 1 : -(vk > 0);
        } else {
            if (l < 0) {
                v = hz_neg(r);
            } else {
                if (k > 24) return 0;
                v = hz_nrice(r, &vk, k);
            }
            run = i + 1 < runs ? hz_ueg(r) : 0;
            if (run > 0x3fffffffu) return 0; /* (runs out of the table are small) */
        }
        if (pos != 0) {
#ifndef DSVCU_EMU
            /* (the slots are 8-byte aligned: cudaMalloc base, 8-byte elements) */
            *(uint2 *) (out + n) = make_uint2(pos, (uint32_t) v);
            n++;
#else
            out[n].pos = pos;
            out[n].v = v;
            n++;
#endif
        }
        cur = pos + 1;
    }
    /* a pair whose bits end at or behind the declared length is the host parser's business
     * (bit positions only grow: looking once, here, is the same as looking after every pair) */
    if ((hz_pos(r) >> 3) >= S.len) return 0;
    if (l < 0) ls1 = n;
    if (l < 1) ls2 = n;
    if (l < 2) ls3 = n;
    hz_align(r);
    if (hz_bits(r, 8) != HZ_EOP || hz_pos(r) > S.len * 8u) return 0;
    meta[HZ_META_NSYM] = n;
    meta[HZ_META_LS + 0] = 0;
    meta[HZ_META_LS + 1] = ls1;
    meta[HZ_META_LS + 2] = ls2;
    meta[HZ_META_LS + 3] = ls3;
    meta[HZ_META_LS + 4] = n;
    meta[HZ_META_DC] = dc;
    return 1;
}

/* ---- side information ---- */

/* a reader over an arbitrarily aligned sub-stream: the window is filled from the 4-byte word
 * that holds the first byte, and the bits in front of it are dropped */
DSVCU_DEV void
hz_open_at(HzRd &r, const uint8_t *base, uint32_t off, uint32_t len)
{
    const uint32_t a = off & ~3u, lead = off & 3u;
    hz_open(r, base + a, len + lead);
    hz_skip(r, (int) lead * 8);
}

/* bits consumed from the start of the SUB-STREAM (hz_open_at) */
DSVCU_DEV uint32_t
hz_pos_at(const HzRd &r, uint32_t off)
{
    return hz_pos(r) - (off & 3u) * 8u;
}

struct HzRle {
    HzRd r;
    uint32_t nz;
};

/* dsv_rle_rd_get (bs.c:277-330): zero-bit run lengths in exp-Golomb */
DSVCU_DEV int
hz_rle_get(HzRle &e)
{
    if (e.nz == 0) {
        e.nz = hz_ueg(e.r);
        return e.nz == 0;
    }
    e.nz--;
    return e.nz == 0;
}

DSVCU_DEV int
hz_iabs(int v)
{
    return v < 0 ? -v : v;
}

DSVCU_DEV int
hz_grad_pick(int left, int top, int topleft) /* dsv.c:324-331 */
{
    const int g = left + top - topleft;
    return hz_iabs(g - left) < hz_iabs(g - top) ? left : top;
}

#define HZ_BD_STABLE 1
#define HZ_BD_SKIP 4
#define HZ_BD_INTRA 16
#define HZ_BD_EPRM_BIT 5
#define HZ_MVF_INTRA 1u
#define HZ_MVF_EPRM 2u
#define HZ_MVF_SKIP 8u

/* One picture: the walk of host/dsv_dec.c read_stability + read_motion.  Returns 1 when no
 * reader went past the end of its sub-stream (what happens behind it depends on how the
 * host's reader is positioned; such pictures are the host's business). */
DSVCU_DEV int
hz_parse_side(const HzJob &J, const HzSide &D)
{
    dsvcu_mv *mvs = (dsvcu_mv *) (J.side_out + D.out_off);
    uint8_t *bd = J.side_out + D.out_off + D.mv_bytes;
    const int nbh = D.nbh, nbv = D.nbv;
    HzRle skip, mode, eprm;
    HzRd rx, ry, rb;
    hz_open_at(skip.r, J.bits, D.off[0], D.len[0] + 8);
    hz_open_at(mode.r, J.bits, D.off[1], D.len[1] + 8);
    hz_open_at(rx, J.bits, D.off[2], D.len[2] + 8);
    hz_open_at(ry, J.bits, D.off[3], D.len[3] + 8);
    hz_open_at(rb, J.bits, D.off[4], D.len[4] + 8);
    hz_open_at(eprm.r, J.bits, D.off[5], D.len[5] + 8);
    skip.nz = mode.nz = eprm.nz = 0;
    const int f_skip = D.flips & 1, f_mode = (D.flips >> 1) & 1, f_eprm = (D.flips >> 2) & 1;
    for (int j = 0; j < nbv; j++) {
        for (int i = 0; i < nbh; i++) {
            const int idx = i + j * nbh;
            dsvcu_mv m;
            m.x = m.y = 0;
            m.flags = 0;
            m.err = 0;
            m.dc = 0;
            m.submask = 0;
            m.pad_[0] = m.pad_[1] = m.pad_[2] = 0;
            if (hz_rle_get(skip) ^ f_skip) {
                m.flags = HZ_MVF_SKIP;
                mvs[idx] = m;
                bd[idx] = (uint8_t) (HZ_BD_SKIP | HZ_BD_STABLE);
                continue;
            }
            const int md = hz_rle_get(mode) ^ f_mode, ep = hz_rle_get(eprm) ^ f_eprm;
            int flags = ep << HZ_BD_EPRM_BIT;
            m.flags = (md ? HZ_MVF_INTRA : 0u) | (ep ? HZ_MVF_EPRM : 0u);
            /* predictor from the left, top and top-left vectors (dsv_movec_pred, dsv.c:333-357) */
            int lx = 0, ly = 0, tx = 0, ty = 0, dx = 0, dy = 0;
            dsvcu_mv L, T;
            L.x = L.y = 0;
            L.flags = 0;
            T = L;
            if (i > 0) {
                L = mvs[idx - 1];
                lx = L.x;
                ly = L.y;
            }
            if (j > 0) {
                T = mvs[idx - nbh];
                tx = T.x;
                ty = T.y;
                if (i > 0) {
                    const dsvcu_mv Q = mvs[idx - nbh - 1];
                    dx = Q.x;
                    dy = Q.y;
                }
            }
            int px = hz_grad_pick(lx, tx, dx), py = hz_grad_pick(ly, ty, dy);
            if (md) {
                px = (px + 2) >> 2;
                py = (py + 2) >> 2;
            }
            int vx = (int16_t) (hz_seg(rx) + px), vy = (int16_t) (hz_seg(ry) + py);
            if (md) {
                vx = (int16_t) (vx * 4); /* intra vectors are full-pel */
                vy = (int16_t) (vy * 4);
                m.submask = hz_bits(rb, 1) ? (uint8_t) 15 : (uint8_t) hz_bits(rb, 4);
                m.dc = hz_bits(rb, 1) ? (uint16_t) (hz_bits(rb, 8) | 0x100u) : (uint16_t) 0;
                flags |= HZ_BD_INTRA;
            }
            m.x = (int16_t) vx;
            m.y = (int16_t) vy;
            mvs[idx] = m;
            /* dsv_neighbordif (dsv.c:359-407): distance to the left and top vectors */
            if (!(hz_iabs(vx) < 2 && hz_iabs(vy) < 2)) {
                int ax = vx, ay = vy, bx = vx, by = vy;
                if (i > 0 && (L.x | L.y) != 0 && !(L.flags & HZ_MVF_SKIP)) {
                    ax = L.x;
                    ay = L.y;
                }
                if (j > 0 && (T.x | T.y) != 0 && !(T.flags & HZ_MVF_SKIP)) {
                    bx = T.x;
                    by = T.y;
                }
                const int nd = hz_iabs(ax - vx) + hz_iabs(ay - vy) + hz_iabs(bx - vx) + hz_iabs(by - vy);
                if (nd / 3 > 8) flags |= HZ_BD_STABLE;
            }
            bd[idx] = (uint8_t) flags;
        }
    }
    return hz_pos_at(skip.r, D.off[0]) <= D.len[0] * 8u && hz_pos_at(mode.r, D.off[1]) <= D.len[1] * 8u &&
           hz_pos_at(rx, D.off[2]) <= D.len[2] * 8u && hz_pos_at(ry, D.off[3]) <= D.len[3] * 8u &&
           hz_pos_at(rb, D.off[4]) <= D.len[4] * 8u && hz_pos_at(eprm.r, D.off[5]) <= D.len[5] * 8u;
}

/* One chain per warp (lane 0 walks it; a chain is latency-bound and wants a scheduler slot,
 * not lanes), HZ_WARPS chains per CTA sharing the prefix table in shared memory (40 KB).
 * Chains 0 .. nspans-1 are coefficient planes, the nsides behind them side information. */
#define HZ_WARPS 8

DSVCU_KERNEL void __launch_bounds__(HZ_WARPS * 32)
k_hzcc_parse(HzJob J)
{
    DSVCU_SHARED uint32_t tab[HZT_ROWS * HZT_SIZE];
    PAR_FOR(i, HZT_ROWS * HZT_SIZE) tab[i] = hzt_entry((uint32_t) i & (HZT_SIZE - 1), i >> HZT_BITS);
    DSVCU_SYNC();
    const int per_cta = DSVCU_NTH >= 32 ? DSVCU_NTH >> 5 : 1;
    const int s = (int) blockIdx.x * per_cta + (DSVCU_TID >> 5);
    if ((DSVCU_TID & 31) != 0 || s >= J.nspans + J.nsides) return;
    if (s >= J.nspans) {
        J.side_ok[s - J.nspans] = hz_parse_side(J, J.sides[s - J.nspans]);
        return;
    }
    int *meta = J.meta + s * HZ_META_WORDS;
    meta[HZ_META_OK] = hz_parse_plane(J, J.spans[s], meta, hz_tab_ref(tab));
}

#endif /* K_HZCC_CUH */
