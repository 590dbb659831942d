/*
 * k_bmc.cuh -- block motion compensation, residual subtract / reconstruct.
 *
 * Replaces reference src/bmc.c: predict (:814-923), luma_qp (:661-769),
 * bilinear_sp (:772-812), avgval/cpyblk (:25-49), subtract (:989-1055) and
 * reconstruct (:925-987).
 *
 * One CTA per (motion block, plane).  The reference-window (+3 px for the
 * 4-tap luma filter) is staged once in shared memory, the separable filter
 * runs out of shared memory, and the residual arithmetic is fused into the
 * same launch: the encoder variant writes prediction + residual (dsv_sub_pred,
 * bmc.c:1057-1070); the decoder variant writes the reconstructed pixel
 * directly (predict + reconstruct of dsv_add_pred, bmc.c:1093-1111) so the
 * prediction never round-trips through HBM.
 */
#ifndef K_BMC_CUH
#define K_BMC_CUH

#include "dsvcu_rt.h"
#include "k_quant.cuh" /* dsvcu_mv + flag bits */

#define BMC_BORDER 32
#define BMC_MAXB 32
#define BMC_THREADS 128

struct BmcPlane {
    const uint8_t *ref; /* pixel (0,0) of the extended reference plane */
    int ref_stride;
    uint8_t *pred;      /* encoder: prediction plane (written); decoder: unused */
    int pred_stride;
    uint8_t *res;       /* residual plane: encoder output; decoder input */
    int res_stride;
    const uint8_t *src; /* encoder: source picture plane (may be the residual plane itself: in place) */
    int src_stride;
    uint8_t *out;       /* decoder: reconstructed output plane */
    int out_stride;
    int w, h;           /* plane size */
    int sh, sv;         /* chroma shifts (0 for luma) */
};

struct BmcArgs {
    BmcPlane pl[3];
    const dsvcu_mv *mvs;
    int nbh, nbv;
    int blk_w, blk_h;
    int tmc;      /* temporal MC flag (fnum % 2) */
    int lossless;
    int mode;     /* 0 = encoder sub_pred, 1 = decoder add_pred */
};

DSVCU_HD int bmc_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
DSVCU_HD int bmc_u8(int v) { return v > 255 ? 255 : (v < 0 ? 0 : v); }
DSVCU_HD int bmc_sar(int v, int s) { return v >> s; }

#define HPF_A(a, b, c, d) ((19 * ((b) + (c))) - (3 * ((a) + (d))))
#define HPF_B(a, b, c, d) ((20 * ((b) + (c))) - (4 * ((a) + (d))))

DSVCU_HD int
bmc_blend(int f, int b, int c, int frac) /* bmc.c:701-714, BF_SHIFT 6, BF_MULADD 32 */
{
    switch (frac) {
        case 0: return (32 * 2 * b + 32) >> 6;
        case 1: return (f + 32 * b + 32) >> 6;
        case 2: return (f * 2 + 32) >> 6;
        default: return (f + 32 * c + 32) >> 6;
    }
}

DSVCU_KERNEL void __launch_bounds__(BMC_THREADS)
k_predict(BmcArgs A)
{
    DSVCU_SHARED uint8_t win[(BMC_MAXB + 3) * (BMC_MAXB + 4)];
    DSVCU_SHARED int16_t tmp[(BMC_MAXB + 3) * BMC_MAXB];
    DSVCU_SHARED uint8_t prd[BMC_MAXB * BMC_MAXB];
    DSVCU_SHARED int sums[4];

    /* a CTA owns one motion block and walks its planes: a third of the CTAs of a launch per
     * (block, plane), which was dominated by CTA start-up for 8x8 chroma blocks */
    for (int c = (int) blockIdx.y; c < 3; c += (int) gridDim.y) {
    const BmcPlane P = A.pl[c];
    const int bi = (int) blockIdx.x % A.nbh, bj = (int) blockIdx.x / A.nbh;
    const dsvcu_mv mv = A.mvs[bi + bj * A.nbh];
    const int bw = A.blk_w >> P.sh, bh = A.blk_h >> P.sv;
    const int lbw = bw >= 32 ? 5 : (bw >= 16 ? 4 : (bw >= 8 ? 3 : (bw >= 4 ? 2 : (bw >= 2 ? 1 : 0))));
    const int x = bi * bw, y = bj * bh;
    const int limx = (P.w - bw) + BMC_BORDER - 1;
    const int limy = (P.h - bh) + BMC_BORDER - 1;
    const int WS = BMC_MAXB + 4; /* window row pitch */
    int px = x + bmc_sar(mv.x, 2 + P.sh);
    int py = y + bmc_sar(mv.y, 2 + P.sv);
    const int intra = mv.flags & MVF_INTRA;

    if (intra) {
        /* D.2: DC fill of the whole block or of masked quadrants */
        px = bmc_clampi(px, -BMC_BORDER, limx);
        py = bmc_clampi(py, -BMC_BORDER, limy);
        PAR_FOR(k, bw * bh) {
            int r = k >> lbw, q = k & (bw - 1); /* bw is a power of two */
            win[r * WS + q] = P.ref[(py + r) * P.ref_stride + px + q];
        }
        PAR_FOR(k, 4) { sums[k] = 0; }
        DSVCU_SYNC();
        const int whole = (mv.submask == 15);
        const int sbw = bw / 2, sbh = bh / 2;
        const int use_dc = (c == 0 && mv.dc);
        if (!use_dc) {
            /* per-quadrant (or whole-block) sums; integer sums are order-free */
            int part[4] = { 0, 0, 0, 0 };
            PAR_FOR(k, bw * bh) {
                int r = k >> lbw, q = k & (bw - 1); /* bw is a power of two */
                int qi = whole ? 0 : ((r >= sbh) * 2 + (q >= sbw));
                part[qi] += win[r * WS + q];
            }
            for (int i = 0; i < 4; i++) {
                if (part[i]) atomicAdd(&sums[i], part[i]);
            }
        }
        DSVCU_SYNC();
        PAR_FOR(k, bw * bh) {
            int r = k >> lbw, q = k & (bw - 1); /* bw is a power of two */
            int qi = (r >= sbh) * 2 + (q >= sbw);
            int v;
            if (whole) {
                v = use_dc ? mv.dc : sums[0] / (bw * bh);
            } else if (mv.submask & (1 << qi)) {
                v = use_dc ? mv.dc : sums[qi] / (sbw * sbh);
            } else {
                v = win[r * WS + q];
            }
            prd[r * BMC_MAXB + q] = (uint8_t) v; /* memset semantics: low 8 bits */
        }
    } else if (c == 0) {
        if (!((mv.x | mv.y) & 3)) {
            px = bmc_clampi(px, -BMC_BORDER, limx);
            py = bmc_clampi(py, -BMC_BORDER, limy);
            PAR_FOR(k, bw * bh) {
                int r = k >> lbw, q = k & (bw - 1); /* bw is a power of two */
                prd[r * BMC_MAXB + q] = P.ref[(py + r) * P.ref_stride + px + q];
            }
        } else {
            /* D.1 separable 4-tap, two sharpnesses (bmc.c:661-769) */
            px = bmc_clampi(px - 1, -BMC_BORDER, limx);
            py = bmc_clampi(py - 1, -BMC_BORDER, limy);
            int adx = mv.x < 0 ? -mv.x : mv.x, ady = mv.y < 0 ? -mv.y : mv.y;
            int large = adx >= 8 || ady >= 8;
            int dx = mv.x & 3, dy = mv.y & 3;
            int dqtx = large || !(dx & 1) || (A.tmc & 1);
            int dqty = large || !(dy & 1) || (A.tmc & 1);
            /* window rows are WS = 36 apart: index it as (row, column) of a 64-wide grid, no division */
            PAR_FOR(k, (bh + 3) * 64) {
                int r = k >> 6, q = k & 63;
                if (q < bw + 3) win[r * WS + q] = P.ref[(py + r) * P.ref_stride + px + q];
            }
            DSVCU_SYNC();
            PAR_FOR(k, (bh + 3) * bw) {
                int r = k >> lbw, q = k & (bw - 1); /* bw is a power of two */
                const uint8_t *s = win + r * WS + q;
                int a = s[0], b = s[1], cc = s[2], d = s[3];
                int f = dqtx ? HPF_A(a, b, cc, d) : HPF_B(a, b, cc, d);
                tmp[r * BMC_MAXB + q] = (int16_t) bmc_blend(f, b, cc, dx);
            }
            DSVCU_SYNC();
            PAR_FOR(k, bh * bw) {
                int r = k >> lbw, q = k & (bw - 1); /* bw is a power of two */
                const int16_t *s = tmp + r * BMC_MAXB + q;
                int a = s[0], b = s[BMC_MAXB], cc = s[2 * BMC_MAXB], d = s[3 * BMC_MAXB];
                int f = dqty ? HPF_A(a, b, cc, d) : HPF_B(a, b, cc, d);
                prd[r * BMC_MAXB + q] = (uint8_t) bmc_u8(bmc_blend(f, b, cc, dy));
            }
        }
    } else {
        /* chroma bilinear at 1/(4<<shift) pel (bmc.c:772-812) */
        px = bmc_clampi(px, -BMC_BORDER, limx);
        py = bmc_clampi(py, -BMC_BORDER, limy);
        int hbits = 2 + P.sh, vbits = 2 + P.sv;
        int hf = 1 << hbits, vf = 1 << vbits;
        int dx = mv.x & (hf - 1), dy = mv.y & (vf - 1);
        if (dx | dy) {
            int f0 = (hf - dx) * (vf - dy), f1 = dx * (vf - dy);
            int f2 = (hf - dx) * dy, f3 = dx * dy;
            int sf = hbits + vbits, af = 1 << (sf - 1);
            PAR_FOR(k, (bh + 1) * 64) {
                int r = k >> 6, q = k & 63;
                if (q < bw + 1) win[r * WS + q] = P.ref[(py + r) * P.ref_stride + px + q];
            }
            DSVCU_SYNC();
            PAR_FOR(k, bh * bw) {
                int r = k >> lbw, q = k & (bw - 1); /* bw is a power of two */
                const uint8_t *s = win + r * WS + q;
                prd[r * BMC_MAXB + q] =
                    (uint8_t) ((f0 * s[0] + f1 * s[1] + f2 * s[WS] + f3 * s[WS + 1] + af) >> sf);
            }
        } else {
            PAR_FOR(k, bw * bh) {
                int r = k >> lbw, q = k & (bw - 1); /* bw is a power of two */
                prd[r * BMC_MAXB + q] = P.ref[(py + r) * P.ref_stride + px + q];
            }
        }
    }
    DSVCU_SYNC();

    const int skip = mv.flags & MVF_SKIP, eprm = mv.flags & MVF_EPRM;
    if (A.mode == 0) {
        /* encoder: store prediction, residual = source - prediction (bmc.c:989-1055) */
        const int noxmit = !intra && (skip || (c == 0 && (mv.flags & MVF_NOXMITY)) ||
                                      (c != 0 && (mv.flags & MVF_NOXMITC)));
        PAR_FOR(k, bw * bh) {
            int r = k >> lbw, q = k & (bw - 1); /* bw is a power of two */
            int p = prd[r * BMC_MAXB + q];
            uint8_t *rp = P.res + (y + r) * P.res_stride + x + q;
            int s = P.src[(y + r) * P.src_stride + x + q], o;
            P.pred[(y + r) * P.pred_stride + x + q] = (uint8_t) p;
            if (A.lossless) {
                o = (s - p + 128) & 0xff;
            } else if (noxmit) {
                o = 128;
            } else if (eprm) {
                o = bmc_u8((s - p + 256) >> 1);
            } else {
                o = bmc_u8(s - p + 128);
            }
            *rp = (uint8_t) o;
        }
    } else {
        /* decoder: out = prediction + residual (bmc.c:925-987) */
        const int plain = !eprm || (!intra && skip);
        PAR_FOR(k, bw * bh) {
            int r = k >> lbw, q = k & (bw - 1); /* bw is a power of two */
            int p = prd[r * BMC_MAXB + q];
            int s = P.res[(y + r) * P.res_stride + x + q], o;
            if (A.lossless) {
                o = (p + s - 128) & 0xff;
            } else if (plain) {
                o = bmc_u8(p + s - 128);
            } else {
                o = bmc_u8(p + (s - 128) * 2);
            }
            P.out[(y + r) * P.out_stride + x + q] = (uint8_t) o;
        }
    }
    DSVCU_SYNC(); /* the staging arrays are reused by the next plane */
    }
}

/* encoder-side reconstruct: res = pred (+) res, in place (dsv_add_res, bmc.c:1072-1090).
 * Pure per-pixel arithmetic steered by the block's flags: one thread per 16-byte segment of a
 * block row (8 / 4 bytes where a chroma block is narrower), 16-byte loads and stores -- planes
 * start 16-byte aligned, strides are multiples of 16, block widths powers of two.  Like the
 * reference it works on whole blocks, i.e. also on the border columns / rows the block grid
 * covers past the picture edge. */
#define REC_THREADS 256

DSVCU_HD uint32_t
bmc_rec4(uint32_t p4, uint32_t s4, int lossless, int plain)
{
    uint32_t o4 = 0;
    for (int k = 0; k < 4; k++) {
        const int p = (int) ((p4 >> (8 * k)) & 255u), s = (int) ((s4 >> (8 * k)) & 255u);
        int o;
        if (lossless) {
            o = (p + s - 128) & 0xff;
        } else if (plain) {
            o = bmc_u8(p + s - 128);
        } else {
            o = bmc_u8(p + (s - 128) * 2);
        }
        o4 |= (uint32_t) o << (8 * k);
    }
    return o4;
}

DSVCU_KERNEL void __launch_bounds__(REC_THREADS)
k_reconstruct(BmcArgs A)
{
    const int c = (int) blockIdx.y;
    const BmcPlane P = A.pl[c];
    const int bw = A.blk_w >> P.sh, bh = A.blk_h >> P.sv;
    const int seg = bw >= 16 ? 16 : bw;            /* bytes per work item: 16, 8 or 4 */
    const int segs_row = (A.nbh * bw) / seg, rows = A.nbv * bh;
    const int total = segs_row * rows;
    for (int k = (int) blockIdx.x * DSVCU_NTH + DSVCU_TID; k < total; k += (int) gridDim.x * DSVCU_NTH) {
        const int row = k / segs_row, x = (k - row * segs_row) * seg;
        const dsvcu_mv *mvp = A.mvs + (x / bw) + (row / bh) * A.nbh;
        const unsigned flags = mvp->flags;
        const int intra = flags & MVF_INTRA, skip = flags & MVF_SKIP, eprm = flags & MVF_EPRM;
        const int plain = !eprm || (!intra && skip);
        const uint8_t *pp = P.pred + (ptrdiff_t) row * P.pred_stride + x;
        uint8_t *rp = P.res + (ptrdiff_t) row * P.res_stride + x;
        if (seg == 16) {
            uint4 p = *(const uint4 *) pp, s = *(const uint4 *) rp, o;
            o.x = bmc_rec4(p.x, s.x, A.lossless, plain);
            o.y = bmc_rec4(p.y, s.y, A.lossless, plain);
            o.z = bmc_rec4(p.z, s.z, A.lossless, plain);
            o.w = bmc_rec4(p.w, s.w, A.lossless, plain);
            *(uint4 *) rp = o;
        } else {
            for (int i = 0; i < seg; i += 4) {
                *(uint32_t *) (rp + i) = bmc_rec4(*(const uint32_t *) (pp + i), *(const uint32_t *) (rp + i), A.lossless, plain);
            }
        }
    }
}

#endif /* K_BMC_CUH */
