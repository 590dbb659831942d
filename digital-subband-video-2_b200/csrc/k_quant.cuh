/*
 * k_quant.cuh -- adaptive quantisation / dequantisation of subband planes.
 *
 * Replaces the arithmetic half of reference src/hzcc.c: lfquant (:88-105) and
 * hfquant (:107-162) are evaluated on the host once per band (scalars);
 * TMQ4POS_P/I (:164-206), quantSUB/quantS (:209-211), dequantS/D (:217-228) and
 * the scan loops of hzcc_enc (:308-439) / hzcc_dec (:520-581) run here.
 *
 * The reference walks LL, then levels 0..2 x {LH,HL,HH} in raster order,
 * quantising in place, and each coefficient looks at its (already quantised)
 * parent and grand-parent.  On the GPU every scan rectangle of a level is
 * processed in parallel.  The only same-level read-after-write hazards are the
 * odd-size "aliased parent" rows/columns: the last
 * row/column of a band whose parent index lands on row/column 0 of a band of
 * the SAME level.  Those elements run in a second launch ("wave B") after the
 * rest of the level ("wave A"); the double-visited spill row is reproduced
 * because rectangles keep the reference's ceil() dimensions and the arithmetic
 * stays in place.
 *
 * Encoder output: a dense scan-order array qv[] (quantised values) which
 * k_compact_* turns into an ordered (scan position, value) list for the host
 * entropy coder.  Decoder input: the same list, produced by the host entropy
 * decoder.
 */
#ifndef K_QUANT_CUH
#define K_QUANT_CUH

#include "dsvcu_rt.h"

/* blockdata bits, reference dsv_internal.h:96-110 */
#define BD_STABLE 1
#define BD_MAINTAIN 2
#define BD_SKIP 4
#define BD_RING 8
#define BD_INTRA 16
#define BD_EPRM 32
#define BD_SIMCMPLX 64
/* DSV_MV flag bits, reference dsv.h:183-190 */
#define MVF_INTRA 1
#define MVF_EPRM 2
#define MVF_MAINTAIN 4
#define MVF_SKIP 8
#define MVF_RINGING 16
#define MVF_NOXMITY 32
#define MVF_NOXMITC 64
#define MVF_SIMCMPLX 128

struct dsvcu_mv { /* layout of DSV_MV, 16 bytes */
    int16_t x, y;
    uint32_t flags;
    uint16_t err;
    uint16_t dc;
    uint8_t submask;
    uint8_t pad_[3];
};

struct dsvcu_sym {
    uint32_t pos; /* position in scan order (hzcc.c C.1 traversal) */
    int32_t v;    /* quantised value, never 0 */
};

struct QuantBand {
    int ox, oy;   /* origin of the scan rectangle inside the plane */
    int pox, poy; /* origin of the parent rectangle */
    int gox, goy; /* origin of the grand-parent rectangle */
    int qp;       /* hfquant() result for this band */
    int scan_base;
};

struct QuantLevel {
    int32_t *coefs;
    int fw;
    int32_t *qv;               /* encoder: dense scan-order output */
    const dsvcu_sym *syms;     /* decoder: symbol list for this plane */
    int sym_begin, sym_end;    /* decoder: symbols of this level */
    int l;                     /* hzcc level 0..2, -1 for the LL part */
    int w, h;                  /* scan rectangle size (ceil dims) */
    int isP, luma, lossless, psy;
    int dbx, dby, nbh;
    const uint8_t *blockdata;
    const dsvcu_mv *mvs;
    int wave;
    QuantBand band[3];
};

/* up to three planes per launch: blockIdx.z selects the plane */
struct QuantJob {
    QuantLevel Q[3];
    int lfq[3]; /* LL step size per plane */
};

struct CompactJob {
    const int32_t *qv[3];
    int n[3];
    int *chunk[3];
    int nchunks[3];
    dsvcu_sym *out[3];
    int *out_n[3];            /* pinned host: symbol count */
    const int32_t *dc_src[3]; /* the plane's DC coefficient ... */
    int *dc_dst[3];           /* ... and where the host wants it */
};

DSVCU_HD int q_sub(int v, int q, int sub) { return ((v >= 0) ? v - sub : v + sub) / q; }
DSVCU_HD int q_deq_s(int v, int q) { return v * q + ((v < 0) ? -(q * 2 / 3) : (q * 2 / 3)); }
DSVCU_HD int q_deq_d(int v, int q) { return v * q + ((v < 0) ? -(q / 2) : (q / 2)); }
DSVCU_HD int q_sign(int x) { return x < 0 ? -1 : (x > 0 ? 1 : 0); }

DSVCU_HD int
q_tmq_p(int tmq, int flags, int parc)
{
    if (parc || (flags & (BD_STABLE | BD_EPRM))) return tmq * 7 >> 3;
    if (flags & BD_INTRA) return tmq * 6 >> 3;
    return tmq;
}

DSVCU_HD int
q_tmq_i(int tmq, int flags, int parc, int l)
{
    int sm = flags & (BD_STABLE | BD_MAINTAIN);
    if (l == 0) return tmq;
    if (l == 2) {
        if (sm == BD_STABLE) return tmq >> 2;
        if (sm == BD_MAINTAIN) return tmq >> ((flags & BD_RING) ? 2 : !parc);
        if (sm == (BD_STABLE | BD_MAINTAIN)) return tmq >> (2 + !parc);
        return tmq;
    }
    if (sm == BD_STABLE) return tmq / 3;
    if (sm == BD_MAINTAIN) return tmq >> ((flags & BD_RING) ? 2 : !parc);
    if (sm == (BD_STABLE | BD_MAINTAIN)) return tmq >> 2;
    return tmq;
}

/* is the parent of (x,y) in band `s` inside a band of the same level? */
DSVCU_HD int
q_parent_aliased(const QuantLevel &Q, const QuantBand &B, int x, int y)
{
    return (B.pox + (x >> 1) >= Q.w) || (B.poy + (y >> 1) >= Q.h);
}

/* ------------------------------------------------------------- encoder */

/* LL part: one step size, no parents (hzcc.c:308-328; lossless :269-283) */
DSVCU_KERNEL void __launch_bounds__(256)
k_quant_ll(QuantJob J)
{
    const QuantLevel &Q = J.Q[blockIdx.z];
    const int qp = J.lfq[blockIdx.z];
    const int total = Q.w * Q.h;
    for (int k = (int) blockIdx.x * DSVCU_NTH + DSVCU_TID; k < total; k += (int) gridDim.x * DSVCU_NTH) {
        int y = k / Q.w, x = k - y * Q.w;
        int32_t *p = Q.coefs + y * Q.fw + x;
        int c = *p, v;
        if (k == 0) { /* DC travels separately (hzcc.c:265, :599-602) */
            Q.qv[0] = 0;
            continue;
        }
        if (Q.lossless) {
            Q.qv[k] = c;
            continue;
        }
        v = Q.isP ? (c / qp) : q_sub(c, qp, -(qp / 6));
        *p = v ? (Q.isP ? q_deq_d(v, qp) : q_deq_s(v, qp)) : 0;
        Q.qv[k] = v;
    }
}

DSVCU_DEV void
quant_hf_one(const QuantLevel &Q, const QuantBand &B, int x, int y)
{
    int32_t *p = Q.coefs + (B.oy + y) * Q.fw + B.ox + x;
    int c = *p, v;
    int scan = B.scan_base + y * Q.w + x;
    if (Q.lossless) {
        Q.qv[scan] = c;
        return;
    }
    int bidx = ((y * Q.dby) >> 14) * Q.nbh + ((x * Q.dbx) >> 14);
    int flags = Q.blockdata[bidx];
    int parc = Q.coefs[(B.poy + (y >> 1)) * Q.fw + B.pox + (x >> 1)];
    int tmq = B.qp;
    if (Q.isP) {
        tmq = q_tmq_p(tmq, flags, parc);
        if (Q.psy) {
            int gparc = Q.coefs[(B.goy + (y >> 2)) * Q.fw + B.gox + (x >> 2)];
            const dsvcu_mv *mv = Q.mvs + bidx;
            int ax = mv->x < 0 ? -mv->x : mv->x, ay = mv->y < 0 ? -mv->y : mv->y;
            if ((!gparc && !parc) || (mv->flags & MVF_EPRM) ||
                ((mv->flags & MVF_MAINTAIN) && ax < 32 && ay < 32)) {
                v = q_sub(c, tmq, tmq >> 3);
            } else if (!parc || !(flags & BD_SIMCMPLX)) {
                v = q_sub(c, tmq, tmq / 6);
            } else {
                v = q_sub(c, tmq, tmq >> 2);
            }
        } else {
            v = c / tmq;
        }
    } else {
        tmq = q_tmq_i(tmq, flags, parc, Q.l);
        if (Q.psy) {
            if (flags & BD_RING) {
                v = q_sub(c, tmq, -(tmq / 6));
            } else if (Q.l == 0) {
                v = q_sub(c, tmq, -(tmq >> 3));
            } else {
                int smf = flags & (BD_MAINTAIN | BD_STABLE);
                int edge = q_sign(parc) == q_sign(c);
                int stp;
                if (smf == 0) {
                    stp = -tmq / 3;
                } else if (edge && smf == BD_STABLE) {
                    stp = tmq >> 3;
                } else {
                    stp = -tmq / 6;
                }
                v = q_sub(c, tmq, stp);
            }
        } else if (!Q.luma) {
            v = q_sub(c, tmq, -(tmq >> 3));
        } else {
            v = c / tmq;
        }
    }
    *p = v ? q_deq_d(v, tmq) : 0;
    Q.qv[scan] = v;
}

/* wave A: every element of the three rectangles whose parent is final */
DSVCU_KERNEL void __launch_bounds__(256)
k_quant_hf(QuantJob J)
{
    const QuantLevel &Q = J.Q[blockIdx.z];
    const QuantBand B = Q.band[blockIdx.y];
    const int total = Q.w * Q.h;
    for (int k = (int) blockIdx.x * DSVCU_NTH + DSVCU_TID; k < total; k += (int) gridDim.x * DSVCU_NTH) {
        int y = k / Q.w, x = k - y * Q.w;
        if (!Q.lossless && q_parent_aliased(Q, B, x, y)) {
            continue;
        }
        quant_hf_one(Q, B, x, y);
    }
}

/* wave B: last column + last row of each rectangle, only if aliased */
DSVCU_KERNEL void __launch_bounds__(256)
k_quant_hf_edge(QuantJob J)
{
    const QuantLevel &Q = J.Q[blockIdx.z];
    const QuantBand B = Q.band[blockIdx.y];
    const int total = Q.w + Q.h - 1;
    for (int k = (int) blockIdx.x * DSVCU_NTH + DSVCU_TID; k < total; k += (int) gridDim.x * DSVCU_NTH) {
        int x, y;
        if (k < Q.h) {
            x = Q.w - 1;
            y = k;
        } else {
            x = k - Q.h;
            y = Q.h - 1;
        }
        if (q_parent_aliased(Q, B, x, y)) {
            quant_hf_one(Q, B, x, y);
        }
    }
}

/* ---------------------------------------------- ordered stream compaction */

#define CMP_CHUNK 4096
#define CMP_THREADS 256

/* pass 1: non-zero count per chunk of the dense scan-order array */
DSVCU_KERNEL void __launch_bounds__(CMP_THREADS)
k_compact_count(CompactJob J)
{
    DSVCU_SHARED int cnt;
    const int32_t *qv = J.qv[blockIdx.y];
    const int n = J.n[blockIdx.y];
    int *chunk_count = J.chunk[blockIdx.y];
    if ((int) blockIdx.x >= J.nchunks[blockIdx.y]) return;
    int base = (int) blockIdx.x * CMP_CHUNK;
    int lim = min(n, base + CMP_CHUNK);
    int local = 0;
    if (DSVCU_TID == 0) cnt = 0;
    DSVCU_SYNC();
    for (int i = base + DSVCU_TID; i < lim; i += DSVCU_NTH) {
        local += (qv[i] != 0);
    }
    atomicAdd(&cnt, local);
    DSVCU_SYNC();
    if (DSVCU_TID == 0) chunk_count[blockIdx.x] = cnt;
}

/* pass 2: exclusive scan of the chunk counts (single block), total -> *out_n */
DSVCU_KERNEL void __launch_bounds__(1024)
k_compact_scan(CompactJob J)
{
    DSVCU_SHARED int part[1024];
    int *chunk_count = J.chunk[blockIdx.x];
    const int nchunks = J.nchunks[blockIdx.x];
    int *out_n = J.out_n[blockIdx.x];
    int per = (nchunks + DSVCU_NTH - 1) / DSVCU_NTH;
    int b = DSVCU_TID * per, e = min(nchunks, b + per);
    int s = 0;
    for (int i = b; i < e; i++) s += chunk_count[i];
    part[DSVCU_TID] = s;
    DSVCU_SYNC();
    if (DSVCU_TID == 0) {
        int acc = 0;
        for (int i = 0; i < DSVCU_NTH; i++) {
            int v = part[i];
            part[i] = acc;
            acc += v;
        }
        *out_n = acc;
        *J.dc_dst[blockIdx.x] = *J.dc_src[blockIdx.x]; /* the DC coefficient is sent raw */
    }
    DSVCU_SYNC();
    s = part[DSVCU_TID];
    for (int i = b; i < e; i++) {
        int v = chunk_count[i];
        chunk_count[i] = s;
        s += v;
    }
}

/* pass 3: ordered scatter.  Each thread owns a contiguous run of its chunk and
 * writes its non-zeros into a shared-memory image of the chunk's output, which
 * the CTA then copies out with consecutive threads writing consecutive
 * symbols: the destination is pinned HOST memory, where scattered 8-byte
 * stores would each become their own PCIe write. */
DSVCU_KERNEL void __launch_bounds__(CMP_THREADS)
k_compact_scatter(CompactJob J)
{
    DSVCU_SHARED int part[CMP_THREADS];
    DSVCU_SHARED dsvcu_sym stage[CMP_CHUNK];
    DSVCU_SHARED int total;
    const int32_t *qv = J.qv[blockIdx.y];
    const int n = J.n[blockIdx.y];
    const int *chunk_off = J.chunk[blockIdx.y];
    dsvcu_sym *out = J.out[blockIdx.y];
    if ((int) blockIdx.x >= J.nchunks[blockIdx.y]) return;
    int base = (int) blockIdx.x * CMP_CHUNK;
    int per = CMP_CHUNK / DSVCU_NTH;
    int b = base + DSVCU_TID * per, e = min(n, b + per);
    int c = 0;
    for (int i = b; i < e; i++) c += (qv[i] != 0);
    part[DSVCU_TID] = c;
    DSVCU_SYNC();
    if (DSVCU_TID == 0) {
        int acc = 0;
        for (int i = 0; i < DSVCU_NTH; i++) {
            int v = part[i];
            part[i] = acc;
            acc += v;
        }
        total = acc;
    }
    DSVCU_SYNC();
    int o = part[DSVCU_TID];
    for (int i = b; i < e; i++) {
        int v = qv[i];
        if (v) {
            stage[o].pos = (uint32_t) i;
            stage[o].v = v;
            o++;
        }
    }
    DSVCU_SYNC();
    {
        dsvcu_sym *dst = out + chunk_off[blockIdx.x];
        const int cnt = total;
        for (int k = DSVCU_TID; k < cnt; k += DSVCU_NTH) dst[k] = stage[k];
    }
}

/* ------------------------------------------------------------- decoder */

/* LL part (hzcc.c:520-533; lossless :479-492) */
DSVCU_DEV void
q_dequant_ll_one(const QuantLevel &Q, int qp, dsvcu_sym sy)
{
    int y = (int) sy.pos / Q.w, x = (int) sy.pos - y * Q.w;
    int v = sy.v;
    if (!Q.lossless) {
        v = Q.isP ? q_deq_d(v, qp) : q_deq_s(v, qp);
    }
    Q.coefs[y * Q.fw + x] = v;
}

/* levels 0..2 (hzcc.c:534-581).  wave 0 skips aliased-parent elements, wave 1
 * handles only those. */
DSVCU_DEV void
q_dequant_hf_one(const QuantLevel &Q, int wave, dsvcu_sym sy)
{
    const int area = Q.w * Q.h;
    int rel = (int) sy.pos - Q.band[0].scan_base;
    int s = rel / area;
    rel -= s * area;
    const QuantBand B = Q.band[s];
    int y = rel / Q.w, x = rel - y * Q.w;
    int v = sy.v;
    if (!Q.lossless) {
        int al = q_parent_aliased(Q, B, x, y);
        if (al != wave) {
            return;
        }
        int flags = Q.blockdata[((y * Q.dby) >> 14) * Q.nbh + ((x * Q.dbx) >> 14)];
        int parc = Q.coefs[(B.poy + (y >> 1)) * Q.fw + B.pox + (x >> 1)];
        int tmq = Q.isP ? q_tmq_p(B.qp, flags, parc) : q_tmq_i(B.qp, flags, parc, Q.l);
        v = q_deq_d(v, tmq);
    } else if (wave) {
        return;
    }
    Q.coefs[(B.oy + y) * Q.fw + B.ox + x] = v;
}

/* one plane, symbol range given by the host (symbols parsed on the host) */
DSVCU_KERNEL void __launch_bounds__(256)
k_dequant_ll(QuantLevel Q, int qp)
{
    for (int k = Q.sym_begin + (int) blockIdx.x * DSVCU_NTH + DSVCU_TID; k < Q.sym_end; k += (int) gridDim.x * DSVCU_NTH) {
        q_dequant_ll_one(Q, qp, Q.syms[k]);
    }
}

DSVCU_KERNEL void __launch_bounds__(256)
k_dequant_hf(QuantLevel Q)
{
    for (int k = Q.sym_begin + (int) blockIdx.x * DSVCU_NTH + DSVCU_TID; k < Q.sym_end; k += (int) gridDim.x * DSVCU_NTH) {
        q_dequant_hf_one(Q, Q.wave, Q.syms[k]);
    }
}

/* The planes of a picture whose symbols were parsed on the device (k_hzcc.cuh): blockIdx.y
 * selects the plane, the symbol ranges of the levels and the DC are read from the parser's
 * meta words {count, level_start[5], dc, ok} -- the host never sees them. */
struct DequantJob {
    QuantLevel Q[3];
    int lfq[3];
    const int *meta[3];
};

DSVCU_KERNEL void __launch_bounds__(256)
k_dequant_ll_m(DequantJob J)
{
    const QuantLevel &Q = J.Q[blockIdx.y];
    const int *meta = J.meta[blockIdx.y];
    const int begin = meta[1], end = meta[2], qp = J.lfq[blockIdx.y];
    for (int k = begin + (int) blockIdx.x * DSVCU_NTH + DSVCU_TID; k < end; k += (int) gridDim.x * DSVCU_NTH) {
        q_dequant_ll_one(Q, qp, Q.syms[k]);
    }
    if (blockIdx.x == 0 && DSVCU_TID == 0) {
        Q.coefs[0] = meta[6]; /* dst->data[0] = LL (hzcc.c:634); no symbol sits at position 0 */
    }
}

DSVCU_KERNEL void __launch_bounds__(256)
k_dequant_hf_m(DequantJob J, int wave)
{
    const QuantLevel &Q = J.Q[blockIdx.y];
    const int *meta = J.meta[blockIdx.y];
    const int begin = meta[2 + Q.l], end = meta[3 + Q.l];
    for (int k = begin + (int) blockIdx.x * DSVCU_NTH + DSVCU_TID; k < end; k += (int) gridDim.x * DSVCU_NTH) {
        q_dequant_hf_one(Q, wave, Q.syms[k]);
    }
}

#endif /* K_QUANT_CUH */
