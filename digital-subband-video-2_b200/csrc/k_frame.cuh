/*
 * k_frame.cuh -- frame helpers the pixel path needs on the device.
 *
 * Replaces reference src/frame.c: extend_plane + downsample_strip (:250-410,
 * the 32-px border whose pixels are 4-sample averages of the nearest edge),
 * dsv_ds2x_frame_luma (:210-234).  dsv_frame_copy (:185-207) is a 2-D device
 * copy followed by k_extend.
 */
#ifndef K_FRAME_CUH
#define K_FRAME_CUH

#include "dsvcu_rt.h"

#define FR_BORDER 32

struct ExtPlane {
    uint8_t *data; /* pixel (0,0) */
    int stride, w, h;
};

struct ExtArgs {
    ExtPlane pl[3];
};

/* average of group g (4 samples, or the remainder group) along an edge:
 * downsample_strip, frame.c:250-355.  p = first sample, d = pitch, n = length */
DSVCU_DEV int
fr_strip(const uint8_t *p, int d, int n, int g)
{
    int len = n & ~3, rem = n & 3;
    if (g * 4 < len) {
        const uint8_t *q = p + (size_t) (g * 4) * d;
        return (q[0] + q[d] + q[2 * d] + q[3 * d] + 2) >> 2;
    }
    int sum = 0;
    for (int i = 0; i < rem; i++) {
        sum += p[(size_t) (len + i) * d];
    }
    return rem ? sum / rem : 0;
}

/* work items per plane: h rows (left+right), ceil(w/4) column groups
 * (top+bottom), 4 corners */
DSVCU_KERNEL void __launch_bounds__(256)
k_extend(ExtArgs A)
{
    const ExtPlane P = A.pl[blockIdx.y];
    const int w = P.w, h = P.h, s = P.stride;
    const int ngc = (w + 3) / 4;
    const int total = h + ngc + 4;
    for (int k = (int) blockIdx.x * DSVCU_NTH + DSVCU_TID; k < total; k += (int) gridDim.x * DSVCU_NTH) {
        if (k < h) {
            int j = k;
            int l = fr_strip(P.data, s, h, j / 4);
            int r = fr_strip(P.data + (w - 1), s, h, j / 4);
            uint8_t *line = P.data + (size_t) j * s;
            for (int i = 0; i < FR_BORDER; i++) {
                line[i - FR_BORDER] = (uint8_t) l;
                line[w + i] = (uint8_t) r;
            }
        } else if (k < h + ngc) {
            int g = k - h;
            int t = fr_strip(P.data, 1, w, g);
            int b = fr_strip(P.data + (size_t) (h - 1) * s, 1, w, g);
            int x0 = g * 4, x1 = min(w, x0 + 4);
            for (int j = 0; j < FR_BORDER; j++) {
                uint8_t *top = P.data - (size_t) (j + 1) * s;
                uint8_t *bot = P.data + (size_t) (h + j) * s;
                for (int x = x0; x < x1; x++) {
                    top[x] = (uint8_t) t;
                    bot[x] = (uint8_t) b;
                }
            }
        } else {
            /* corners average the two adjacent strip ends (frame.c:377-380);
             * the right/bottom ends use the last FULL group */
            int cidx = k - h - ngc;
            int ts0 = fr_strip(P.data, 1, w, 0), ts1 = fr_strip(P.data, 1, w, w / 4 - 1);
            int bs0 = fr_strip(P.data + (size_t) (h - 1) * s, 1, w, 0);
            int bs1 = fr_strip(P.data + (size_t) (h - 1) * s, 1, w, w / 4 - 1);
            int ls0 = fr_strip(P.data, s, h, 0), ls1 = fr_strip(P.data, s, h, h / 4 - 1);
            int rs0 = fr_strip(P.data + (w - 1), s, h, 0), rs1 = fr_strip(P.data + (w - 1), s, h, h / 4 - 1);
            int v, cx, cy;
            if (cidx == 0) {
                v = (ts0 + ls0 + 1) >> 1; cx = -FR_BORDER; cy = -FR_BORDER;
            } else if (cidx == 1) {
                v = (ts1 + rs0 + 1) >> 1; cx = w; cy = -FR_BORDER;
            } else if (cidx == 2) {
                v = (ls1 + bs0 + 1) >> 1; cx = -FR_BORDER; cy = h;
            } else {
                v = (bs1 + rs1 + 1) >> 1; cx = w; cy = h;
            }
            for (int j = 0; j < FR_BORDER; j++) {
                uint8_t *o = P.data + (ptrdiff_t) (cy + j) * s + cx;
                for (int i = 0; i < FR_BORDER; i++) {
                    o[i] = (uint8_t) v;
                }
            }
        }
    }
}

/* 2x2 box downsample of luma (frame.c:210-234); reads the source border when
 * the source height/width is odd */
DSVCU_KERNEL void __launch_bounds__(256)
k_ds2x(uint8_t *dst, int ds, int dw, int dh, const uint8_t *src, int ss)
{
    const int total = dw * dh;
    for (int k = (int) blockIdx.x * DSVCU_NTH + DSVCU_TID; k < total; k += (int) gridDim.x * DSVCU_NTH) {
        int j = k / dw, i = k - j * dw;
        const uint8_t *sp = src + (size_t) (2 * j) * ss + 2 * i;
        dst[(size_t) j * ds + i] = (uint8_t) ((sp[0] + sp[1] + sp[ss] + sp[ss + 1] + 2) >> 2);
    }
}

#endif /* K_FRAME_CUH */
