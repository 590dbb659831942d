/*
 * dsv_session.h -- whole-stream entry points of the B200 build (plain C ABI).
 *
 * One call = what one `dsv2 e ...` / `dsv2 d ...` process of the reference does
 * (reference src/dsv_main.c:547-905 encode(), :959-1120 decode()), on buffers
 * instead of files, plus the closed-GOP sharded form of the reference's
 * parallel_encode_yuv.sh (:31-52): N fresh encoder instances on consecutive
 * chunks, outputs concatenated in order -- here as host threads that each own a
 * CUDA context on one of the visible GPUs instead of N processes.
 *
 * Frames cross this boundary as tightly packed planar YUV (Y, U, V; the raw
 * .yuv layout the reference CLI reads and writes).
 */
#ifndef DSV2_B200_DSV_SESSION_H
#define DSV2_B200_DSV_SESSION_H

#include <stddef.h>
#include <stdint.h>
#include "dsv.h"
#include "dsv_encoder.h"

#ifdef __cplusplus
extern "C" {
#endif

/* encoder options: the reference CLI's parameter table (dsv_main.c:111-233),
 * same names, same defaults, same units */
typedef struct {
    int w, h, fmt; /* fmt = DSV_SUBSAMP_* */
    int fps_num, fps_den, aspect_num, aspect_den;
    int qp;        /* 0..100 percent, 100 = lossless, -1 = default (85) */
    int effort;    /* 0..10 */
    int gop;       /* -1 = frame rate, 0 = intra only */
    int rc_mode;   /* 0 CRF, 1 ABR, 2 CQP */
    int rc_pergop;
    int kbps;      /* ABR only; 0 = estimate from qp */
    int minqstep, maxqstep;
    int minqp, maxqp, iminqp; /* percent, -1 = auto */
    int stabref, scd, tempaq, bszx, bszy, scpct, skipthresh, varint, psy, dib;
    int ifilter, pfilter, psharp, ipct, pyrlevels;
    int noeos;
} dsv_enc_opts;

void dsv_enc_opts_default(dsv_enc_opts *o, int w, int h, int fmt, int fps_num, int fps_den);
/* dsv_enc_init + the CLI's way of turning the option table into a DSV_ENCODER
 * configuration (dsv_main.c:573-723); dsv_enc_start / dsv_enc then work as usual */
void dsv_enc_configure(DSV_ENCODER *enc, const dsv_enc_opts *o);

/* the CUDA device used by encoder / decoder instances created by the calling
 * thread from now on (default: $DSV_CUDA_DEVICE or 0) */
void dsv_set_thread_device(int device);
int dsv_get_thread_device(void);

/* pinned host memory for frame / stream buffers (optional; any memory works) */
void *dsv_pinned_alloc(size_t bytes);
void dsv_pinned_free(void *p);

/* `dsv2 e`: encodes nframes pictures.  `exhausted` != 0 says the input ended
 * with these frames (the reference then appends an EOS packet even with
 * -noeos=1, dsv_main.c:797).  *out is malloc'ed; the caller frees it. */
int dsv_encode_buffer(const dsv_enc_opts *o, const uint8_t *yuv, int nframes, int exhausted, uint8_t **out,
                      size_t *out_len);

/* parallel_encode_yuv.sh: chunks of `chunk` frames, each coded by a fresh
 * encoder with -noeos=1 semantics, by `nthreads` workers spread round-robin
 * over `ndevices` GPUs (devices[] lists them; NULL = 0..ndevices-1); chunk k's
 * bytes land at their place in the concatenation. */
int dsv_encode_sharded(const dsv_enc_opts *o, const uint8_t *yuv, int nframes, int chunk, int nthreads,
                       const int *devices, int ndevices, uint8_t **out, size_t *out_len);

/* persistent worker pool: `nthreads` host threads spread round-robin over the
 * listed GPUs, each keeping its CUDA context objects between calls.  The
 * one-shot dsv_encode_sharded / dsv_decode_sharded build a pool per call. */
typedef struct dsv_pool dsv_pool;
dsv_pool *dsv_pool_create(int nthreads, const int *devices, int ndevices);
void dsv_pool_destroy(dsv_pool *pool);
int dsv_pool_threads(dsv_pool *pool);
/* `yuv` may be host, pinned or DEVICE memory (unified addressing) */
int dsv_pool_encode(dsv_pool *pool, const dsv_enc_opts *o, const uint8_t *yuv, int nframes, int chunk, uint8_t **out,
                    size_t *out_len);
/* frames are written to caller memory `dst` (host, pinned or DEVICE) */
int dsv_pool_decode(dsv_pool *pool, const uint8_t *dsv, size_t len, uint8_t *dst, size_t dst_cap, int *nframes,
                    DSV_META *meta);

/* Where the whole-stream decoders (dsv_pool_decode*, dsv_decode_*) entropy-decode the
 * coefficient planes: 1 (default; any negative value restores it) = on the device, in batches
 * of pictures one batch ahead of the reconstruction (csrc/k_hzcc.cuh; long pictures and planes
 * the device parser does not accept stay on the host, see dsv_set_device_entropy_limits);
 * 0 = all of them on the host threads.  Same output either way.  Process-wide; returns the
 * previous setting. */
int dsv_set_device_entropy_decode(int on);
/* Which pictures of a batch the device parser takes (host/dsv_dec.c, preparse_classify).  By
 * default a small model decides from the picture sizes whether a picture's chain is over by
 * the time the decoder reaches it.  With early_bytes > 0 fixed sizes are used instead: packets
 * up to early_bytes are parsed together and wanted first; longer ones up to late_bytes form a
 * second part that only they wait for, unless they are among the first late_from pictures of
 * the batch; everything else stays on the host.  early_bytes <= 0 restores the model.
 * Process-wide. */
void dsv_set_device_entropy_limits(long early_bytes, long late_bytes, int late_from);

/* the same with the frames in pinned memory allocated by the call (dsv_pinned_free) */
int dsv_pool_decode_alloc(dsv_pool *pool, const uint8_t *dsv, size_t len, uint8_t **yuv, size_t *yuv_len, int *nframes,
                          DSV_META *meta);

/* `dsv2 d`: decodes a stream into packed frames.  *yuv is malloc'ed (or pinned
 * when `pinned` != 0: free with dsv_pinned_free). */
int dsv_decode_buffer(const uint8_t *dsv, size_t len, int pinned, uint8_t **yuv, size_t *yuv_len, int *nframes,
                      DSV_META *meta);

/* closed-GOP sharded decode: the stream is cut at metadata packets (every
 * chunk of a sharded encode starts with one), segments decoded by `nthreads`
 * workers over `ndevices` GPUs, frames written in stream order. */
int dsv_decode_sharded(const uint8_t *dsv, size_t len, int nthreads, const int *devices, int ndevices, int pinned,
                       uint8_t **yuv, size_t *yuv_len, int *nframes, DSV_META *meta);

#ifdef __cplusplus
}
#endif
#endif
