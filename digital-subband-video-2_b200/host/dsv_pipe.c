/*
 * dsv_pipe.c -- whole-stream drivers (include/dsv_session.h).
 *
 * dsv_encode_buffer / dsv_decode_buffer do what one run of the reference CLI
 * does (src/dsv_main.c:547-905, :959-1120) on memory buffers.  The sharded
 * forms reproduce parallel_encode_yuv.sh (:31-52): independent closed-GOP
 * chunks, each coded by a fresh encoder instance, concatenated in order.  The
 * shell script forks processes; here a chunk is a job taken by one of a set of
 * host threads, and each thread owns one CUDA context (stream, device frames)
 * on one of the GPUs, so several chunks are in flight per GPU and their kernels
 * overlap.  There is no exchange between chunks, hence no collective.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "dsv_host.h"
#include "../../include/dsv_encoder.h"
#include "../../include/dsv_decoder.h"
#include "../../include/dsv_session.h"

/* internal hooks of dsv_enc.c / dsv_dec.c */
void dsv_enc_recycle(DSV_ENCODER *from, DSV_ENCODER *to);
void dsv_dec_direct_output(uint8_t *dst);
void dsv_dec_set_async(int on);
int dsv_dec_flush(DSV_DECODER *d);
int dsv_dec_preparse(DSV_DECODER *d, const uint8_t *const *pkt, const size_t *len, int n);
void dsv_dec_use_parsed(int set, int idx);
int dsv_dec_preparse_pending(DSV_DECODER *d, int set);

static __thread int tls_device = -1;

void
dsv_set_thread_device(int device)
{
    tls_device = device;
}

int
dsv_get_thread_device(void)
{
    if (tls_device < 0) {
        const char *e = getenv("DSV_CUDA_DEVICE");
        return e ? atoi(e) : 0;
    }
    return tls_device;
}

void *
dsv_pinned_alloc(size_t bytes)
{
    return dsvcu_host_alloc(bytes);
}

void
dsv_pinned_free(void *p)
{
    dsvcu_host_free(p);
}

void
dsv_enc_opts_default(dsv_enc_opts *o, int w, int h, int fmt, int fps_num, int fps_den)
{
    memset(o, 0, sizeof(*o));
    o->w = w;
    o->h = h;
    o->fmt = fmt;
    o->fps_num = fps_num;
    o->fps_den = fps_den;
    o->aspect_num = o->aspect_den = 1;
    o->qp = -1;
    o->effort = DSV_MAX_EFFORT;
    o->gop = -1;
    o->rc_mode = DSV_RATE_CONTROL_CRF;
    o->minqstep = DSV_USER_QUAL_TO_RC_QUAL(1) / 2;
    o->maxqstep = DSV_USER_QUAL_TO_RC_QUAL(1) / 4;
    o->minqp = o->maxqp = o->iminqp = -1;
    o->scd = 1;
    o->tempaq = 1;
    o->bszx = o->bszy = -1;
    o->scpct = 85;
    o->varint = 1;
    o->psy = DSV_PSY_ALL;
    o->dib = 1;
    o->ifilter = 1;
    o->pfilter = -1;
    o->psharp = 1;
    o->ipct = 90;
}

/* heuristic bytes/s for a quality percent (reference util.c:21-57) */
static unsigned
guess_bitrate(int quality, int gop, const DSV_META *md)
{
    int fps = (md->fps_num + md->fps_den / 2) / md->fps_den;
    int bpf, scale;
    switch (md->subsamp) {
        case DSV_SUBSAMP_422:
        case DSV_SUBSAMP_UYVY: bpf = 352 * 288 * 2; break;
        case DSV_SUBSAMP_420:
        case DSV_SUBSAMP_411: bpf = 352 * 288 * 3 / 2; break;
        case DSV_SUBSAMP_410: bpf = 352 * 288 * 9 / 8; break;
        default: bpf = 352 * 288 * 3; break;
    }
    if (gop == DSV_GOP_INTRA) {
        bpf *= 4;
    }
    if (md->width < 320 && md->height < 240) {
        bpf /= 4;
    }
    scale = (((md->width + md->height) / 2) << 8) / 352;
    bpf = bpf * scale >> 8;
    return (unsigned) ((bpf * fps) / (26 - quality / 4)) * 3 / 2;
}

static int
guess_quality(int bps, int gop, const DSV_META *md) /* util.c:59-76 */
{
    int q, bestq = 50, best = INT_MAX;
    for (q = 0; q < 100; q++) {
        int dif = abs((int) guess_bitrate(q, gop, md) - bps);
        if (dif < best) {
            bestq = q;
            best = dif;
        }
    }
    return CLAMP(bestq, 0, 99);
}

static int
pct_or_auto(int pct)
{
    return pct < 0 ? -1 : DSV_USER_QUAL_TO_RC_QUAL(pct);
}

/* option table -> encoder configuration, as the CLI does it
 * (reference dsv_main.c:573-723) */
void
dsv_enc_configure(DSV_ENCODER *enc, const dsv_enc_opts *o)
{
    DSV_META md;
    int fps, bps;

    dsv_enc_init(enc);
    memset(&md, 0, sizeof(md));
    md.width = o->w;
    md.height = o->h;
    md.subsamp = o->fmt;
    md.fps_num = o->fps_num;
    md.fps_den = o->fps_den > 0 ? o->fps_den : 1;
    md.aspect_num = o->aspect_num;
    md.aspect_den = o->aspect_den;
    md.inter_sharpen = o->psharp;
    fps = (md.fps_num + md.fps_den / 2) / md.fps_den;
    if (fps <= 0) {
        md.fps_num = md.fps_den = 1;
        fps = 1;
    }
    dsv_enc_set_metadata(enc, &md);

    enc->gop = o->gop < 0 ? fps : o->gop;
    enc->scene_change_pct = o->scpct;
    enc->do_scd = o->scd;
    enc->intra_pct_thresh = o->ipct;
    enc->skip_block_thresh = o->skipthresh;
    enc->rc_mode = o->rc_mode;
    enc->rc_pergop = o->rc_pergop;
    bps = o->kbps * 1024;
    if (o->qp < 0) {
        int pct = (enc->rc_mode != DSV_RATE_CONTROL_ABR || bps == 0) ? 85 : guess_quality(bps, enc->gop, &md);
        enc->quality = DSV_USER_QUAL_TO_RC_QUAL(pct);
    } else {
        enc->quality = DSV_USER_QUAL_TO_RC_QUAL(o->qp);
    }
    enc->bitrate = bps ? (unsigned) bps : guess_bitrate(enc->quality * 100 / DSV_RC_QUAL_MAX, enc->gop, &md);
    enc->min_q_step = o->minqstep;
    enc->max_q_step = o->maxqstep;
    enc->min_quality = pct_or_auto(o->minqp);
    enc->max_quality = pct_or_auto(o->maxqp);
    enc->min_I_frame_quality = pct_or_auto(o->iminqp);
    if (enc->rc_mode == DSV_RATE_CONTROL_CRF) {
        if (enc->min_quality < 0) enc->min_quality = enc->quality - DSV_USER_QUAL_TO_RC_QUAL(5);
        if (enc->min_I_frame_quality < 0) enc->min_I_frame_quality = enc->quality - DSV_USER_QUAL_TO_RC_QUAL(2);
    } else {
        if (enc->min_quality < 0) enc->min_quality = 0;
        if (enc->min_I_frame_quality < 0) enc->min_I_frame_quality = DSV_USER_QUAL_TO_RC_QUAL(5);
    }
    if (enc->max_quality < 0) {
        enc->max_quality = DSV_RC_QUAL_MAX;
    }
    enc->min_quality = CLAMP(enc->min_quality, 0, DSV_RC_QUAL_MAX);
    enc->min_I_frame_quality = CLAMP(enc->min_I_frame_quality, 0, DSV_RC_QUAL_MAX);
    enc->max_quality = CLAMP(enc->max_quality, 0, DSV_RC_QUAL_MAX);
    enc->pyramid_levels = o->pyrlevels;
    enc->stable_refresh = o->stabref ? (unsigned) o->stabref : (unsigned) CLAMP(fps, 1, 60);
    enc->do_temporal_aq = o->tempaq;
    enc->variable_i_interval = o->varint;
    enc->block_size_override_x = o->bszx;
    enc->block_size_override_y = o->bszy;
    enc->effort = o->effort;
    enc->do_psy = o->psy;
    enc->do_dark_intra_boost = o->dib;
    enc->do_intra_filter = o->ifilter;
    enc->do_inter_filter = o->pfilter;
}

static size_t
frame_bytes(int w, int h, int fmt)
{
    size_t cw = (size_t) DSV_ROUND_SHIFT(w, DSV_FORMAT_H_SHIFT(fmt));
    size_t ch = (size_t) DSV_ROUND_SHIFT(h, DSV_FORMAT_V_SHIFT(fmt));
    return (size_t) w * h + 2 * cw * ch;
}

typedef struct {
    uint8_t *data;
    size_t len, cap;
} BYTES;

static int
bytes_append(BYTES *b, const uint8_t *p, size_t n)
{
    if (b->len + n > b->cap) {
        size_t ncap = b->cap ? b->cap * 2 : (1 << 16);
        uint8_t *nd;
        while (ncap < b->len + n) {
            ncap *= 2;
        }
        nd = realloc(b->data, ncap);
        if (!nd) {
            return -1;
        }
        b->data = nd;
        b->cap = ncap;
    }
    memcpy(b->data + b->len, p, n);
    b->len += n;
    return 0;
}

/* one encoder instance over `nframes` pictures -> `out` (appended) */
static int
run_encoder(DSV_ENCODER *enc, const dsv_enc_opts *o, const uint8_t *yuv, int nframes, int write_eos, BYTES *out)
{
    size_t fsz = frame_bytes(o->w, o->h, o->fmt);
    DSV_BUF bufs[4];
    int f, i, n;

    dsv_enc_start(enc);
    for (f = 0; f < nframes; f++) {
        DSV_FRAME *fr = dsv_load_planar_frame(o->fmt, (void *) (yuv + (size_t) f * fsz), o->w, o->h);
        n = dsv_enc(enc, fr, bufs) & DSV_ENC_NUM_BUFS;
        if (n == 0) {
            return -1;
        }
        for (i = 0; i < n; i++) {
            int r = bytes_append(out, bufs[i].data, bufs[i].len);
            dsv_buf_free(&bufs[i]);
            if (r) {
                return -1;
            }
        }
    }
    if (write_eos) {
        dsv_enc_end_of_stream(enc, bufs);
        bytes_append(out, bufs[0].data, bufs[0].len);
        dsv_buf_free(&bufs[0]);
    }
    return 0;
}

int
dsv_encode_buffer(const dsv_enc_opts *o, const uint8_t *yuv, int nframes, int exhausted, uint8_t **out, size_t *out_len)
{
    DSV_ENCODER enc;
    BYTES b;
    int r;
    memset(&b, 0, sizeof(b));
    dsv_enc_configure(&enc, o);
    r = run_encoder(&enc, o, yuv, nframes, !o->noeos || (exhausted && nframes > 0), &b);
    dsv_enc_free(&enc);
    if (r) {
        free(b.data);
        return -1;
    }
    *out = b.data;
    *out_len = b.len;
    return 0;
}

/* ------------------------------------------------------------ worker pool */

/* A pool is a set of persistent host threads, each bound to one GPU and each
 * keeping its CUDA context objects (stream, device frames, pinned staging)
 * alive between jobs.  A job is "run fn(k) for k in [0, n)"; workers take
 * indices from a shared counter, so chunks/segments balance themselves. */
typedef struct WORKER WORKER;
typedef void (*job_fn)(WORKER *w, void *arg, int k);

struct dsv_pool {
    int nthreads;
    WORKER *wk;
    pthread_mutex_t lock;
    pthread_cond_t wake, done;
    job_fn fn;
    void *arg;
    int n, next, running, generation, quit;
};

struct WORKER {
    struct dsv_pool *pool;
    pthread_t th;
    int device;
    int seen; /* last generation served */
    DSV_ENCODER enc_keep; /* finished encoder whose device buffers get recycled */
    int have_enc;
    DSV_DECODER dec;
};

static void *
pool_worker(void *p)
{
    WORKER *w = p;
    struct dsv_pool *pl = w->pool;
    dsv_set_thread_device(w->device);
    pthread_mutex_lock(&pl->lock);
    for (;;) {
        while (!pl->quit && (pl->generation == w->seen || pl->next >= pl->n)) {
            if (pl->generation != w->seen) {
                w->seen = pl->generation; /* nothing left for me in this job */
            }
            pthread_cond_wait(&pl->wake, &pl->lock);
        }
        if (pl->quit) {
            break;
        }
        while (pl->next < pl->n) {
            int k = pl->next++;
            pl->running++;
            pthread_mutex_unlock(&pl->lock);
            pl->fn(w, pl->arg, k);
            pthread_mutex_lock(&pl->lock);
            pl->running--;
        }
        w->seen = pl->generation;
        if (pl->running == 0) {
            pthread_cond_broadcast(&pl->done);
        }
    }
    pthread_mutex_unlock(&pl->lock);
    if (w->have_enc) {
        dsv_enc_free(&w->enc_keep);
    }
    dsv_dec_free(&w->dec);
    return NULL;
}

dsv_pool *
dsv_pool_create(int nthreads, const int *devices, int ndevices)
{
    struct dsv_pool *pl;
    int i;
    if (nthreads <= 0) {
        return NULL;
    }
    if (ndevices <= 0) {
        ndevices = 1;
    }
    pl = calloc(1, sizeof(*pl));
    pl->wk = calloc((size_t) nthreads, sizeof(WORKER));
    pthread_mutex_init(&pl->lock, NULL);
    pthread_cond_init(&pl->wake, NULL);
    pthread_cond_init(&pl->done, NULL);
    for (i = 0; i < nthreads; i++) {
        pl->wk[i].pool = pl;
        pl->wk[i].device = devices ? devices[i % ndevices] : (ndevices > 1 ? i % ndevices : dsv_get_thread_device());
        if (pthread_create(&pl->wk[i].th, NULL, pool_worker, &pl->wk[i])) {
            break;
        }
        pl->nthreads++;
    }
    if (pl->nthreads == 0) {
        free(pl->wk);
        free(pl);
        return NULL;
    }
    return pl;
}

void
dsv_pool_destroy(dsv_pool *pl)
{
    int i;
    if (!pl) {
        return;
    }
    pthread_mutex_lock(&pl->lock);
    pl->quit = 1;
    pthread_cond_broadcast(&pl->wake);
    pthread_mutex_unlock(&pl->lock);
    for (i = 0; i < pl->nthreads; i++) {
        pthread_join(pl->wk[i].th, NULL);
    }
    pthread_mutex_destroy(&pl->lock);
    pthread_cond_destroy(&pl->wake);
    pthread_cond_destroy(&pl->done);
    free(pl->wk);
    free(pl);
}

int
dsv_pool_threads(dsv_pool *pl)
{
    return pl ? pl->nthreads : 0;
}

static void
pool_run(struct dsv_pool *pl, job_fn fn, void *arg, int n)
{
    pthread_mutex_lock(&pl->lock);
    pl->fn = fn;
    pl->arg = arg;
    pl->n = n;
    pl->next = 0;
    pl->generation++;
    pthread_cond_broadcast(&pl->wake);
    while (pl->next < pl->n || pl->running > 0) {
        pthread_cond_wait(&pl->done, &pl->lock);
    }
    pthread_mutex_unlock(&pl->lock);
}

/* ----------------------------------------------------------- sharded encode */

typedef struct {
    const dsv_enc_opts *o;
    const uint8_t *yuv;
    int nframes, chunk;
    BYTES *parts;
    int failed;
} ENC_JOB;

static void
encode_chunk(WORKER *w, void *arg, int k)
{
    ENC_JOB *j = arg;
    DSV_ENCODER enc;
    size_t fsz = frame_bytes(j->o->w, j->o->h, j->o->fmt);
    int first = k * j->chunk;
    int n = MIN(j->chunk, j->nframes - first);
    /* a fresh encoder per chunk (frame numbers, rate control and block
     * statistics restart, parallel_encode_yuv.sh:34-41), but the device
     * buffers of the worker's previous encoder are handed over */
    dsv_enc_configure(&enc, j->o);
    if (w->have_enc) {
        dsv_enc_recycle(&w->enc_keep, &enc);
        dsv_enc_free(&w->enc_keep);
        w->have_enc = 0;
    }
    if (run_encoder(&enc, j->o, j->yuv + (size_t) first * fsz, n, 0, &j->parts[k])) {
        j->failed = 1;
    }
    w->enc_keep = enc;
    w->have_enc = 1;
}

int
dsv_pool_encode(dsv_pool *pl, const dsv_enc_opts *o, const uint8_t *yuv, int nframes, int chunk, uint8_t **out,
                size_t *out_len)
{
    ENC_JOB job;
    size_t total = 0, off = 0;
    int i, nchunks;

    *out = NULL;
    *out_len = 0;
    if (!pl || chunk <= 0 || nframes <= 0) {
        return -1;
    }
    memset(&job, 0, sizeof(job));
    nchunks = (nframes + chunk - 1) / chunk;
    job.o = o;
    job.yuv = yuv;
    job.nframes = nframes;
    job.chunk = chunk;
    job.parts = calloc((size_t) nchunks, sizeof(BYTES));
    pool_run(pl, encode_chunk, &job, nchunks);
    for (i = 0; i < nchunks; i++) {
        total += job.parts[i].len;
    }
    if (!job.failed) {
        *out = malloc(total ? total : 1);
        for (i = 0; i < nchunks; i++) {
            memcpy(*out + off, job.parts[i].data, job.parts[i].len);
            off += job.parts[i].len;
        }
        *out_len = total;
    }
    for (i = 0; i < nchunks; i++) {
        free(job.parts[i].data);
    }
    free(job.parts);
    return job.failed ? -1 : 0;
}

int
dsv_encode_sharded(const dsv_enc_opts *o, const uint8_t *yuv, int nframes, int chunk, int nthreads, const int *devices,
                   int ndevices, uint8_t **out, size_t *out_len)
{
    dsv_pool *pl;
    int r, nchunks;
    if (chunk <= 0 || nframes <= 0 || nthreads <= 0) {
        return -1;
    }
    nchunks = (nframes + chunk - 1) / chunk;
    pl = dsv_pool_create(MIN(nthreads, nchunks), devices, ndevices);
    if (!pl) {
        return -1;
    }
    r = dsv_pool_encode(pl, o, yuv, nframes, chunk, out, out_len);
    dsv_pool_destroy(pl);
    return r;
}

/* ------------------------------------------------------------------ decode */

typedef struct {
    size_t off, len; /* packet position in the stream */
    int type;
} PKT;

/* walk the packet chain by the next-link field (reference dsv_main.c:912-957) */
static int
index_packets(const uint8_t *d, size_t len, PKT **out)
{
    PKT *v = NULL;
    int n = 0, cap = 0;
    size_t off = 0;
    while (off + DSV_PACKET_HDR_SIZE <= len) {
        const uint8_t *p = d + off;
        size_t size;
        if (p[0] != DSV_FOURCC_0 || p[1] != DSV_FOURCC_1 || p[2] != DSV_FOURCC_2 || p[3] != DSV_FOURCC_3) {
            break;
        }
        size = ((size_t) p[10] << 24) | ((size_t) p[11] << 16) | ((size_t) p[12] << 8) | p[13];
        if (size == 0) {
            size = DSV_PACKET_HDR_SIZE; /* EOS */
        }
        if (size < DSV_PACKET_HDR_SIZE || off + size > len) {
            break;
        }
        if (n == cap) {
            cap = cap ? cap * 2 : 256;
            v = realloc(v, (size_t) cap * sizeof(PKT));
        }
        v[n].off = off;
        v[n].len = size;
        v[n].type = p[DSV_PACKET_TYPE_OFFSET];
        n++;
        off += size;
    }
    *out = v;
    return n;
}

/* metadata of a stream = its first metadata packet, parsed by the decoder.
 * The whole-stream drivers write every picture into one array of equally sized
 * frames, so EVERY metadata packet is parsed here and a stream whose geometry
 * changes on the way is refused (the reference CLI sizes each frame from the
 * metadata current at that point, dsv_main.c:1008-1060; a fixed-size output
 * cannot represent that, and silently keeping the first size would overrun it) */
static int
probe_meta(const uint8_t *d, const PKT *pk, int npk, DSV_META *meta)
{
    DSV_DECODER dec;
    int i, have = 0;
    memset(&dec, 0, sizeof(dec));
    for (i = 0; i < npk; i++) {
        if (pk[i].type == DSV_PT_META) {
            DSV_BUF b;
            DSV_FRAME *fr;
            DSV_FNUM fn;
            dsv_mk_buf(&b, (int) pk[i].len);
            memcpy(b.data, d + pk[i].off, pk[i].len);
            if (dsv_dec(&dec, &b, &fr, &fn) != DSV_DEC_GOT_META) {
                return -1;
            }
            if (!have) {
                *meta = dec.vidmeta;
                have = 1;
            } else if (dec.vidmeta.width != meta->width || dec.vidmeta.height != meta->height ||
                       dec.vidmeta.subsamp != meta->subsamp) {
                DSV_ERROR(("picture geometry changes inside the stream (%dx%d fmt %d -> %dx%d fmt %d): not supported "
                           "by the whole-stream drivers", meta->width, meta->height, meta->subsamp, dec.vidmeta.width,
                           dec.vidmeta.height, dec.vidmeta.subsamp));
                return -1;
            }
        }
    }
    return have ? 0 : -1;
}

/* pictures per device-side entropy decode (dsv_dec_preparse): a closed GOP or two */
#define PREPARSE_MAX_PICS 64
#define PREPARSE_MAX_BYTES ((size_t) 24 << 20)

/* 1: the coefficient planes of the whole-stream decoders are entropy-decoded on the device
 * (measured on one B200 + 16 cores, 1080p: 1 / 8 / 32 instances 678 / 4716 / 10 514 frames/s
 * against 647 / 4260 / 9575 with the host parser; with 4 cores for 32 instances 5781 against
 * 3545); 0: on the host threads */
static volatile int g_device_entropy = 1;

int
dsv_set_device_entropy_decode(int on)
{
    const int was = g_device_entropy;
    g_device_entropy = on < 0 ? 1 : !!on;
    return was;
}

typedef struct {
    int begin, end; /* packets [begin, end), all of them pictures */
    int slot;       /* where dsv_dec_preparse put them; -1: the host parses */
} BATCH;

/* hand the coefficient planes of the pictures from packet `from` on (up to the next packet
 * that is not a picture) to the device parser */
static void
batch_begin(DSV_DECODER *dec, const uint8_t *d, const PKT *pk, int from, int last, int on_device, BATCH *b)
{
    const uint8_t *bp[PREPARSE_MAX_PICS];
    size_t bl[PREPARSE_MAX_PICS], bytes = 0;
    int j, n = 0;
    for (j = from; j < last && n < PREPARSE_MAX_PICS && DSV_PT_IS_PIC(pk[j].type); j++) {
        if (n && bytes + pk[j].len > PREPARSE_MAX_BYTES) {
            break;
        }
        bp[n] = d + pk[j].off;
        bl[n] = pk[j].len;
        bytes += pk[j].len;
        n++;
    }
    b->begin = from;
    b->end = j;
    /* on a device error the host parser takes over (and will report it) */
    b->slot = (dec->got_metadata && on_device) ? dsv_dec_preparse(dec, bp, bl, n) : -1;
}

/* decode packets [first, last) with `dec`; frames are written to dst one after
 * the other.  returns the number of frames written */
static int
decode_range(DSV_DECODER *dec, const uint8_t *d, const PKT *pk, int first, int last, uint8_t *dst, size_t fsz,
             int on_device)
{
    BATCH cur, nxt;
    int i, nfr = 0, cur_k = 0, have_next = 0;
    cur.begin = cur.end = first;
    cur.slot = nxt.slot = -1;
    /* pictures are queued without waiting: parsing the next packet overlaps the device work */
    dsv_dec_set_async(1);
    for (i = first; i < last; i++) {
        DSV_BUF b;
        DSV_FRAME *fr = NULL;
        DSV_FNUM fn;
        int code, is_pic = DSV_PT_IS_PIC(pk[i].type);
        if (is_pic && i >= cur.end) {
            if (have_next && nxt.begin == i) {
                cur = nxt;
            } else {
                batch_begin(dec, d, pk, i, last, on_device, &cur);
            }
            have_next = 0;
            cur_k = 0;
        }
        dsv_mk_buf(&b, (int) pk[i].len);
        memcpy(b.data, d + pk[i].off, pk[i].len);
        if (is_pic) {
            /* never arm a direct copy for a geometry other than the one dst was sized for */
            if (dec->got_metadata &&
                frame_bytes(dec->vidmeta.width, dec->vidmeta.height, dec->vidmeta.subsamp) != fsz) {
                DSV_ERROR(("picture packet with unexpected geometry: segment abandoned"));
                dsv_buf_free(&b);
                break;
            }
            dsv_dec_direct_output(dst + (size_t) nfr * fsz);
            dsv_dec_use_parsed(cur.slot, cur.slot >= 0 ? cur_k : -1);
            cur_k++;
        }
        code = dsv_dec(dec, &b, &fr, &fn);
        dsv_dec_direct_output(NULL);
        dsv_dec_use_parsed(0, -1);
        if (code == DSV_DEC_EOS) {
            break;
        }
        if (code == DSV_DEC_OK && fr) {
            nfr++;
            dsv_frame_ref_dec(fr);
        }
        /* one batch ahead: as soon as the result of the current batch has been collected, the
         * planes of the pictures behind it go to the device, to be parsed while the current
         * ones are reconstructed */
        if (is_pic && !have_next && cur.end < last && DSV_PT_IS_PIC(pk[cur.end].type) &&
            (cur.slot < 0 || !dsv_dec_preparse_pending(dec, cur.slot))) {
            batch_begin(dec, d, pk, cur.end, last, on_device, &nxt);
            have_next = 1;
        }
    }
    dsv_dec_set_async(0);
    if (dsv_dec_flush(dec)) {
        return 0;
    }
    return nfr;
}

static int
count_pictures(const PKT *pk, int first, int last)
{
    int i, n = 0;
    for (i = first; i < last; i++) {
        n += !!DSV_PT_IS_PIC(pk[i].type);
    }
    return n;
}

typedef struct {
    const uint8_t *d;
    const PKT *pk;
    int *seg_first, *seg_last, *seg_frame0; /* per segment */
    uint8_t *dst;
    size_t fsz;
    int *seg_done; /* frames actually decoded per segment */
    int device_entropy;
} DEC_JOB;

static void
decode_segment(WORKER *w, void *arg, int k)
{
    DEC_JOB *j = arg;
    j->seg_done[k] = decode_range(&w->dec, j->d, j->pk, j->seg_first[k], j->seg_last[k],
                                  j->dst + (size_t) j->seg_frame0[k] * j->fsz, j->fsz, j->device_entropy);
}

/* frames are written to `dst` (host, pinned or DEVICE memory) when it is given
 * and large enough for dst_cap bytes; with dst == NULL *yuv is allocated */
static int
pool_decode(dsv_pool *pl, const uint8_t *dsv, size_t len, uint8_t *dst, size_t dst_cap, int pinned, uint8_t **yuv,
            size_t *yuv_len, int *nframes, DSV_META *meta)
{
    PKT *pk = NULL;
    DEC_JOB job;
    DSV_META md;
    int npk, i, nseg = 0, total, ok = 1, outn = 0;

    if (yuv) {
        *yuv = NULL;
    }
    *yuv_len = 0;
    *nframes = 0;
    npk = index_packets(dsv, len, &pk);
    if (npk <= 0 || probe_meta(dsv, pk, npk, &md)) {
        free(pk);
        return -1;
    }
    if (meta) {
        *meta = md;
    }
    memset(&job, 0, sizeof(job));
    job.seg_first = calloc((size_t) npk + 1, sizeof(int));
    job.seg_last = calloc((size_t) npk + 1, sizeof(int));
    job.seg_frame0 = calloc((size_t) npk + 1, sizeof(int));
    job.seg_done = calloc((size_t) npk + 1, sizeof(int));
    /* a segment starts at every metadata packet: the picture that follows has
     * no reference (closed GOP), so segments decode independently */
    for (i = 0; i < npk; i++) {
        if (pk[i].type == DSV_PT_META || nseg == 0) {
            job.seg_first[nseg] = i;
            if (nseg) {
                job.seg_last[nseg - 1] = i;
            }
            nseg++;
        }
    }
    job.seg_last[nseg - 1] = npk;
    total = 0;
    for (i = 0; i < nseg; i++) {
        job.seg_frame0[i] = total;
        total += count_pictures(pk, job.seg_first[i], job.seg_last[i]);
    }
    job.d = dsv;
    job.pk = pk;
    job.fsz = frame_bytes(md.width, md.height, md.subsamp);
    job.device_entropy = g_device_entropy;
    if (dst) {
        if (dst_cap < job.fsz * (size_t) total) {
            ok = 0;
        }
        job.dst = dst;
    } else {
        job.dst = pinned ? dsv_pinned_alloc(job.fsz * (size_t) MAX(total, 1)) : malloc(job.fsz * (size_t) MAX(total, 1));
        if (!job.dst) {
            ok = 0;
        }
    }
    if (ok) {
        pool_run(pl, decode_segment, &job, nseg);
        /* frames of a damaged segment may be missing: close the gaps (host
         * destinations only; a device destination keeps segment positions) */
        for (i = 0; i < nseg; i++) {
            if (!dst && job.seg_done[i] > 0 && outn != job.seg_frame0[i]) {
                memmove(job.dst + (size_t) outn * job.fsz, job.dst + (size_t) job.seg_frame0[i] * job.fsz,
                        (size_t) job.seg_done[i] * job.fsz);
            }
            outn += job.seg_done[i];
        }
        if (yuv) {
            *yuv = job.dst;
        }
        *yuv_len = (size_t) outn * job.fsz;
        *nframes = outn;
    }
    free(job.seg_first);
    free(job.seg_last);
    free(job.seg_frame0);
    free(job.seg_done);
    free(pk);
    return ok ? 0 : -1;
}

int
dsv_pool_decode(dsv_pool *pl, const uint8_t *dsv, size_t len, uint8_t *dst, size_t dst_cap, int *nframes, DSV_META *meta)
{
    size_t n;
    if (!pl || !dst) {
        return -1;
    }
    return pool_decode(pl, dsv, len, dst, dst_cap, 0, NULL, &n, nframes, meta);
}

/* the same into pinned memory allocated by the call (dsv_pinned_free) */
int
dsv_pool_decode_alloc(dsv_pool *pl, const uint8_t *dsv, size_t len, uint8_t **yuv, size_t *yuv_len, int *nframes,
                      DSV_META *meta)
{
    if (!pl || !yuv) {
        return -1;
    }
    return pool_decode(pl, dsv, len, NULL, 0, 1, yuv, yuv_len, nframes, meta);
}

int
dsv_decode_sharded(const uint8_t *dsv, size_t len, int nthreads, const int *devices, int ndevices, int pinned,
                   uint8_t **yuv, size_t *yuv_len, int *nframes, DSV_META *meta)
{
    dsv_pool *pl = dsv_pool_create(MAX(nthreads, 1), devices, ndevices);
    int r;
    if (!pl) {
        return -1;
    }
    r = pool_decode(pl, dsv, len, NULL, 0, pinned, yuv, yuv_len, nframes, meta);
    dsv_pool_destroy(pl);
    return r;
}

int
dsv_decode_buffer(const uint8_t *dsv, size_t len, int pinned, uint8_t **yuv, size_t *yuv_len, int *nframes, DSV_META *meta)
{
    return dsv_decode_sharded(dsv, len, 1, NULL, 1, pinned, yuv, yuv_len, nframes, meta);
}
