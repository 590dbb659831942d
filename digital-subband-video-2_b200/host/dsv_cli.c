/*
 * dsv_cli.c -- `dsv2cu`: command line of the B200 build.
 *
 * Same verbs, option names, defaults and exit codes as the reference CLI
 * (src/dsv_main.c: `dsv2 e ...` / `dsv2 d ...`, option table :111-247, exit
 * status -2 = "input exhausted" :904), so scripts written for `dsv2` --
 * including parallel_encode_yuv.sh, which relies on -sfr/-nfr/-noeos and on the
 * exit status -- work unchanged.  Extensions (all optional):
 *     -gpus=N -threads=T -chunk=K   closed-GOP sharded encode / decode inside
 *                                   one process (what parallel_encode_yuv.sh
 *                                   does with N processes), K frames per chunk
 *     -dev=D                        CUDA device for the single-instance path
 *
 * Everything streams: the single-instance paths read, code and write one frame /
 * packet at a time like the reference (dsv_main.c:744-792, :1000-1109); the
 * sharded paths work on batches of `threads` chunks (encode) or closed GOPs
 * (decode), so memory is bounded by the batch, not by the length of the clip.
 * Not implemented: stdin / stdout piping.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "dsv_host.h"
#include "../../include/dsv_encoder.h"
#include "../../include/dsv_decoder.h"
#include "../../include/dsv_session.h"

typedef struct {
    const char *name;
    int *dst;
} OPT;

static const char *g_inp = NULL, *g_out = NULL;
static int g_yes = 0, g_verbose = 0;

static int
fmt_to_subsamp(int f)
{
    switch (f) {
        case 0: return DSV_SUBSAMP_444;
        case 1: return DSV_SUBSAMP_422;
        case 2: return DSV_SUBSAMP_420;
        case 3: return DSV_SUBSAMP_411;
        case 4: return DSV_SUBSAMP_410;
        case 5: return DSV_SUBSAMP_UYVY;
        default: return DSV_SUBSAMP_420;
    }
}

static int
parse_common(const char *a)
{
    if (!strncmp(a, "-inp=", 5)) {
        g_inp = a + 5;
    } else if (!strncmp(a, "-out=", 5)) {
        g_out = a + 5;
    } else if (!strcmp(a, "-y")) {
        g_yes = 1;
    } else if (!strcmp(a, "-v")) {
        g_verbose = 1;
    } else if (!strncmp(a, "-l", 2) && a[2] >= '0' && a[2] <= '4' && a[3] == 0) {
        dsv_set_log_level(a[2] - '0');
    } else {
        return 0;
    }
    return 1;
}

static int
parse_opts(int argc, char **argv, const OPT *tab)
{
    int i, k;
    for (i = 2; i < argc; i++) {
        const char *a = argv[i];
        const char *eq;
        if (parse_common(a)) {
            continue;
        }
        eq = strchr(a, '=');
        if (a[0] != '-' || !eq) {
            fprintf(stderr, "unrecognized argument %s\n", a);
            return -1;
        }
        for (k = 0; tab[k].name; k++) {
            size_t n = strlen(tab[k].name);
            if ((size_t) (eq - a - 1) == n && !strncmp(a + 1, tab[k].name, n)) {
                *tab[k].dst = atoi(eq + 1);
                break;
            }
        }
        if (!tab[k].name) {
            fprintf(stderr, "unrecognized option %s\n", a);
            return -1;
        }
    }
    if (!g_inp || !g_out) {
        fprintf(stderr, "need -inp= and -out=\n");
        return -1;
    }
    return 0;
}

static int
confirm_overwrite(const char *path)
{
    FILE *f;
    int c;
    if (g_yes || !(f = fopen(path, "rb"))) {
        return 1;
    }
    fclose(f);
    printf("\n--- file (%s) already exists, overwrite? (y/n) ", path);
    c = getchar();
    return c == 'y' || c == 'Y';
}

static size_t
frame_size(int w, int h, int subsamp)
{
    size_t cw = (size_t) DSV_ROUND_SHIFT(w, DSV_FORMAT_H_SHIFT(subsamp));
    size_t ch = (size_t) DSV_ROUND_SHIFT(h, DSV_FORMAT_V_SHIFT(subsamp));
    return (size_t) w * h + 2 * cw * ch;
}

static int
read_frame(FILE *in, uint8_t *dst, const dsv_enc_opts *o, int y4m)
{
    return y4m ? dsv_y4m_read_frame(in, dst, o->w, o->h, o->fmt) : dsv_yuv_read_seq(in, dst, o->w, o->h, o->fmt);
}

/* ------------------------------------------------------------------ encode */

/* one encoder instance, one frame at a time (reference encode(), dsv_main.c:744-800) */
static int
encode_streaming(const dsv_enc_opts *o, FILE *in, FILE *fo, int y4m, int nfr, int *pgot, int *pexhausted)
{
    DSV_ENCODER enc;
    DSV_BUF bufs[4];
    size_t fsz = frame_size(o->w, o->h, o->fmt);
    uint8_t *pic = malloc(fsz ? fsz : 1);
    size_t total = 0;
    int got = 0, exhausted = 0, i, n, rc = 0;

    if (!pic) {
        return -1;
    }
    dsv_enc_configure(&enc, o);
    dsv_enc_start(&enc);
    while (nfr < 0 || got < nfr) {
        DSV_FRAME *fr;
        if (read_frame(in, pic, o, y4m) < 0) {
            exhausted = 1;
            break;
        }
        fr = dsv_load_planar_frame(o->fmt, pic, o->w, o->h);
        n = dsv_enc(&enc, fr, bufs) & DSV_ENC_NUM_BUFS;
        if (n == 0) {
            rc = -1;
            break;
        }
        for (i = 0; i < n; i++) {
            if (fwrite(bufs[i].data, 1, bufs[i].len, fo) != bufs[i].len) {
                rc = -1;
            }
            total += bufs[i].len;
            dsv_buf_free(&bufs[i]);
        }
        got++;
        if (g_verbose) {
            printf("encoded frame %d\r", got);
            fflush(stdout);
        }
    }
    /* end of stream packet: always without -noeos=1, and with it when the input ran out
     * after at least one picture (dsv_main.c:797) */
    if (rc == 0 && (!o->noeos || (exhausted && total > 0))) {
        dsv_enc_end_of_stream(&enc, bufs);
        if (fwrite(bufs[0].data, 1, bufs[0].len, fo) != bufs[0].len) {
            rc = -1;
        }
        dsv_buf_free(&bufs[0]);
    }
    dsv_enc_free(&enc);
    free(pic);
    *pgot = got;
    *pexhausted = exhausted;
    return rc;
}

/* closed-GOP chunks, `threads` of them in flight: batches of threads x chunk frames */
static int
encode_chunked(const dsv_enc_opts *o, FILE *in, FILE *fo, int y4m, int nfr, int chunk, int threads, int gpus, int *pgot,
               int *pexhausted)
{
    size_t fsz = frame_size(o->w, o->h, o->fmt);
    int batch = threads * chunk, got = 0, exhausted = 0, rc = 0, last_len = 0;
    uint8_t *frames = dsv_pinned_alloc(fsz * (size_t) batch);
    dsv_pool *pool = dsv_pool_create(threads, NULL, gpus > 0 ? gpus : 1);
    dsv_enc_opts oc = *o;

    oc.noeos = 1; /* chunks never carry their own end-of-stream packet */
    if (!frames || !pool) {
        return -1;
    }
    while (!exhausted && (nfr < 0 || got < nfr)) {
        int n = 0, want = batch;
        uint8_t *out = NULL;
        size_t out_len = 0;
        if (nfr >= 0 && nfr - got < want) {
            want = nfr - got;
        }
        while (n < want) {
            if (read_frame(in, frames + (size_t) n * fsz, o, y4m) < 0) {
                exhausted = 1;
                break;
            }
            n++;
        }
        if (n == 0) {
            break;
        }
        if (dsv_pool_encode(pool, &oc, frames, n, chunk, &out, &out_len)) {
            rc = -1;
            break;
        }
        if (fwrite(out, 1, out_len, fo) != out_len) {
            rc = -1;
        }
        /* size of the last packet written (the end-of-stream packet links back to it) */
        {
            size_t off = 0;
            while (off + DSV_PACKET_HDR_SIZE <= out_len) {
                const uint8_t *q = out + off;
                size_t sz = ((size_t) q[10] << 24) | ((size_t) q[11] << 16) | ((size_t) q[12] << 8) | q[13];
                if (sz < DSV_PACKET_HDR_SIZE || off + sz > out_len) {
                    break;
                }
                last_len = (int) sz;
                off += sz;
            }
        }
        free(out);
        got += n;
        if (g_verbose) {
            printf("encoded %d frames\r", got);
            fflush(stdout);
        }
    }
    /* parallel_encode_yuv.sh semantics: with -noeos=1 only the process whose chunk runs into the
     * end of the input appends an end-of-stream packet (dsv_main.c:797), i.e. there is one
     * exactly when the last chunk is short; without -noeos the stream always ends with one */
    if (rc == 0 && got > 0 && (!o->noeos || (exhausted && got % chunk != 0))) {
        DSV_ENCODER enc;
        DSV_BUF b;
        dsv_enc_configure(&enc, &oc);
        enc.prev_link = last_len;
        dsv_enc_end_of_stream(&enc, &b);
        if (fwrite(b.data, 1, b.len, fo) != b.len) {
            rc = -1;
        }
        dsv_buf_free(&b);
    }
    dsv_pool_destroy(pool);
    dsv_pinned_free(frames);
    *pgot = got;
    *pexhausted = exhausted;
    return rc;
}

static int
do_encode(int argc, char **argv)
{
    dsv_enc_opts o;
    int fmt = 2, sfr = 0, nfr = -1, y4m = 0, gpus = 0, threads = 0, chunk = 0, dev = -1;
    int got = 0, exhausted = 0, r, i;
    size_t fsz;
    uint8_t *skip;
    FILE *in, *fo;
    OPT tab[] = {
        { "qp", &o.qp }, { "effort", &o.effort }, { "w", &o.w }, { "h", &o.h }, { "gop", &o.gop }, { "fmt", &fmt },
        { "nfr", &nfr }, { "sfr", &sfr }, { "noeos", &o.noeos }, { "fps_num", &o.fps_num }, { "fps_den", &o.fps_den },
        { "aspect_num", &o.aspect_num }, { "aspect_den", &o.aspect_den }, { "ipct", &o.ipct },
        { "pyrlevels", &o.pyrlevels }, { "rc_mode", &o.rc_mode }, { "rc_pergop", &o.rc_pergop }, { "kbps", &o.kbps },
        { "minqstep", &o.minqstep }, { "maxqstep", &o.maxqstep }, { "minqp", &o.minqp }, { "maxqp", &o.maxqp },
        { "iminqp", &o.iminqp }, { "stabref", &o.stabref }, { "scd", &o.scd }, { "tempaq", &o.tempaq },
        { "bszx", &o.bszx }, { "bszy", &o.bszy }, { "scpct", &o.scpct }, { "skipthresh", &o.skipthresh },
        { "varint", &o.varint }, { "psy", &o.psy }, { "dib", &o.dib }, { "y4m", &y4m }, { "ifilter", &o.ifilter },
        { "pfilter", &o.pfilter }, { "psharp", &o.psharp }, { "gpus", &gpus }, { "threads", &threads },
        { "chunk", &chunk }, { "dev", &dev }, { NULL, NULL }
    };

    dsv_enc_opts_default(&o, 352, 288, DSV_SUBSAMP_420, 30, 1);
    if (parse_opts(argc, argv, tab)) {
        return EXIT_FAILURE;
    }
    o.fmt = fmt_to_subsamp(fmt);
    in = fopen(g_inp, "rb");
    if (!in) {
        printf("error opening input file %s\n", g_inp);
        return EXIT_FAILURE;
    }
    if (y4m && dsv_y4m_read_hdr(in, &o.w, &o.h, &o.fmt, &o.fps_num, &o.fps_den, &o.aspect_num, &o.aspect_den)) {
        printf("bad Y4M file %s\n", g_inp);
        return EXIT_FAILURE;
    }
    if (o.w <= 0 || o.h <= 0 || (o.w & 1) || (o.h & 1)) {
        DSV_ERROR(("unsupported dimensions: %dx%d", o.w, o.h));
        return EXIT_FAILURE;
    }
    if (!confirm_overwrite(g_out)) {
        return EXIT_FAILURE;
    }
    fsz = frame_size(o.w, o.h, o.fmt);
    /* skip the frames in front of -sfr */
    skip = malloc(fsz ? fsz : 1);
    for (i = 0; i < sfr && !exhausted; i++) {
        if (read_frame(in, skip, &o, y4m) < 0) {
            exhausted = 1;
        }
    }
    free(skip);
    fo = fopen(g_out, "wb");
    if (!fo) {
        printf("error opening output file %s\n", g_out);
        return EXIT_FAILURE;
    }
    if (dev >= 0) {
        dsv_set_thread_device(dev);
    }
    if (exhausted) {
        r = 0;
    } else if (chunk > 0) {
        r = encode_chunked(&o, in, fo, y4m, nfr, chunk, threads > 0 ? threads : 8, gpus, &got, &exhausted);
    } else {
        r = encode_streaming(&o, in, fo, y4m, nfr, &got, &exhausted);
    }
    fclose(in);
    if (fclose(fo) || r) {
        DSV_ERROR(("encode failed: %s", dsvcu_last_error()));
        return EXIT_FAILURE;
    }
    if (g_verbose) {
        printf("encoded %d frames\n", got);
    }
    return exhausted ? -2 : EXIT_SUCCESS;
}

/* ------------------------------------------------------------------ decode */

/* chroma converters of -out420p (reference util.c:78-153): horizontal / vertical pair
 * averages with the last sample repeated, sample doubling for 4:1:0 */
static void
halve_h(const DSV_PLANE *s, DSV_PLANE *d)
{
    int i, j;
    for (j = 0; j < s->h; j++) {
        const uint8_t *sp = s->data + (size_t) j * s->stride;
        uint8_t *dp = d->data + (size_t) j * d->stride;
        for (i = 0; i < s->w; i += 2) {
            int n = i < s->w - 1 ? i + 1 : s->w - 1;
            dp[i >> 1] = (uint8_t) ((sp[i] + sp[n] + 1) >> 1);
        }
    }
}

/* vertical pair average; xrep = 2 additionally doubles every sample horizontally (4:1:1 -> 4:2:0) */
static void
halve_v(const DSV_PLANE *s, DSV_PLANE *d, int xrep)
{
    int i, j;
    for (j = 0; j < s->h; j += 2) {
        int n = j < s->h - 1 ? j + 1 : s->h - 1;
        const uint8_t *a = s->data + (size_t) j * s->stride, *b = s->data + (size_t) n * s->stride;
        uint8_t *dp = d->data + (size_t) (j >> 1) * d->stride;
        for (i = 0; i < s->w * xrep; i++) {
            dp[i] = (uint8_t) ((a[i / xrep] + b[i / xrep] + 1) >> 1);
        }
    }
}

static void
double_both(const DSV_PLANE *s, DSV_PLANE *d)
{
    int i, j;
    for (j = 0; j < s->h * 2; j++) {
        for (i = 0; i < s->w * 2; i++) {
            d->data[i + (size_t) d->stride * j] = s->data[(i >> 1) + (size_t) s->stride * (j >> 1)];
        }
    }
}

static DSV_FRAME *
to_420(DSV_FRAME *f, int subsamp)
{
    DSV_FRAME *o = dsv_mk_frame(DSV_SUBSAMP_420, f->width, f->height, 0);
    int c, y;
    for (c = 1; c < 3; c++) {
        if (subsamp == DSV_SUBSAMP_444) {
            DSV_FRAME *t = dsv_mk_frame(DSV_SUBSAMP_422, f->width, f->height, 0);
            halve_h(&f->planes[c], &t->planes[c]);
            halve_v(&t->planes[c], &o->planes[c], 1);
            dsv_frame_ref_dec(t);
        } else if (subsamp == DSV_SUBSAMP_422 || subsamp == DSV_SUBSAMP_UYVY) {
            halve_v(&f->planes[c], &o->planes[c], 1);
        } else if (subsamp == DSV_SUBSAMP_411) {
            halve_v(&f->planes[c], &o->planes[c], 2);
        } else if (subsamp == DSV_SUBSAMP_410) {
            double_both(&f->planes[c], &o->planes[c]);
        }
    }
    for (y = 0; y < o->planes[0].h; y++) {
        memcpy(o->planes[0].data + (size_t) y * o->planes[0].stride, f->planes[0].data + (size_t) y * f->planes[0].stride,
               (size_t) f->planes[0].w);
    }
    return o;
}

/* reads the next packet (size from its next-link field, dsv_main.c:912-957) */
static int
read_packet(FILE *in, DSV_BUF *b)
{
    uint8_t hdr[DSV_PACKET_HDR_SIZE];
    size_t size;
    if (fread(hdr, 1, DSV_PACKET_HDR_SIZE, in) != DSV_PACKET_HDR_SIZE) {
        return -1;
    }
    size = ((size_t) hdr[10] << 24) | ((size_t) hdr[11] << 16) | ((size_t) hdr[12] << 8) | hdr[13];
    if (size == 0) {
        size = DSV_PACKET_HDR_SIZE; /* end of stream */
    }
    if (size < DSV_PACKET_HDR_SIZE || size > (1u << 30)) {
        return -1;
    }
    dsv_mk_buf(b, (int) size);
    memcpy(b->data, hdr, DSV_PACKET_HDR_SIZE);
    if (size > DSV_PACKET_HDR_SIZE && fread(b->data + DSV_PACKET_HDR_SIZE, 1, size - DSV_PACKET_HDR_SIZE, in) != size - DSV_PACKET_HDR_SIZE) {
        dsv_buf_free(b);
        return -1;
    }
    return 0;
}

static void
write_frame(FILE *fo, DSV_FRAME *f, const DSV_META *md, int subsamp, int y4m, int *first)
{
    if (y4m) {
        if (*first) {
            dsv_y4m_write_hdr(fo, md->width, md->height, subsamp, md->fps_num, md->fps_den, md->aspect_num, md->aspect_den);
            *first = 0;
        }
        dsv_y4m_write_frame_hdr(fo);
    }
    if (dsv_yuv_write_seq(fo, f->planes) < 0) {
        DSV_ERROR(("failed to write frame"));
    }
}

/* one decoder instance, one packet at a time (reference decode(), dsv_main.c:1000-1109) */
static int
decode_streaming(FILE *in, FILE *fo, int y4m, int out420p, int postsharp, int drawinfo)
{
    DSV_DECODER dec;
    DSV_META md;
    int have_meta = 0, first = 1, nfr = 0;

    memset(&dec, 0, sizeof(dec));
    memset(&md, 0, sizeof(md));
    dec.draw_info = drawinfo;
    for (;;) {
        DSV_BUF b;
        DSV_FRAME *fr = NULL;
        DSV_FNUM fn;
        int code;
        if (read_packet(in, &b) < 0) {
            DSV_ERROR(("error reading packet"));
            break;
        }
        code = dsv_dec(&dec, &b, &fr, &fn);
        if (code == DSV_DEC_GOT_META) {
            if (!have_meta) {
                md = dec.vidmeta;
                have_meta = 1;
            }
            continue;
        }
        if (code == DSV_DEC_EOS) {
            break;
        }
        if (code != DSV_DEC_OK || !fr) {
            continue;
        }
        if (!have_meta) {
            DSV_ERROR(("no metadata!"));
            dsv_frame_ref_dec(fr);
            break;
        }
        if (out420p && md.subsamp != DSV_SUBSAMP_420) {
            DSV_FRAME *f420 = to_420(fr, md.subsamp);
            if (postsharp) {
                dsv_post_process(&f420->planes[0]);
            }
            write_frame(fo, f420, &md, DSV_SUBSAMP_420, y4m, &first);
            dsv_frame_ref_dec(f420);
        } else {
            if (postsharp) {
                dsv_post_process(&fr->planes[0]);
            }
            write_frame(fo, fr, &md, md.subsamp, y4m, &first);
        }
        dsv_frame_ref_dec(fr);
        nfr++;
        if (g_verbose) {
            printf("\rdecoded frame %d", nfr);
            fflush(stdout);
        }
    }
    dsv_dec_free(&dec);
    return nfr;
}

/* closed GOPs on `threads` decoder instances: the stream is read packet by packet and handed
 * to the pool in batches of about 2 x threads closed GOPs (a GOP starts at a metadata packet) */
static int
decode_batched(FILE *in, FILE *fo, int y4m, int threads, int gpus)
{
    dsv_pool *pool = dsv_pool_create(threads, NULL, gpus > 0 ? gpus : 1);
    uint8_t *buf = NULL;
    size_t len = 0, cap = 0;
    int first = 1, total = 0, done = 0, gops = 0, have_pending = 0, failed = 0;
    DSV_BUF pending;

    if (!pool) {
        return -1;
    }
    memset(&pending, 0, sizeof(pending));
    while (!done && !failed) {
        DSV_BUF b;
        int got = 0;
        if (have_pending) {
            b = pending;
            have_pending = 0;
            got = 1;
        } else if (read_packet(in, &b) == 0) {
            got = 1;
        } else {
            done = 1;
        }
        if (got) {
            const int type = b.data[DSV_PACKET_TYPE_OFFSET];
            if (type == DSV_PT_META && gops >= 2 * threads && len > 0) {
                pending = b; /* opens the next batch */
                have_pending = 1;
            } else {
                if (len + b.len > cap) {
                    cap = (len + b.len) * 2 + (1 << 20);
                    buf = realloc(buf, cap);
                    if (!buf) {
                        failed = 1;
                        break;
                    }
                }
                memcpy(buf + len, b.data, b.len);
                len += b.len;
                gops += type == DSV_PT_META;
                done = type == DSV_PT_EOS;
                dsv_buf_free(&b);
                if (!done) {
                    continue;
                }
            }
        }
        if (len > 0) {
            uint8_t *yuv = NULL;
            size_t yuv_len = 0, fsz;
            DSV_META md;
            int n = 0, i;
            if (dsv_pool_decode_alloc(pool, buf, len, &yuv, &yuv_len, &n, &md)) {
                DSV_ERROR(("decode failed: %s", dsvcu_last_error()));
                failed = 1;
                break;
            }
            fsz = frame_size(md.width, md.height, md.subsamp);
            for (i = 0; i < n; i++) {
                if (y4m) {
                    if (first) {
                        dsv_y4m_write_hdr(fo, md.width, md.height, md.subsamp, md.fps_num, md.fps_den, md.aspect_num,
                                          md.aspect_den);
                        first = 0;
                    }
                    dsv_y4m_write_frame_hdr(fo);
                }
                if (fwrite(yuv + (size_t) i * fsz, 1, fsz, fo) != fsz) {
                    failed = 1;
                }
            }
            dsv_pinned_free(yuv);
            total += n;
            len = 0;
            gops = 0;
            if (g_verbose) {
                printf("\rdecoded %d frames", total);
                fflush(stdout);
            }
        }
    }
    if (have_pending) {
        dsv_buf_free(&pending);
    }
    free(buf);
    dsv_pool_destroy(pool);
    return failed ? -1 : total;
}

static int
do_decode(int argc, char **argv)
{
    int out420p = 0, y4m = 0, postsharp = 0, drawinfo = 0, gpus = 0, threads = 1, dev = -1, nfr;
    FILE *in, *fo;
    OPT tab[] = { { "out420p", &out420p }, { "y4m", &y4m }, { "postsharp", &postsharp }, { "drawinfo", &drawinfo },
                  { "gpus", &gpus }, { "threads", &threads }, { "dev", &dev }, { NULL, NULL } };

    if (parse_opts(argc, argv, tab)) {
        return EXIT_FAILURE;
    }
    in = fopen(g_inp, "rb");
    if (!in) {
        printf("error opening input file %s\n", g_inp);
        return EXIT_FAILURE;
    }
    if (!confirm_overwrite(g_out)) {
        return EXIT_FAILURE;
    }
    fo = fopen(g_out, "wb");
    if (!fo) {
        printf("error opening output file %s\n", g_out);
        return EXIT_FAILURE;
    }
    if (dev >= 0) {
        dsv_set_thread_device(dev);
    }
    if (threads > 1 && !out420p && !postsharp && !drawinfo) {
        nfr = decode_batched(in, fo, y4m, threads, gpus);
    } else {
        nfr = decode_streaming(in, fo, y4m, out420p, postsharp, drawinfo);
    }
    fclose(in);
    fclose(fo);
    if (nfr < 0) {
        return EXIT_FAILURE;
    }
    if (g_verbose) {
        printf("\ndecoded %d frames\n", nfr);
    }
    return EXIT_SUCCESS;
}

int
main(int argc, char **argv)
{
    if (argc < 2 || (argv[1][0] != 'e' && argv[1][0] != 'd')) {
        printf("usage: %s e|d -inp=<file> -out=<file> [-y] [-v] [-l<0-4>] [-name=value ...]\n"
               "  option names and defaults are those of the reference dsv2 CLI\n"
               "  extensions: -gpus=N -threads=T -chunk=K (closed-GOP sharding), -dev=D\n",
               argv[0]);
        return EXIT_FAILURE;
    }
    return argv[1][0] == 'e' ? do_encode(argc, argv) : do_decode(argc, argv);
}
