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
 * Not implemented: -out420p and -drawinfo of the decoder, stdin/stdout piping.
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
do_encode(int argc, char **argv)
{
    dsv_enc_opts o;
    int fmt = 2, sfr = 0, nfr = -1, y4m = 0, gpus = 0, threads = 0, chunk = 0, dev = -1;
    int got = 0, exhausted = 0, r, i;
    uint8_t *frames = NULL, *out = NULL;
    size_t fsz, cap = 0, out_len = 0;
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
    /* read frames [sfr, sfr + nfr) */
    for (i = 0; nfr < 0 || got < nfr; i++) {
        uint8_t *dst;
        if ((size_t) (got + 1) * fsz > cap) {
            cap = cap ? cap * 2 : fsz * 16;
            frames = realloc(frames, cap);
            if (!frames) {
                return EXIT_FAILURE;
            }
        }
        dst = frames + (size_t) got * fsz;
        r = y4m ? dsv_y4m_read_frame(in, dst, o.w, o.h, o.fmt) : dsv_yuv_read_seq(in, dst, o.w, o.h, o.fmt);
        if (r < 0) {
            exhausted = 1;
            break;
        }
        if (i >= sfr) {
            got++;
        }
    }
    fclose(in);
    if (dev >= 0) {
        dsv_set_thread_device(dev);
    }
    if (chunk > 0 && got > 0) {
        int nth = threads > 0 ? threads : 8;
        r = dsv_encode_sharded(&o, frames, got, chunk, nth, NULL, gpus > 0 ? gpus : 1, &out, &out_len);
    } else {
        r = dsv_encode_buffer(&o, frames, got, exhausted, &out, &out_len);
    }
    free(frames);
    if (r) {
        DSV_ERROR(("encode failed: %s", dsvcu_last_error()));
        return EXIT_FAILURE;
    }
    fo = fopen(g_out, "wb");
    if (!fo || fwrite(out, 1, out_len, fo) != out_len) {
        printf("error writing %s\n", g_out);
        return EXIT_FAILURE;
    }
    fclose(fo);
    free(out);
    if (g_verbose) {
        printf("encoded %d frames to %lu bytes\n", got, (unsigned long) out_len);
    }
    return exhausted ? -2 : EXIT_SUCCESS;
}

static int
do_decode(int argc, char **argv)
{
    int out420p = 0, y4m = 0, postsharp = 0, drawinfo = 0, gpus = 0, threads = 1, dev = -1;
    uint8_t *dsv, *yuv = NULL;
    size_t len, yuv_len = 0, fsz;
    int nfr = 0, i;
    DSV_META md;
    FILE *in, *fo;
    OPT tab[] = { { "out420p", &out420p }, { "y4m", &y4m }, { "postsharp", &postsharp }, { "drawinfo", &drawinfo },
                  { "gpus", &gpus }, { "threads", &threads }, { "dev", &dev }, { NULL, NULL } };

    if (parse_opts(argc, argv, tab)) {
        return EXIT_FAILURE;
    }
    if (out420p || drawinfo || postsharp) {
        DSV_WARNING(("-out420p / -drawinfo / -postsharp are not implemented by dsv2cu"));
    }
    in = fopen(g_inp, "rb");
    if (!in) {
        printf("error opening input file %s\n", g_inp);
        return EXIT_FAILURE;
    }
    fseek(in, 0, SEEK_END);
    len = (size_t) ftell(in);
    fseek(in, 0, SEEK_SET);
    dsv = malloc(len ? len : 1);
    if (!dsv || fread(dsv, 1, len, in) != len) {
        return EXIT_FAILURE;
    }
    fclose(in);
    if (!confirm_overwrite(g_out)) {
        return EXIT_FAILURE;
    }
    if (dev >= 0) {
        dsv_set_thread_device(dev);
    }
    if (dsv_decode_sharded(dsv, len, threads > 0 ? threads : 1, NULL, gpus > 0 ? gpus : 1, 0, &yuv, &yuv_len, &nfr, &md)) {
        DSV_ERROR(("decode failed: %s", dsvcu_last_error()));
        return EXIT_FAILURE;
    }
    free(dsv);
    fo = fopen(g_out, "wb");
    if (!fo) {
        printf("error opening output file %s\n", g_out);
        return EXIT_FAILURE;
    }
    fsz = frame_size(md.width, md.height, md.subsamp);
    if (y4m) {
        dsv_y4m_write_hdr(fo, md.width, md.height, md.subsamp, md.fps_num, md.fps_den, md.aspect_num, md.aspect_den);
    }
    for (i = 0; i < nfr; i++) {
        if (y4m) {
            dsv_y4m_write_frame_hdr(fo);
        }
        fwrite(yuv + (size_t) i * fsz, 1, fsz, fo);
    }
    fclose(fo);
    free(yuv);
    if (g_verbose) {
        printf("decoded %d frames\n", nfr);
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
