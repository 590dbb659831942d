/*
 * dsv_core.c -- host plumbing shared by encoder, decoder and CLI: logging,
 * counted allocation, byte buffers, host frames, raw YUV / Y4M file I/O.
 *
 * API-compatible restatement of the host helpers of the reference
 * (src/dsv.c:19-322, src/frame.c:19-207, src/util.c:184-488).  None of this is
 * on the pixel hot path; pixel operators live in csrc/ and are reached through
 * dsv_cuda.h.
 */
#include <stdlib.h>
#include <string.h>
#include "dsv_host.h"

/* ---------------------------------------------------------------- logging */

char *dsv_lvlname[DSV_LEVEL_DEBUG + 1] = { "NONE", "ERROR", "WARNING", "INFO", "DEBUG" };
static int g_loglevel = DSV_LEVEL_ERROR;

void
dsv_set_log_level(int level)
{
    g_loglevel = level;
}

int
dsv_get_log_level(void)
{
    return g_loglevel;
}

/* ------------------------------------------------------------- allocation */

/* every block carries a 16-byte header with its size so dsv_free can keep the
 * statistics; memory is zeroed (callers rely on it, reference dsv.c:41-107) */
static unsigned g_allocated = 0, g_freed = 0, g_bytes = 0, g_peak = 0;

void *
dsv_alloc(int size)
{
    unsigned char *p;
    if (size < 0) {
        return NULL;
    }
    p = calloc(1, (size_t) size + 16);
    if (!p) {
        return NULL;
    }
    *(int *) p = size;
    /* encoder / decoder instances run on several host threads */
    __atomic_add_fetch(&g_allocated, 1, __ATOMIC_RELAXED);
    {
        unsigned now = __atomic_add_fetch(&g_bytes, (unsigned) size, __ATOMIC_RELAXED);
        if (now > g_peak) {
            g_peak = now;
        }
    }
    return p + 16;
}

void
dsv_free(void *ptr)
{
    unsigned char *p = ptr;
    if (!p) {
        return;
    }
    p -= 16;
    __atomic_sub_fetch(&g_bytes, (unsigned) *(int *) p, __ATOMIC_RELAXED);
    __atomic_add_fetch(&g_freed, 1, __ATOMIC_RELAXED);
    free(p);
}

void
dsv_memory_report(void)
{
    DSV_DEBUG(("n alloc: %u", g_allocated));
    DSV_DEBUG(("n freed: %u", g_freed));
    DSV_DEBUG(("alloc - freed = %d", (int) (g_allocated - g_freed)));
    DSV_DEBUG(("bytes remaining: %u", g_bytes));
    DSV_DEBUG(("peak alloc: %u", g_peak));
}

void
dsv_mk_buf(DSV_BUF *buf, int size)
{
    memset(buf, 0, sizeof(*buf));
    buf->data = dsv_alloc(size);
    buf->len = (unsigned) size;
}

void
dsv_buf_free(DSV_BUF *buf)
{
    if (buf->data) {
        dsv_free(buf->data);
        buf->data = NULL;
    }
}

/* ------------------------------------------------------------ host frames */

void
dsv_mk_coefs(DSV_COEFS *c, int format, int width, int height)
{
    int cw = DSV_ROUND_SHIFT(width, DSV_FORMAT_H_SHIFT(format));
    int ch = DSV_ROUND_SHIFT(height, DSV_FORMAT_V_SHIFT(format));
    int n0, n1;
    cw = (cw + 1) & ~1;
    ch = (ch + 1) & ~1;
    c[0].width = width;
    c[0].height = height;
    c[1].width = c[2].width = cw;
    c[1].height = c[2].height = ch;
    n0 = width * height;
    n1 = cw * ch;
    c[0].data = dsv_alloc((n0 + 2 * n1) * (int) sizeof(DSV_SBC));
    c[1].data = c[0].data + n0;
    c[2].data = c[1].data + n1;
}

DSV_FRAME *
dsv_mk_frame(int format, int width, int height, int border)
{
    DSV_FRAME *f = dsv_alloc(sizeof(*f));
    int ext = border ? DSV_MAX_BLOCK_SIZE : 0;
    int i, off = 0;
    int w[3], h[3];

    f->refcount = 1;
    f->format = format;
    f->width = width;
    f->height = height;
    f->border = !!border;
    w[0] = width;
    h[0] = height;
    w[1] = w[2] = DSV_ROUND_SHIFT(width, DSV_FORMAT_H_SHIFT(format));
    h[1] = h[2] = DSV_ROUND_SHIFT(height, DSV_FORMAT_V_SHIFT(format));
    for (i = 0; i < 3; i++) {
        DSV_PLANE *p = &f->planes[i];
        p->format = format;
        p->w = w[i];
        p->h = h[i];
        p->stride = (int) DSV_ROUND_POW2(w[i] + ext * 2, 4);
        p->len = p->stride * (h[i] + ext * 2);
        off += p->len;
    }
    f->alloc = dsv_alloc(off);
    off = 0;
    for (i = 0; i < 3; i++) {
        DSV_PLANE *p = &f->planes[i];
        p->data = f->alloc + off + p->stride * ext + ext;
        off += p->len;
    }
    return f;
}

DSV_FRAME *
dsv_load_planar_frame(int format, void *data, int width, int height)
{
    /* wraps caller memory; the frame does not own the pixels */
    DSV_FRAME *f = dsv_alloc(sizeof(*f));
    uint8_t *p = data;
    int i;

    f->refcount = 1;
    f->format = format;
    f->width = width;
    f->height = height;
    for (i = 0; i < 3; i++) {
        DSV_PLANE *pl = &f->planes[i];
        pl->format = format;
        pl->w = i ? DSV_ROUND_SHIFT(width, DSV_FORMAT_H_SHIFT(format)) : width;
        pl->h = i ? DSV_ROUND_SHIFT(height, DSV_FORMAT_V_SHIFT(format)) : height;
        pl->stride = pl->w;
        pl->len = pl->stride * pl->h;
        pl->data = p;
        p += pl->len;
    }
    return f;
}

DSV_FRAME *
dsv_frame_ref_inc(DSV_FRAME *frame)
{
    DSV_ASSERT(frame && frame->refcount > 0);
    frame->refcount++;
    return frame;
}

void
dsv_frame_ref_dec(DSV_FRAME *frame)
{
    DSV_ASSERT(frame && frame->refcount > 0);
    if (--frame->refcount == 0) {
        if (frame->alloc) {
            dsv_free(frame->alloc);
        }
        dsv_free(frame);
    }
}

void
dsv_plane_xy(DSV_FRAME *frame, DSV_PLANE *out, int c, int x, int y)
{
    DSV_PLANE *p = frame->planes + c;
    out->format = p->format;
    out->data = DSV_GET_XY(p, x, y);
    out->stride = p->stride;
    out->w = MAX(0, p->w - x);
    out->h = MAX(0, p->h - y);
}

void
dsv_host_copy_planes(DSV_FRAME *dst, DSV_FRAME *src)
{
    int c, i;
    for (c = 0; c < 3; c++) {
        uint8_t *sp = src->planes[c].data, *dp = dst->planes[c].data;
        for (i = 0; i < dst->planes[c].h; i++) {
            memcpy(dp, sp, (size_t) src->planes[c].w);
            sp += src->planes[c].stride;
            dp += dst->planes[c].stride;
        }
    }
}

/* Host-side frame copies (reference frame.c:185-207, :436-446 semantics for
 * the visible area).  Bordered HOST frames are not part of this build -- the
 * padded pictures the operators read live on the device and are extended there
 * (dsvcu_extend_frame) -- so a host copy never synthesises border pixels. */
void
dsv_frame_copy(DSV_FRAME *dst, DSV_FRAME *src)
{
    dsv_host_copy_planes(dst, src);
}

DSV_FRAME *
dsv_clone_frame(DSV_FRAME *f, int border)
{
    DSV_FRAME *d = dsv_mk_frame(f->format, f->width, f->height, border);
    dsv_host_copy_planes(d, f);
    return d;
}

/* ------------------------------------------------------------ raw YUV I/O */

static size_t
yuv_frame_bytes(int w, int h, int subsamp)
{
    size_t cw = (size_t) DSV_ROUND_SHIFT(w, DSV_FORMAT_H_SHIFT(subsamp));
    size_t chh = (size_t) DSV_ROUND_SHIFT(h, DSV_FORMAT_V_SHIFT(subsamp));
    return (size_t) w * h + 2 * cw * chh;
}

int
dsv_yuv_write_seq(FILE *out, DSV_PLANE *p)
{
    int c, y;
    if (!out) {
        return -1;
    }
    for (c = 0; c < 3; c++) {
        for (y = 0; y < p[c].h; y++) {
            if (fwrite(DSV_GET_LINE(&p[c], y), 1, (size_t) p[c].w, out) != (size_t) p[c].w) {
                return -1;
            }
        }
    }
    return 0;
}

int
dsv_yuv_write(FILE *out, int fno, DSV_PLANE *p)
{
    size_t fsz;
    if (!out || fno < 0) {
        return -1;
    }
    fsz = (size_t) p[0].w * p[0].h + (size_t) p[1].w * p[1].h + (size_t) p[2].w * p[2].h;
    if (fseek(out, (long) (fno * fsz), SEEK_SET)) {
        return -1;
    }
    return dsv_yuv_write_seq(out, p);
}

/* packed UYVY 4:2:2 (2 bytes per pixel: U Y V Y) -> planar Y, U, V (reference dsv.c:142-176) */
static int
read_uyvy(FILE *in, uint8_t *o, int w, int h)
{
    uint8_t *yp = o, *up = o + (size_t) w * h, *vp = up + (size_t) (w / 2) * h;
    size_t row = (size_t) w * 2;
    uint8_t *line = malloc(row ? row : 1);
    int i, j, ok = 0;
    if (!line) {
        return -1;
    }
    for (j = 0; j < h; j++) {
        const uint8_t *t = line;
        if (fread(line, 1, row, in) != row) {
            ok = -1;
            break;
        }
        for (i = 0; i < w / 2; i++, t += 4) {
            *up++ = t[0];
            *yp++ = t[1];
            *vp++ = t[2];
            *yp++ = t[3];
        }
    }
    free(line);
    return ok;
}

int
dsv_yuv_read_seq(FILE *in, uint8_t *o, int w, int h, int subsamp)
{
    size_t n = yuv_frame_bytes(w, h, subsamp);
    if (!in) {
        return -1;
    }
    if (subsamp == DSV_SUBSAMP_UYVY) {
        return read_uyvy(in, o, w, h);
    }
    if (fread(o, 1, n, in) != n) {
        return -1;
    }
    return 0;
}

int
dsv_yuv_read(FILE *in, int fno, uint8_t *o, int w, int h, int subsamp)
{
    size_t n = yuv_frame_bytes(w, h, subsamp);
    if (!in || fno < 0) {
        return -1;
    }
    if (fseek(in, (long) (fno * n), SEEK_SET)) {
        return -1;
    }
    return dsv_yuv_read_seq(in, o, w, h, subsamp);
}

/* --------------------------------------------------------------- Y4M I/O */

/* parses "YUV4MPEG2 W.. H.. F..:.. A..:.. I. C..." (subset of util.c:184-307) */
int
dsv_y4m_read_hdr(FILE *in, int *w, int *h, int *subsamp, int *fpsn, int *fpsd, int *aspn, int *aspd)
{
    char line[256], *tok;
    int n = 0, c;
    while ((c = fgetc(in)) != EOF && c != '\n' && n < 255) {
        line[n++] = (char) c;
    }
    line[n] = 0;
    if (strncmp(line, "YUV4MPEG2", 9)) {
        return -1;
    }
    *subsamp = DSV_SUBSAMP_420;
    *fpsn = 30;
    *fpsd = 1;
    *aspn = *aspd = 1;
    for (tok = strtok(line + 9, " "); tok; tok = strtok(NULL, " ")) {
        switch (tok[0]) {
            case 'W': *w = atoi(tok + 1); break;
            case 'H': *h = atoi(tok + 1); break;
            case 'F': sscanf(tok + 1, "%d:%d", fpsn, fpsd); break;
            case 'A': sscanf(tok + 1, "%d:%d", aspn, aspd); break;
            case 'C':
                if (!strncmp(tok + 1, "444", 3)) {
                    *subsamp = DSV_SUBSAMP_444;
                } else if (!strncmp(tok + 1, "422", 3)) {
                    *subsamp = DSV_SUBSAMP_422;
                } else if (!strncmp(tok + 1, "411", 3)) {
                    *subsamp = DSV_SUBSAMP_411;
                } else if (!strncmp(tok + 1, "420", 3)) {
                    *subsamp = DSV_SUBSAMP_420;
                } else {
                    return -1;
                }
                break;
            default: break;
        }
    }
    if (*aspn == 0 || *aspd == 0) {
        *aspn = *aspd = 1;
    }
    return 0;
}

int
dsv_y4m_read_frame(FILE *in, uint8_t *o, int w, int h, int subsamp)
{
    int c, n = 0;
    char tag[8];
    while ((c = fgetc(in)) != EOF && c != '\n') {
        if (n < 7) {
            tag[n++] = (char) c;
        }
    }
    tag[n < 7 ? n : 7] = 0;
    if (c == EOF || strncmp(tag, "FRAME", 5)) {
        return -1;
    }
    return dsv_yuv_read_seq(in, o, w, h, subsamp);
}

void
dsv_y4m_write_hdr(FILE *out, int w, int h, int subsamp, int fpsn, int fpsd, int aspn, int aspd)
{
    const char *c = "420";
    if (subsamp == DSV_SUBSAMP_444) {
        c = "444";
    } else if (subsamp == DSV_SUBSAMP_422 || subsamp == DSV_SUBSAMP_UYVY) {
        c = "422";
    } else if (subsamp == DSV_SUBSAMP_411) {
        c = "411";
    } else if (subsamp == DSV_SUBSAMP_410) {
        c = "410";
    }
    fprintf(out, "YUV4MPEG2 W%d H%d F%d:%d A%d:%d Ip C%s\n", w, h, fpsn, fpsd, aspn, aspd, c);
}

void
dsv_y4m_write_frame_hdr(FILE *out)
{
    fwrite("FRAME\n", 1, 6, out);
}
