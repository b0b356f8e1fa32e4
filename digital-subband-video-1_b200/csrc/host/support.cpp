/*
 * support.cpp -- host-side support symbols of the drop-in ABI (include/dsv1_b200.h): the zeroing
 * allocator, logging, DSV_BUF, host DSV_FRAME objects and raw YUV file I/O, plus the three CLI
 * helpers from util.c.  Pure host plumbing (not GPU work); the
 * reference CLI links every one of these, so they exist with the same names and behaviour.
 */
#include "dsv1_b200.h"

#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../common.cuh"

extern "C" {

/* ---- logging (dsv.c:19-39) ---------------------------------------------------------------- */
char *dsv_lvlname[DSV_LEVEL_DEBUG + 1] = {(char *) "NONE", (char *) "ERROR", (char *) "WARNING", (char *) "INFO",
                                          (char *) "DEBUG"};
static std::atomic<int> g_log_level{DSV_LEVEL_ERROR};
void dsv_set_log_level(int level) { g_log_level.store(level); }
int dsv_get_log_level(void) { return g_log_level.load(); }

/* ---- allocator (dsv.c:41-77): zero-filled blocks with a 16-byte size header; thread-safe stats -- */
static std::atomic<unsigned> g_nalloc{0}, g_nfree{0}, g_balloc{0}, g_bfree{0};

void *dsv_alloc(int size)
{
    uint8_t *p = (uint8_t *) calloc(1, (size_t) size + 16);
    if (!p) {
        DSV_ERROR(("out of memory (%d bytes)", size));
        return NULL;
    }
    *(int32_t *) p = size;
    g_nalloc++;
    g_balloc += (unsigned) size;
    return p + 16;
}

void dsv_free(void *ptr)
{
    if (!ptr) {
        return;
    }
    uint8_t *p = (uint8_t *) ptr - 16;
    g_nfree++;
    g_bfree += (unsigned) *(int32_t *) p;
    free(p);
}

void dsv_memory_report(void)
{
    DSV_DEBUG(("n alloc: %u", g_nalloc.load()));
    DSV_DEBUG(("n freed: %u", g_nfree.load()));
    DSV_DEBUG(("alloc bytes: %u", g_balloc.load()));
    DSV_DEBUG(("freed bytes: %u", g_bfree.load()));
    DSV_DEBUG(("bytes not freed: %d", (int) (g_balloc.load() - g_bfree.load())));
}

/* ---- DSV_BUF (dsv.c:172-187) ---------------------------------------------------------------- */
void dsv_mk_buf(DSV_BUF *buf, int size)
{
    buf->data = (unsigned char *) dsv_alloc(size);
    buf->len = (unsigned) size;
}

void dsv_buf_free(DSV_BUF *buf)
{
    if (buf->data) {
        dsv_free(buf->data);
        buf->data = NULL;
    }
}

/* ---- host frames (frame.c:29-221, 263-295) --------------------------------------------------- */
static void set_plane(DSV_PLANE *p, int format, int w, int h, int stride, int hs, int vs, int rows)
{
    p->format = format;
    p->w = w;
    p->h = h;
    p->stride = stride;
    p->len = stride * rows;
    p->hs = hs;
    p->vs = vs;
}

DSV_FRAME *dsv_mk_frame(int format, int width, int height, int border)
{
    DSV_FRAME *f = (DSV_FRAME *) dsv_alloc(sizeof(DSV_FRAME));
    const int ext = border ? DSV_MAX_BLOCK_SIZE : 0;
    const int hs = DSV_FORMAT_H_SHIFT(format), vs = DSV_FORMAT_V_SHIFT(format);
    const int cw = DSV_ROUND_SHIFT(width, hs), ch = DSV_ROUND_SHIFT(height, vs);
    f->refcount = 1;
    f->format = format;
    f->width = width;
    f->height = height;
    f->border = !!border;
    set_plane(&f->planes[0], format, width, height, (int) DSV_ROUND_POW2(width + 2 * ext, 4), 0, 0, height + 2 * ext);
    set_plane(&f->planes[1], format, cw, ch, (int) DSV_ROUND_POW2(cw + 2 * ext, 4), hs, vs, ch + 2 * ext);
    set_plane(&f->planes[2], format, cw, ch, (int) DSV_ROUND_POW2(cw + 2 * ext, 4), hs, vs, ch + 2 * ext);
    f->alloc = (uint8_t *) dsv_alloc(f->planes[0].len + f->planes[1].len + f->planes[2].len);
    uint8_t *at = f->alloc;
    for (int c = 0; c < 3; c++) {
        f->planes[c].data = at + f->planes[c].stride * ext + ext;
        at += f->planes[c].len;
    }
    return f;
}

/* wraps caller memory laid out Y,U,V tightly packed; no copy (frame.c:122-164) */
DSV_FRAME *dsv_load_planar_frame(int format, void *data, int width, int height)
{
    DSV_FRAME *f = (DSV_FRAME *) dsv_alloc(sizeof(DSV_FRAME));
    const int hs = DSV_FORMAT_H_SHIFT(format), vs = DSV_FORMAT_V_SHIFT(format);
    const int cw = DSV_ROUND_SHIFT(width, hs), ch = DSV_ROUND_SHIFT(height, vs);
    f->refcount = 1;
    f->format = format;
    f->width = width;
    f->height = height;
    set_plane(&f->planes[0], format, width, height, width, 0, 0, height);
    set_plane(&f->planes[1], format, cw, ch, cw, hs, vs, ch);
    set_plane(&f->planes[2], format, cw, ch, cw, hs, vs, ch);
    f->planes[0].data = (uint8_t *) data;
    f->planes[1].data = f->planes[0].data + f->planes[0].len;
    f->planes[2].data = f->planes[1].data + f->planes[1].len;
    return f;
}

} /* extern "C" */

/* ---- pinned picture pool ---------------------------------------------------------------------------
 * Pictures handed out by dsv_dec live in page-locked memory so the decoder's device->host copy is a straight
 * DMA at PCIe speed (a pageable destination is staged through a bounce buffer at a fraction of that).
 * Page-locking is slow, so the storage is recycled: dsv_frame_ref_dec returns it here instead of freeing. */
namespace {
struct PinnedPool {
    std::mutex m;
    std::unordered_map<size_t, std::vector<uint8_t *>> idle;
    std::unordered_map<uint8_t *, size_t> owned;
};
PinnedPool &pinned_pool()
{
    static PinnedPool *p = new PinnedPool(); /* never destroyed: frames may outlive static destruction order */
    return *p;
}
uint8_t *pinned_get(size_t bytes)
{
    PinnedPool &pp = pinned_pool();
    {
        std::lock_guard<std::mutex> lk(pp.m);
        auto it = pp.idle.find(bytes);
        if (it != pp.idle.end() && !it->second.empty()) {
            uint8_t *p = it->second.back();
            it->second.pop_back();
            return p;
        }
    }
    uint8_t *p = nullptr;
    CUDA_CHECK(cudaMallocHost(&p, bytes));
    std::lock_guard<std::mutex> lk(pp.m);
    pp.owned[p] = bytes;
    return p;
}
bool pinned_put(uint8_t *p)
{
    PinnedPool &pp = pinned_pool();
    std::lock_guard<std::mutex> lk(pp.m);
    auto it = pp.owned.find(p);
    if (it == pp.owned.end()) {
        return false;
    }
    pp.idle[it->second].push_back(p);
    return true;
}
} // namespace

namespace dsv {
DSV_FRAME *mk_frame_pinned(int format, int width, int height);
}

/* bordered frame (same layout as dsv_mk_frame(..., 1)) in pinned memory; contents are NOT cleared */
DSV_FRAME *dsv::mk_frame_pinned(int format, int width, int height)
{
    DSV_FRAME *f = (DSV_FRAME *) dsv_alloc(sizeof(DSV_FRAME));
    const int ext = DSV_MAX_BLOCK_SIZE;
    const int hs = DSV_FORMAT_H_SHIFT(format), vs = DSV_FORMAT_V_SHIFT(format);
    const int cw = DSV_ROUND_SHIFT(width, hs), ch = DSV_ROUND_SHIFT(height, vs);
    f->refcount = 1;
    f->format = format;
    f->width = width;
    f->height = height;
    f->border = 1;
    set_plane(&f->planes[0], format, width, height, (int) DSV_ROUND_POW2(width + 2 * ext, 4), 0, 0, height + 2 * ext);
    set_plane(&f->planes[1], format, cw, ch, (int) DSV_ROUND_POW2(cw + 2 * ext, 4), hs, vs, ch + 2 * ext);
    set_plane(&f->planes[2], format, cw, ch, (int) DSV_ROUND_POW2(cw + 2 * ext, 4), hs, vs, ch + 2 * ext);
    f->alloc = pinned_get((size_t) f->planes[0].len + f->planes[1].len + f->planes[2].len);
    uint8_t *at = f->alloc;
    for (int c = 0; c < 3; c++) {
        f->planes[c].data = at + f->planes[c].stride * ext + ext;
        at += f->planes[c].len;
    }
    return f;
}

extern "C" {

DSV_FRAME *dsv_frame_ref_inc(DSV_FRAME *frame)
{
    DSV_ASSERT(frame && frame->refcount > 0);
    frame->refcount++;
    return frame;
}

void dsv_frame_ref_dec(DSV_FRAME *frame)
{
    DSV_ASSERT(frame && frame->refcount > 0);
    if (--frame->refcount == 0) {
        if (frame->alloc && !pinned_put(frame->alloc)) {
            dsv_free(frame->alloc);
        }
        dsv_free(frame);
    }
}

DSV_FRAME *dsv_extend_frame(DSV_FRAME *frame)
{
    if (!frame->border) {
        return frame;
    }
    const int B = DSV_MAX_BLOCK_SIZE;
    for (int c = 0; c < 3; c++) {
        DSV_PLANE *p = &frame->planes[c];
        for (int y = 0; y < p->h; y++) {
            uint8_t *line = DSV_GET_LINE(p, y);
            memset(line - B, line[0], B);
            memset(line + p->w, line[p->w - 1], B);
        }
        for (int j = 1; j <= B; j++) {
            memcpy(DSV_GET_XY(p, -B, -j), DSV_GET_XY(p, -B, 0), p->w + 2 * B);
            memcpy(DSV_GET_XY(p, -B, p->h - 1 + j), DSV_GET_XY(p, -B, p->h - 1), p->w + 2 * B);
        }
    }
    return frame;
}

void dsv_frame_copy(DSV_FRAME *dst, DSV_FRAME *src)
{
    for (int c = 0; c < 3; c++) {
        DSV_PLANE *s = &src->planes[c], *d = &dst->planes[c];
        for (int y = 0; y < d->h; y++) {
            memcpy(DSV_GET_LINE(d, y), DSV_GET_LINE(s, y), s->w);
        }
    }
    if (dst->border) {
        dsv_extend_frame(dst);
    }
}

DSV_FRAME *dsv_clone_frame(DSV_FRAME *s, int border)
{
    DSV_FRAME *d = dsv_mk_frame(s->format, s->width, s->height, border);
    dsv_frame_copy(d, s);
    return d;
}

/* ---- the remaining frame helpers dsv.h declares (host frames, host arithmetic: they are part of the public header,
 * the codec itself does these steps in frame_ops.cu / sbt_inv.cu on the device) ------------------------------------ */
void dsv_frame_add(DSV_FRAME *dst, DSV_FRAME *src) /* bmc.c:304-316: dst = clamp(dst + src - 128) */
{
    for (int c = 0; c < 3; c++) {
        DSV_PLANE *s = &src->planes[c], *d = &dst->planes[c];
        for (int y = 0; y < d->h; y++) {
            uint8_t *o = DSV_GET_LINE(d, y);
            const uint8_t *a = DSV_GET_LINE(s, y);
            for (int x = 0; x < d->w; x++) {
                const int v = o[x] + a[x] - 128;
                o[x] = (uint8_t) (v < 0 ? 0 : (v > 255 ? 255 : v));
            }
        }
    }
}

int dsv_frame_avg_luma(DSV_FRAME *frame) /* frame.c:223-238 */
{
    const DSV_PLANE *p = &frame->planes[0];
    long long acc = 0;
    for (int y = 0; y < p->h; y++) {
        const uint8_t *line = DSV_GET_LINE(p, y);
        for (int x = 0; x < p->w; x++) {
            acc += line[x];
        }
    }
    return (int) (acc / ((long long) p->w * p->h));
}

void dsv_ds2x_frame_luma(DSV_FRAME *dst, DSV_FRAME *src) /* frame.c:240-261: rounded 2x2 mean, luma only */
{
    const DSV_PLANE *s = &src->planes[0];
    DSV_PLANE *d = &dst->planes[0];
    for (int y = 0; y < d->h; y++) {
        const uint8_t *r0 = DSV_GET_LINE(s, 2 * y), *r1 = r0 + s->stride; /* an odd source size reads its border */
        uint8_t *o = DSV_GET_LINE(d, y);
        for (int x = 0; x < d->w; x++) {
            o[x] = (uint8_t) ((r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1] + 2) >> 2);
        }
    }
}

DSV_FRAME *dsv_extend_frame_luma(DSV_FRAME *frame) /* frame.c:297-327 */
{
    if (!frame->border) {
        return frame;
    }
    const int B = DSV_MAX_BLOCK_SIZE;
    DSV_PLANE *p = &frame->planes[0];
    for (int y = 0; y < p->h; y++) {
        uint8_t *line = DSV_GET_LINE(p, y);
        memset(line - B, line[0], B);
        memset(line + p->w, line[p->w - 1], B);
    }
    for (int j = 1; j <= B; j++) {
        memcpy(DSV_GET_XY(p, -B, -j), DSV_GET_XY(p, -B, 0), p->w + 2 * B);
        memcpy(DSV_GET_XY(p, -B, p->h - 1 + j), DSV_GET_XY(p, -B, p->h - 1), p->w + 2 * B);
    }
    return frame;
}

void dsv_plane_xy(DSV_FRAME *frame, DSV_PLANE *out, int c, int x, int y) /* frame.c:329-342: a window into plane c */
{
    const DSV_PLANE *p = &frame->planes[c];
    out->format = p->format;
    out->data = DSV_GET_XY(p, x, y);
    out->stride = p->stride;
    out->w = p->w - x > 0 ? p->w - x : 0;
    out->h = p->h - y > 0 ? p->h - y : 0;
    out->hs = p->hs;
    out->vs = p->vs;
}

void dsv_mk_coefs(DSV_COEFS *c, int format, int width, int height)
{
    const int hs = DSV_FORMAT_H_SHIFT(format), vs = DSV_FORMAT_V_SHIFT(format);
    const int cw = (int) DSV_ROUND_POW2(DSV_ROUND_SHIFT(width, hs), 1);
    const int ch = (int) DSV_ROUND_POW2(DSV_ROUND_SHIFT(height, vs), 1);
    const int n0 = width * height, n1 = cw * ch;
    c[0].width = width;
    c[0].height = height;
    c[1].width = c[2].width = cw;
    c[1].height = c[2].height = ch;
    c[0].data = (DSV_SBC *) dsv_alloc((n0 + 2 * n1) * (int) sizeof(DSV_SBC));
    c[1].data = c[0].data + n0;
    c[2].data = c[1].data + n1;
}

/* ---- raw planar YUV files (dsv.c:98-170) ------------------------------------------------------ */
static size_t yuv_frame_bytes(int w, int h, int subsamp, size_t *chroma)
{
    size_t npix = (size_t) w * h, c = 0;
    switch (subsamp) {
        case DSV_SUBSAMP_444: c = npix; break;
        case DSV_SUBSAMP_422: c = (size_t) (w / 2) * h; break;
        case DSV_SUBSAMP_420:
        case DSV_SUBSAMP_411: c = npix / 4; break;
        default:
            DSV_ERROR(("unsupported format"));
            DSV_ASSERT(0);
    }
    *chroma = c;
    return npix + 2 * c;
}

int dsv_yuv_write(FILE *out, int fno, DSV_PLANE *p)
{
    if (!out || fno < 0) {
        return -1;
    }
    size_t fsz = (size_t) p[0].w * p[0].h + (size_t) p[1].w * p[1].h + (size_t) p[2].w * p[2].h;
    if (fseek(out, (long) (fsz * (size_t) fno), SEEK_SET)) {
        return -1;
    }
    for (int c = 0; c < 3; c++) {
        for (int y = 0; y < p[c].h; y++) {
            if (fwrite(DSV_GET_LINE(&p[c], y), (size_t) p[c].w, 1, out) != 1) {
                return -1;
            }
        }
    }
    return 0;
}

int dsv_yuv_read(FILE *in, int fno, uint8_t *o, int width, int height, int subsamp)
{
    if (!in || fno < 0) {
        return -1;
    }
    size_t chroma, fsz = yuv_frame_bytes(width, height, subsamp, &chroma);
    /* the reference seeks to fno*npix*{3, 2, 3/2}: identical to fno*fsz for even dimensions */
    size_t npix = (size_t) width * height, off;
    switch (subsamp) {
        case DSV_SUBSAMP_444: off = (size_t) fno * npix * 3; break;
        case DSV_SUBSAMP_422: off = (size_t) fno * npix * 2; break;
        default: off = (size_t) fno * npix * 3 / 2; break;
    }
    if (fseek(in, (long) off, SEEK_SET)) {
        return -1;
    }
    return fread(o, 1, fsz, in) == fsz ? 0 : -1;
}

/* ---- CLI helpers (util.c:21-93) ------------------------------------------------------------------ */
unsigned estimate_bitrate(int quality, int gop, DSV_META *md)
{
    int fps = (md->fps_num + md->fps_den / 2) / md->fps_den;
    int bpf = 352 * 288 * 3 / 2; /* bytes per CIF frame at 4:2:0 / 4:1:1 */
    if (md->subsamp == DSV_SUBSAMP_444) {
        bpf = 352 * 288 * 3;
    } else if (md->subsamp == DSV_SUBSAMP_422) {
        bpf = 352 * 288 * 2;
    }
    if (gop == DSV_GOP_INTRA) {
        bpf *= 4;
    }
    if (md->width < 320 && md->height < 240) {
        bpf /= 4;
    }
    int ratio = (((md->width + md->height) / 2) << 8) / 352;
    bpf = bpf * ratio >> 8;
    return (unsigned) (((bpf * fps) / (26 - quality / 4)) * 3 / 2);
}

void conv444to422(DSV_PLANE *s, DSV_PLANE *d)
{
    for (int y = 0; y < s->h; y++) {
        const uint8_t *sp = DSV_GET_LINE(s, y);
        uint8_t *dp = DSV_GET_LINE(d, y);
        for (int x = 0; x < s->w; x += 2) {
            int nx = x + 1 < s->w ? x + 1 : s->w - 1;
            dp[x >> 1] = (uint8_t) ((sp[x] + sp[nx] + 1) >> 1);
        }
    }
}

void conv422to420(DSV_PLANE *s, DSV_PLANE *d)
{
    for (int y = 0; y < s->h; y += 2) {
        int ny = y + 1 < s->h ? y + 1 : s->h - 1;
        const uint8_t *a = DSV_GET_LINE(s, y), *b = DSV_GET_LINE(s, ny);
        uint8_t *dp = DSV_GET_LINE(d, y >> 1);
        for (int x = 0; x < s->w; x++) {
            dp[x] = (uint8_t) ((a[x] + b[x] + 1) >> 1);
        }
    }
}

} /* extern "C" */
