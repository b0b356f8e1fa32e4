/*
 * tools/api_harness.c -- a CALLER of the dsv_encoder.h / dsv_decoder.h API.
 *
 * This file is the drop-in proof: the same source is compiled twice,
 *   (a) by oracle/Makefile against the reference's own headers and objects
 *       (-I/root/reference, symbols prefixed ref_)  -> oracle/_ref/libdsv1ref.so
 *   (b) by the product build against include/compat/*.h and libdsv1_b200.so
 *       (symbols prefixed dsvh_)
 * and does exactly what the reference CLI's encode()/decode() loops do
 * (dsv_main.c:423-560 and dsv_main.c:567-721) but memory-to-memory, so that
 * tests and bench.py can drive either implementation through ctypes.
 *
 * Encoder fields are set the way the CLI sets them (dsv_main.c:463-489); the
 * caller passes already-converted values (quality 0..2047, rc_mode as in
 * dsv_encoder.h:32-33, i.e. 0 = CRF, 1 = ABR).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "dsv.h"
#include "dsv_encoder.h"
#include "dsv_decoder.h"
#include "util.h"

#ifndef HPFX
#error "compile with -DHPFX=ref_ or -DHPFX=dsvh_"
#endif
#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
#define FN(name) CAT(HPFX, name)

enum {
    CFG_W, CFG_H, CFG_SUBSAMP, CFG_FPS_NUM, CFG_FPS_DEN, CFG_ASPECT_NUM, CFG_ASPECT_DEN,
    CFG_GOP, CFG_QUALITY, CFG_RC_MODE, CFG_BITRATE, CFG_DO_SCD, CFG_SCD_DELTA, CFG_INTRA_PCT,
    CFG_PYR_LEVELS, CFG_STABLE_REFRESH, CFG_MAX_Q_STEP, CFG_MIN_QUALITY, CFG_MAX_QUALITY,
    CFG_MIN_I_QUALITY, CFG_HM_NUDGE, CFG_COUNT
};

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

static long frame_bytes(int w, int h, int subsamp)
{
    int hs = DSV_FORMAT_H_SHIFT(subsamp), vs = DSV_FORMAT_V_SHIFT(subsamp);
    long cw = (w + (1 << hs) - 1) >> hs, ch = (h + (1 << vs) - 1) >> vs;
    return (long) w * h + 2 * cw * ch;
}

int FN(cfg_count)(void) { return CFG_COUNT; }

/*
 * Encode nframes packed planar frames; the concatenated packets (META/PIC.../EOS)
 * are appended to out.  Returns the stream length, or -1 if out_cap is too small.
 * pkt_len (optional, 2*nframes+1 entries) receives each packet's length,
 * *npkt their count, *seconds the wall time spent inside dsv_enc calls.
 */
long
FN(encode_sequence)(const int *cfg, const uint8_t *yuv, int nframes,
                    uint8_t *out, long out_cap, int *pkt_len, int *npkt, double *seconds)
{
    DSV_ENCODER enc;
    DSV_META md;
    DSV_BUF bufs[4];
    long fsz, pos = 0;
    int f, i, n, np = 0;
    double t0, acc = 0.0;
    uint8_t *picture;

    memset(&md, 0, sizeof(md));
    md.width = cfg[CFG_W];
    md.height = cfg[CFG_H];
    md.subsamp = cfg[CFG_SUBSAMP];
    md.fps_num = cfg[CFG_FPS_NUM];
    md.fps_den = cfg[CFG_FPS_DEN];
    md.aspect_num = cfg[CFG_ASPECT_NUM];
    md.aspect_den = cfg[CFG_ASPECT_DEN];
    fsz = frame_bytes(md.width, md.height, md.subsamp);

    dsv_enc_init(&enc);
    dsv_enc_set_metadata(&enc, &md);
    enc.gop = cfg[CFG_GOP];
    enc.scene_change_delta = cfg[CFG_SCD_DELTA];
    enc.do_scd = cfg[CFG_DO_SCD];
    enc.intra_pct_thresh = cfg[CFG_INTRA_PCT];
    enc.quality = cfg[CFG_QUALITY];
    enc.rc_mode = cfg[CFG_RC_MODE];
    enc.bitrate = (unsigned) cfg[CFG_BITRATE];
    enc.max_q_step = cfg[CFG_MAX_Q_STEP];
    enc.min_quality = cfg[CFG_MIN_QUALITY];
    enc.max_quality = cfg[CFG_MAX_QUALITY];
    enc.min_I_frame_quality = cfg[CFG_MIN_I_QUALITY];
    enc.rc_high_motion_nudge = cfg[CFG_HM_NUDGE];
    enc.pyramid_levels = cfg[CFG_PYR_LEVELS];
    enc.stable_refresh = (unsigned) cfg[CFG_STABLE_REFRESH];
    dsv_enc_start(&enc);

    /* the CLI reuses one picture buffer for every frame (dsv_main.c:461,511-515) */
    picture = (uint8_t *) malloc((size_t) fsz + 64);
    for (f = 0; f < nframes; f++) {
        DSV_FRAME *frame;
        memcpy(picture, yuv + (long) f * fsz, (size_t) fsz);
        frame = dsv_load_planar_frame(md.subsamp, picture, md.width, md.height);
        t0 = now_s();
        n = dsv_enc(&enc, frame, bufs) & DSV_ENC_NUM_BUFS;
        acc += now_s() - t0;
        for (i = 0; i < n; i++) {
            if (pos + (long) bufs[i].len > out_cap) {
                pos = -1;
            } else if (pos >= 0) {
                memcpy(out + pos, bufs[i].data, bufs[i].len);
                pos += bufs[i].len;
                if (pkt_len) {
                    pkt_len[np] = (int) bufs[i].len;
                }
                np++;
            }
            dsv_buf_free(&bufs[i]);
        }
    }
    dsv_enc_end_of_stream(&enc, bufs);
    if (pos >= 0 && pos + (long) bufs[0].len <= out_cap) {
        memcpy(out + pos, bufs[0].data, bufs[0].len);
        pos += bufs[0].len;
        if (pkt_len) {
            pkt_len[np] = (int) bufs[0].len;
        }
        np++;
    } else {
        pos = -1;
    }
    dsv_buf_free(&bufs[0]);
    dsv_enc_free(&enc);
    free(picture);
    if (npkt) {
        *npkt = np;
    }
    if (seconds) {
        *seconds = acc;
    }
    return pos;
}

/*
 * Decode a whole in-memory .dsv stream.  Frames are written packed planar at
 * offset fnum*frame_bytes (dsv.c:98-129 semantics).  Returns the number of
 * frames decoded, or a negative value on a malformed container.
 * meta_out (7 ints) receives width,height,subsamp,fps_num,fps_den,aspect_num,aspect_den.
 */
int
FN(decode_stream_ex)(const uint8_t *stream, long len, uint8_t *yuv_out, long out_cap,
                     int *meta_out, double *seconds, int draw_info, int to_420p)
{
    DSV_DECODER dec;
    DSV_META *meta = NULL;
    long pos = 0;
    int nfr = 0;
    double t0, acc = 0.0;

    memset(&dec, 0, sizeof(dec));
    dec.draw_info = draw_info; /* dsv_main.c:639 */
    while (pos + DSV_PACKET_HDR_SIZE <= len) {
        const uint8_t *hdr = stream + pos;
        DSV_BUF buffer;
        DSV_FRAME *frame = NULL;
        DSV_FNUM fno = 0;
        long size;
        int code;

        if (hdr[0] != DSV_FOURCC_0 || hdr[1] != DSV_FOURCC_1 || hdr[2] != DSV_FOURCC_2 || hdr[3] != DSV_FOURCC_3) {
            nfr = -4;
            break;
        }
        size = ((long) hdr[DSV_PACKET_NEXT_OFFSET] << 24) | ((long) hdr[DSV_PACKET_NEXT_OFFSET + 1] << 16)
             | ((long) hdr[DSV_PACKET_NEXT_OFFSET + 2] << 8) | (long) hdr[DSV_PACKET_NEXT_OFFSET + 3];
        if (size == 0) {
            size = DSV_PACKET_HDR_SIZE;
        }
        if (size < DSV_PACKET_HDR_SIZE || pos + size > len) {
            nfr = -3;
            break;
        }
        dsv_mk_buf(&buffer, (int) size);
        memcpy(buffer.data, hdr, (size_t) size);
        pos += size;

        t0 = now_s();
        code = dsv_dec(&dec, &buffer, &frame, &fno);
        acc += now_s() - t0;

        if (code == DSV_DEC_GOT_META) {
            if (!meta) {
                meta = dsv_get_metadata(&dec);
            }
            continue;
        }
        if (code == DSV_DEC_EOS) {
            break;
        }
        if (code != DSV_DEC_OK || frame == NULL || meta == NULL) {
            continue;
        }
        {
            /* -out420p (dsv_main.c:674-699): chroma goes 444 -> 422 -> 420 through the util.c filters */
            const int conv = to_420p && meta->subsamp != DSV_SUBSAMP_420;
            DSV_FRAME *f420 = NULL;
            long fsz = frame_bytes(meta->width, meta->height, conv ? DSV_SUBSAMP_420 : meta->subsamp);
            long off = (long) fno * fsz;
            int c, y;
            if (conv) {
                f420 = dsv_mk_frame(DSV_SUBSAMP_420, frame->width, frame->height, 0);
                if (meta->subsamp == DSV_SUBSAMP_444) {
                    DSV_FRAME *f422 = dsv_mk_frame(DSV_SUBSAMP_422, frame->width, frame->height, 0);
                    for (c = 1; c < 3; c++) {
                        conv444to422(&frame->planes[c], &f422->planes[c]);
                        conv422to420(&f422->planes[c], &f420->planes[c]);
                    }
                    dsv_frame_ref_dec(f422);
                } else {
                    for (c = 1; c < 3; c++) {
                        conv422to420(&frame->planes[c], &f420->planes[c]);
                    }
                }
            }
            if (off + fsz <= out_cap) {
                uint8_t *o = yuv_out + off;
                for (c = 0; c < 3; c++) {
                    DSV_PLANE *p = (conv && c > 0) ? &f420->planes[c] : &frame->planes[c];
                    for (y = 0; y < p->h; y++) {
                        memcpy(o, DSV_GET_LINE(p, y), (size_t) p->w);
                        o += p->w;
                    }
                }
                nfr++;
            }
            if (f420) {
                dsv_frame_ref_dec(f420);
            }
        }
        dsv_frame_ref_dec(frame);
    }
    dsv_dec_free(&dec);
    if (meta) {
        if (meta_out) {
            meta_out[0] = meta->width;
            meta_out[1] = meta->height;
            meta_out[2] = meta->subsamp;
            meta_out[3] = meta->fps_num;
            meta_out[4] = meta->fps_den;
            meta_out[5] = meta->aspect_num;
            meta_out[6] = meta->aspect_den;
        }
        dsv_free(meta);
    }
    if (seconds) {
        *seconds = acc;
    }
    return nfr;
}

int
FN(decode_stream)(const uint8_t *stream, long len, uint8_t *yuv_out, long out_cap,
                  int *meta_out, double *seconds)
{
    return FN(decode_stream_ex)(stream, len, yuv_out, out_cap, meta_out, seconds, 0, 0);
}

/*
 * dsv_hme through its exported interface (dsv_encoder.h:122-132): the caller owns the pyramids.  Builds bordered
 * frames for `levels` + 1 levels of src and ref (2x2 rounded box filter on luma, like the encoder's pyramid), runs
 * dsv_hme and returns the intra percentage; mv_out receives the vector fields of all levels, level 0 first
 * ((levels + 1) * nblocks * sizeof(DSV_MV) bytes).
 */
static DSV_FRAME *harness_half(DSV_FRAME *prev)
{
    int w = (prev->width + 1) / 2, h = (prev->height + 1) / 2, x, y;
    DSV_FRAME *f = dsv_mk_frame(prev->format, w, h, 1);
    DSV_PLANE *s = &prev->planes[0], *d = &f->planes[0];
    for (y = 0; y < h; y++) {
        for (x = 0; x < w; x++) {
            /* may read one sample into the (replicated) border of the level above */
            int a = DSV_GET_LINE(s, 2 * y)[2 * x], b = DSV_GET_LINE(s, 2 * y)[2 * x + 1];
            int c = DSV_GET_LINE(s, 2 * y + 1)[2 * x], e = DSV_GET_LINE(s, 2 * y + 1)[2 * x + 1];
            DSV_GET_LINE(d, y)[x] = (uint8_t) ((a + b + c + e + 2) >> 2);
        }
    }
    return dsv_extend_frame(f);
}

static DSV_FRAME *harness_bordered(const uint8_t *yuv, int w, int h, int subsamp)
{
    DSV_FRAME *wrap = dsv_load_planar_frame(subsamp, (void *) yuv, w, h);
    DSV_FRAME *f = dsv_clone_frame(wrap, 1);
    dsv_frame_ref_dec(wrap);
    return f;
}

int
FN(hme_api)(const uint8_t *src_yuv, const uint8_t *ref_yuv, int w, int h, int subsamp, int blk_w, int blk_h,
            int levels, uint8_t *mv_out)
{
    DSV_HME hme;
    DSV_PARAMS params;
    DSV_META meta;
    int i, pct, nblk;

    if (levels < 0 || levels > DSV_MAX_PYRAMID_LEVELS) {
        return -1;
    }
    memset(&hme, 0, sizeof(hme));
    memset(&params, 0, sizeof(params));
    memset(&meta, 0, sizeof(meta));
    meta.width = w;
    meta.height = h;
    meta.subsamp = subsamp;
    params.vidmeta = &meta;
    params.has_ref = 1;
    params.is_ref = 1;
    params.blk_w = blk_w;
    params.blk_h = blk_h;
    params.nblocks_h = (w + blk_w - 1) / blk_w;
    params.nblocks_v = (h + blk_h - 1) / blk_h;
    nblk = params.nblocks_h * params.nblocks_v;
    hme.params = &params;
    hme.levels = levels;
    hme.src[0] = harness_bordered(src_yuv, w, h, subsamp);
    hme.ref[0] = harness_bordered(ref_yuv, w, h, subsamp);
    for (i = 1; i <= levels; i++) {
        hme.src[i] = harness_half(hme.src[i - 1]);
        hme.ref[i] = harness_half(hme.ref[i - 1]);
    }
    pct = dsv_hme(&hme);
    for (i = 0; i <= levels; i++) {
        memcpy(mv_out + (size_t) i * nblk * sizeof(DSV_MV), hme.mvf[i], (size_t) nblk * sizeof(DSV_MV));
        dsv_free(hme.mvf[i]);
        dsv_frame_ref_dec(hme.src[i]);
        dsv_frame_ref_dec(hme.ref[i]);
    }
    return pct;
}

/*
 * The frame helpers of dsv.h that the codec path itself does not go through (dsv_ds2x_frame_luma,
 * dsv_extend_frame_luma, dsv_frame_avg_luma, dsv_frame_add, dsv_plane_xy): exercised on one picture, everything
 * they produce is written to out so that the two libraries can be compared byte for byte.  Host-only: needs no GPU.
 * Returns the number of bytes written (or -1 if out_cap is too small).
 */
long
FN(frame_helpers_probe)(const uint8_t *yuv, int w, int h, int subsamp, uint8_t *out, long out_cap)
{
    DSV_FRAME *a = harness_bordered(yuv, w, h, subsamp);
    DSV_FRAME *half = dsv_mk_frame(subsamp, (w + 1) / 2, (h + 1) / 2, 1);
    DSV_FRAME *sum = dsv_clone_frame(a, 0), *other = dsv_clone_frame(a, 0);
    DSV_PLANE win, *hp = &half->planes[0];
    long n = 0;
    int c, y, meta[8];
    int B = 64; /* DSV_FRAME_BORDER */

    dsv_ds2x_frame_luma(half, a);
    dsv_extend_frame_luma(half);
    /* make `other` differ from `sum` so that both clamps of the add are reached */
    for (y = 0; y < other->planes[0].h; y++) {
        uint8_t *line = DSV_GET_LINE(&other->planes[0], y);
        int x;
        for (x = 0; x < other->planes[0].w; x++) {
            line[x] = (uint8_t) (255 - line[x] + ((x ^ y) & 63));
        }
    }
    dsv_frame_add(sum, other);
    memset(&win, 0, sizeof(win));
    dsv_plane_xy(a, &win, 1, 3, 2);
    meta[0] = dsv_frame_avg_luma(a);
    meta[1] = dsv_frame_avg_luma(half);
    meta[2] = win.w;
    meta[3] = win.h;
    meta[4] = win.stride;
    meta[5] = (int) (win.data - a->planes[1].data);
    meta[6] = win.hs * 16 + win.vs;
    meta[7] = win.format;
    if ((long) sizeof(meta) + (long) (hp->w + 2 * B) * (hp->h + 2 * B) + frame_bytes(w, h, subsamp) > out_cap) {
        n = -1;
    } else {
        memcpy(out, meta, sizeof(meta));
        n = (long) sizeof(meta);
        for (y = -B; y < hp->h + B; y++) { /* the half-size luma with its border */
            memcpy(out + n, DSV_GET_XY(hp, -B, y), (size_t) (hp->w + 2 * B));
            n += hp->w + 2 * B;
        }
        for (c = 0; c < 3; c++) {
            DSV_PLANE *p = &sum->planes[c];
            for (y = 0; y < p->h; y++) {
                memcpy(out + n, DSV_GET_LINE(p, y), (size_t) p->w);
                n += p->w;
            }
        }
    }
    dsv_frame_ref_dec(a);
    dsv_frame_ref_dec(half);
    dsv_frame_ref_dec(sum);
    dsv_frame_ref_dec(other);
    return n;
}
