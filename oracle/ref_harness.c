/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE (not product code).
 *
 * Flat (ctypes-friendly) entry points into the UNMODIFIED reference, compiled
 * by oracle/Makefile from the sources where they lie under /root/reference
 * (-I/root/reference; nothing is copied).  The result is oracle/_ref/libdsv1ref.so.
 *
 * Every function here only marshals arguments into the reference's own structs
 * and calls the reference's own per-subsystem entry points
 * (dsv_internal.h:94-109, dsv_encoder.h:132).  The same flat signatures are
 * implemented by oracle/dsv1_port.c (port_*) and by the CUDA library (dsvk_*),
 * so a parity test is "call three libraries with the same arguments".
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "dsv.h"
#include "dsv_internal.h"
#include "dsv_encoder.h"
#include "dsv_decoder.h"

/* ---- helpers ------------------------------------------------------------ */

static int fmt_hs(int subsamp) { return DSV_FORMAT_H_SHIFT(subsamp); }
static int fmt_vs(int subsamp) { return DSV_FORMAT_V_SHIFT(subsamp); }

/* bordered reference frame filled from packed planar YUV, borders extended */
static DSV_FRAME *
frame_from_packed(const uint8_t *yuv, int w, int h, int subsamp, int border)
{
    DSV_FRAME *src = dsv_load_planar_frame(subsamp, (void *) yuv, w, h);
    DSV_FRAME *f = dsv_mk_frame(subsamp, w, h, border);
    dsv_frame_copy(f, src); /* extends when f has a border (frame.c:218-220) */
    dsv_frame_ref_dec(src);
    return f;
}

static void
frame_to_packed(DSV_FRAME *f, uint8_t *yuv)
{
    int c, y;
    for (c = 0; c < 3; c++) {
        DSV_PLANE *p = &f->planes[c];
        for (y = 0; y < p->h; y++) {
            memcpy(yuv, DSV_GET_LINE(p, y), p->w);
            yuv += p->w;
        }
    }
}

/* ---- (1) subband transform --------------------------------------------- */

/* pix: ph rows of `stride` bytes, at least cw valid columns (sbt.c:583-591 reads cw) */
int
ref_fwd_sbt(const uint8_t *pix, int stride, int pw, int ph, int cw, int ch, int isP, int32_t *coef_out)
{
    DSV_PLANE p;
    DSV_COEFS c;
    memset(&p, 0, sizeof(p));
    p.data = (uint8_t *) pix;
    p.stride = stride;
    p.w = pw;
    p.h = ph;
    c.width = cw;
    c.height = ch;
    c.data = coef_out;
    memset(coef_out, 0, sizeof(int32_t) * cw * ch); /* dsv_mk_coefs zeroes (frame.c:58) */
    dsv_fwd_sbt(&p, &c, isP);
    return 0;
}

/* coef is clobbered exactly as the reference clobbers it */
int
ref_inv_sbt(int32_t *coef, int cw, int ch, int q, int isP, int c, uint8_t *pix_out, int stride, int pw, int ph)
{
    DSV_PLANE p;
    DSV_COEFS co;
    memset(&p, 0, sizeof(p));
    p.data = pix_out;
    p.stride = stride;
    p.w = pw;
    p.h = ph;
    co.width = cw;
    co.height = ch;
    co.data = coef;
    dsv_inv_sbt(&p, &co, q, isP, c);
    return 0;
}

int ref_get_quant(int q, int isP, int level) { return dsv_get_quant(q, isP, level); }
int ref_lb2(unsigned n) { return dsv_lb2(n); }

/* ---- (4) quantiser + HZCC ------------------------------------------------ */

static void
mk_stab(DSV_STABILITY *stab, DSV_PARAMS *prm, DSV_META *md, const uint8_t *stable, int nbh, int nbv, int isP, int c)
{
    memset(prm, 0, sizeof(*prm));
    memset(md, 0, sizeof(*md));
    prm->vidmeta = md;
    prm->nblocks_h = nbh;
    prm->nblocks_v = nbv;
    stab->params = prm;
    stab->stable_blocks = (unsigned char *) stable;
    stab->cur_plane = (unsigned char) c;
    stab->isP = (unsigned char) isP;
}

/* returns number of bytes written to out (out must be zeroed, out_cap large enough);
 * coef is updated in place with the dequantised values (hzcc.c:172-184 etc.) */
int
ref_encode_plane(int32_t *coef, int cw, int ch, int q, int isP, int c,
                 const uint8_t *stable, int nbh, int nbv, uint8_t *out, int out_cap)
{
    DSV_STABILITY stab;
    DSV_PARAMS prm;
    DSV_META md;
    DSV_COEFS co;
    DSV_BS bs;
    (void) out_cap;
    mk_stab(&stab, &prm, &md, stable, nbh, nbv, isP, c);
    co.width = cw;
    co.height = ch;
    co.data = coef;
    dsv_bs_init(&bs, out);
    dsv_encode_plane(&bs, &co, q, &stab);
    return dsv_bs_ptr(&bs);
}

/* `in` points just after the 32-bit plen field, as in dsv_decoder.c:402-407 */
int
ref_decode_plane(const uint8_t *in, int plen, int cw, int ch, int q, int isP, int c,
                 const uint8_t *stable, int nbh, int nbv, int32_t *coef_out)
{
    DSV_STABILITY stab;
    DSV_PARAMS prm;
    DSV_META md;
    DSV_COEFS co;
    mk_stab(&stab, &prm, &md, stable, nbh, nbv, isP, c);
    co.width = cw;
    co.height = ch;
    co.data = coef_out;
    memset(coef_out, 0, sizeof(int32_t) * cw * ch); /* dsv_decoder.c:405 */
    dsv_decode_plane((uint8_t *) in, plen, &co, q, &stab);
    return 0;
}

/* ---- (2) pyramid + hierarchical motion estimation ----------------------- */

/* One pyramid level (luma only), output packed w'*h' (frame.c:240-261 + dsv_encoder.c:194-217) */
int
ref_pyramid(const uint8_t *yuv, int w, int h, int subsamp, int levels, uint8_t *out, int *out_w, int *out_h)
{
    DSV_FRAME *prev = frame_from_packed(yuv, w, h, subsamp, 1);
    int i, y;
    dsv_extend_frame(prev);
    for (i = 0; i < levels; i++) {
        DSV_FRAME *f = dsv_mk_frame(subsamp, DSV_ROUND_SHIFT(w, i + 1), DSV_ROUND_SHIFT(h, i + 1), 1);
        dsv_ds2x_frame_luma(f, prev);
        dsv_extend_frame_luma(f);
        for (y = 0; y < f->planes[0].h; y++) {
            memcpy(out, DSV_GET_LINE(&f->planes[0], y), f->planes[0].w);
            out += f->planes[0].w;
        }
        out_w[i] = f->planes[0].w;
        out_h[i] = f->planes[0].h;
        dsv_frame_ref_dec(prev);
        prev = f;
    }
    dsv_frame_ref_dec(prev);
    return 0;
}

int
ref_avg_luma(const uint8_t *y, int w, int h)
{
    DSV_FRAME f;
    memset(&f, 0, sizeof(f));
    f.planes[0].data = (uint8_t *) y;
    f.planes[0].stride = w;
    f.planes[0].w = w;
    f.planes[0].h = h;
    return dsv_frame_avg_luma(&f);
}

/* src/ref: packed planar ORIGINAL frames (dsv_encoder.c:231-236).
 * mv_out: nbh*nbv DSV_MV records (12 bytes each, dsv.h:137-150).
 * returns intra percentage (hme.c:740). */
int
ref_hme(const uint8_t *src_yuv, const uint8_t *ref_yuv, int w, int h, int subsamp,
        int blk_w, int blk_h, int levels, void *mv_out)
{
    DSV_PARAMS prm;
    DSV_META md;
    DSV_HME hme;
    DSV_FRAME *sf[DSV_MAX_PYRAMID_LEVELS + 1], *rf[DSV_MAX_PYRAMID_LEVELS + 1];
    int i, pct, nb;

    memset(&prm, 0, sizeof(prm));
    memset(&md, 0, sizeof(md));
    md.width = w;
    md.height = h;
    md.subsamp = subsamp;
    prm.vidmeta = &md;
    prm.blk_w = blk_w;
    prm.blk_h = blk_h;
    prm.nblocks_h = DSV_DIV_ROUND(w, blk_w);
    prm.nblocks_v = DSV_DIV_ROUND(h, blk_h);
    nb = prm.nblocks_h * prm.nblocks_v;

    sf[0] = frame_from_packed(src_yuv, w, h, subsamp, 1);
    rf[0] = frame_from_packed(ref_yuv, w, h, subsamp, 1);
    for (i = 0; i < levels; i++) {
        sf[i + 1] = dsv_mk_frame(subsamp, DSV_ROUND_SHIFT(w, i + 1), DSV_ROUND_SHIFT(h, i + 1), 1);
        dsv_ds2x_frame_luma(sf[i + 1], sf[i]);
        dsv_extend_frame_luma(sf[i + 1]);
        rf[i + 1] = dsv_mk_frame(subsamp, DSV_ROUND_SHIFT(w, i + 1), DSV_ROUND_SHIFT(h, i + 1), 1);
        dsv_ds2x_frame_luma(rf[i + 1], rf[i]);
        dsv_extend_frame_luma(rf[i + 1]);
    }
    memset(&hme, 0, sizeof(hme));
    hme.levels = levels;
    hme.params = &prm;
    for (i = 0; i <= levels; i++) {
        hme.src[i] = sf[i];
        hme.ref[i] = rf[i];
    }
    pct = dsv_hme(&hme);
    memcpy(mv_out, hme.mvf[0], sizeof(DSV_MV) * nb);
    for (i = 0; i <= levels; i++) {
        dsv_free(hme.mvf[i]);
        dsv_frame_ref_dec(sf[i]);
        dsv_frame_ref_dec(rf[i]);
    }
    return pct;
}

int ref_sizeof_mv(void) { return (int) sizeof(DSV_MV); }

/* ---- (3) block motion compensation -------------------------------------- */

static void
mk_params(DSV_PARAMS *prm, DSV_META *md, int w, int h, int subsamp, int blk_w, int blk_h)
{
    memset(prm, 0, sizeof(*prm));
    memset(md, 0, sizeof(*md));
    md->width = w;
    md->height = h;
    md->subsamp = subsamp;
    prm->vidmeta = md;
    prm->blk_w = blk_w;
    prm->blk_h = blk_h;
    prm->nblocks_h = DSV_DIV_ROUND(w, blk_w);
    prm->nblocks_v = DSV_DIV_ROUND(h, blk_h);
}

/* encoder side (dsv_encoder.c:657-660): pred_out = prediction, resid_out = clamp(inp - pred + 128) */
int
ref_sub_pred(const void *mvs, int w, int h, int subsamp, int blk_w, int blk_h,
             const uint8_t *inp_yuv, const uint8_t *ref_yuv, uint8_t *pred_out, uint8_t *resid_out)
{
    DSV_PARAMS prm;
    DSV_META md;
    DSV_FRAME *inp, *ref, *dif;
    mk_params(&prm, &md, w, h, subsamp, blk_w, blk_h);
    inp = frame_from_packed(inp_yuv, w, h, subsamp, 1);
    ref = frame_from_packed(ref_yuv, w, h, subsamp, 1);
    dif = dsv_mk_frame(subsamp, w, h, 1);
    dsv_sub_pred((DSV_MV *) mvs, &prm, dif, inp, ref);
    frame_to_packed(dif, pred_out);
    frame_to_packed(inp, resid_out);
    dsv_frame_ref_dec(inp);
    dsv_frame_ref_dec(ref);
    dsv_frame_ref_dec(dif);
    return 0;
}

/* decoder side (dsv_decoder.c:432): out = clamp(pred + resid - 128) */
int
ref_add_pred(const void *mvs, int w, int h, int subsamp, int blk_w, int blk_h,
             const uint8_t *resid_yuv, const uint8_t *ref_yuv, uint8_t *out_yuv)
{
    DSV_PARAMS prm;
    DSV_META md;
    DSV_FRAME *res, *ref, *out;
    mk_params(&prm, &md, w, h, subsamp, blk_w, blk_h);
    res = frame_from_packed(resid_yuv, w, h, subsamp, 1);
    ref = frame_from_packed(ref_yuv, w, h, subsamp, 1);
    out = dsv_mk_frame(subsamp, w, h, 1);
    dsv_add_pred((DSV_MV *) mvs, &prm, res, out, ref);
    frame_to_packed(out, out_yuv);
    dsv_frame_ref_dec(res);
    dsv_frame_ref_dec(ref);
    dsv_frame_ref_dec(out);
    return 0;
}
