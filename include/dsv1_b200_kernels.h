/*
 * dsv1_b200_kernels.h -- kernel-level C ABI of libdsv1_b200.so (HOST buffers in, HOST buffers out).
 *
 * One entry point per hot-path subsystem of the reference (dsv_internal.h:94-109, dsv_encoder.h:132),
 * with flat pointer+size signatures so that a test, a benchmark or a foreign-language binding can
 * drive a single subsystem.  Every call copies its inputs to the GPU, runs the CUDA kernels and
 * copies the result back; there is no host implementation behind any of these symbols.
 * The same signatures are implemented by the checker libraries (oracle/ref_harness.c: ref_*,
 * oracle/dsv1_port.c: port_*) -- those are test infrastructure, never linked into this library.
 *
 * Return value: 0 (or a byte/element count where stated) on success, negative on bad arguments.
 */
#ifndef DSV1_B200_KERNELS_H
#define DSV1_B200_KERNELS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* dsv_fwd_sbt (sbt.c:630-651): pix = ph rows of `stride` bytes with >= cw valid columns;
 * coef_out = cw*ch int32, dense.  cw, ch even and >= 16. */
int dsvk_fwd_sbt(const uint8_t *pix, int stride, int pw, int ph, int cw, int ch, int isP, int32_t *coef_out);

/* dsv_inv_sbt (sbt.c:653-714): c = plane index (0 = luma: smoothing-filtered inverse), q = frame quant. */
int dsvk_inv_sbt(int32_t *coef_io, int cw, int ch, int q, int isP, int c, uint8_t *pix_out, int stride, int pw, int ph);

/* dsv_fwd_sbt followed by the quantise + in-place dequantise half of hzcc_enc (hzcc.c:156-281), fused
 * as the encoder runs it.  dv_out (optional) receives the first-visit symbols of positions hzcc scans
 * twice; returns their count. */
int dsvk_fwd_sbt_q(const uint8_t *pix, int stride, int pw, int ph, int cw, int ch, int isP, int c, int q,
                   const uint8_t *stable, int nbh, int nbv, int32_t *coef_out, int32_t *dv_out);

/* dsv_encode_plane (hzcc.c:449-476): raw coefficients in, plane bytes (plen field first) out,
 * coef_io <- dequantised.  Returns the number of bytes written. */
int dsvk_encode_plane(int32_t *coef_io, int cw, int ch, int q, int isP, int c, const uint8_t *stable,
                      int nbh, int nbv, uint8_t *out, int out_cap);

/* dsv_decode_plane (hzcc.c:478-496): `in` points just after the 32-bit plen field. */
int dsvk_decode_plane(const uint8_t *in, int plen, int cw, int ch, int q, int isP, int c,
                      const uint8_t *stable, int nbh, int nbv, int32_t *coef_out);

/* mk_pyramid (dsv_encoder.c:194-217): luma pyramid levels 1..levels of a packed planar frame, packed
 * one after another into out; out_w/out_h receive each level's size. */
int dsvk_pyramid(const uint8_t *yuv, int w, int h, int subsamp, int levels, uint8_t *out, int *out_w, int *out_h);

/* dsv_hme (hme.c:730-741) on two packed planar ORIGINAL frames; mv_out = nbh*nbv DSV_MV records
 * (12 bytes each, dsv.h:137-150).  Returns the intra-block percentage. */
int dsvk_hme(const uint8_t *src_yuv, const uint8_t *ref_yuv, int w, int h, int subsamp, int blk_w, int blk_h,
             int levels, void *mv_out);

/* dsv_sub_pred (bmc.c:318-331): pred_out = prediction, resid_out = clamp(inp - pred + 128). */
int dsvk_sub_pred(const void *mvs, int w, int h, int subsamp, int blk_w, int blk_h, const uint8_t *inp_yuv,
                  const uint8_t *ref_yuv, uint8_t *pred_out, uint8_t *resid_out);

/* dsv_add_pred (bmc.c:333-346): out = clamp(pred + resid - 128). */
int dsvk_add_pred(const void *mvs, int w, int h, int subsamp, int blk_w, int blk_h, const uint8_t *resid_yuv,
                  const uint8_t *ref_yuv, uint8_t *out_yuv);

#ifdef __cplusplus
}
#endif
#endif /* DSV1_B200_KERNELS_H */
