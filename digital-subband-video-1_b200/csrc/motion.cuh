/*
 * motion.cuh -- interfaces of the motion kernels: hierarchical motion estimation (hme.cu, replaces
 * hme.c:378-741) and half-pel block motion compensation (bmc.cu, replaces bmc.c:204-346).
 */
#pragma once
#include "frame.cuh"

namespace dsv {

/* device twin of DSV_MV (dsv.h:137-150): 12 bytes */
struct DevMV {
    int16_t x, y;
    uint8_t mode, submask, lo_var, lo_tex, high_detail;
    uint8_t pad[3];
};

struct MotionGeom {
    int w, h;           /* luma size */
    int hs, vs;         /* chroma shifts */
    int blk_w, blk_h;
    int nbh, nbv;
    int levels;         /* pyramid levels (dsv_encoder.c:602-613) */
};

/*
 * src[0]/ref[0]: full-size padded ORIGINAL frames (luma + chroma); src[i]/ref[i], i >= 1: pyramid level i
 * (luma only).  mvf[l]: nbh*nbv DevMV scratch per level; the result is mvf[0].  aux: nbh*nbv int2 scratch.
 * *d_nintra receives the number of intra blocks at level 0 (hme.c:727,740).
 */
void hme_launch(const MotionGeom &g, const DevFrame *src, const DevFrame *ref, DevMV *const *mvf, int2 *aux,
                int *d_nintra, cudaStream_t st);

/* prediction from `ref` into `pred` (may be null: not kept) for all three planes, fused with
 *   mode 1 (encoder, dsv_sub_pred):  io = clamp(io - pred + 128)
 *   mode 2 (decoder, dsv_add_pred):  io = clamp(pred + io - 128)  (io holds the residual on entry) */
void bmc_launch(const MotionGeom &g, const DevMV *mv, const DevFrame &ref, const DevFrame *pred, const DevFrame &io,
                int mode, cudaStream_t st);

/* dst = clamp(dst + src - 128) on w x h of every plane (dsv_frame_add, bmc.c:304-316) */
void frame_add_launch(const DevFrame &dst, const DevFrame &src, cudaStream_t st);

} // namespace dsv
