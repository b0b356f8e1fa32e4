/*
 * motion.cuh -- interfaces of the motion kernels: hierarchical motion estimation (hme.cu, replaces
 * hme.c:378-741) and half-pel block motion compensation (bmc.cu, replaces bmc.c:204-346).
 * Both are batched over lanes (independent sequences in lock step): one argument record per lane in a
 * device array, lane = blockIdx.z.
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

struct HmePlane {
    const uint8_t *p;
    int stride, w, h;
};
struct HmeArgs {
    HmePlane src, ref;               /* luma at this level */
    HmePlane srcU, srcV, refU, refV; /* level 0 only */
    const DevMV *parent;             /* level + 1 field or null */
    DevMV *out;
    int2 *aux;                       /* level 0: (luma_tex, src_var) per block for the neighbour pass */
    int *nintra;
    int level, blk_w, blk_h, nbh, nbv, hs, vs;
};

/*
 * src[0]/ref[0]: full-size padded ORIGINAL frames (luma + chroma); src[i]/ref[i], i >= 1: pyramid level i
 * (luma only).  mvf[l]: nbh*nbv DevMV scratch per level; the result is mvf[0].  aux: nbh*nbv int2 scratch.
 * *d_nintra receives the number of intra blocks at level 0 (hme.c:727,740).
 */
void hme_fill_args(HmeArgs *A, const MotionGeom &g, int level, const DevFrame *src, const DevFrame *ref,
                   DevMV *const *mvf, int2 *aux, int *d_nintra);
/* d_args: device array [(levels + 1) x n], entry [level * n + lane] */
void hme_launch(const HmeArgs *d_args, int n, const MotionGeom &g, cudaStream_t st);

struct BmcPlane {
    const uint8_t *ref;
    uint8_t *pred; /* may be null: prediction not kept (decoder) */
    const uint8_t *in;
    uint8_t *out;
    int rstride, pstride, istride, ostride;
    int w, h;
};
struct BmcArgs {
    BmcPlane pl[3];
    const DevMV *mv;
    int blk_w, blk_h, nbh, nbv, hs, vs, mode;
};
/* prediction from `ref` (kept in `pred` unless null) for all three planes, fused with
 *   mode 1 (encoder, dsv_sub_pred):  out = clamp(in - pred + 128)
 *   mode 2 (decoder, dsv_add_pred):  out = clamp(pred + in - 128)   (in may alias out) */
void bmc_fill_args(BmcArgs *a, const MotionGeom &g, const DevMV *d_mv, const DevFrame &ref, const DevFrame *pred,
                   const DevFrame &in, const DevFrame &out, int mode);
void bmc_launch(const BmcArgs *d_args, int n, const MotionGeom &g, cudaStream_t st);

} // namespace dsv
