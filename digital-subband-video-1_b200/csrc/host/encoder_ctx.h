/* encoder_ctx.h -- device-side state behind a DSV_ENCODER / DSV_DECODER (hung off their `ref` slots). */
#pragma once
#include <vector>

#include "../frame.cuh"
#include "../hzcc.cuh"
#include "../motion.cuh"
#include "../sbt.cuh"

namespace dsv {

struct CodecGeom {
    int w, h, subsamp, hs, vs;
    int pw[3], ph[3]; /* plane sizes */
    int cw[3], ch[3]; /* coefficient plane sizes (frame.c:29-61) */
    size_t coef_off[3], coef_total;
    size_t frame_bytes; /* packed planar frame */
    int blk_w, blk_h, nbh, nbv, nblk;
};

/* coefficient planes + kernel job tables of one picture in flight */
struct CoderBufs {
    int32_t *coef;
    int32_t *llx[3];
    int32_t *dv[3];
    uint8_t *d_stab;
    SbtJob *d_sjobs;
    HzJob *d_hjobs;
    HzChunk *d_chunks;
    HzFrame *d_frame;
    SbtJob sj[3];
    HzJob hj[3];
    int total_tiles, total_chunks;
    size_t lo_smem;
};

void plan_geometry(CodecGeom *g, int w, int h, int subsamp);
void plan_blocks(CodecGeom *g, int blk_w, int blk_h);
void coder_alloc(CoderBufs *c, const CodecGeom &g);
void coder_free(CoderBufs *c);
void coder_setup_jobs(CoderBufs *c, const CodecGeom &g, const DevFrame &pix, int quant, int isP, int do_quant,
                      cudaStream_t st);
void predict_mv(const DevMV *mvs, int nbh, int x, int y, int *px, int *py);

struct EncCtx {
    CodecGeom g;
    CoderBufs cb;
    cudaStream_t st = 0;
    bool inter = false;
    int have_ref = 0;
    int cur = 0;
    DevFrame xf, pred, pad[2], recon[2], pyr[2][5];
    DevMV *d_mvf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int2 *d_aux = nullptr;
    uint8_t *d_pkt = nullptr;
    uint8_t *d_misc = nullptr;
    size_t pkt_cap = 0;
    unsigned pkt_dirty = 0;
    uint8_t *h_in = nullptr, *h_pkt = nullptr, *h_misc = nullptr;
    DevMV *h_mv = nullptr;
    HzFrame *h_frame = nullptr;
};

} // namespace dsv
