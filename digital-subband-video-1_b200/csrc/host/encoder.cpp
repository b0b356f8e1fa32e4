/*
 * encoder.cpp -- the dsv_encoder.h API (dsv_encoder.c:696-854) on top of the CUDA kernels.
 *
 * Host code here does only what is serial and tiny in the reference: GOP / frame-type decisions
 * (dsv_encoder.c:575-694), the CRF quality->quant map (dsv_encoder.c:162-165), the stability tracker
 * and its ZBRLE map (dsv_encoder.c:329-408), motion-vector prediction + side-info sub-streams
 * (dsv_encoder.c:256-327, dsv.c:189-231) and packet framing (dsv_encoder.c:170-192,410-536).
 * Every per-pixel / per-coefficient / per-bit stage is a kernel launch on the encoder's own stream:
 *   H2D -> extend/pyramid -> HME -> [sync: MVs] -> BMC -> SBT fwd + quant -> HZCC -> SBT inv -> recon.
 * There is no CPU implementation of those stages in this library.
 */
#include "dsv1_b200.h"

#include "../frame.cuh"
#include "../hzcc.cuh"
#include "../motion.cuh"
#include "../sbt.cuh"
#include "bits.h"
#include "encoder_ctx.h"

using namespace dsv;

namespace dsv {

static int size4dim(int dim) /* dsv_encoder.c:556-572 */
{
    if (dim > 1280) return 64;
    if (dim > 1024) return 48;
    if (dim > 704) return 32;
    if (dim > 352) return 24;
    return 16;
}

void plan_geometry(CodecGeom *g, int w, int h, int subsamp)
{
    g->w = w;
    g->h = h;
    g->subsamp = subsamp;
    g->hs = DSV_FORMAT_H_SHIFT(subsamp);
    g->vs = DSV_FORMAT_V_SHIFT(subsamp);
    g->pw[0] = w;
    g->ph[0] = h;
    g->pw[1] = g->pw[2] = ceil_shift(w, g->hs);
    g->ph[1] = g->ph[2] = ceil_shift(h, g->vs);
    g->cw[0] = w;
    g->ch[0] = h;
    g->cw[1] = g->cw[2] = (g->pw[1] + 1) & ~1; /* dsv_mk_coefs: chroma rounded up to even */
    g->ch[1] = g->ch[2] = (g->ph[1] + 1) & ~1;
    g->coef_off[0] = 0;
    g->coef_off[1] = (size_t) g->cw[0] * g->ch[0];
    g->coef_off[2] = g->coef_off[1] + (size_t) g->cw[1] * g->ch[1];
    g->coef_total = g->coef_off[2] + (size_t) g->cw[2] * g->ch[2];
    g->frame_bytes = (size_t) w * h + 2 * (size_t) g->pw[1] * g->ph[1];
}

void plan_blocks(CodecGeom *g, int blk_w, int blk_h)
{
    g->blk_w = blk_w;
    g->blk_h = blk_h;
    g->nbh = ceil_div(g->w, blk_w);
    g->nbv = ceil_div(g->h, blk_h);
    g->nblk = g->nbh * g->nbv;
}

/* device buffers shared by the encoder and decoder pipelines: coefficient planes + job tables */
void coder_alloc(CoderBufs *c, const CodecGeom &g)
{
    CUDA_CHECK(cudaMalloc(&c->coef, g.coef_total * sizeof(int32_t)));
    for (int p = 0; p < 3; p++) {
        CUDA_CHECK(cudaMalloc(&c->llx[p], sbt_llx_elems(g.cw[p], g.ch[p]) * sizeof(int32_t)));
        CUDA_CHECK(cudaMalloc(&c->dv[p], sbt_dv_elems(g.cw[p], g.ch[p]) * sizeof(int32_t)));
        CUDA_CHECK(cudaMemset(c->dv[p], 0, sbt_dv_elems(g.cw[p], g.ch[p]) * sizeof(int32_t)));
    }
    CUDA_CHECK(cudaMalloc(&c->d_sjobs, 3 * sizeof(SbtJob)));
    CUDA_CHECK(cudaMalloc(&c->d_hjobs, 3 * sizeof(HzJob)));
    CUDA_CHECK(cudaMalloc(&c->d_frame, sizeof(HzFrame)));
    CUDA_CHECK(cudaMalloc(&c->d_stab, (size_t) imax(g.nblk, 1)));
    int chunks = 0;
    for (int p = 0; p < 3; p++) {
        HzRegions r;
        hz_fill_regions(&r, g.cw[p], g.ch[p]);
        chunks += ceil_div(r.base[HZ_NREG], HZ_CHUNK);
    }
    c->total_chunks = chunks;
    CUDA_CHECK(cudaMalloc(&c->d_chunks, (size_t) chunks * sizeof(HzChunk)));
    c->lo_smem = 0;
    for (int p = 0; p < 3; p++) {
        size_t s = sbt_lo_smem_bytes(g.cw[p], g.ch[p]);
        c->lo_smem = s > c->lo_smem ? s : c->lo_smem;
    }
}

void coder_free(CoderBufs *c)
{
    cudaFree(c->coef);
    for (int p = 0; p < 3; p++) {
        cudaFree(c->llx[p]);
        cudaFree(c->dv[p]);
    }
    cudaFree(c->d_sjobs);
    cudaFree(c->d_hjobs);
    cudaFree(c->d_frame);
    cudaFree(c->d_stab);
    cudaFree(c->d_chunks);
    memset(c, 0, sizeof(*c));
}

/* fill + upload the three plane jobs of one picture */
void coder_setup_jobs(CoderBufs *c, const CodecGeom &g, const DevFrame &pix, int quant, int isP, int do_quant,
                      cudaStream_t st)
{
    int tile_base = 0, chunk_base = 0;
    for (int p = 0; p < 3; p++) {
        SbtJob &s = c->sj[p];
        memset(&s, 0, sizeof(s));
        sbt_fill_geometry(&s, g.pw[p], g.ph[p], g.cw[p], g.ch[p], isP, p);
        sbt_fill_quant(&s, quant, isP, p, g.nbh, g.nbv);
        s.pix = pix.p[p];
        s.pstride = pix.stride[p];
        s.coef = c->coef + g.coef_off[p];
        s.llx = c->llx[p];
        s.dv = c->dv[p];
        s.stable = c->d_stab;
        s.do_quant = do_quant;
        s.tile_base = tile_base;
        tile_base += s.tiles_x * s.tiles_y;

        HzJob &h = c->hj[p];
        memset(&h, 0, sizeof(h));
        hz_fill_job(&h, g.cw[p], g.ch[p], quant, isP, p, g.nbh, g.nbv);
        h.coef = s.coef;
        h.dv = s.dv;
        h.stable = c->d_stab;
        h.chunk_base = chunk_base;
        h.frame = 0;
        chunk_base += h.nchunks;
    }
    c->total_tiles = tile_base;
    CUDA_CHECK(cudaMemcpyAsync(c->d_sjobs, c->sj, sizeof(c->sj), cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(c->d_hjobs, c->hj, sizeof(c->hj), cudaMemcpyHostToDevice, st));
}

} // namespace dsv

/* ============================================================================================== */

static EncCtx *enc_ctx(DSV_ENCODER *enc) { return reinterpret_cast<EncCtx *>(enc->ref); }

static void enc_ctx_destroy(EncCtx *c)
{
    if (!c) {
        return;
    }
    cudaStreamSynchronize(c->st);
    coder_free(&c->cb);
    devframe_free(&c->xf);
    devframe_free(&c->pred);
    for (int i = 0; i < 2; i++) {
        devframe_free(&c->pad[i]);
        devframe_free(&c->recon[i]);
        for (int l = 0; l < DSV_MAX_PYRAMID_LEVELS; l++) {
            devframe_free(&c->pyr[i][l]);
        }
    }
    for (int l = 0; l <= DSV_MAX_PYRAMID_LEVELS; l++) {
        cudaFree(c->d_mvf[l]);
    }
    cudaFree(c->d_aux);
    cudaFree(c->d_pkt);
    cudaFree(c->d_misc);
    cudaFreeHost(c->h_in);
    cudaFreeHost(c->h_pkt);
    cudaFreeHost(c->h_mv);
    cudaFreeHost(c->h_misc);
    cudaFreeHost(c->h_frame);
    cudaStreamDestroy(c->st);
    delete c;
}

/* created on the first dsv_enc call, when metadata, gop and pyramid settings are final */
static EncCtx *enc_ctx_create(DSV_ENCODER *enc)
{
    EncCtx *c = new EncCtx();
    const DSV_META &md = enc->vidmeta;
    if ((md.width & 1) || (md.height & 1) || md.width < 16 || md.height < 16) {
        DSV_ERROR(("unsupported dimensions %dx%d: width and height must be even and >= 16", md.width, md.height));
        exit(-1);
    }
    plan_geometry(&c->g, md.width, md.height, md.subsamp);
    plan_blocks(&c->g, iclamp(size4dim(md.width) & ~7, DSV_MIN_BLOCK_SIZE, DSV_MAX_BLOCK_SIZE),
                iclamp(size4dim(md.height) & ~7, DSV_MIN_BLOCK_SIZE, DSV_MAX_BLOCK_SIZE));
    const CodecGeom &g = c->g;
    c->inter = enc->gop != DSV_GOP_INTRA;

    CUDA_CHECK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    coder_alloc(&c->cb, g);
    devframe_alloc(&c->xf, g.w, g.h, g.subsamp);

    /* packet upper bound as in dsv_encoder.c:472-491 */
    size_t ub = (size_t) g.w * g.h;
    ub *= (g.subsamp == DSV_SUBSAMP_444) ? 6 : (g.subsamp == DSV_SUBSAMP_422) ? 4 : 2;
    c->pkt_cap = ub + 4096 + (size_t) g.nblk * 48;
    CUDA_CHECK(cudaMalloc(&c->d_pkt, c->pkt_cap));
    CUDA_CHECK(cudaMemset(c->d_pkt, 0, c->pkt_cap));
    c->pkt_dirty = 0;
    CUDA_CHECK(cudaMalloc(&c->d_misc, 64));
    CUDA_CHECK(cudaMallocHost(&c->h_in, g.frame_bytes));
    CUDA_CHECK(cudaMallocHost(&c->h_pkt, c->pkt_cap));
    CUDA_CHECK(cudaMallocHost(&c->h_mv, sizeof(DevMV) * (size_t) g.nblk));
    CUDA_CHECK(cudaMallocHost(&c->h_misc, 64));
    CUDA_CHECK(cudaMallocHost(&c->h_frame, sizeof(HzFrame)));

    if (c->inter) {
        devframe_alloc(&c->pred, g.w, g.h, g.subsamp);
        for (int i = 0; i < 2; i++) {
            devframe_alloc(&c->pad[i], g.w, g.h, g.subsamp);
            devframe_alloc(&c->recon[i], g.w, g.h, g.subsamp);
            for (int l = 0; l < enc->pyramid_levels; l++) {
                devframe_alloc(&c->pyr[i][l], ceil_shift(g.w, l + 1), ceil_shift(g.h, l + 1), g.subsamp);
            }
        }
        for (int l = 0; l <= enc->pyramid_levels; l++) {
            CUDA_CHECK(cudaMalloc(&c->d_mvf[l], sizeof(DevMV) * (size_t) g.nblk));
            CUDA_CHECK(cudaMemset(c->d_mvf[l], 0, sizeof(DevMV) * (size_t) g.nblk));
        }
        CUDA_CHECK(cudaMalloc(&c->d_aux, sizeof(int2) * (size_t) g.nblk));
    }
    return c;
}

/* ---- packet pieces -------------------------------------------------------------------------- */

static void put_packet_hdr(BitWriter &bw, int pkt_type) /* dsv_encoder.c:410-424 */
{
    bw.put_bits(8, DSV_FOURCC_0);
    bw.put_bits(8, DSV_FOURCC_1);
    bw.put_bits(8, DSV_FOURCC_2);
    bw.put_bits(8, DSV_FOURCC_3);
    bw.put_bits(8, DSV_VERSION_MINOR);
    bw.put_bits(8, (uint32_t) pkt_type);
    bw.put_bits(32, 0);
    bw.put_bits(32, 0);
}

static void put_be32(uint8_t *p, unsigned v)
{
    p[0] = (uint8_t) (v >> 24);
    p[1] = (uint8_t) (v >> 16);
    p[2] = (uint8_t) (v >> 8);
    p[3] = (uint8_t) v;
}

static void set_links(DSV_ENCODER *enc, DSV_BUF *buf, int is_eos) /* dsv_encoder.c:170-192 */
{
    unsigned next = is_eos ? 0 : buf->len;
    put_be32(buf->data + DSV_PACKET_PREV_OFFSET, (unsigned) enc->prev_link);
    put_be32(buf->data + DSV_PACKET_NEXT_OFFSET, next);
    enc->prev_link = (int) next;
}

static void make_metadata_packet(DSV_ENCODER *enc, DSV_BUF *buf) /* dsv_encoder.c:426-461 */
{
    const DSV_META &m = enc->vidmeta;
    dsv_mk_buf(buf, 64);
    BitWriter bw(buf->data);
    put_packet_hdr(bw, DSV_PT_META);
    bw.put_ueg((uint32_t) m.width);
    bw.put_ueg((uint32_t) m.height);
    bw.put_ueg((uint32_t) m.subsamp);
    bw.put_ueg((uint32_t) m.fps_num);
    bw.put_ueg((uint32_t) m.fps_den);
    bw.put_ueg((uint32_t) m.aspect_num);
    bw.put_ueg((uint32_t) m.aspect_den);
    bw.align();
    buf->len = bw.byte_pos();
    put_be32(buf->data + DSV_PACKET_NEXT_OFFSET, buf->len);
}

/* stability tracker + its ZBRLE map (dsv_encoder.c:329-408); also fills enc->stable_blocks */
static void put_stable_blocks(DSV_ENCODER *enc, EncCtx *c, int isP, const DevMV *mvs, BitWriter &bw)
{
    const int nblk = c->g.nblk;
    std::vector<uint8_t> tmp((size_t) nblk * 4 + 64, 0);
    RleWriter rle(tmp.data());
    if (enc->refresh_ctr >= enc->stable_refresh) {
        enc->refresh_ctr = 0;
        memset(enc->stability, 0, sizeof(*enc->stability) * (size_t) nblk);
    }
    int avgdiv = (int) enc->refresh_ctr;
    if (avgdiv <= 0) {
        avgdiv = 1;
    }
    for (int i = 0; i < nblk; i++) {
        int stable = 0, intra = 0;
        if (isP) {
            const DevMV &mv = mvs[i];
            if (mv.mode == DSV_MODE_INTER) {
                enc->stability[i].x += iabs(mv.x) >> 2;
                enc->stability[i].y += iabs(mv.y) >> 2;
                stable = mv.high_detail;
                int ax = enc->stability[i].x / avgdiv, ay = enc->stability[i].y / avgdiv;
                stable |= (ax == 0 && ay == 0 && !mv.lo_tex && !mv.lo_var);
            } else {
                intra = 1;
            }
            if (mv.lo_tex || mv.lo_var) { /* not worth bits: poison the accumulators */
                enc->stability[i].x = 0x3fff;
                enc->stability[i].y = 0x3fff;
            }
        } else {
            int ax = enc->stability[i].x / avgdiv, ay = enc->stability[i].y / avgdiv;
            stable = (ax == 0 && ay == 0);
        }
        enc->stable_blocks[i] = (unsigned char) (stable | (intra << 1));
        rle.put(stable & 1);
    }
    bw.align();
    unsigned bytes = rle.finish();
    bw.put_ueg(bytes);
    bw.align();
    bw.concat(tmp.data(), bytes);
}

static int mv_pred1(int left, int top, int topleft) /* dsv.c:189-197 */
{
    int dif = left + top - topleft;
    return iabs(dif - left) < iabs(dif - top) ? left : top;
}

void dsv::predict_mv(const DevMV *mvs, int nbh, int x, int y, int *px, int *py) /* dsv.c:199-231 */
{
    int vx[3] = {0, 0, 0}, vy[3] = {0, 0, 0};
    if (x > 0) {
        const DevMV &m = mvs[y * nbh + x - 1];
        if (m.mode == DSV_MODE_INTER) { vx[0] = m.x; vy[0] = m.y; }
    }
    if (y > 0) {
        const DevMV &m = mvs[(y - 1) * nbh + x];
        if (m.mode == DSV_MODE_INTER) { vx[1] = m.x; vy[1] = m.y; }
    }
    if (x > 0 && y > 0) {
        const DevMV &m = mvs[(y - 1) * nbh + x - 1];
        if (m.mode == DSV_MODE_INTER) { vx[2] = m.x; vy[2] = m.y; }
    }
    *px = mv_pred1(vx[0], vx[1], vx[2]);
    *py = mv_pred1(vy[0], vy[1], vy[2]);
}

/* four byte-aligned sub-streams: modes (ZBRLE), MV x, MV y (SEG of prediction error), intra masks
 * (dsv_encoder.c:256-327) */
static void put_motion(EncCtx *c, const DevMV *mvs, BitWriter &bw)
{
    const int nbh = c->g.nbh, nbv = c->g.nbv;
    const size_t ub = (size_t) nbh * nbv * 32;
    std::vector<uint8_t> b_mode(ub, 0), b_x(ub, 0), b_y(ub, 0), b_mask(ub, 0);
    RleWriter rle(b_mode.data());
    BitWriter wx(b_x.data()), wy(b_y.data()), wm(b_mask.data());
    for (int j = 0; j < nbv; j++) {
        for (int i = 0; i < nbh; i++) {
            const DevMV &mv = mvs[j * nbh + i];
            rle.put(mv.mode);
            if (mv.mode == DSV_MODE_INTER) {
                int px, py;
                predict_mv(mvs, nbh, i, j, &px, &py);
                wx.put_seg(mv.x - px);
                wy.put_seg(mv.y - py);
            } else if (mv.submask == DSV_MASK_ALL_INTRA) {
                wm.put_bit(1);
            } else {
                wm.put_bit(0);
                wm.put_bits(4, mv.submask);
            }
        }
    }
    unsigned n_mode = rle.finish();
    wx.align();
    wy.align();
    wm.align();
    const uint8_t *data[4] = {b_mode.data(), b_x.data(), b_y.data(), b_mask.data()};
    unsigned len[4] = {n_mode, wx.byte_pos(), wy.byte_pos(), wm.byte_pos()};
    for (int s = 0; s < 4; s++) {
        bw.align();
        bw.put_ueg(len[s]);
        bw.align();
        bw.concat(data[s], len[s]);
    }
}

/* ---- public API ------------------------------------------------------------------------------ */

extern "C" void dsv_enc_init(DSV_ENCODER *enc) /* defaults: dsv_encoder.c:696-722 */
{
    memset(enc, 0, sizeof(*enc));
    enc->prev_gop = (DSV_FNUM) -1;
    enc->quality = DSV_QUALITY_PERCENT(85);
    enc->gop = 24;
    enc->pyramid_levels = 0;
    enc->rc_mode = DSV_RATE_CONTROL_CRF;
    enc->bitrate = INT_MAX;
    enc->max_q_step = DSV_MAX_QUALITY / 200;
    enc->min_quality = DSV_QUALITY_PERCENT(1);
    enc->max_quality = DSV_QUALITY_PERCENT(95);
    enc->min_I_frame_quality = DSV_QUALITY_PERCENT(5);
    enc->rc_high_motion_nudge = 1;
    enc->intra_pct_thresh = 50;
    enc->stable_refresh = 14;
    enc->scene_change_delta = 4;
    enc->do_scd = 1;
}

extern "C" void dsv_enc_start(DSV_ENCODER *enc) /* dsv_encoder.c:724-734 */
{
    enc->quality = iclamp(enc->quality, 0, DSV_MAX_QUALITY);
    if (enc->rc_mode != DSV_RATE_CONTROL_CRF) {
        enc->rc_quant = (unsigned) enc->quality;
        enc->avg_P_frame_q = enc->quality * 4 / 5;
    }
    enc->force_metadata = 1;
}

extern "C" void dsv_enc_free(DSV_ENCODER *enc)
{
    enc_ctx_destroy(enc_ctx(enc));
    enc->ref = NULL;
    if (enc->stability) {
        dsv_free(enc->stability);
        enc->stability = NULL;
    }
    if (enc->stable_blocks) {
        dsv_free(enc->stable_blocks);
        enc->stable_blocks = NULL;
    }
}

extern "C" void dsv_enc_set_metadata(DSV_ENCODER *enc, DSV_META *md) { enc->vidmeta = *md; }
extern "C" void dsv_enc_force_metadata(DSV_ENCODER *enc) { enc->force_metadata = 1; }

extern "C" void dsv_enc_end_of_stream(DSV_ENCODER *enc, DSV_BUF *bufs) /* dsv_encoder.c:765-778 */
{
    dsv_mk_buf(&bufs[0], DSV_PACKET_HDR_SIZE);
    BitWriter bw(bufs[0].data);
    put_packet_hdr(bw, DSV_PT_EOS);
    set_links(enc, &bufs[0], 1);
}

/* ABR controller (dsv_encoder.c:84-160): host-only, needs every previous packet's size, hence 1 GPU */
static int rate_control_quality(DSV_ENCODER *enc, int isP, int forced_intra)
{
    int q = (int) enc->rc_quant;
    if (enc->rc_mode == DSV_RATE_CONTROL_CRF) {
        q = enc->quality;
        enc->rc_quant = (unsigned) q;
        return q;
    }
    const DSV_META &vf = enc->vidmeta;
    int fps = (vf.fps_num << 5) / vf.fps_den;
    if (fps == 0) {
        fps = 1;
    }
    int needed_bpf = (int) (((enc->bitrate << 5) / (unsigned) fps) >> 3);
    int bpf = enc->bpf_avg ? enc->bpf_avg : needed_bpf;
    int dir = (bpf - needed_bpf) > 0 ? -1 : 1;
    int delta = (iabs(bpf - needed_bpf) << 9) / needed_bpf;
    int nudged = 0;
    if (dir == 1) {
        delta *= 2;
    }
    if (enc->rc_high_motion_nudge) {
        if (isP && enc->last_P_frame_over) {
            delta = (delta + 1) * 2;
            dir = -1;
            nudged = 1;
        } else if (enc->back_into_range) {
            delta = (delta + 1) * 2;
            dir = 1;
            nudged = 1;
        }
    }
    delta = (q * delta) >> 9;
    enc->max_q_step = iclamp(enc->max_q_step, 1, DSV_MAX_QUALITY);
    delta = imin(delta, nudged ? enc->max_q_step * 16 : enc->max_q_step);
    q += delta * dir;
    int low_p = iclamp(enc->avg_P_frame_q - DSV_QUALITY_PERCENT(4), enc->min_quality, enc->max_quality);
    int minq = isP ? low_p : enc->min_I_frame_quality;
    if (forced_intra) {
        if (q < DSV_QUALITY_PERCENT(60)) {
            q += DSV_QUALITY_PERCENT(15);
        } else if (q < DSV_QUALITY_PERCENT(70)) {
            q += DSV_QUALITY_PERCENT(8);
        } else if (q < DSV_QUALITY_PERCENT(75)) {
            q += DSV_QUALITY_PERCENT(3);
        }
        q = iclamp(q, 0, enc->max_quality - DSV_QUALITY_PERCENT(5));
    }
    q = iclamp(q, minq, enc->max_quality);
    q = iclamp(q, 0, DSV_MAX_QUALITY);
    enc->rc_quant = (unsigned) q;
    return q;
}

static void rate_control_update(DSV_ENCODER *enc, int isP, unsigned pkt_len) /* dsv_encoder.c:816-848 */
{
    if (enc->rc_mode == DSV_RATE_CONTROL_CRF) {
        return;
    }
    enc->bpf_total += pkt_len;
    enc->bpf_reset++;
    if (isP) {
        enc->total_P_frame_q += (int) enc->rc_quant;
        enc->avg_P_frame_q = enc->total_P_frame_q / (int) enc->bpf_reset;
        unsigned fps = (unsigned) ((enc->vidmeta.fps_num << 5) / enc->vidmeta.fps_den);
        if (fps == 0) {
            fps = 1;
        }
        unsigned needed = ((enc->bitrate << 5) / fps) >> 3;
        int under = pkt_len < (needed * 3 / 4);
        int over = pkt_len > (needed * 7 / 8);
        enc->back_into_range = (enc->last_P_frame_over && under);
        enc->last_P_frame_over = over;
    } else {
        enc->last_P_frame_over = 0;
        enc->back_into_range = 0;
    }
    enc->bpf_avg = (int) (enc->bpf_total / enc->bpf_reset);
    if (enc->bpf_reset >= DSV_BPF_RESET) {
        enc->bpf_total = (unsigned) enc->bpf_avg;
        enc->total_P_frame_q = enc->total_P_frame_q / (int) enc->bpf_reset;
        enc->bpf_reset = 1;
    }
}

extern "C" int dsv_enc(DSV_ENCODER *enc, DSV_FRAME *frame, DSV_BUF *bufs)
{
    if (bufs == NULL) {
        DSV_ERROR(("null buffer list passed to encoder!"));
        return 0;
    }
    const int w = enc->vidmeta.width, h = enc->vidmeta.height;

    /* block size / pyramid depth (dsv_encoder.c:588-613) -- fixed for the life of the encoder */
    const int blk_w = iclamp(size4dim(w) & ~7, DSV_MIN_BLOCK_SIZE, DSV_MAX_BLOCK_SIZE);
    const int blk_h = iclamp(size4dim(h) & ~7, DSV_MIN_BLOCK_SIZE, DSV_MAX_BLOCK_SIZE);
    const int nbh = ceil_div(w, blk_w), nbv = ceil_div(h, blk_h);
    if (enc->stability == NULL) {
        enc->stability = (decltype(enc->stability)) dsv_alloc((int) sizeof(*enc->stability) * nbh * nbv);
        enc->stable_blocks = (unsigned char *) dsv_alloc(nbh * nbv);
    }
    if (enc->pyramid_levels == 0) {
        int lvls = lb2((unsigned) imin(w, h));
        int maxdim = imax(nbh, nbv);
        while ((1 << lvls) > maxdim) {
            lvls--;
        }
        enc->pyramid_levels = iclamp(lvls, 3, DSV_MAX_PYRAMID_LEVELS);
    }
    if (enc->ref == NULL) {
        enc->ref = reinterpret_cast<DSV_ENCDATA *>(enc_ctx_create(enc));
    }
    EncCtx *c = enc_ctx(enc);
    const CodecGeom &g = c->g;
    cudaStream_t st = c->st;
    const DSV_FNUM fnum = enc->next_fnum++;
    const int cur = c->cur, prev = cur ^ 1;

    /* ---- input: caller memory -> pinned staging -> device (valid only during this call) ---- */
    {
        uint8_t *o = c->h_in;
        for (int p = 0; p < 3; p++) {
            const DSV_PLANE *pl = &frame->planes[p];
            for (int y = 0; y < g.ph[p]; y++) {
                memcpy(o, pl->data + (size_t) y * pl->stride, (size_t) g.pw[p]);
                o += g.pw[p];
            }
        }
        const DevFrame &dst = c->inter ? c->pad[cur] : c->xf;
        const uint8_t *s = c->h_in;
        for (int p = 0; p < 3; p++) {
            CUDA_CHECK(cudaMemcpy2DAsync(dst.p[p], dst.stride[p], s, g.pw[p], g.pw[p], g.ph[p], cudaMemcpyHostToDevice, st));
            s += (size_t) g.pw[p] * g.ph[p];
        }
        frame_extend_launch(dst, 3, st); /* clone + extend (dsv_encoder.c:617-623, frame.c:218-220) */
    }
    dsv_frame_ref_dec(frame); /* the encoder owns the reference it was given (dsv_encoder.c:38-40,801) */

    /* ---- GOP bookkeeping (dsv_encoder.c:624-652) ---- */
    int gop_start = 0, is_ref = 0, has_ref = 0, forced_intra = 0;
    if (enc->force_metadata || ((enc->prev_gop + (DSV_FNUM) enc->gop) <= fnum)) {
        gop_start = 1;
        enc->prev_gop = fnum;
        enc->force_metadata = 0;
    }
    if (c->inter) {
        is_ref = 1;
        has_ref = !gop_start && c->have_ref;
        if (!gop_start && !c->have_ref) {
            DSV_ASSERT(0 && "P frame without a reference");
        }
        /* pyramid of the ORIGINAL frame: used by this frame's search and by the next frame as its reference */
        const DevFrame *below = &c->pad[cur];
        for (int l = 0; l < enc->pyramid_levels; l++) {
            frame_down2_luma_launch(*below, c->pyr[cur][l], st);
            below = &c->pyr[cur][l];
        }
        unsigned long long *d_sum = reinterpret_cast<unsigned long long *>(c->d_misc);
        int *d_nintra = reinterpret_cast<int *>(c->d_misc + 8);
        int need_sync = 0;
        if (enc->do_scd) {
            frame_sum_luma_launch(c->pyr[cur][enc->pyramid_levels - 1], d_sum, st);
            need_sync = 1;
        }
        if (has_ref) { /* speculative: a scene change below simply discards the vectors */
            MotionGeom mg = {g.w, g.h, g.hs, g.vs, g.blk_w, g.blk_h, g.nbh, g.nbv, enc->pyramid_levels};
            DevFrame src[DSV_MAX_PYRAMID_LEVELS + 1], ref[DSV_MAX_PYRAMID_LEVELS + 1];
            src[0] = c->pad[cur];
            ref[0] = c->pad[prev];
            for (int l = 0; l < enc->pyramid_levels; l++) {
                src[l + 1] = c->pyr[cur][l];
                ref[l + 1] = c->pyr[prev][l];
            }
            hme_launch(mg, src, ref, c->d_mvf, c->d_aux, d_nintra, st);
            CUDA_CHECK(cudaMemcpyAsync(c->h_mv, c->d_mvf[0], sizeof(DevMV) * (size_t) g.nblk, cudaMemcpyDeviceToHost, st));
            need_sync = 1;
        }
        if (need_sync) {
            CUDA_CHECK(cudaMemcpyAsync(c->h_misc, c->d_misc, 16, cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(cudaStreamSynchronize(st));
        }
        if (enc->do_scd) { /* check_scene_change, dsv_encoder.c:538-554 */
            const DevFrame &top = c->pyr[cur][enc->pyramid_levels - 1];
            int al = (int) (*reinterpret_cast<unsigned long long *>(c->h_misc) / (unsigned long long) (top.w[0] * top.h[0]));
            if (iabs(enc->prev_avg_luma - al) > enc->scene_change_delta) {
                has_ref = 0;
                forced_intra = 1;
            }
            enc->prev_avg_luma = al;
        }
        if (has_ref) { /* motion_est's verdict, dsv_encoder.c:246-253 */
            int nintra = *reinterpret_cast<int *>(c->h_misc + 8);
            int pct = nintra * 100 / g.nblk;
            forced_intra = 0;
            if (pct > enc->intra_pct_thresh) {
                has_ref = 0;
                forced_intra = 1;
            }
        }
    }
    const int isP = has_ref;
    const int quality = rate_control_quality(enc, isP, forced_intra);
    const int quant = DSV_MAX_QUALITY - ((DSV_MAX_QUALITY - 5) * quality / DSV_MAX_QUALITY); /* dsv_encoder.c:165 */

    /* ---- residual formation (dsv_encoder.c:657-660) ---- */
    if (c->inter) {
        CUDA_CHECK(cudaMemcpyAsync(c->xf.alloc, c->pad[cur].alloc, c->xf.bytes, cudaMemcpyDeviceToDevice, st));
        if (has_ref) {
            MotionGeom mg = {g.w, g.h, g.hs, g.vs, g.blk_w, g.blk_h, g.nbh, g.nbv, enc->pyramid_levels};
            bmc_launch(mg, c->d_mvf[0], c->recon[prev], &c->pred, c->xf, 1, st);
        }
    }

    /* ---- packet head on the host: header, frame number, block size, stability, motion ---- */
    memset(c->h_pkt, 0, 256 + (size_t) g.nblk * 48);
    BitWriter bw(c->h_pkt);
    put_packet_hdr(bw, DSV_MAKE_PT(is_ref, has_ref));
    bw.align();
    bw.put_bits(32, fnum);
    bw.align();
    bw.put_ueg((uint32_t) (g.blk_w >> 2));
    bw.put_ueg((uint32_t) (g.blk_h >> 2));
    bw.align();
    put_stable_blocks(enc, c, isP, c->h_mv, bw);
    if (has_ref) {
        bw.align();
        put_motion(c, c->h_mv, bw);
    }
    bw.align();
    bw.put_bits(DSV_MAX_QP_BITS, (uint32_t) quant);
    bw.align(); /* dsv_encode_plane aligns before each plane (hzcc.c:457) */
    const unsigned head_bytes = bw.byte_pos();

    /* ---- device side of encode_picture (dsv_encoder.c:513-526) ---- */
    if (c->pkt_dirty) {
        CUDA_CHECK(cudaMemsetAsync(c->d_pkt, 0, imin((int) c->pkt_cap, (int) c->pkt_dirty + 64), st));
    }
    CUDA_CHECK(cudaMemcpyAsync(c->d_pkt, c->h_pkt, head_bytes, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(c->cb.d_stab, enc->stable_blocks, (size_t) g.nblk, cudaMemcpyHostToDevice, st));
    coder_setup_jobs(&c->cb, g, c->xf, quant, isP, 1, st);
    HzFrame hf;
    memset(&hf, 0, sizeof(hf));
    hf.pkt = c->d_pkt;
    hf.start_byte = head_bytes;
    hf.nplanes = 3;
    hf.job[0] = 0; hf.job[1] = 1; hf.job[2] = 2;
    *c->h_frame = hf;
    CUDA_CHECK(cudaMemcpyAsync(c->cb.d_frame, c->h_frame, sizeof(HzFrame), cudaMemcpyHostToDevice, st));
    sbt_fwd_launch(c->cb.d_sjobs, 3, c->cb.total_tiles, c->cb.lo_smem, st);
    hzcc_enc_launch(c->cb.d_hjobs, 3, c->cb.d_chunks, c->cb.total_chunks, c->cb.d_frame, 1, st);
    CUDA_CHECK(cudaMemcpyAsync(c->h_frame, c->cb.d_frame, sizeof(HzFrame), cudaMemcpyDeviceToHost, st));
    if (is_ref) {
        /* closed loop: reconstruct exactly what the decoder will (dsv_encoder.c:525,662-674) */
        sbt_inv_launch(c->cb.d_sjobs, 3, c->cb.total_tiles, c->cb.lo_smem, !isP, st);
        if (has_ref) {
            frame_add_launch(c->xf, c->pred, st);
        }
        frame_copy_launch(c->recon[cur], c->xf, st);
        frame_extend_launch(c->recon[cur], 3, st);
    }
    CUDA_CHECK(cudaStreamSynchronize(st));
    const unsigned total = c->h_frame->total_bytes;
    if (total > c->pkt_cap - 64) {
        DSV_ERROR(("packet exceeds the output bound"));
        exit(-1);
    }
    DSV_BUF outbuf;
    dsv_mk_buf(&outbuf, (int) total + 8);
    outbuf.len = total;
    memcpy(outbuf.data, c->h_pkt, head_bytes);
    CUDA_CHECK(cudaMemcpyAsync(outbuf.data + head_bytes, c->d_pkt + head_bytes, total - head_bytes, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    c->pkt_dirty = total;
    if (is_ref) {
        c->have_ref = 1;
        c->cur ^= 1;
    }

    int nbuf = 0;
    if (gop_start) {
        make_metadata_packet(enc, &bufs[nbuf++]);
    }
    bufs[nbuf++] = outbuf;
    if (isP) {
        enc->refresh_ctr++;
    }
    rate_control_update(enc, isP, outbuf.len);
    set_links(enc, &bufs[nbuf - 1], 0);
    return nbuf;
}
