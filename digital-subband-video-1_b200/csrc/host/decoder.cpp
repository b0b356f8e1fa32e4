/*
 * decoder.cpp -- the dsv_decoder.h API (dsv_decoder.c:22-145,244-472) on top of the CUDA kernels.
 *
 * Host code parses only what is serial and tiny: packet header, metadata, the ZBRLE stability map and
 * the four motion sub-streams (<= 2040 blocks), plus SEG(DC) / nruns / first run of each plane head.
 * The packet is copied to the device once; coefficient parsing (parallel bit-FSM), dequantisation,
 * the inverse subband transform, motion compensation + reconstruction and the border extension of the
 * new reference all run as kernels on the decoder's own stream:
 *   H2D packet -> hzcc parse x3 -> SBT inverse x3 -> BMC + add (P) -> extend (refs) -> D2H frame.
 * There is no CPU implementation of those stages in this library.
 */
#include "dsv1_b200.h"

#include "../frame.cuh"
#include "../hzcc.cuh"
#include "../hzcc_dec.cuh"
#include "../motion.cuh"
#include "../sbt.cuh"
#include "bits.h"
#include "encoder_ctx.h"

using namespace dsv;

namespace dsv {

/* device context behind DSV_DECODER.ref (the reference keeps its DSV_IMAGE there, dsv_decoder.h:26-36) */
struct DecCtx {
    CodecGeom g;
    CoderBufs cb;
    HzDecBufs hz;
    cudaStream_t st = 0;
    DevFrame out[2]; /* out[cur] is being decoded, out[cur ^ 1] is the reference picture */
    int cur = 0;
    int have_ref = 0;
    uint8_t *d_pkt = nullptr, *h_pkt = nullptr;
    size_t pkt_cap = 0;
    DevMV *d_mv = nullptr, *h_mv = nullptr;
    uint8_t *h_stab = nullptr;
    uint8_t *h_out = nullptr;
    int max_nblk = 0;
};

} // namespace dsv

static DecCtx *dec_ctx(DSV_DECODER *d) { return reinterpret_cast<DecCtx *>(d->ref); }

static void dec_ctx_destroy(DecCtx *c)
{
    if (!c) {
        return;
    }
    cudaStreamSynchronize(c->st);
    coder_free(&c->cb);
    hzdec_free(&c->hz);
    devframe_free(&c->out[0]);
    devframe_free(&c->out[1]);
    cudaFree(c->d_pkt);
    cudaFree(c->d_mv);
    cudaFreeHost(c->h_pkt);
    cudaFreeHost(c->h_mv);
    cudaFreeHost(c->h_stab);
    cudaFreeHost(c->h_out);
    cudaStreamDestroy(c->st);
    delete c;
}

static DecCtx *dec_ctx_create(const DSV_META &md)
{
    DecCtx *c = new DecCtx();
    plan_geometry(&c->g, md.width, md.height, md.subsamp);
    plan_blocks(&c->g, DSV_MIN_BLOCK_SIZE, DSV_MIN_BLOCK_SIZE); /* worst case; the real size is per picture */
    const CodecGeom &g = c->g;
    c->max_nblk = g.nblk;
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    coder_alloc(&c->cb, g);
    HzDecPlan pl[3];
    for (int p = 0; p < 3; p++) {
        hzdec_plan(&pl[p], g.cw[p], g.ch[p]);
    }
    hzdec_alloc(&c->hz, pl);
    devframe_alloc(&c->out[0], g.w, g.h, g.subsamp);
    devframe_alloc(&c->out[1], g.w, g.h, g.subsamp);
    /* a picture packet holds three planes of at most 2 * 4 * cw * ch bytes each (dsv_decoder.c:397-401) */
    c->pkt_cap = g.coef_total * 8 + 4096 + (size_t) g.nblk * 64;
    CUDA_CHECK(cudaMalloc(&c->d_pkt, c->pkt_cap + 64));
    CUDA_CHECK(cudaMallocHost(&c->h_pkt, c->pkt_cap + 64));
    CUDA_CHECK(cudaMalloc(&c->d_mv, sizeof(DevMV) * (size_t) g.nblk));
    CUDA_CHECK(cudaMallocHost(&c->h_mv, sizeof(DevMV) * (size_t) g.nblk));
    CUDA_CHECK(cudaMallocHost(&c->h_stab, (size_t) g.nblk));
    CUDA_CHECK(cudaMallocHost(&c->h_out, g.frame_bytes));
    return c;
}

static bool meta_supported(const DSV_META &m)
{
    if (m.width < 16 || m.height < 16 || (m.width & 1) || (m.height & 1) || m.width > 16384 || m.height > 16384) {
        return false;
    }
    return m.subsamp == DSV_SUBSAMP_444 || m.subsamp == DSV_SUBSAMP_422 || m.subsamp == DSV_SUBSAMP_420 ||
           m.subsamp == DSV_SUBSAMP_411;
}

static int read_packet_hdr(BitReader &br) /* dsv_decoder.c:21-48 */
{
    int c0 = (int) br.get_bits(8), c1 = (int) br.get_bits(8), c2 = (int) br.get_bits(8), c3 = (int) br.get_bits(8);
    if (c0 != DSV_FOURCC_0 || c1 != DSV_FOURCC_1 || c2 != DSV_FOURCC_2 || c3 != DSV_FOURCC_3) {
        DSV_ERROR(("bad 4cc (%c %c %c %c)\n", c0, c1, c2, c3));
        return -1;
    }
    br.get_bits(8); /* minor version */
    int type = (int) br.get_bits(8);
    br.get_bits(32);
    br.get_bits(32);
    return type;
}

/* B.2.3.1 stability map (dsv_decoder.c:126-145) */
static void read_stability(BitReader &br, const uint8_t *pkt, unsigned pkt_len, uint8_t *stab, int nblk)
{
    br.align();
    unsigned len = br.get_ueg();
    br.align();
    unsigned at = br.byte_pos();
    RleReader rle(pkt + (at < pkt_len ? at : pkt_len), at < pkt_len ? pkt_len - at : 0);
    br.skip_bytes(len);
    for (int i = 0; i < nblk; i++) {
        stab[i] = (uint8_t) rle.get();
    }
}

/* B.2.3.2 motion data (dsv_decoder.c:72-124) */
static void read_motion(BitReader &br, const uint8_t *pkt, unsigned pkt_len, DevMV *mvs, uint8_t *stab, int nbh, int nbv)
{
    unsigned start[4];
    br.align();
    for (int s = 0; s < 4; s++) {
        unsigned len = br.get_ueg();
        br.align();
        start[s] = br.byte_pos() < pkt_len ? br.byte_pos() : pkt_len;
        br.skip_bytes(len);
    }
    RleReader mode(pkt + start[0], pkt_len - start[0]);
    BitReader bx(pkt + start[1], pkt_len - start[1]), by(pkt + start[2], pkt_len - start[2]);
    BitReader bm(pkt + start[3], pkt_len - start[3]);
    memset(mvs, 0, sizeof(DevMV) * (size_t) nbh * nbv);
    for (int j = 0; j < nbv; j++) {
        for (int i = 0; i < nbh; i++) {
            DevMV &mv = mvs[j * nbh + i];
            mv.mode = (uint8_t) mode.get();
            if (mv.mode == DSV_MODE_INTER) {
                int px, py;
                predict_mv(mvs, nbh, i, j, &px, &py);
                mv.x = (int16_t) (bx.get_seg() + px);
                mv.y = (int16_t) (by.get_seg() + py);
            } else {
                mv.submask = bm.get_bit() ? DSV_MASK_ALL_INTRA : (uint8_t) bm.get_bits(4);
                stab[j * nbh + i] |= 2;
            }
        }
    }
}

extern "C" DSV_META *dsv_get_metadata(DSV_DECODER *d)
{
    DSV_META *m = (DSV_META *) dsv_alloc(sizeof(DSV_META));
    *m = d->vidmeta;
    return m;
}

extern "C" void dsv_dec_free(DSV_DECODER *d)
{
    dec_ctx_destroy(dec_ctx(d));
    d->ref = NULL;
}

extern "C" int dsv_dec(DSV_DECODER *d, DSV_BUF *buffer, DSV_FRAME **out, DSV_FNUM *fn)
{
    *fn = (DSV_FNUM) -1;
    const uint8_t *pkt = buffer->data;
    const unsigned pkt_len = buffer->len;
    BitReader br(pkt, pkt_len);
    const int pkt_type = read_packet_hdr(br);
    if (pkt_type == -1) {
        dsv_buf_free(buffer);
        return DSV_DEC_ERROR;
    }
    if (!DSV_PT_IS_PIC(pkt_type)) {
        int ret = DSV_DEC_ERROR;
        if (pkt_type == DSV_PT_META) { /* dsv_decoder.c:50-70 */
            DSV_META *m = &d->vidmeta;
            m->width = (int) br.get_ueg();
            m->height = (int) br.get_ueg();
            m->subsamp = (int) br.get_ueg();
            m->fps_num = (int) br.get_ueg();
            m->fps_den = (int) br.get_ueg();
            m->aspect_num = (int) br.get_ueg();
            m->aspect_den = (int) br.get_ueg();
            d->got_metadata = 1;
            ret = DSV_DEC_GOT_META;
        } else if (pkt_type == DSV_PT_EOS) {
            ret = DSV_DEC_EOS;
        }
        dsv_buf_free(buffer);
        return ret;
    }
    if (!d->got_metadata) {
        DSV_WARNING(("no metadata, skipping frame"));
        dsv_buf_free(buffer);
        return DSV_DEC_OK;
    }
    const DSV_META &md = d->vidmeta;
    if (!meta_supported(md)) {
        DSV_ERROR(("unsupported picture format %dx%d subsamp %d", md.width, md.height, md.subsamp));
        dsv_buf_free(buffer);
        return DSV_DEC_ERROR;
    }
    DecCtx *c = dec_ctx(d);
    if (c && (c->g.w != md.width || c->g.h != md.height || c->g.subsamp != md.subsamp)) {
        dec_ctx_destroy(c); /* new sequence parameters: references of the old size are useless */
        c = nullptr;
    }
    if (!c) {
        c = dec_ctx_create(md);
        d->ref = reinterpret_cast<DSV_IMAGE *>(c);
    }

    const int has_ref = DSV_PT_HAS_REF(pkt_type), is_ref = DSV_PT_IS_REF(pkt_type);
    br.align();
    const DSV_FNUM fno = br.get_bits(32);
    br.align();
    const int blk_w = (int) (br.get_ueg() << 2), blk_h = (int) (br.get_ueg() << 2);
    if (blk_w < DSV_MIN_BLOCK_SIZE || blk_h < DSV_MIN_BLOCK_SIZE || blk_w > DSV_MAX_BLOCK_SIZE || blk_h > DSV_MAX_BLOCK_SIZE) {
        dsv_buf_free(buffer);
        return DSV_DEC_ERROR;
    }
    plan_blocks(&c->g, blk_w, blk_h);
    const CodecGeom &g = c->g;
    cudaStream_t st = c->st;

    read_stability(br, pkt, pkt_len, c->h_stab, g.nblk);
    if (has_ref) {
        read_motion(br, pkt, pkt_len, c->h_mv, c->h_stab, g.nbh, g.nbv);
    }
    br.align();
    const int quant = (int) br.get_bits(DSV_MAX_QP_BITS);

    /* plane directory (dsv_decoder.c:383-413) */
    HzPlaneData pd[3];
    int nplanes = 0;
    for (int p = 0; p < 3; p++) {
        br.align();
        const int plen = (int) br.get_bits(32);
        br.align();
        const int framesz = g.cw[p] * g.ch[p] * (int) sizeof(int32_t);
        if (plen <= 0 || plen > framesz * 2) {
            DSV_ERROR(("plane length was strange: %d", plen));
            break;
        }
        const unsigned at = br.byte_pos();
        if (at >= pkt_len) {
            DSV_ERROR(("plane starts past the end of the packet"));
            break;
        }
        hzdec_parse_head(pkt + at, pkt_len - at, (unsigned) plen, &pd[p]);
        pd[p].body = c->d_pkt + at;
        br.skip_bytes((unsigned) plen);
        nplanes++;
    }

    /* ---- device side ---- */
    if ((size_t) pkt_len > c->pkt_cap) {
        DSV_ERROR(("packet larger than any valid picture (%u bytes)", pkt_len));
        dsv_buf_free(buffer);
        return DSV_DEC_ERROR;
    }
    memcpy(c->h_pkt, pkt, pkt_len);
    memset(c->h_pkt + pkt_len, 0, 64);
    CUDA_CHECK(cudaMemcpyAsync(c->d_pkt, c->h_pkt, (size_t) pkt_len + 64, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(c->cb.d_stab, c->h_stab, (size_t) g.nblk, cudaMemcpyHostToDevice, st));
    if (has_ref) {
        CUDA_CHECK(cudaMemcpyAsync(c->d_mv, c->h_mv, sizeof(DevMV) * (size_t) g.nblk, cudaMemcpyHostToDevice, st));
    }
    const DevFrame &cur = c->out[c->cur];
    const DevFrame &prev = c->out[c->cur ^ 1];
    coder_setup_jobs(&c->cb, g, cur, quant, has_ref, 0, st);
    CUDA_CHECK(cudaMemsetAsync(c->cb.coef, 0, g.coef_total * sizeof(int32_t), st)); /* dsv_decoder.c:405 */
    if (nplanes > 0) {
        hzdec_launch(&c->hz, c->cb.hj, pd, nplanes, st);
    }
    /* planes that were never coded stay as an all-zero coefficient plane here; the reference leaves
     * the (zeroed) residual plane untouched instead -- only reachable with a corrupt plen */
    sbt_inv_launch(c->cb.d_sjobs, 3, c->cb.total_tiles, c->cb.lo_smem, !has_ref, st);

    *fn = fno;
    if (has_ref) {
        if (!c->have_ref) {
            DSV_WARNING(("reference frame not found"));
            CUDA_CHECK(cudaStreamSynchronize(st));
            return DSV_DEC_ERROR; /* the reference also keeps the packet buffer on this path (dsv_decoder.c:424-427) */
        }
        MotionGeom mg = {g.w, g.h, g.hs, g.vs, g.blk_w, g.blk_h, g.nbh, g.nbv, 0};
        bmc_launch(mg, c->d_mv, prev, nullptr, cur, 2, st);
    }
    if (is_ref) {
        frame_extend_launch(cur, 3, st); /* dsv_decoder.c:438-440 */
    }
    /* output: packed planes -> pinned staging -> a host frame the caller owns one reference of */
    {
        uint8_t *o = c->h_out;
        for (int p = 0; p < 3; p++) {
            CUDA_CHECK(cudaMemcpy2DAsync(o, g.pw[p], cur.p[p], cur.stride[p], g.pw[p], g.ph[p], cudaMemcpyDeviceToHost, st));
            o += (size_t) g.pw[p] * g.ph[p];
        }
    }
    DSV_FRAME *f = dsv_mk_frame(md.subsamp, md.width, md.height, 1);
    CUDA_CHECK(cudaStreamSynchronize(st));
    {
        const uint8_t *s = c->h_out;
        for (int p = 0; p < 3; p++) {
            DSV_PLANE *pl = &f->planes[p];
            for (int y = 0; y < pl->h; y++) {
                memcpy(DSV_GET_LINE(pl, y), s, (size_t) pl->w);
                s += pl->w;
            }
        }
    }
    if (is_ref) {
        c->have_ref = 1;
        c->cur ^= 1;
    }
    dsv_buf_free(buffer);
    *out = f;
    return DSV_DEC_OK;
}
