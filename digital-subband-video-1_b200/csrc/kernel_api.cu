/*
 * kernel_api.cu -- dsvk_*: kernel-level C ABI with HOST buffers (declared in include/dsv1_b200_kernels.h).
 *
 * One entry point per hot-path subsystem, with the same flat signatures as the checker libraries
 * (oracle/ref_harness.c, oracle/dsv1_port.c) so the parity tests call all three alike.  Every call
 * copies its inputs to the GPU, runs the CUDA kernels and copies the result back: there is no host
 * implementation behind these symbols.
 */
#include "common.cuh"
#include "sbt.cuh"
#include "hzcc.cuh"
#include "hzcc_dec.cuh"

#include <vector>

using namespace dsv;

namespace {

struct DevBuf {
    void *p = nullptr;
    explicit DevBuf(size_t n) { CUDA_CHECK(cudaMalloc(&p, n ? n : 1)); }
    ~DevBuf() { cudaFree(p); }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

/* device plane in the reference's bordered layout: returns pointer to sample (0,0) */
struct DevPlane {
    DevBuf buf;
    int stride, rows;
    uint8_t *origin;
    DevPlane(int w, int h)
        : buf((size_t) frame_stride(w) * (h + 2 * DSV_BORDER) + 256), stride(frame_stride(w)), rows(h + 2 * DSV_BORDER)
    {
        CUDA_CHECK(cudaMemset(buf.p, 0, (size_t) stride * rows + 256));
        origin = buf.as<uint8_t>() + (size_t) stride * DSV_BORDER + DSV_BORDER;
    }
};

} // namespace

extern "C" int dsvk_fwd_sbt(const uint8_t *pix, int stride, int pw, int ph, int cw, int ch, int isP, int32_t *coef_out)
{
    DSV_API_BEGIN
    if ((cw & 1) || (ch & 1) || cw < 16 || ch < 16) {
        return -1;
    }
    DevPlane dp(cw, ph);
    CUDA_CHECK(cudaMemcpy2D(dp.origin, dp.stride, pix, stride, cw, ph, cudaMemcpyHostToDevice));
    DevBuf coef((size_t) cw * ch * 4), llx(sbt_llx_elems(cw, ch) * 4), dv(sbt_dv_elems(cw, ch) * 4), jobs(sizeof(SbtJob));
    CUDA_CHECK(cudaMemset(coef.p, 0, (size_t) cw * ch * 4));
    SbtJob j;
    memset(&j, 0, sizeof(j));
    sbt_fill_geometry(&j, pw, ph, cw, ch, isP, 0);
    j.pix = dp.origin;
    j.pstride = dp.stride;
    j.coef = coef.as<int32_t>();
    j.llx = llx.as<int32_t>();
    j.dv = dv.as<int32_t>();
    j.do_quant = 0;
    const SbtDims dims = sbt_assign_tiles(&j, 1);
    CUDA_CHECK(cudaMemcpy(jobs.p, &j, sizeof(j), cudaMemcpyHostToDevice));
    sbt_fwd_launch(jobs.as<SbtJob>(), dims, sbt_lo_smem_bytes(cw, ch), 0);
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(coef_out, coef.p, (size_t) cw * ch * 4, cudaMemcpyDeviceToHost));
    return 0;
    DSV_API_END(-100)
}

extern "C" int dsvk_inv_sbt(int32_t *coef_io, int cw, int ch, int q, int isP, int c, uint8_t *pix_out, int stride, int pw, int ph)
{
    DSV_API_BEGIN
    if ((cw & 1) || (ch & 1) || cw < 16 || ch < 16) {
        return -1;
    }
    DevPlane dp(cw, ph);
    DevBuf coef((size_t) cw * ch * 4), llx(sbt_llx_elems(cw, ch) * 4), jobs(sizeof(SbtJob));
    CUDA_CHECK(cudaMemcpy(coef.p, coef_io, (size_t) cw * ch * 4, cudaMemcpyHostToDevice));
    SbtJob j;
    memset(&j, 0, sizeof(j));
    sbt_fill_geometry(&j, pw, ph, cw, ch, isP, c);
    sbt_fill_quant(&j, q, isP, c, 1, 1);
    j.pix = dp.origin;
    j.pstride = dp.stride;
    j.opix = dp.origin;
    j.ostride = dp.stride;
    j.coef = coef.as<int32_t>();
    j.llx = llx.as<int32_t>();
    const SbtDims dims = sbt_assign_tiles(&j, 1);
    CUDA_CHECK(cudaMemcpy(jobs.p, &j, sizeof(j), cudaMemcpyHostToDevice));
    sbt_inv_launch(jobs.as<SbtJob>(), dims, sbt_lo_smem_bytes(cw, ch), 0);
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy2D(pix_out, stride, dp.origin, dp.stride, pw, ph, cudaMemcpyDeviceToHost));
    return 0;
    DSV_API_END(-100)
}

/* forward transform with the fused quantiser, exactly as the encoder runs it (do_quant = 1) */
extern "C" int dsvk_fwd_sbt_q(const uint8_t *pix, int stride, int pw, int ph, int cw, int ch, int isP, int c, int q,
                              const uint8_t *stable, int nbh, int nbv, int32_t *coef_out, int32_t *dv_out)
{
    DSV_API_BEGIN
    if ((cw & 1) || (ch & 1) || cw < 16 || ch < 16) {
        return -1;
    }
    DevPlane dp(cw, ph);
    CUDA_CHECK(cudaMemcpy2D(dp.origin, dp.stride, pix, stride, cw, ph, cudaMemcpyHostToDevice));
    DevBuf coef((size_t) cw * ch * 4), llx(sbt_llx_elems(cw, ch) * 4), dv(sbt_dv_elems(cw, ch) * 4), jobs(sizeof(SbtJob));
    DevBuf stab((size_t) nbh * nbv);
    CUDA_CHECK(cudaMemset(coef.p, 0, (size_t) cw * ch * 4));
    CUDA_CHECK(cudaMemset(dv.p, 0, sbt_dv_elems(cw, ch) * 4));
    CUDA_CHECK(cudaMemcpy(stab.p, stable, (size_t) nbh * nbv, cudaMemcpyHostToDevice));
    SbtJob j;
    memset(&j, 0, sizeof(j));
    sbt_fill_geometry(&j, pw, ph, cw, ch, isP, c);
    sbt_fill_quant(&j, q, isP, c, nbh, nbv);
    j.pix = dp.origin;
    j.pstride = dp.stride;
    j.coef = coef.as<int32_t>();
    j.llx = llx.as<int32_t>();
    j.dv = dv.as<int32_t>();
    j.stable = stab.as<uint8_t>();
    j.do_quant = 1;
    const SbtDims dims = sbt_assign_tiles(&j, 1);
    CUDA_CHECK(cudaMemcpy(jobs.p, &j, sizeof(j), cudaMemcpyHostToDevice));
    sbt_fwd_launch(jobs.as<SbtJob>(), dims, sbt_lo_smem_bytes(cw, ch), 0);
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(coef_out, coef.p, (size_t) cw * ch * 4, cudaMemcpyDeviceToHost));
    if (dv_out) {
        CUDA_CHECK(cudaMemcpy(dv_out, dv.p, (size_t) j.dg.total * 4, cudaMemcpyDeviceToHost));
    }
    return j.dg.total;
    DSV_API_END(-100)
}

/* dsv_encode_plane semantics (hzcc.c:449-476): raw coefficients in, plane bytes out, coef <- dequantised */
extern "C" int dsvk_encode_plane(int32_t *coef_io, int cw, int ch, int q, int isP, int c, const uint8_t *stable,
                                 int nbh, int nbv, uint8_t *out, int out_cap)
{
    DSV_API_BEGIN
    if ((cw & 1) || (ch & 1) || cw < 16 || ch < 16) {
        return -1;
    }
    HzJob j;
    memset(&j, 0, sizeof(j));
    hz_fill_job(&j, cw, ch, q, isP, c, nbh, nbv);
    const size_t cap = (size_t) cw * ch * 8 + 256;
    DevBuf coef((size_t) cw * ch * 4), dv(sbt_dv_elems(cw, ch) * 4), stab((size_t) nbh * nbv), pkt(cap);
    DevBuf jobs(sizeof(HzJob)), chunks(sizeof(HzChunk) * j.nchunks), frames(sizeof(HzFrame));
    CUDA_CHECK(cudaMemcpy(coef.p, coef_io, (size_t) cw * ch * 4, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(stab.p, stable, (size_t) nbh * nbv, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemset(dv.p, 0, sbt_dv_elems(cw, ch) * 4));
    CUDA_CHECK(cudaMemset(pkt.p, 0, cap));
    j.coef = coef.as<int32_t>();
    j.dv = dv.as<int32_t>();
    j.stable = stab.as<uint8_t>();
    j.chunk_base = 0;
    j.frame = 0;
    HzFrame f;
    memset(&f, 0, sizeof(f));
    f.pkt = pkt.as<uint8_t>();
    f.start_byte = 0;
    f.cap = (unsigned) cap;
    f.nplanes = 1;
    CUDA_CHECK(cudaMemcpy(jobs.p, &j, sizeof(j), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(frames.p, &f, sizeof(f), cudaMemcpyHostToDevice));
    hzcc_quant_launch(jobs.as<HzJob>(), 1, cw * ch, 0);
    hzcc_enc_launch(jobs.as<HzJob>(), 1, chunks.as<HzChunk>(), j.nchunks, frames.as<HzFrame>(), 1, 0);
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(&f, frames.p, sizeof(f), cudaMemcpyDeviceToHost));
    if (f.overflow || (int) f.total_bytes > out_cap) {
        return -2;
    }
    CUDA_CHECK(cudaMemcpy(out, pkt.p, f.total_bytes, cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(coef_io, coef.p, (size_t) cw * ch * 4, cudaMemcpyDeviceToHost));
    return (int) f.total_bytes;
    DSV_API_END(-100)
}

/* dsv_decode_plane semantics (hzcc.c:478-496): `in` points just after the 32-bit plen field */
extern "C" int dsvk_decode_plane(const uint8_t *in, int plen, int cw, int ch, int q, int isP, int c,
                                 const uint8_t *stable, int nbh, int nbv, int32_t *coef_out)
{
    DSV_API_BEGIN
    if ((cw & 1) || (ch & 1) || cw < 16 || ch < 16 || plen <= 0) {
        return -1;
    }
    HzJob j;
    memset(&j, 0, sizeof(j));
    hz_fill_job(&j, cw, ch, q, isP, c, nbh, nbv);
    DevBuf coef((size_t) cw * ch * 4), stab((size_t) nbh * nbv), body((size_t) plen + 64);
    CUDA_CHECK(cudaMemset(coef.p, 0, (size_t) cw * ch * 4));
    CUDA_CHECK(cudaMemset(body.p, 0, (size_t) plen + 64));
    CUDA_CHECK(cudaMemcpy(body.p, in, (size_t) plen, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(stab.p, stable, (size_t) nbh * nbv, cudaMemcpyHostToDevice));
    j.coef = coef.as<int32_t>();
    j.stable = stab.as<uint8_t>();
    HzPlaneData pd;
    hzdec_parse_head(in, (unsigned) plen, (unsigned) plen, &pd);
    pd.body = body.as<uint8_t>();
    HzDecPlan pl;
    hzdec_plan(&pl, cw, ch);
    HzDecPlaneBufs bufs;
    hzdec_plane_alloc(&bufs, pl);
    HzDecDims dims;
    std::vector<uint8_t> slot(hzdec_job_size());
    hzdec_fill_job(slot.data(), j, pd, bufs, &dims);
    DevBuf djob(hzdec_job_size());
    CUDA_CHECK(cudaMemcpy(djob.p, slot.data(), hzdec_job_size(), cudaMemcpyHostToDevice));
    hzdec_launch_jobs(djob.p, dims, 0);
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(coef_out, coef.p, (size_t) cw * ch * 4, cudaMemcpyDeviceToHost));
    hzdec_plane_free(&bufs);
    return 0;
    DSV_API_END(-100)
}

/* ---- motion: pyramid, HME, BMC ------------------------------------------------------------------ */
#include "frame.cuh"
#include "motion.cuh"
#include "dsv1_b200.h"

namespace {

/* packed planar YUV (host) -> bordered device frame, borders replicated (dsv_frame_copy + extend) */
void upload_frame(DevFrame *f, const uint8_t *yuv, int w, int h, int subsamp)
{
    devframe_alloc(f, w, h, subsamp);
    const uint8_t *s = yuv;
    for (int c = 0; c < 3; c++) {
        CUDA_CHECK(cudaMemcpy2D(f->p[c], f->stride[c], s, f->w[c], f->w[c], f->h[c], cudaMemcpyHostToDevice));
        s += (size_t) f->w[c] * f->h[c];
    }
    frame_extend_launch(*f, 3, 0);
}

void download_frame(const DevFrame &f, uint8_t *yuv)
{
    uint8_t *o = yuv;
    for (int c = 0; c < 3; c++) {
        CUDA_CHECK(cudaMemcpy2D(o, f.w[c], f.p[c], f.stride[c], f.w[c], f.h[c], cudaMemcpyDeviceToHost));
        o += (size_t) f.w[c] * f.h[c];
    }
}

MotionGeom motion_geom(int w, int h, int subsamp, int blk_w, int blk_h, int levels)
{
    MotionGeom g;
    g.w = w;
    g.h = h;
    g.hs = (subsamp >> 2) & 3;
    g.vs = subsamp & 3;
    g.blk_w = blk_w;
    g.blk_h = blk_h;
    g.nbh = ceil_div(w, blk_w);
    g.nbv = ceil_div(h, blk_h);
    g.levels = levels;
    return g;
}

} // namespace

extern "C" int dsvk_pyramid(const uint8_t *yuv, int w, int h, int subsamp, int levels, uint8_t *out, int *out_w, int *out_h)
{
    DSV_API_BEGIN
    if (levels < 1 || levels > 5) {
        return -1;
    }
    DevFrame f[6];
    upload_frame(&f[0], yuv, w, h, subsamp);
    for (int l = 0; l < levels; l++) {
        devframe_alloc(&f[l + 1], ceil_shift(w, l + 1), ceil_shift(h, l + 1), subsamp);
        frame_down2_luma_launch(f[l], f[l + 1], 0);
    }
    CUDA_CHECK(cudaDeviceSynchronize());
    for (int l = 1; l <= levels; l++) {
        CUDA_CHECK(cudaMemcpy2D(out, f[l].w[0], f[l].p[0], f[l].stride[0], f[l].w[0], f[l].h[0], cudaMemcpyDeviceToHost));
        out += (size_t) f[l].w[0] * f[l].h[0];
        out_w[l - 1] = f[l].w[0];
        out_h[l - 1] = f[l].h[0];
    }
    for (int l = 0; l <= levels; l++) {
        devframe_free(&f[l]);
    }
    return 0;
    DSV_API_END(-100)
}

extern "C" int dsvk_hme(const uint8_t *src_yuv, const uint8_t *ref_yuv, int w, int h, int subsamp, int blk_w, int blk_h,
                        int levels, void *mv_out)
{
    DSV_API_BEGIN
    if (levels < 0 || levels > 5) {
        return -1;
    }
    const MotionGeom g = motion_geom(w, h, subsamp, blk_w, blk_h, levels);
    const int nblk = g.nbh * g.nbv;
    DevFrame sf[6], rf[6];
    upload_frame(&sf[0], src_yuv, w, h, subsamp);
    upload_frame(&rf[0], ref_yuv, w, h, subsamp);
    for (int l = 0; l < levels; l++) {
        devframe_alloc(&sf[l + 1], ceil_shift(w, l + 1), ceil_shift(h, l + 1), subsamp);
        devframe_alloc(&rf[l + 1], ceil_shift(w, l + 1), ceil_shift(h, l + 1), subsamp);
        frame_down2_luma_launch(sf[l], sf[l + 1], 0);
        frame_down2_luma_launch(rf[l], rf[l + 1], 0);
    }
    DevMV *mvf[6];
    for (int l = 0; l <= levels; l++) {
        CUDA_CHECK(cudaMalloc(&mvf[l], sizeof(DevMV) * (size_t) nblk));
        CUDA_CHECK(cudaMemset(mvf[l], 0, sizeof(DevMV) * (size_t) nblk));
    }
    DevBuf aux(sizeof(int2) * (size_t) nblk), cnt(sizeof(int)), dargs(sizeof(HmeArgs) * (size_t) (levels + 1));
    CUDA_CHECK(cudaMemset(cnt.p, 0, sizeof(int)));
    std::vector<HmeArgs> ha((size_t) levels + 1);
    for (int l = 0; l <= levels; l++) {
        hme_fill_args(&ha[(size_t) l], g, l, sf, rf, mvf, aux.as<int2>(), cnt.as<int>());
    }
    CUDA_CHECK(cudaMemcpy(dargs.p, ha.data(), sizeof(HmeArgs) * ha.size(), cudaMemcpyHostToDevice));
    hme_launch(dargs.as<HmeArgs>(), 1, g, 0);
    CUDA_CHECK(cudaDeviceSynchronize());
    int nintra = 0;
    CUDA_CHECK(cudaMemcpy(&nintra, cnt.p, sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(mv_out, mvf[0], sizeof(DevMV) * (size_t) nblk, cudaMemcpyDeviceToHost));
    for (int l = 0; l <= levels; l++) {
        cudaFree(mvf[l]);
        devframe_free(&sf[l]);
        devframe_free(&rf[l]);
    }
    return nintra * 100 / nblk;
    DSV_API_END(-100)
}

/* a caller's host frame -> device frame, border included when the caller's frame has one (the search reads it) */
static void upload_host_frame(DevFrame *f, const DSV_FRAME *src)
{
    devframe_alloc(f, src->width, src->height, src->format);
    for (int c = 0; c < 3; c++) {
        const DSV_PLANE &pl = src->planes[c];
        if (src->border && pl.stride >= pl.w + 2 * DSV_BORDER) {
            const int rows = pl.h + 2 * DSV_BORDER, cols = pl.w + 2 * DSV_BORDER;
            CUDA_CHECK(cudaMemcpy2D(f->p[c] - (ptrdiff_t) DSV_BORDER * f->stride[c] - DSV_BORDER, (size_t) f->stride[c],
                                    pl.data - (ptrdiff_t) DSV_BORDER * pl.stride - DSV_BORDER, (size_t) pl.stride, (size_t) cols, (size_t) rows,
                                    cudaMemcpyHostToDevice));
        } else {
            CUDA_CHECK(cudaMemcpy2D(f->p[c], (size_t) f->stride[c], pl.data, (size_t) pl.stride, (size_t) pl.w, (size_t) pl.h, cudaMemcpyHostToDevice));
        }
    }
    if (!src->border) {
        frame_extend_launch(*f, 3, 0);
    }
}

/* drop-in for the reference's exported dsv_hme (dsv_encoder.h:122-132, hme.c:730-741) */
extern "C" int dsv_hme(DSV_HME *hme)
{
    DSV_API_BEGIN
    const DSV_PARAMS *pr = hme->params;
    const int levels = hme->levels;
    if (levels < 0 || levels > DSV_MAX_PYRAMID_LEVELS) {
        return 0;
    }
    const DSV_FRAME *top = hme->src[0];
    MotionGeom g = motion_geom(top->width, top->height, top->format, pr->blk_w, pr->blk_h, levels);
    g.nbh = pr->nblocks_h;
    g.nbv = pr->nblocks_v;
    const int nblk = g.nbh * g.nbv;
    DevFrame sf[DSV_MAX_PYRAMID_LEVELS + 1], rf[DSV_MAX_PYRAMID_LEVELS + 1];
    DevMV *mvf[DSV_MAX_PYRAMID_LEVELS + 1];
    for (int l = 0; l <= levels; l++) {
        upload_host_frame(&sf[l], hme->src[l]);
        upload_host_frame(&rf[l], hme->ref[l]);
        CUDA_CHECK(cudaMalloc(&mvf[l], sizeof(DevMV) * (size_t) nblk));
        CUDA_CHECK(cudaMemset(mvf[l], 0, sizeof(DevMV) * (size_t) nblk));
    }
    DevBuf aux(sizeof(int2) * (size_t) nblk), cnt(sizeof(int)), dargs(sizeof(HmeArgs) * (size_t) (levels + 1));
    CUDA_CHECK(cudaMemset(cnt.p, 0, sizeof(int)));
    std::vector<HmeArgs> ha((size_t) levels + 1);
    for (int l = 0; l <= levels; l++) {
        hme_fill_args(&ha[(size_t) l], g, l, sf, rf, mvf, aux.as<int2>(), cnt.as<int>());
    }
    CUDA_CHECK(cudaMemcpy(dargs.p, ha.data(), sizeof(HmeArgs) * ha.size(), cudaMemcpyHostToDevice));
    hme_launch(dargs.as<HmeArgs>(), 1, g, 0);
    CUDA_CHECK(cudaDeviceSynchronize());
    int nintra = 0;
    CUDA_CHECK(cudaMemcpy(&nintra, cnt.p, sizeof(int), cudaMemcpyDeviceToHost));
    static_assert(sizeof(DSV_MV) == sizeof(DevMV), "DevMV is the device twin of DSV_MV");
    for (int l = 0; l <= levels; l++) {
        hme->mvf[l] = (DSV_MV *) dsv_alloc((int) (sizeof(DSV_MV) * (size_t) nblk));
        CUDA_CHECK(cudaMemcpy(hme->mvf[l], mvf[l], sizeof(DevMV) * (size_t) nblk, cudaMemcpyDeviceToHost));
        cudaFree(mvf[l]);
        devframe_free(&sf[l]);
        devframe_free(&rf[l]);
    }
    return nintra * 100 / nblk;
    DSV_API_END(-100)
}

extern "C" int dsvk_sub_pred(const void *mvs, int w, int h, int subsamp, int blk_w, int blk_h, const uint8_t *inp_yuv,
                             const uint8_t *ref_yuv, uint8_t *pred_out, uint8_t *resid_out)
{
    DSV_API_BEGIN
    const MotionGeom g = motion_geom(w, h, subsamp, blk_w, blk_h, 0);
    DevFrame inp, ref, pred;
    upload_frame(&inp, inp_yuv, w, h, subsamp);
    upload_frame(&ref, ref_yuv, w, h, subsamp);
    devframe_alloc(&pred, w, h, subsamp);
    DevBuf mv(sizeof(DevMV) * (size_t) g.nbh * g.nbv);
    CUDA_CHECK(cudaMemcpy(mv.p, mvs, sizeof(DevMV) * (size_t) g.nbh * g.nbv, cudaMemcpyHostToDevice));
    BmcArgs ba;
    bmc_fill_args(&ba, g, mv.as<DevMV>(), ref, &pred, inp, inp, 1);
    DevBuf dargs(sizeof(BmcArgs));
    CUDA_CHECK(cudaMemcpy(dargs.p, &ba, sizeof(ba), cudaMemcpyHostToDevice));
    bmc_launch(dargs.as<BmcArgs>(), 1, g, 0);
    CUDA_CHECK(cudaDeviceSynchronize());
    download_frame(pred, pred_out);
    download_frame(inp, resid_out);
    devframe_free(&inp);
    devframe_free(&ref);
    devframe_free(&pred);
    return 0;
    DSV_API_END(-100)
}

extern "C" int dsvk_add_pred(const void *mvs, int w, int h, int subsamp, int blk_w, int blk_h, const uint8_t *resid_yuv,
                             const uint8_t *ref_yuv, uint8_t *out_yuv)
{
    DSV_API_BEGIN
    const MotionGeom g = motion_geom(w, h, subsamp, blk_w, blk_h, 0);
    DevFrame io, ref;
    upload_frame(&io, resid_yuv, w, h, subsamp);
    upload_frame(&ref, ref_yuv, w, h, subsamp);
    DevBuf mv(sizeof(DevMV) * (size_t) g.nbh * g.nbv);
    CUDA_CHECK(cudaMemcpy(mv.p, mvs, sizeof(DevMV) * (size_t) g.nbh * g.nbv, cudaMemcpyHostToDevice));
    BmcArgs ba;
    bmc_fill_args(&ba, g, mv.as<DevMV>(), ref, nullptr, io, io, 2);
    DevBuf dargs(sizeof(BmcArgs));
    CUDA_CHECK(cudaMemcpy(dargs.p, &ba, sizeof(ba), cudaMemcpyHostToDevice));
    bmc_launch(dargs.as<BmcArgs>(), 1, g, 0);
    CUDA_CHECK(cudaDeviceSynchronize());
    download_frame(io, out_yuv);
    devframe_free(&io);
    devframe_free(&ref);
    return 0;
    DSV_API_END(-100)
}
