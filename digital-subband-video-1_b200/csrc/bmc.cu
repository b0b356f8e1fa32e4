/*
 * bmc.cu -- half-pel block motion compensation, fused with residual formation (encoder) or
 * reconstruction (decoder).  Replaces compensate / hpelL / hpel / avgval / cpyzero / subf / addf /
 * dsv_sub_pred / dsv_add_pred (bmc.c:29-346).
 *
 * Streaming formulation: the unit of work is a strip of 16 samples x R rows of ONE motion block (R = 8 luma,
 * 4 chroma), one strip per thread, the strips of all blocks / planes / lanes flattened into one grid.  A thread
 * keeps everything in registers -- no shared memory, no barriers:
 *
 *   luma    every inter block goes through the SAME separable 4-tap path whatever its half-pel phase: the phase
 *           only selects the tap words, (-1,9,9,-1) or (0,16,0,0), per direction.  With 16 = the taps' DC gain
 *           the four cases of hpelL (bmc.c:124-174) are reproduced exactly:
 *             (16 (9(b+c)-(a+d)) + 128) >> 8 == (9(b+c)-(a+d) + 8) >> 4   and   (256 p + 128) >> 8 == p.
 *           So warps never diverge on the phase.  The H pass is one dp4a per sample on funnel-shifted words of the
 *           row; rows slide through a register window of vertically paired 16-bit H results, the V pass is two
 *           dp2a per sample.
 *   chroma  bilinear (bmc.c:58-110) as one dp4a per sample with phase-selected weights summing to 4:
 *             (4a + 2) >> 2 == a,  (2a + 2b + 2) >> 2 == (a + b + 1) >> 1,  (a + b + c + d + 2) >> 2.
 *   intra   (bmc.c:255-298) blocks take the co-located reference through the same path (zero vector) and
 *           overwrite the flagged quadrants with the block / quadrant means that bmc_means_kernel (one warp per
 *           intra block, a launch that exits at once for inter blocks) left in a small table.
 *
 * The prediction never makes a round trip through HBM on the decoder side (mode 2: io = clamp(pred + io - 128));
 * the encoder keeps it (mode 1) because the closed-loop reconstruction adds it back in the inverse transform's
 * store (sbt_inv.cu).  Plane rows are 16-byte aligned, so whole strips move as 16-byte loads / stores; strips cut
 * by the picture edge or by block widths that are not multiples of 16 fall back to words / bytes.
 *
 * Reads outside the picture go through the 64-sample replicated border exactly like the reference
 * (position clamp bmc.c:221-249); filter taps that step one sample past the border see the same
 * neighbouring bytes because DevFrame keeps the reference's strides and plane order.  Taps with weight zero
 * read (and ignore) bytes the reference does not touch; they stay inside the frame allocation's guard bands.
 */
#include "motion.cuh"

namespace dsv {

#define BMC_THREADS 256
#define BMC_RL 8 /* luma rows per strip */
#define BMC_RC 4 /* chroma rows per strip */

/* c + sum of the four unsigned bytes of w times the four signed bytes of taps */
DSV_HD int dp4a_us(unsigned w, unsigned taps, int c)
{
#if defined(__CUDA_ARCH__)
    int r;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(taps), "r"(c));
    return r;
#else
    int r = c;
    for (int i = 0; i < 4; i++) {
        r += (int) ((w >> (8 * i)) & 0xffu) * (int) (int8_t) ((taps >> (8 * i)) & 0xffu);
    }
    return r;
#endif
}

/* out = clamp(in -/+ pred +/- 128) on four packed samples */
DSV_D unsigned bmc_combine4(unsigned cur, unsigned pw, int mode)
{
    const int p0 = byte_of(pw, 0), p1 = byte_of(pw, 1), p2 = byte_of(pw, 2), p3 = byte_of(pw, 3);
    if (mode == 1) {
        return pack_u8x4(byte_of(cur, 0) - p0 + 128, byte_of(cur, 1) - p1 + 128, byte_of(cur, 2) - p2 + 128, byte_of(cur, 3) - p3 + 128);
    }
    return pack_u8x4(byte_of(cur, 0) + p0 - 128, byte_of(cur, 1) + p1 - 128, byte_of(cur, 2) + p2 - 128, byte_of(cur, 3) + p3 - 128);
}

/* one row of a strip: 16 predicted samples in pw[4] (sample 0 in the lowest byte of pw[0]); n = valid samples */
DSV_D void bmc_store_row(const BmcPlane &P, int mode, int gx, int gy, int n, bool vec16, const unsigned pw[4])
{
    const size_t io = (size_t) gy * P.istride + gx, oo = (size_t) gy * P.ostride + gx;
    if (vec16) { /* n == 16 and every row of the three frames 16-byte aligned at gx */
        const uint4 cur = *reinterpret_cast<const uint4 *>(P.in + io);
        uint4 o;
        o.x = bmc_combine4(cur.x, pw[0], mode);
        o.y = bmc_combine4(cur.y, pw[1], mode);
        o.z = bmc_combine4(cur.z, pw[2], mode);
        o.w = bmc_combine4(cur.w, pw[3], mode);
        *reinterpret_cast<uint4 *>(P.out + oo) = o;
        if (P.pred) {
            *reinterpret_cast<uint4 *>(P.pred + (size_t) gy * P.pstride + gx) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (4 * k + e < n) {
                const int pv = byte_of(pw[k], e);
                const int cur = P.in[io + 4 * k + e];
                P.out[oo + 4 * k + e] = mode == 1 ? clamp_u8(cur - pv + 128) : clamp_u8(pv + cur - 128);
                if (P.pred) {
                    P.pred[(size_t) gy * P.pstride + gx + 4 * k + e] = (uint8_t) pv;
                }
            }
        }
    }
}

/* intra blocks (bmc.c:255-298): the flagged quadrants become their mean, the others keep the co-located reference
 * already in pw; odd edge blocks: the quadrants do not cover the last row / column (zeroed frame in the reference) */
DSV_D void bmc_intra_row(unsigned pw[4], const DevMV &mv, unsigned means, int lx, int ly, int cw, int ch)
{
    const bool whole = mv.submask == 15;
    const int sbw = cw / 2, sbh = ch / 2;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        unsigned w = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int x = lx + 4 * k + e;
            unsigned v;
            if (whole) {
                v = means & 0xffu;
            } else if (x >= 2 * sbw || ly >= 2 * sbh) {
                v = 0;
            } else {
                const int q = (x >= sbw ? 1 : 0) | (ly >= sbh ? 2 : 0);
                v = (mv.submask & (1 << q)) ? ((means >> (8 * q)) & 0xffu) : (unsigned) byte_of(pw[k], e);
            }
            w |= v << (8 * e);
        }
        pw[k] = w;
    }
}

struct BmcStrip {
    int b;          /* block index */
    int x, y;       /* block origin in the plane */
    int cw, ch;     /* clipped block size */
    int lx, ly;     /* strip origin inside the block */
    int n, rows;    /* valid samples per row / rows */
    int px, py;     /* clamped integer reference position of the block */
    int xh, yh;     /* half-pel flags */
    bool intra, vec16;
    DevMV mv;
};

template <int R> DSV_D bool bmc_strip_setup(const BmcArgs &a, const int c, const int u, BmcStrip &s)
{
    const BmcPlane &P = a.pl[c];
    const int sh = c ? a.hs : 0, sv = c ? a.vs : 0;
    const int bw = a.blk_w >> sh, bh = a.blk_h >> sv;
    const int segs = (bw + 15) >> 4, rgs = (bh + R - 1) / R, upb = segs * rgs;
    if (u >= upb * a.nbh * a.nbv) {
        return false;
    }
    s.b = u / upb;
    const int r = u - s.b * upb;
    const int rg = r / segs, seg = r - rg * segs;
    const int j = s.b / a.nbh, i = s.b - j * a.nbh;
    s.x = i * bw;
    s.y = j * bh;
    s.cw = (s.x + bw >= P.w) ? P.w - s.x : bw;
    s.ch = (s.y + bh >= P.h) ? P.h - s.y : bh;
    s.lx = 16 * seg;
    s.ly = R * rg;
    if (s.lx >= s.cw || s.ly >= s.ch) {
        return false;
    }
    s.n = imin(16, s.cw - s.lx);
    s.rows = imin(R, s.ch - s.ly);
    s.mv = a.mv[s.b];
    s.intra = s.mv.mode != 0;
    const int dx = s.intra ? 0 : (s.mv.x >> sh), dy = s.intra ? 0 : (s.mv.y >> sv);
    s.px = iclamp(s.x + (dx >> 1), -DSV_BORDER, (P.w - bw) + DSV_BORDER - 1); /* bmc.c:240-249 */
    s.py = iclamp(s.y + (dy >> 1), -DSV_BORDER, (P.h - bh) + DSV_BORDER - 1);
    if (s.intra) { /* co-located, never clamped (bmc.c:262,287) */
        s.px = s.x;
        s.py = s.y;
    }
    s.xh = dx & 1;
    s.yh = dy & 1;
    const uintptr_t al = reinterpret_cast<uintptr_t>(P.in) | reinterpret_cast<uintptr_t>(P.out) | reinterpret_cast<uintptr_t>(P.pred) |
                         (uintptr_t) (unsigned) (P.istride | P.ostride | P.pstride | (s.x + s.lx));
    s.vec16 = s.n == 16 && (al & 15) == 0;
    return true;
}

/* the forward transform of a plane with odd width reads one column past it (sbt.c:583-591); in the reference
 * that column of the residual frame still holds the replicated INPUT border (dsv_encoder.c:657-659: xf = copy
 * of the padded input, then only w x h is replaced) */
DSV_D void bmc_edge_column(const BmcPlane &P, const BmcStrip &s, int mode)
{
    if (mode == 1 && s.x + s.cw == P.w && s.lx + s.n == s.cw) {
        for (int k = 0; k < s.rows; k++) {
            const int gy = s.y + s.ly + k;
            P.out[(size_t) gy * P.ostride + P.w] = P.in[(size_t) gy * P.istride + P.w];
        }
    }
}

DSV_D void bmc_luma_strip(const BmcArgs &a, const int u)
{
    BmcStrip s = {};
    if (!bmc_strip_setup<BMC_RL>(a, 0, u, s)) {
        return;
    }
    const BmcPlane &P = a.pl[0];
    const unsigned means = s.intra ? a.means[s.b] : 0u;
    /* taps: bytes (p[-1], p[0], p[1], p[2]); pairs (row a, row b) and (row c, row d) */
    const unsigned th = s.xh ? 0xff0909ffu : 0x00001000u;
    const unsigned t_ab = s.yh ? 0x09ffu : 0x1000u, t_cd = s.yh ? 0xff09u : 0x0000u;
    const int rs = P.rstride;
    /* first byte needed: one row above and one sample left of the strip's reference position */
    const uint8_t *src = P.ref + (ptrdiff_t) (s.py + s.ly - 1) * rs + (s.px + s.lx - 1);
    const unsigned bsh = ((unsigned) reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
    const unsigned *q = reinterpret_cast<const unsigned *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t) 3);
    const int qs = rs >> 2; /* strides are multiples of 16 */

    int hprev[16];
    unsigned pr0[16], pr1[16]; /* vertical pairs (row r-3, r-2) and (row r-2, r-1) of 16-bit H results */
#pragma unroll
    for (int e = 0; e < 16; e++) {
        hprev[e] = 0;
        pr0[e] = pr1[e] = 0;
    }
#pragma unroll
    for (int r = 0; r < BMC_RL + 3; r++) {
        if (r >= s.rows + 3) {
            break;
        }
        {
            const unsigned *qr = q + (ptrdiff_t) r * qs;
            const unsigned w0 = qr[0], w1 = qr[1], w2 = qr[2], w3 = qr[3], w4 = qr[4], w5 = qr[5];
            unsigned v[5];
            v[0] = __funnelshift_r(w0, w1, bsh);
            v[1] = __funnelshift_r(w1, w2, bsh);
            v[2] = __funnelshift_r(w2, w3, bsh);
            v[3] = __funnelshift_r(w3, w4, bsh);
            v[4] = __funnelshift_r(w4, w5, bsh);
            int h[16];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                h[4 * k + 0] = dp4a_us(v[k], th, 0);
                h[4 * k + 1] = dp4a_us(__funnelshift_r(v[k], v[k + 1], 8), th, 0);
                h[4 * k + 2] = dp4a_us(__funnelshift_r(v[k], v[k + 1], 16), th, 0);
                h[4 * k + 3] = dp4a_us(__funnelshift_r(v[k], v[k + 1], 24), th, 0);
            }
            unsigned pr2[16]; /* pair (row r-1, row r) */
#pragma unroll
            for (int e = 0; e < 16; e++) {
                pr2[e] = __byte_perm((unsigned) hprev[e], (unsigned) h[e], 0x5410);
                hprev[e] = h[e];
            }
            if (r >= 3) { /* output row r-3: rows a,b = pair (r-3, r-2), rows c,d = pair (r-1, r) */
                unsigned pw[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int v0 = dp2a_lo_s16(pr2[4 * k + 0], t_cd, dp2a_lo_s16(pr0[4 * k + 0], t_ab, 128)) >> 8;
                    const int v1 = dp2a_lo_s16(pr2[4 * k + 1], t_cd, dp2a_lo_s16(pr0[4 * k + 1], t_ab, 128)) >> 8;
                    const int v2 = dp2a_lo_s16(pr2[4 * k + 2], t_cd, dp2a_lo_s16(pr0[4 * k + 2], t_ab, 128)) >> 8;
                    const int v3 = dp2a_lo_s16(pr2[4 * k + 3], t_cd, dp2a_lo_s16(pr0[4 * k + 3], t_ab, 128)) >> 8;
                    pw[k] = pack_u8x4(v0, v1, v2, v3);
                }
                if (s.intra) {
                    bmc_intra_row(pw, s.mv, means, s.lx, s.ly + r - 3, s.cw, s.ch);
                }
                bmc_store_row(P, a.mode, s.x + s.lx, s.y + s.ly + r - 3, s.n, s.vec16, pw);
            }
#pragma unroll
            for (int e = 0; e < 16; e++) {
                pr0[e] = pr1[e];
                pr1[e] = pr2[e];
            }
        }
    }
    bmc_edge_column(P, s, a.mode);
}

DSV_D void bmc_chroma_strip(const BmcArgs &a, const int c, const int u)
{
    BmcStrip s = {};
    if (!bmc_strip_setup<BMC_RC>(a, c, u, s)) {
        return;
    }
    const BmcPlane &P = a.pl[c];
    const unsigned means = s.intra ? a.means[(size_t) c * a.nbh * a.nbv + s.b] : 0u;
    /* weights for (p[0], p[1], p[rs], p[rs + 1]) */
    const unsigned wt = s.xh ? (s.yh ? 0x01010101u : 0x00000202u) : (s.yh ? 0x00020002u : 0x00000004u);
    const int rs = P.rstride;
    const uint8_t *src = P.ref + (ptrdiff_t) (s.py + s.ly) * rs + (s.px + s.lx);
    const unsigned bsh = ((unsigned) reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
    const unsigned *q = reinterpret_cast<const unsigned *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t) 3);
    const int qs = rs >> 2;

    unsigned top[16]; /* (p[e], p[e + 1]) of the row above in the two low bytes */
#pragma unroll
    for (int e = 0; e < 16; e++) {
        top[e] = 0;
    }
#pragma unroll
    for (int r = 0; r < BMC_RC + 1; r++) {
        if (r >= s.rows + 1) {
            break;
        }
        {
            const unsigned *qr = q + (ptrdiff_t) r * qs;
            const unsigned w0 = qr[0], w1 = qr[1], w2 = qr[2], w3 = qr[3], w4 = qr[4];
            unsigned v[5];
            v[0] = __funnelshift_r(w0, w1, bsh);
            v[1] = __funnelshift_r(w1, w2, bsh);
            v[2] = __funnelshift_r(w2, w3, bsh);
            v[3] = __funnelshift_r(w3, w4, bsh);
            v[4] = w4 >> bsh; /* only sample 16 is needed */
            unsigned cur[16];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                cur[4 * k + 0] = v[k];
                cur[4 * k + 1] = __funnelshift_r(v[k], v[k + 1], 8);
                cur[4 * k + 2] = __funnelshift_r(v[k], v[k + 1], 16);
                cur[4 * k + 3] = __funnelshift_r(v[k], v[k + 1], 24);
            }
            if (r >= 1) {
                unsigned pw[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int v0 = dp4a_us(__byte_perm(top[4 * k + 0], cur[4 * k + 0], 0x5410), wt, 2) >> 2;
                    const int v1 = dp4a_us(__byte_perm(top[4 * k + 1], cur[4 * k + 1], 0x5410), wt, 2) >> 2;
                    const int v2 = dp4a_us(__byte_perm(top[4 * k + 2], cur[4 * k + 2], 0x5410), wt, 2) >> 2;
                    const int v3 = dp4a_us(__byte_perm(top[4 * k + 3], cur[4 * k + 3], 0x5410), wt, 2) >> 2;
                    pw[k] = pack_u8x4(v0, v1, v2, v3);
                }
                if (s.intra) {
                    bmc_intra_row(pw, s.mv, means, s.lx, s.ly + r - 1, s.cw, s.ch);
                }
                bmc_store_row(P, a.mode, s.x + s.lx, s.y + s.ly + r - 1, s.n, s.vec16, pw);
            }
#pragma unroll
            for (int e = 0; e < 16; e++) {
                top[e] = cur[e];
            }
        }
    }
    bmc_edge_column(P, s, a.mode);
}

/* grid.x = [luma strips | U strips | V strips] in CTAs of BMC_THREADS strips, grid.y = lane */
__global__ void __launch_bounds__(BMC_THREADS, 2) bmc_kernel(const BmcArgs *args, int ctas_l, int ctas_c)
{
    const BmcArgs &a = args[blockIdx.y];
    int cta = blockIdx.x;
    if (cta < ctas_l) {
        bmc_luma_strip(a, cta * BMC_THREADS + threadIdx.x);
        return;
    }
    cta -= ctas_l;
    const int c = cta < ctas_c ? 1 : 2;
    if (c == 2) {
        cta -= ctas_c;
    }
    bmc_chroma_strip(a, c, cta * BMC_THREADS + threadIdx.x);
}

/* means of the intra blocks: one warp per block, planes in turn; a word per (plane, block) holds the four quadrant
 * means (avgval, bmc.c:176-190), or the block mean four times for an all-intra block */
__global__ void __launch_bounds__(BMC_THREADS) bmc_means_kernel(const BmcArgs *args)
{
    const BmcArgs &a = args[blockIdx.y];
    const int nblk = a.nbh * a.nbv;
    const int b = blockIdx.x * (BMC_THREADS / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= nblk) {
        return;
    }
    const DevMV mv = a.mv[b];
    if (mv.mode == 0) {
        return;
    }
    const int j = b / a.nbh, i = b - j * a.nbh;
    for (int c = 0; c < 3; c++) {
        const BmcPlane &P = a.pl[c];
        const int sh = c ? a.hs : 0, sv = c ? a.vs : 0;
        const int bw = a.blk_w >> sh, bh = a.blk_h >> sv;
        const int x = i * bw, y = j * bh;
        const int cw = (x + bw >= P.w) ? P.w - x : bw;
        const int ch = (y + bh >= P.h) ? P.h - y : bh;
        const bool whole = mv.submask == 15;
        const int sbw = whole ? cw : cw / 2, sbh = whole ? ch : ch / 2;
        unsigned word = 0;
        for (int qd = 0; qd < (whole ? 1 : 4); qd++) {
            unsigned acc = 0;
            if (whole || (mv.submask & (1 << qd))) {
                const uint8_t *r0 = P.ref + (ptrdiff_t) (y + (qd >> 1) * sbh) * P.rstride + x + (qd & 1) * sbw;
                for (int ly = 0; ly < sbh; ly++) {
                    for (int lx = lane; lx < sbw; lx += 32) {
                        acc += r0[(ptrdiff_t) ly * P.rstride + lx];
                    }
                }
            }
            acc = __reduce_add_sync(0xffffffffu, acc);
            const unsigned m = (sbw > 0 && sbh > 0) ? acc / (unsigned) (sbw * sbh) : 0u;
            word |= m << (8 * qd);
        }
        if (whole) {
            word *= 0x01010101u;
        }
        if (lane == 0 && cw > 0 && ch > 0) {
            a.means[(size_t) c * nblk + b] = word;
        }
    }
}

void bmc_fill_args(BmcArgs *a, const MotionGeom &g, const DevMV *d_mv, uint32_t *d_means, const DevFrame &ref, const DevFrame *pred,
                   const DevFrame &in, const DevFrame &out, int mode)
{
    for (int c = 0; c < 3; c++) {
        a->pl[c].ref = ref.p[c];
        a->pl[c].rstride = ref.stride[c];
        a->pl[c].pred = pred ? pred->p[c] : nullptr;
        a->pl[c].pstride = pred ? pred->stride[c] : 0;
        a->pl[c].in = in.p[c];
        a->pl[c].istride = in.stride[c];
        a->pl[c].out = out.p[c];
        a->pl[c].ostride = out.stride[c];
        a->pl[c].w = out.w[c];
        a->pl[c].h = out.h[c];
    }
    a->mv = d_mv;
    a->means = d_means;
    a->blk_w = g.blk_w;
    a->blk_h = g.blk_h;
    a->nbh = g.nbh;
    a->nbv = g.nbv;
    a->hs = g.hs;
    a->vs = g.vs;
    a->mode = mode;
}

static int bmc_ctas(int bw, int bh, int rows, int nblk)
{
    const long long units = (long long) ((bw + 15) >> 4) * ((bh + rows - 1) / rows) * nblk;
    return (int) ((units + BMC_THREADS - 1) / BMC_THREADS);
}

void bmc_launch(const BmcArgs *d_args, int n, const MotionGeom &g, cudaStream_t st)
{
    if (n <= 0) {
        return;
    }
    const int nblk = g.nbh * g.nbv;
    const int ctas_l = bmc_ctas(g.blk_w, g.blk_h, BMC_RL, nblk);
    const int ctas_c = bmc_ctas(g.blk_w >> g.hs, g.blk_h >> g.vs, BMC_RC, nblk);
    DSV_LAUNCH(bmc_means_kernel, dim3((nblk + BMC_THREADS / 32 - 1) / (BMC_THREADS / 32), n), dim3(BMC_THREADS), 0, st, d_args);
    KERNEL_CHECK();
    DSV_LAUNCH(bmc_kernel, dim3(ctas_l + 2 * ctas_c, n), dim3(BMC_THREADS), 0, st, d_args, ctas_l, ctas_c);
    KERNEL_CHECK();
}

} // namespace dsv
