/*
 * bmc.cu -- half-pel block motion compensation, fused with residual formation (encoder) or
 * reconstruction (decoder).  Replaces compensate / hpelL / hpel / avgval / cpyzero / subf / addf /
 * dsv_sub_pred / dsv_add_pred (bmc.c:29-346).
 *
 * Streaming formulation: the unit of work is a strip of 8 samples x R rows of ONE motion block (R = 16 luma,
 * 8 chroma), one strip per thread, the strips of all blocks / planes / lanes flattened into one grid.  A thread
 * keeps everything in registers -- no shared memory, no barriers -- and keeps BMC_PF reference rows plus three
 * rows of the current picture in flight (explicit prefetch rings):
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
 *           overwrite the flagged quadrants with the block / quadrant means, which the strip's thread sums itself
 *           (a separate pass over all blocks measured 13-45 us per launch to find, usually, nothing).
 *
 * The prediction never makes a round trip through HBM on the decoder side (mode 2: io = clamp(pred + io - 128));
 * the encoder keeps it (mode 1) because the closed-loop reconstruction adds it back in the inverse transform's
 * store (sbt_inv.cu).  Plane rows are 16-byte aligned, so whole strip rows move as 8-byte loads / stores; strips cut
 * by the picture edge or by block widths that are not multiples of 8 fall back to bytes.
 *
 * Reads outside the picture go through the 64-sample replicated border exactly like the reference
 * (position clamp bmc.c:221-249); filter taps that step one sample past the border see the same
 * neighbouring bytes because DevFrame keeps the reference's strides and plane order.  Taps with weight zero
 * read (and ignore) bytes the reference does not touch; they stay inside the frame allocation's guard bands.
 */
#include "motion.cuh"

namespace dsv {

#define BMC_THREADS 256
#define BMC_W 8   /* samples per strip row */
#define BMC_RL 16 /* luma rows per strip */
#define BMC_RC 8  /* chroma rows per strip */
#ifndef BMC_PF
#define BMC_PF 4  /* reference rows in flight per thread */
#endif
#ifndef BMC_MINB
#define BMC_MINB 3 /* CTAs per SM the register budget is held to: 369 us per 64 HD pictures at 2 (96 registers), 322 at 3 (80) */
#endif

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

/* four packed samples: mode 1 clamp(cur - pred + 128), mode 2 clamp(pred + cur - 128) (subf / addf, bmc.c:29-57).
 * With both operands moved to signed bytes (x ^ 0x80 == x - 128) these are the signed saturating byte
 * subtraction / addition, moved back: a dozen logic instructions per word instead of unpack - add - repack. */
DSV_D unsigned bmc_combine4(unsigned cur, unsigned pw, int mode)
{
    const unsigned a = cur ^ 0x80808080u, b = pw ^ 0x80808080u;
    return (mode == 1 ? __vsubss4(a, b) : __vaddss4(a, b)) ^ 0x80808080u;
}

/* one row of a strip: 8 predicted samples in pw[2] (sample 0 in the lowest byte of pw[0]); n = valid samples;
 * cur = the co-located samples of `in` when vec8 (fetched ahead by the caller) */
DSV_D void bmc_store_row(const BmcPlane &P, int mode, int gx, int gy, int n, bool vec8, uint2 cur, const unsigned pw[2])
{
    const size_t oo = (size_t) gy * P.ostride + gx;
    if (vec8) { /* n == 8 and every row of the three frames 8-byte aligned at gx */
        *reinterpret_cast<uint2 *>(P.out + oo) = make_uint2(bmc_combine4(cur.x, pw[0], mode), bmc_combine4(cur.y, pw[1], mode));
        if (P.pred) {
            *reinterpret_cast<uint2 *>(P.pred + (size_t) gy * P.pstride + gx) = make_uint2(pw[0], pw[1]);
        }
        return;
    }
    const size_t io = (size_t) gy * P.istride + gx;
#pragma unroll
    for (int k = 0; k < 2; k++) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (4 * k + e < n) {
                const int pv = byte_of(pw[k], e);
                const int c = P.in[io + 4 * k + e];
                P.out[oo + 4 * k + e] = mode == 1 ? clamp_u8(c - pv + 128) : clamp_u8(pv + c - 128);
                if (P.pred) {
                    P.pred[(size_t) gy * P.pstride + gx + 4 * k + e] = (uint8_t) pv;
                }
            }
        }
    }
}

/* intra blocks (bmc.c:255-298): the flagged quadrants become their mean, the others keep the co-located reference
 * already in pw; odd edge blocks: the quadrants do not cover the last row / column (zeroed frame in the reference) */
DSV_D void bmc_intra_row(unsigned pw[2], const DevMV &mv, unsigned means, int lx, int ly, int cw, int ch)
{
    const bool whole = mv.submask == 15;
    const int sbw = cw / 2, sbh = ch / 2;
#pragma unroll
    for (int k = 0; k < 2; k++) {
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

/* block / quadrant means of an intra block (avgval, bmc.c:176-190), computed by the strip's own thread: a word holds
 * the four quadrant means, or the block mean four times for an all-intra block.  Intra blocks are rare (a picture with
 * more than intra_pct of them is coded as an I picture), so every strip of the block redoing the sums costs less than
 * a separate pass over all blocks would. */
DSV_D unsigned bmc_rect_mean(const uint8_t *r0, int rstride, int sw, int sh)
{
    if (sw <= 0 || sh <= 0) {
        return 0u;
    }
    unsigned acc = 0;
    for (int ly = 0; ly < sh; ly++) {
        const uint8_t *p = r0 + (ptrdiff_t) ly * rstride;
        int lx = 0;
        for (; lx + 4 <= sw; lx += 4) {
            acc = (unsigned) dp4a_us(ld4u(p + lx), 0x01010101u, (int) acc);
        }
        if (lx < sw) {
            acc = (unsigned) dp4a_us(ld4u(p + lx) & (0xffffffffu >> (8 * (4 - (sw - lx)))), 0x01010101u, (int) acc);
        }
    }
    return acc / (unsigned) (sw * sh);
}
DSV_D unsigned bmc_intra_means(const BmcPlane &P, const DevMV &mv, int x, int y, int cw, int ch)
{
    if (mv.submask == 15) {
        return bmc_rect_mean(P.ref + (ptrdiff_t) y * P.rstride + x, P.rstride, cw, ch) * 0x01010101u;
    }
    const int sbw = cw / 2, sbh = ch / 2;
    unsigned word = 0;
    for (int qd = 0; qd < 4; qd++) {
        if (mv.submask & (1 << qd)) {
            word |= bmc_rect_mean(P.ref + (ptrdiff_t) (y + (qd >> 1) * sbh) * P.rstride + x + (qd & 1) * sbw, P.rstride, sbw, sbh) << (8 * qd);
        }
    }
    return word;
}

struct BmcStrip {
    int b;          /* block index */
    int x, y;       /* block origin in the plane */
    int cw, ch;     /* clipped block size */
    int lx, ly;     /* strip origin inside the block */
    int n, rows;    /* valid samples per row / rows */
    int px, py;     /* clamped integer reference position of the block */
    int xh, yh;     /* half-pel flags */
    bool intra, vec8;
    DevMV mv;
};

/* NOTE: the argument record lives in global memory and the kernel stores through pointers the compiler cannot
 * tell apart from it, so every field is copied to a local ONCE (P by value, mode, means pointer); reading a.x
 * in the row loop would be re-fetched from memory after every store. */
template <int R> DSV_D bool bmc_strip_setup(const BmcArgs &a, const BmcPlane &P, const int c, const int u, BmcStrip &s)
{
    const int sh = c ? a.hs : 0, sv = c ? a.vs : 0;
    const int bw = a.blk_w >> sh, bh = a.blk_h >> sv;
    const int segs = (bw + BMC_W - 1) / BMC_W, rgs = (bh + R - 1) / R, upb = segs * rgs;
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
    s.lx = BMC_W * seg;
    s.ly = R * rg;
    if (s.lx >= s.cw || s.ly >= s.ch) {
        return false;
    }
    s.n = imin(BMC_W, s.cw - s.lx);
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
    s.vec8 = s.n == BMC_W && (al & 7) == 0;
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

/* the co-located row of `in` for the 8-byte path (rows past the strip repeat its last row: the load stays valid and
 * unconditional, so it can be issued rows ahead of its use) */
DSV_D uint2 bmc_fetch_in(const BmcPlane &P, const BmcStrip &s, int t)
{
    if (!s.vec8) {
        return make_uint2(0u, 0u);
    }
    const int gy = s.y + s.ly + imin(t, s.rows - 1);
    return *reinterpret_cast<const uint2 *>(P.in + (size_t) gy * P.istride + s.x + s.lx);
}

DSV_D void bmc_luma_strip(const BmcArgs &a, const int u)
{
    const BmcPlane P = a.pl[0];
    const int mode = a.mode;
    BmcStrip s = {};
    if (!bmc_strip_setup<BMC_RL>(a, P, 0, u, s)) {
        return;
    }
    const unsigned means = s.intra ? bmc_intra_means(P, s.mv, s.x, s.y, s.cw, s.ch) : 0u;
    /* taps: bytes (p[-1], p[0], p[1], p[2]); pairs (row a, row b) and (row c, row d) */
    const unsigned th = s.xh ? 0xff0909ffu : 0x00001000u;
    const unsigned t_ab = s.yh ? 0x09ffu : 0x1000u, t_cd = s.yh ? 0xff09u : 0x0000u;
    const int rs = P.rstride;
    /* first byte needed: one row above and one sample left of the strip's reference position */
    const uint8_t *src = P.ref + (ptrdiff_t) (s.py + s.ly - 1) * rs + (s.px + s.lx - 1);
    const unsigned bsh = ((unsigned) reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
    const unsigned *q = reinterpret_cast<const unsigned *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t) 3);
    const int qs = rs >> 2; /* strides are multiples of 16 */
    const int last = s.rows + 2; /* last reference row this strip needs; later rows repeat it (loads stay in bounds) */

    /* reference rows travel through a ring of BMC_PF raw rows (bytes p[-1..10] = 4 aligned words), the rows of `in`
     * through a ring of 3: enough loads in flight per thread to cover the memory latency at 2-3 CTAs per SM */
    unsigned raw[BMC_PF][4];
#pragma unroll
    for (int d = 0; d < BMC_PF; d++) {
        const unsigned *qr = q + (ptrdiff_t) imin(d, last) * qs;
        raw[d][0] = qr[0], raw[d][1] = qr[1], raw[d][2] = qr[2], raw[d][3] = qr[3];
    }
    uint2 cur[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        cur[d] = bmc_fetch_in(P, s, d);
    }
    int hprev[BMC_W];
    unsigned pr0[BMC_W], pr1[BMC_W]; /* vertical pairs (row r-3, r-2) and (row r-2, r-1) of 16-bit H results */
#pragma unroll
    for (int e = 0; e < BMC_W; e++) {
        hprev[e] = 0;
        pr0[e] = pr1[e] = 0;
    }
#pragma unroll
    for (int r = 0; r < BMC_RL + 3; r++) {
        const unsigned w0 = raw[r % BMC_PF][0], w1 = raw[r % BMC_PF][1], w2 = raw[r % BMC_PF][2], w3 = raw[r % BMC_PF][3];
        if (r + BMC_PF < BMC_RL + 3) {
            const unsigned *qr = q + (ptrdiff_t) imin(r + BMC_PF, last) * qs;
            raw[r % BMC_PF][0] = qr[0], raw[r % BMC_PF][1] = qr[1], raw[r % BMC_PF][2] = qr[2], raw[r % BMC_PF][3] = qr[3];
        }
        unsigned v[3];
        v[0] = __funnelshift_r(w0, w1, bsh);
        v[1] = __funnelshift_r(w1, w2, bsh);
        v[2] = __funnelshift_r(w2, w3, bsh);
        int h[BMC_W];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            h[4 * k + 0] = dp4a_us(v[k], th, 0);
            h[4 * k + 1] = dp4a_us(__funnelshift_r(v[k], v[k + 1], 8), th, 0);
            h[4 * k + 2] = dp4a_us(__funnelshift_r(v[k], v[k + 1], 16), th, 0);
            h[4 * k + 3] = dp4a_us(__funnelshift_r(v[k], v[k + 1], 24), th, 0);
        }
        unsigned pr2[BMC_W]; /* pair (row r-1, row r) */
#pragma unroll
        for (int e = 0; e < BMC_W; e++) {
            pr2[e] = __byte_perm((unsigned) hprev[e], (unsigned) h[e], 0x5410);
            hprev[e] = h[e];
        }
        if (r >= 3) { /* output row r-3: rows a,b = pair (r-3, r-2), rows c,d = pair (r-1, r) */
            const int t = r - 3;
            unsigned pw[2];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int v0 = dp2a_lo_s16(pr2[4 * k + 0], t_cd, dp2a_lo_s16(pr0[4 * k + 0], t_ab, 128)) >> 8;
                const int v1 = dp2a_lo_s16(pr2[4 * k + 1], t_cd, dp2a_lo_s16(pr0[4 * k + 1], t_ab, 128)) >> 8;
                const int v2 = dp2a_lo_s16(pr2[4 * k + 2], t_cd, dp2a_lo_s16(pr0[4 * k + 2], t_ab, 128)) >> 8;
                const int v3 = dp2a_lo_s16(pr2[4 * k + 3], t_cd, dp2a_lo_s16(pr0[4 * k + 3], t_ab, 128)) >> 8;
                pw[k] = pack_u8x4(v0, v1, v2, v3);
            }
            const uint2 c = cur[t % 3];
            if (t + 3 < BMC_RL) {
                cur[t % 3] = bmc_fetch_in(P, s, t + 3);
            }
            if (t < s.rows) {
                if (s.intra) {
                    bmc_intra_row(pw, s.mv, means, s.lx, s.ly + t, s.cw, s.ch);
                }
                bmc_store_row(P, mode, s.x + s.lx, s.y + s.ly + t, s.n, s.vec8, c, pw);
            }
        }
#pragma unroll
        for (int e = 0; e < BMC_W; e++) {
            pr0[e] = pr1[e];
            pr1[e] = pr2[e];
        }
    }
    bmc_edge_column(P, s, mode);
}

DSV_D void bmc_chroma_strip(const BmcArgs &a, const int c, const int u)
{
    const BmcPlane P = a.pl[c];
    const int mode = a.mode;
    BmcStrip s = {};
    if (!bmc_strip_setup<BMC_RC>(a, P, c, u, s)) {
        return;
    }
    const unsigned means = s.intra ? bmc_intra_means(P, s.mv, s.x, s.y, s.cw, s.ch) : 0u;
    /* weights for (p[0], p[1], p[rs], p[rs + 1]) */
    const unsigned wt = s.xh ? (s.yh ? 0x01010101u : 0x00000202u) : (s.yh ? 0x00020002u : 0x00000004u);
    const int rs = P.rstride;
    const uint8_t *src = P.ref + (ptrdiff_t) (s.py + s.ly) * rs + (s.px + s.lx);
    const unsigned bsh = ((unsigned) reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
    const unsigned *q = reinterpret_cast<const unsigned *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t) 3);
    const int qs = rs >> 2;
    const int last = s.rows;

    unsigned raw[BMC_PF][3]; /* bytes p[0..8] */
#pragma unroll
    for (int d = 0; d < BMC_PF; d++) {
        const unsigned *qr = q + (ptrdiff_t) imin(d, last) * qs;
        raw[d][0] = qr[0], raw[d][1] = qr[1], raw[d][2] = qr[2];
    }
    uint2 cur[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        cur[d] = bmc_fetch_in(P, s, d);
    }
    unsigned top[BMC_W]; /* (p[e], p[e + 1]) of the row above in the two low bytes */
#pragma unroll
    for (int e = 0; e < BMC_W; e++) {
        top[e] = 0;
    }
#pragma unroll
    for (int r = 0; r < BMC_RC + 1; r++) {
        const unsigned w0 = raw[r % BMC_PF][0], w1 = raw[r % BMC_PF][1], w2 = raw[r % BMC_PF][2];
        if (r + BMC_PF < BMC_RC + 1) {
            const unsigned *qr = q + (ptrdiff_t) imin(r + BMC_PF, last) * qs;
            raw[r % BMC_PF][0] = qr[0], raw[r % BMC_PF][1] = qr[1], raw[r % BMC_PF][2] = qr[2];
        }
        unsigned v[3];
        v[0] = __funnelshift_r(w0, w1, bsh);
        v[1] = __funnelshift_r(w1, w2, bsh);
        v[2] = w2 >> bsh; /* only sample 8 is needed */
        unsigned row[BMC_W];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            row[4 * k + 0] = v[k];
            row[4 * k + 1] = __funnelshift_r(v[k], v[k + 1], 8);
            row[4 * k + 2] = __funnelshift_r(v[k], v[k + 1], 16);
            row[4 * k + 3] = __funnelshift_r(v[k], v[k + 1], 24);
        }
        if (r >= 1) {
            const int t = r - 1;
            unsigned pw[2];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int v0 = dp4a_us(__byte_perm(top[4 * k + 0], row[4 * k + 0], 0x5410), wt, 2) >> 2;
                const int v1 = dp4a_us(__byte_perm(top[4 * k + 1], row[4 * k + 1], 0x5410), wt, 2) >> 2;
                const int v2 = dp4a_us(__byte_perm(top[4 * k + 2], row[4 * k + 2], 0x5410), wt, 2) >> 2;
                const int v3 = dp4a_us(__byte_perm(top[4 * k + 3], row[4 * k + 3], 0x5410), wt, 2) >> 2;
                pw[k] = pack_u8x4(v0, v1, v2, v3);
            }
            const uint2 cc = cur[t % 3];
            if (t + 3 < BMC_RC) {
                cur[t % 3] = bmc_fetch_in(P, s, t + 3);
            }
            if (t < s.rows) {
                if (s.intra) {
                    bmc_intra_row(pw, s.mv, means, s.lx, s.ly + t, s.cw, s.ch);
                }
                bmc_store_row(P, mode, s.x + s.lx, s.y + s.ly + t, s.n, s.vec8, cc, pw);
            }
        }
#pragma unroll
        for (int e = 0; e < BMC_W; e++) {
            top[e] = row[e];
        }
    }
    bmc_edge_column(P, s, mode);
}

/* grid.x = [luma strips | U strips | V strips] in CTAs of BMC_THREADS strips, grid.y = lane */
__global__ void __launch_bounds__(BMC_THREADS, BMC_MINB) bmc_kernel(const BmcArgs *args, int ctas_l, int ctas_c)
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

void bmc_fill_args(BmcArgs *a, const MotionGeom &g, const DevMV *d_mv, const DevFrame &ref, const DevFrame *pred,
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
    const long long units = (long long) ((bw + BMC_W - 1) / BMC_W) * ((bh + rows - 1) / rows) * nblk;
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
    DSV_LAUNCH(bmc_kernel, dim3(ctas_l + 2 * ctas_c, n), dim3(BMC_THREADS), 0, st, d_args, ctas_l, ctas_c);
    KERNEL_CHECK();
}

} // namespace dsv
