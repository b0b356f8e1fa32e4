/*
 * bmc.cu -- half-pel block motion compensation, fused with residual formation (encoder) or
 * reconstruction (decoder).  Replaces compensate / hpelL / hpel / avgval / cpyzero / subf / addf /
 * dsv_sub_pred / dsv_add_pred (bmc.c:29-346).
 *
 * Streaming formulation: the unit of work is a strip of 8 samples x up to 24 rows of ONE motion block, one strip per thread, the strips of all blocks / planes / lanes flattened into one grid.  A thread
 * keeps its filter state in registers, shares nothing and meets no barrier; the rows it will need are fetched
 * BMC_D rows ahead by cp.async into thread-private shared-memory slots:
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

#ifndef BMC_THREADS
#define BMC_THREADS 256
#endif
#define BMC_W 8   /* samples per strip row */
#ifndef BMC_RL
#define BMC_RL 24 /* luma rows per strip */
#endif
#ifndef BMC_RC
#define BMC_RC 24 /* chroma rows per strip */
#endif
#ifndef BMC_D
#define BMC_D 6   /* rows in flight per thread (cp.async ring depth); a multiple of 3 */
#endif
#ifndef BMC_MINB
#define BMC_MINB 3 /* CTAs per SM the register budget is held to */
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

/*
 * Strips that are cut (fewer than 8 samples or unaligned: picture edge, block widths that are not multiples of 8) or
 * belong to an intra block take this plain per-sample path: the reference's formulas as written (bmc.c:58-174,
 * 255-298).  Out of line, so the streaming path below carries none of its state.
 */
template <bool LUMA> __device__ __noinline__ void bmc_slow_strip(const BmcPlane &P, const BmcStrip &s, int mode)
{
    const unsigned means = s.intra ? bmc_intra_means(P, s.mv, s.x, s.y, s.cw, s.ch) : 0u;
    const bool whole = s.mv.submask == 15;
    const int sbw = s.cw / 2, sbh = s.ch / 2;
    const int rs = P.rstride;
    for (int t = 0; t < s.rows; t++) {
        const int ly = s.ly + t, gy = s.y + ly;
        for (int e = 0; e < s.n; e++) {
            const int lx = s.lx + e, gx = s.x + lx;
            const uint8_t *p = P.ref + (ptrdiff_t) (s.py + ly) * rs + s.px + lx;
            int pv;
            if (LUMA) {
                if (!s.xh && !s.yh) {
                    pv = p[0];
                } else if (!s.xh) {
                    pv = clamp_u8((9 * (p[0] + p[rs]) - (p[-rs] + p[2 * rs]) + 8) >> 4);
                } else if (!s.yh) {
                    pv = clamp_u8((9 * (p[0] + p[1]) - (p[-1] + p[2]) + 8) >> 4);
                } else {
                    int h[4];
                    for (int k = 0; k < 4; k++) {
                        const uint8_t *q = p + (ptrdiff_t) (k - 1) * rs;
                        h[k] = 9 * (q[0] + q[1]) - (q[-1] + q[2]);
                    }
                    pv = clamp_u8((9 * (h[1] + h[2]) - (h[0] + h[3]) + 128) >> 8);
                }
            } else {
                if (!s.xh && !s.yh) {
                    pv = p[0];
                } else if (!s.xh) {
                    pv = (p[0] + p[rs] + 1) >> 1;
                } else if (!s.yh) {
                    pv = (p[0] + p[1] + 1) >> 1;
                } else {
                    pv = (p[0] + p[1] + p[rs] + p[rs + 1] + 2) >> 2;
                }
            }
            if (s.intra) { /* pv is the co-located sample here (zero vector) */
                if (whole) {
                    pv = (int) (means & 0xffu);
                } else if (lx >= 2 * sbw || ly >= 2 * sbh) {
                    pv = 0; /* odd edge blocks: the quadrants do not cover the last row / column (zeroed frame) */
                } else {
                    const int qd = (lx >= sbw ? 1 : 0) | (ly >= sbh ? 2 : 0);
                    if (s.mv.submask & (1 << qd)) {
                        pv = (int) ((means >> (8 * qd)) & 0xffu);
                    }
                }
            }
            const int c = P.in[(size_t) gy * P.istride + gx];
            P.out[(size_t) gy * P.ostride + gx] = mode == 1 ? clamp_u8(c - pv + 128) : clamp_u8(pv + c - 128);
            if (P.pred) {
                P.pred[(size_t) gy * P.pstride + gx] = (uint8_t) pv;
            }
        }
    }
}

/*
 * Rows in flight.  Reference rows (bytes p[-1..10] of a luma strip row = 4 aligned words, 3 for chroma) and the
 * co-located rows of `in` (8 bytes) are fetched BMC_D rows ahead with cp.async into thread-private slots of a
 * shared-memory ring -- reference words as [slot][word][thread], `in` rows as [slot][thread] word pairs, so both the
 * asynchronous writes and the reads are conflict-free.  Nothing in flight holds a register.  The row loop runs
 * BMC_D row bodies per trip (ring slots and the period-3 rotation of the pair registers are then compile-time
 * names; BMC_D is a multiple of 3) plus single-row trips for the remainder.
 * Group g = reference row g + `in` row g - LAG (LAG = rows between a reference row arriving and the output row it
 * completes); one group is committed per row body, empty past the end.
 */
struct BmcFeed {
    unsigned *ref;      /* this thread's column of [BMC_D][4][BMC_THREADS] */
    unsigned *cur;      /* this thread's pair in [BMC_D][BMC_THREADS][2] */
    const unsigned *q;  /* next reference row to issue (aligned words) */
    const uint8_t *in;  /* next row of `in` to issue */
    int qs, istride;
    int nref, rows;     /* reference rows / rows of `in` this strip needs: later groups are empty */
};

template <int NW, int LAG> DSV_D void bmc_issue(BmcFeed &f, int slot, int g)
{
    if (g < f.nref) {
        unsigned *d = f.ref + slot * 4 * BMC_THREADS;
#pragma unroll
        for (int k = 0; k < NW; k++) {
            cp_async4(d + k * BMC_THREADS, f.q + k);
        }
        f.q += f.qs;
        if (g >= LAG) { /* g - LAG < rows follows from g < nref = rows + LAG */
            cp_async8(f.cur + slot * 2 * BMC_THREADS, f.in);
            f.in += f.istride;
        }
    }
    cp_async_commit();
}

/* store one output row of the 8-byte path and step the pointers */
template <int MODE> DSV_D void bmc_emit(uint8_t *&outp, uint8_t *&predp, int ostride, int pstride, uint2 cur, const unsigned pw[2])
{
    *reinterpret_cast<uint2 *>(outp) = make_uint2(bmc_combine4(cur.x, pw[0], MODE), bmc_combine4(cur.y, pw[1], MODE));
    outp += ostride;
    if (MODE == 1) { /* the encoder keeps the prediction */
        *reinterpret_cast<uint2 *>(predp) = make_uint2(pw[0], pw[1]);
        predp += pstride;
    }
}

struct BmcOut {
    uint8_t *outp, *predp;
    int ostride, pstride, rows;
};

/* one reference row of a luma strip: wait for it, refill its ring slot, H pass, pair it with the row above, and
 * (from the fourth row on) finish the output row three rows up.  rd: pair (r - 3, r - 2); wr receives (r - 1, r). */
template <int MODE>
DSV_D void bmc_luma_row(BmcFeed &f, BmcOut &o, int slot, int r, unsigned bsh, unsigned th, unsigned t_ab, unsigned t_cd,
                        int (&hprev)[BMC_W], const unsigned (&rd)[BMC_W], unsigned (&wr)[BMC_W])
{
    cp_async_wait<BMC_D - 1>();
    const unsigned *sl = f.ref + slot * 4 * BMC_THREADS;
    const unsigned w0 = sl[0], w1 = sl[BMC_THREADS], w2 = sl[2 * BMC_THREADS], w3 = sl[3 * BMC_THREADS];
    const uint2 c = *reinterpret_cast<const uint2 *>(f.cur + slot * 2 * BMC_THREADS);
    bmc_issue<4, 3>(f, slot, r + BMC_D);
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
    unsigned pnew[BMC_W];
#pragma unroll
    for (int e = 0; e < BMC_W; e++) {
        pnew[e] = __byte_perm((unsigned) hprev[e], (unsigned) h[e], 0x5410);
        hprev[e] = h[e];
    }
    if (r >= 3 && r - 3 < o.rows) {
        unsigned pw[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int v0 = dp2a_lo_s16(pnew[4 * k + 0], t_cd, dp2a_lo_s16(rd[4 * k + 0], t_ab, 128)) >> 8;
            const int v1 = dp2a_lo_s16(pnew[4 * k + 1], t_cd, dp2a_lo_s16(rd[4 * k + 1], t_ab, 128)) >> 8;
            const int v2 = dp2a_lo_s16(pnew[4 * k + 2], t_cd, dp2a_lo_s16(rd[4 * k + 2], t_ab, 128)) >> 8;
            const int v3 = dp2a_lo_s16(pnew[4 * k + 3], t_cd, dp2a_lo_s16(rd[4 * k + 3], t_ab, 128)) >> 8;
            pw[k] = pack_u8x4(v0, v1, v2, v3);
        }
        bmc_emit<MODE>(o.outp, o.predp, o.ostride, o.pstride, c, pw);
    }
#pragma unroll
    for (int e = 0; e < BMC_W; e++) {
        wr[e] = pnew[e];
    }
}

template <int MODE> DSV_D void bmc_luma_strip(const BmcArgs &a, const int u, unsigned *smem)
{
    const BmcPlane P = a.pl[0];
    BmcStrip s = {};
    if (!bmc_strip_setup<BMC_RL>(a, P, 0, u, s)) {
        return;
    }
    bmc_edge_column(P, s, MODE);
    if (s.intra || !s.vec8) {
        const BmcStrip sc = s; /* only the copy's address escapes: s itself stays in registers */
        const BmcPlane Pc = P;
        bmc_slow_strip<true>(Pc, sc, MODE);
        return;
    }
    /* taps: bytes (p[-1], p[0], p[1], p[2]); pairs (row a, row b) and (row c, row d) */
    const unsigned th = s.xh ? 0xff0909ffu : 0x00001000u;
    const unsigned t_ab = s.yh ? 0x09ffu : 0x1000u, t_cd = s.yh ? 0xff09u : 0x0000u;
    /* first byte needed: one row above and one sample left of the strip's reference position */
    const uint8_t *src = P.ref + (ptrdiff_t) (s.py + s.ly - 1) * P.rstride + (s.px + s.lx - 1);
    const unsigned bsh = ((unsigned) reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
    BmcOut o;
    o.rows = s.rows;
    o.outp = P.out + (size_t) (s.y + s.ly) * P.ostride + s.x + s.lx;
    o.predp = MODE == 1 ? P.pred + (size_t) (s.y + s.ly) * P.pstride + s.x + s.lx : nullptr;
    o.ostride = P.ostride;
    o.pstride = P.pstride;
    BmcFeed f;
    f.ref = smem + threadIdx.x;
    f.cur = smem + BMC_D * 4 * BMC_THREADS + 2 * threadIdx.x;
    f.q = reinterpret_cast<const unsigned *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t) 3);
    f.qs = P.rstride >> 2; /* strides are multiples of 16 */
    f.in = P.in + (size_t) (s.y + s.ly) * P.istride + s.x + s.lx;
    f.istride = P.istride;
    f.nref = s.rows + 3;
    f.rows = s.rows;
    const int nref = f.nref;

#pragma unroll
    for (int g = 0; g < BMC_D; g++) {
        bmc_issue<4, 3>(f, g, g);
    }
    int hprev[BMC_W];
    unsigned p0[BMC_W], p1[BMC_W], p2[BMC_W]; /* vertical pairs of 16-bit H results: pair j = rows (j, j + 1) in p<j % 3> */
#pragma unroll
    for (int e = 0; e < BMC_W; e++) {
        hprev[e] = 0;
        p0[e] = p1[e] = p2[e] = 0;
    }
    int r = 0;
#pragma unroll 1
    for (; r + BMC_D <= nref; r += BMC_D) { /* r % 3 == 0: row r reads pair r - 3 in p0 and writes pair r - 1 into p2, ... */
#pragma unroll
        for (int jj = 0; jj < BMC_D; jj += 3) {
            bmc_luma_row<MODE>(f, o, jj + 0, r + jj + 0, bsh, th, t_ab, t_cd, hprev, p0, p2);
            bmc_luma_row<MODE>(f, o, jj + 1, r + jj + 1, bsh, th, t_ab, t_cd, hprev, p1, p0);
            bmc_luma_row<MODE>(f, o, jj + 2, r + jj + 2, bsh, th, t_ab, t_cd, hprev, p2, p1);
        }
    }
#pragma unroll 1
    for (int slot = 0; r < nref; r++, slot++) { /* remainder: one row per trip, the pair registers rotate by moves */
        bmc_luma_row<MODE>(f, o, slot, r, bsh, th, t_ab, t_cd, hprev, p0, p2);
#pragma unroll
        for (int e = 0; e < BMC_W; e++) {
            const unsigned t = p0[e];
            p0[e] = p1[e];
            p1[e] = p2[e];
            p2[e] = t;
        }
    }
}

/* one reference row of a chroma strip; top: (p[e], p[e + 1]) of the row above in the two low bytes */
template <int MODE> DSV_D void bmc_chroma_row(BmcFeed &f, BmcOut &o, int slot, int r, unsigned bsh, unsigned wt, unsigned (&top)[BMC_W])
{
    cp_async_wait<BMC_D - 1>();
    const unsigned *sl = f.ref + slot * 4 * BMC_THREADS;
    const unsigned w0 = sl[0], w1 = sl[BMC_THREADS], w2 = sl[2 * BMC_THREADS];
    const uint2 cc = *reinterpret_cast<const uint2 *>(f.cur + slot * 2 * BMC_THREADS);
    bmc_issue<3, 1>(f, slot, r + BMC_D);
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
    if (r >= 1 && r - 1 < o.rows) {
        unsigned pw[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int v0 = dp4a_us(__byte_perm(top[4 * k + 0], row[4 * k + 0], 0x5410), wt, 2) >> 2;
            const int v1 = dp4a_us(__byte_perm(top[4 * k + 1], row[4 * k + 1], 0x5410), wt, 2) >> 2;
            const int v2 = dp4a_us(__byte_perm(top[4 * k + 2], row[4 * k + 2], 0x5410), wt, 2) >> 2;
            const int v3 = dp4a_us(__byte_perm(top[4 * k + 3], row[4 * k + 3], 0x5410), wt, 2) >> 2;
            pw[k] = pack_u8x4(v0, v1, v2, v3);
        }
        bmc_emit<MODE>(o.outp, o.predp, o.ostride, o.pstride, cc, pw);
    }
#pragma unroll
    for (int e = 0; e < BMC_W; e++) {
        top[e] = row[e];
    }
}

template <int MODE> DSV_D void bmc_chroma_strip(const BmcArgs &a, const int c, const int u, unsigned *smem)
{
    const BmcPlane P = a.pl[c];
    BmcStrip s = {};
    if (!bmc_strip_setup<BMC_RC>(a, P, c, u, s)) {
        return;
    }
    bmc_edge_column(P, s, MODE);
    if (s.intra || !s.vec8) {
        const BmcStrip sc = s;
        const BmcPlane Pc = P;
        bmc_slow_strip<false>(Pc, sc, MODE);
        return;
    }
    /* weights for (p[0], p[1], p[rs], p[rs + 1]) */
    const unsigned wt = s.xh ? (s.yh ? 0x01010101u : 0x00000202u) : (s.yh ? 0x00020002u : 0x00000004u);
    const uint8_t *src = P.ref + (ptrdiff_t) (s.py + s.ly) * P.rstride + (s.px + s.lx);
    const unsigned bsh = ((unsigned) reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
    BmcOut o;
    o.rows = s.rows;
    o.outp = P.out + (size_t) (s.y + s.ly) * P.ostride + s.x + s.lx;
    o.predp = MODE == 1 ? P.pred + (size_t) (s.y + s.ly) * P.pstride + s.x + s.lx : nullptr;
    o.ostride = P.ostride;
    o.pstride = P.pstride;
    BmcFeed f;
    f.ref = smem + threadIdx.x;
    f.cur = smem + BMC_D * 4 * BMC_THREADS + 2 * threadIdx.x;
    f.q = reinterpret_cast<const unsigned *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t) 3);
    f.qs = P.rstride >> 2;
    f.in = P.in + (size_t) (s.y + s.ly) * P.istride + s.x + s.lx;
    f.istride = P.istride;
    f.nref = s.rows + 1;
    f.rows = s.rows;
    const int nref = f.nref;

#pragma unroll
    for (int g = 0; g < BMC_D; g++) {
        bmc_issue<3, 1>(f, g, g);
    }
    unsigned top[BMC_W];
#pragma unroll
    for (int e = 0; e < BMC_W; e++) {
        top[e] = 0;
    }
    int r = 0;
#pragma unroll 1
    for (; r + BMC_D <= nref; r += BMC_D) {
#pragma unroll
        for (int jj = 0; jj < BMC_D; jj++) {
            bmc_chroma_row<MODE>(f, o, jj, r + jj, bsh, wt, top);
        }
    }
#pragma unroll 1
    for (int slot = 0; r < nref; r++, slot++) {
        bmc_chroma_row<MODE>(f, o, slot, r, bsh, wt, top);
    }
}

template <int MODE> DSV_D void bmc_strips(const BmcArgs &a, int ctas_l, int ctas_c, unsigned *smem)
{
    int cta = blockIdx.x;
    if (cta < ctas_l) {
        bmc_luma_strip<MODE>(a, cta * BMC_THREADS + threadIdx.x, smem);
        return;
    }
    cta -= ctas_l;
    const int c = cta < ctas_c ? 1 : 2;
    if (c == 2) {
        cta -= ctas_c;
    }
    bmc_chroma_strip<MODE>(a, c, cta * BMC_THREADS + threadIdx.x, smem);
}

/* grid.x = [luma strips | U strips | V strips] in CTAs of BMC_THREADS strips, grid.y = lane */
__global__ void __launch_bounds__(BMC_THREADS, BMC_MINB) bmc_kernel(const BmcArgs *args, int ctas_l, int ctas_c)
{
    DSV_DYN_SMEM(unsigned, smem); /* BMC_D x (4 + 2) x BMC_THREADS words */
    const BmcArgs &a = args[blockIdx.y];
    if (a.mode == 1) { /* one mode per launch: uniform */
        bmc_strips<1>(a, ctas_l, ctas_c, smem);
    } else {
        bmc_strips<2>(a, ctas_l, ctas_c, smem);
    }
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
    const size_t smem = (size_t) BMC_D * 6 * BMC_THREADS * sizeof(unsigned);
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(bmc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr_set = true;
    }
    DSV_LAUNCH(bmc_kernel, dim3(ctas_l + 2 * ctas_c, n), dim3(BMC_THREADS), smem, st, d_args, ctas_l, ctas_c);
    KERNEL_CHECK();
}

} // namespace dsv
