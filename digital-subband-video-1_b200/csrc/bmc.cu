/*
 * bmc.cu -- half-pel block motion compensation, fused with residual formation (encoder) or
 * reconstruction (decoder).  Replaces compensate / hpelL / hpel / avgval / cpyzero / subf / addf /
 * dsv_sub_pred / dsv_add_pred / dsv_frame_add (bmc.c:29-346).
 *
 *   bmc_kernel        one CTA per motion block and plane (grid = nbh x nbv x 3 planes x lanes; one BmcArgs per lane).  Inter blocks:
 *                     luma 4-tap (-1,9,9,-1) half-pel filter, the HV phase through an int16
 *                     H-filtered strip staged in shared memory (bmc.c:124-174); chroma bilinear
 *                     (bmc.c:58-110).  Intra blocks / quadrants: integer mean of the co-located
 *                     reference (sub)block (bmc.c:256-298), one block-wide reduction.
 *                     The prediction never makes a round trip through HBM on the decoder side
 *                     (mode 2: io = clamp(pred + io - 128)); the encoder keeps it (mode 1) because the
 *                     closed-loop reconstruction adds it back after the inverse transform.
 *                     The closed-loop add-back itself is recon_kernel in frame_ops.cu.
 *
 * Reads outside the picture go through the 64-sample replicated border exactly like the reference
 * (position clamp bmc.c:221-249); filter taps that step one sample past the border see the same
 * neighbouring bytes because DevFrame keeps the reference's strides and plane order.
 */
#include "motion.cuh"

namespace dsv {

#define BMC_THREADS 256



DSV_D unsigned block_sum_u32(unsigned v, unsigned *scratch /* >= 33 */)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = __reduce_add_sync(0xffffffffu, v);
    __syncthreads();
    if (lane == 0) {
        scratch[wid] = v;
    }
    __syncthreads();
    unsigned t = 0;
    for (int i = 0; i < nw; i++) {
        t += scratch[i];
    }
    return t;
}

/* combine 4 predicted samples (packed, sample 0 in the lowest byte) with the current word and store (prediction
 * word kept when asked) */
DSV_D void bmc_store4(const BmcPlane &P, int mode, int gx, int gy, int nvalid, bool vec, unsigned pw)
{
    const size_t io = (size_t) gy * P.istride + gx, oo = (size_t) gy * P.ostride + gx;
    const int p0 = byte_of(pw, 0), p1 = byte_of(pw, 1), p2 = byte_of(pw, 2), p3 = byte_of(pw, 3);
    if (vec && nvalid == 4) {
        const unsigned cur = *reinterpret_cast<const unsigned *>(P.in + io);
        unsigned out;
        if (mode == 1) {
            out = pack_u8x4(byte_of(cur, 0) - p0 + 128, byte_of(cur, 1) - p1 + 128, byte_of(cur, 2) - p2 + 128, byte_of(cur, 3) - p3 + 128);
        } else {
            out = pack_u8x4(byte_of(cur, 0) + p0 - 128, byte_of(cur, 1) + p1 - 128, byte_of(cur, 2) + p2 - 128, byte_of(cur, 3) + p3 - 128);
        }
        *reinterpret_cast<unsigned *>(P.out + oo) = out;
        if (P.pred) {
            *reinterpret_cast<unsigned *>(P.pred + (size_t) gy * P.pstride + gx) = pw;
        }
        return;
    }
    const int pv[4] = {p0, p1, p2, p3};
#pragma unroll
    for (int e = 0; e < 4; e++) {
        if (e < nvalid) {
            const int cur = P.in[io + e];
            P.out[oo + e] = mode == 1 ? clamp_u8(cur - pv[e] + 128) : clamp_u8(pv[e] + cur - 128);
            if (P.pred) {
                P.pred[(size_t) gy * P.pstride + gx + e] = (uint8_t) pv[e];
            }
        }
    }
}

/* one motion block of one plane; a thread owns 4 horizontally adjacent samples (frames keep rows 4-byte aligned
 * at multiples of 4) */
DSV_D void bmc_block_plane(const BmcArgs &a, const int c, int16_t *hbuf, unsigned *scratch, int *s_avg)
{
    const BmcPlane &P = a.pl[c];
    const int sh = c ? a.hs : 0, sv = c ? a.vs : 0;
    const int bw = a.blk_w >> sh, bh = a.blk_h >> sv;
    const int i = blockIdx.x, j = blockIdx.y;
    const int x = i * bw, y = j * bh;
    const int cw = (x + bw >= P.w) ? P.w - x : bw;
    const int ch = (y + bh >= P.h) ? P.h - y : bh;
    if (cw <= 0 || ch <= 0) {
        return;
    }
    const DevMV mv = a.mv[j * a.nbh + i];
    const int tid = threadIdx.x;
    const int mode = a.mode;
    const int words = (cw + 3) >> 2;
    const int wsh = words > 1 ? 32 - __clz(words - 1) : 0; /* ceil(log2(words)) */
    const int wmask = (1 << wsh) - 1;
    const bool walign = ((x & 3) == 0) && ((P.istride | P.ostride | P.pstride) & 3) == 0 &&
                        ((reinterpret_cast<uintptr_t>(P.in) | reinterpret_cast<uintptr_t>(P.out) | reinterpret_cast<uintptr_t>(P.pred)) & 3) == 0;
    if (mode == 1 && x + cw == P.w) {
        /* the forward transform of a plane with odd width reads one column past it (sbt.c:583-591); in the
         * reference that column of the residual frame still holds the replicated INPUT border
         * (dsv_encoder.c:657-659: xf = copy of the padded input, then only w x h is replaced) */
        for (int k = tid; k < ch; k += BMC_THREADS) {
            P.out[(size_t) (y + k) * P.ostride + P.w] = P.in[(size_t) (y + k) * P.istride + P.w];
        }
    }

    if (mv.mode == 0) { /* DSV_MODE_INTER, bmc.c:240-254 */
        const int dx = mv.x >> sh, dy = mv.y >> sv;
        const int limx = (P.w - bw) + DSV_BORDER - 1, limy = (P.h - bh) + DSV_BORDER - 1;
        const int px = iclamp(x + (dx >> 1), -DSV_BORDER, limx);
        const int py = iclamp(y + (dy >> 1), -DSV_BORDER, limy);
        const int phase = ((dx & 1) << 1) | (dy & 1);
        const uint8_t *r0 = P.ref + (ptrdiff_t) py * P.rstride + px;
        const int rs = P.rstride;
        const int hstride = words * 4; /* int16 per staged row */
        if (c == 0 && phase == 3) {
            for (int k = tid; k < ((ch + 3) << wsh); k += BMC_THREADS) {
                const int ly = k >> wsh, wx = k & wmask;
                if (wx >= words) {
                    continue;
                }
                /* bytes p[-1..6] of the row: three aligned words, shifted once; the taps of sample e are the four
                 * bytes starting at e */
                const uint8_t *p = r0 + (ptrdiff_t) (ly - 1) * rs + 4 * wx - 1;
                const unsigned *q = reinterpret_cast<const unsigned *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t) 3);
                const unsigned bsh = ((unsigned) reinterpret_cast<uintptr_t>(p) & 3u) * 8u;
                const unsigned q0 = q[0], q1 = q[1], q2 = q[2];
                const unsigned wa = __funnelshift_r(q0, q1, bsh), wb = __funnelshift_r(q1, q2, bsh);
                const int h0 = hp_taps_u8x4(wa), h1 = hp_taps_u8x4(__funnelshift_r(wa, wb, 8));
                const int h2 = hp_taps_u8x4(__funnelshift_r(wa, wb, 16)), h3 = hp_taps_u8x4(__funnelshift_r(wa, wb, 24));
                *reinterpret_cast<uint2 *>(hbuf + ly * hstride + 4 * wx) =
                    make_uint2(__byte_perm((unsigned) h0, (unsigned) h1, 0x5410), __byte_perm((unsigned) h2, (unsigned) h3, 0x5410));
            }
            __syncthreads();
        }
        for (int k = tid; k < (ch << wsh); k += BMC_THREADS) {
            const int ly = k >> wsh, wx = k & wmask;
            if (wx >= words) {
                continue;
            }
            const int lx = 4 * wx;
            const uint8_t *p = r0 + (ptrdiff_t) ly * rs + lx;
            unsigned pw; /* the four predicted samples, saturated to 8 bits */
            if (phase == 0) {
                pw = ld4u(p);
            } else if (c == 0) {
                if (phase == 1) {
                    int v0, v1, v2, v3;
                    const unsigned wa = ld4u(p - rs), wb = ld4u(p), wc = ld4u(p + rs), wd = ld4u(p + 2 * rs);
                    /* 4x4 byte transpose: column e of the four rows becomes one word */
                    const unsigned ab0 = __byte_perm(wa, wb, 0x5140), ab1 = __byte_perm(wa, wb, 0x7362);
                    const unsigned cd0 = __byte_perm(wc, wd, 0x5140), cd1 = __byte_perm(wc, wd, 0x7362);
                    v0 = (hp_taps_u8x4(__byte_perm(ab0, cd0, 0x5410)) + 8) >> 4;
                    v1 = (hp_taps_u8x4(__byte_perm(ab0, cd0, 0x7632)) + 8) >> 4;
                    v2 = (hp_taps_u8x4(__byte_perm(ab1, cd1, 0x5410)) + 8) >> 4;
                    v3 = (hp_taps_u8x4(__byte_perm(ab1, cd1, 0x7632)) + 8) >> 4;
                    pw = pack_u8x4(v0, v1, v2, v3);
                } else if (phase == 2) {
                    int v0, v1, v2, v3;
                    const unsigned wa = ld4u(p - 1), wb = ld4u(p + 3);
                    v0 = (hp_taps_u8x4(wa) + 8) >> 4; /* saturated by the pack below */
                    v1 = (hp_taps_u8x4(__funnelshift_r(wa, wb, 8)) + 8) >> 4;
                    v2 = (hp_taps_u8x4(__funnelshift_r(wa, wb, 16)) + 8) >> 4;
                    v3 = (hp_taps_u8x4(__funnelshift_r(wa, wb, 24)) + 8) >> 4;
                    pw = pack_u8x4(v0, v1, v2, v3);
                } else {
                    int v0, v1, v2, v3;
                    const int16_t *b = hbuf + ly * hstride + lx;
                    /* rows ly-1 .. ly+2 of the 16-bit H-filtered image; (row a, row b) and (row c, row d) of one column
                     * are paired into a word each and run through dp2a with the taps (-1, 9) and (9, -1) */
                    const uint2 ra = *reinterpret_cast<const uint2 *>(b), rb = *reinterpret_cast<const uint2 *>(b + hstride);
                    const uint2 rc = *reinterpret_cast<const uint2 *>(b + 2 * hstride), rd = *reinterpret_cast<const uint2 *>(b + 3 * hstride);
                    const unsigned t_ab = 0x09ffu, t_cd = 0xff09u;
                    v0 = dp2a_lo_s16(__byte_perm(rc.x, rd.x, 0x5410), t_cd, dp2a_lo_s16(__byte_perm(ra.x, rb.x, 0x5410), t_ab, 128)) >> 8;
                    v1 = dp2a_lo_s16(__byte_perm(rc.x, rd.x, 0x7632), t_cd, dp2a_lo_s16(__byte_perm(ra.x, rb.x, 0x7632), t_ab, 128)) >> 8;
                    v2 = dp2a_lo_s16(__byte_perm(rc.y, rd.y, 0x5410), t_cd, dp2a_lo_s16(__byte_perm(ra.y, rb.y, 0x5410), t_ab, 128)) >> 8;
                    v3 = dp2a_lo_s16(__byte_perm(rc.y, rd.y, 0x7632), t_cd, dp2a_lo_s16(__byte_perm(ra.y, rb.y, 0x7632), t_ab, 128)) >> 8;
                    pw = pack_u8x4(v0, v1, v2, v3);
                }
            } else {
                if (phase == 1) {
                    pw = avg_up_u8x4(ld4u(p), ld4u(p + rs));
                } else if (phase == 2) {
                    pw = avg_up_u8x4(ld4u(p), ld4u(p + 1));
                } else {
                    int v0, v1, v2, v3;
                    const unsigned wa = ld4u(p), wb = ld4u(p + 1), wc = ld4u(p + rs), wd = ld4u(p + rs + 1);
                    v0 = (byte_of(wa, 0) + byte_of(wb, 0) + byte_of(wc, 0) + byte_of(wd, 0) + 2) >> 2;
                    v1 = (byte_of(wa, 1) + byte_of(wb, 1) + byte_of(wc, 1) + byte_of(wd, 1) + 2) >> 2;
                    v2 = (byte_of(wa, 2) + byte_of(wb, 2) + byte_of(wc, 2) + byte_of(wd, 2) + 2) >> 2;
                    v3 = (byte_of(wa, 3) + byte_of(wb, 3) + byte_of(wc, 3) + byte_of(wd, 3) + 2) >> 2;
                    pw = pack_u8x4(v0, v1, v2, v3);
                }
            }
            bmc_store4(P, mode, x + lx, y + ly, imin(4, cw - lx), walign, pw);
        }
        return;
    }

    /* intra block: mean of the co-located reference block, whole or per quadrant (bmc.c:255-298) */
    const bool whole = mv.submask == 15;
    const int sbw = whole ? cw : cw / 2, sbh = whole ? ch : ch / 2;
    const uint8_t *r0 = P.ref + (ptrdiff_t) y * P.rstride + x;
    const int nq = whole ? 1 : 4;
    for (int q = 0; q < nq; q++) {
        const int qx = (q & 1) * sbw, qy = (q >> 1) * sbh;
        unsigned acc = 0;
        if (whole || (mv.submask & (1 << q))) {
            for (int k = tid; k < sbw * sbh; k += BMC_THREADS) {
                const int ly = k / sbw, lx = k - ly * sbw;
                acc += r0[(ptrdiff_t) (qy + ly) * P.rstride + qx + lx];
            }
        }
        const unsigned tot = block_sum_u32(acc, scratch);
        if (tid == 0) {
            s_avg[q] = (sbw > 0 && sbh > 0) ? (int) (tot / (unsigned) (sbw * sbh)) : 0;
        }
    }
    __syncthreads();
    for (int k = tid; k < (ch << wsh); k += BMC_THREADS) {
        const int ly = k >> wsh, wx = k & wmask;
        if (wx >= words) {
            continue;
        }
        int v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int lx = 4 * wx + e;
            if (whole) {
                v[e] = s_avg[0];
            } else if (lx >= 2 * sbw || ly >= 2 * sbh) {
                v[e] = 0; /* odd edge blocks: the quadrants do not cover the last row/column (zeroed frame) */
            } else {
                const int q = (lx >= sbw ? 1 : 0) | (ly >= sbh ? 2 : 0);
                v[e] = (mv.submask & (1 << q)) ? s_avg[q] : (lx < cw ? (int) r0[(ptrdiff_t) ly * P.rstride + lx] : 0);
            }
        }
        bmc_store4(P, mode, x + 4 * wx, y + ly, imin(4, cw - 4 * wx), walign, pack_u8x4(v[0], v[1], v[2], v[3]));
    }
}

/* one CTA per motion block and lane, all three planes in turn (a chroma block alone is too little work to pay
 * for a CTA launch) */
/* 6 CTAs per SM (40 registers): the kernel is latency-bound, measured 268 us per 32 HD pictures against 308 us at 48
 * registers / 5 CTAs and 290 us at 32 registers with spills; unrolling the word loop is slower as well */
__global__ void __launch_bounds__(BMC_THREADS, 6) bmc_kernel(const BmcArgs *args)
{
    __shared__ __align__(8) int16_t hbuf[(DSV_BORDER + 3) * DSV_BORDER]; /* (bh + 3) x bw, bmc.c:127 */
    __shared__ unsigned scratch[40];
    __shared__ int s_avg[4];
    const BmcArgs &a = args[blockIdx.z];
    for (int c = 0; c < 3; c++) {
        bmc_block_plane(a, c, hbuf, scratch, s_avg);
        __syncthreads(); /* hbuf / s_avg are reused by the next plane */
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

void bmc_launch(const BmcArgs *d_args, int n, int nbh, int nbv, cudaStream_t st)
{
    if (n > 0) {
        DSV_LAUNCH(bmc_kernel, dim3(nbh, nbv, n), dim3(BMC_THREADS), 0, st, d_args);
        KERNEL_CHECK();
    }
}

} // namespace dsv
