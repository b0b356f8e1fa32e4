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


DSV_D int hpf4(int a, int b, int c, int d) { return 9 * (b + c) - (a + d); }

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

/* combine 4 predicted samples with the current word and store (prediction word kept when asked) */
DSV_D void bmc_store4(const BmcPlane &P, int mode, int gx, int gy, int nvalid, bool vec, int p0, int p1, int p2, int p3)
{
    const size_t io = (size_t) gy * P.istride + gx, oo = (size_t) gy * P.ostride + gx;
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
            *reinterpret_cast<unsigned *>(P.pred + (size_t) gy * P.pstride + gx) =
                (unsigned) p0 | ((unsigned) p1 << 8) | ((unsigned) p2 << 16) | ((unsigned) p3 << 24);
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
                const uint8_t *p = r0 + (ptrdiff_t) (ly - 1) * rs + 4 * wx;
                const unsigned wa = ld4u(p - 1), wb = ld4u(p + 3); /* p[-1..2], p[3..6] */
                const int b0 = byte_of(wa, 0), b1 = byte_of(wa, 1), b2 = byte_of(wa, 2), b3 = byte_of(wa, 3);
                const int b4 = byte_of(wb, 0), b5 = byte_of(wb, 1), b6 = byte_of(wb, 2);
                int16_t *d = hbuf + ly * hstride + 4 * wx;
                d[0] = (int16_t) hpf4(b0, b1, b2, b3);
                d[1] = (int16_t) hpf4(b1, b2, b3, b4);
                d[2] = (int16_t) hpf4(b2, b3, b4, b5);
                d[3] = (int16_t) hpf4(b3, b4, b5, b6);
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
            int v0, v1, v2, v3;
            if (phase == 0) {
                const unsigned w = ld4u(p);
                v0 = byte_of(w, 0); v1 = byte_of(w, 1); v2 = byte_of(w, 2); v3 = byte_of(w, 3);
            } else if (c == 0) {
                if (phase == 1) {
                    const unsigned wa = ld4u(p - rs), wb = ld4u(p), wc = ld4u(p + rs), wd = ld4u(p + 2 * rs);
                    v0 = clamp_u8((hpf4(byte_of(wa, 0), byte_of(wb, 0), byte_of(wc, 0), byte_of(wd, 0)) + 8) >> 4);
                    v1 = clamp_u8((hpf4(byte_of(wa, 1), byte_of(wb, 1), byte_of(wc, 1), byte_of(wd, 1)) + 8) >> 4);
                    v2 = clamp_u8((hpf4(byte_of(wa, 2), byte_of(wb, 2), byte_of(wc, 2), byte_of(wd, 2)) + 8) >> 4);
                    v3 = clamp_u8((hpf4(byte_of(wa, 3), byte_of(wb, 3), byte_of(wc, 3), byte_of(wd, 3)) + 8) >> 4);
                } else if (phase == 2) {
                    const unsigned wa = ld4u(p - 1), wb = ld4u(p + 3);
                    const int b0 = byte_of(wa, 0), b1 = byte_of(wa, 1), b2 = byte_of(wa, 2), b3 = byte_of(wa, 3);
                    const int b4 = byte_of(wb, 0), b5 = byte_of(wb, 1), b6 = byte_of(wb, 2);
                    v0 = clamp_u8((hpf4(b0, b1, b2, b3) + 8) >> 4);
                    v1 = clamp_u8((hpf4(b1, b2, b3, b4) + 8) >> 4);
                    v2 = clamp_u8((hpf4(b2, b3, b4, b5) + 8) >> 4);
                    v3 = clamp_u8((hpf4(b3, b4, b5, b6) + 8) >> 4);
                } else {
                    const int16_t *b = hbuf + ly * hstride + lx;
                    const short4 ra = *reinterpret_cast<const short4 *>(b), rb = *reinterpret_cast<const short4 *>(b + hstride);
                    const short4 rc = *reinterpret_cast<const short4 *>(b + 2 * hstride), rd = *reinterpret_cast<const short4 *>(b + 3 * hstride);
                    v0 = clamp_u8((hpf4(ra.x, rb.x, rc.x, rd.x) + 128) >> 8);
                    v1 = clamp_u8((hpf4(ra.y, rb.y, rc.y, rd.y) + 128) >> 8);
                    v2 = clamp_u8((hpf4(ra.z, rb.z, rc.z, rd.z) + 128) >> 8);
                    v3 = clamp_u8((hpf4(ra.w, rb.w, rc.w, rd.w) + 128) >> 8);
                }
            } else {
                unsigned w;
                if (phase == 1) {
                    w = avg_up_u8x4(ld4u(p), ld4u(p + rs));
                    v0 = byte_of(w, 0); v1 = byte_of(w, 1); v2 = byte_of(w, 2); v3 = byte_of(w, 3);
                } else if (phase == 2) {
                    w = avg_up_u8x4(ld4u(p), ld4u(p + 1));
                    v0 = byte_of(w, 0); v1 = byte_of(w, 1); v2 = byte_of(w, 2); v3 = byte_of(w, 3);
                } else {
                    const unsigned wa = ld4u(p), wb = ld4u(p + 1), wc = ld4u(p + rs), wd = ld4u(p + rs + 1);
                    v0 = (byte_of(wa, 0) + byte_of(wb, 0) + byte_of(wc, 0) + byte_of(wd, 0) + 2) >> 2;
                    v1 = (byte_of(wa, 1) + byte_of(wb, 1) + byte_of(wc, 1) + byte_of(wd, 1) + 2) >> 2;
                    v2 = (byte_of(wa, 2) + byte_of(wb, 2) + byte_of(wc, 2) + byte_of(wd, 2) + 2) >> 2;
                    v3 = (byte_of(wa, 3) + byte_of(wb, 3) + byte_of(wc, 3) + byte_of(wd, 3) + 2) >> 2;
                }
            }
            bmc_store4(P, mode, x + lx, y + ly, imin(4, cw - lx), walign, v0, v1, v2, v3);
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
        bmc_store4(P, mode, x + 4 * wx, y + ly, imin(4, cw - 4 * wx), walign, v[0], v[1], v[2], v[3]);
    }
}

/* one CTA per motion block and lane, all three planes in turn (a chroma block alone is too little work to pay
 * for a CTA launch) */
__global__ void __launch_bounds__(BMC_THREADS) bmc_kernel(const BmcArgs *args)
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
