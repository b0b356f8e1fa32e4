/*
 * sbt_fwd.cu -- forward subband transform fused with adaptive quantise + dequantise.
 *
 * Replaces dsv_fwd_sbt (sbt.c:576-651: p2sbc, fwd_b4t_2d, fwd) and the quantiser half of
 * hzcc_enc (hzcc.c:156-281: quant/dequant/quantH/dequantH with the in-place dequantised
 * write-back the encoder's own inverse transform consumes).
 *
 *   sbt_fwd_tile_kernel  one CTA per 128x64-sample tile; u8 samples are staged once in
 *                        shared memory (16-byte coalesced loads), levels 1..nlt run out of
 *                        shared memory, every high-band coefficient is quantised in
 *                        registers and stored once (row-contiguous 128 B per warp);
 *                        the tile's LL_nlt (4x2 values) goes to the llx hand-over array.
 *   sbt_fwd_lo_kernel    one CTA per plane; LL_nlt (<= 32 KB at 4K) lives in shared memory
 *                        and levels nlt+1..L ping-pong there.
 *
 * HBM traffic per plane: w*h bytes read + 4*cw*ch bytes written -- the algorithmic minimum.
 */
#include "sbt.cuh"
#include "quant.cuh"

namespace dsv {

/* forward Haar butterfly for one pair with the reference's edge rules (sbt.c:290-347) */
DSV_D void haar_fwd_pair(int x0, int x1, int x2, int x3, bool col2, bool row2, bool scale,
                         int &ll, int &lh, int &hl, int &hh)
{
    if (col2 && row2) {
        ll = x0 + x1 + x2 + x3;
        lh = x0 - x1 + x2 - x3;
        hl = x0 + x1 - x2 - x3;
        hh = x0 - x1 - x2 + x3;
    } else if (row2) {
        ll = 2 * (x0 + x2);
        hl = 2 * (x0 - x2);
        lh = hh = 0;
    } else if (col2) {
        ll = 2 * (x0 + x1);
        lh = 2 * (x0 - x1);
        hl = hh = 0;
    } else {
        ll = 4 * x0;
        lh = hl = hh = 0;
    }
    if (scale) {
        ll = ll_down(ll);
    }
}

#define FWD_IN_STRIDE 132 /* int16 elements per staged row (I frames: 1 + 128 + 1 halo, padded) */

__global__ void __launch_bounds__(SBT_TILE_THREADS) sbt_fwd_tile_kernel(const SbtJob *jobs, int njobs)
{
    __shared__ SbtJob J;
    __shared__ int32_t s_r0[(SBT_TH + 2) * FWD_IN_STRIDE / 2 + 64];
    __shared__ int32_t s_r1[(SBT_TH + 2) * SBT_TW / 2];
    __shared__ uint8_t s_stab[SBT_STAB_SMEM];
    __shared__ int s_job;
    const int tid = threadIdx.x;

    if (tid == 0) {
        s_job = sbt_find_job(jobs, njobs, (int) blockIdx.x);
    }
    __syncthreads();
    {
        const int *src = reinterpret_cast<const int *>(&jobs[s_job]);
        int *dst = reinterpret_cast<int *>(&J);
        for (int i = tid; i < (int) (sizeof(SbtJob) / sizeof(int)); i += SBT_TILE_THREADS) {
            dst[i] = src[i];
        }
    }
    __syncthreads();

    const int t = (int) blockIdx.x - J.tile_base;
    const int tx = t % J.tiles_x, ty = t / J.tiles_x;
    const int gx0 = tx * SBT_TW, gy0 = ty * SBT_TH;
    const int cw = J.cw, ch = J.ch;
    const bool isI = !J.isP;
    const uint8_t *stab = J.stable;
    if (J.do_quant) {
        int nb = J.pq.nbh * J.pq.nbv;
        if (nb <= SBT_STAB_SMEM) {
            for (int i = tid; i < nb; i += SBT_TILE_THREADS) {
                s_stab[i] = J.stable[i];
            }
            stab = s_stab;
        }
    }

    int16_t *in = reinterpret_cast<int16_t *>(s_r0);
    int32_t *llA, *llB;

    /* ---- stage samples: d = pix - 128, zero for rows >= ph (sbt.c:576-592) -------------- */
    {
        const int halo = isI ? 1 : 0;
        const int rows = SBT_TH + 2 * halo;
        const int istride = isI ? FWD_IN_STRIDE : SBT_TW;
        for (int task = tid; task < rows * (SBT_TW / 16); task += SBT_TILE_THREADS) {
            int lr = task >> 3, ck = task & 7;
            int gr = gy0 - halo + lr;
            if (isI) { /* B4T edge rule on rows: x[-1] := x[1], x[n] := x[n-1] */
                gr = gr == -1 ? 1 : (gr == ch ? ch - 1 : gr);
            }
            bool rvalid = gr >= 0 && gr < ch && gr < J.ph;
            int gc = gx0 + ck * 16;
            int16_t *dst = in + lr * istride + halo + ck * 16;
            const uint8_t *src = J.pix + (size_t) (rvalid ? gr : 0) * J.pstride + gc;
            if (rvalid && gc + 16 <= cw && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
                uint4 v = *reinterpret_cast<const uint4 *>(src);
                unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    dst[e] = (int16_t) ((int) ((w[e >> 2] >> (8 * (e & 3))) & 0xff) - 128);
                }
            } else {
#pragma unroll 4
                for (int e = 0; e < 16; e++) {
                    int c = gc + e;
                    if (isI && c == cw) {
                        c = cw - 1;
                    }
                    dst[e] = (rvalid && c < cw) ? (int16_t) ((int) J.pix[(size_t) gr * J.pstride + c] - 128) : (int16_t) 0;
                }
            }
        }
        if (isI) { /* halo columns gx0-1 and gx0+128 */
            for (int task = tid; task < rows * 2; task += SBT_TILE_THREADS) {
                int lr = task >> 1, side = task & 1;
                int gr = gy0 - 1 + lr;
                gr = gr == -1 ? 1 : (gr == ch ? ch - 1 : gr);
                bool rvalid = gr >= 0 && gr < ch && gr < J.ph;
                int c = side ? gx0 + SBT_TW : gx0 - 1;
                c = c == -1 ? 1 : (c == cw ? cw - 1 : c);
                bool ok = rvalid && c >= 0 && c < cw;
                in[lr * FWD_IN_STRIDE + (side ? SBT_TW + 1 : 0)] =
                    ok ? (int16_t) ((int) J.pix[(size_t) gr * J.pstride + c] - 128) : (int16_t) 0;
            }
        }
    }
    __syncthreads();

    /* ---- level 1 ------------------------------------------------------------------------- */
    if (isI) {
        /* B4T rows (sbt.c:91-126): hb[r][k] = L, hb[r][64+k] = H for the 66 staged rows */
        int16_t *hb = reinterpret_cast<int16_t *>(s_r1);
        for (int task = tid; task < (SBT_TH + 2) * (SBT_TW / 2); task += SBT_TILE_THREADS) {
            int lr = task >> 6, k = task & 63;
            const int16_t *p = in + lr * FWD_IN_STRIDE + 2 * k;
            int xp = p[0], a = p[1], b = p[2], xn = p[3];
            hb[lr * SBT_TW + k] = (int16_t) rnd_shift<1>(3 * a + 3 * b - xp - xn);
            hb[lr * SBT_TW + 64 + k] = (int16_t) rnd_shift<1>(xp - 3 * a + 3 * b - xn);
        }
        __syncthreads();
        /* B4T columns (sbt.c:166-201) on the row-transformed data; LL1 stays in shared memory */
        llA = s_r0;
        llB = s_r0 + (SBT_TW / 2) * (SBT_TH / 2);
        for (int task = tid; task < SBT_TW * (SBT_TH / 2); task += SBT_TILE_THREADS) {
            int c = task & 127, m = task >> 7;
            const int16_t *p = hb + (2 * m) * SBT_TW + c;
            int xp = p[0], a = p[SBT_TW], b = p[2 * SBT_TW], xn = p[3 * SBT_TW];
            int lo = rnd_shift<1>(3 * a + 3 * b - xp - xn);
            int hi = rnd_shift<1>(xp - 3 * a + 3 * b - xn);
            int gm = ty * (SBT_TH / 2) + m;
            int k = c & 63, gk = tx * (SBT_TW / 2) + k;
            bool valid = gm < (ch >> 1) && gk < (cw >> 1);
            if (c < 64) {
                llA[m * 64 + k] = lo;
                if (valid) {
                    emit_h(J, J.do_quant != 0, stab, 1, 2, gk, gm, hi);
                }
            } else if (valid) {
                emit_h(J, J.do_quant != 0, stab, 1, 1, gk, gm, lo);
                emit_h(J, J.do_quant != 0, stab, 1, 3, gk, gm, hi);
            }
        }
    } else {
        /* Haar level 1 straight from the staged samples (sbt.c:268-349, no LL scaling at P level 1) */
        llA = s_r1;
        llB = s_r1 + (SBT_TW / 2) * (SBT_TH / 2);
        const int ws = cw, hs = ch, wo = sbt_wo(cw, 1), ho = sbt_wo(ch, 1);
        for (int task = tid; task < (SBT_TW / 2) * (SBT_TH / 2); task += SBT_TILE_THREADS) {
            int ix = task & 63, iy = task >> 6;
            int gx = tx * (SBT_TW / 2) + ix, gy = ty * (SBT_TH / 2) + iy;
            if (gx < wo && gy < ho) {
                const int16_t *p = in + (2 * iy) * SBT_TW + 2 * ix;
                bool col2 = 2 * gx + 1 < ws, row2 = 2 * gy + 1 < hs;
                int ll, lh, hl, hh;
                haar_fwd_pair(p[0], col2 ? p[1] : 0, row2 ? p[SBT_TW] : 0, (col2 && row2) ? p[SBT_TW + 1] : 0,
                              col2, row2, false, ll, lh, hl, hh);
                llA[iy * 64 + ix] = ll;
                if (col2) {
                    emit_h(J, J.do_quant != 0, stab, 1, 1, gx, gy, lh);
                }
                if (row2) {
                    emit_h(J, J.do_quant != 0, stab, 1, 2, gx, gy, hl);
                }
                if (col2 && row2) {
                    emit_h(J, J.do_quant != 0, stab, 1, 3, gx, gy, hh);
                }
            }
        }
    }
    __syncthreads();

    /* ---- levels 2..nlt: Haar on the in-tile LL (LL scaled by 4/5: I always, P for level > 1) ---- */
    int iw = SBT_TW / 2, ih = SBT_TH / 2;
    for (int lvl = 2; lvl <= J.nlt; lvl++) {
        const int ow = iw >> 1, oh = ih >> 1;
        const int ws = sbt_ws(cw, lvl), hs = sbt_ws(ch, lvl), wo = sbt_wo(cw, lvl), ho = sbt_wo(ch, lvl);
        for (int task = tid; task < ow * oh; task += SBT_TILE_THREADS) {
            int ix = task % ow, iy = task / ow;
            int gx = tx * ow + ix, gy = ty * oh + iy;
            if (gx < wo && gy < ho) {
                const int32_t *p = llA + (2 * iy) * iw + 2 * ix;
                bool col2 = 2 * gx + 1 < ws, row2 = 2 * gy + 1 < hs;
                int ll, lh, hl, hh;
                haar_fwd_pair(p[0], col2 ? p[1] : 0, row2 ? p[iw] : 0, (col2 && row2) ? p[iw + 1] : 0,
                              col2, row2, true, ll, lh, hl, hh);
                llB[iy * ow + ix] = ll;
                if (col2) {
                    emit_h(J, J.do_quant != 0, stab, lvl, 1, gx, gy, lh);
                }
                if (row2) {
                    emit_h(J, J.do_quant != 0, stab, lvl, 2, gx, gy, hl);
                }
                if (col2 && row2) {
                    emit_h(J, J.do_quant != 0, stab, lvl, 3, gx, gy, hh);
                }
            }
        }
        __syncthreads();
        int32_t *tsw = llA;
        llA = llB;
        llB = tsw;
        iw = ow;
        ih = oh;
    }

    /* ---- hand LL_nlt to the lo kernel ---------------------------------------------------- */
    {
        const int wo = sbt_wo(cw, J.nlt), ho = sbt_wo(ch, J.nlt);
        for (int task = tid; task < iw * ih; task += SBT_TILE_THREADS) {
            int ix = task % iw, iy = task / iw;
            int gx = tx * iw + ix, gy = ty * ih + iy;
            if (gx < wo && gy < ho) {
                J.llx[gy * wo + gx] = llA[iy * iw + ix];
            }
        }
    }
}

__global__ void __launch_bounds__(SBT_LO_THREADS) sbt_fwd_lo_kernel(const SbtJob *jobs)
{
    DSV_DYN_SMEM(int32_t, sm);
    __shared__ SbtJob J;
    const int tid = threadIdx.x;
    {
        const int *src = reinterpret_cast<const int *>(&jobs[blockIdx.x]);
        int *dst = reinterpret_cast<int *>(&J);
        for (int i = tid; i < (int) (sizeof(SbtJob) / sizeof(int)); i += SBT_LO_THREADS) {
            dst[i] = src[i];
        }
    }
    __syncthreads();
    const int cw = J.cw, ch = J.ch;
    int ws = sbt_wo(cw, J.nlt), hs = sbt_wo(ch, J.nlt);
    int32_t *A = sm, *B = sm + ws * hs;
    for (int i = tid; i < ws * hs; i += SBT_LO_THREADS) {
        A[i] = J.llx[i];
    }
    __syncthreads();
    for (int lvl = J.nlt + 1; lvl <= J.lvls; lvl++) {
        const int wo = sbt_wo(cw, lvl), ho = sbt_wo(ch, lvl);
        for (int task = tid; task < wo * ho; task += SBT_LO_THREADS) {
            int ix = task % wo, iy = task / wo;
            const int32_t *p = A + (2 * iy) * ws + 2 * ix;
            bool col2 = 2 * ix + 1 < ws, row2 = 2 * iy + 1 < hs;
            int ll, lh, hl, hh;
            haar_fwd_pair(p[0], col2 ? p[1] : 0, row2 ? p[ws] : 0, (col2 && row2) ? p[ws + 1] : 0,
                          col2, row2, true, ll, lh, hl, hh);
            B[iy * wo + ix] = ll;
            if (col2) {
                emit_h(J, J.do_quant != 0, J.stable, lvl, 1, ix, iy, lh);
            }
            if (row2) {
                emit_h(J, J.do_quant != 0, J.stable, lvl, 2, ix, iy, hl);
            }
            if (col2 && row2) {
                emit_h(J, J.do_quant != 0, J.stable, lvl, 3, ix, iy, hh);
            }
        }
        __syncthreads();
        int32_t *tsw = A;
        A = B;
        B = tsw;
        ws = wo;
        hs = ho;
    }
    if (tid == 0) {
        J.coef[0] = A[0]; /* DC is carried unquantised (hzcc.c:462-465) */
    }
}

void sbt_fwd_launch(const SbtJob *d_jobs, int njobs, int total_tiles, size_t lo_smem, cudaStream_t st,
                    cudaEvent_t ev0, cudaEvent_t ev1)
{
    if (lo_smem > 48 * 1024) {
        CUDA_CHECK(cudaFuncSetAttribute(sbt_fwd_lo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) lo_smem));
    }
    if (ev0) {
        CUDA_CHECK(cudaEventRecord(ev0, st));
    }
    DSV_LAUNCH(sbt_fwd_tile_kernel, dim3(total_tiles), dim3(SBT_TILE_THREADS), 0, st, d_jobs, njobs);
    KERNEL_CHECK();
    if (ev1) {
        CUDA_CHECK(cudaEventRecord(ev1, st));
    }
    DSV_LAUNCH(sbt_fwd_lo_kernel, dim3(njobs), dim3(SBT_LO_THREADS), lo_smem, st, d_jobs);
    KERNEL_CHECK();
}

} // namespace dsv
