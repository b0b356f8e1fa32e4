/*
 * sbt_fwd.cu -- forward subband transform fused with adaptive quantise + dequantise.
 *
 * Replaces dsv_fwd_sbt (sbt.c:576-651: p2sbc, fwd_b4t_2d, fwd) and the quantiser half of
 * hzcc_enc (hzcc.c:156-281: quant/dequant/quantH/dequantH with the in-place dequantised
 * write-back the encoder's own inverse transform consumes).
 *
 *   sbt_fwd_tile_kernel  one CTA per 128x64-sample tile.  Level 1 reads the u8 samples straight from
 *                        global memory (8-byte loads, a thread owns 4 adjacent quads), quantises the
 *                        three high bands in registers (one stability lookup per quad) and stores
 *                        each band with one 16-byte store; only LL goes to shared memory, where
 *                        levels 2..nlt run.  I frames: B4T row pass from global into an int16 strip,
 *                        column pass from shared memory.  The tile's LL_nlt (4x2 values) goes to the
 *                        llx hand-over array.
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

/* ---- register-level quantiser for the hot levels --------------------------------------------------- */

/* top level (transform level 1): magnitude shift quantise + dequantise (hzcc.c:114-135) */
DSV_D int requant_p2(int v, int sh)
{
    const int m = (iabs(v) >> sh) << sh;
    return v < 0 ? -m : m;
}
/* dead-zone quantise + dequantise (hzcc.c:94-128): (r*2q + q) >> 1 == r*q + (q >> 1) */
DSV_D int requant_dz(int v, int q, const FastDiv &two_q)
{
    const unsigned m = (unsigned) iabs(v) * 2u;
    if (m <= (unsigned) q) {
        return 0;
    }
    const int sym = (int) fastdiv(m + 1u, two_q); /* may still be 0 just above the dead zone */
    if (sym == 0) {
        return 0;
    }
    const int r = sym * q + (q >> 1);
    return v < 0 ? -r : r;
}

DSV_D void store_n(int32_t *dst, const int *v, int n, int nvalid)
{
    if (nvalid == n && n == 4 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        *reinterpret_cast<int4 *>(dst) = make_int4(v[0], v[1], v[2], v[3]);
    } else if (nvalid == n && n == 2 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
        *reinterpret_cast<int2 *>(dst) = make_int2(v[0], v[1]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i < n && i < nvalid) {
                dst[i] = v[i];
            }
        }
    }
}

/*
 * Quantise (optionally) and store N horizontally adjacent quads' high bands of transform level `lvl`
 * (1..3) at band-local (bx .. bx+N-1, by).  The stability flag is looked up once per quad and shared by
 * its three bands; levels whose scan regions overlap the next level's (odd sizes, SURVEY.md Appendix
 * B-1) take the generic per-coefficient path.
 */
template <int N>
DSV_D int emit_quads(const SbtJob &J, const uint8_t *stab, int lvl, int wo, int ho, int bx, int by, int nvalid,
                     int *lh, int *hl, int *hh)
{
    const int cw = J.cw;
    if (J.do_quant) {
        if (lvl <= 2 && (J.dg.dvx[lvl] >= 0 || J.dg.dvy[lvl] >= 0)) {
#pragma unroll
            for (int i = 0; i < N; i++) {
                if (i < nvalid) {
                    emit_h(J, true, stab, lvl, 1, bx + i, by, lh[i]);
                    emit_h(J, true, stab, lvl, 2, bx + i, by, hl[i]);
                    emit_h(J, true, stab, lvl, 3, bx + i, by, hh[i]);
                }
            }
            return 1; /* values go through the generic path: reported as "may be non-zero" */
        }
        const int rowterm = ((by * J.pq.dby[lvl]) >> 14) * J.pq.nbh;
        const int dbx = J.pq.dbx[lvl];
#pragma unroll
        for (int i = 0; i < N; i++) {
            const int f = stab[rowterm + (((bx + i) * dbx) >> 14)];
            if (lvl == 1) {
                const int sh = f ? J.pq.sh_hq : J.pq.sh_plain;
                lh[i] = requant_p2(lh[i], sh);
                hl[i] = requant_p2(hl[i], sh);
                hh[i] = requant_p2(hh[i], sh);
            } else {
                const int sel = (f & 2) ? 2 : (f ? 1 : 0);
                const LevelQ &L = J.pq.lv[3 - lvl];
                const int q = L.q[sel];
                const FastDiv fd = L.fd[sel];
                lh[i] = requant_dz(lh[i], q, fd);
                hl[i] = requant_dz(hl[i], q, fd);
                hh[i] = requant_dz(hh[i], q, fd);
            }
        }
    }
    int32_t *row0 = J.coef + (size_t) by * cw + bx, *row1 = J.coef + (size_t) (ho + by) * cw + bx;
    store_n(row0 + wo, lh, N, nvalid);
    store_n(row1, hl, N, nvalid);
    store_n(row1 + wo, hh, N, nvalid);
    int any = 0;
    if (nvalid == N) { /* the usual case: no per-quad predicate */
#pragma unroll
        for (int i = 0; i < N; i++) {
            any |= lh[i] | hl[i] | hh[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (i < nvalid) {
                any |= lh[i] | hl[i] | hh[i];
            }
        }
    }
    return any;
}

#define FWD_HB_STRIDE 136 /* int16 per row of the B4T row-pass buffer: 64 L + 64 H, padded */

/* sample (r, c) of the plane as the transform sees it: pix - 128 inside w x ph, 0 below the picture */
DSV_D int fwd_sample(const SbtJob &J, int r, int c)
{
    return r < J.ph ? (int) J.pix[(size_t) r * J.pstride + c] - 128 : 0;
}

/* one-barrier prologue shared by the tile kernels: every thread finds the job (same broadcast loads), the
 * job record is copied to shared memory */
template <bool MID> DSV_D int load_job(SbtJob *sJ, const SbtJob *jobs, const SbtDims &dims)
{
    int tile;
    const int job = sbt_locate<MID>(jobs, dims, (int) blockIdx.x, &tile);
    const int *src = reinterpret_cast<const int *>(&jobs[job]);
    int *dst = reinterpret_cast<int *>(sJ);
    for (int i = threadIdx.x; i < (int) (sizeof(SbtJob) / sizeof(int)); i += blockDim.x) {
        dst[i] = src[i];
    }
    __syncthreads();
    return tile;
}

/* 8 samples of rows r0, r0+1 starting at column c0 as two pairs of packed words; samples outside w x ph read as
 * 128 (i.e. 0 after the -128 offset, sbt.c:576-592), nv = valid quads (1..4) */
DSV_D void load_quads_slow(const SbtJob &J, int r0, int c0, int nv, unsigned &w00, unsigned &w01, unsigned &w10, unsigned &w11)
{
    w00 = w01 = w10 = w11 = 0x80808080u;
    const bool ok0 = r0 < J.ph, ok1 = r0 + 1 < J.ph;
    const uint8_t *p0 = J.pix + (size_t) r0 * J.pstride + c0, *p1 = p0 + J.pstride;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        if (e < 2 * nv) {
            const unsigned a = ok0 ? p0[e] : 128u, b = ok1 ? p1[e] : 128u;
            const unsigned m = ~(0xffu << (8 * (e & 3)));
            if (e < 4) {
                w00 = (w00 & m) | (a << (8 * (e & 3)));
                w10 = (w10 & m) | (b << (8 * (e & 3)));
            } else {
                w01 = (w01 & m) | (a << (8 * (e & 3)));
                w11 = (w11 & m) | (b << (8 * (e & 3)));
            }
        }
    }
}

/*
 * Streaming kernel: levels 1 and 2 of one 128x64-sample tile.  LL_2 (32x16 values) goes to the job's ll2
 * hand-over plane for the mid kernel.
 */
#ifndef SBT_FWD_MINB
#define SBT_FWD_MINB 8 /* CTAs per SM, measured per 32 HD pictures: 8 -> 118.4 us, 6 (compiler default) -> 120.0, 5 -> 125.0 */
#endif
__global__ void __launch_bounds__(SBT_TILE_THREADS, SBT_FWD_MINB) sbt_fwd_tile_kernel(const SbtJob *jobs, const SbtDims dims)
{
    DSV_DYN_SMEM(int16_t, s_hb); /* I frames only: B4T row-pass strip, (64 + 2) x FWD_HB_STRIDE int16 */
    __shared__ SbtJob J;
    __shared__ __align__(16) int32_t s_ll1[(SBT_TW / 2) * (SBT_TH / 2)];
    const int tid = threadIdx.x;
    const int t = load_job<false>(&J, jobs, dims);

    const int ty = (int) fastdiv((unsigned) t, J.tiles_x_fd), tx = t - ty * J.tiles_x;
    const int gx0 = tx * SBT_TW, gy0 = ty * SBT_TH;
    const int cw = J.cw, ch = J.ch;
    const bool isI = !J.isP;
    const uint8_t *stab = J.stable;
    const int wo1 = cw >> 1, ho1 = ch >> 1; /* cw, ch are even */
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(J.pix) | (uintptr_t) J.pstride) & 7) == 0;
    int32_t *llA = s_ll1;
    int any1 = isI ? 1 : 0, any2 = 0; /* tile flags (sbt.cuh): level 1 of an I picture is dense anyway */

    /* ---- level 1 ------------------------------------------------------------------------- */
    if (!isI) {
        /* Haar straight from global memory: a thread owns 4 adjacent quads (8 samples x 2 rows, two 8-byte
         * loads) of two tile rows 16 quad-rows apart; no LL scaling at level 1 of P frames (sbt.c:20-22,268-349) */
        unsigned w0[2][2], w1[2][2];
        int nv[2], gxs[2], gys[2];
#pragma unroll
        for (int it = 0; it < 2; it++) {
            const int g = tid + it * SBT_TILE_THREADS;
            const int qrow = g >> 4, qc = (g & 15) * 4;
            const int gy = ty * (SBT_TH / 2) + qrow, gx = tx * (SBT_TW / 2) + qc;
            gxs[it] = gx;
            gys[it] = gy;
            nv[it] = (gy >= ho1 || gx >= wo1) ? 0 : imin(4, wo1 - gx);
            const int r0 = 2 * gy, c0 = 2 * gx;
            if (nv[it] == 4 && vec_ok && r0 + 1 < J.ph) {
                const uint2 a = *reinterpret_cast<const uint2 *>(J.pix + (size_t) r0 * J.pstride + c0);
                const uint2 b = *reinterpret_cast<const uint2 *>(J.pix + (size_t) (r0 + 1) * J.pstride + c0);
                w0[it][0] = a.x; w0[it][1] = a.y; w1[it][0] = b.x; w1[it][1] = b.y;
            } else if (nv[it] > 0) {
                load_quads_slow(J, r0, c0, nv[it], w0[it][0], w0[it][1], w1[it][0], w1[it][1]);
            }
        }
#pragma unroll
        for (int it = 0; it < 2; it++) {
            if (nv[it] == 0) {
                continue;
            }
            const int g = tid + it * SBT_TILE_THREADS;
            const int qrow = g >> 4, qc = (g & 15) * 4;
            int ll[4], lh[4], hl[4], hh[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const unsigned wa = w0[it][i >> 1], wb = w1[it][i >> 1];
                const int sft = (i & 1) * 16;
                const int x0 = (int) ((wa >> sft) & 0xff), x1 = (int) ((wa >> (sft + 8)) & 0xff);
                const int x2 = (int) ((wb >> sft) & 0xff), x3 = (int) ((wb >> (sft + 8)) & 0xff);
                const int a = x0 + x1, b = x0 - x1, c = x2 + x3, d = x2 - x3;
                ll[i] = a + c - 512; /* the -128 offsets cancel in the three high bands */
                hl[i] = a - c;
                lh[i] = b + d;
                hh[i] = b - d;
            }
            *reinterpret_cast<int4 *>(llA + qrow * (SBT_TW / 2) + qc) = make_int4(ll[0], ll[1], ll[2], ll[3]);
            any1 |= emit_quads<4>(J, stab, 1, wo1, ho1, gxs[it], gys[it], nv[it], lh, hl, hh);
        }
    } else {
        /* B4T rows (sbt.c:91-126) straight from global memory: s_hb[r][k] = L, s_hb[r][64 + k] = H for the
         * 64 tile rows plus one halo row above and below; edge rule x[-1] := x[1], x[n] := x[n-1] */
        for (int g = tid; g < (SBT_TH + 2) * (SBT_TW / 8); g += SBT_TILE_THREADS) {
            const int lr = g >> 4, k0 = (g & 15) * 4;
            int gr = gy0 - 1 + lr;
            gr = gr == -1 ? 1 : (gr == ch ? ch - 1 : gr);
            const int c0 = gx0 + 2 * k0; /* first of the 8 samples this thread transforms */
            int16_t *dst = s_hb + lr * FWD_HB_STRIDE + k0;
            if (gr > ch || c0 >= cw) {
                continue;
            }
            int x[10];
            if (vec_ok && c0 + 8 <= cw && gr < J.ph) {
                const uint8_t *row = J.pix + (size_t) gr * J.pstride;
                const uint2 a = *reinterpret_cast<const uint2 *>(row + c0);
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    x[1 + e] = (int) ((a.x >> (8 * e)) & 0xff) - 128;
                    x[5 + e] = (int) ((a.y >> (8 * e)) & 0xff) - 128;
                }
                x[0] = (int) row[c0 == 0 ? 1 : c0 - 1] - 128;
                x[9] = (int) row[c0 + 8 == cw ? cw - 1 : c0 + 8] - 128;
            } else {
#pragma unroll
                for (int e = 0; e < 10; e++) {
                    int c = c0 - 1 + e;
                    c = c == -1 ? 1 : (c >= cw ? cw - 1 : c);
                    x[e] = fwd_sample(J, gr, c);
                }
            }
            const int nk = imin(4, (cw - c0) >> 1);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k < nk) {
                    const int xp = x[2 * k], a = x[2 * k + 1], b = x[2 * k + 2], xn = x[2 * k + 3];
                    dst[k] = (int16_t) rnd_shift<1>(3 * a + 3 * b - xp - xn);
                    dst[64 + k] = (int16_t) rnd_shift<1>(xp - 3 * a + 3 * b - xn);
                }
            }
        }
        __syncthreads();
        /* B4T columns (sbt.c:166-201): a thread owns 4 adjacent columns of one output row pair */
        for (int g = tid; g < (SBT_TW / 4) * (SBT_TH / 2); g += SBT_TILE_THREADS) {
            const int m = g >> 5, c4 = (g & 31) * 4;
            const bool hcol = c4 >= 64;
            const int k0 = c4 & 63;
            const int gm = ty * (SBT_TH / 2) + m, gk = tx * (SBT_TW / 2) + k0;
            if (gm >= ho1 || gk >= wo1) {
                continue;
            }
            const int nv = imin(4, wo1 - gk);
            const int16_t *p = s_hb + (2 * m) * FWD_HB_STRIDE + c4;
            int lo[4], hi[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int xp = p[i], a = p[FWD_HB_STRIDE + i], b = p[2 * FWD_HB_STRIDE + i], xn = p[3 * FWD_HB_STRIDE + i];
                lo[i] = rnd_shift<1>(3 * a + 3 * b - xp - xn);
                hi[i] = rnd_shift<1>(xp - 3 * a + 3 * b - xn);
            }
            if (!hcol) {
                *reinterpret_cast<int4 *>(llA + m * (SBT_TW / 2) + k0) = make_int4(lo[0], lo[1], lo[2], lo[3]);
            }
            /* low columns carry (LL, HL), high columns (LH, HH) */
            int32_t *row0 = J.coef + (size_t) gm * cw + gk, *row1 = J.coef + (size_t) (ho1 + gm) * cw + gk;
            if (J.do_quant && (J.dg.dvx[1] >= 0 || J.dg.dvy[1] >= 0)) {
                /* odd next-level size: some of these positions are scanned twice */
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (i < nv) {
                        if (!hcol) {
                            emit_h(J, true, stab, 1, 2, gk + i, gm, hi[i]);
                        } else {
                            emit_h(J, true, stab, 1, 1, gk + i, gm, lo[i]);
                            emit_h(J, true, stab, 1, 3, gk + i, gm, hi[i]);
                        }
                    }
                }
                continue;
            }
            if (J.do_quant) {
                const int rowterm = ((gm * J.pq.dby[1]) >> 14) * J.pq.nbh;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int f = stab[rowterm + (((gk + i) * J.pq.dbx[1]) >> 14)];
                    const int sh = f ? J.pq.sh_hq : J.pq.sh_plain;
                    hi[i] = requant_p2(hi[i], sh);
                    if (hcol) {
                        lo[i] = requant_p2(lo[i], sh);
                    }
                }
            }
            if (!hcol) {
                store_n(row1, hi, 4, nv); /* HL */
            } else {
                store_n(row0 + wo1, lo, 4, nv); /* LH */
                store_n(row1 + wo1, hi, 4, nv); /* HH */
            }
        }
    }
    __syncthreads();

    /* ---- level 2: a thread owns 2 adjacent quads (LL scaled by 4/5: I always, P for level > 1) ---- */
    {
        const int iw = SBT_TW / 2;
        const int ow = iw >> 1, oh = SBT_TH / 4;
        const int ws = sbt_ws(cw, 2), hs = sbt_ws(ch, 2), wo = sbt_wo(cw, 2), ho = sbt_wo(ch, 2);
        int32_t *ll2 = J.llx + J.ll2_off;
        const int g = tid; /* (ow / 2) * oh == 256 */
        const int iy = g >> 4, ix = (g & 15) * 2;
        const int gx = tx * ow + ix, gy = ty * oh + iy;
        if (gx < wo && gy < ho) {
            const int nv = imin(2, wo - gx);
            const bool row2 = 2 * gy + 1 < hs;
            const int4 a = *reinterpret_cast<const int4 *>(llA + (2 * iy) * iw + 2 * ix);
            const int4 b = *reinterpret_cast<const int4 *>(llA + (2 * iy + 1) * iw + 2 * ix);
            int ll[2], lh[2], hl[2], hh[2];
            {
                const bool col2 = 2 * gx + 1 < ws;
                haar_fwd_pair(a.x, col2 ? a.y : 0, row2 ? b.x : 0, (col2 && row2) ? b.y : 0, col2, row2, true, ll[0], lh[0], hl[0], hh[0]);
            }
            {
                const bool col2 = 2 * gx + 3 < ws;
                haar_fwd_pair(a.z, col2 ? a.w : 0, row2 ? b.z : 0, (col2 && row2) ? b.w : 0, col2, row2, true, ll[1], lh[1], hl[1], hh[1]);
            }
            store_n(ll2 + (size_t) gy * wo + gx, ll, 2, nv);
            const bool all2 = row2 && (2 * gx + 2 * nv - 1 < ws);
            if (all2) {
                any2 |= emit_quads<2>(J, stab, 2, wo, ho, gx, gy, nv, lh, hl, hh);
            } else {
                any2 = 1; /* odd edges: generic path */
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const bool col2 = i < nv && 2 * (gx + i) + 1 < ws;
                    if (i >= nv) {
                        continue;
                    }
                    if (col2) {
                        emit_h(J, J.do_quant != 0, stab, 2, 1, gx + i, gy, lh[i]);
                    }
                    if (row2) {
                        emit_h(J, J.do_quant != 0, stab, 2, 2, gx + i, gy, hl[i]);
                    }
                    if (col2 && row2) {
                        emit_h(J, J.do_quant != 0, stab, 2, 3, gx + i, gy, hh[i]);
                    }
                }
            }
        }
    }
    if (J.tflags) {
        /* the flag bytes were zeroed before the launch: a warp that stored a non-zero coefficient ORs its bits in (one
         * atomic per warp at most, none at all for the empty level-1 bands of a P picture) -- no block-wide barrier */
        const unsigned m1 = __ballot_sync(0xffffffffu, any1 != 0), m2 = __ballot_sync(0xffffffffu, any2 != 0);
        const unsigned bits = (m1 ? 1u : 0u) | (m2 ? 2u : 0u);
        if (bits && (tid & 31) == 0) {
            const uintptr_t a = reinterpret_cast<uintptr_t>(J.tflags + t);
            atomicOr(reinterpret_cast<unsigned *>(a & ~(uintptr_t) 3), bits << (8 * (a & 3)));
        }
    }
}

/*
 * Mid kernel: levels 3..nlt of one 128x64 block of LL_2 (= a 512x256-sample region), one quad per thread
 * per level (6 % of the coefficients); LL_nlt goes to the llx hand-over array.
 */
__global__ void __launch_bounds__(SBT_TILE_THREADS) sbt_fwd_mid_kernel(const SbtJob *jobs, const SbtDims dims)
{
    __shared__ SbtJob J;
    __shared__ __align__(16) int32_t s_a[SBT_TW * SBT_TH];
    __shared__ __align__(16) int32_t s_b[(SBT_TW / 2) * (SBT_TH / 2)];
    const int tid = threadIdx.x;
    const int t = load_job<true>(&J, jobs, dims);
    const int ty = (int) fastdiv((unsigned) t, J.mtiles_x_fd), tx = t - ty * J.mtiles_x;
    const int cw = J.cw, ch = J.ch;
    const uint8_t *stab = J.stable;
    {
        const int w2 = sbt_wo(cw, 2), h2 = sbt_wo(ch, 2);
        const int32_t *ll2 = J.llx + J.ll2_off;
        for (int i = tid; i < SBT_TW * SBT_TH; i += SBT_TILE_THREADS) {
            const int lx = i & (SBT_TW - 1), ly = i >> 7;
            const int x = tx * SBT_TW + lx, y = ty * SBT_TH + ly;
            s_a[i] = (x < w2 && y < h2) ? ll2[(size_t) y * w2 + x] : 0;
        }
    }
    __syncthreads();
    int32_t *llA = s_a, *llB = s_b;
    int iw = SBT_TW, ih = SBT_TH;
    for (int lvl = SBT_HI + 1; lvl <= J.nlt; lvl++) {
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

void sbt_fwd_launch(const SbtJob *d_jobs, const SbtDims &dims, size_t lo_smem, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1)
{
    if (dims.njobs <= 0) {
        return;
    }
    if (lo_smem > 48 * 1024) {
        CUDA_CHECK(cudaFuncSetAttribute(sbt_fwd_lo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) lo_smem));
    }
    const size_t hb = dims.any_intra ? (size_t) (SBT_TH + 2) * FWD_HB_STRIDE * sizeof(int16_t) : 0;
    if (ev0) {
        CUDA_CHECK(cudaEventRecord(ev0, st));
    }
    DSV_LAUNCH(sbt_fwd_tile_kernel, dim3(dims.tiles), dim3(SBT_TILE_THREADS), hb, st, d_jobs, dims);
    KERNEL_CHECK();
    if (ev1) {
        CUDA_CHECK(cudaEventRecord(ev1, st));
    }
    DSV_LAUNCH(sbt_fwd_mid_kernel, dim3(dims.mtiles), dim3(SBT_TILE_THREADS), 0, st, d_jobs, dims);
    KERNEL_CHECK();
    DSV_LAUNCH(sbt_fwd_lo_kernel, dim3(dims.njobs), dim3(SBT_LO_THREADS), lo_smem, st, d_jobs);
    KERNEL_CHECK();
}

} // namespace dsv
