/*
 * sbt_inv.cu -- inverse subband transform, int32 coefficients -> u8 samples.
 *
 * Replaces dsv_inv_sbt (sbt.c:653-714): inv (luma, smoothing-filtered Haar, sbt.c:438-574),
 * inv_simple (chroma, sbt.c:352-435), inv_b4t_2d (level 1 of I frames, sbt.c:129-163,
 * 204-238, 253-265) and sbc2int (sbt.c:594-614).
 *
 *   sbt_inv_lo_kernel    one CTA per plane: levels L..nlt+1 in shared memory, result LL_nlt
 *                        -> llx hand-over array.
 *   sbt_inv_tile_kernel  one CTA per 128x64-sample tile: levels nlt..1 out of shared memory.
 *                        Because the filtered inverse looks at the neighbouring LL values
 *                        (and B4T at neighbouring L/H), each level's LL window carries a
 *                        one-coefficient halo that is recomputed per tile (<= 15 % extra
 *                        coefficient reads, all L2 hits); samples are clamped, packed and
 *                        stored with 16-byte writes.
 *
 * Quirk kept on purpose: the "next LL" neighbour of the last full
 * pair of an even-sized level is the array element right after the LL quadrant, i.e. LH[0]
 * of that row / HL row 0 of that column; both are read from the coefficient array.
 */
#include "sbt.cuh"

namespace dsv {

/* sbt.c:480-503: nudge a high-band coefficient towards the local LL gradient */
DSV_D int smooth_nudge(int c, int lp, int ln, int hb, int bound)
{
    int mx = c - ln, mn = lp - c;
    if (mn > mx) {
        int t = mn;
        mn = mx;
        mx = t;
    }
    mx = imin(mx, 0);
    mn = imax(mn, 0);
    if (mx == mn) {
        return hb;
    }
    int t = rnd_shift<2>(lp - ln);
    t = iclamp(t, mx, mn);
    t = rnd_shift<1>(t - 2 * hb);
    return hb + iclamp(t, -bound, bound);
}

struct Win {
    int a, b;   /* x range [a,b) in LL coordinates of this level */
    int ha, hb; /* y range */
    int pa, pb; /* pair range x at this level */
    int qa, qb; /* pair range y */
};

#define INV_W1 66
#define INV_H1 34
#define INV_LL1_ELEMS (INV_W1 * INV_H1)
/* window storage for LL_1..LL_5: 66x34, 36x20, 20x12, 12x8, 8x6 */
#define INV_OFF2 (INV_LL1_ELEMS)
#define INV_OFF3 (INV_OFF2 + 36 * 20)
#define INV_OFF4 (INV_OFF3 + 20 * 12)
#define INV_OFF5 (INV_OFF4 + 12 * 8)
#define INV_WIN_ELEMS (INV_OFF5 + 8 * 6)
/* I frames: three more level-1 band windows + the column-pass output (2 x 66 columns x 64 rows) */
#define INV_I_EXTRA (3 * INV_LL1_ELEMS + 2 * INV_W1 * SBT_TH)
#define INV_OUT_BYTES (SBT_TW * SBT_TH)

static size_t inv_tile_smem(bool anyI)
{
    return (size_t) (INV_WIN_ELEMS + (anyI ? INV_I_EXTRA : 0)) * sizeof(int32_t) + INV_OUT_BYTES;
}

__global__ void __launch_bounds__(SBT_TILE_THREADS) sbt_inv_tile_kernel(const SbtJob *jobs, int njobs)
{
    DSV_DYN_SMEM(int32_t, sm);
    __shared__ SbtJob J;
    __shared__ Win W[SBT_NLT + 1];
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
    const bool filtered = J.plane == 0;
    const int nlt = J.nlt;

    int32_t *win[SBT_NLT + 1];
    win[1] = sm;
    win[2] = sm + INV_OFF2;
    win[3] = sm + INV_OFF3;
    win[4] = sm + INV_OFF4;
    win[5] = sm + INV_OFF5;
    int32_t *ibase = sm + INV_WIN_ELEMS; /* I frames only */
    uint8_t *outb = reinterpret_cast<uint8_t *>(sm + INV_WIN_ELEMS + (isI ? INV_I_EXTRA : 0));

    /* ---- window geometry, top-down from the sample tile ---------------------------------- */
    if (tid == 0) {
        int a = gx0, b = imin(gx0 + SBT_TW, cw), ha = gy0, hb = imin(gy0 + SBT_TH, ch);
        W[0].a = a; W[0].b = b; W[0].ha = ha; W[0].hb = hb;
        for (int l = 1; l <= nlt; l++) {
            int halo = (l == 1 && isI) ? 1 : (filtered ? 1 : 0);
            int wo = sbt_wo(cw, l), ho = sbt_wo(ch, l);
            Win w;
            w.pa = a >> 1; w.pb = ((b - 1) >> 1) + 1;
            w.qa = ha >> 1; w.qb = ((hb - 1) >> 1) + 1;
            w.a = imax(w.pa - halo, 0); w.b = imin(w.pb + halo, wo);
            w.ha = imax(w.qa - halo, 0); w.hb = imin(w.qb + halo, ho);
            W[l] = w;
            a = w.a; b = w.b; ha = w.ha; hb = w.hb;
        }
    }
    __syncthreads();

    /* ---- load LL_nlt window from the lo kernel's hand-over array -------------------------- */
    {
        const Win w = W[nlt];
        const int ww = w.b - w.a, wh = w.hb - w.ha, wo = sbt_wo(cw, nlt);
        for (int i = tid; i < ww * wh; i += SBT_TILE_THREADS) {
            int x = i % ww, y = i / ww;
            win[nlt][y * ww + x] = J.llx[(w.ha + y) * wo + w.a + x];
        }
    }
    __syncthreads();

    /* ---- Haar levels nlt .. 2 (and 1 for P frames) --------------------------------------- */
    const int last_haar = isI ? 2 : 1;
    for (int lvl = nlt; lvl >= last_haar; lvl--) {
        const Win w = W[lvl];
        const Win o = W[lvl - 1];
        const int ww = w.b - w.a;
        const int oww = o.b - o.a;
        const int ws = sbt_ws(cw, lvl), hs = sbt_ws(ch, lvl), wo = sbt_wo(cw, lvl), ho = sbt_wo(ch, lvl);
        const bool scale = lvl > 1;
        const int bound = J.hqp[lvl];
        const int32_t *src = win[lvl];
        const int npx = w.pb - w.pa, npy = w.qb - w.qa;
        for (int task = tid; task < npx * npy; task += SBT_TILE_THREADS) {
            const int jx = w.pa + task % npx, jy = w.qa + task / npx;
            const bool col2 = 2 * jx + 1 < ws, row2 = 2 * jy + 1 < hs;
            const int32_t *pc = src + (jy - w.ha) * ww + (jx - w.a);
            int LL = scale ? ll_up(pc[0]) : pc[0];
            int v00, v01 = 0, v10 = 0, v11 = 0;
            if (col2 && row2) {
                int LH = J.coef[(size_t) jy * cw + wo + jx];
                int HL = J.coef[(size_t) (ho + jy) * cw + jx];
                int HH = J.coef[(size_t) (ho + jy) * cw + wo + jx];
                if (filtered) {
                    if (jx > 0) {
                        int lp = pc[-1];
                        int ln = (jx + 1 < wo) ? pc[1] : J.coef[(size_t) jy * cw + wo];
                        if (scale) {
                            lp = ll_up(lp);
                            ln = ll_up(ln);
                        }
                        LH = smooth_nudge(LL, lp, ln, LH, bound);
                    }
                    if (jy > 0) {
                        int lp = pc[-ww];
                        int ln = (jy + 1 < ho) ? pc[ww] : J.coef[(size_t) ho * cw + jx];
                        if (scale) {
                            lp = ll_up(lp);
                            ln = ll_up(ln);
                        }
                        HL = smooth_nudge(LL, lp, ln, HL, bound);
                    }
                }
                v00 = div4_trunc(LL + LH + HL + HH);
                v01 = div4_trunc(LL - LH + HL - HH);
                v10 = div4_trunc(LL + LH - HL - HH);
                v11 = div4_trunc(LL - LH - HL + HH);
            } else if (row2) {
                int HL = J.coef[(size_t) (ho + jy) * cw + jx];
                v00 = div4_trunc(LL + HL);
                v10 = div4_trunc(LL - HL);
            } else if (col2) {
                int LH = J.coef[(size_t) jy * cw + wo + jx];
                v00 = div4_trunc(LL + LH);
                v01 = div4_trunc(LL - LH);
            } else {
                v00 = div4_trunc(LL);
            }
            const int ox = 2 * jx, oy = 2 * jy;
            if (lvl > 1) {
                int32_t *dst = win[lvl - 1];
                bool x0ok = ox >= o.a && ox < o.b, x1ok = col2 && ox + 1 >= o.a && ox + 1 < o.b;
                bool y0ok = oy >= o.ha && oy < o.hb, y1ok = row2 && oy + 1 >= o.ha && oy + 1 < o.hb;
                if (y0ok) {
                    if (x0ok) dst[(oy - o.ha) * oww + (ox - o.a)] = v00;
                    if (x1ok) dst[(oy - o.ha) * oww + (ox + 1 - o.a)] = v01;
                }
                if (y1ok) {
                    if (x0ok) dst[(oy + 1 - o.ha) * oww + (ox - o.a)] = v10;
                    if (x1ok) dst[(oy + 1 - o.ha) * oww + (ox + 1 - o.a)] = v11;
                }
            } else { /* sbc2int: +128, clamp (sbt.c:594-614); tile-local staging for wide stores */
                int lx = ox - gx0, ly = oy - gy0;
                outb[ly * SBT_TW + lx] = clamp_u8(v00 + 128);
                if (col2) outb[ly * SBT_TW + lx + 1] = clamp_u8(v01 + 128);
                if (row2) {
                    outb[(ly + 1) * SBT_TW + lx] = clamp_u8(v10 + 128);
                    if (col2) outb[(ly + 1) * SBT_TW + lx + 1] = clamp_u8(v11 + 128);
                }
            }
        }
        __syncthreads();
    }

    /* ---- level 1 of I frames: inverse B4T, columns then rows (sbt.c:253-265) --------------- */
    if (isI) {
        const Win w = W[1];
        const int ww = w.b - w.a, wh = w.hb - w.ha;
        const int wo = cw >> 1, ho = ch >> 1;
        int32_t *bLH = ibase, *bHL = ibase + INV_LL1_ELEMS, *bHH = ibase + 2 * INV_LL1_ELEMS;
        int32_t *vL = ibase + 3 * INV_LL1_ELEMS;      /* [64][66] column-pass output, low columns  */
        int32_t *vH = vL + INV_W1 * SBT_TH;           /* same for the high columns                 */
        const int32_t *bLL = win[1];
        for (int i = tid; i < ww * wh; i += SBT_TILE_THREADS) {
            int x = i % ww, y = i / ww;
            int bx = w.a + x, by = w.ha + y;
            bLH[i] = J.coef[(size_t) by * cw + wo + bx];
            bHL[i] = J.coef[(size_t) (ho + by) * cw + bx];
            bHH[i] = J.coef[(size_t) (ho + by) * cw + wo + bx];
        }
        __syncthreads();
        const int rows = W[0].hb - W[0].ha, cols = W[0].b - W[0].a;
        /* column pass: out[2m] = r8(L[m-1]+3L[m]+H[m-1]-3H[m]); out[2m+1] = r8(3L[m]+L[m+1]+3H[m]-H[m+1]) */
        for (int task = tid; task < 2 * ww * rows; task += SBT_TILE_THREADS) {
            int c = task % (2 * ww), ly = task / (2 * ww);
            int y = gy0 + ly, m = y >> 1;
            bool hcol = c >= ww;
            int x = hcol ? c - ww : c;
            const int32_t *Lc = (hcol ? bLH : bLL) + x;
            const int32_t *Hc = (hcol ? bHH : bHL) + x;
            int m0 = imax(m - 1, 0) - w.ha, m1 = m - w.ha, m2 = imin(m + 1, ho - 1) - w.ha;
            int r;
            if (!(y & 1)) {
                r = rnd_shift<3>(Lc[m0 * ww] + 3 * Lc[m1 * ww] + Hc[m0 * ww] - 3 * Hc[m1 * ww]);
            } else {
                r = rnd_shift<3>(3 * Lc[m1 * ww] + Lc[m2 * ww] + 3 * Hc[m1 * ww] - Hc[m2 * ww]);
            }
            (hcol ? vH : vL)[ly * INV_W1 + x] = r;
        }
        __syncthreads();
        /* row pass + sbc2int */
        for (int task = tid; task < cols * rows; task += SBT_TILE_THREADS) {
            int lx = task % cols, ly = task / cols;
            int x = gx0 + lx, k = x >> 1;
            const int32_t *L = vL + ly * INV_W1, *H = vH + ly * INV_W1;
            int k0 = imax(k - 1, 0) - w.a, k1 = k - w.a, k2 = imin(k + 1, wo - 1) - w.a;
            int r;
            if (!(x & 1)) {
                r = rnd_shift<3>(L[k0] + 3 * L[k1] + H[k0] - 3 * H[k1]);
            } else {
                r = rnd_shift<3>(3 * L[k1] + L[k2] + 3 * H[k1] - H[k2]);
            }
            outb[ly * SBT_TW + lx] = clamp_u8(r + 128);
        }
        __syncthreads();
    }

    /* ---- store the sample tile (only pw x ph is written, sbt.c:603-613) -------------------- */
    {
        const int rows = imin(SBT_TH, J.ph - gy0), cols = imin(SBT_TW, J.pw - gx0);
        for (int task = tid; task < SBT_TH * (SBT_TW / 16); task += SBT_TILE_THREADS) {
            int ly = task >> 3, ck = task & 7;
            if (ly >= rows || ck * 16 >= cols) {
                continue;
            }
            uint8_t *dst = J.opix + (size_t) (gy0 + ly) * J.ostride + gx0 + ck * 16;
            const uint8_t *srcb = outb + ly * SBT_TW + ck * 16;
            if (ck * 16 + 16 <= cols && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(srcb);
            } else {
                for (int e = 0; e < 16 && ck * 16 + e < cols; e++) {
                    dst[e] = srcb[e];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(SBT_LO_THREADS) sbt_inv_lo_kernel(const SbtJob *jobs)
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
    const int cw = J.cw, ch = J.ch, nlt = J.nlt;
    const bool filtered = J.plane == 0;
    int32_t *R0 = sm, *R1 = sm + sbt_wo(cw, nlt) * sbt_wo(ch, nlt);
    /* LL_k lives in R0 when (k - nlt) is even so that LL_nlt ends up in R0 */
    int32_t *A = ((J.lvls - nlt) & 1) ? R1 : R0;
    if (tid == 0) {
        A[0] = J.coef[0];
    }
    __syncthreads();
    for (int lvl = J.lvls; lvl > nlt; lvl--) {
        int32_t *B = (A == R0) ? R1 : R0;
        const int ws = sbt_ws(cw, lvl), hs = sbt_ws(ch, lvl), wo = sbt_wo(cw, lvl), ho = sbt_wo(ch, lvl);
        const int bound = J.hqp[lvl < 16 ? lvl : 15];
        for (int task = tid; task < wo * ho; task += SBT_LO_THREADS) {
            const int jx = task % wo, jy = task / wo;
            const bool col2 = 2 * jx + 1 < ws, row2 = 2 * jy + 1 < hs;
            const int32_t *pc = A + jy * wo + jx;
            int LL = ll_up(pc[0]);
            int32_t *d = B + (2 * jy) * ws + 2 * jx;
            if (col2 && row2) {
                int LH = J.coef[(size_t) jy * cw + wo + jx];
                int HL = J.coef[(size_t) (ho + jy) * cw + jx];
                int HH = J.coef[(size_t) (ho + jy) * cw + wo + jx];
                if (filtered) {
                    if (jx > 0) {
                        int lp = ll_up(pc[-1]);
                        int ln = ll_up((jx + 1 < wo) ? pc[1] : J.coef[(size_t) jy * cw + wo]);
                        LH = smooth_nudge(LL, lp, ln, LH, bound);
                    }
                    if (jy > 0) {
                        int lp = ll_up(pc[-wo]);
                        int ln = ll_up((jy + 1 < ho) ? pc[wo] : J.coef[(size_t) ho * cw + jx]);
                        HL = smooth_nudge(LL, lp, ln, HL, bound);
                    }
                }
                d[0] = div4_trunc(LL + LH + HL + HH);
                d[1] = div4_trunc(LL - LH + HL - HH);
                d[ws] = div4_trunc(LL + LH - HL - HH);
                d[ws + 1] = div4_trunc(LL - LH - HL + HH);
            } else if (row2) {
                int HL = J.coef[(size_t) (ho + jy) * cw + jx];
                d[0] = div4_trunc(LL + HL);
                d[ws] = div4_trunc(LL - HL);
            } else if (col2) {
                int LH = J.coef[(size_t) jy * cw + wo + jx];
                d[0] = div4_trunc(LL + LH);
                d[1] = div4_trunc(LL - LH);
            } else {
                d[0] = div4_trunc(LL);
            }
        }
        __syncthreads();
        A = B;
    }
    {
        const int n = sbt_wo(cw, nlt) * sbt_wo(ch, nlt);
        for (int i = tid; i < n; i += SBT_LO_THREADS) {
            J.llx[i] = A[i];
        }
    }
}

void sbt_inv_launch(const SbtJob *d_jobs, int njobs, int total_tiles, size_t lo_smem, bool any_intra, cudaStream_t st,
                    cudaEvent_t ev0, cudaEvent_t ev1)
{
    size_t tile_smem = inv_tile_smem(any_intra);
    if (lo_smem > 48 * 1024) {
        CUDA_CHECK(cudaFuncSetAttribute(sbt_inv_lo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) lo_smem));
    }
    CUDA_CHECK(cudaFuncSetAttribute(sbt_inv_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) tile_smem));
    DSV_LAUNCH(sbt_inv_lo_kernel, dim3(njobs), dim3(SBT_LO_THREADS), lo_smem, st, d_jobs);
    KERNEL_CHECK();
    if (ev0) {
        CUDA_CHECK(cudaEventRecord(ev0, st));
    }
    DSV_LAUNCH(sbt_inv_tile_kernel, dim3(total_tiles), dim3(SBT_TILE_THREADS), tile_smem, st, d_jobs, njobs);
    KERNEL_CHECK();
    if (ev1) {
        CUDA_CHECK(cudaEventRecord(ev1, st));
    }
}

} // namespace dsv
