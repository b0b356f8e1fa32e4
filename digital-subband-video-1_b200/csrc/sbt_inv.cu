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

/* same, on the two LL differences mn0 = lp - c and mx0 = c - ln (so lp - ln = mn0 + mx0) */
DSV_D int smooth_nudge_d(int mn0, int mx0, int hb, int bound)
{
    const int mx = imin(imax(mn0, mx0), 0), mn = imax(imin(mn0, mx0), 0);
    if (mx == mn) {
        return hb;
    }
    int t = iclamp(rnd_shift<2>(mn0 + mx0), mx, mn);
    t = rnd_shift<1>(t - 2 * hb);
    return hb + iclamp(t, -bound, bound);
}

struct Win {
    int a, b;   /* x range [a,b) in LL coordinates of this level */
    int ha, hb; /* y range */
    int pa, pb; /* pair range x at this level */
    int qa, qb; /* pair range y */
};

/* window storage for the LL arrays 1..5 levels above the kernel's base level: 66x34, 36x20, 20x12, 12x8, 8x6 */
#define INV_W1 66
#define INV_H1 34
#define INV_LL1_ELEMS (INV_W1 * INV_H1)
#define INV_OFF2 (INV_LL1_ELEMS)
#define INV_OFF3 (INV_OFF2 + 36 * 20)
#define INV_OFF4 (INV_OFF3 + 20 * 12)
#define INV_OFF5 (INV_OFF4 + 12 * 8)
#define INV_WIN_ELEMS (INV_OFF5 + 8 * 6)
/* I frames: column-pass output of the inverse B4T, [64 rows][low 72 | high 72]: up to 66 window columns per half,
 * shifted so that the tile's own 64 start at a multiple of 4 */
#define INV_VHALF 72
#define INV_VSTRIDE (2 * INV_VHALF)
#define INV_I_EXTRA (SBT_TH * INV_VSTRIDE)

static size_t inv_tile_smem(bool anyI)
{
    const size_t generic = (size_t) INV_OFF3 + (anyI ? INV_I_EXTRA : 0);
    const size_t fast = (size_t) 72 * 36 + 36 * 20; /* INV_FAST_ELEMS: interior tiles of P pictures */
    return (generic > fast ? generic : fast) * sizeof(int32_t);
}

DSV_D unsigned pack4_u8(int a, int b, int c, int d) { return pack_u8x4(a + 128, b + 128, c + 128, d + 128); }

/* store 8 reconstructed samples of one output row (sbc2int, sbt.c:594-614): only pw x ph is written */
DSV_D void store_row8(const SbtJob &J, int oy, int ox, const int *v)
{
    if (oy >= J.ph || ox >= J.pw) {
        return;
    }
    uint8_t *dst = J.opix + (size_t) oy * J.ostride + ox;
    if (J.addp) {
        const uint8_t *ap = J.addp + (size_t) oy * J.addstride + ox;
        if (ox + 8 <= J.pw && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(ap)) & 7) == 0) {
            const uint2 p = *reinterpret_cast<const uint2 *>(ap);
            int o[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                o[e] = clamp_u8(v[e] + 128) + byte_of(e < 4 ? p.x : p.y, e & 3) - 128;
            }
            *reinterpret_cast<uint2 *>(dst) = make_uint2(pack_u8x4(o[0], o[1], o[2], o[3]), pack_u8x4(o[4], o[5], o[6], o[7]));
        } else {
#pragma unroll
            for (int e = 0; e < 8; e++) {
                if (ox + e < J.pw) {
                    dst[e] = clamp_u8(clamp_u8(v[e] + 128) + ap[e] - 128);
                }
            }
        }
        return;
    }
    if (ox + 8 <= J.pw && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
        *reinterpret_cast<uint2 *>(dst) = make_uint2(pack4_u8(v[0], v[1], v[2], v[3]), pack4_u8(v[4], v[5], v[6], v[7]));
    } else {
#pragma unroll
        for (int e = 0; e < 8; e++) {
            if (ox + e < J.pw) {
                dst[e] = clamp_u8(v[e] + 128);
            }
        }
    }
}

/* 8 samples of one output row whose residual is exactly zero: 128, or the prediction itself */
DSV_D void store_row8_zero(const SbtJob &J, int oy, int ox)
{
    if (oy >= J.ph || ox >= J.pw) {
        return;
    }
    uint8_t *dst = J.opix + (size_t) oy * J.ostride + ox;
    const uint8_t *ap = J.addp ? J.addp + (size_t) oy * J.addstride + ox : nullptr;
    if (ox + 8 <= J.pw && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(ap)) & 7) == 0) {
        *reinterpret_cast<uint2 *>(dst) = ap ? *reinterpret_cast<const uint2 *>(ap) : make_uint2(0x80808080u, 0x80808080u);
    } else {
        for (int e = 0; e < 8 && ox + e < J.pw; e++) {
            dst[e] = ap ? ap[e] : (uint8_t) 128;
        }
    }
}

/* smooth_nudge_d for a zero high-band coefficient: non-zero only where LL is strictly monotonic across the pair */
DSV_D int smooth_nudge_z(int mn0, int mx0, int bound)
{
    const int mx = imin(imax(mn0, mx0), 0), mn = imax(imin(mn0, mx0), 0);
    const int t = iclamp(rnd_shift<2>(mn0 + mx0), mx, mn); /* mx == mn == 0 clamps t to 0, and 0 nudges to 0 */
    return iclamp(rnd_shift<1>(t), -bound, bound);
}

/*
 * Level 1 of a P picture for one tile: a thread owns 4 adjacent pairs = 8 x 2 output samples (cw, ch are even, so
 * every pair is complete; no LL scaling at level 1 of P frames).  EMPTY: the tile's three level-1 band blocks are all
 * zero (tile flag): no coefficient is loaded, and a warp whose whole LL neighbourhood is zero as well stores the
 * zero residual (128, or the prediction) without any arithmetic.
 */
template <bool EMPTY> DSV_D void inv_level1_p(const SbtJob &J, const Win &w, const int32_t *LLw, int tx, int ty, int tid)
{
    const int cw = J.cw, ch = J.ch;
    const bool filtered = J.plane == 0;
    const int ww = w.b - w.a;
    const int wo = cw >> 1, ho = ch >> 1;
    const int bound = J.hqp[1];
#pragma unroll 1
    for (int g = tid; g < (SBT_TW / 8) * (SBT_TH / 2); g += SBT_TILE_THREADS) {
        const int qrow = g >> 4, qc = (g & 15) * 4;
        const int jy = ty * (SBT_TH / 2) + qrow, jx = tx * (SBT_TW / 2) + qc;
        const bool active = jy < ho && jx < wo;
        if (!EMPTY && !active) {
            continue;
        }
        const int32_t *pc = LLw + (jy - w.ha) * ww + (jx - w.a);
        int r0[8], r1[8];
        int Lc[4];
        if (EMPTY) {
            /* the warp's trip count is uniform (512 groups, 256 threads); inactive lanes vote "zero" */
            int z = 0;
            if (active) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    Lc[i] = pc[i];
                    z |= Lc[i];
                }
                if (filtered) {
                    z |= pc[4] | (jx > 0 ? pc[-1] : 0);
                    if (jy > 0) {
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            z |= pc[i - ww] | pc[i + ww];
                        }
                    }
                }
            }
            if (__all_sync(0xffffffffu, z == 0)) {
                if (active) {
                    store_row8_zero(J, 2 * jy, 2 * jx);
                    store_row8_zero(J, 2 * jy + 1, 2 * jx);
                }
                continue;
            }
            if (!active) {
                continue;
            }
            int h[4] = {0, 0, 0, 0}, v[4] = {0, 0, 0, 0};
            if (filtered) {
                int d[5];
                d[0] = (jx > 0 ? pc[-1] : Lc[0]) - Lc[0];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    d[i + 1] = Lc[i] - Lc[i + 1];
                }
                d[4] = Lc[3] - pc[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (i > 0 || jx > 0) {
                        h[i] = smooth_nudge_z(d[i], d[i + 1], bound);
                    }
                }
                if (jy > 0) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        v[i] = smooth_nudge_z(pc[i - ww] - Lc[i], Lc[i] - pc[i + ww], bound);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int sa = Lc[i] + h[i], sb = Lc[i] - h[i];
                r0[2 * i] = div4_trunc(sa + v[i]);
                r0[2 * i + 1] = div4_trunc(sb + v[i]);
                r1[2 * i] = div4_trunc(sa - v[i]);
                r1[2 * i + 1] = div4_trunc(sb - v[i]);
            }
        } else {
            const int nv = imin(4, wo - jx);
            const int32_t *pLH = J.coef + (size_t) jy * cw + wo + jx;
            const int32_t *pHL = J.coef + (size_t) (ho + jy) * cw + jx;
            const int32_t *pHH = pHL + wo;
            int LH[4] = {0, 0, 0, 0}, HL[4] = {0, 0, 0, 0}, HH[4] = {0, 0, 0, 0};
            if (nv == 4 && ((reinterpret_cast<uintptr_t>(pLH) | reinterpret_cast<uintptr_t>(pHL) | reinterpret_cast<uintptr_t>(pHH)) & 15) == 0) {
                const int4 a = *reinterpret_cast<const int4 *>(pLH), b = *reinterpret_cast<const int4 *>(pHL), c = *reinterpret_cast<const int4 *>(pHH);
                LH[0] = a.x; LH[1] = a.y; LH[2] = a.z; LH[3] = a.w;
                HL[0] = b.x; HL[1] = b.y; HL[2] = b.z; HL[3] = b.w;
                HH[0] = c.x; HH[1] = c.y; HH[2] = c.z; HH[3] = c.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (i < nv) {
                        LH[i] = pLH[i];
                        HL[i] = pHL[i];
                        HH[i] = pHH[i];
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                Lc[i] = pc[i];
            }
            if (filtered) {
                /* horizontal nudges share the LL differences of neighbouring pairs; the window already holds the
                 * reference's "next LL" values past the quadrant, only the first column / row of the plane is special */
                int d[5];
                d[0] = (jx > 0 ? pc[-1] : Lc[0]) - Lc[0];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    d[i + 1] = Lc[i] - Lc[i + 1];
                }
                d[4] = Lc[3] - pc[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (i > 0 || jx > 0) {
                        LH[i] = smooth_nudge_d(d[i], d[i + 1], LH[i], bound);
                    }
                }
                if (jy > 0) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        HL[i] = smooth_nudge_d(pc[i - ww] - Lc[i], Lc[i] - pc[i + ww], HL[i], bound);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int LL = Lc[i];
                const int sa = LL + LH[i], sb = LL - LH[i], sc = HL[i] + HH[i], sd = HL[i] - HH[i];
                r0[2 * i] = div4_trunc(sa + sc);
                r0[2 * i + 1] = div4_trunc(sb + sd);
                r1[2 * i] = div4_trunc(sa - sc);
                r1[2 * i + 1] = div4_trunc(sb - sd);
            }
        }
        store_row8(J, 2 * jy, 2 * jx, r0);
        store_row8(J, 2 * jy + 1, 2 * jx, r1);
    }
}

/* ---- interior tiles of P pictures: compile-time geometry ---------------------------------------------------------
 * A tile that has a full tile on every side needs none of the edge rules: every pair is complete, every nudge has
 * both neighbours, every 8-sample row store is in bounds and aligned.  The windows then have fixed shapes, so the
 * index arithmetic is constant-folded and the bounds tests (a third of the generic path's instructions) disappear.
 *   LL_2 window  (32 + 4H) x (16 + 4H) at pitch FW2, element (jx2, jy2) of the plane at column jx2 - 32 tx + 2H, ...
 *   LL_1 window  pitch FW1: element jx1 at column jx1 - 64 tx + 4 (the four values a thread owns are 16-byte aligned),
 *                row jy1 - 32 ty + 2
 * H = 1 for the filtered (luma) inverse, whose nudges look one LL value to each side. */
#define FW1 72
#define FH1 36
#define FW2 36
#define FH2 20
#define INV_FAST_ELEMS (FW1 * FH1 + FW2 * FH2)

DSV_D int clamp_s8(int v) { return imax(imin(v, 127), -128); }

/* 8 reconstructed samples: clamp_u8(v + 128), or with a prediction clamp_u8(clamp_u8(v + 128) + p - 128)
 * == clamp_u8(clamp(v, -128, 127) + p): the byte of p is added by a dp4a whose other operand selects it */
template <bool ADDP> DSV_D void store8_fast(uint8_t *dst, const uint8_t *ap, const int *v)
{
    if (ADDP) {
        const uint2 p = *reinterpret_cast<const uint2 *>(ap);
        int o[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            o[e] = (int) __dp4a(e < 4 ? p.x : p.y, 1u << (8 * (e & 3)), (unsigned) clamp_s8(v[e]));
        }
        *reinterpret_cast<uint2 *>(dst) = make_uint2(pack_u8x4(o[0], o[1], o[2], o[3]), pack_u8x4(o[4], o[5], o[6], o[7]));
    } else {
        *reinterpret_cast<uint2 *>(dst) = make_uint2(pack4_u8(v[0], v[1], v[2], v[3]), pack4_u8(v[4], v[5], v[6], v[7]));
    }
}

template <bool FILT, bool ADDP, bool EMPTY1>
DSV_D void inv_tile_fast_p(const SbtJob &J, int tx, int ty, int f2, int32_t *sm, int tid)
{
    constexpr int H = FILT ? 1 : 0;
    const int cw = J.cw;
    const int wo1 = cw >> 1, ho1 = J.ch >> 1, wo2 = sbt_wo(cw, 2), ho2 = sbt_wo(J.ch, 2);
    int32_t *win1 = sm, *win2 = sm + FW1 * FH1;
    /* level 2 works on 34 x 18 pairs (the tile's 32 x 16 and the ring level 1 looks at), one pair per thread and
     * round.  The band coefficients of all of a thread's pairs are requested before anything else: their latency
     * overlaps the LL_2 window fetch and the barrier behind it. */
    constexpr int NPX = 32 + 2 * H, NPY = 16 + 2 * H;
    constexpr int NIT = (NPX * NPY + SBT_TILE_THREADS - 1) / SBT_TILE_THREADS;
    int cLH[NIT], cHL[NIT], cHH[NIT];
    {
        const int32_t *bLH = J.coef + (size_t) (ty * 16 - H) * cw + wo2 + tx * 32 - H;
        const int32_t *bHL = J.coef + (size_t) (ho2 + ty * 16 - H) * cw + tx * 32 - H;
#pragma unroll
        for (int it = 0; it < NIT; it++) {
            const int p = tid + it * SBT_TILE_THREADS;
            cLH[it] = cHL[it] = cHH[it] = 0;
            if (f2 && p < NPX * NPY) {
                const int py = p / NPX, px = p - py * NPX;
                const size_t o = (size_t) py * cw + px;
                cLH[it] = bLH[o];
                cHL[it] = bHL[o];
                cHH[it] = bHL[o + wo2];
            }
        }
    }
    {
        const int32_t *ll2 = J.llx + J.ll2_off + (size_t) (ty * 16 - 2 * H) * wo2 + tx * 32 - 2 * H;
        constexpr int W2 = 32 + 4 * H, H2 = 16 + 4 * H;
        for (int i = tid; i < W2 * H2; i += SBT_TILE_THREADS) {
            const int y = i / W2, x = i - y * W2;
            win2[y * FW2 + x] = ll_up(ll2[(size_t) y * wo2 + x]); /* LL scaled by 5/4 (sbt.c:20-22) */
        }
    }
    __syncthreads();
    {
        const int bound = J.hqp[2];
#pragma unroll
        for (int it = 0; it < NIT; it++) {
            const int p = tid + it * SBT_TILE_THREADS;
            if (p >= NPX * NPY) {
                break;
            }
            const int py = p / NPX, px = p - py * NPX;
            const int32_t *pc = win2 + (py + H) * FW2 + px + H;
            const int LL = pc[0];
            int LH = cLH[it], HL = cHL[it];
            const int HH = cHH[it];
            if (FILT) {
                LH = smooth_nudge_d(pc[-1] - LL, LL - pc[1], LH, bound);
                HL = smooth_nudge_d(pc[-FW2] - LL, LL - pc[FW2], HL, bound);
            }
            const int sa = LL + LH, sb = LL - LH, sc = HL + HH, sd = HL - HH;
            int32_t *dst = win1 + (2 * py - 2 * H + 2) * FW1 + 2 * px - 2 * H + 4;
            *reinterpret_cast<int2 *>(dst) = make_int2(div4_trunc(sa + sc), div4_trunc(sb + sd));
            *reinterpret_cast<int2 *>(dst + FW1) = make_int2(div4_trunc(sa - sc), div4_trunc(sb - sd));
        }
    }
    __syncthreads();
    /* level 1: a thread owns 4 adjacent pairs = 8 x 2 samples, twice */
    const int bound = J.hqp[1];
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
        const int g = tid + it * SBT_TILE_THREADS;
        const int qrow = g >> 4, qc = (g & 15) * 4;
        const int jy = ty * (SBT_TH / 2) + qrow, jx = tx * (SBT_TW / 2) + qc;
        const int32_t *pc = win1 + (qrow + 2) * FW1 + qc + 4;
        const int4 c4 = *reinterpret_cast<const int4 *>(pc);
        const int Lc[4] = {c4.x, c4.y, c4.z, c4.w};
        int up[4] = {0, 0, 0, 0}, dn[4] = {0, 0, 0, 0}, lf = 0, rt = 0;
        if (FILT) {
            const int4 u4 = *reinterpret_cast<const int4 *>(pc - FW1), d4 = *reinterpret_cast<const int4 *>(pc + FW1);
            up[0] = u4.x; up[1] = u4.y; up[2] = u4.z; up[3] = u4.w;
            dn[0] = d4.x; dn[1] = d4.y; dn[2] = d4.z; dn[3] = d4.w;
            lf = pc[-1];
            rt = pc[4];
        }
        uint8_t *dst = J.opix + (size_t) (2 * jy) * J.ostride + 2 * jx;
        const uint8_t *ap = ADDP ? J.addp + (size_t) (2 * jy) * J.addstride + 2 * jx : nullptr;
        int LH[4] = {0, 0, 0, 0}, HL[4] = {0, 0, 0, 0}, HH[4] = {0, 0, 0, 0};
        if (EMPTY1) {
            /* the tile's level-1 blocks are empty: nothing to load, and a warp whose LL neighbourhood is zero too
             * has a zero residual */
            int z = Lc[0] | Lc[1] | Lc[2] | Lc[3];
            if (FILT) {
                z |= lf | rt | up[0] | up[1] | up[2] | up[3] | dn[0] | dn[1] | dn[2] | dn[3];
            }
            if (__all_sync(0xffffffffu, z == 0)) {
                if (ADDP) {
                    *reinterpret_cast<uint2 *>(dst) = *reinterpret_cast<const uint2 *>(ap);
                    *reinterpret_cast<uint2 *>(dst + J.ostride) = *reinterpret_cast<const uint2 *>(ap + J.addstride);
                } else {
                    *reinterpret_cast<uint2 *>(dst) = make_uint2(0x80808080u, 0x80808080u);
                    *reinterpret_cast<uint2 *>(dst + J.ostride) = make_uint2(0x80808080u, 0x80808080u);
                }
                continue;
            }
            if (FILT) {
                const int d[5] = {lf - Lc[0], Lc[0] - Lc[1], Lc[1] - Lc[2], Lc[2] - Lc[3], Lc[3] - rt};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    LH[i] = smooth_nudge_z(d[i], d[i + 1], bound);
                    HL[i] = smooth_nudge_z(up[i] - Lc[i], Lc[i] - dn[i], bound);
                }
            }
        } else {
            const int32_t *pLH = J.coef + (size_t) jy * cw + wo1 + jx;
            const int32_t *pHL = J.coef + (size_t) (ho1 + jy) * cw + jx;
            const int4 a = *reinterpret_cast<const int4 *>(pLH), b = *reinterpret_cast<const int4 *>(pHL),
                       c = *reinterpret_cast<const int4 *>(pHL + wo1);
            LH[0] = a.x; LH[1] = a.y; LH[2] = a.z; LH[3] = a.w;
            HL[0] = b.x; HL[1] = b.y; HL[2] = b.z; HL[3] = b.w;
            HH[0] = c.x; HH[1] = c.y; HH[2] = c.z; HH[3] = c.w;
            if (FILT) {
                const int d[5] = {lf - Lc[0], Lc[0] - Lc[1], Lc[1] - Lc[2], Lc[2] - Lc[3], Lc[3] - rt};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    LH[i] = smooth_nudge_d(d[i], d[i + 1], LH[i], bound);
                    HL[i] = smooth_nudge_d(up[i] - Lc[i], Lc[i] - dn[i], HL[i], bound);
                }
            }
        }
        int r0[8], r1[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int sa = Lc[i] + LH[i], sb = Lc[i] - LH[i], sc = HL[i] + HH[i], sd = HL[i] - HH[i];
            r0[2 * i] = div4_trunc(sa + sc);
            r0[2 * i + 1] = div4_trunc(sb + sd);
            r1[2 * i] = div4_trunc(sa - sc);
            r1[2 * i + 1] = div4_trunc(sb - sd);
        }
        store8_fast<ADDP>(dst, ap, r0);
        store8_fast<ADDP>(dst + J.ostride, ADDP ? ap + J.addstride : nullptr, r1);
    }
}

/* may tile (tx, ty) take the fast path?  (uniform; thread 0) */
DSV_D bool inv_tile_is_interior(const SbtJob &J, int tx, int ty)
{
    if (!J.isP || tx < 1 || ty < 1) {
        return false;
    }
    const int cw = J.cw, ch = J.ch;
    const int wo1 = cw >> 1, ho1 = ch >> 1, wo2 = sbt_wo(cw, 2), ho2 = sbt_wo(ch, 2);
    /* the windows (tile + ring) stay inside LL_1 / LL_2 proper, away from the planes' last pairs and their quirks */
    if (tx * 64 + 66 > wo1 || ty * 32 + 34 > ho1 || tx * 32 + 34 > wo2 || ty * 16 + 18 > ho2) {
        return false;
    }
    if ((tx + 1) * SBT_TW > J.pw || (ty + 1) * SBT_TH > J.ph) {
        return false;
    }
    /* 16-byte coefficient loads, 8-byte sample stores */
    uintptr_t a = reinterpret_cast<uintptr_t>(J.opix) | (uintptr_t) J.ostride;
    if (J.addp) {
        a |= reinterpret_cast<uintptr_t>(J.addp) | (uintptr_t) J.addstride;
    }
    return (a & 7) == 0 && (reinterpret_cast<uintptr_t>(J.coef) & 15) == 0 && (cw & 7) == 0;
}

template <bool MID> DSV_D int inv_load_job(SbtJob *sJ, const SbtJob *jobs, const SbtDims &dims)
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

/* window geometry, top-down from the tile of level-`base` outputs (base 0 = samples) to level `top` */
DSV_D void inv_windows(const SbtJob &J, Win *W, int base, int top, int gx0, int gy0)
{
    const bool isI = !J.isP, filtered = J.plane == 0;
    int a = gx0, b = imin(gx0 + SBT_TW, base ? sbt_wo(J.cw, base) : J.cw);
    int ha = gy0, hb = imin(gy0 + SBT_TH, base ? sbt_wo(J.ch, base) : J.ch);
    W[base].a = a; W[base].b = b; W[base].ha = ha; W[base].hb = hb;
    for (int l = base + 1; l <= top; l++) {
        int halo = (l == 1 && isI) ? 1 : (filtered ? 1 : 0);
        int wo = sbt_wo(J.cw, l), ho = sbt_wo(J.ch, l);
        Win w;
        w.pa = a >> 1; w.pb = ((b - 1) >> 1) + 1;
        w.qa = ha >> 1; w.qb = ((hb - 1) >> 1) + 1;
        /* level 1 of a filtered P plane: one extra column / row past the LL quadrant holds the value the
         * reference reads as "next LL" of the last pair, so the streaming loop has no
         * edge cases */
        const int ext = (filtered && ((l == 1 && !isI) || (l == 2 && base == 0))) ? 1 : 0;
        w.a = imax(w.pa - halo, 0); w.b = imin(w.pb + halo, wo + ext);
        w.ha = imax(w.qa - halo, 0); w.hb = imin(w.qb + halo, ho + ext);
        W[l] = w;
        a = w.a; b = imin(w.b, wo); ha = w.ha; hb = imin(w.hb, ho); /* the extra column / row is not produced by level l+1 */
    }
}

/*
 * One inverse Haar level, one pair per thread (inv / inv_simple, sbt.c:352-574): LL window `src` of level lvl
 * -> values of level lvl-1, written either into the next window (dst_win, geometry o) or, when dst_plane is
 * given, into a dense global LL plane of row pitch dst_pitch restricted to the rectangle o.
 */
DSV_D void inv_haar_level(const SbtJob &J, int lvl, const Win &w, const Win &o, const int32_t *src, int32_t *dst_win,
                          int32_t *dst_plane, int dst_pitch)
{
    const int cw = J.cw, ch = J.ch;
    const bool filtered = J.plane == 0;
    const int ww = w.b - w.a;
    const int oww = o.b - o.a;
    const int ws = sbt_ws(cw, lvl), hs = sbt_ws(ch, lvl), wo = sbt_wo(cw, lvl), ho = sbt_wo(ch, lvl);
    const bool scale = lvl > 1; /* LL scaled by 5/4 at every level this function is used for (sbt.c:20-22) */
    const int bound = J.hqp[lvl];
    const int npx = w.pb - w.pa, npy = w.qb - w.qa;
    for (int task = threadIdx.x; task < npx * npy; task += blockDim.x) {
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
        const bool x0ok = ox >= o.a && ox < o.b, x1ok = col2 && ox + 1 >= o.a && ox + 1 < o.b;
        const bool y0ok = oy >= o.ha && oy < o.hb, y1ok = row2 && oy + 1 >= o.ha && oy + 1 < o.hb;
        if (dst_plane) {
            if (y0ok) {
                if (x0ok) dst_plane[(size_t) oy * dst_pitch + ox] = v00;
                if (x1ok) dst_plane[(size_t) oy * dst_pitch + ox + 1] = v01;
            }
            if (y1ok) {
                if (x0ok) dst_plane[(size_t) (oy + 1) * dst_pitch + ox] = v10;
                if (x1ok) dst_plane[(size_t) (oy + 1) * dst_pitch + ox + 1] = v11;
            }
        } else {
            if (y0ok) {
                if (x0ok) dst_win[(oy - o.ha) * oww + (ox - o.a)] = v00;
                if (x1ok) dst_win[(oy - o.ha) * oww + (ox + 1 - o.a)] = v01;
            }
            if (y1ok) {
                if (x0ok) dst_win[(oy + 1 - o.ha) * oww + (ox - o.a)] = v10;
                if (x1ok) dst_win[(oy + 1 - o.ha) * oww + (ox + 1 - o.a)] = v11;
            }
        }
    }
}

/*
 * Mid kernel: levels nlt..3 for one 128x64 block of LL_2.  LL_nlt window from the lo kernel's hand-over array,
 * result into the dense LL_2 hand-over plane.
 */
__global__ void __launch_bounds__(SBT_TILE_THREADS) sbt_inv_mid_kernel(const SbtJob *jobs, const SbtDims dims)
{
    __shared__ int32_t sm[INV_WIN_ELEMS];
    __shared__ SbtJob J;
    __shared__ Win W[SBT_NLT + 1];
    const int tid = threadIdx.x;
    const int t = inv_load_job<true>(&J, jobs, dims);
    const int ty = (int) fastdiv((unsigned) t, J.mtiles_x_fd), tx = t - ty * J.mtiles_x;
    const int nlt = J.nlt;
    int32_t *win[SBT_NLT + 1];
    win[SBT_HI + 1] = sm;
    win[SBT_HI + 2] = sm + INV_OFF2;
    win[SBT_HI + 3] = sm + INV_OFF3;
    if (tid == 0) {
        inv_windows(J, W, SBT_HI, nlt, tx * SBT_TW, ty * SBT_TH);
    }
    __syncthreads();
    {
        const Win w = W[nlt];
        const int ww = w.b - w.a, wh = w.hb - w.ha, wo = sbt_wo(J.cw, nlt);
        for (int i = tid; i < ww * wh; i += SBT_TILE_THREADS) {
            int x = i % ww, y = i / ww;
            win[nlt][y * ww + x] = J.llx[(w.ha + y) * wo + w.a + x];
        }
    }
    __syncthreads();
    for (int lvl = nlt; lvl > SBT_HI; lvl--) {
        const bool last = lvl == SBT_HI + 1;
        inv_haar_level(J, lvl, W[lvl], W[lvl - 1], win[lvl], last ? nullptr : win[lvl - 1], last ? J.llx + J.ll2_off : nullptr,
                       sbt_wo(J.cw, SBT_HI));
        __syncthreads();
    }
}

/*
 * Streaming kernel: levels 2 and 1 of one 128x64-sample tile.  Level 2 (with its one-coefficient halo) one pair
 * per thread from the LL_2 hand-over plane; level 1 with a thread owning 4 adjacent pairs = 8x2 output samples,
 * 16-byte coefficient loads and 8-byte sample stores.
 */
#ifndef SBT_INV_MINB
#define SBT_INV_MINB 6 /* CTAs per SM, measured per 32 HD pictures (decoder / encoder with fused prediction add):
                          6 -> 181 / 213 us, 8 -> 183 / 211, 5 -> 183 / 219, 4 -> 202 / 245 */
#endif
#ifndef SBT_INV_MINB_I
#define SBT_INV_MINB_I 4 /* tiles of I pictures (inverse B4T through a 37 KB column buffer): 64 registers instead of 40
                            take the 64-lane launch from 934 to 652 us; the P path loses 20 % at that occupancy */
#endif
/* The body is compiled twice: INTRA = false with the P pictures' occupancy, true with the I pictures'.  A launch whose
 * jobs are all of one kind (the usual lock-step case) runs one of them; a mixed launch runs both over the same grid and
 * every CTA leaves at once when its tile belongs to the other kind. */
template <bool INTRA> DSV_D void sbt_inv_tile_body(const SbtJob *jobs, const SbtDims &dims)
{
    DSV_DYN_SMEM(int32_t, sm);
    __shared__ SbtJob J;
    __shared__ Win W[SBT_HI + 1];
    const int tid = threadIdx.x;
    {
        int tile;
        if ((jobs[sbt_locate<false>(jobs, dims, (int) blockIdx.x, &tile)].isP == 0) != INTRA) {
            return;
        }
    }
    const int t = inv_load_job<false>(&J, jobs, dims);

    const int ty = (int) fastdiv((unsigned) t, J.tiles_x_fd), tx = t - ty * J.tiles_x;
    const int gx0 = tx * SBT_TW, gy0 = ty * SBT_TH;
    const int cw = J.cw, ch = J.ch;
    constexpr bool isI = INTRA; /* checked against the job above */
    const bool filtered = J.plane == 0;

    int32_t *win1 = sm, *win2 = sm + INV_OFF2;
    int32_t *ibase = sm + INV_OFF3; /* I frames only */
    __shared__ int s_f1, s_f2; /* tile flags (sbt.cuh): level-1 blocks of this tile, level-2 blocks of the tile and its halo */
    __shared__ int s_fast;
    if (tid == 0) {
        s_fast = (!INTRA && inv_tile_is_interior(J, tx, ty)) ? 1 : 0;
        if (!s_fast) {
            inv_windows(J, W, 0, SBT_HI, gx0, gy0);
        }
        int f1 = 1, f2 = 1;
        if (J.tflags) {
            f1 = J.tflags[t] & 1;
            f2 = 0;
            for (int ny = imax(ty - 1, 0); ny <= imin(ty + 1, J.tiles_y - 1); ny++) {
                for (int nx = imax(tx - 1, 0); nx <= imin(tx + 1, J.tiles_x - 1); nx++) {
                    f2 |= J.tflags[ny * J.tiles_x + nx] & 2;
                }
            }
        }
        s_f1 = f1;
        s_f2 = f2;
    }
    __syncthreads();
    if (!INTRA && s_fast) {
        const int sel = (filtered ? 4 : 0) | (J.addp ? 2 : 0) | (s_f1 ? 0 : 1);
        switch (sel) {
            case 0: inv_tile_fast_p<false, false, false>(J, tx, ty, s_f2, sm, tid); break;
            case 1: inv_tile_fast_p<false, false, true>(J, tx, ty, s_f2, sm, tid); break;
            case 2: inv_tile_fast_p<false, true, false>(J, tx, ty, s_f2, sm, tid); break;
            case 3: inv_tile_fast_p<false, true, true>(J, tx, ty, s_f2, sm, tid); break;
            case 4: inv_tile_fast_p<true, false, false>(J, tx, ty, s_f2, sm, tid); break;
            case 5: inv_tile_fast_p<true, false, true>(J, tx, ty, s_f2, sm, tid); break;
            case 6: inv_tile_fast_p<true, true, false>(J, tx, ty, s_f2, sm, tid); break;
            default: inv_tile_fast_p<true, true, true>(J, tx, ty, s_f2, sm, tid); break;
        }
        return;
    }
    {
        /* LL_2 window, already scaled by 5/4 (sbt.c:20-22) and, for filtered planes, extended by the column / row
         * the reference reads as "next LL" of the last pair (Appendix B-2) */
        const Win w = W[2];
        const int ww = w.b - w.a, wh = w.hb - w.ha, wo = sbt_wo(cw, 2), ho = sbt_wo(ch, 2);
        const int32_t *ll2 = J.llx + J.ll2_off;
        const int wmag = magic20(ww);
        for (int i = tid; i < ww * wh; i += SBT_TILE_THREADS) {
            const int y = (i * wmag) >> 20, x = i - y * ww; /* i / ww for i < 2^10, ww <= 36 */
            const int gx = w.a + x, gy = w.ha + y;
            int v = 0;
            if (gx < wo && gy < ho) {
                v = ll2[(size_t) gy * wo + gx];
            } else if (gx == wo && gy < ho) {
                v = J.coef[(size_t) gy * cw + wo];
            } else if (gy == ho && gx < wo) {
                v = J.coef[(size_t) ho * cw + gx];
            }
            win2[i] = ll_up(v);
        }
    }
    __syncthreads();
    {
        /* level 2 (Haar with LL scaling, both frame types): one pair per thread into the LL_1 window */
        const Win w = W[2], o = W[1];
        const int ww = w.b - w.a, oww = o.b - o.a;
        const int ws = sbt_ws(cw, 2), hs = sbt_ws(ch, 2), wo = sbt_wo(cw, 2), ho = sbt_wo(ch, 2);
        const int bound = J.hqp[2];
        const int f2 = s_f2;
        const int npx = w.pb - w.pa, npy = w.qb - w.qa;
        const int mag = magic20(npx);
        const int o_b = imin(o.b, ws), o_hb = imin(o.hb, hs); /* level 2 produces LL_1 proper only */
        for (int task = tid; task < npx * npy; task += SBT_TILE_THREADS) {
            const int ty2 = (task * mag) >> 20, tx2 = task - ty2 * npx;
            const int jx = w.pa + tx2, jy = w.qa + ty2;
            const bool col2 = 2 * jx + 1 < ws, row2 = 2 * jy + 1 < hs;
            const int32_t *pc = win2 + (jy - w.ha) * ww + (jx - w.a);
            const int LL = pc[0];
            int v00, v01 = 0, v10 = 0, v11 = 0;
            if (col2 && row2) {
                int LH = 0, HL = 0, HH = 0;
                if (f2) {
                    const int32_t *pb = J.coef + (size_t) jy * cw + wo + jx;
                    LH = pb[0];
                    const int32_t *pb2 = J.coef + (size_t) (ho + jy) * cw + jx;
                    HL = pb2[0];
                    HH = pb2[wo];
                }
                if (filtered) {
                    if (jx > 0) {
                        LH = smooth_nudge_d(pc[-1] - LL, LL - pc[1], LH, bound);
                    }
                    if (jy > 0) {
                        HL = smooth_nudge_d(pc[-ww] - LL, LL - pc[ww], HL, bound);
                    }
                }
                const int sa = LL + LH, sb = LL - LH, sc = HL + HH, sd = HL - HH;
                v00 = div4_trunc(sa + sc);
                v01 = div4_trunc(sb + sd);
                v10 = div4_trunc(sa - sc);
                v11 = div4_trunc(sb - sd);
            } else if (row2) {
                const int HL = f2 ? J.coef[(size_t) (ho + jy) * cw + jx] : 0;
                v00 = div4_trunc(LL + HL);
                v10 = div4_trunc(LL - HL);
            } else if (col2) {
                const int LH = f2 ? J.coef[(size_t) jy * cw + wo + jx] : 0;
                v00 = div4_trunc(LL + LH);
                v01 = div4_trunc(LL - LH);
            } else {
                v00 = div4_trunc(LL);
            }
            const int ox = 2 * jx, oy = 2 * jy;
            const bool x0ok = ox >= o.a && ox < o_b, x1ok = col2 && ox + 1 >= o.a && ox + 1 < o_b;
            const bool y0ok = oy >= o.ha && oy < o_hb, y1ok = row2 && oy + 1 >= o.ha && oy + 1 < o_hb;
            int32_t *dst = win1 + (oy - o.ha) * oww + (ox - o.a);
            if (y0ok) {
                if (x0ok) dst[0] = v00;
                if (x1ok) dst[1] = v01;
            }
            if (y1ok) {
                if (x0ok) dst[oww] = v10;
                if (x1ok) dst[oww + 1] = v11;
            }
        }
    }
    __syncthreads();
    if (!isI && filtered) {
        const Win w = W[1];
        const int ww = w.b - w.a, wo = cw >> 1, ho = ch >> 1;
        if (w.b == wo + 1) { /* right plane edge: "next LL" of the last pair of row y is LH[0] of that row */
            for (int y = w.ha + tid; y < imin(w.hb, ho); y += SBT_TILE_THREADS) {
                win1[(y - w.ha) * ww + (wo - w.a)] = J.coef[(size_t) y * cw + wo];
            }
        }
        if (w.hb == ho + 1) { /* bottom plane edge: HL row 0 */
            for (int x = w.a + tid; x < imin(w.b, wo); x += SBT_TILE_THREADS) {
                win1[(ho - w.ha) * ww + (x - w.a)] = J.coef[(size_t) ho * cw + x];
            }
        }
    }
    __syncthreads();

    /* ---- level 1 of P frames (inv_level1_p) ----------------------------------------------------------- */
    if (!isI) {
        if (s_f1) {
            inv_level1_p<false>(J, W[1], win1, tx, ty, tid);
        } else {
            inv_level1_p<true>(J, W[1], win1, tx, ty, tid);
        }
    }

    /* ---- level 1 of I frames: inverse B4T, columns then rows (sbt.c:129-163,204-238,253-265) ----------
     * out[2m]   = r8(L[m-1] + 3L[m] + H[m-1] - 3H[m]),  out[2m+1] = r8(3L[m] + L[m+1] + 3H[m] - H[m+1]),
     * L[-1] := L[0], L[n/2] := L[n/2-1] (same for H).  Column pass: a thread walks down one column of the
     * LL|HL (low) or LH|HH (high) pair with a sliding 3-row window (coalesced 4-byte loads across the warp);
     * row pass: a thread owns 8 adjacent output samples of one row and stores them with one 8-byte store. */
    if (isI) {
        const Win w = W[1];
        const int ww = w.b - w.a;
        const int wo = cw >> 1, ho = ch >> 1;
        int32_t *vbuf = ibase; /* [SBT_TH][INV_VSTRIDE]: low columns at voff.., high columns at INV_VHALF + voff.. */
        const int32_t *bLL = win1;
        const int rows = W[0].hb - W[0].ha; /* output rows of this tile (<= 64, even) */
        const int m_lo = gy0 >> 1, nm = rows >> 1;
        /* Column pass.  tasks: (column in window) x (low | high) x (runs of 4 row pairs).  A task first requests all
         * the rows it needs (6 x 2 independent loads: one memory latency instead of one per row pair), then slides.
         * vbuf keeps the tile's own columns 16-byte aligned: window column x sits at x + voff. */
        const int voff = 4 - (tx * (SBT_TW / 2) - w.a);
        constexpr int RUN = 4;
        const int nrun = (nm + RUN - 1) / RUN;
        for (int task = tid; task < 2 * ww * nrun; task += SBT_TILE_THREADS) {
            const int part = task / (2 * ww), c = task - part * 2 * ww;
            const bool hcol = c >= ww;
            const int x = hcol ? c - ww : c; /* window column */
            const int gx = w.a + x;          /* band column */
            const int mA = m_lo + part * RUN, mB = imin(m_lo + nm, mA + RUN);
            if (mA >= mB) {
                continue;
            }
            const int32_t *gL = J.coef + wo + gx, *gH = J.coef + (size_t) ho * cw + (hcol ? wo : 0) + gx;
            const int32_t *sL = bLL + x - w.ha * ww;
            int Lr[RUN + 2], Hr[RUN + 2];
#pragma unroll
            for (int i = 0; i < RUN + 2; i++) {
                const int m = iclamp(mA - 1 + i, 0, ho - 1); /* L[-1] := L[0], L[n/2] := L[n/2-1] */
                Lr[i] = hcol ? gL[(size_t) m * cw] : sL[m * ww];
                Hr[i] = gH[(size_t) m * cw];
            }
            int32_t *vcol = vbuf + (hcol ? INV_VHALF : 0) + x + voff + 2 * (mA - m_lo) * INV_VSTRIDE;
#pragma unroll
            for (int i = 0; i < RUN; i++) {
                if (mA + i < mB) {
                    vcol[(2 * i) * INV_VSTRIDE] = rnd_shift<3>(Lr[i] + 3 * Lr[i + 1] + Hr[i] - 3 * Hr[i + 1]);
                    vcol[(2 * i + 1) * INV_VSTRIDE] = rnd_shift<3>(3 * Lr[i + 1] + Lr[i + 2] + 3 * Hr[i + 1] - Hr[i + 2]);
                }
            }
        }
        __syncthreads();
        /* Row pass: the thread's four columns k0..k0+3 are one aligned 16-byte shared-memory load per band */
        for (int g = tid; g < (SBT_TW / 8) * SBT_TH; g += SBT_TILE_THREADS) {
            const int ly = g >> 4, k0l = (g & 15) * 4;
            const int k0 = tx * (SBT_TW / 2) + k0l; /* first of 4 band columns -> 8 samples */
            if (ly >= rows || k0 >= wo) {
                continue;
            }
            const int32_t *L = vbuf + ly * INV_VSTRIDE + voff - w.a, *H = L + INV_VHALF;
            int v[8];
            if (k0 + 4 <= wo) {
                const int4 l4 = *reinterpret_cast<const int4 *>(L + k0), h4 = *reinterpret_cast<const int4 *>(H + k0);
                const int kp = imax(k0 - 1, 0), kn = imin(k0 + 4, wo - 1);
                const int Lv[6] = {L[kp], l4.x, l4.y, l4.z, l4.w, L[kn]}, Hv[6] = {H[kp], h4.x, h4.y, h4.z, h4.w, H[kn]};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    v[2 * i] = rnd_shift<3>(Lv[i] + 3 * Lv[i + 1] + Hv[i] - 3 * Hv[i + 1]);
                    v[2 * i + 1] = rnd_shift<3>(3 * Lv[i + 1] + Lv[i + 2] + 3 * Hv[i + 1] - Hv[i + 2]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int k = imin(k0 + i, wo - 1);
                    const int kp = imax(k - 1, 0), kn = imin(k + 1, wo - 1);
                    v[2 * i] = rnd_shift<3>(L[kp] + 3 * L[k] + H[kp] - 3 * H[k]);
                    v[2 * i + 1] = rnd_shift<3>(3 * L[k] + L[kn] + 3 * H[k] - H[kn]);
                }
            }
            store_row8(J, gy0 + ly, 2 * k0, v);
        }
    }
}

__global__ void __launch_bounds__(SBT_TILE_THREADS, SBT_INV_MINB) sbt_inv_tile_kernel(const SbtJob *jobs, const SbtDims dims)
{
    sbt_inv_tile_body<false>(jobs, dims);
}
__global__ void __launch_bounds__(SBT_TILE_THREADS, SBT_INV_MINB_I) sbt_inv_tile_intra_kernel(const SbtJob *jobs, const SbtDims dims)
{
    sbt_inv_tile_body<true>(jobs, dims);
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

void sbt_inv_launch(const SbtJob *d_jobs, const SbtDims &dims, size_t lo_smem, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1)
{
    if (dims.njobs <= 0) {
        return;
    }
    const size_t tile_smem = inv_tile_smem(false), tile_smem_i = inv_tile_smem(true);
    if (lo_smem > 48 * 1024) {
        CUDA_CHECK(cudaFuncSetAttribute(sbt_inv_lo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) lo_smem));
    }
    if (dims.any_intra) { /* static shared memory (job record, windows) counts against the 48 KB default too */
        CUDA_CHECK(cudaFuncSetAttribute(sbt_inv_tile_intra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) tile_smem_i));
    }
    DSV_LAUNCH(sbt_inv_lo_kernel, dim3(dims.njobs), dim3(SBT_LO_THREADS), lo_smem, st, d_jobs);
    KERNEL_CHECK();
    DSV_LAUNCH(sbt_inv_mid_kernel, dim3(dims.mtiles), dim3(SBT_TILE_THREADS), 0, st, d_jobs, dims);
    KERNEL_CHECK();
    if (ev0) {
        CUDA_CHECK(cudaEventRecord(ev0, st));
    }
    if (dims.any_inter) {
        DSV_LAUNCH(sbt_inv_tile_kernel, dim3(dims.tiles), dim3(SBT_TILE_THREADS), tile_smem, st, d_jobs, dims);
        KERNEL_CHECK();
    }
    if (dims.any_intra) {
        DSV_LAUNCH(sbt_inv_tile_intra_kernel, dim3(dims.tiles), dim3(SBT_TILE_THREADS), tile_smem_i, st, d_jobs, dims);
        KERNEL_CHECK();
    }
    if (ev1) {
        CUDA_CHECK(cudaEventRecord(ev1, st));
    }
}

} // namespace dsv
