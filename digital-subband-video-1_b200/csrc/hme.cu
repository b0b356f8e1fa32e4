/*
 * hme.cu -- hierarchical motion estimation.  Replaces refine_level / dsv_hme and their helpers
 * (hme.c:32-741): candidate inheritance from the parent level, 9-point full-pel SAD search, and at
 * level 0 the half-pel refinement on the 14x14 centre patch, the block statistics and the
 * intra / inter decision with its quadrant mask.
 *
 *   hme_level_kernel   levels > 0: one WARP per visited block, four blocks per CTA, no block barrier.
 *                      The source block is staged once in shared memory as aligned 32-bit words;
 *                      every candidate's SAD is accumulated with __vsadu4 over unaligned reference
 *                      words (two aligned ld.global.nc + funnel shift), reduced with warp
 *                      reductions; the FIRST minimum in the reference's candidate order wins
 *                      (strict '<', hme.c:503,531).
 *   hme_l0_kernel      level 0: the same search, then the half-pel image of the 16x16 reference
 *                      patch (hme.c:350-376) as three phase planes in shared memory (H, V, HV: a
 *                      candidate's 14 samples of a row are consecutive bytes), 8 half-pel SADs on
 *                      packed words with a patch row per lane, and 22 block sums (variance /
 *                      texture / chroma variance / patch textures) reduced in one pass.  The
 *                      decision cascade (hme.c:651-718) runs on every lane -- its inputs are
 *                      warp-uniform -- and only the blocks it marks intra take the two passes
 *                      nothing else needs: the reduced-range test and the quadrant good-vs-evil
 *                      metric.  All unsigned 32-bit wrap-arounds of the reference are kept.
 *   hme_neigh_kernel   high_detail needs the left / top / top-left blocks' final mode and flags
 *                      (hme.c:621-648): a second, one-thread-per-block pass.
 *
 * Search order, tie-breaking, clamps and the "last candidate" default follow SURVEY.md section 3.3
 * and Appendix B-5/B-6 exactly; MV fields come out byte-identical to the reference's DSV_MV records.
 */
#include "motion.cuh"

namespace dsv {

#define HME_WARPS 4 /* blocks per CTA: one warp each, no block-wide barrier anywhere */
#define HME_THREADS (32 * HME_WARPS)
#define HME_SRC_STRIDE 64 /* bytes per staged block row */
#ifndef HME_UR_CAND
#define HME_UR_CAND 8 /* reference rows requested ahead in the candidate SAD loops */
#endif
#ifndef HME_UR_9PT
#define HME_UR_9PT 4 /* A/B at 64 lanes: (8, 4) 668 us, (4, 2) 678, (8, 8) 676, (1, 1) 807 */
#endif
#define HME_PRAGMA(x) _Pragma(#x)
#define HME_UNROLL(n) HME_PRAGMA(unroll n)
#define HP_SAD_SZ 14
#define HP_DIM 16

/* the search offsets in the reference's candidate order (hme.c:505-541 full-pel, hme.c:560-575 half-pel), packed two
 * bits per entry (value + 1) so that picking the winner's offset is a shift instead of a table in local memory */
template <int N> constexpr unsigned pack_offsets(const int (&v)[N])
{
    unsigned r = 0;
    for (int i = 0; i < N; i++) {
        r |= (unsigned) (v[i] + 1) << (2 * i);
    }
    return r;
}
constexpr int FP_X[9] = {0, 1, -1, 0, 0, -1, 1, -1, 1}, FP_Y[9] = {0, 0, 0, 1, -1, -1, -1, 1, 1};
constexpr int HP_X[8] = {1, -1, 0, 0, -1, 1, -1, 1}, HP_Y[8] = {0, 0, 1, -1, -1, -1, 1, 1};
constexpr unsigned FP_XP = pack_offsets(FP_X), FP_YP = pack_offsets(FP_Y), HP_XP = pack_offsets(HP_X), HP_YP = pack_offsets(HP_Y);
DSV_D int offset_of(unsigned packed, int m) { return (int) ((packed >> (2 * m)) & 3u) - 1; }

/* per-warp shared memory of the search: the staged source block and the candidate list */
struct HmeSearchSmem {
    uint8_t src[64 * HME_SRC_STRIDE];
    int cx[8], cy[8], valid[8];
    int n;
};
/* level 0 adds the half-pel image of the 16x16 reference patch */
#define HP_PLANE (HP_DIM * HP_DIM) /* a half-pel phase plane of the patch: 16 rows of 16 bytes */
struct HmeL0Smem {
    HmeSearchSmem s;
    __align__(16) uint8_t tmp[3 * HP_PLANE];    /* phase planes H (x + 1/2), V (y + 1/2), HV */
    __align__(16) uint8_t refblk[HP_SAD_SZ * 16]; /* the chosen reference patch, 16 bytes per row */
    int16_t hbuf[(HP_DIM + 4) * HP_DIM];
};

/* ld4u on frame memory the kernel only reads (ld.global.nc instead of a generic load) */
DSV_D unsigned ld4g(const uint8_t *p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const unsigned *q = reinterpret_cast<const unsigned *>(a & ~(uintptr_t) 3);
    return __funnelshift_r(__ldg(q), __ldg(q + 1), (unsigned) (a & 3) * 8);
}
/* 16 bytes at an arbitrary address of such memory: five aligned words, four funnel shifts */
DSV_D void ld16g(const uint8_t *p, unsigned &w0, unsigned &w1, unsigned &w2, unsigned &w3)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const unsigned *q = reinterpret_cast<const unsigned *>(a & ~(uintptr_t) 3);
    const unsigned sh = (unsigned) (a & 3) * 8;
    const unsigned a0 = __ldg(q), a1 = __ldg(q + 1), a2 = __ldg(q + 2), a3 = __ldg(q + 3), a4 = __ldg(q + 4);
    w0 = __funnelshift_r(a0, a1, sh);
    w1 = __funnelshift_r(a1, a2, sh);
    w2 = __funnelshift_r(a2, a3, sh);
    w3 = __funnelshift_r(a3, a4, sh);
}
/* SAD of the 14 bytes s0..s3 (bytes 14, 15 of s3 clear) against bytes OFF .. OFF + 13 of the 16-byte row r */
template <int OFF> DSV_D unsigned sad14(const uint4 &r, unsigned s0, unsigned s1, unsigned s2, unsigned s3)
{
    if (OFF == 0) {
        return __vsadu4(s0, r.x) + __vsadu4(s1, r.y) + __vsadu4(s2, r.z) + __vsadu4(s3, r.w & 0xffffu);
    }
    return __vsadu4(s0, __funnelshift_r(r.x, r.y, 8)) + __vsadu4(s1, __funnelshift_r(r.y, r.z, 8)) +
           __vsadu4(s2, __funnelshift_r(r.z, r.w, 8)) + __vsadu4(s3, (r.w >> 8) & 0xffffu);
}

/* A block's words are dealt to the warp as (word column, row group): lanes_per_row = the power of two >= words,
 * 32 / lanes_per_row row groups of consecutive rows.  A lane keeps its word column and walks its rows top to
 * bottom, so vertically adjacent data slides through registers and the 9-point search loads every reference row
 * once instead of three times. */
struct BlockGeom {
    int bx, by, bw, bh, words;
    unsigned tail_mask;
    int wx, r0, r1, rpg; /* this lane: word column, rows [r0, r1), rows per group (uniform trip count) */
    bool active;         /* wx < words */
    unsigned m;          /* byte mask of this lane's word (the block's last word column may be partial) */
};

template <int N> DSV_D void warp_reduce_n(unsigned (&acc)[N])
{
#pragma unroll
    for (int k = 0; k < N; k++) {
        acc[k] = __reduce_add_sync(0xffffffffu, acc[k]);
    }
}

/* stage the source block (bytes past bw zero) and deal the words to the lanes; false: the block lies outside */
DSV_D bool block_setup(const HmeArgs &A, int i, int j, BlockGeom &G, uint8_t *s_src, int lane)
{
    G.bx = (i * A.blk_w) >> A.level;
    G.by = (j * A.blk_h) >> A.level;
    if (G.bx >= A.src.w || G.by >= A.src.h) {
        return false;
    }
    G.bw = imin(A.src.w - G.bx, A.blk_w);
    G.bh = imin(A.src.h - G.by, A.blk_h);
    G.words = (G.bw + 3) >> 2;
    const int tail = G.bw & 3;
    G.tail_mask = tail ? ((1u << (8 * tail)) - 1u) : 0xffffffffu;
    const int lg = G.words <= 1 ? 0 : 32 - __clz(G.words - 1);
    const int rgs = 32 >> lg;
    G.rpg = (G.bh + rgs - 1) / rgs;
    G.wx = lane & ((1 << lg) - 1);
    G.active = G.wx < G.words;
    G.r0 = imin(G.bh, (lane >> lg) * G.rpg);
    G.r1 = G.active ? imin(G.bh, G.r0 + G.rpg) : G.r0;
    G.m = (G.wx == G.words - 1) ? G.tail_mask : 0xffffffffu;
    const uint8_t *p = A.src.p + (ptrdiff_t) (G.by + G.r0) * A.src.stride + G.bx + 4 * G.wx;
    for (int r = G.r0; r < G.r1; r++, p += A.src.stride) {
        *reinterpret_cast<unsigned *>(s_src + r * HME_SRC_STRIDE + 4 * G.wx) = ld4g(p) & G.m;
    }
    __syncwarp();
    return true;
}

/*
 * Candidate selection + 9-point full-pel search (hme.c:439-541).  Every lane of the warp calls;
 * the result (full-pel dx, dy at this level and the winning SAD) is returned to all lanes.
 */
DSV_D void search_block(const HmeArgs &A, int i, int j, const BlockGeom &G, HmeSearchSmem &S, int lane, int &odx, int &ody, int &obest)
{
    const int level = A.level;
    const int W = A.ref.w, H = A.ref.h;
    const int rs = A.ref.stride;
    const uint8_t *s_src = S.src;

    if (lane == 0) {
        int n = 0;
        int call[8];
        call[n] = 0;
        S.cx[n] = 0;
        S.cy[n] = 0;
        n++;
        if (A.parent) {
            const int step = 1 << level;
            const int pmask = ~((step << 1) - 1);
            const int pi = i & pmask, pj = j & pmask;
            const int ptx[5] = {0, -2, 2, 0, 0}, pty[5] = {0, 0, 0, -2, 2};
            for (int m = 0; m < 5; m++) {
                const int x = pi + ptx[m] * step, y = pj + pty[m] * step;
                if (x >= 0 && x < A.nbh && y >= 0 && y < A.nbv) {
                    const DevMV pm = A.parent[x + y * A.nbh];
                    const int all = (int) ((unsigned) (uint16_t) pm.x | ((unsigned) (uint16_t) pm.y << 16));
                    if (all) {
                        bool exists = false;
                        for (int k = 0; k < n; k++) {
                            exists |= call[k] == all;
                        }
                        if (!exists) {
                            call[n] = all;
                            S.cx[n] = pm.x;
                            S.cy[n] = pm.y;
                            n++;
                        }
                    }
                }
            }
        }
        for (int k = 0; k < n; k++) {
            const int dx = S.cx[k] >> level, dy = S.cy[k] >> level;
            const int x = G.bx + dx, y = G.by + dy;
            S.valid[k] = !(x < -DSV_BORDER || y < -DSV_BORDER || x + G.bw > W + DSV_BORDER || y + G.bh > H + DSV_BORDER);
        }
        S.n = n;
    }
    __syncwarp();
    const int n = S.n;
    int best_k = n - 1;
    if (n > 1) {
        unsigned acc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 6; k++) {
            if (k < n && S.valid[k]) {
                const int dx = S.cx[k] >> level, dy = S.cy[k] >> level;
                const uint8_t *p = A.ref.p + (ptrdiff_t) (G.by + dy + G.r0) * rs + G.bx + dx + 4 * G.wx;
                unsigned t = 0;
                HME_UNROLL(HME_UR_CAND)
                for (int r = G.r0; r < G.r1; r++, p += rs) {
                    t += __vsadu4(*reinterpret_cast<const unsigned *>(s_src + r * HME_SRC_STRIDE + 4 * G.wx), ld4g(p) & G.m);
                }
                acc[k] = t;
            }
        }
        warp_reduce_n<6>(acc);
        int best_score = 0x7fffffff;
        for (int k = 0; k < n; k++) {
            if (S.valid[k] && best_score > (int) acc[k]) {
                best_score = (int) acc[k];
                best_k = k;
            }
        }
    }
    int dx = S.cx[best_k] >> level, dy = S.cy[best_k] >> level;
    dx = iclamp(dx, -G.bw - G.bx, W - G.bx);
    dy = iclamp(dy, -G.bh - G.by, H - G.by);
    const int xx = G.bx + dx, yy = G.by + dy;
    unsigned acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (G.r1 > G.r0) {
        /* reference rows r0-1 .. r1 of this lane's word column, each loaded once as bytes [-1, 5) around the word:
         * the three horizontal candidates are the windows at byte offsets 0, 1, 2; a reference row t meets the
         * source rows t (yf = 0), t - 1 (yf = +1) and t + 1 (yf = -1) */
        const uint8_t *p = A.ref.p + (ptrdiff_t) (yy + G.r0 - 1) * rs + xx + 4 * G.wx - 1;
        const unsigned *q = reinterpret_cast<const unsigned *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t) 3);
        const unsigned bsh = ((unsigned) reinterpret_cast<uintptr_t>(p) & 3u) * 8u;
        const int qs = rs >> 2;
        const uint8_t *sp = s_src + 4 * G.wx;
        unsigned s_prev = 0, s_cur = 0, s_next = *reinterpret_cast<const unsigned *>(sp + G.r0 * HME_SRC_STRIDE);
        HME_UNROLL(HME_UR_9PT)
        for (int t = G.r0 - 1; t <= G.r1; t++, q += qs) {
            s_prev = s_cur;
            s_cur = s_next;
            if (t + 1 < G.r1) {
                s_next = *reinterpret_cast<const unsigned *>(sp + (t + 1) * HME_SRC_STRIDE);
            }
            const unsigned q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
            const unsigned lo = __funnelshift_r(q0, q1, bsh), hi = __funnelshift_r(q1, q2, bsh);
            const unsigned ra = lo & G.m, rb = __funnelshift_r(lo, hi, 8) & G.m, rc = __funnelshift_r(lo, hi, 16) & G.m;
            if (t >= G.r0 && t < G.r1) {
                acc[2] += __vsadu4(s_cur, ra);
                acc[0] += __vsadu4(s_cur, rb);
                acc[1] += __vsadu4(s_cur, rc);
            }
            if (t > G.r0) { /* source row t - 1 */
                acc[7] += __vsadu4(s_prev, ra);
                acc[3] += __vsadu4(s_prev, rb);
                acc[8] += __vsadu4(s_prev, rc);
            }
            if (t + 1 < G.r1) { /* source row t + 1 */
                acc[5] += __vsadu4(s_next, ra);
                acc[4] += __vsadu4(s_next, rb);
                acc[6] += __vsadu4(s_next, rc);
            }
        }
    }
    warp_reduce_n<9>(acc);
    int best = 0x7fffffff, m = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) {
        if (best > (int) acc[k]) {
            best = (int) acc[k];
            m = k;
        }
    }
    odx = dx + offset_of(FP_XP, m);
    ody = dy + offset_of(FP_YP, m);
    obest = best;
}

DSV_D void store_mv(DevMV *dst, int x, int y, int mode, int submask, int lo_var, int lo_tex)
{
    DevMV m;
    m.x = (int16_t) x;
    m.y = (int16_t) y;
    m.mode = (uint8_t) mode;
    m.submask = (uint8_t) submask;
    m.lo_var = (uint8_t) lo_var;
    m.lo_tex = (uint8_t) lo_tex;
    m.high_detail = 0;
    m.pad[0] = m.pad[1] = m.pad[2] = 0;
    *dst = m;
}

__global__ void __launch_bounds__(HME_THREADS) hme_level_kernel(const HmeArgs *args, int nbx, int nby)
{
    const HmeArgs &A = args[blockIdx.z];
    __shared__ __align__(16) HmeSearchSmem smem[HME_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wb = (int) blockIdx.x * HME_WARPS + warp;
    if (wb >= nbx * nby) {
        return;
    }
    const int step = 1 << A.level;
    const int bj = wb / nbx;
    const int i = (wb - bj * nbx) * step, j = bj * step;
    HmeSearchSmem &S = smem[warp];
    BlockGeom G;
    if (!block_setup(A, i, j, G, S.src, lane)) {
        if (lane == 0) {
            store_mv(&A.out[i + j * A.nbh], 0, 0, 0, 0, 0, 0);
        }
        return;
    }
    int dx, dy, best;
    search_block(A, i, j, G, S, lane, dx, dy, best);
    if (lane == 0) {
        store_mv(&A.out[i + j * A.nbh], dx << A.level, dy << A.level, 0, 0, 0, 0);
    }
}

enum {
    SUM_S, SUM_SS, SUM_SH, SUM_SV,          /* source block: block_analysis */
    SUM_RS, SUM_RSS,                        /* zero-MV reference block: y_sqrvar, block_intra_test mean */
    SUM_CSU, SUM_CSSU, SUM_CSV, SUM_CSSV,   /* source chroma */
    SUM_CRU, SUM_CRSU, SUM_CRV, SUM_CRSV,   /* reference chroma */
    SUM_PSH, SUM_PSV, SUM_PAV, SUM_PAVS,    /* source 14x14 patch: block_texture */
    SUM_QSH, SUM_QSV, SUM_QAV, SUM_QAVS,    /* chosen reference patch */
    SUM_COUNT
};

/* Sums and sums of squares of the block's four chroma rectangles (source U, V, reference U, V: c_maxvar's inputs,
 * hme.c:270-300) into sum[2k], sum[2k + 1].  One walk serves the four planes, so four loads are in flight per step;
 * the (row, column) of a lane's item advances by additions instead of a division per item. */
DSV_D void warp_chroma_moments(const HmeArgs &A, int cbx, int cby, int cw, int ch, int lane, unsigned *sum)
{
    if (cw <= 0 || ch <= 0) {
        return;
    }
    const HmePlane *pl[4] = {&A.srcU, &A.srcV, &A.refU, &A.refV};
    const uint8_t *p0[4];
    int st[4];
    unsigned al = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        st[k] = pl[k]->stride;
        p0[k] = pl[k]->p + (ptrdiff_t) cby * st[k] + cbx;
        al |= (unsigned) reinterpret_cast<uintptr_t>(p0[k]) | (unsigned) st[k];
    }
    if ((cw & 3) == 0) {
        const bool v16 = ((cw | (int) al) & 15) == 0; /* 16 samples per item, else 4 */
        const int per = v16 ? cw >> 4 : cw >> 2;
        int ly = lane / per, x = lane - ly * per;
        const int dq = 32 / per, dr = 32 - dq * per;
        if (v16) {
            while (ly < ch) {
                uint4 v[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    v[k] = __ldg(reinterpret_cast<const uint4 *>(p0[k] + (ptrdiff_t) ly * st[k]) + x);
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    sum[2 * k] += __vsadu4(v[k].x, 0u) + __vsadu4(v[k].y, 0u) + __vsadu4(v[k].z, 0u) + __vsadu4(v[k].w, 0u);
                    sum[2 * k + 1] = __dp4a(v[k].x, v[k].x, __dp4a(v[k].y, v[k].y, __dp4a(v[k].z, v[k].z, __dp4a(v[k].w, v[k].w, sum[2 * k + 1]))));
                }
                ly += dq;
                x += dr;
                if (x >= per) {
                    x -= per;
                    ly++;
                }
            }
            return;
        }
        while (ly < ch) {
            unsigned w[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                w[k] = ld4g(p0[k] + (ptrdiff_t) ly * st[k] + 4 * x);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                sum[2 * k] += __vsadu4(w[k], 0u);
                sum[2 * k + 1] = __dp4a(w[k], w[k], sum[2 * k + 1]);
            }
            ly += dq;
            x += dr;
            if (x >= per) {
                x -= per;
                ly++;
            }
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        unsigned s1 = 0, s2 = 0;
        for (int ly = 0; ly < ch; ly++) {
            for (int lx = lane; lx < cw; lx += 32) {
                const unsigned v = p0[k][(ptrdiff_t) ly * st[k] + lx];
                s1 += v;
                s2 += v * v;
            }
        }
        sum[2 * k] += s1;
        sum[2 * k + 1] += s2;
    }
}

/* intra_metric (hme.c:87-134) of the block's four quadrants against the zero-MV reference block, summed over the warp:
 * ge[q] = good, ge[4 + q] = evil.  Only the blocks the cascade has already marked intra get here. */
DSV_D void quadrant_metric(const BlockGeom &G, const uint8_t *s_src, const uint8_t *ref0, int rs, int lane, unsigned (&ge)[8])
{
    const int sbw = G.bw / 2, sbh = G.bh / 2;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        ge[k] = 0;
    }
    if ((G.bw & 7) == 0) {
        /* sbw is a multiple of 4: a word lies in one quadrant column, so a lane keeps two running sums (all rows, top
         * rows); the bytes left of a word come from the neighbouring lane, the row above is last step's word */
        unsigned g_all = 0, e_all = 0, g_top = 0, e_top = 0;
        const int lx0 = 4 * G.wx;
        const int qxi = lx0 >= sbw;
        const int rbase = G.active ? G.r0 : 0; /* idle lanes walk along (shuffles are warp-wide) */
        unsigned w_up = 0, r_up = 0;
        if (G.active && G.r0 > 0 && G.r0 < G.bh) {
            w_up = *reinterpret_cast<const unsigned *>(s_src + (G.r0 - 1) * HME_SRC_STRIDE + lx0);
            r_up = ld4g(ref0 + (ptrdiff_t) (G.r0 - 1) * rs + lx0);
        }
        for (int it = 0; it < G.rpg; it++) {
            const int ly = rbase + it;
            const bool on = G.active && ly < G.r1;
            unsigned w = 0, r = 0;
            if (on) {
                w = *reinterpret_cast<const unsigned *>(s_src + ly * HME_SRC_STRIDE + lx0);
                r = ld4g(ref0 + (ptrdiff_t) ly * rs + lx0);
            }
            const unsigned w_l = __shfl_up_sync(0xffffffffu, w, 1), r_l = __shfl_up_sync(0xffffffffu, r, 1);
            if (on) {
                if (ly < 2 * sbh) {
                    const int qyi = ly >= sbh;
                    const int qi0 = lx0 - qxi * sbw, qj = ly - qyi * sbh;
                    const unsigned wl = (w << 8) | (qi0 == 0 ? (w & 0xffu) : (w_l >> 24));
                    const unsigned rl = (r << 8) | (qi0 == 0 ? (r & 0xffu) : (r_l >> 24));
                    const unsigned ua = qj == 0 ? w : w_up;
                    const unsigned ub = qj == 0 ? r : r_up;
                    unsigned g = __vsadu4(w, wl) + __vsadu4(w, ua) + __vsadu4(r, rl) + __vsadu4(r, ub);
                    /* |w - r| > 2 counts as evil; 0 / 1 / 2 add 192 / 128 / 96 = 192 - 64 v + 32 (v >> 1) to good */
                    const unsigned d = __vabsdiffu4(w, r);
                    const unsigned big = __vcmpgtu4(d, 0x02020202u);
                    const unsigned sm = d & ~big;
                    const unsigned e = __vsadu4(d & big, 0u);
                    g += 192u * (4u - ((unsigned) __popc(big) >> 3)) - 64u * __vsadu4(sm, 0u) + 32u * (unsigned) __popc(sm & 0x02020202u);
                    g_all += g;
                    e_all += e;
                    g_top += qyi ? 0u : g;
                    e_top += qyi ? 0u : e;
                }
                w_up = w;
                r_up = r;
            }
        }
        const unsigned g_bot = g_all - g_top, e_bot = e_all - e_top;
        ge[0] = qxi ? 0u : g_top;
        ge[1] = qxi ? g_top : 0u;
        ge[2] = qxi ? 0u : g_bot;
        ge[3] = qxi ? g_bot : 0u;
        ge[4] = qxi ? 0u : e_top;
        ge[5] = qxi ? e_top : 0u;
        ge[6] = qxi ? 0u : e_bot;
        ge[7] = qxi ? e_bot : 0u;
    } else {
        /* general widths (blocks cut by the picture edge): rows in turn, columns to lanes */
        for (int ly = 0; ly < 2 * sbh; ly++) {
            const int qyi = ly >= sbh;
            const int qj = ly - qyi * sbh;
            unsigned good0 = 0, good1 = 0, evil0 = 0, evil1 = 0;
            for (int lx = lane; lx < 2 * sbw; lx += 32) {
                const uint8_t *sp = s_src + ly * HME_SRC_STRIDE + lx;
                const uint8_t *rp = ref0 + (ptrdiff_t) ly * rs + lx;
                const int pa = sp[0], pb = rp[0];
                const int qxi = lx >= sbw;
                const int qi = lx - qxi * sbw;
                const int la = qi == 0 ? pa : sp[-1], lb = qi == 0 ? pb : rp[-1];
                const int ua = qj == 0 ? pa : sp[-HME_SRC_STRIDE], ub = qj == 0 ? pb : rp[-rs];
                unsigned good = (unsigned) (iabs(pa - la) + iabs(pa - ua) + iabs(pb - lb) + iabs(pb - ub));
                unsigned evil = 0;
                const int dif = iabs(pa - pb);
                if (dif > 2) {
                    evil = (unsigned) dif;
                } else {
                    good += dif == 0 ? 192u : (dif == 1 ? 128u : 96u);
                }
                good0 += qxi ? 0u : good;
                good1 += qxi ? good : 0u;
                evil0 += qxi ? 0u : evil;
                evil1 += qxi ? evil : 0u;
            }
            ge[0] += qyi ? 0u : good0;
            ge[1] += qyi ? 0u : good1;
            ge[2] += qyi ? good0 : 0u;
            ge[3] += qyi ? good1 : 0u;
            ge[4] += qyi ? 0u : evil0;
            ge[5] += qyi ? 0u : evil1;
            ge[6] += qyi ? evil0 : 0u;
            ge[7] += qyi ? evil1 : 0u;
        }
    }
    warp_reduce_n<8>(ge);
}

/* occupancy A/B at 64 lanes (round 2): no cap (80 registers, 6 CTAs of 4 warps per SM) 668 us; 8 CTAs / 64 registers
 * 685 us; 10 CTAs / 48 registers 798 us; an explicit minimum of 1 CTA lets ptxas spend more registers: 807 us */
#ifdef HME_L0_MINB
__global__ void __launch_bounds__(HME_THREADS, HME_L0_MINB) hme_l0_kernel(const HmeArgs *args)
#else
__global__ void __launch_bounds__(HME_THREADS) hme_l0_kernel(const HmeArgs *args)
#endif
{
    const HmeArgs &A = args[blockIdx.z];
    __shared__ __align__(16) HmeL0Smem smem[HME_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wb = (int) blockIdx.x * HME_WARPS + warp;
    if (wb >= A.nbh * A.nbv) {
        return;
    }
    const int j = wb / A.nbh, i = wb - j * A.nbh;
    HmeL0Smem &L = smem[warp];
    uint8_t *s_src = L.s.src;
    int16_t *s_hbuf = L.hbuf;
    uint8_t *s_tmp = L.tmp, *s_refblk = L.refblk;
    DevMV *out = &A.out[i + j * A.nbh];
    BlockGeom G;
    if (!block_setup(A, i, j, G, s_src, lane)) {
        if (lane == 0) {
            store_mv(out, 0, 0, 0, 0, 0, 0);
            A.aux[i + j * A.nbh] = make_int2(0, 0);
        }
        return;
    }
    int fx, fy, best;
    search_block(A, i, j, G, L.s, lane, fx, fy, best);

    const int rs = A.ref.stride, ss = A.src.stride;
    const unsigned yarea = (unsigned) (G.bw * G.bh);
    const unsigned yareasq = yarea * yarea;
    const int cx = G.bx + ((G.bw >> 1) - (HP_SAD_SZ / 2)), cy = G.by + ((G.bh >> 1) - (HP_SAD_SZ / 2));
    int mvx = fx, mvy = fy;
    bool has_hp = false;

    /* the 14 bytes of row (lane & 15) of the source's centre patch (lanes 0-13 and 16-29), bytes 14 / 15 clear */
    const int prow = lane & 15;
    unsigned s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    if (prow < HP_SAD_SZ) {
        ld16g(A.src.p + (ptrdiff_t) (cy + prow) * ss + cx, s0, s1, s2, s3);
        s3 &= 0xffffu;
    }
    unsigned *s_refblk_w = reinterpret_cast<unsigned *>(s_refblk);

    /* ---- half-pel refinement (hme.c:551-591) ---- */
    if (best > A.blk_w * A.blk_h) {
        int best_hp = (int) ((unsigned) (best * (HP_SAD_SZ * HP_SAD_SZ)) / yarea);
        const uint8_t *rp0 = A.ref.p + (ptrdiff_t) (cy + mvy - 1) * rs + (cx + mvx - 1); /* hpel()'s `ref` */
        /* H-filtered rows -1 .. 18 of the patch, four columns per lane and step: bytes p[-1..6] of the row */
        for (int k = lane; k < (HP_DIM + 4) * (HP_DIM / 4); k += 32) {
            const int jj = k >> 2, i0 = (k & 3) * 4;
            const uint8_t *p = rp0 + (ptrdiff_t) (jj - 1) * rs + i0 - 1;
            const unsigned wa = ld4g(p), wb2 = ld4g(p + 4);
            int16_t *d = s_hbuf + jj * HP_DIM + i0;
            d[0] = (int16_t) hp_taps_u8x4(wa);
            d[1] = (int16_t) hp_taps_u8x4(__funnelshift_r(wa, wb2, 8));
            d[2] = (int16_t) hp_taps_u8x4(__funnelshift_r(wa, wb2, 16));
            d[3] = (int16_t) hp_taps_u8x4(__funnelshift_r(wa, wb2, 24));
        }
        __syncwarp();
        /* The half-pel image of the patch (hme.c:350-376) as three phase planes of 16 x 16 bytes -- H (x + 1/2),
         * V (y + 1/2), HV -- so that a candidate's 14 samples of a row are consecutive bytes; the full-pel phase is
         * never a candidate.  Four columns per lane and step; the H filter of patch row jj is hbuf row jj + 1. */
        for (int k = lane; k < HP_DIM * (HP_DIM / 4); k += 32) {
            const int jj = k >> 2, i0 = (k & 3) * 4;
            const uint8_t *p = rp0 + (ptrdiff_t) jj * rs + i0;
            const unsigned ra = ld4g(p - rs), rb = ld4g(p), rc = ld4g(p + rs), rd = ld4g(p + 2 * rs);
            const int16_t *b = s_hbuf + jj * HP_DIM + i0;
            unsigned hw = 0, vw = 0, xw = 0;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int v = 9 * (byte_of(rb, e) + byte_of(rc, e)) - (byte_of(ra, e) + byte_of(rd, e));
                vw |= (unsigned) clamp_u8((v + 8) >> 4) << (8 * e);
                hw |= (unsigned) clamp_u8((b[HP_DIM + e] + 8) >> 4) << (8 * e);
                xw |= (unsigned) clamp_u8((9 * (b[HP_DIM + e] + b[2 * HP_DIM + e]) - (b[e] + b[3 * HP_DIM + e]) + 128) >> 8) << (8 * e);
            }
            unsigned *d = reinterpret_cast<unsigned *>(s_tmp) + k;
            d[0] = hw;
            d[HP_PLANE / 4] = vw;
            d[2 * (HP_PLANE / 4)] = xw;
        }
        __syncwarp();
        /* candidate (xh, yh), sample (ii, jj) = plane[jj + (yh >= 0)][ii + (xh >= 0)]: lanes 0-13 take the H / V
         * candidates of row jj = lane, lanes 16-29 the four HV candidates of row jj = lane - 16 */
        unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        if (prow < HP_SAD_SZ) {
            const uint4 *pl = reinterpret_cast<const uint4 *>(s_tmp);
            if (lane < 16) {
                const uint4 h = pl[prow + 1], v1 = pl[HP_DIM + prow + 1], v0 = pl[HP_DIM + prow];
                a0 = sad14<1>(h, s0, s1, s2, s3);
                a1 = sad14<0>(h, s0, s1, s2, s3);
                a2 = sad14<1>(v1, s0, s1, s2, s3);
                a3 = sad14<1>(v0, s0, s1, s2, s3);
            } else {
                const uint4 x0 = pl[2 * HP_DIM + prow], x1 = pl[2 * HP_DIM + prow + 1];
                a0 = sad14<0>(x0, s0, s1, s2, s3);
                a1 = sad14<1>(x0, s0, s1, s2, s3);
                a2 = sad14<0>(x1, s0, s1, s2, s3);
                a3 = sad14<1>(x1, s0, s1, s2, s3);
            }
        }
        const bool hv = lane >= 16;
        unsigned acc[8] = {hv ? 0u : a0, hv ? 0u : a1, hv ? 0u : a2, hv ? 0u : a3, hv ? a0 : 0u, hv ? a1 : 0u, hv ? a2 : 0u, hv ? a3 : 0u};
        warp_reduce_n<8>(acc);
        int m = -1;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (best_hp > (int) acc[k]) {
                best_hp = (int) acc[k];
                m = k;
            }
        }
        mvx <<= 1;
        mvy <<= 1;
        if (m != -1) {
            mvx += offset_of(HP_XP, m);
            mvy += offset_of(HP_YP, m);
            /* the winner's samples: its plane from the first row (yh >= 0) and byte (xh >= 0) on */
            const int plane = m < 2 ? 0 : (m < 4 ? 1 : 2), dy = (0xc7 >> m) & 1, dx = (0xad >> m) & 1;
            const unsigned *pw = reinterpret_cast<const unsigned *>(s_tmp + plane * HP_PLANE + dy * HP_DIM);
            for (int k = lane; k < HP_SAD_SZ * 4; k += 32) {
                const unsigned lo = pw[k], hi = pw[k + 1];
                s_refblk_w[k] = dx ? __funnelshift_r(lo, hi, 8) : lo;
            }
            has_hp = true;
            best = (int) ((unsigned) best_hp * yarea / (unsigned) (HP_SAD_SZ * HP_SAD_SZ));
        }
    } else {
        mvx <<= 1;
        mvy <<= 1;
    }
    if (!has_hp) {
        const uint8_t *rp = A.ref.p + (ptrdiff_t) (cy + (mvy >> 1)) * rs + cx + (mvx >> 1);
        for (int k = lane; k < HP_SAD_SZ * 4; k += 32) {
            s_refblk_w[k] = ld4g(rp + (ptrdiff_t) (k >> 2) * rs + 4 * (k & 3));
        }
    }
    __syncwarp();

    /* ---- block sums ---- */
    unsigned sum[SUM_COUNT];
#pragma unroll
    for (int k = 0; k < SUM_COUNT; k++) {
        sum[k] = 0;
    }
    const int sbw = G.bw / 2, sbh = G.bh / 2;
    const uint8_t *ref0 = A.ref.p + (ptrdiff_t) G.by * rs + G.bx; /* zero-MV reference block */
    if ((G.bw & 7) == 0) {
        /* four samples per step: the block sums of block_analysis / y_sqrvar are sums of absolute differences and
         * squares, i.e. __vsadu4 / __dp4a on packed words.  A lane walks its rows downwards, so the row above is last
         * step's word; the byte right of a word comes from the neighbouring lane. */
        const int lx0 = 4 * G.wx;
        const int rbase = G.active ? G.r0 : 0; /* idle lanes walk along (shuffles are warp-wide) */
        unsigned w_up = 0;
        if (G.active && G.r0 > 0 && G.r0 < G.bh) {
            w_up = *reinterpret_cast<const unsigned *>(s_src + (G.r0 - 1) * HME_SRC_STRIDE + lx0);
        }
        for (int it = 0; it < G.rpg; it++) {
            const int ly = rbase + it;
            const bool on = G.active && ly < G.r1;
            unsigned w = 0, r = 0;
            if (on) {
                w = *reinterpret_cast<const unsigned *>(s_src + ly * HME_SRC_STRIDE + lx0);
                r = ld4g(ref0 + (ptrdiff_t) ly * rs + lx0);
            }
            const unsigned w_r = __shfl_down_sync(0xffffffffu, w, 1);
            if (on) {
                const unsigned nxt = lx0 + 4 < G.bw ? (w_r & 0xffu) : (w >> 24);
                const unsigned wr = (w >> 8) | (nxt << 24);
                const unsigned up = ly == 0 ? w : w_up;
                sum[SUM_S] += __vsadu4(w, 0u);
                sum[SUM_SS] = __dp4a(w, w, sum[SUM_SS]);
                sum[SUM_SH] += __vsadu4(w, wr);
                sum[SUM_SV] += __vsadu4(w, up);
                sum[SUM_RS] += __vsadu4(r, 0u);
                sum[SUM_RSS] = __dp4a(r, r, sum[SUM_RSS]);
                w_up = w;
            }
        }
    } else {
        /* general widths (blocks cut by the picture edge): rows in turn, columns to lanes */
        for (int ly = 0; ly < G.bh; ly++) {
            for (int lx = lane; lx < G.bw; lx += 32) {
                const uint8_t *sp = s_src + ly * HME_SRC_STRIDE + lx;
                const int pa = sp[0], pb = ref0[(ptrdiff_t) ly * rs + lx];
                const int right = lx == G.bw - 1 ? pa : sp[1];
                const int up = ly == 0 ? pa : sp[-HME_SRC_STRIDE];
                sum[SUM_S] += (unsigned) pa;
                sum[SUM_SS] += (unsigned) (pa * pa);
                sum[SUM_SH] += (unsigned) iabs(pa - right);
                sum[SUM_SV] += (unsigned) iabs(pa - up);
                sum[SUM_RS] += (unsigned) pb;
                sum[SUM_RSS] += (unsigned) (pb * pb);
            }
        }
    }
    warp_chroma_moments(A, i * (A.blk_w >> A.hs), j * (A.blk_h >> A.vs), G.bw >> A.hs, G.bh >> A.vs, lane, &sum[SUM_CSU]);
    { /* block_texture on the two 14x14 patches (hme.c:179-209): lanes 0-13 a row of the source patch each, lanes
       * 16-29 a row of the chosen reference patch; the row above comes from the lane below */
        unsigned t0 = s0, t1 = s1, t2 = s2, t3 = s3;
        if (lane >= 16 && prow < HP_SAD_SZ) {
            const uint4 r = reinterpret_cast<const uint4 *>(s_refblk)[prow];
            t0 = r.x;
            t1 = r.y;
            t2 = r.z;
            t3 = r.w & 0xffffu;
        }
        unsigned u0 = __shfl_up_sync(0xffffffffu, t0, 1), u1 = __shfl_up_sync(0xffffffffu, t1, 1);
        unsigned u2 = __shfl_up_sync(0xffffffffu, t2, 1), u3 = __shfl_up_sync(0xffffffffu, t3, 1);
        if (prow == 0) {
            u0 = t0;
            u1 = t1;
            u2 = t2;
            u3 = t3;
        }
        if (prow < HP_SAD_SZ) {
            const unsigned last = t3 >> 8; /* sample 13 is its own right neighbour */
            const unsigned th = __vsadu4(t0, __funnelshift_r(t0, t1, 8)) + __vsadu4(t1, __funnelshift_r(t1, t2, 8)) +
                                __vsadu4(t2, __funnelshift_r(t2, t3, 8)) + __vsadu4(t3, last | (last << 8));
            const unsigned tv = __vsadu4(t0, u0) + __vsadu4(t1, u1) + __vsadu4(t2, u2) + __vsadu4(t3, u3);
            const unsigned av = __vsadu4(t0, 0u) + __vsadu4(t1, 0u) + __vsadu4(t2, 0u) + __vsadu4(t3, 0u);
            const unsigned avs = __dp4a(t0, t0, __dp4a(t1, t1, __dp4a(t2, t2, __dp4a(t3, t3, 0u))));
            const bool q = lane >= 16;
            sum[SUM_PSH] += q ? 0u : th;
            sum[SUM_PSV] += q ? 0u : tv;
            sum[SUM_PAV] += q ? 0u : av;
            sum[SUM_PAVS] += q ? 0u : avs;
            sum[SUM_QSH] += q ? th : 0u;
            sum[SUM_QSV] += q ? tv : 0u;
            sum[SUM_QAV] += q ? av : 0u;
            sum[SUM_QAVS] += q ? avs : 0u;
        }
    }
    warp_reduce_n<SUM_COUNT>(sum);

    /* ---- the decision cascade (hme.c:651-718).  The sums are warp-uniform, so every lane runs it and the rare intra
     * blocks take the two passes only they need (the range test and the quadrant metric) as a whole warp ---- */
    const unsigned area = yarea;
    const unsigned luma_tex = ((sum[SUM_SH] + sum[SUM_SV]) / 2u) / area;
    const unsigned luma_var = sum[SUM_SS] - (sum[SUM_S] * sum[SUM_S]) / area;
    const int lo_tex = luma_tex <= 2, lo_var = luma_var < yareasq;
    const unsigned pn = HP_SAD_SZ * HP_SAD_SZ;
    const int src_tex = (int) (((sum[SUM_PSH] + sum[SUM_PSV]) / 2u) / pn);
    const int src_avg = (int) (sum[SUM_PAV] / pn);
    const int src_var = (int) (sum[SUM_PAVS] - (sum[SUM_PAV] * sum[SUM_PAV]) / pn);
    const int ref_tex = (int) (((sum[SUM_QSH] + sum[SUM_QSV]) / 2u) / pn);
    const int ref_avg = (int) (sum[SUM_QAV] / pn);
    const int ref_var = (int) (sum[SUM_QAVS] - (sum[SUM_QAV] * sum[SUM_QAV]) / pn);
    const unsigned ubest = (unsigned) best;
    bool intra = false;
    if (src_tex < 2 && (sum[SUM_RSS] - (sum[SUM_RS] * sum[SUM_RS]) / area) > luma_var * 2u) {
        intra = true;
    } else if (ref_var > src_var * 2) {
        intra = true;
    } else if (src_tex == 0 && ref_tex != 0) {
        intra = true;
    } else if (iabs(src_avg - ref_avg) > 8) {
        intra = true;
    } else if (luma_tex <= 10 && ubest > yareasq / 16u) {
        intra = true;
    } else {
        const unsigned carea = (unsigned) ((G.bw >> A.hs) * (G.bh >> A.vs));
        if (carea) {
            const unsigned vsu = sum[SUM_CSSU] - (sum[SUM_CSU] * sum[SUM_CSU]) / carea;
            const unsigned vsv = sum[SUM_CSSV] - (sum[SUM_CSV] * sum[SUM_CSV]) / carea;
            const unsigned vru = sum[SUM_CRSU] - (sum[SUM_CRU] * sum[SUM_CRU]) / carea;
            const unsigned vrv = sum[SUM_CRSV] - (sum[SUM_CRV] * sum[SUM_CRV]) / carea;
            const unsigned cvarS = vsu > vsv ? vsu : vsv, cvarR = vru > vrv ? vru : vrv;
            intra = cvarR > 4u * cvarS;
        }
    }
    int mode = 0, submask = 0;
    if (intra) {
        /* block_intra_test (hme.c:141-177): any sample the reduced-range intra path cannot represent.
         * clamp_u8(ravg + clamp_u8(p - ravg + 128) - 128) != p  <=>  the inner clamp is active  <=>  p outside
         * [ravg - 128, ravg + 127]; four samples per step from the staged words */
        const int ravg = (int) sum[SUM_RS] / (G.bw * G.bh);
        const int lo = ravg - 128, hi = ravg + 127;
        int bad = 0;
        if (lo > 0 || hi < 255) {
            const int tail = G.bw & 3;
            const int nb = (G.wx == G.words - 1 && tail) ? tail : 4;
            if (tail == 0) { /* whole words: only one of the two bounds can lie inside 0..255 */
                const unsigned bound = (unsigned) (lo > 0 ? lo : hi) * 0x01010101u;
                unsigned out = 0;
                for (int r = G.r0; r < G.r1; r++) {
                    const unsigned w = *reinterpret_cast<const unsigned *>(s_src + r * HME_SRC_STRIDE + 4 * G.wx);
                    out |= lo > 0 ? __vcmpltu4(w, bound) : __vcmpgtu4(w, bound);
                }
                bad = out != 0u;
            } else {
                for (int r = G.r0; r < G.r1; r++) {
                    const unsigned w = *reinterpret_cast<const unsigned *>(s_src + r * HME_SRC_STRIDE + 4 * G.wx);
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int p = byte_of(w, e);
                        bad |= (e < nb) & ((p < lo) | (p > hi));
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, bad)) {
            submask = 15;
            if (src_tex > 1) {
                unsigned ge[8];
                quadrant_metric(G, s_src, ref0, rs, lane, ge);
                const unsigned wgt = (unsigned) ((sbw + sbh) >> 1);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (ge[q] >= wgt * ge[4 + q]) {
                        submask &= ~(1 << q);
                    }
                }
            }
            if (submask) {
                mode = 1;
                if (lane == 0) {
                    atomicAdd(A.nintra, 1);
                }
            }
        }
    }
    if (lane == 0) {
        store_mv(out, mvx, mvy, mode, submask, lo_var, lo_tex);
        A.aux[i + j * A.nbh] = make_int2((int) luma_tex, src_var);
    }
}

/* high_detail from the already-final left / top / top-left neighbours (hme.c:607-648) */
__global__ void __launch_bounds__(128) hme_neigh_kernel(const HmeArgs *args)
{
    const HmeArgs &A = args[blockIdx.z];
    DevMV *mv = A.out;
    const int2 *aux = A.aux;
    const int nbh = A.nbh, nbv = A.nbv;
    const int b = (int) (blockIdx.x * blockDim.x + threadIdx.x);
    if (b >= nbh * nbv) {
        return;
    }
    const int i = b % nbh, j = b / nbh;
    unsigned thresh_tex = 1;
    int thresh_var = HP_SAD_SZ * HP_SAD_SZ;
    auto detailed = [&](int x, int y) {
        const DevMV m = mv[y * nbh + x];
        return m.mode == 0 && !m.lo_tex && !m.lo_var;
    };
    if (i > 0 && detailed(i - 1, j)) {
        thresh_var *= HP_SAD_SZ;
        thresh_tex++;
    }
    if (j > 0 && detailed(i, j - 1)) {
        thresh_var *= HP_SAD_SZ;
        thresh_tex++;
    }
    if (i > 0 && j > 0 && detailed(i - 1, j - 1)) {
        thresh_var *= HP_SAD_SZ / 4;
        thresh_tex++;
    }
    const int2 a = aux[b];
    mv[b].high_detail = (uint8_t) ((unsigned) a.x > thresh_tex && a.y > thresh_var);
}

static HmePlane mk_plane(const DevFrame &f, int c)
{
    HmePlane p;
    p.p = f.p[c];
    p.stride = f.stride[c];
    p.w = f.w[c];
    p.h = f.h[c];
    return p;
}

void hme_fill_args(HmeArgs *A, const MotionGeom &g, int level, const DevFrame *src, const DevFrame *ref,
                   DevMV *const *mvf, int2 *aux, int *d_nintra)
{
    memset(A, 0, sizeof(*A));
    A->src = mk_plane(src[level], 0);
    A->ref = mk_plane(ref[level], 0);
    A->parent = level < g.levels ? mvf[level + 1] : nullptr;
    A->out = mvf[level];
    A->aux = aux;
    A->nintra = d_nintra;
    A->level = level;
    A->blk_w = g.blk_w;
    A->blk_h = g.blk_h;
    A->nbh = g.nbh;
    A->nbv = g.nbv;
    A->hs = g.hs;
    A->vs = g.vs;
    if (level == 0) {
        A->srcU = mk_plane(src[0], 1);
        A->srcV = mk_plane(src[0], 2);
        A->refU = mk_plane(ref[0], 1);
        A->refV = mk_plane(ref[0], 2);
    }
}

/* d_args[level * n + lane], levels g.levels .. 0; *nintra of every lane must be zero on entry */
void hme_launch(const HmeArgs *d_args, int n, const MotionGeom &g, cudaStream_t st)
{
    if (n <= 0) {
        return;
    }
    for (int level = g.levels; level >= 0; level--) {
        const HmeArgs *a = d_args + (size_t) level * n;
        if (level > 0) {
            const int step = 1 << level;
            const int nbx = ceil_div(g.nbh, step), nby = ceil_div(g.nbv, step);
            DSV_LAUNCH(hme_level_kernel, dim3(ceil_div(nbx * nby, HME_WARPS), 1, n), dim3(HME_THREADS), 0, st, a, nbx, nby);
        } else {
            DSV_LAUNCH(hme_l0_kernel, dim3(ceil_div(g.nbh * g.nbv, HME_WARPS), 1, n), dim3(HME_THREADS), 0, st, a);
        }
        KERNEL_CHECK();
    }
    DSV_LAUNCH(hme_neigh_kernel, dim3(ceil_div(g.nbh * g.nbv, 128), 1, n), dim3(128), 0, st, d_args);
    KERNEL_CHECK();
}

} // namespace dsv
