/*
 * hzcc_enc.cu -- HZCC coefficient coder, encoder side: hzcc_enc's token stream (hzcc.c:137-293)
 * and dsv_encode_plane's framing (hzcc.c:449-476) as three data-parallel passes over the scan order.
 *
 *   hzcc_scan_kernel    one warp per 2048-position chunk: coalesced 16-byte zero tests mark the groups of
 *                       4 positions that hold something, the marked groups are visited in scan order
 *                       (the quantised symbol of a non-zero is re-derived from the dequantised
 *                       coefficient the SBT epilogue stored): count non-zeros, find the chunk's
 *                       first/last non-zero and the bits of all groups that are fully determined
 *                       inside the chunk.
 *   hzcc_prefix_kernel  one CTA per frame: exclusive scans over the chunk summaries (previous
 *                       non-zero, bit offsets), plane framing (plen, SEG(DC), nruns, final NEG,
 *                       0x55) and plane base offsets; yields the packet length.
 *   hzcc_pack_kernel    per chunk again: every non-zero ORs its group (UEG(run) ++ NEG(prev))
 *                       at its bit offset into the zeroed packet (MSB-first, bs.c:76-91).  With a
 *                       list scratch (HzJob.dense, the encoder always passes one) the scan pass has
 *                       left the chunk's non-zeros there -- one (offset, symbol) list in scan order
 *                       for a sparse chunk, per-lane lists for a dense one (those are emitted by
 *                       hzcc_pack_dense_kernel) -- and no coefficient is read a second time.
 */
#include "hzcc.cuh"
#include "scan.cuh"
#include "quant.cuh"

namespace dsv {

void hz_fill_regions(HzRegions *r, int cw, int ch)
{
    int n = 0, pos = 0;
    auto add = [&](int x0, int y0, int sw, int sh, int lvl) {
        r->base[n] = pos;
        r->x0[n] = x0; r->y0[n] = y0; r->sw[n] = sw; r->sh[n] = sh; r->lvl[n] = lvl;
        r->fdw[n] = make_fastdiv(sw);
        pos += sw * sh;
        n++;
    };
    add(0, 0, ceil_shift(cw, 3), ceil_shift(ch, 3), 4);
    for (int l = 0; l < 3; l++) {
        int sw = ceil_shift(cw, 3 - l), sh = ceil_shift(ch, 3 - l);
        add(sw, 0, sw, sh, 3 - l);
        add(0, sh, sw, sh, 3 - l);
        add(sw, sh, sw, sh, 3 - l);
    }
    r->base[n] = pos;
}

void hz_fill_job(HzJob *j, int cw, int ch, int q, int isP, int plane, int nbh, int nbv)
{
    SbtJob t;
    memset(&t, 0, sizeof(t));
    sbt_fill_geometry(&t, cw, ch, cw, ch, isP, plane);
    sbt_fill_quant(&t, q, isP, plane, nbh, nbv);
    j->cw = cw;
    j->ch = ch;
    j->plane = plane;
    j->isP = isP;
    j->pq = t.pq;
    j->dg = t.dg;
    hz_fill_regions(&j->rg, cw, ch);
    j->nchunks = ceil_div(j->rg.base[HZ_NREG], HZ_CHUNK);
}

/*
 * Stand-alone quantise + dequantise of a raw coefficient plane, in place -- the same emit_h the SBT
 * epilogue uses, for callers that hand in coefficients directly (dsv_encode_plane semantics,
 * hzcc.c:449-476).  One thread per coefficient.
 */
__global__ void __launch_bounds__(256) hzcc_quant_kernel(const HzJob *jobs)
{
    const HzJob &J = jobs[blockIdx.y];
    const int i = (int) (blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= J.cw * J.ch) {
        return;
    }
    const int ax = i % J.cw, ay = i / J.cw;
    for (int lvl = 1; lvl < 32; lvl++) {
        const int wo = sbt_wo(J.cw, lvl), ho = sbt_wo(J.ch, lvl);
        if (ax >= wo || ay >= ho) {
            const int band = (ax >= wo ? 1 : 0) | (ay >= ho ? 2 : 0);
            emit_h(J, true, J.stable, lvl, band, ax - ((band & 1) ? wo : 0), ay - ((band & 2) ? ho : 0), J.coef[i]);
            return;
        }
        if (wo == 1 && ho == 1) {
            return; /* (0,0): DC */
        }
    }
}

void hzcc_quant_launch(const HzJob *d_jobs, int njobs, int max_elems, cudaStream_t st)
{
    DSV_LAUNCH(hzcc_quant_kernel, dim3(ceil_div(max_elems, 256), njobs), dim3(256), 0, st, d_jobs);
    KERNEL_CHECK();
}

/* chunk -> job.  The engines launch the planes Y,U,V of many pictures of ONE format: chunk counts repeat with period
 * 3, so the job is arithmetic (map.per_pic > 0); a per-CTA binary search over the job table costs several dependent
 * global loads -- more than half the lifetime of a CTA that handles 2048 coefficients.  The search remains for
 * irregular launches (kernel-level API). */
DSV_D int hz_job_of_chunk(const HzJob *jobs, int njobs, int chunk, const HzMap &map)
{
    if (map.per_pic > 0) {
        const int pic = (int) fastdiv((unsigned) chunk, map.per_pic_fd), r = chunk - pic * map.per_pic;
        return 3 * pic + (r >= map.c0) + (r >= map.c0 + map.c1);
    }
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].chunk_base <= chunk) {
            lo = mid;
        } else {
            hi = mid - 1;
        }
    }
    return lo;
}

/* position inside the scan order: region + region-local coordinates */
struct HzCursor {
    int r, x, y;
};

DSV_D void hz_locate(const HzRegions &rg, int s, HzCursor &c)
{
    int r = 0;
    while (r < HZ_NREG - 1 && s >= rg.base[r + 1]) {
        r++;
    }
    const int k = s - rg.base[r];
    c.r = r;
    c.y = (int) fastdiv((unsigned) k, rg.fdw[r]);
    c.x = k - c.y * rg.sw[r];
}

DSV_D void hz_advance(const HzRegions &rg, HzCursor &c)
{
    if (++c.x == rg.sw[c.r]) {
        c.x = 0;
        if (++c.y == rg.sh[c.r]) {
            c.y = 0;
            c.r++;
        }
    }
}

/* quantised symbol of the coefficient v != 0 at (x, y) of region r: the LL region away from the DC, or a detail band
 * position that is not the first visit of a double-visited row / column */
DSV_D int hz_symbol_plain(const HzJob &J, int r, int x, int y, int v)
{
    if (r == 0) {
        return dz_quant(v, J.pq.ll_q, J.pq.ll_fd);
    }
    const int lvl = J.rg.lvl[r];
    int f = J.stable[((y * J.pq.dby[lvl]) >> 14) * J.pq.nbh + ((x * J.pq.dbx[lvl]) >> 14)];
    if (lvl == 1) {
        return p2_quant(v, f ? J.pq.sh_hq : J.pq.sh_plain);
    }
    int sel = (f & 2) ? 2 : (f ? 1 : 0);
    const LevelQ &Lq = J.pq.lv[3 - lvl];
    return dz_quant(v, Lq.q[sel], Lq.fd[sel]);
}

/* quantised symbol at the cursor (which must be inside the plane's scan order): re-derived from the dequantised
 * coefficient the SBT epilogue stored; zero coefficients (the vast majority) cost one load and a compare */
DSV_D int hz_symbol_at(const HzJob &J, const HzCursor &c)
{
    const HzRegions &rg = J.rg;
    const int r = c.r, x = c.x, y = c.y;
    const int ax = rg.x0[r] + x, ay = rg.y0[r] + y;
    const int lvl = rg.lvl[r];
    if (r == 0) {
        if ((x | y) == 0) {
            return 0; /* DC travels separately (hzcc.c:166,462-465) */
        }
        int v = J.coef[(size_t) ay * J.cw + ax];
        return v ? dz_quant(v, J.pq.ll_q, J.pq.ll_fd) : 0;
    }
    if (lvl >= 2) { /* first visit of a position that the next hzcc level scans again */
        const DvGeom &g = J.dg;
        const int L = lvl - 1;
        if (g.dvx[L] >= 0 || g.dvy[L] >= 0) {
            bool col = (ax == g.dvx[L]) && (ay < g.dvey[L]);
            bool row = (ay == g.dvy[L]) && (ax < g.dvex[L]);
            if (col || row) {
                return J.dv[col ? g.col_base[L] + ay : g.row_base[L] + ax];
            }
        }
    }
    int v = J.coef[(size_t) ay * J.cw + ax];
    return v ? hz_symbol_plain(J, r, x, y, v) : 0;
}
/* the symbols of the four positions of a group hz_group_plain accepted, from its one 16-byte load: the four
 * derivations (block-flag load + division) are independent of each other */
DSV_D void hz_symbols_of_group(const HzJob &J, const HzCursor &c, const int4 &v, int (&sym)[4])
{
    const int vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; e++) {
        sym[e] = vv[e] ? hz_symbol_plain(J, c.r, c.x + e, c.y, vv[e]) : 0;
    }
}

DSV_D unsigned long long mk_key(int pos, int sym) { return ((unsigned long long) (unsigned) pos << 32) | (unsigned) sym; }
DSV_D int key_pos(unsigned long long k) { return (int) (k >> 32); }
DSV_D int key_sym(unsigned long long k) { return (int) (unsigned) k; }
#define KEY_NONE 0xFFFFFFFF00000000ull /* pos = -1 */

DSV_D unsigned group_bits(int pos, int prev_pos, int prev_sym)
{
    unsigned b = (unsigned) ueg_len((unsigned) (pos - prev_pos - 1));
    if (prev_pos >= 0) {
        b += (unsigned) neg_len(prev_sym);
    }
    return b;
}

/*
 * One WARP per 2048-position chunk, in two phases.  (The previous formulation, 8 positions per thread with block-wide
 * scans, spent ~50 instructions per position on set-up and barriers: 97 000 CTAs per launch, each caching the job
 * record for 2048 coefficients, while well under 1 % of a P picture's symbols are non-zero.)
 *
 *   sweep   the chunk as 512 groups of 4 scan positions; group 32k + l belongs to lane l in step k, so a step's 32
 *           groups are 512 contiguous bytes wherever the chunk stays inside one region row: coalesced 16-byte loads,
 *           all 16 steps independent.  A group that is contiguous and 16-byte aligned in the coefficient plane
 *           (inside one row of one region, away from the DC and the double-visited row / column) and all zero is
 *           done; a ballot per step leaves a 512-bit map of the groups that need a look.
 *   rounds  32 marked groups at a time, in scan order: a lane derives the symbols of its group's positions
 *           (hz_symbol_at re-derives the quantised symbol from the dequantised coefficient the SBT epilogue stored)
 *           and the warp strings the non-zeros together with shuffles; the visitor (scan: summary, pack: bit
 *           offsets + emission) sees every lane's up-to-4 non-zeros.
 */
#define HZW_WARPS (HZ_THREADS / 32)
#define HZW_STEPS (HZ_CHUNK / 128)
#define HZW_ITEMS (HZ_CHUNK / 32)
#ifndef HZW_BATCH
#define HZW_BATCH 4 /* sweep loads in flight per lane */
#endif
#ifndef HZW_DENSE
#define HZW_DENSE 128 /* marked groups (of 512) from which a chunk is walked lane by lane (see hz_walk); A/B on 64 HD I
                         * pictures: 64 -> scan 820 us, 128 -> 822 us, 320 -> 1180 us (+1.2 ms in the sparse pack kernel);
                         * requesting the walk's next group one step ahead changes nothing (816 us) */
#endif

DSV_D int nth_set_bit(unsigned m, int n)
{
    for (int i = 0; i < n; i++) {
        m &= m - 1;
    }
    return __ffs((int) m) - 1;
}

/* can the 4 positions starting at the cursor be zero-tested with one aligned 16-byte load? (then *p is its address) */
DSV_D bool hz_group_plain(const HzJob &J, const HzCursor &c, const int32_t **p)
{
    const HzRegions &rg = J.rg;
    const int r = c.r;
    if (c.x + 4 > rg.sw[r] || (r == 0 && (c.x | c.y) == 0)) {
        return false;
    }
    const int ax = rg.x0[r] + c.x, ay = rg.y0[r] + c.y, lvl = rg.lvl[r];
    if (r > 0 && lvl >= 2) { /* first visits of double-visited positions live in the side buffer */
        const DvGeom &g = J.dg;
        const int L = lvl - 1;
        if ((g.dvx[L] >= ax && g.dvx[L] < ax + 4 && ay < g.dvey[L]) || (ay == g.dvy[L] && ax < g.dvex[L])) {
            return false;
        }
    }
    *p = J.coef + (size_t) ay * J.cw + ax;
    return (reinterpret_cast<uintptr_t>(*p) & 15) == 0;
}

/* a lane's group in a round: its non-zeros in scan order */
struct HzGroup {
    int cnt;
    int pos[4], sym[4];
};

template <class Visitor> DSV_D void hz_chunk_rounds(const HzJob &J, int cbase, int total, int lane, Visitor &V)
{
    const HzRegions &rg = J.rg;
    unsigned marks[HZW_STEPS];
    {
        HzCursor c;
        int pos = cbase + 4 * lane;
        if (pos < total) {
            hz_locate(rg, pos, c);
        }
        /* HZW_BATCH loads are requested before the first of them is looked at: a chunk then waits for
         * HZW_STEPS / HZW_BATCH memory latencies instead of one per step */
#pragma unroll
        for (int kb = 0; kb < HZW_STEPS; kb += HZW_BATCH) {
            int4 v[HZW_BATCH];
            int st[HZW_BATCH]; /* 0: past the end, 1: has to be looked at, 2: loaded */
#pragma unroll
            for (int i = 0; i < HZW_BATCH; i++) {
                st[i] = 0;
                v[i] = make_int4(0, 0, 0, 0);
                if (pos < total) {
                    st[i] = 1;
                    const int32_t *p;
                    if (pos + 4 <= total && hz_group_plain(J, c, &p)) {
                        v[i] = *reinterpret_cast<const int4 *>(p);
                        st[i] = 2;
                    }
                    /* 128 positions on */
                    pos += 128;
                    if (pos < total) {
                        c.x += 128;
                        while (c.x >= rg.sw[c.r]) {
                            c.x -= rg.sw[c.r];
                            if (++c.y == rg.sh[c.r]) {
                                c.y = 0;
                                c.r++;
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < HZW_BATCH; i++) {
                const bool mark = st[i] == 1 || (st[i] == 2 && (v[i].x | v[i].y | v[i].z | v[i].w) != 0);
                marks[kb + i] = __ballot_sync(0xffffffffu, mark);
            }
        }
    }
    int G = 0;
#pragma unroll
    for (int k = 0; k < HZW_STEPS; k++) {
        G += __popc(marks[k]);
    }
    if (G >= HZW_DENSE) {
        V.dense(J, cbase + lane * HZW_ITEMS, imin(total, cbase + HZ_CHUNK), lane);
        return;
    }
    for (int g0 = 0; g0 < G; g0 += 32) {
        HzGroup grp;
        grp.cnt = 0;
        int n = g0 + lane;
        if (n < G) {
            int k = 0;
#pragma unroll
            for (int kk = 0; kk < HZW_STEPS; kk++) { /* marks[] is indexed by constants only: it stays in registers */
                const int c = __popc(marks[kk]);
                if (n >= 0 && n < c) {
                    k = 128 * kk + 4 * nth_set_bit(marks[kk], n);
                    n = -1;
                } else if (n >= 0) {
                    n -= c;
                }
            }
            const int pos = cbase + k;
            HzCursor c;
            hz_locate(rg, pos, c);
            const int e_end = imin(4, total - pos);
            const int32_t *p;
            if (e_end == 4 && hz_group_plain(J, c, &p)) {
                int sy[4];
                hz_symbols_of_group(J, c, *reinterpret_cast<const int4 *>(p), sy);
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    if (sy[e]) {
                        grp.pos[grp.cnt] = pos + e;
                        grp.sym[grp.cnt] = sy[e];
                        grp.cnt++;
                    }
                }
            } else {
                for (int e = 0; e < e_end; e++) {
                    const int sym = hz_symbol_at(J, c);
                    hz_advance(rg, c);
                    if (sym) {
                        grp.pos[grp.cnt] = pos + e;
                        grp.sym[grp.cnt] = sym;
                        grp.cnt++;
                    }
                }
            }
        }
        V.round(grp, lane);
    }
}

/* last non-zero of the lanes before this one (KEY_NONE if there is none); *all = the last of the whole warp */
DSV_D unsigned long long warp_prev_key(unsigned long long key, int lane, unsigned long long *all)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long n = __shfl_up_sync(0xffffffffu, key, o);
        if (lane >= o) {
            key = OpMaxS64::apply(n, key);
        }
    }
    *all = __shfl_sync(0xffffffffu, key, 31);
    const unsigned long long ex = __shfl_up_sync(0xffffffffu, key, 1);
    return lane == 0 ? KEY_NONE : ex;
}

/* what a lane contributes to its chunk: its non-zeros' count, first position, last (position, symbol), and the bits of
 * the groups whose predecessor is the lane's own */
struct HzSummary {
    int cnt, first_pos;
    unsigned long long last_key;
    unsigned bits;
};
DSV_D HzSummary hz_group_summary(const HzGroup &g)
{
    HzSummary s;
    s.cnt = g.cnt;
    s.first_pos = g.cnt ? g.pos[0] : -1;
    s.last_key = KEY_NONE;
    s.bits = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        if (e < g.cnt) {
            s.last_key = mk_key(g.pos[e], g.sym[e]);
            if (e > 0) {
                s.bits += group_bits(g.pos[e], g.pos[e - 1], g.sym[e - 1]);
            }
        }
    }
    return s;
}

/*
 * Dense chunks (I pictures: most groups hold something) skip the rounds: a lane walks its own 64 consecutive scan
 * positions, so the per-round shuffles are paid once per chunk instead of once per 32 groups.
 */

template <class F> DSV_D void hz_walk(const HzJob &J, int base, int total, F on_nonzero)
{
    if (base >= total) {
        return;
    }
    const HzRegions &rg = J.rg;
    const int end = imin(base + HZW_ITEMS, total);
    HzCursor c;
    hz_locate(rg, base, c);
    int pos = base;
    while (pos < end) {
        const int32_t *p;
        if (pos + 4 <= end && hz_group_plain(J, c, &p)) {
            const int4 v = *reinterpret_cast<const int4 *>(p);
            if ((v.x | v.y | v.z | v.w) != 0) {
                int sy[4];
                hz_symbols_of_group(J, c, v, sy);
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    if (sy[e]) {
                        on_nonzero(pos + e, sy[e]);
                    }
                }
            }
            pos += 4;
            c.x += 4;
            if (c.x == rg.sw[c.r]) {
                c.x = 0;
                if (++c.y == rg.sh[c.r]) {
                    c.y = 0;
                    c.r++;
                }
            }
            continue;
        }
        const int sym = hz_symbol_at(J, c);
        hz_advance(rg, c);
        if (sym) {
            on_nonzero(pos, sym);
        }
        pos++;
    }
}
DSV_D HzSummary hz_walk_summary(const HzJob &J, int base, int total)
{
    HzSummary s;
    s.cnt = 0;
    s.first_pos = -1;
    s.last_key = KEY_NONE;
    s.bits = 0;
    int prev_pos = -1, prev_sym = 0;
    hz_walk(J, base, total, [&](int pos, int sym) {
        if (prev_pos >= 0) {
            s.bits += group_bits(pos, prev_pos, prev_sym);
        } else {
            s.first_pos = pos;
        }
        prev_pos = pos;
        prev_sym = sym;
        s.cnt++;
    });
    if (s.cnt) {
        s.last_key = mk_key(prev_pos, prev_sym);
    }
    return s;
}

/* chunk summary, accumulated over rounds (or one walk) */
struct HzScanAcc {
    unsigned long long carry = KEY_NONE; /* last non-zero of the chunk so far */
    unsigned cnt = 0, bits = 0;
    int first = -1;
    DSV_D void add(const HzSummary &s, int lane)
    {
        unsigned long long all;
        unsigned long long ex = warp_prev_key(s.last_key, lane, &all);
        if (key_pos(ex) < 0) {
            ex = carry;
        }
        unsigned b = s.bits;
        int f = 0x7fffffff;
        if (s.cnt) {
            if (key_pos(ex) >= 0) {
                b += group_bits(s.first_pos, key_pos(ex), key_sym(ex));
            } else {
                f = s.first_pos; /* the chunk's first non-zero: its group depends on earlier chunks */
            }
        }
        cnt += __reduce_add_sync(0xffffffffu, (unsigned) s.cnt);
        bits += __reduce_add_sync(0xffffffffu, b);
        f = __reduce_min_sync(0xffffffffu, f);
        if (f != 0x7fffffff) {
            first = f;
        }
        carry = OpMaxS64::apply(carry, all);
    }
};
struct HzScanVisitor {
    HzScanAcc acc;
    int saved = 0;           /* which lists of this chunk were written to the job's scratch (HzChunk.dense) */
    uint8_t *sparse = nullptr; /* the chunk's scratch if sparse lists are wanted */
    int nlist = 0;
    DSV_D void round(const HzGroup &g, int lane)
    {
        if (sparse) { /* append the round's non-zeros, in scan order: lanes in turn, a lane's group in order */
            unsigned incl = (unsigned) g.cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) {
                    incl += n;
                }
            }
            const int at = nlist + (int) incl - g.cnt;
            uint16_t *lp = reinterpret_cast<uint16_t *>(sparse + HZ_DENSE_OFF);
            int *ls = reinterpret_cast<int *>(sparse + HZ_DENSE_SYM);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (e < g.cnt) {
                    lp[at + e] = (uint16_t) (g.pos[e] & (HZ_CHUNK - 1));
                    ls[at + e] = g.sym[e];
                }
            }
            nlist += (int) __shfl_sync(0xffffffffu, incl, 31);
            saved = 2;
        }
        acc.add(hz_group_summary(g), lane);
    }
    DSV_D void dense(const HzJob &J, int base, int total, int lane)
    {
        if (!J.dense || J.list_mode == HZ_LISTS_SPARSE) {
            acc.add(hz_walk_summary(J, base, total), lane);
            return;
        }
        uint8_t *sc = J.dense + (size_t) ((base - lane * HZW_ITEMS) / HZ_CHUNK) * HZ_DENSE_BYTES;
        uint8_t *off = sc + HZ_DENSE_OFF + lane;
        int *sym = reinterpret_cast<int *>(sc + HZ_DENSE_SYM) + lane;
        HzSummary s;
        s.cnt = 0;
        s.first_pos = -1;
        s.last_key = KEY_NONE;
        s.bits = 0;
        int pp = -1, ps = 0;
        hz_walk(J, base, total, [&](int pos, int sy) {
            if (pp >= 0) {
                s.bits += group_bits(pos, pp, ps);
            } else {
                s.first_pos = pos;
            }
            off[s.cnt * 32] = (uint8_t) (pos - base);
            sym[s.cnt * 32] = sy;
            pp = pos;
            ps = sy;
            s.cnt++;
        });
        if (s.cnt) {
            s.last_key = mk_key(pp, ps);
        }
        reinterpret_cast<unsigned *>(sc + HZ_DENSE_BITS)[lane] = s.bits;
        sc[HZ_DENSE_CNT + lane] = (uint8_t) s.cnt;
        saved = 1;
        acc.add(s, lane);
    }
};

/* Tile flags (sbt.cuh): a chunk that lies inside one level-1 or level-2 band whose tiles over the chunk's rows are all
 * flagged empty holds no symbol -- decided from a few flag bytes, without touching the 8 KB of coefficients.  (P
 * pictures at qp85: every level-1 band, 3/4 of the plane.)  Warp-uniform. */
DSV_D bool hz_chunk_flagged_empty(const HzJob &J, int cbase, int total, int lane)
{
    if (!J.tflags) {
        return false;
    }
    const HzRegions &rg = J.rg;
    const int last = imin(cbase + HZ_CHUNK, total) - 1;
    int r = 0;
    while (r < HZ_NREG - 1 && cbase >= rg.base[r + 1]) {
        r++;
    }
    const int lvl = rg.lvl[r];
    if (r == 0 || lvl > 2 || last >= rg.base[r + 1]) {
        return false;
    }
    if (lvl == 2 && (J.dg.dvx[1] >= 0 || J.dg.dvy[1] >= 0)) {
        return false; /* first visits of double-visited positions come from the side buffer, not from the plane */
    }
    const int y0 = (int) fastdiv((unsigned) (cbase - rg.base[r]), rg.fdw[r]), y1 = (int) fastdiv((unsigned) (last - rg.base[r]), rg.fdw[r]);
    const int sh = lvl == 1 ? 5 : 4; /* a 128x64 tile holds 32 band rows of level 1, 16 of level 2 */
    const int n = ((y1 >> sh) - (y0 >> sh) + 1) * J.tiles_x;
    const uint8_t *f = J.tflags + (y0 >> sh) * J.tiles_x;
    int any = 0;
    for (int i = lane; i < n; i += 32) {
        any |= f[i] & lvl; /* bit 0 for level 1, bit 1 for level 2 */
    }
    return !__any_sync(0xffffffffu, any != 0);
}

__global__ void __launch_bounds__(HZ_THREADS) hzcc_scan_kernel(const HzJob *jobs, int njobs, HzChunk *chunks, int total_chunks, const HzMap map)
{
    const int lane = threadIdx.x & 31;
    const int chunk = (int) blockIdx.x * HZW_WARPS + (threadIdx.x >> 5);
    if (chunk >= total_chunks) {
        return;
    }
    const HzJob &J = jobs[hz_job_of_chunk(jobs, njobs, chunk, map)];
    HzScanVisitor V;
    const int cbase = (chunk - J.chunk_base) * HZ_CHUNK;
    if (J.dense && J.list_mode != HZ_LISTS_DENSE) {
        V.sparse = J.dense + (size_t) (chunk - J.chunk_base) * HZ_DENSE_BYTES;
    }
    if (!hz_chunk_flagged_empty(J, cbase, J.rg.base[HZ_NREG], lane)) {
        hz_chunk_rounds(J, cbase, J.rg.base[HZ_NREG], lane, V);
    }
    if (lane == 0) {
        HzChunk c;
        c.cnt = (int) V.acc.cnt;
        c.dense = V.saved;
        c.bits_inner = V.acc.bits;
        c.first_pos = V.acc.first;
        c.last_pos = key_pos(V.acc.carry);
        c.last_sym = key_sym(V.acc.carry);
        c.prev_pos = -1;
        c.prev_sym = 0;
        c.bit_off = 0;
        chunks[chunk] = c;
    }
}

/* plain (single-writer) MSB-first bit store into a zeroed byte buffer */
DSV_D void put_bits_plain(uint8_t *buf, unsigned long long bitpos, int len, unsigned long long code)
{
    for (int i = len - 1; i >= 0; i--, bitpos++) {
        if ((code >> i) & 1) {
            buf[bitpos >> 3] |= (uint8_t) (0x80u >> (bitpos & 7));
        }
    }
}
DSV_D void put_u32_plain(uint8_t *buf, unsigned at, unsigned v)
{
    buf[at] = (uint8_t) (v >> 24);
    buf[at + 1] = (uint8_t) (v >> 16);
    buf[at + 2] = (uint8_t) (v >> 8);
    buf[at + 3] = (uint8_t) v;
}

#define HZP_THREADS 1024
__global__ void __launch_bounds__(HZP_THREADS) hzcc_prefix_kernel(const HzJob *jobs, HzChunk *chunks, HzFrame *frames)
{
    __shared__ unsigned long long scratch[40];
    __shared__ unsigned long long s_carry_key, s_carry_bits;
    __shared__ unsigned s_P;
    __shared__ int s_overflow;
    const int tid = threadIdx.x;
    HzFrame &F = frames[blockIdx.x];
    if (tid == 0) {
        s_P = F.start_byte;
        s_overflow = 0;
    }
    __syncthreads();

    for (int p = 0; p < F.nplanes && !s_overflow; p++) {
        const HzJob &J = jobs[F.job[p]];
        HzChunk *ck = chunks + J.chunk_base;
        const int n = J.nchunks;
        const unsigned P = s_P;
        const int dc = J.coef[0];
        const unsigned nruns_at = P + 4 + (unsigned) ((seg_len(dc) + 7) >> 3);
        const unsigned long long token_base = (unsigned long long) (nruns_at + 4) * 8ull;

        if (tid == 0) {
            s_carry_key = KEY_NONE;
            s_carry_bits = 0;
        }
        __syncthreads();
        /* pass A: previous non-zero of every chunk, then the chunk's total bits */
        for (int b0 = 0; b0 < n; b0 += HZP_THREADS) {
            const int c = b0 + tid;
            unsigned long long key = KEY_NONE, tot;
            if (c < n && ck[c].cnt > 0) {
                key = mk_key(ck[c].last_pos, ck[c].last_sym);
            }
            unsigned long long ex = block_scan_excl<OpMaxS64>(key, scratch, &tot);
            ex = OpMaxS64::apply(ex, s_carry_key);
            if (c < n) {
                ck[c].prev_pos = key_pos(ex);
                ck[c].prev_sym = key_sym(ex);
            }
            __syncthreads();
            if (tid == 0) {
                s_carry_key = OpMaxS64::apply(s_carry_key, tot);
            }
            __syncthreads();
        }
        /* pass B: bit offsets and symbol counts (count in the high bits of one 64-bit sum is unsafe for
         * large planes, so run the two sums separately) */
        unsigned long long total_bits = 0, total_cnt = 0;
        for (int b0 = 0; b0 < n; b0 += HZP_THREADS) {
            const int c = b0 + tid;
            unsigned long long bits = 0, cnt = 0, tot;
            if (c < n) {
                cnt = (unsigned long long) ck[c].cnt;
                bits = ck[c].bits_inner;
                if (ck[c].cnt > 0) {
                    bits += group_bits(ck[c].first_pos, ck[c].prev_pos, ck[c].prev_sym);
                }
            }
            unsigned long long ex = block_scan_excl<OpAdd64>(bits, scratch, &tot);
            if (c < n) {
                ck[c].bit_off = token_base + s_carry_bits + ex;
            }
            total_bits = s_carry_bits + tot;
            __syncthreads();
            if (tid == 0) {
                s_carry_bits = total_bits;
            }
            unsigned long long ctot;
            block_scan_incl<OpAdd64>(cnt, scratch, &ctot);
            total_cnt += ctot;
            __syncthreads();
        }
        /* framing (hzcc.c:151-154,283-292,457-474).  The bit total is known before a single token is packed: a
         * picture that does not fit its packet buffer (the reference's w*h*{2,4,6} heuristic overflows its heap
         * there) is refused here, and hzcc_pack_kernel skips it */
        if (tid == 0) {
            unsigned long long end = token_base + total_bits;
            if (((end + 64 + 7) >> 3) + 16 > (unsigned long long) F.cap) {
                s_overflow = 1;
            }
        }
        __syncthreads();
        if (tid == 0 && !s_overflow) {
            uint8_t *pkt = F.pkt;
            unsigned long long end = token_base + total_bits;
            put_bits_plain(pkt, (unsigned long long) (P + 4) * 8ull, seg_len(dc),
                           dc ? ((ueg_code((unsigned) iabs(dc)) << 1) | (dc < 0 ? 1ull : 0ull)) : ueg_code(0));
            put_u32_plain(pkt, nruns_at, (unsigned) total_cnt);
            if (total_cnt > 0) {
                int lastv = key_sym(s_carry_key);
                put_bits_plain(pkt, end, neg_len(lastv), neg_code(lastv));
                end += (unsigned long long) neg_len(lastv);
            }
            unsigned end_byte = (unsigned) ((end + 7) >> 3);
            pkt[end_byte] = 0x55;
            end_byte += 1;
            put_u32_plain(pkt, P, end_byte - P - 4);
            F.plane_bytes[p] = end_byte - P;
            F.plane_nruns[p] = (unsigned) total_cnt;
            s_P = end_byte;
        }
        __syncthreads();
    }
    if (tid == 0) {
        F.overflow = (unsigned) s_overflow;
        F.total_bytes = s_overflow ? F.start_byte : s_P;
    }
}

/* OR `len` (<= 64) bits of `code` (right-aligned) at bit position bitpos, MSB first, 32-bit atomics */
DSV_D void or_bits_atomic(unsigned *words, unsigned long long bitpos, int len, unsigned long long code)
{
    while (len > 0) {
        const unsigned long long wi = bitpos >> 5;
        const int off = (int) (bitpos & 31), room = 32 - off;
        const int n = len < room ? len : room;
        unsigned piece = (unsigned) ((code >> (len - n)) & (n == 32 ? 0xFFFFFFFFull : ((1ull << n) - 1)));
        unsigned be = piece << (room - n);
        atomicOr(&words[wi], __byte_perm(be, 0, 0x0123)); /* big-endian bit order in little-endian words */
        bitpos += (unsigned long long) n;
        len -= n;
    }
}

/* a lane's contiguous piece of the bit stream: whole 32-bit words are stored plainly (the packet starts zeroed and
 * nobody else owns a bit of them), only the first and last word, shared with the neighbours, are OR-ed atomically */
struct HzBitWriter {
    unsigned *words;
    unsigned long long wi; /* next word */
    unsigned long long acc;
    int n;                 /* pending bits in acc (low n bits), < 32 after put() */
    bool first;
    DSV_D void begin(unsigned *w, unsigned long long bitpos)
    {
        words = w;
        wi = bitpos >> 5;
        n = (int) (bitpos & 31); /* the bits before our start belong to someone else: zeros here, OR-ed in */
        acc = 0;
        first = true;
    }
    DSV_D void flush_word(unsigned be)
    {
        const unsigned v = __byte_perm(be, 0, 0x0123); /* big-endian bit order in little-endian words */
        if (first) {
            atomicOr(&words[wi], v);
            first = false;
        } else {
            words[wi] = v;
        }
        wi++;
    }
    DSV_D void put32(int len, unsigned code) /* len <= 32 */
    {
        acc = (acc << len) | (unsigned long long) code;
        n += len;
        if (n >= 32) {
            flush_word((unsigned) (acc >> (n - 32)));
            n -= 32;
        }
    }
    DSV_D void put(int len, unsigned long long code) /* len <= 64 */
    {
        if (len > 32) {
            put32(len - 32, (unsigned) (code >> 32));
            put32(32, (unsigned) code);
        } else {
            put32(len, (unsigned) code);
        }
    }
    DSV_D void end()
    {
        if (n > 0) {
            atomicOr(&words[wi], __byte_perm((unsigned) (acc << (32 - n)), 0, 0x0123));
        }
    }
};

/* bit offsets, accumulated over rounds (or one walk) */
struct HzPackAcc {
    unsigned long long carry; /* last non-zero before this round: from the previous chunks at first */
    unsigned long long off;   /* bit position of this round's first group */
    DSV_D void place(const HzSummary &s, int lane, int &prev_pos, int &prev_sym, unsigned long long &at)
    {
        unsigned long long all;
        unsigned long long ex = warp_prev_key(s.last_key, lane, &all);
        if (key_pos(ex) < 0) {
            ex = carry;
        }
        prev_pos = key_pos(ex);
        prev_sym = key_sym(ex);
        unsigned mybits = s.bits;
        if (s.cnt) {
            mybits += group_bits(s.first_pos, prev_pos, prev_sym);
        }
        unsigned incl = mybits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) {
                incl += n;
            }
        }
        at = off + (unsigned long long) (incl - mybits);
        off += (unsigned long long) __shfl_sync(0xffffffffu, incl, 31);
        carry = OpMaxS64::apply(carry, all);
    }
};
/* Pack pass of a dense chunk whose lists the scan pass saved (HzJob.dense): no coefficient is read.  A free function
 * with its own accumulator, so that the sparse path's visitor never has its address taken. */
DSV_D void hz_pack_from_lists(const HzJob &J, const HzChunk &C, unsigned *words, int cbase, int lane)
{
    HzPackAcc acc;
    acc.carry = mk_key(C.prev_pos, C.prev_sym);
    acc.off = C.bit_off;
    const uint8_t *sc = J.dense + (size_t) (cbase / HZ_CHUNK) * HZ_DENSE_BYTES;
    const uint8_t *off = sc + HZ_DENSE_OFF + lane;
    const int *sym = reinterpret_cast<const int *>(sc + HZ_DENSE_SYM) + lane;
    const int base = cbase + lane * HZW_ITEMS;
    HzSummary s;
    s.cnt = sc[HZ_DENSE_CNT + lane];
    s.bits = reinterpret_cast<const unsigned *>(sc + HZ_DENSE_BITS)[lane];
    s.first_pos = s.cnt ? base + off[0] : -1;
    s.last_key = s.cnt ? mk_key(base + off[(s.cnt - 1) * 32], sym[(s.cnt - 1) * 32]) : KEY_NONE;
    int prev_pos, prev_sym;
    unsigned long long at;
    acc.place(s, lane, prev_pos, prev_sym, at);
    if (s.cnt == 0) {
        return;
    }
    HzBitWriter bw;
    bw.begin(words, at);
    for (int k = 0; k < s.cnt; k++) {
        const int pos = base + off[k * 32];
        const unsigned run = (unsigned) (pos - prev_pos - 1);
        bw.put(ueg_len(run), ueg_code(run));
        if (prev_pos >= 0) {
            bw.put(neg_len(prev_sym), neg_code(prev_sym));
        }
        prev_pos = pos;
        prev_sym = sym[k * 32];
    }
    bw.end();
}

/* Pack pass of a sparse chunk whose list the scan pass saved: 32 non-zeros per round, one per lane */
DSV_D void hz_pack_from_sparse_list(const HzJob &J, const HzChunk &C, unsigned *words, int cbase, int lane)
{
    const uint8_t *sc = J.dense + (size_t) (cbase / HZ_CHUNK) * HZ_DENSE_BYTES;
    const uint16_t *lp = reinterpret_cast<const uint16_t *>(sc + HZ_DENSE_OFF);
    const int *ls = reinterpret_cast<const int *>(sc + HZ_DENSE_SYM);
    int carry_pos = C.prev_pos, carry_sym = C.prev_sym;
    unsigned long long off = C.bit_off;
    for (int e0 = 0; e0 < C.cnt; e0 += 32) {
        const int e = e0 + lane;
        const bool on = e < C.cnt;
        int pos = -1, sym = 0;
        if (on) {
            pos = cbase + lp[e];
            sym = ls[e];
        }
        int prev_pos = __shfl_up_sync(0xffffffffu, pos, 1), prev_sym = __shfl_up_sync(0xffffffffu, sym, 1);
        if (lane == 0) {
            prev_pos = carry_pos;
            prev_sym = carry_sym;
        }
        const unsigned bits = on ? group_bits(pos, prev_pos, prev_sym) : 0u;
        unsigned incl = bits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) {
                incl += n;
            }
        }
        if (on) { /* the non-zero ORs its group (UEG(run) ++ NEG(prev)) at its bit offset */
            unsigned long long at = off + (unsigned long long) (incl - bits);
            const unsigned run = (unsigned) (pos - prev_pos - 1);
            const int l = ueg_len(run);
            or_bits_atomic(words, at, l, ueg_code(run));
            if (prev_pos >= 0) {
                or_bits_atomic(words, at + (unsigned long long) l, neg_len(prev_sym), neg_code(prev_sym));
            }
        }
        off += (unsigned long long) __shfl_sync(0xffffffffu, incl, 31);
        const int last = imin(31, C.cnt - e0 - 1);
        carry_pos = __shfl_sync(0xffffffffu, pos, last);
        carry_sym = __shfl_sync(0xffffffffu, sym, last);
    }
}

struct HzPackVisitor {
    HzPackAcc acc;
    unsigned *words;
    DSV_D void round(const HzGroup &g, int lane)
    {
        int prev_pos, prev_sym;
        unsigned long long at;
        acc.place(hz_group_summary(g), lane, prev_pos, prev_sym, at);
#pragma unroll
        for (int e = 0; e < 4; e++) { /* every non-zero ORs its group (UEG(run) ++ NEG(prev)) at its bit offset */
            if (e < g.cnt) {
                const unsigned run = (unsigned) (g.pos[e] - prev_pos - 1);
                int l = ueg_len(run);
                or_bits_atomic(words, at, l, ueg_code(run));
                at += (unsigned long long) l;
                if (prev_pos >= 0) {
                    l = neg_len(prev_sym);
                    or_bits_atomic(words, at, l, neg_code(prev_sym));
                    at += (unsigned long long) l;
                }
                prev_pos = g.pos[e];
                prev_sym = g.sym[e];
            }
        }
    }
    DSV_D void dense(const HzJob &J, int base, int total, int lane)
    {
        /* one walk: the lane's non-zeros (offset inside its 64 positions, symbol) are kept in a small per-lane list
         * (local memory, L1-resident) while the summary is formed; once the lane's bit offset is known the codes are
         * emitted from the list -- no second pass over the coefficients, no second symbol derivation */
        unsigned char off[HZW_ITEMS];
        int sym[HZW_ITEMS];
        HzSummary s;
        s.cnt = 0;
        s.first_pos = -1;
        s.last_key = KEY_NONE;
        s.bits = 0;
        {
            int pp = -1, ps = 0;
            hz_walk(J, base, total, [&](int pos, int sy) {
                if (pp >= 0) {
                    s.bits += group_bits(pos, pp, ps);
                } else {
                    s.first_pos = pos;
                }
                off[s.cnt] = (unsigned char) (pos - base);
                sym[s.cnt] = sy;
                pp = pos;
                ps = sy;
                s.cnt++;
            });
            if (s.cnt) {
                s.last_key = mk_key(pp, ps);
            }
        }
        int prev_pos, prev_sym;
        unsigned long long at;
        acc.place(s, lane, prev_pos, prev_sym, at);
        if (s.cnt == 0) {
            return;
        }
        HzBitWriter bw;
        bw.begin(words, at);
        for (int k = 0; k < s.cnt; k++) {
            const int pos = base + off[k];
            const unsigned run = (unsigned) (pos - prev_pos - 1);
            bw.put(ueg_len(run), ueg_code(run));
            if (prev_pos >= 0) {
                bw.put(neg_len(prev_sym), neg_code(prev_sym));
            }
            prev_pos = pos;
            prev_sym = sym[k];
        }
        bw.end();
    }
};

/* Dense chunks whose lists the scan pass saved (I pictures) are packed by hzcc_pack_dense_kernel: with that path
 * inside, this kernel needs 64 registers instead of 40 and the sparse (P picture) path loses its occupancy
 * (103 -> 137 us per 64 pictures). */
__global__ void __launch_bounds__(HZ_THREADS) hzcc_pack_kernel(const HzJob *jobs, int njobs, const HzChunk *chunks,
                                                               const HzFrame *frames, int total_chunks, const HzMap map)
{
    const int lane = threadIdx.x & 31;
    const int chunk = (int) blockIdx.x * HZW_WARPS + (threadIdx.x >> 5);
    if (chunk >= total_chunks) {
        return;
    }
    const HzChunk C = chunks[chunk];
    if (C.cnt == 0) {
        return; /* nothing to write; uniform for the warp */
    }
    const HzJob &J = jobs[hz_job_of_chunk(jobs, njobs, chunk, map)];
    if (frames[J.frame].overflow) {
        return; /* refused by the prefix pass: nothing of this picture is written */
    }
    if (C.dense == 1) {
        return; /* hzcc_pack_dense_kernel's */
    }
    if (C.dense == 2) {
        hz_pack_from_sparse_list(J, C, reinterpret_cast<unsigned *>(frames[J.frame].pkt), (chunk - J.chunk_base) * HZ_CHUNK, lane);
        return;
    }
    HzPackVisitor V;
    V.acc.carry = mk_key(C.prev_pos, C.prev_sym);
    V.acc.off = C.bit_off;
    V.words = reinterpret_cast<unsigned *>(frames[J.frame].pkt);
    hz_chunk_rounds(J, (chunk - J.chunk_base) * HZ_CHUNK, J.rg.base[HZ_NREG], lane, V);
}

__global__ void __launch_bounds__(HZ_THREADS) hzcc_pack_dense_kernel(const HzJob *jobs, int njobs, const HzChunk *chunks,
                                                                     const HzFrame *frames, int total_chunks, const HzMap map)
{
    const int lane = threadIdx.x & 31;
    const int chunk = (int) blockIdx.x * HZW_WARPS + (threadIdx.x >> 5);
    if (chunk >= total_chunks) {
        return;
    }
    const HzChunk C = chunks[chunk];
    if (C.cnt == 0 || C.dense != 1) {
        return;
    }
    const HzJob &J = jobs[hz_job_of_chunk(jobs, njobs, chunk, map)];
    if (frames[J.frame].overflow) {
        return;
    }
    hz_pack_from_lists(J, C, reinterpret_cast<unsigned *>(frames[J.frame].pkt), (chunk - J.chunk_base) * HZ_CHUNK, lane);
}

void hzcc_enc_launch(const HzJob *d_jobs, int njobs, HzChunk *d_chunks, int total_chunks,
                     HzFrame *d_frames, int nframes, cudaStream_t st, int chunks_per_pic, int chunks_y, int chunks_u, int any_dense)
{
    HzMap map;
    map.per_pic = chunks_per_pic;
    map.c0 = chunks_y;
    map.c1 = chunks_u;
    map.per_pic_fd = make_fastdiv(chunks_per_pic > 0 ? chunks_per_pic : 1);
    DSV_LAUNCH(hzcc_scan_kernel, dim3(ceil_div(total_chunks, HZW_WARPS)), dim3(HZ_THREADS), 0, st, d_jobs, njobs, d_chunks, total_chunks, map);
    KERNEL_CHECK();
    DSV_LAUNCH(hzcc_prefix_kernel, dim3(nframes), dim3(HZP_THREADS), 0, st, d_jobs, d_chunks, d_frames);
    KERNEL_CHECK();
    DSV_LAUNCH(hzcc_pack_kernel, dim3(ceil_div(total_chunks, HZW_WARPS)), dim3(HZ_THREADS), 0, st, d_jobs, njobs, d_chunks, d_frames, total_chunks, map);
    KERNEL_CHECK();
    if (any_dense) { /* some job has list scratch (HzJob.dense): its dense chunks were left to this kernel */
        DSV_LAUNCH(hzcc_pack_dense_kernel, dim3(ceil_div(total_chunks, HZW_WARPS)), dim3(HZ_THREADS), 0, st, d_jobs, njobs, d_chunks, d_frames, total_chunks, map);
        KERNEL_CHECK();
    }
}

} // namespace dsv
