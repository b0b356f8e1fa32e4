/*
 * hzcc_dec.cu -- HZCC coefficient coder, decoder side (hzcc_dec, hzcc.c:295-435; dsv_decode_plane,
 * hzcc.c:478-496; exp-Golomb readers, bs.c:147-219) as a parallel bit-FSM parse.
 *
 * The token stream R0, R1,V0, R2,V1, ... R(n-1),V(n-2), V(n-1) has no synchronisation markers, but it
 * is recognised by a 7-state automaton over single bits (interleaved exp-Golomb: flag,data pairs, a
 * terminating 1-flag, and a sign bit after every V).  Function composition of per-word transition
 * tables is associative, so token boundaries fall out of a prefix scan:
 *
 *   hzdec_fsm_kernel     per 32-bit word: exit state + number of token ends for each entry state;
 *                        composed over the CTA (8192 bits) -> one table per CTA.
 *   hzdec_link_kernel    one CTA per plane: scan of the CTA tables from the known start state
 *                        -> entry state and token index of every CTA.
 *   hzdec_token_kernel   per word again, now with known entry state: every token that STARTS in the
 *                        word is decoded by that thread (reading ahead if it is longer) into runs[]
 *                        or vals[]; validity against plen (hzcc.c:337-339) via atomicMin.
 *   hzdec_runsum_kernel / hzdec_scatter_kernel
 *                        inclusive scan of (run+1) gives each non-zero's scan position; it is mapped
 *                        to (region, x, y), dequantised with that position's quantiser and written
 *                        into the zeroed coefficient plane.  Positions scanned twice (SURVEY.md
 *                        Appendix B-1) keep the decoder's semantics -- later scan position wins, an
 *                        earlier value survives only where the later one is absent -- through
 *                        atomicCAS on the first visit.
 * R0 (the first token), SEG(DC) and nruns are read on the host while it parses the packet head.
 */
#include "hzcc_dec.cuh"
#include "scan.cuh"

namespace dsv {

enum { S_R0 = 0, S_RF = 1, S_RD = 2, S_V0 = 3, S_VF = 4, S_VD = 5, S_VS = 6, S_NUM = 7 };

struct FsmSum {
    unsigned exits;      /* 3 bits per entry state */
    unsigned cnt[S_NUM]; /* token ends seen when entering in that state */
};

DSV_D int fsm_step(int s, int bit, unsigned &ends)
{
    switch (s) {
        case S_R0:
        case S_RF:
            if (bit) {
                ends++;
                return S_V0;
            }
            return S_RD;
        case S_RD:
            return S_RF;
        case S_V0:
        case S_VF:
            return bit ? S_VS : S_VD;
        case S_VD:
            return S_VF;
        default: /* S_VS: the sign bit */
            ends++;
            return S_R0;
    }
}

DSV_D FsmSum fsm_identity()
{
    FsmSum r;
    r.exits = 0;
#pragma unroll
    for (int s = 0; s < S_NUM; s++) {
        r.exits |= (unsigned) s << (3 * s);
        r.cnt[s] = 0;
    }
    return r;
}

/* a happens first, then b */
DSV_D FsmSum fsm_compose(const FsmSum &a, const FsmSum &b)
{
    FsmSum r;
    r.exits = 0;
#pragma unroll
    for (int s = 0; s < S_NUM; s++) {
        int e = (a.exits >> (3 * s)) & 7;
        r.exits |= ((b.exits >> (3 * e)) & 7) << (3 * s);
        r.cnt[s] = a.cnt[s] + b.cnt[e];
    }
    return r;
}

DSV_D FsmSum fsm_shfl_up(const FsmSum &v, int d)
{
    FsmSum r;
    r.exits = __shfl_up_sync(0xffffffffu, v.exits, d);
#pragma unroll
    for (int s = 0; s < S_NUM; s++) {
        r.cnt[s] = __shfl_up_sync(0xffffffffu, v.cnt[s], d);
    }
    return r;
}

struct HzDecJob {
    HzJob hz;            /* geometry, quantisers, coef plane */
    const uint8_t *body; /* plane bytes after the plen field (device) */
    unsigned plen;       /* declared length */
    unsigned avail;      /* bytes that may be read from body (to the end of the packet) */
    unsigned tok_bit0;   /* bit position (from body) of the first token after R0 */
    int nruns, first_run, dc;
    int ntok;            /* tokens after R0 = 2*nruns - 1 (0 if nruns == 0) */
    int32_t *runs, *vals; /* cap entries each */
    int cap;
    unsigned *first_bad; /* smallest k whose V token ends at or after plen */
    FsmSum *cta_sum;     /* fsm_ncta entries */
    unsigned *cta_entry; /* per CTA: entry state | token ends before << 3 */
    int fsm_cta_base, fsm_ncta;
    int scan_blk_base, scan_nblk;
    unsigned long long *blk_sum; /* scan_nblk entries */
};

#define HZD_THREADS 256
#define HZD_WORD_BITS 32
#define HZD_CTA_BITS (HZD_THREADS * HZD_WORD_BITS)

DSV_D unsigned body_bit(const HzDecJob &J, unsigned long long pos)
{
    unsigned long long byte = pos >> 3;
    if (byte >= J.avail) {
        return 0;
    }
    return (J.body[byte] >> (7 - (pos & 7))) & 1u;
}
/* 32 bits starting at bit position pos (MSB first), zero past the readable area */
DSV_D unsigned body_word(const HzDecJob &J, unsigned long long pos)
{
    unsigned long long byte = pos >> 3;
    unsigned long long acc = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        unsigned long long b = (byte + i < J.avail) ? J.body[byte + i] : 0;
        acc = (acc << 8) | b;
    }
    return (unsigned) (acc >> (8 - (pos & 7)));
}

template <class JT> DSV_D int dec_job_of(const JT *jobs, int njobs, int blk, int JT::*base)
{
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].*base <= blk) {
            lo = mid;
        } else {
            hi = mid - 1;
        }
    }
    return lo;
}

/*
 * Per-thread table for its 32-bit word.  Walking the word bit by bit for each of the five distinguishable entry
 * states is a chain of 160 dependent steps -- and P pictures are so small that a launch has only a few CTAs per
 * picture, so that chain IS the kernel's run time.  Each CTA therefore first builds a byte-wise transition table in
 * shared memory (state, byte) -> (exit state | token ends << 3): 7 x 16 nibble entries by stepping, the 7 x 256 byte
 * entries by composing two nibbles; a word is then four lookups per entry state.
 */
#define FSM_TAB_BYTES (S_NUM * 256)
DSV_D void fsm_build_table(uint8_t *tab /* shared, FSM_TAB_BYTES */, uint8_t *nib /* shared, S_NUM * 16 */)
{
    for (int e = threadIdx.x; e < S_NUM * 16; e += blockDim.x) {
        int s = e >> 4;
        unsigned ends = 0;
#pragma unroll
        for (int i = 3; i >= 0; i--) {
            s = fsm_step(s, (e >> i) & 1, ends);
        }
        nib[e] = (uint8_t) ((unsigned) s | (ends << 3));
    }
    __syncthreads();
    for (int e = threadIdx.x; e < FSM_TAB_BYTES; e += blockDim.x) {
        const unsigned hi = nib[((e >> 8) << 4) | ((e >> 4) & 15)];
        const unsigned lo = nib[((hi & 7u) << 4) | (e & 15)];
        tab[e] = (uint8_t) ((lo & 7u) | (((hi >> 3) + (lo >> 3)) << 3)); /* at most 6 token ends per byte */
    }
    __syncthreads();
}

DSV_D FsmSum word_summary(unsigned w, const uint8_t *tab)
{
    FsmSum r;
    r.exits = 0;
    const int reps[5] = {S_RF, S_RD, S_VF, S_VD, S_VS};
    unsigned ex[5], ct[5];
#pragma unroll
    for (int c = 0; c < 5; c++) {
        unsigned s = (unsigned) reps[c], ends = 0;
#pragma unroll
        for (int b = 3; b >= 0; b--) {
            const unsigned t = tab[(s << 8) | ((w >> (8 * b)) & 0xffu)];
            s = t & 7u;
            ends += t >> 3;
        }
        ex[c] = s;
        ct[c] = ends;
    }
    const int cls[S_NUM] = {0, 0, 1, 2, 2, 3, 4};
#pragma unroll
    for (int s = 0; s < S_NUM; s++) {
        r.exits |= ex[cls[s]] << (3 * s);
        r.cnt[s] = ct[cls[s]];
    }
    return r;
}

/* inclusive scan of per-thread tables over the CTA; returns this thread's EXCLUSIVE prefix, *total = CTA table */
DSV_D FsmSum block_scan_fsm(const FsmSum &mine, FsmSum *warp_tab /* [8] shared */, FsmSum *total)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    FsmSum inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        FsmSum n = fsm_shfl_up(inc, o);
        if (lane >= o) {
            inc = fsm_compose(n, inc);
        }
    }
    FsmSum ex = fsm_shfl_up(inc, 1);
    if (lane == 0) {
        ex = fsm_identity();
    }
    if (lane == 31) {
        warp_tab[wid] = inc;
    }
    __syncthreads();
    FsmSum pre = fsm_identity();
    for (int w = 0; w < wid; w++) {
        pre = fsm_compose(pre, warp_tab[w]);
    }
    ex = fsm_compose(pre, ex);
    if (total) {
        FsmSum t = fsm_identity();
        for (int w = 0; w < (int) (blockDim.x >> 5); w++) {
            t = fsm_compose(t, warp_tab[w]);
        }
        *total = t;
    }
    __syncthreads();
    return ex;
}

__global__ void __launch_bounds__(HZD_THREADS) hzdec_fsm_kernel(const HzDecJob *jobs, int njobs)
{
    __shared__ FsmSum warp_tab[8];
    __shared__ uint8_t s_tab[FSM_TAB_BYTES], s_nib[S_NUM * 16];
    const int jid = dec_job_of(jobs, njobs, (int) blockIdx.x, &HzDecJob::fsm_cta_base);
    const HzDecJob &J = jobs[jid];
    const int cta = (int) blockIdx.x - J.fsm_cta_base;
    const unsigned long long pos = (unsigned long long) J.tok_bit0 + (unsigned long long) cta * HZD_CTA_BITS +
                                   (unsigned long long) threadIdx.x * HZD_WORD_BITS;
    const unsigned w = body_word(J, pos); /* in flight while the table is built */
    fsm_build_table(s_tab, s_nib);
    FsmSum mine = word_summary(w, s_tab);
    FsmSum total;
    block_scan_fsm(mine, warp_tab, &total);
    if (threadIdx.x == 0) {
        J.cta_sum[cta] = total;
    }
}

/* one CTA per plane: entry state / token count of every fsm CTA, from the known state after R0 */
__global__ void __launch_bounds__(HZD_THREADS) hzdec_link_kernel(const HzDecJob *jobs)
{
    __shared__ FsmSum warp_tab[8];
    __shared__ unsigned s_state, s_count;
    const HzDecJob &J = jobs[blockIdx.x];
    if (threadIdx.x == 0) {
        s_state = S_R0; /* after R0 comes R1: an R token starts (hzcc.c:173-181 order) */
        s_count = 0;
        *J.first_bad = 0x7fffffffu;
    }
    __syncthreads();
    for (int b0 = 0; b0 < J.fsm_ncta; b0 += HZD_THREADS) {
        const int c = b0 + (int) threadIdx.x;
        FsmSum mine = c < J.fsm_ncta ? J.cta_sum[c] : fsm_identity();
        FsmSum total;
        FsmSum ex = block_scan_fsm(mine, warp_tab, &total);
        const unsigned st = s_state, ct = s_count;
        if (c < J.fsm_ncta) {
            unsigned e = (ex.exits >> (3 * st)) & 7;
            J.cta_entry[c] = e | ((ct + ex.cnt[st]) << 3);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            s_state = (total.exits >> (3 * st)) & 7;
            s_count = ct + total.cnt[st];
        }
        __syncthreads();
    }
    /* tokens that never complete inside the readable bits are as good as truncated (hzcc.c:337-339) */
    if (threadIdx.x == 0 && s_count < (unsigned) J.ntok) {
        atomicMin(J.first_bad, s_count >> 1);
    }
}

__global__ void __launch_bounds__(HZD_THREADS) hzdec_token_kernel(const HzDecJob *jobs, int njobs)
{
    __shared__ FsmSum warp_tab[8];
    __shared__ uint8_t s_tab[FSM_TAB_BYTES], s_nib[S_NUM * 16];
    const int jid = dec_job_of(jobs, njobs, (int) blockIdx.x, &HzDecJob::fsm_cta_base);
    const HzDecJob &J = jobs[jid];
    const int cta = (int) blockIdx.x - J.fsm_cta_base;
    const unsigned long long pos0 = (unsigned long long) J.tok_bit0 + (unsigned long long) cta * HZD_CTA_BITS +
                                    (unsigned long long) threadIdx.x * HZD_WORD_BITS;
    const unsigned w = body_word(J, pos0);
    fsm_build_table(s_tab, s_nib);
    FsmSum mine = word_summary(w, s_tab);
    FsmSum ex = block_scan_fsm(mine, warp_tab, nullptr);
    const unsigned ce = J.cta_entry[cta];
    const unsigned cst = ce & 7;
    int state = (int) ((ex.exits >> (3 * cst)) & 7);
    long long tok = (long long) (ce >> 3) + ex.cnt[cst]; /* token ends before this word */
    const bool at_start = (state == S_R0 || state == S_V0);
    /* index of the token currently open (if mid-token) is `tok`; our first own token is the next one */
    long long next_tok = at_start ? tok : tok + 1;
    const long long last_tok = (long long) J.ntok - 1;

    const unsigned long long end = pos0 + HZD_WORD_BITS;
    /* a well-formed token (32-bit value) is at most 66 bits long; anything longer is corrupt data, and letting
     * every thread chase a never-ending token to the end of the packet would be quadratic work */
    unsigned long long hard_end = (unsigned long long) J.avail * 8ull + 64ull;
    if (hard_end > end + 160ull) {
        hard_end = end + 160ull;
    }
    const int limit = hard_end > pos0 ? (int) (hard_end - pos0) : 0; /* <= 192 bits from the start of our word */
    /* the 64 bits behind our word are the neighbours' words: a token that runs past the word end is finished out of
     * registers (the last two lanes of a warp fetch theirs), only corrupt data goes further and reads bit by bit */
    const int lane = (int) threadIdx.x & 31;
    unsigned hi = w, mid = __shfl_down_sync(0xffffffffu, w, 1), lo = __shfl_down_sync(0xffffffffu, w, 2);
    if (lane == 31) {
        mid = body_word(J, pos0 + 32);
    }
    if (lane >= 30) {
        lo = body_word(J, pos0 + 64);
    }
    /* next state for (state, bit), 3 bits each: R0/RF: 0 -> RD, 1 -> V0; RD -> RF; V0/VF: 0 -> VD, 1 -> VS; VD -> VF; VS -> R0 */
    const unsigned long long tbl = 2ull | (3ull << 3) | (2ull << 6) | (3ull << 9) | (1ull << 12) | (1ull << 15) | (5ull << 18) |
                                   (6ull << 21) | (5ull << 24) | (6ull << 27) | (4ull << 30) | (4ull << 33);
    int next_i = (int) next_tok;
    const int last_i = (int) last_tok;
    bool own = false; /* currently inside a token this thread owns */
    unsigned v = 1;
    int i = 0; /* bits consumed from pos0 */
    while (i < limit) {
        const bool starting = (state == S_R0 || state == S_V0);
        if (starting) {
            if (i >= HZD_WORD_BITS || next_i > last_i) {
                break; /* tokens starting beyond our word belong to the next thread */
            }
            own = true;
            v = 1;
            if (next_i == last_i) {
                state = S_V0; /* the final token is V(n-1) even though an R is due (hzcc.c:283-285) */
            }
        }
        if (!own && i >= HZD_WORD_BITS) {
            break;
        }
        unsigned bit;
        if (i < 96) {
            bit = hi >> 31;
            hi = (hi << 1) | (mid >> 31);
            mid = (mid << 1) | (lo >> 31);
            lo <<= 1;
        } else {
            bit = body_bit(J, pos0 + (unsigned long long) i);
        }
        const int prev = state;
        state = (int) ((tbl >> (3 * (2 * prev + (int) bit))) & 7ull);
        i++;
        if (own) {
            if (prev == S_RD || prev == S_VD) {
                v = (v << 1) | bit;
            }
            const bool r_end = (prev == S_R0 || prev == S_RF) && bit; /* token j = next_i -> R_{j/2+1} */
            const bool v_end = prev == S_VS;                          /* V_k: token 2k+1, or the final token 2n-2 */
            if (r_end || v_end) {
                const int k = (next_i >> 1) + (r_end ? 1 : 0);
                int val = r_end ? (int) (v - 1u) : (int) v; /* UEG, UEG + 1 (bs.c:214) */
                if (v_end && val && bit) {
                    val = -val;
                }
                if (k < J.cap) {
                    (r_end ? J.runs : J.vals)[k] = val;
                }
                if (v_end && ((pos0 + (unsigned long long) i) >> 3) >= (unsigned long long) J.plen) { /* byte pointer after the read (hzcc.c:337) */
                    atomicMin(J.first_bad, (unsigned) k);
                }
                own = false;
                next_i++;
            }
        }
    }
}

#define HZS_THREADS 256
#define HZS_ITEMS 4
#define HZS_BLOCK (HZS_THREADS * HZS_ITEMS)

DSV_D unsigned long long run_plus1(const HzDecJob &J, int k)
{
    if (k >= J.nruns || k >= J.cap) {
        return 0;
    }
    unsigned r = k == 0 ? (unsigned) J.first_run : (unsigned) J.runs[k];
    return (unsigned long long) r + 1ull;
}

__global__ void __launch_bounds__(HZS_THREADS) hzdec_runsum_kernel(const HzDecJob *jobs, int njobs)
{
    __shared__ unsigned long long scratch[40];
    const int jid = dec_job_of(jobs, njobs, (int) blockIdx.x, &HzDecJob::scan_blk_base);
    const HzDecJob &J = jobs[jid];
    const int blk = (int) blockIdx.x - J.scan_blk_base;
    unsigned long long s = 0, tot;
    for (int i = 0; i < HZS_ITEMS; i++) {
        s += run_plus1(J, blk * HZS_BLOCK + (int) threadIdx.x * HZS_ITEMS + i);
    }
    block_scan_incl<OpAdd64>(s, scratch, &tot);
    if (threadIdx.x == 0) {
        J.blk_sum[blk] = tot;
    }
}

/* one CTA per plane: exclusive scan of the block sums, in place */
__global__ void __launch_bounds__(1024) hzdec_blkscan_kernel(const HzDecJob *jobs)
{
    __shared__ unsigned long long scratch[40];
    __shared__ unsigned long long s_carry;
    const HzDecJob &J = jobs[blockIdx.x];
    if (threadIdx.x == 0) {
        s_carry = 0;
    }
    __syncthreads();
    for (int b0 = 0; b0 < J.scan_nblk; b0 += 1024) {
        const int i = b0 + (int) threadIdx.x;
        unsigned long long v = i < J.scan_nblk ? J.blk_sum[i] : 0, tot;
        unsigned long long ex = block_scan_excl<OpAdd64>(v, scratch, &tot);
        const unsigned long long carry = s_carry;
        if (i < J.scan_nblk) {
            J.blk_sum[i] = carry + ex;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            s_carry = carry + tot;
        }
        __syncthreads();
    }
}

/* dequantise value v for scan position s and store it (hzcc.c:327-432) */
DSV_D void scatter_one(const HzDecJob &D, unsigned long long s, int v)
{
    const HzJob &J = D.hz;
    const HzRegions &rg = J.rg;
    if (s >= (unsigned long long) rg.base[HZ_NREG] || s == 0) {
        return; /* past the plane, or the DC slot which is overwritten afterwards (hzcc.c:495) */
    }
    int r = 0;
    while ((int) s >= rg.base[r + 1]) {
        r++;
    }
    const int k = (int) s - rg.base[r];
    const int y = (int) fastdiv((unsigned) k, rg.fdw[r]);
    const int x = k - y * rg.sw[r];
    const int ax = rg.x0[r] + x, ay = rg.y0[r] + y;
    const int lvl = rg.lvl[r];
    int out;
    bool first_visit = false;
    if (r == 0) {
        out = dz_dequant(v, J.pq.ll_q);
    } else {
        const int f = J.stable[((y * J.pq.dby[lvl]) >> 14) * J.pq.nbh + ((x * J.pq.dbx[lvl]) >> 14)];
        if (lvl == 1) {
            out = p2_dequant(v, f ? J.pq.sh_hq : J.pq.sh_plain);
        } else {
            const int sel = (f & 2) ? 2 : (f ? 1 : 0);
            out = dz_dequant(v, J.pq.lv[3 - lvl].q[sel]);
            const DvGeom &g = J.dg;
            const int L = lvl - 1;
            first_visit = ((ax == g.dvx[L]) && (ay < g.dvey[L])) || ((ay == g.dvy[L]) && (ax < g.dvex[L]));
        }
    }
    if (J.tflags && r > 0 && lvl <= 2) {
        /* tile flag (sbt.cuh): this tile's level-`lvl` blocks are no longer empty.  Byte inside an aligned word. */
        const uint8_t *fb = J.tflags + (y >> (6 - lvl)) * J.tiles_x + (x >> (7 - lvl));
        if (!(*fb & lvl)) {
            const uintptr_t a = reinterpret_cast<uintptr_t>(fb);
            atomicOr(reinterpret_cast<unsigned *>(a & ~(uintptr_t) 3), (unsigned) lvl << (8 * (a & 3)));
        }
    }
    int32_t *dst = J.coef + (size_t) ay * J.cw + ax;
    if (first_visit) {
        /* the next hzcc level scans this element again and overwrites it if it codes a value there */
        atomicCAS(reinterpret_cast<unsigned *>(dst), 0u, (unsigned) out);
    } else {
        *dst = out;
    }
}

__global__ void __launch_bounds__(HZS_THREADS) hzdec_scatter_kernel(const HzDecJob *jobs, int njobs)
{
    __shared__ unsigned long long scratch[40];
    const int jid = dec_job_of(jobs, njobs, (int) blockIdx.x, &HzDecJob::scan_blk_base);
    const HzDecJob &J = jobs[jid];
    const int blk = (int) blockIdx.x - J.scan_blk_base;
    const int k0 = blk * HZS_BLOCK + (int) threadIdx.x * HZS_ITEMS;
    unsigned long long rp[HZS_ITEMS], s = 0, tot;
    for (int i = 0; i < HZS_ITEMS; i++) {
        rp[i] = run_plus1(J, k0 + i);
        s += rp[i];
    }
    unsigned long long pos = J.blk_sum[blk] + block_scan_excl<OpAdd64>(s, scratch, &tot);
    const unsigned bad = *J.first_bad;
    const int n = J.nruns < J.cap ? J.nruns : J.cap;
    for (int i = 0; i < HZS_ITEMS; i++) {
        const int k = k0 + i;
        pos += rp[i];
        if (k < n && (unsigned) k < bad) {
            scatter_one(J, pos - 1ull, J.vals[k]);
        }
    }
    if (blk == 0 && threadIdx.x == 0) {
        J.hz.coef[0] = J.dc;
    }
}

/* DC when a plane has no tokens at all (no scatter blocks are launched for it) */
__global__ void hzdec_dc_kernel(const HzDecJob *jobs, int njobs)
{
    int j = (int) (blockIdx.x * blockDim.x + threadIdx.x);
    if (j < njobs) {
        jobs[j].hz.coef[0] = jobs[j].dc;
    }
}

/*
 * Coefficient planes start zeroed (dsv_decoder.c:405).  Instead of clearing 4 * cw * ch bytes per plane and picture,
 * the clean-up follows the tile flags the previous picture's scatter left behind: a tile clears the level-1 / level-2
 * blocks it flagged and its share of the small level >= 3 corner (always), then puts its flag back to the plane's
 * base value.  P pictures at qp85 touch no level-1 block at all: 3/4 of the plane is neither cleared nor, in the
 * inverse transform, read.  One CTA per tile and plane.
 */
#define CLEAN_TPC 8 /* tiles per CTA: a P picture leaves ~8 KB to clear per tile, too little for a CTA of its own
                        (one tile per CTA: 49 000 CTAs per 64 HD pictures, launch-rate bound at 70 us) */
__global__ void __launch_bounds__(256) hzdec_clean_kernel(const HzCleanItem *items)
{
    __shared__ int s_f[CLEAN_TPC];
    const HzCleanItem &C = items[blockIdx.y];
    const int ntiles = C.tiles_x * C.tiles_y, t0 = (int) blockIdx.x * CLEAN_TPC;
    if (t0 >= ntiles) {
        return;
    }
    if (threadIdx.x < CLEAN_TPC) {
        const int t = t0 + (int) threadIdx.x;
        s_f[threadIdx.x] = t < ntiles ? C.tflags[t] : 0;
    }
    __syncthreads(); /* every thread works from the copy: the flags themselves go back to the base value now */
    if (threadIdx.x < CLEAN_TPC && t0 + (int) threadIdx.x < ntiles) {
        C.tflags[t0 + threadIdx.x] = (uint8_t) C.base;
    }
#pragma unroll 1
    for (int k = 0; k < CLEAN_TPC && t0 + k < ntiles; k++) {
        const int t = t0 + k, f = s_f[k];
        const int ty = t / C.tiles_x, tx = t - ty * C.tiles_x;
        /* rectangle q: 0 = the corner [0, x2) x [0, y2) in 32x16 shares, 1..3 = level-2 regions, 4..6 = level-1 regions */
#pragma unroll 1
        for (int q = 0; q < 7; q++) {
            const int lvl = q < 4 ? 2 : 1;
            if (q > 0 && !(f & lvl)) {
                continue;
            }
            const int ls = 7 - lvl;                  /* log2 of the block width: 32 or 64 coefficients */
            const int bw = 1 << ls, bh = 64 >> lvl;
            const int rx = q == 0 ? 0 : C.rx[q - 1], ry = q == 0 ? 0 : C.ry[q - 1];
            const int rw = q == 0 ? C.x2 : C.rw[q - 1], rh = q == 0 ? C.y2 : C.rh[q - 1];
            const int xa = tx * bw, ya = ty * bh;
            const int w = imin(bw, rw - xa), nrow = imin(bh, rh - ya);
            if (w <= 0 || nrow <= 0) {
                continue;
            }
            int32_t *base = C.coef + (size_t) (ry + ya) * C.cw + rx + xa;
            if (((reinterpret_cast<uintptr_t>(base) | (uintptr_t) (C.cw * 4)) & 15) == 0 && (w & 3) == 0) {
                const int vs = ls - 2; /* 16-byte stores per full row: 8 or 16 */
                for (int i = (int) threadIdx.x; i < (nrow << vs); i += 256) {
                    const int yy = i >> vs, xx = (i & ((1 << vs) - 1)) * 4;
                    if (xx < w) {
                        *reinterpret_cast<int4 *>(base + (size_t) yy * C.cw + xx) = make_int4(0, 0, 0, 0);
                    }
                }
            } else {
                for (int i = (int) threadIdx.x; i < (nrow << ls); i += 256) {
                    const int yy = i >> ls, xx = i & (bw - 1);
                    if (xx < w) {
                        base[(size_t) yy * C.cw + xx] = 0;
                    }
                }
            }
        }
    }
}

int hz_flag_base(const DvGeom &g)
{
    /* levels with double-visited positions (odd sizes one level up) exchange values with the
     * neighbouring level's blocks: their flags stay set, nothing is skipped there */
    return ((g.dvx[1] >= 0 || g.dvy[1] >= 0) ? 1 : 0) | ((g.dvx[2] >= 0 || g.dvy[2] >= 0) ? 2 : 0);
}

void hzdec_fill_clean(HzCleanItem *c, const HzJob &h, int tiles_y)
{
    memset(c, 0, sizeof(*c));
    c->coef = h.coef;
    c->tflags = const_cast<uint8_t *>(h.tflags);
    c->cw = h.cw;
    c->tiles_x = h.tiles_x;
    c->tiles_y = tiles_y;
    c->base = hz_flag_base(h.dg);
    c->x2 = h.rg.x0[4]; /* level-2 LH starts where the corner ends */
    c->y2 = h.rg.y0[5];
    for (int q = 0; q < 6; q++) {
        c->rx[q] = h.rg.x0[4 + q];
        c->ry[q] = h.rg.y0[4 + q];
        /* a region may be one column / row wider than the plane's band when the size one level down is odd: stay inside the plane */
        c->rw[q] = imin(h.rg.sw[4 + q], h.cw - h.rg.x0[4 + q]);
        c->rh[q] = imin(h.rg.sh[4 + q], h.ch - h.rg.y0[4 + q]);
    }
}

void hzdec_clean_launch(const HzCleanItem *d_items, int n, int max_tiles, cudaStream_t st)
{
    if (n > 0 && max_tiles > 0) {
        DSV_LAUNCH(hzdec_clean_kernel, dim3((unsigned) ceil_div(max_tiles, CLEAN_TPC), (unsigned) n), dim3(256), 0, st, d_items);
        KERNEL_CHECK();
    }
}

size_t hzdec_job_size() { return sizeof(HzDecJob); }

} // namespace dsv

#include "hzcc_dec.cuh"
#include "host/bits.h"

namespace dsv {

void hzdec_parse_head(const uint8_t *host_body, unsigned avail, unsigned plen, HzPlaneData *pd)
{
    BitReader br(host_body, avail);
    pd->dc = br.get_seg();
    br.align();
    pd->nruns = (int) br.get_bits(32);
    br.align();
    pd->first_run = pd->nruns > 0 ? (int) br.get_ueg() : 0;
    pd->tok_bit0 = (unsigned) br.pos;
    pd->plen = plen;
    pd->avail = avail;
    pd->body = nullptr;
}

void hzdec_plan(HzDecPlan *pl, int cw, int ch)
{
    HzRegions r;
    hz_fill_regions(&r, cw, ch);
    pl->cap = r.base[HZ_NREG];
    pl->max_bits = (size_t) cw * ch * 8 * 8 + 4096; /* plen <= 2 * framesz (dsv_decoder.c:397-401) */
    pl->max_fsm_cta = (int) ((pl->max_bits + HZD_CTA_BITS - 1) / HZD_CTA_BITS) + 1;
    pl->max_scan_blk = ceil_div(pl->cap, HZS_BLOCK) + 1;
}

void hzdec_plane_alloc(HzDecPlaneBufs *b, const HzDecPlan &pl)
{
    CUDA_CHECK(cudaMalloc(&b->runs, (size_t) (pl.cap + 8) * 4));
    CUDA_CHECK(cudaMalloc(&b->vals, (size_t) (pl.cap + 8) * 4));
    CUDA_CHECK(cudaMalloc(&b->cta_sum, (size_t) pl.max_fsm_cta * sizeof(FsmSum)));
    CUDA_CHECK(cudaMalloc(&b->cta_entry, (size_t) pl.max_fsm_cta * 4));
    CUDA_CHECK(cudaMalloc(&b->blk_sum, (size_t) pl.max_scan_blk * 8));
    CUDA_CHECK(cudaMalloc(&b->first_bad, 4));
    b->plan = pl;
}

void hzdec_plane_free(HzDecPlaneBufs *b)
{
    cudaFree(b->runs);
    cudaFree(b->vals);
    cudaFree(b->cta_sum);
    cudaFree(b->cta_entry);
    cudaFree(b->blk_sum);
    cudaFree(b->first_bad);
    memset(b, 0, sizeof(*b));
}

/* fill one job record (host memory, hzdec_job_size() bytes) and advance the launch-wide block counters */
void hzdec_fill_job(void *slot, const HzJob &hz, const HzPlaneData &pd, const HzDecPlaneBufs &b, HzDecDims *dims)
{
    HzDecJob &J = *reinterpret_cast<HzDecJob *>(slot);
    memset(&J, 0, sizeof(J));
    J.hz = hz;
    J.body = pd.body;
    J.plen = pd.plen;
    J.avail = pd.avail;
    J.tok_bit0 = pd.tok_bit0;
    J.nruns = pd.nruns;
    J.first_run = pd.first_run;
    J.dc = pd.dc;
    J.cap = b.plan.cap;
    int n = J.nruns < 0 ? 0 : (J.nruns > J.cap ? J.cap : J.nruns);
    J.ntok = J.nruns > 0 ? (J.nruns > 0x3fffffff ? 0x7ffffffe : 2 * J.nruns - 1) : 0;
    J.runs = b.runs;
    J.vals = b.vals;
    J.first_bad = b.first_bad;
    J.cta_sum = reinterpret_cast<FsmSum *>(b.cta_sum);
    J.cta_entry = b.cta_entry;
    J.blk_sum = b.blk_sum;
    /* bits that can hold usable tokens: up to plen (a token ending at/after plen invalidates the rest) */
    unsigned long long lim = (unsigned long long) (J.plen < J.avail ? J.plen : J.avail) * 8ull + 64ull;
    unsigned long long nbits = lim > J.tok_bit0 ? lim - J.tok_bit0 : 0;
    J.fsm_ncta = J.ntok > 0 ? (int) ((nbits + HZD_CTA_BITS - 1) / HZD_CTA_BITS) : 0;
    if (J.fsm_ncta > b.plan.max_fsm_cta) {
        J.fsm_ncta = b.plan.max_fsm_cta;
    }
    J.fsm_cta_base = dims->fsm_total;
    dims->fsm_total += J.fsm_ncta;
    J.scan_nblk = ceil_div(n, HZS_BLOCK);
    J.scan_blk_base = dims->scan_total;
    dims->scan_total += J.scan_nblk;
    dims->njobs++;
}

/* d_jobs: device array of dims.njobs job records (planes of every picture in flight), coef planes zeroed */
void hzdec_launch_jobs(const void *d_jobs, const HzDecDims &dims, cudaStream_t st)
{
    const HzDecJob *dj = reinterpret_cast<const HzDecJob *>(d_jobs);
    const int nj = dims.njobs;
    if (nj <= 0) {
        return;
    }
    if (dims.fsm_total > 0) {
        DSV_LAUNCH(hzdec_fsm_kernel, dim3(dims.fsm_total), dim3(HZD_THREADS), 0, st, dj, nj);
        KERNEL_CHECK();
    }
    DSV_LAUNCH(hzdec_link_kernel, dim3(nj), dim3(HZD_THREADS), 0, st, dj);
    KERNEL_CHECK();
    if (dims.fsm_total > 0) {
        DSV_LAUNCH(hzdec_token_kernel, dim3(dims.fsm_total), dim3(HZD_THREADS), 0, st, dj, nj);
        KERNEL_CHECK();
    }
    DSV_LAUNCH(hzdec_dc_kernel, dim3(ceil_div(nj, 32)), dim3(32), 0, st, dj, nj);
    KERNEL_CHECK();
    if (dims.scan_total > 0) {
        DSV_LAUNCH(hzdec_runsum_kernel, dim3(dims.scan_total), dim3(HZS_THREADS), 0, st, dj, nj);
        KERNEL_CHECK();
        DSV_LAUNCH(hzdec_blkscan_kernel, dim3(nj), dim3(1024), 0, st, dj);
        KERNEL_CHECK();
        DSV_LAUNCH(hzdec_scatter_kernel, dim3(dims.scan_total), dim3(HZS_THREADS), 0, st, dj, nj);
        KERNEL_CHECK();
    }
}

} // namespace dsv
