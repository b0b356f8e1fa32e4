/*
 * hzcc_enc.cu -- HZCC coefficient coder, encoder side: hzcc_enc's token stream (hzcc.c:137-293)
 * and dsv_encode_plane's framing (hzcc.c:449-476) as three data-parallel passes over the scan order.
 *
 *   hzcc_scan_kernel    per 2048-position chunk: gather symbols in scan order (coalesced row
 *                       segments; the quantised symbol is re-derived from the dequantised
 *                       coefficient the SBT epilogue stored), count non-zeros, find the chunk's
 *                       first/last non-zero and the bits of all groups that are fully determined
 *                       inside the chunk.
 *   hzcc_prefix_kernel  one CTA per frame: exclusive scans over the chunk summaries (previous
 *                       non-zero, bit offsets), plane framing (plen, SEG(DC), nruns, final NEG,
 *                       0x55) and plane base offsets; yields the packet length.
 *   hzcc_pack_kernel    per chunk again: every non-zero ORs its group (UEG(run) ++ NEG(prev))
 *                       at its bit offset into the zeroed packet (MSB-first, bs.c:76-91).
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
    if (!v) {
        return 0;
    }
    int f = J.stable[((y * J.pq.dby[lvl]) >> 14) * J.pq.nbh + ((x * J.pq.dbx[lvl]) >> 14)];
    if (lvl == 1) {
        return p2_quant(v, f ? J.pq.sh_hq : J.pq.sh_plain);
    }
    int sel = (f & 2) ? 2 : (f ? 1 : 0);
    const LevelQ &Lq = J.pq.lv[3 - lvl];
    return dz_quant(v, Lq.q[sel], Lq.fd[sel]);
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

/* load the chunk's symbols and find, for every thread, the last non-zero before its first item */
DSV_D void chunk_load(const HzJob &J, int chunk_local, int sym[HZ_ITEMS], int &base,
                      unsigned long long *scratch, unsigned long long &excl_key, unsigned long long &chunk_last)
{
    base = chunk_local * HZ_CHUNK + (int) threadIdx.x * HZ_ITEMS;
    unsigned long long mine = KEY_NONE;
    const int total = J.rg.base[HZ_NREG];
    HzCursor cur;
    bool all_zero = false;
    if (base < total) {
        hz_locate(J.rg, base, cur);
        /* the thread's 8 positions usually sit in one row of one region, contiguous in memory: two 16-byte loads
         * tell whether there is anything to quantise at all (rarely, in a P picture) */
        const HzRegions &rg = J.rg;
        const int r = cur.r, lvl = rg.lvl[r];
        const bool dv = r > 0 && lvl >= 2 && (J.dg.dvx[lvl - 1] >= 0 || J.dg.dvy[lvl - 1] >= 0);
        if (!dv && cur.x + HZ_ITEMS <= rg.sw[r] && !(r == 0 && (cur.x | cur.y) == 0)) {
            const int32_t *p = J.coef + (size_t) (rg.y0[r] + cur.y) * J.cw + rg.x0[r] + cur.x;
            if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
                const int4 a = *reinterpret_cast<const int4 *>(p), b = *reinterpret_cast<const int4 *>(p + 4);
                all_zero = (a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) == 0;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < HZ_ITEMS; i++) {
        sym[i] = 0;
        if (!all_zero && base + i < total) {
            sym[i] = hz_symbol_at(J, cur);
            hz_advance(J.rg, cur);
        }
        if (sym[i]) {
            mine = mk_key(base + i, sym[i]);
        }
    }
    excl_key = block_scan_excl<OpMaxS64>(mine, scratch, &chunk_last);
}

#define HZ_CPB 1 /* chunks per CTA (measured: 8 is slower; most chunks of a P picture still hold a few non-zeros) */

/* (re)load the job record of `chunk` into shared memory unless the cached one already covers it */
DSV_D void hz_cache_job(HzJob *sJ, int *s_have, const HzJob *jobs, int njobs, int chunk, const HzMap &map)
{
    __syncthreads(); /* everybody is done with the previous chunk (and with *sJ) */
    const bool hit = *s_have && chunk >= sJ->chunk_base && chunk < sJ->chunk_base + sJ->nchunks;
    __syncthreads();
    if (!hit) {
        const int jid = hz_job_of_chunk(jobs, njobs, chunk, map);
        const int *src = reinterpret_cast<const int *>(&jobs[jid]);
        int *dst = reinterpret_cast<int *>(sJ);
        for (int i = threadIdx.x; i < (int) (sizeof(HzJob) / sizeof(int)); i += HZ_THREADS) {
            dst[i] = src[i];
        }
        if (threadIdx.x == 0) {
            *s_have = 1;
        }
        __syncthreads();
    }
}

#ifndef HZ_SCAN_MINB
#define HZ_SCAN_MINB 8 /* measured: 8 -> 252 us, default (40 registers) -> 268 */
#endif
__global__ void __launch_bounds__(HZ_THREADS, HZ_SCAN_MINB) hzcc_scan_kernel(const HzJob *jobs, int njobs, HzChunk *chunks, int total_chunks, const HzMap map)
{
    __shared__ HzJob J;
    __shared__ unsigned long long scratch[40];
    __shared__ int s_first, s_have;
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_have = 0;
    }
    for (int ci = 0; ci < HZ_CPB; ci++) {
        const int chunk = (int) blockIdx.x * HZ_CPB + ci;
        if (chunk >= total_chunks) {
            return;
        }
        hz_cache_job(&J, &s_have, jobs, njobs, chunk, map);
        if (tid == 0) {
            s_first = -1;
        }
        __syncthreads();

        int sym[HZ_ITEMS], base;
        unsigned long long excl, last;
        chunk_load(J, chunk - J.chunk_base, sym, base, scratch, excl, last);

        int prev_pos = key_pos(excl), prev_sym = key_sym(excl);
        unsigned bits = 0, cnt = 0;
#pragma unroll
        for (int i = 0; i < HZ_ITEMS; i++) {
            if (sym[i]) {
                if (prev_pos >= 0) {
                    bits += group_bits(base + i, prev_pos, prev_sym);
                } else {
                    s_first = base + i; /* exactly one thread sees the chunk's first non-zero */
                }
                prev_pos = base + i;
                prev_sym = sym[i];
                cnt++;
            }
        }
        unsigned long long tot;
        block_scan_incl<OpAdd64>(((unsigned long long) cnt << 40) | bits, scratch, &tot);
        if (tid == 0) {
            HzChunk c;
            c.cnt = (int) (tot >> 40);
            c.bits_inner = (unsigned) (tot & 0xFFFFFFFFFFull);
            c.first_pos = s_first;
            c.last_pos = key_pos(last);
            c.last_sym = key_sym(last);
            c.prev_pos = -1;
            c.prev_sym = 0;
            c.bit_off = 0;
            chunks[chunk] = c;
        }
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
    const int tid = threadIdx.x;
    HzFrame &F = frames[blockIdx.x];
    if (tid == 0) {
        s_P = F.start_byte;
    }
    __syncthreads();

    for (int p = 0; p < F.nplanes; p++) {
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
        /* framing (hzcc.c:151-154,283-292,457-474) */
        if (tid == 0) {
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
        F.total_bytes = s_P;
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

#ifndef HZ_PACK_MINB
#define HZ_PACK_MINB 8 /* measured: 8 -> 193 us, 6 -> 202, default (48 registers) -> 216 */
#endif
__global__ void __launch_bounds__(HZ_THREADS, HZ_PACK_MINB) hzcc_pack_kernel(const HzJob *jobs, int njobs, const HzChunk *chunks,
                                                               const HzFrame *frames, int total_chunks, const HzMap map)
{
    __shared__ HzJob J;
    __shared__ unsigned long long scratch[40];
    __shared__ int s_have;
    const int tid = threadIdx.x;
    if (tid == 0) {
        s_have = 0;
    }
    __syncthreads();
    for (int ci = 0; ci < HZ_CPB; ci++) {
        const int chunk = (int) blockIdx.x * HZ_CPB + ci;
        if (chunk >= total_chunks) {
            return;
        }
        const HzChunk C = chunks[chunk];
        if (C.cnt == 0) {
            continue; /* nothing to write (most chunks of a P picture); uniform for the whole block */
        }
        hz_cache_job(&J, &s_have, jobs, njobs, chunk, map);
        int sym[HZ_ITEMS], base;
        unsigned long long excl, last;
        chunk_load(J, chunk - J.chunk_base, sym, base, scratch, excl, last);

        int prev_pos = key_pos(excl), prev_sym = key_sym(excl);
        if (prev_pos < 0) { /* nothing earlier in this chunk: continue from the previous chunks */
            prev_pos = C.prev_pos;
            prev_sym = C.prev_sym;
        }
        const int pp0 = prev_pos, ps0 = prev_sym;
        unsigned long long bits = 0;
    #pragma unroll
        for (int i = 0; i < HZ_ITEMS; i++) {
            if (sym[i]) {
                bits += group_bits(base + i, prev_pos, prev_sym);
                prev_pos = base + i;
                prev_sym = sym[i];
            }
        }
        unsigned long long tot;
        unsigned long long off = C.bit_off + block_scan_excl<OpAdd64>(bits, scratch, &tot);
        unsigned *words = reinterpret_cast<unsigned *>(frames[J.frame].pkt);
        prev_pos = pp0;
        prev_sym = ps0;
    #pragma unroll
        for (int i = 0; i < HZ_ITEMS; i++) {
            if (sym[i]) {
                unsigned run = (unsigned) (base + i - prev_pos - 1);
                int l = ueg_len(run);
                or_bits_atomic(words, off, l, ueg_code(run));
                off += (unsigned long long) l;
                if (prev_pos >= 0) {
                    l = neg_len(prev_sym);
                    or_bits_atomic(words, off, l, neg_code(prev_sym));
                    off += (unsigned long long) l;
                }
                prev_pos = base + i;
                prev_sym = sym[i];
            }
        }
    }
}

void hzcc_enc_launch(const HzJob *d_jobs, int njobs, HzChunk *d_chunks, int total_chunks,
                     HzFrame *d_frames, int nframes, cudaStream_t st, int chunks_per_pic, int chunks_y, int chunks_u)
{
    HzMap map;
    map.per_pic = chunks_per_pic;
    map.c0 = chunks_y;
    map.c1 = chunks_u;
    map.per_pic_fd = make_fastdiv(chunks_per_pic > 0 ? chunks_per_pic : 1);
    DSV_LAUNCH(hzcc_scan_kernel, dim3(ceil_div(total_chunks, HZ_CPB)), dim3(HZ_THREADS), 0, st, d_jobs, njobs, d_chunks, total_chunks, map);
    KERNEL_CHECK();
    DSV_LAUNCH(hzcc_prefix_kernel, dim3(nframes), dim3(HZP_THREADS), 0, st, d_jobs, d_chunks, d_frames);
    KERNEL_CHECK();
    DSV_LAUNCH(hzcc_pack_kernel, dim3(ceil_div(total_chunks, HZ_CPB)), dim3(HZ_THREADS), 0, st, d_jobs, njobs, d_chunks, d_frames, total_chunks, map);
    KERNEL_CHECK();
}

} // namespace dsv
