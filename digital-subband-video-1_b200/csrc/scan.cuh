/* scan.cuh -- block-wide scans built on warp shuffles (all threads of the block must call). */
#pragma once
#include "common.cuh"

namespace dsv {

struct OpAdd64 {
    DSV_D static unsigned long long ident() { return 0ull; }
    DSV_D static unsigned long long apply(unsigned long long a, unsigned long long b) { return a + b; }
};
/* signed max on 64-bit keys (used for "right-most valid element": key = pos << 32 | payload, pos = -1 when absent) */
struct OpMaxS64 {
    DSV_D static unsigned long long ident() { return 0x8000000000000000ull; }
    DSV_D static unsigned long long apply(unsigned long long a, unsigned long long b)
    {
        return ((long long) a > (long long) b) ? a : b;
    }
};

/*
 * Inclusive scan of v across the block; *total receives the block aggregate.
 * warp_scratch: shared array of at least 33 unsigned long long.
 */
template <class Op>
DSV_D unsigned long long block_scan_incl(unsigned long long v, unsigned long long *warp_scratch,
                                         unsigned long long *total)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) {
            v = Op::apply(n, v);
        }
    }
    if (lane == 31) {
        warp_scratch[wid] = v;
    }
    __syncthreads();
    if (wid == 0) {
        unsigned long long w = lane < nw ? warp_scratch[lane] : Op::ident();
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long n = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) {
                w = Op::apply(n, w);
            }
        }
        warp_scratch[lane] = w; /* inclusive over warps */
        if (lane == 31) {
            warp_scratch[32] = w;
        }
    }
    __syncthreads();
    if (wid > 0) {
        v = Op::apply(warp_scratch[wid - 1], v);
    }
    *total = warp_scratch[nw - 1];
    __syncthreads(); /* scratch may be reused by the caller */
    return v;
}

/* exclusive variant: returns the aggregate of all elements strictly before this thread */
template <class Op>
DSV_D unsigned long long block_scan_excl(unsigned long long v, unsigned long long *warp_scratch,
                                         unsigned long long *total)
{
    unsigned long long inc = block_scan_incl<Op>(v, warp_scratch, total);
    unsigned long long prev = __shfl_up_sync(0xffffffffu, inc, 1);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (lane == 0) {
        /* need the inclusive value of the last lane of the previous warp: recompute from scratch is gone,
         * so derive it: inclusive(prev thread) = exclusive(this thread).  Use a second shared exchange. */
        prev = Op::ident();
    }
    /* lanes 0 of warps > 0 are patched below through shared memory */
    __shared__ unsigned long long s_last[33];
    if (lane == 31) {
        s_last[wid] = inc;
    }
    __syncthreads();
    if (lane == 0 && wid > 0) {
        prev = s_last[wid - 1];
    }
    __syncthreads();
    return prev;
}

} // namespace dsv
