/*
 * hzcc.cuh -- job descriptors and bit-code helpers for the HZCC coefficient coder kernels
 * (hzcc_enc.cu, hzcc_dec.cu).  Reference: hzcc.c:137-435 (scan order, tokens), bs.c:128-219
 * (interleaved exp-Golomb codes).
 *
 * Scan order of a plane: region 0 = "LL" (ceil(w/8) x ceil(h/8)), then for
 * hzcc level l = 0,1,2 (transform level 3,2,1) the LH, HL, HH regions, each raster.  Token stream for
 * non-zero symbols v_0..v_{n-1} at scan positions s_0 < ... < s_{n-1}:
 *     UEG(run_0), [UEG(run_k), NEG(v_{k-1})] for k = 1..n-1, NEG(v_{n-1}),   run_k = s_k - s_{k-1} - 1.
 * "Group k" below = the bits non-zero k is responsible for: UEG(run_k) followed by NEG(v_{k-1}) (k > 0).
 */
#pragma once
#include "sbt.cuh"

namespace dsv {

#define HZ_NREG 10
#define HZ_THREADS 256
#define HZ_ITEMS 8
#define HZ_CHUNK (HZ_THREADS * HZ_ITEMS)

struct HzRegions {
    int base[HZ_NREG + 1]; /* first scan position of each region; base[10] = total */
    int x0[HZ_NREG], y0[HZ_NREG], sw[HZ_NREG], sh[HZ_NREG];
    int lvl[HZ_NREG];      /* transform level 3,2,1; 4 for the LL region */
    FastDiv fdw[HZ_NREG];  /* division by sw */
};

struct HzJob {
    int32_t *coef; /* dequantised coefficients (encoder: read; decoder: written) */
    int32_t *dv;   /* first-visit symbols (encoder) / values (decoder) of double-visited positions */
    const uint8_t *stable;
    const uint8_t *tflags; /* optional tile flags of the plane (sbt.cuh): clear bits let the scan skip chunks unread */
    int tiles_x;
    /* optional scratch, HZ_DENSE_BYTES per chunk of the plane: the scan pass leaves the non-zeros of a DENSE chunk
     * (I pictures) there -- per lane: bit count, count, then (offset, symbol) lists, lane-interleaved -- so that the
     * pack pass emits straight from the lists instead of walking the coefficients and re-deriving every symbol */
    uint8_t *dense;
    /* what the scan pass leaves in that scratch: HZ_LISTS_DENSE (0, the default) = the lists of dense chunks;
     * HZ_LISTS_BOTH = also the (offset, symbol) list of every sparse chunk, in scan order, so that the pack pass reads
     * no coefficient at all; HZ_LISTS_SPARSE = sparse chunks only (P pictures: their handful of dense chunks is not
     * worth the launch of the dense pack kernel) */
    int list_mode;
    int cw, ch;
    int plane, isP;
    int chunk_base, nchunks; /* this plane's chunks inside the launch-wide chunk arrays */
    int frame;               /* index into the per-frame arrays */
    PlaneQ pq;
    DvGeom dg;
    HzRegions rg;
};

/* per-chunk summary produced by the scan pass */
#define HZ_DENSE_BITS 0     /* unsigned bits[32] */
#define HZ_DENSE_CNT 128    /* uint8 cnt[32] */
#define HZ_DENSE_OFF 256    /* uint8 off[64][32] */
#define HZ_DENSE_SYM 2304   /* int sym[64][32] */
#define HZ_DENSE_BYTES (2304 + 64 * 32 * 4)
/* a sparse chunk's list lives in the same scratch: uint16 offset[<= 508] at HZ_DENSE_OFF, int sym[<= 508] at HZ_DENSE_SYM */
enum { HZ_LISTS_DENSE = 0, HZ_LISTS_SPARSE = 1, HZ_LISTS_BOTH = 2 };

struct HzChunk {
    int cnt;            /* non-zero symbols in the chunk */
    int dense;          /* the chunk's lists are in the job's scratch: 1 = per-lane lists of a dense chunk, 2 = one sparse list */
    int first_pos;      /* scan position of the first / last non-zero, -1 if none */
    int last_pos;
    int last_sym;
    unsigned bits_inner; /* bits of all groups except the chunk's first */
    /* filled by the prefix pass */
    int prev_pos;       /* last non-zero before the chunk (-1 if none) and its symbol */
    int prev_sym;
    unsigned long long bit_off; /* absolute bit position (in the packet) of the chunk's first group */
};

/* per-frame packet state */
struct HzFrame {
    uint8_t *pkt;        /* zeroed packet buffer (device) */
    unsigned start_byte; /* where plane 0 begins (after header/side info written by the host) */
    unsigned cap;        /* bytes allocated at pkt: nothing is written at or beyond pkt + cap */
    unsigned overflow;   /* out: the coded picture does not fit (total_bytes is then start_byte and no plane byte is valid) */
    unsigned total_bytes; /* out: packet length after the three planes */
    unsigned plane_bytes[3];
    unsigned plane_nruns[3];
    int job[3];          /* indices of the frame's plane jobs */
    int nplanes;         /* 3 for a picture; 1 for the single-plane test entry point */
};

void hz_fill_regions(HzRegions *r, int cw, int ch);
void hz_fill_job(HzJob *j, int cw, int ch, int q, int isP, int plane, int nbh, int nbv);
void hzcc_quant_launch(const HzJob *d_jobs, int njobs, int max_elems, cudaStream_t st);
/* chunk -> job without a search when the jobs are the planes Y,U,V of pictures of one format (per_pic == 0: search) */
struct HzMap {
    int per_pic, c0, c1;
    FastDiv per_pic_fd;
};
/* chunks_per_pic > 0 declares that regular layout: job 3k+p covers chunks [k * chunks_per_pic + offset_p, ...) with
 * chunks_y / chunks_u chunks in the first two planes */
void hzcc_enc_launch(const HzJob *d_jobs, int njobs, HzChunk *d_chunks, int total_chunks,
                     HzFrame *d_frames, int nframes, cudaStream_t st, int chunks_per_pic = 0, int chunks_y = 0, int chunks_u = 0,
                     int any_dense = 0);

/* ---- interleaved exp-Golomb code construction (bs.c:128-145) -------------------------------- */
DSV_HD unsigned long long spread_bits(unsigned v)
{
    unsigned long long x = v;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}
DSV_HD int ilog2_u32(unsigned x) /* floor(log2 x), x >= 1 */
{
#ifdef __CUDA_ARCH__
    return 31 - __clz((int) x);
#endif
    int n = 0;
    while (x >> (n + 1)) {
        n++;
    }
    return n;
}
DSV_HD int ueg_len(unsigned v) { return 2 * ilog2_u32(v + 1) + 1; }
DSV_HD int neg_len(int v) { return 2 * ilog2_u32((unsigned) iabs(v)) + 2; } /* UEG(|v|-1) + sign */
DSV_HD int seg_len(int v) { return ueg_len((unsigned) iabs(v)) + (v != 0); }
/* right-aligned code words */
DSV_HD unsigned long long ueg_code(unsigned v)
{
    unsigned x = v + 1;
    int n = ilog2_u32(x);
    return (spread_bits(x & ((1u << n) - 1)) << 1) | 1ull;
}
DSV_HD unsigned long long neg_code(int v) { return (ueg_code((unsigned) iabs(v) - 1) << 1) | (v < 0 ? 1ull : 0ull); }

} // namespace dsv
