/*
 * sbt.cuh -- job descriptors for the subband-transform kernels (sbt_fwd.cu, sbt_inv.cu).
 *
 * One "job" = one plane of one frame.  A launch takes an array of jobs in device
 * memory so that planes x frames x sequences are batched into one grid (a single
 * 1080p plane is only a few microseconds of HBM traffic).
 *
 * Decomposition (reference: sbt.c:617-714), three kernels per direction:
 *   "tile" (hi)  levels 1-2 per 128x64-sample tile: the streaming part (94 % of the coefficients), no
 *                serial tail -- every thread has work at both levels;
 *   "mid"        levels 3..nlt (nlt = min(5, levels)) per 128x64 block of LL_2;
 *   "lo"         the remaining levels on the small LL_nlt array, one CTA per plane.
 * They hand LL_2 and LL_nlt over through the job's `llx` scratch array (LL_nlt first, LL_2 at ll2_off).
 */
#pragma once
#include "common.cuh"

namespace dsv {

#define SBT_TW 128 /* tile width  in samples */
#define SBT_TH 64  /* tile height in samples */
#define SBT_NLT 5  /* levels done before the lo kernel */
#define SBT_HI 2   /* levels done by the streaming tile kernel */
#define SBT_TILE_THREADS 256
#define SBT_LO_THREADS 1024
#define SBT_STAB_SMEM 2048

/* adaptive quantiser table for one plane of one frame (hzcc.c:50-92,196-258) */
struct LevelQ {
    int q[3]; /* [0] plain, [1] stable, [2] intra block */
    FastDiv fd[3];
};
struct PlaneQ {
    int ll_q; /* "LL" region: levels >= 4 (hzcc.c:158-188) */
    FastDiv ll_fd;
    LevelQ lv[2];          /* hzcc level 0,1 == transform level 3,2 */
    int sh_plain, sh_hq;   /* hzcc level 2 == transform level 1: shift amounts */
    int dbx[4], dby[4];    /* 14-bit fixed-point block steps per transform level 1..3 (hzcc.c:196-197) */
    int nbh, nbv;
};

/* geometry of the positions hzcc visits twice, per transform level l in {2,1}:
 * element (ax,ay) written by level l is ALSO the last column/row of level l+1's scan region when
 * ax == dvx[l] (rows < dvey[l]) or ay == dvy[l] (cols < dvex[l]).  -1 = no overlap. */
struct DvGeom {
    int dvx[3], dvy[3], dvex[3], dvey[3];
    int col_base[3], row_base[3]; /* offsets into the plane's dv side buffer */
    int total;
};

struct SbtJob {
    /* pixels: bordered-frame layout, pix -> sample (0,0): forward input */
    uint8_t *pix;
    int pstride;
    /* inverse output (may be a different frame than the forward input, e.g. the new reference) */
    uint8_t *opix;
    int ostride;
    /* inverse only, optional: a prediction plane added on the way out, opix = clamp(sample + addp - 128)
     * (dsv_add_pred / addf, bmc.c:304-346) -- the encoder's closed-loop reconstruction without a separate pass */
    const uint8_t *addp;
    int addstride;
    int pw, ph;
    /* coefficients: dense, stride == cw */
    int32_t *coef;
    int cw, ch;
    int32_t *llx; /* hand-over scratch: LL_nlt (wo_nlt * ho_nlt ints), then LL_2 at ll2_off */
    int ll2_off;
    int32_t *dv;  /* first-visit symbols of double-visited positions (encoder) / values (decoder) */
    /* optional, one byte per 128x64 tile (tiles_x * tiles_y): bit 0 = the tile's level-1 band blocks (64x32 each of
     * LH, HL, HH) hold a non-zero coefficient, bit 1 = its level-2 blocks (32x16) do.  Written by the forward tile
     * kernel's quantiser (encoder) or by the entropy decoder's scatter pass; a clear bit lets the inverse tile
     * kernel, the HZCC scan and the decoder's clean-up skip the block without reading it.  At qp85 the level-1
     * bands of P pictures are zero by construction (|LH| <= 4 * 255 < 2^10 = the quantiser step). */
    uint8_t *tflags;
    const uint8_t *stable;
    int lvls, nlt;
    int isP;
    int plane;    /* 0 = luma (filtered inverse), 1/2 = chroma */
    int quant;    /* frame quant (for the inverse's nudge bounds) */
    int do_quant; /* forward: fuse quantise+dequantise into the epilogue */
    int tiles_x, tiles_y, tile_base;
    FastDiv tiles_x_fd, mtiles_x_fd; /* tile index -> (row, column) without an integer divide per thread */
    int mtiles_x, mtiles_y, mtile_base; /* mid kernel: 128x64 blocks of LL_2 */
    int hqp[16];  /* inverse: nudge bound per level (sbt.c:677-696), index = level */
    PlaneQ pq;
    DvGeom dg;
};

/* host helpers (sbt_host.cu) */
int sbt_num_levels(int cw, int ch);
void sbt_fill_geometry(SbtJob *j, int pw, int ph, int cw, int ch, int isP, int plane);
void sbt_fill_quant(SbtJob *j, int q, int isP, int plane, int nbh, int nbv);
size_t sbt_llx_elems(int cw, int ch);
size_t sbt_dv_elems(int cw, int ch);

/* launches; jobs is a DEVICE array, total_tiles = sum of tiles over jobs */
/* grid sizes of one launch over an array of jobs */
struct SbtDims {
    int njobs = 0, tiles = 0, mtiles = 0;
    bool any_intra = false, any_inter = false;
    /* uniform launches (the engines: planes Y,U,V of many pictures of one format): jobs come in groups of gsz
     * with identical tile counts per position, so tile -> job is arithmetic; gsz == 0: binary search */
    int gsz = 0, tg = 0, c0 = 0, c1 = 0, mtg = 0, mc0 = 0, mc1 = 0;
    FastDiv tg_fd = {1ull << 31, 31, 1}, mtg_fd = {1ull << 31, 31, 1};
};
/* assigns tile_base / mtile_base of jobs[0..n) (host copies, in launch order) and returns the totals */
SbtDims sbt_assign_tiles(SbtJob *jobs, int n);

/* ev0/ev1 (optional) are recorded right before / after the tile kernel, for live roofline timing */
void sbt_fwd_launch(const SbtJob *d_jobs, const SbtDims &dims, size_t lo_smem, cudaStream_t st,
                    cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr);
void sbt_inv_launch(const SbtJob *d_jobs, const SbtDims &dims, size_t lo_smem, cudaStream_t st,
                    cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr);
size_t sbt_lo_smem_bytes(int cw, int ch);

/* flat tile index -> (job, tile within job) */
template <bool MID> DSV_D int sbt_locate(const SbtJob *jobs, const SbtDims &d, int bid, int *tile);

/* flat tile index -> job (jobs are sorted by tile_base / mtile_base) */
template <bool MID> DSV_D int sbt_find_job(const SbtJob *jobs, int njobs, int tile)
{
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if ((MID ? jobs[mid].mtile_base : jobs[mid].tile_base) <= tile) {
            lo = mid;
        } else {
            hi = mid - 1;
        }
    }
    return lo;
}

template <bool MID> DSV_D int sbt_locate(const SbtJob *jobs, const SbtDims &d, int bid, int *tile)
{
    if (d.gsz) {
        const int tg = MID ? d.mtg : d.tg, c0 = MID ? d.mc0 : d.c0, c1 = MID ? d.mc1 : d.c1;
        const int grp = (int) fastdiv((unsigned) bid, MID ? d.mtg_fd : d.tg_fd), r = bid - grp * tg;
        const int k = (r >= c0) + (r >= c0 + c1);
        *tile = r - (k == 0 ? 0 : (k == 1 ? c0 : c0 + c1));
        return grp * d.gsz + k;
    }
    const int job = sbt_find_job<MID>(jobs, d.njobs, bid);
    *tile = bid - (MID ? jobs[job].mtile_base : jobs[job].tile_base);
    return job;
}

/* per-level geometry helpers */
DSV_HD int sbt_ws(int dim, int lvl) { return ceil_shift(dim, lvl - 1); } /* input size of level lvl */
DSV_HD int sbt_wo(int dim, int lvl) { return ceil_shift(dim, lvl); }     /* LL size / H offset */

} // namespace dsv
