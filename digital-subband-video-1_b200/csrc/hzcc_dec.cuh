/* hzcc_dec.cuh -- host interface of the HZCC decode kernels (hzcc_dec.cu). */
#pragma once
#include "hzcc.cuh"

namespace dsv {

struct HzDecPlan {
    int cap;          /* scan positions of the plane = max non-zeros that can land */
    size_t max_bits;
    int max_fsm_cta, max_scan_blk;
};

/* where one plane's coded bytes sit inside the packet that was copied to the device */
struct HzPlaneData {
    const uint8_t *body; /* device pointer just after the 32-bit plen field */
    unsigned plen;       /* as read from the stream */
    unsigned avail;      /* readable bytes from body to the end of the packet */
    unsigned tok_bit0;   /* first bit after R0 */
    int nruns, first_run, dc;
};

/* device scratch of one coefficient plane in flight */
struct HzDecPlaneBufs {
    int32_t *runs, *vals;
    void *cta_sum;
    unsigned *cta_entry;
    unsigned long long *blk_sum;
    unsigned *first_bad;
    HzDecPlan plan;
};

/* launch-wide totals accumulated by hzdec_fill_job */
struct HzDecDims {
    int njobs = 0, fsm_total = 0, scan_total = 0;
};

/* flag-guided clean-up of one coefficient plane before it is decoded into (hzdec_clean_kernel) */
struct HzCleanItem {
    int32_t *coef;
    uint8_t *tflags;
    int cw, tiles_x, tiles_y, base;
    int x2, y2;                     /* the level >= 3 corner, always cleared */
    int rx[6], ry[6], rw[6], rh[6]; /* level-2 (0..2) and level-1 (3..5) regions LH, HL, HH */
};
int hz_flag_base(const DvGeom &g);
void hzdec_fill_clean(HzCleanItem *c, const HzJob &h, int tiles_y);
void hzdec_clean_launch(const HzCleanItem *d_items, int n, int max_tiles, cudaStream_t st);

/* host: read SEG(DC), nruns and the first run from the plane head (hzcc.c:309-316,485-486) */
void hzdec_parse_head(const uint8_t *host_body, unsigned avail, unsigned plen, HzPlaneData *pd);
void hzdec_plan(HzDecPlan *pl, int cw, int ch);
void hzdec_plane_alloc(HzDecPlaneBufs *b, const HzDecPlan &pl);
void hzdec_plane_free(HzDecPlaneBufs *b);
size_t hzdec_job_size();
void hzdec_fill_job(void *slot, const HzJob &hz, const HzPlaneData &pd, const HzDecPlaneBufs &b, HzDecDims *dims);
void hzdec_launch_jobs(const void *d_jobs, const HzDecDims &dims, cudaStream_t st);

} // namespace dsv
