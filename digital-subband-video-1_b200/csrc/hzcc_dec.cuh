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

struct HzDecBufs {
    void *d_jobs;
    void *h_jobs;
    int32_t *runs[3], *vals[3];
    void *cta_sum[3];
    unsigned *cta_entry[3];
    unsigned long long *blk_sum[3];
    unsigned *first_bad[3];
    HzDecPlan plan[3];
};

/* host: read SEG(DC), nruns and the first run from the plane head (hzcc.c:309-316,485-486) */
void hzdec_parse_head(const uint8_t *host_body, unsigned avail, unsigned plen, HzPlaneData *pd);
void hzdec_plan(HzDecPlan *pl, int cw, int ch);
void hzdec_alloc(HzDecBufs *b, const HzDecPlan pl[3]);
void hzdec_free(HzDecBufs *b);
void hzdec_launch(HzDecBufs *b, const HzJob *hz, const HzPlaneData *pd, int nplanes, cudaStream_t st);

} // namespace dsv
