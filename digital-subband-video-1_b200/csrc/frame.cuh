/*
 * frame.cuh -- device frames in the reference's bordered layout (frame.c:63-120): three u8 planes in one
 * allocation, 64-sample replicated border on every side, stride = roundup16(w + 128), planes back to
 * back.  Keeping the exact layout means in-bounds-but-past-border reads of the motion search / half-pel
 * filters see the same bytes as the reference; a zeroed guard band in front of
 * and behind the allocation makes the few reads that leave the reference's allocation well defined.
 */
#pragma once
#include "common.cuh"

namespace dsv {

struct DevFrame {
    uint8_t *alloc = nullptr; /* cudaMalloc'd block incl. guards */
    uint8_t *p[3] = {nullptr, nullptr, nullptr}; /* sample (0,0) of each plane */
    int stride[3] = {0, 0, 0};
    int w[3] = {0, 0, 0}, h[3] = {0, 0, 0};
    size_t bytes = 0;
};

#define DSV_GUARD_BYTES (4 * 4096)

void devframe_alloc(DevFrame *f, int width, int height, int subsamp);
void devframe_free(DevFrame *f);

/* replicate the border of planes [0, nplanes) (dsv_extend_frame / dsv_extend_frame_luma, frame.c:263-327) */
void frame_extend_launch(const DevFrame &f, int nplanes, cudaStream_t st);
/* 2x2 rounded box filter of the luma plane incl. its new border (dsv_ds2x_frame_luma + extend, frame.c:240-261) */
void frame_down2_luma_launch(const DevFrame &src, const DevFrame &dst, cudaStream_t st);
/* sum of the luma plane -> *d_sum (unsigned long long); caller divides (dsv_frame_avg_luma, frame.c:223-238) */
void frame_sum_luma_launch(const DevFrame &f, unsigned long long *d_sum, cudaStream_t st);
/* plane-wise copy w x h (dsv_frame_copy without the extension, frame.c:199-217) */
void frame_copy_launch(const DevFrame &dst, const DevFrame &src, cudaStream_t st);

} // namespace dsv
