/*
 * out420.cu -- the CLI's -out420p conversion of decoder output (dsv_main.c:674-699) as a device kernel, so that a
 * 4:4:4 / 4:2:2 stream decoded through the batch API leaves the GPU as 4:2:0 (half / two thirds of the D2H bytes).
 *
 * The reference converts chroma in two separate passes over host frames (util.c:54-93):
 *   conv444to422: d[x]    = (s[2x] + s[min(2x+1, w-1)] + 1) >> 1            per row
 *   conv422to420: d[y][x] = (s[2y][x] + s[min(2y+1, h-1)][x] + 1) >> 1      per column
 * 4:4:4 takes both (the intermediate is rounded to u8 in between), 4:2:2 the second only.  For 4:1:1 the CLI runs
 * the second pass on the quarter-width planes as they are: the right half of the 4:2:0 chroma planes stays zero
 * (dsv_mk_frame memory is zeroed); reproduced as is.  Luma is copied unchanged (pack_kernel).
 *
 * One thread produces 4 horizontally adjacent output samples from a 2 x 8 (4:4:4) or 2 x 4 source patch.
 */
#include "frame.cuh"

namespace dsv {

#define O4_BX 64
#define O4_BY 4

DSV_D int o4_src(const PlaneRef &S, int x, int y, int hpass)
{
    const uint8_t *row = S.p + (size_t) y * S.stride;
    if (hpass) {
        const int x1 = 2 * x + 1 < S.w ? 2 * x + 1 : S.w - 1;
        return (row[2 * x] + row[x1] + 1) >> 1;
    }
    return row[x];
}

__global__ void __launch_bounds__(O4_BX *O4_BY) to420_kernel(const To420Item *items)
{
    const To420Item it = items[blockIdx.z];
    const PlaneRef S = it.src;
    const int x0 = (int) (blockIdx.x * O4_BX + threadIdx.x) * 4, y = (int) (blockIdx.y * O4_BY + threadIdx.y);
    if (x0 >= it.dw || y >= it.dh) {
        return;
    }
    const int sw = it.hpass ? (S.w + 1) >> 1 : S.w; /* width after the horizontal pass */
    const int y0 = 2 * y, y1 = 2 * y + 1 < S.h ? 2 * y + 1 : S.h - 1;
    uint8_t *dst = it.dst + (size_t) y * it.dw + x0;
    unsigned v[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int x = x0 + e;
        v[e] = x < sw ? (unsigned) ((o4_src(S, x, y0, it.hpass) + o4_src(S, x, y1, it.hpass) + 1) >> 1) : 0u;
    }
    if (x0 + 4 <= it.dw && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
        *reinterpret_cast<unsigned *>(dst) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
        return;
    }
    for (int e = 0; e < 4 && x0 + e < it.dw; e++) {
        dst[e] = (uint8_t) v[e];
    }
}

void to420_launch(const To420Item *d_items, int n, int max_dw, int max_dh, cudaStream_t st)
{
    if (n > 0) {
        DSV_LAUNCH(to420_kernel, dim3(ceil_div(ceil_div(max_dw, 4), O4_BX), ceil_div(max_dh, O4_BY), n), dim3(O4_BX, O4_BY), 0, st, d_items);
        KERNEL_CHECK();
    }
}

} // namespace dsv
