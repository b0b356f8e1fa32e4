/*
 * overlay.cu -- the decoder's debug overlay (DSV_DECODER.draw_info, dsv_decoder.h:38-41) as a device kernel.
 *
 * What the reference does (dsv_decoder.c:147-243, 441-447): on P pictures it clones the output frame WITHOUT a
 * border (so the decoder's own reference stays clean) and paints, block by block in raster order, a black grid,
 * a dashed mark on "stable" blocks, the motion vector as a Bresenham line and a white dot per intra quadrant.
 * Later blocks paint over earlier ones, so the result depends on the order.  The intra dots are written without
 * a bounds check: on edge blocks they land past the end of the luma plane of the border-less clone -- in row
 * padding, in a following luma row or in the chroma planes that follow luma in the same allocation
 * (frame.c:63-120).  The kernel reproduces that by addressing the clone linearly and mapping the byte back to a
 * (plane, row, column).
 *
 * Here: the caller has already packed the picture (three dense planes, anywhere in device memory); one CTA per
 * picture walks the blocks in raster order.  The primitives of one block are painted in parallel, two barriers
 * per block keep the order between blocks.  This is a debugging feature; it is sized for correctness.
 */
#include "frame.cuh"
#include "motion.cuh"

namespace dsv {

#define OV_THREADS 256
#define OV_STABHQ 1
#define OV_MOVECS 2
#define OV_IBLOCK 4
#define OV_MODE_INTER 0 /* DSV_MODE_INTER, DSV_MODE_INTRA (dsv.h:129-130) */
#define OV_MODE_INTRA 1

/* a byte of the reference's border-less clone, addressed from luma sample (0,0) */
static __device__ __forceinline__ void clone_store(const DrawItem &it, long off, uint8_t v)
{
    if (off < 0) {
        return;
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const long rs = (it.w[c] + 15) & ~15; /* stride of the clone: roundup16(w), no border */
        const long len = rs * it.h[c];
        if (off < len) {
            const int row = (int) (off / rs), col = (int) (off % rs);
            if (col < it.w[c]) {
                it.dst[c][(size_t) row * it.stride[c] + col] = v;
            }
            return;
        }
        off -= len;
    }
    /* past the clone's allocation: undefined in the reference, dropped here */
}

__global__ void __launch_bounds__(OV_THREADS) overlay_kernel(const DrawItem *items)
{
    const DrawItem it = items[blockIdx.x];
    uint8_t *luma = it.dst[0];
    const int ls = it.stride[0], w = it.w[0], h = it.h[0];
    const int bw = it.blk_w, bh = it.blk_h;
    const int tid = threadIdx.x;
    for (int j = 0; j < it.nbv; j++) {
        const int y = j * bh;
        for (int x = tid; x < w; x += OV_THREADS) { /* dsv_decoder.c:201 (the row's padding is not visible) */
            luma[(size_t) y * ls + x] = 0;
        }
        __syncthreads();
        for (int i = 0; i < it.nbh; i++) {
            const DevMV mv = it.mvs[j * it.nbh + i];
            const int x = i * bw;
            if (x < w) {
                for (int k = y + tid; k < y + bh && k < h; k += OV_THREADS) {
                    luma[(size_t) k * ls + x] = 0;
                }
            }
            if ((it.mode & OV_STABHQ) && (it.stab[j * it.nbh + i] & 1)) {
                const int a = x + bw / 2, b = y + bh / 2;
                for (int k = -(bw / 4) + tid; k <= bw / 4; k += OV_THREADS) {
                    if (b >= 0 && b < h && a + k >= 0 && a + k < w) {
                        luma[(size_t) b * ls + a + k] = (uint8_t) ((k & 1) * 255);
                    }
                }
            }
            __syncthreads();
            if (tid == 0) {
                if ((it.mode & OV_MOVECS) && mv.mode == OV_MODE_INTER) {
                    /* error-driven line from the block centre to centre + vector (vector units taken as samples) */
                    int x0 = x + bw / 2, y0 = y + bh / 2;
                    const int x1 = x0 + mv.x, y1 = y0 + mv.y;
                    const int dx = abs(x1 - x0), dy = abs(y1 - y0);
                    const int sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1;
                    int err = dx - dy;
                    bool first = true; /* the reference paints the start (even for a zero vector) but not the end point */
                    while (first || x0 != x1 || y0 != y1) {
                        if (y0 >= 0 && y0 < h && x0 >= 0 && x0 < w) {
                            luma[(size_t) y0 * ls + x0] = 0;
                        }
                        if (x0 == x1 && y0 == y1) {
                            break;
                        }
                        first = false;
                        const int e2 = 2 * err;
                        if (e2 > -dy) {
                            err -= dy;
                            x0 += sx;
                        }
                        if (e2 < dx) {
                            err += dx;
                            y0 += sy;
                        }
                    }
                }
                if ((it.mode & OV_IBLOCK) && mv.mode == OV_MODE_INTRA) {
                    const long rs = (w + 15) & ~15;
                    for (int q = 0; q < 4; q++) {
                        if (mv.submask & (1 << q)) { /* DSV_MASK_INTRA00..11 = 1,2,4,8 (dsv.h:131-135) */
                            const int a = x + bw * ((q & 1) ? 3 : 1) / 4;
                            const int b = y + bh * ((q & 2) ? 3 : 1) / 4;
                            clone_store(it, (long) b * rs + a, 255);
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
}

void overlay_launch(const DrawItem *d_items, int n, cudaStream_t st)
{
    if (n > 0) {
        DSV_LAUNCH(overlay_kernel, dim3(n), dim3(OV_THREADS), 0, st, d_items);
        KERNEL_CHECK();
    }
}

} // namespace dsv
