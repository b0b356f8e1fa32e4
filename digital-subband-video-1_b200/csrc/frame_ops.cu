/* frame_ops.cu -- border extension, luma pyramid level, luma sum, plane copies (frame.c). */
#include "frame.cuh"

namespace dsv {

void devframe_alloc(DevFrame *f, int width, int height, int subsamp)
{
    const int hs = (subsamp >> 2) & 3, vs = subsamp & 3;
    const int cw = ceil_shift(width, hs), ch = ceil_shift(height, vs);
    size_t len[3];
    f->w[0] = width; f->h[0] = height;
    f->w[1] = f->w[2] = cw; f->h[1] = f->h[2] = ch;
    for (int c = 0; c < 3; c++) {
        f->stride[c] = frame_stride(f->w[c]);
        len[c] = (size_t) f->stride[c] * (f->h[c] + 2 * DSV_BORDER);
    }
    f->bytes = len[0] + len[1] + len[2] + 2 * DSV_GUARD_BYTES;
    CUDA_CHECK(cudaMalloc(&f->alloc, f->bytes));
    CUDA_CHECK(cudaMemset(f->alloc, 0, f->bytes));
    uint8_t *at = f->alloc + DSV_GUARD_BYTES;
    for (int c = 0; c < 3; c++) {
        f->p[c] = at + (size_t) f->stride[c] * DSV_BORDER + DSV_BORDER;
        at += len[c];
    }
}

void devframe_free(DevFrame *f)
{
    if (f->alloc) {
        cudaFree(f->alloc);
        f->alloc = nullptr;
    }
}

struct PlaneRef {
    uint8_t *p;
    int stride, w, h;
};
struct ExtendArgs {
    PlaneRef pl[3];
    int n;
};

/* out(x,y) = in(clamp(x,0,w-1), clamp(y,0,h-1)) for every border sample; 16 bytes per thread */
__global__ void __launch_bounds__(256) frame_extend_kernel(ExtendArgs a)
{
    const PlaneRef P = a.pl[blockIdx.z];
    const int chunks = (P.w + 2 * DSV_BORDER + 15) >> 4;
    const int ck = (int) (blockIdx.x * blockDim.x + threadIdx.x);
    const int y = (int) blockIdx.y - DSV_BORDER;
    if (ck >= chunks || y >= P.h + DSV_BORDER) {
        return;
    }
    const int x0 = ck * 16 - DSV_BORDER;
    const bool yin = y >= 0 && y < P.h;
    if (yin && x0 >= 0 && x0 + 16 <= P.w) {
        return; /* interior */
    }
    const int sy = iclamp(y, 0, P.h - 1);
    const uint8_t *src = P.p + (size_t) sy * P.stride;
    uint8_t *dst = P.p + (ptrdiff_t) y * P.stride + x0;
    const int xend = P.w + DSV_BORDER;
#pragma unroll 4
    for (int e = 0; e < 16; e++) {
        int x = x0 + e;
        if (x < xend && !(yin && x >= 0 && x < P.w)) {
            dst[e] = src[iclamp(x, 0, P.w - 1)];
        }
    }
}

void frame_extend_launch(const DevFrame &f, int nplanes, cudaStream_t st)
{
    ExtendArgs a;
    int maxw = 0, maxh = 0;
    a.n = nplanes;
    for (int c = 0; c < 3; c++) {
        a.pl[c].p = f.p[c]; a.pl[c].stride = f.stride[c]; a.pl[c].w = f.w[c]; a.pl[c].h = f.h[c];
        if (c < nplanes) {
            maxw = imax(maxw, f.w[c]);
            maxh = imax(maxh, f.h[c]);
        }
    }
    dim3 grid(ceil_div(ceil_div(maxw + 2 * DSV_BORDER, 16), 256), maxh + 2 * DSV_BORDER, nplanes);
    DSV_LAUNCH(frame_extend_kernel, grid, dim3(256), 0, st, a);
    KERNEL_CHECK();
}

/* dst(i,j) = (s(2i,2j) + s(2i+1,2j) + s(2i,2j+1) + s(2i+1,2j+1) + 2) >> 2; may read one sample into the
 * source border (odd source sizes), which is why the source must be extended first */
__global__ void __launch_bounds__(256) frame_down2_kernel(PlaneRef s, PlaneRef d)
{
    const int x = (int) (blockIdx.x * blockDim.x + threadIdx.x), y = (int) blockIdx.y;
    if (x >= d.w || y >= d.h) {
        return;
    }
    const uint8_t *sp = s.p + (size_t) (2 * y) * s.stride + 2 * x;
    d.p[(size_t) y * d.stride + x] = (uint8_t) ((sp[0] + sp[1] + sp[s.stride] + sp[s.stride + 1] + 2) >> 2);
}

void frame_down2_luma_launch(const DevFrame &src, const DevFrame &dst, cudaStream_t st)
{
    PlaneRef s{src.p[0], src.stride[0], src.w[0], src.h[0]}, d{dst.p[0], dst.stride[0], dst.w[0], dst.h[0]};
    DSV_LAUNCH(frame_down2_kernel, dim3(ceil_div(d.w, 256), d.h), dim3(256), 0, st, s, d);
    KERNEL_CHECK();
    frame_extend_launch(dst, 1, st);
}

__global__ void __launch_bounds__(256) frame_sum_kernel(PlaneRef s, unsigned long long *out)
{
    unsigned acc = 0;
    const int y = (int) blockIdx.x;
    for (int x = threadIdx.x; x < s.w; x += 256) {
        acc += s.p[(size_t) y * s.stride + x];
    }
    acc = __reduce_add_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && acc) {
        atomicAdd(out, (unsigned long long) acc);
    }
}

void frame_sum_luma_launch(const DevFrame &f, unsigned long long *d_sum, cudaStream_t st)
{
    PlaneRef s{f.p[0], f.stride[0], f.w[0], f.h[0]};
    CUDA_CHECK(cudaMemsetAsync(d_sum, 0, sizeof(unsigned long long), st));
    DSV_LAUNCH(frame_sum_kernel, dim3(s.h), dim3(256), 0, st, s, d_sum);
    KERNEL_CHECK();
}

void frame_copy_launch(const DevFrame &dst, const DevFrame &src, cudaStream_t st)
{
    for (int c = 0; c < 3; c++) {
        CUDA_CHECK(cudaMemcpy2DAsync(dst.p[c], dst.stride[c], src.p[c], src.stride[c], src.w[c], src.h[c],
                                     cudaMemcpyDeviceToDevice, st));
    }
}

} // namespace dsv
