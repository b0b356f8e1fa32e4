/*
 * frame_ops.cu -- batched frame plumbing kernels: ingest (packed -> bordered + border), pack (bordered ->
 * packed), border extension, luma pyramid level (2x2 box filter + border in one pass), luma sum.
 * Replaces dsv_frame_copy / dsv_clone_frame / dsv_extend_frame[_luma] / dsv_ds2x_frame_luma /
 * dsv_frame_avg_luma (frame.c:199-327); dsv_frame_add (bmc.c:304-316) is fused into the inverse transform's
 * store (sbt_inv.cu: SbtJob.addp).
 *
 * Every kernel takes a device array of items (one per plane of every frame in flight); threads own 16
 * consecutive bytes of one bordered row: interior chunks move as one 16-byte load/store, border chunks
 * replicate the nearest edge sample (out(x,y) = in(clamp x, clamp y)).
 */
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

void StepArena::create(size_t bytes)
{
    cap = bytes;
    CUDA_CHECK(cudaMallocHost(&h, bytes));
    CUDA_CHECK(cudaMalloc(&d, bytes));
    used = uploaded = 0;
}
void StepArena::destroy()
{
    if (h) {
        cudaFreeHost(h);
        cudaFree(d);
        h = d = nullptr;
    }
}
void *StepArena::push(size_t bytes, void **dev)
{
    const size_t at = (used + 15) & ~(size_t) 15;
    if (at + bytes > cap) {
        fprintf(stderr, "[dsv1_b200] step arena exhausted (%zu + %zu > %zu)\n", at, bytes, cap);
        abort();
    }
    used = at + bytes;
    *dev = d + at;
    return h + at;
}
bool StepArena::take_upload(CopyItem *it)
{
    if (used <= uploaded) {
        return false;
    }
    const size_t from = uploaded & ~(size_t) 15;
    it->dst = d + from;
    it->src = h + from;
    it->bytes = used - from;
    uploaded = used;
    return true;
}
void StepArena::upload(cudaStream_t st)
{
    if (used > uploaded) {
        const size_t from = uploaded & ~(size_t) 15;
        copy1_launch(d + from, h + from, used - from, st); /* SM copy from mapped pinned memory, see CopyItem */
        uploaded = used;
    }
}

#define FO_BX 64
#define FO_BY 4
#define FO_RY 1 /* row groups per CTA (measured: more rows per CTA is slower -- fewer CTAs in flight to hide latency) */
static dim3 fo_grid(int max_w, int max_h, int n, bool bordered)
{
    const int cols = bordered ? max_w + 2 * DSV_BORDER : max_w, rows = bordered ? max_h + 2 * DSV_BORDER : max_h;
    return dim3(ceil_div(ceil_div(cols, 16), FO_BX), ceil_div(rows, FO_BY * FO_RY), n);
}
/* body runs once per (16-byte chunk ck, row) owned by the thread */
#define FO_FOREACH_ROW()                                                 \
    const int ck = (int) (blockIdx.x * FO_BX + threadIdx.x);             \
    for (int row = (int) (blockIdx.y * FO_BY * FO_RY + threadIdx.y), fo_i = 0; fo_i < FO_RY; fo_i++, row += FO_BY)

DSV_D bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

#ifndef ING_RY
#define ING_RY 4 /* rows per thread in ingest_kernel: their 16-byte loads are all requested before the first store */
#endif
__global__ void __launch_bounds__(FO_BX *FO_BY) ingest_kernel(const IngestItem *items)
{
    const IngestItem it = items[blockIdx.z];
    const PlaneRef D = it.dst;
    const int ck = (int) (blockIdx.x * FO_BX + threadIdx.x);
    const int x0 = ck * 16 - DSV_BORDER;
    if (x0 >= D.w + DSV_BORDER) {
        return;
    }
    const int row0 = (int) (blockIdx.y * FO_BY * ING_RY + threadIdx.y);
    /* A chunk that lies wholly in the left / right border repeats the row's first / last sample (one byte load, one
     * 16-byte store); an interior chunk of a plane whose rows keep the alignment is one 16-byte load.  Either way the
     * ING_RY independent loads are requested before the first store.  (x0 is a multiple of 16.) */
    enum { CK_MIXED, CK_INTERIOR, CK_LEFT, CK_RIGHT };
    int kind = CK_MIXED;
    if (aligned16(D.p) && (D.stride & 15) == 0) {
        if (x0 + 16 <= 0) {
            kind = CK_LEFT;
        } else if (x0 >= D.w) {
            kind = x0 + 16 <= D.w + DSV_BORDER ? CK_RIGHT : CK_MIXED;
        } else if (x0 + 16 <= D.w && (D.w & 15) == 0 && aligned16(it.src)) {
            kind = CK_INTERIOR;
        }
    }
    if (kind != CK_MIXED) {
        uint4 v[ING_RY];
#pragma unroll
        for (int i = 0; i < ING_RY; i++) {
            const int y = row0 + i * FO_BY - DSV_BORDER;
            if (y < D.h + DSV_BORDER) {
                const uint8_t *srow = it.src + (size_t) iclamp(y, 0, D.h - 1) * D.w;
                if (kind == CK_INTERIOR) {
                    v[i] = *reinterpret_cast<const uint4 *>(srow + x0);
                } else {
                    const unsigned b = (unsigned) (kind == CK_LEFT ? srow[0] : srow[D.w - 1]) * 0x01010101u;
                    v[i] = make_uint4(b, b, b, b);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < ING_RY; i++) {
            const int y = row0 + i * FO_BY - DSV_BORDER;
            if (y < D.h + DSV_BORDER) {
                *reinterpret_cast<uint4 *>(D.p + (ptrdiff_t) y * D.stride + x0) = v[i];
            }
        }
        return;
    }
    for (int i = 0; i < ING_RY; i++) {
        const int y = row0 + i * FO_BY - DSV_BORDER;
        if (y >= D.h + DSV_BORDER) {
            break;
        }
        const int sy = iclamp(y, 0, D.h - 1);
        const uint8_t *src = it.src + (size_t) sy * D.w;
        uint8_t *dst = D.p + (ptrdiff_t) y * D.stride + x0;
        if (x0 >= 0 && x0 + 16 <= D.w && aligned16(src + x0)) {
            *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(src + x0);
            continue;
        }
        const int xend = D.w + DSV_BORDER;
#pragma unroll 4
        for (int e = 0; e < 16; e++) {
            const int x = x0 + e;
            if (x < xend) {
                dst[e] = src[iclamp(x, 0, D.w - 1)];
            }
        }
    }
}

#ifndef PK_RY
#define PK_RY 4 /* rows per thread in pack_kernel */
#endif
__global__ void __launch_bounds__(FO_BX *FO_BY) pack_kernel(const PackItem *items)
{
    const PackItem it = items[blockIdx.z];
    const PlaneRef S = it.src;
    const int ck = (int) (blockIdx.x * FO_BX + threadIdx.x);
    if (ck * 16 >= S.w) {
        return;
    }
    const int row0 = (int) (blockIdx.y * FO_BY * PK_RY + threadIdx.y);
    if ((S.w & 15) == 0 && aligned16(it.dst)) { /* every chunk of the packed plane is aligned: PK_RY loads, then the stores */
        uint4 v[PK_RY];
#pragma unroll
        for (int i = 0; i < PK_RY; i++) {
            const int y = row0 + i * FO_BY;
            if (y < S.h) {
                v[i] = *reinterpret_cast<const uint4 *>(S.p + (size_t) y * S.stride + ck * 16);
            }
        }
#pragma unroll
        for (int i = 0; i < PK_RY; i++) {
            const int y = row0 + i * FO_BY;
            if (y < S.h) {
                *reinterpret_cast<uint4 *>(it.dst + (size_t) y * S.w + ck * 16) = v[i];
            }
        }
        return;
    }
    for (int i = 0; i < PK_RY; i++) {
        const int x0 = ck * 16, y = row0 + i * FO_BY;
        if (y >= S.h) {
            break;
        }
        const uint8_t *src = S.p + (size_t) y * S.stride + x0;
        uint8_t *dst = it.dst + (size_t) y * S.w + x0;
        if (x0 + 16 <= S.w && aligned16(dst)) {
            *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(src);
            continue;
        }
        for (int e = 0; e < 16 && x0 + e < S.w; e++) {
            dst[e] = src[e];
        }
    }
}

/* Border replication only touches the frame's rim: the work items are the 16-byte chunks of the 2 x 64 border rows
 * (full bordered width, copied from the first / last row) followed by 16 slots per interior row of which 9 cover
 * the left and right border (4 chunks left, up to 5 right when w is not a multiple of 16). */
#define EXT_SIDE_SLOTS 16
__global__ void __launch_bounds__(FO_BX *FO_BY) extend_kernel(const PlaneRef *items)
{
    const PlaneRef P = items[blockIdx.z];
    const int C = (P.w + 2 * DSV_BORDER + 15) >> 4;
    const int nb = 2 * DSV_BORDER * C;
    const int idx = (int) (blockIdx.x * (FO_BX * FO_BY) + threadIdx.y * FO_BX + threadIdx.x);
    int x0, y;
    if (idx < nb) {
        const int r = idx / C;
        x0 = (idx - r * C) * 16 - DSV_BORDER;
        y = r < DSV_BORDER ? r - DSV_BORDER : P.h + (r - DSV_BORDER);
    } else {
        const int j = idx - nb, c = j & (EXT_SIDE_SLOTS - 1);
        y = j / EXT_SIDE_SLOTS;
        if (y >= P.h || c >= 9) {
            return;
        }
        x0 = c < 4 ? 16 * c - DSV_BORDER : (P.w & ~15) + 16 * (c - 4);
    }
    const bool yin = y >= 0 && y < P.h;
    const int sy = iclamp(y, 0, P.h - 1);
    const uint8_t *src = P.p + (size_t) sy * P.stride;
    uint8_t *dst = P.p + (ptrdiff_t) y * P.stride + x0;
    const int xend = P.w + DSV_BORDER;
    if (aligned16(dst) && x0 + 16 <= xend) {
        if (x0 + 16 <= 0 || x0 >= P.w) { /* all 16 samples repeat the row's first / last sample */
            const unsigned v = (x0 < 0 ? src[0] : src[P.w - 1]) * 0x01010101u;
            *reinterpret_cast<uint4 *>(dst) = make_uint4(v, v, v, v);
            return;
        }
        if (!yin && x0 >= 0 && x0 + 16 <= P.w && aligned16(src + x0)) {
            *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(src + x0);
            return;
        }
    }
#pragma unroll 4
    for (int e = 0; e < 16; e++) {
        const int x = x0 + e;
        if (x < xend && !(yin && x >= 0 && x < P.w)) {
            dst[e] = src[iclamp(x, 0, P.w - 1)];
        }
    }
}

/* dst(i,j) = (s(2i,2j) + s(2i+1,2j) + s(2i,2j+1) + s(2i+1,2j+1) + 2) >> 2 (frame.c:240-261); may read one
 * sample into the source border (odd source sizes), which is why the source must be extended first */
__global__ void __launch_bounds__(FO_BX *FO_BY) down2_kernel(const Down2Item *items)
{
    const Down2Item it = items[blockIdx.z];
    const PlaneRef S = it.src, D = it.dst;
    FO_FOREACH_ROW()
    {
        const int x0 = ck * 16 - DSV_BORDER, y = row - DSV_BORDER;
        if (x0 >= D.w + DSV_BORDER || y >= D.h + DSV_BORDER) {
            continue;
        }
        const int sy = iclamp(y, 0, D.h - 1);
        const uint8_t *s0 = S.p + (size_t) (2 * sy) * S.stride, *s1 = s0 + S.stride;
        uint8_t *dst = D.p + (ptrdiff_t) y * D.stride + x0;
        if (x0 >= 0 && x0 + 16 <= D.w) {
            const uint4 a0 = *reinterpret_cast<const uint4 *>(s0 + 2 * x0), a1 = *reinterpret_cast<const uint4 *>(s0 + 2 * x0 + 16);
            const uint4 b0 = *reinterpret_cast<const uint4 *>(s1 + 2 * x0), b1 = *reinterpret_cast<const uint4 *>(s1 + 2 * x0 + 16);
            const unsigned ra[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const unsigned rb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            unsigned o[4];
    #pragma unroll
            for (int k = 0; k < 4; k++) {
                unsigned r = 0;
    #pragma unroll
                for (int e = 0; e < 4; e++) { /* output sample 4k+e <- source samples 8k+2e, 8k+2e+1 */
                    const unsigned wa = ra[2 * k + (e >> 1)], wb = rb[2 * k + (e >> 1)];
                    const int sh = (e & 1) * 16;
                    const unsigned v = ((wa >> sh) & 0xff) + ((wa >> (sh + 8)) & 0xff) + ((wb >> sh) & 0xff) + ((wb >> (sh + 8)) & 0xff) + 2;
                    r |= (v >> 2) << (8 * e);
                }
                o[k] = r;
            }
            *reinterpret_cast<uint4 *>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
            continue;
        }
        if ((x0 + 16 <= 0 || (x0 >= D.w && x0 + 16 <= D.w + DSV_BORDER)) && aligned16(dst)) {
            /* wholly inside the left / right border: the row's first / last output sample, sixteen times */
            const int sx = x0 < 0 ? 0 : 2 * (D.w - 1);
            const unsigned v = (unsigned) ((s0[sx] + s0[sx + 1] + s1[sx] + s1[sx + 1] + 2) >> 2) * 0x01010101u;
            *reinterpret_cast<uint4 *>(dst) = make_uint4(v, v, v, v);
            continue;
        }
        const int xend = D.w + DSV_BORDER;
    #pragma unroll 4
        for (int e = 0; e < 16; e++) {
            const int x = x0 + e;
            if (x < xend) {
                const int sx = 2 * iclamp(x, 0, D.w - 1);
                dst[e] = (uint8_t) ((s0[sx] + s0[sx + 1] + s1[sx] + s1[sx + 1] + 2) >> 2);
            }
        }
    }
}

__global__ void __launch_bounds__(256) sum_kernel(const SumItem *items)
{
    const SumItem it = items[blockIdx.y];
    const PlaneRef S = it.src;
    const int y = (int) blockIdx.x;
    if (y >= S.h) {
        return;
    }
    unsigned acc = 0;
    for (int x = threadIdx.x; x < S.w; x += 256) {
        acc += S.p[(size_t) y * S.stride + x];
    }
    acc = __reduce_add_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && acc) {
        atomicAdd(it.out, (unsigned long long) acc);
    }
}

DSV_D unsigned add4_clamp(unsigned a, unsigned b)
{
    unsigned r = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int v = (int) ((a >> (8 * e)) & 0xff) + (int) ((b >> (8 * e)) & 0xff) - 128;
        r |= (unsigned) clamp_u8(v) << (8 * e);
    }
    return r;
}

#define ZERO_PER_CTA (256 * 16 * 8)
__global__ void __launch_bounds__(256) zero_kernel(const ZeroItem *items)
{
    const ZeroItem it = items[blockIdx.y];
    const size_t base = (size_t) blockIdx.x * ZERO_PER_CTA;
    uint8_t *p = reinterpret_cast<uint8_t *>(it.p);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const size_t off = base + ((size_t) k * 256 + threadIdx.x) * 16;
        if (off + 16 <= it.bytes) {
            *reinterpret_cast<uint4 *>(p + off) = make_uint4(0, 0, 0, 0);
        } else if (off < it.bytes) {
            for (size_t e = off; e < it.bytes; e++) {
                p[e] = 0;
            }
        }
    }
}

void zero_launch(const ZeroItem *d_items, int n, size_t max_bytes, cudaStream_t st)
{
    if (n > 0 && max_bytes > 0) {
        DSV_LAUNCH(zero_kernel, dim3((unsigned) ((max_bytes + ZERO_PER_CTA - 1) / ZERO_PER_CTA), n), dim3(256), 0, st, d_items);
        KERNEL_CHECK();
    }
}

#define COPY_PER_CTA (256 * 16 * 4)
/* chunks are aligned to the DESTINATION (16-byte stores, important when it is host memory across PCIe); the source
 * is read with whatever alignment it has */
DSV_D void copy_span(uint8_t *dst, const uint8_t *src, size_t bytes, size_t base)
{
    const size_t lead = (size_t) ((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
    const bool src_al = ((reinterpret_cast<uintptr_t>(src) + lead) & 15) == 0;
    if (base == 0 && threadIdx.x == 0) {
        for (size_t e = 0; e < lead && e < bytes; e++) {
            dst[e] = src[e];
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const size_t off = lead + base + ((size_t) k * 256 + threadIdx.x) * 16;
        if (off >= bytes) {
            return;
        }
        if (off + 16 <= bytes) {
            uint4 v;
            if (src_al) {
                v = *reinterpret_cast<const uint4 *>(src + off);
            } else {
                v = make_uint4(ld4u(src + off), ld4u(src + off + 4), ld4u(src + off + 8), ld4u(src + off + 12));
            }
            *reinterpret_cast<uint4 *>(dst + off) = v;
        } else {
            for (size_t e = off; e < bytes; e++) {
                dst[e] = src[e];
            }
        }
    }
}
__global__ void __launch_bounds__(256) copy_kernel(const CopyItem *items)
{
    const CopyItem it = items[blockIdx.y];
    copy_span(reinterpret_cast<uint8_t *>(it.dst), reinterpret_cast<const uint8_t *>(it.src), it.bytes, (size_t) blockIdx.x * COPY_PER_CTA);
}
__global__ void __launch_bounds__(256) copy1_kernel(CopyItem it)
{
    copy_span(reinterpret_cast<uint8_t *>(it.dst), reinterpret_cast<const uint8_t *>(it.src), it.bytes, (size_t) blockIdx.x * COPY_PER_CTA);
}

struct CopyItemsN {
    CopyItem it[COPYN_MAX];
};
__global__ void __launch_bounds__(256) copyn_kernel(const CopyItemsN its)
{
    const CopyItem it = its.it[blockIdx.y];
    copy_span(reinterpret_cast<uint8_t *>(it.dst), reinterpret_cast<const uint8_t *>(it.src), it.bytes, (size_t) blockIdx.x * COPY_PER_CTA);
}
void copyn_launch(const CopyItem *items, int n, cudaStream_t st)
{
    CopyItemsN its;
    memset(&its, 0, sizeof(its));
    int m = 0;
    size_t max_bytes = 0;
    for (int i = 0; i < n; i++) {
        if (items[i].bytes > 0) {
            if (m == COPYN_MAX) { /* more than one launch's worth */
                copyn_launch(items + i, n - i, st);
                break;
            }
            its.it[m++] = items[i];
            max_bytes = items[i].bytes > max_bytes ? items[i].bytes : max_bytes;
        }
    }
    if (m > 0) {
        DSV_LAUNCH(copyn_kernel, dim3((unsigned) ((max_bytes + COPY_PER_CTA - 1) / COPY_PER_CTA), (unsigned) m), dim3(256), 0, st, its);
        KERNEL_CHECK();
    }
}

void copy_launch(const CopyItem *items, int n, size_t max_bytes, cudaStream_t st)
{
    if (n > 0 && max_bytes > 0) {
        DSV_LAUNCH(copy_kernel, dim3((unsigned) ((max_bytes + COPY_PER_CTA - 1) / COPY_PER_CTA), n), dim3(256), 0, st, items);
        KERNEL_CHECK();
    }
}
void copy1_launch(void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    if (bytes > 0) {
        CopyItem it;
        it.dst = dst;
        it.src = src;
        it.bytes = bytes;
        DSV_LAUNCH(copy1_kernel, dim3((unsigned) ((bytes + COPY_PER_CTA - 1) / COPY_PER_CTA)), dim3(256), 0, st, it);
        KERNEL_CHECK();
    }
}

void ingest_launch(const IngestItem *d_items, int n, int max_w, int max_h, cudaStream_t st)
{
    if (n > 0) {
        dim3 grid = fo_grid(max_w, max_h, n, true);
        grid.y = (unsigned) ceil_div(max_h + 2 * DSV_BORDER, FO_BY * ING_RY);
        DSV_LAUNCH(ingest_kernel, grid, dim3(FO_BX, FO_BY), 0, st, d_items);
        KERNEL_CHECK();
    }
}
void pack_launch(const PackItem *d_items, int n, int max_w, int max_h, cudaStream_t st)
{
    if (n > 0) {
        dim3 grid = fo_grid(max_w, max_h, n, false);
        grid.y = (unsigned) ceil_div(max_h, FO_BY * PK_RY);
        DSV_LAUNCH(pack_kernel, grid, dim3(FO_BX, FO_BY), 0, st, d_items);
        KERNEL_CHECK();
    }
}
void extend_launch(const PlaneRef *d_items, int n, int max_w, int max_h, cudaStream_t st)
{
    if (n > 0) {
        const int items = 2 * DSV_BORDER * ((max_w + 2 * DSV_BORDER + 15) >> 4) + EXT_SIDE_SLOTS * max_h;
        DSV_LAUNCH(extend_kernel, dim3(ceil_div(items, FO_BX * FO_BY), 1, n), dim3(FO_BX, FO_BY), 0, st, d_items);
        KERNEL_CHECK();
    }
}
void down2_launch(const Down2Item *d_items, int n, int max_w, int max_h, cudaStream_t st)
{
    if (n > 0) {
        DSV_LAUNCH(down2_kernel, fo_grid(max_w, max_h, n, true), dim3(FO_BX, FO_BY), 0, st, d_items);
        KERNEL_CHECK();
    }
}
void sum_launch(const SumItem *d_items, int n, int max_h, cudaStream_t st)
{
    if (n > 0) {
        DSV_LAUNCH(sum_kernel, dim3(max_h, n), dim3(256), 0, st, d_items);
        KERNEL_CHECK();
    }
}

/* ---- single-frame conveniences (kernel-level API): build a one-off list in device memory ---- */
template <typename T> static T *upload_items(const T *items, int n, cudaStream_t st)
{
    T *d;
    CUDA_CHECK(cudaMalloc(&d, sizeof(T) * (size_t) n));
    CUDA_CHECK(cudaMemcpyAsync(d, items, sizeof(T) * (size_t) n, cudaMemcpyHostToDevice, st));
    return d;
}

void frame_extend_launch(const DevFrame &f, int nplanes, cudaStream_t st)
{
    PlaneRef it[3];
    int mw = 0, mh = 0;
    for (int c = 0; c < nplanes; c++) {
        it[c] = plane_ref(f, c);
        mw = imax(mw, f.w[c]);
        mh = imax(mh, f.h[c]);
    }
    PlaneRef *d = upload_items(it, nplanes, st);
    extend_launch(d, nplanes, mw, mh, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(d);
}

void frame_down2_luma_launch(const DevFrame &src, const DevFrame &dst, cudaStream_t st)
{
    Down2Item it;
    it.src = plane_ref(src, 0);
    it.dst = plane_ref(dst, 0);
    Down2Item *d = upload_items(&it, 1, st);
    down2_launch(d, 1, dst.w[0], dst.h[0], st);
    CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(d);
}

} // namespace dsv
