/*
 * tests/emu/cuda_emu.h -- TEST TOOLING ONLY.  Never part of the product build.
 *
 * A minimal CPU emulation of the CUDA execution model, just enough to compile the
 * product's .cu sources with g++ (-x c++ -include cuda_emu.h -DDSV_CPU_EMU) and run
 * the kernels' *logic* in this GPU-less container before spending B200 time:
 *   - one OS thread per CUDA thread of the running block, blocks run one after another;
 *   - __syncthreads / warp shuffles / ballots via std::barrier (threads that return
 *     early drop out, as on the device);
 *   - static __shared__ becomes `static` (valid because blocks run sequentially);
 *   - atomics via the GCC __atomic builtins; cudaMalloc/Memcpy/Memset map to libc.
 * The resulting tests/emu/_build/libdsv1_emu.so is loaded ONLY by the "emu" tests
 * (tests/test_emu_*.py).  libdsv1_b200.so (nvcc, sm_100a) has no such path.
 */
#pragma once
#ifdef DSV_CPU_EMU

#include <algorithm>
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_e { unsigned x, y, z; };
struct uchar4 { unsigned char x, y, z, w; };
struct char4 { signed char x, y, z, w; };
struct short2 { short x, y; };
struct short4 { short x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }
static inline short2 make_short2(short x, short y) { return short2{x, y}; }

namespace emu {
struct BlockCtx {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<std::unique_ptr<std::barrier<>>> wbar;
    std::vector<unsigned long long> xchg; /* one slot per thread for shuffles */
    unsigned nthreads = 0;
};
extern BlockCtx *g_blk;
extern thread_local uint3_e t_threadIdx, t_blockIdx;
extern thread_local unsigned t_lin; /* linear thread id in block */
extern dim3 g_blockDim, g_gridDim;
extern unsigned char *g_dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
} // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define warpSize 32

static inline void __syncthreads() { emu::g_blk->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::g_blk->wbar[emu::t_lin >> 5]->arrive_and_wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

/* warp exchange: every lane of the (converged) warp must call */
template <typename T> static inline T emu_warp_read(T v, unsigned src_lane)
{
    static_assert(sizeof(T) <= 8, "shuffle payload");
    unsigned long long raw = 0;
    unsigned base = emu::t_lin & ~31u;
    unsigned wsz = std::min(32u, emu::g_blk->nthreads - base);
    memcpy(&raw, &v, sizeof(T));
    emu::g_blk->xchg[emu::t_lin] = raw;
    __syncwarp();
    unsigned long long got = emu::g_blk->xchg[base + (src_lane < wsz ? src_lane : (emu::t_lin & 31))];
    __syncwarp();
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
    unsigned lane = emu::t_lin & 31;
    unsigned s = (lane & ~(unsigned) (width - 1)) + ((unsigned) src & (unsigned) (width - 1));
    return emu_warp_read(v, s);
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32)
{
    unsigned lane = emu::t_lin & 31;
    unsigned lo = lane & ~(unsigned) (width - 1);
    unsigned s = (lane - lo >= d) ? lane - d : lane;
    return emu_warp_read(v, s);
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32)
{
    unsigned lane = emu::t_lin & 31;
    unsigned hi = (lane | (unsigned) (width - 1));
    unsigned s = (lane + d <= hi) ? lane + d : lane;
    return emu_warp_read(v, s);
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32)
{
    (void) width;
    return emu_warp_read(v, (emu::t_lin & 31) ^ (unsigned) m);
}
static inline unsigned __ballot_sync(unsigned, int pred)
{
    unsigned base = emu::t_lin & ~31u;
    unsigned wsz = std::min(32u, emu::g_blk->nthreads - base);
    emu::g_blk->xchg[emu::t_lin] = pred ? 1 : 0;
    __syncwarp();
    unsigned r = 0;
    for (unsigned i = 0; i < wsz; i++) {
        r |= (unsigned) (emu::g_blk->xchg[base + i] & 1) << i;
    }
    __syncwarp();
    return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p)
{
    unsigned base = emu::t_lin & ~31u;
    unsigned wsz = std::min(32u, emu::g_blk->nthreads - base);
    unsigned full = wsz == 32 ? 0xffffffffu : ((1u << wsz) - 1);
    return __ballot_sync(m, p) == full;
}
/* block-wide OR of a predicate with barrier semantics: every thread of the block must call */
static inline int __syncthreads_or(int pred)
{
    emu::g_blk->xchg[emu::t_lin] = pred ? 1 : 0;
    __syncthreads();
    int r = 0;
    for (unsigned i = 0; i < emu::g_blk->nthreads; i++) {
        r |= (int) (emu::g_blk->xchg[i] & 1);
    }
    __syncthreads();
    return r;
}
static inline unsigned __activemask() { return 0xffffffffu; }
static inline int __reduce_add_sync(unsigned m, int v)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(m, v, o);
    return v;
}
static inline unsigned __reduce_add_sync(unsigned m, unsigned v)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(m, v, o);
    return v;
}
static inline int __reduce_max_sync(unsigned m, int v)
{
    for (int o = 16; o; o >>= 1) v = std::max(v, __shfl_xor_sync(m, v, o));
    return v;
}
static inline int __reduce_min_sync(unsigned m, int v)
{
    for (int o = 16; o; o >>= 1) v = std::min(v, __shfl_xor_sync(m, v, o));
    return v;
}
static inline unsigned __reduce_min_sync(unsigned m, unsigned v)
{
    for (int o = 16; o; o >>= 1) v = std::min(v, __shfl_xor_sync(m, v, o));
    return v;
}
static inline unsigned __reduce_or_sync(unsigned m, unsigned v)
{
    for (int o = 16; o; o >>= 1) v |= __shfl_xor_sync(m, v, o);
    return v;
}

/* bit / SIMD intrinsics */
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned) v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long) v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v)
{
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s)
{
    unsigned long long t = ((unsigned long long) b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned byte = (unsigned) ((t >> (8 * (sel & 7))) & 0xff);
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0;
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __vabsdiffu4(unsigned a, unsigned b)
{
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        int x = (a >> (8 * i)) & 0xff, y = (b >> (8 * i)) & 0xff;
        r |= (unsigned) (x > y ? x - y : y - x) << (8 * i);
    }
    return r;
}
static inline unsigned __vaddss4(unsigned a, unsigned b)
{
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        int v = (int) (signed char) (a >> (8 * i)) + (int) (signed char) (b >> (8 * i));
        v = v > 127 ? 127 : (v < -128 ? -128 : v);
        r |= (unsigned) (v & 0xff) << (8 * i);
    }
    return r;
}
static inline unsigned __vsubss4(unsigned a, unsigned b)
{
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        int v = (int) (signed char) (a >> (8 * i)) - (int) (signed char) (b >> (8 * i));
        v = v > 127 ? 127 : (v < -128 ? -128 : v);
        r |= (unsigned) (v & 0xff) << (8 * i);
    }
    return r;
}
static inline unsigned __vcmpgtu4(unsigned a, unsigned b)
{
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (((a >> (8 * i)) & 0xff) > ((b >> (8 * i)) & 0xff) ? 0xffu : 0u) << (8 * i);
    return r;
}
static inline unsigned __vcmpltu4(unsigned a, unsigned b) { return __vcmpgtu4(b, a); }
static inline unsigned __vsadu4(unsigned a, unsigned b)
{
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        int x = (a >> (8 * i)) & 0xff, y = (b >> (8 * i)) & 0xff;
        r += (unsigned) (x > y ? x - y : y - x);
    }
    return r;
}
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c)
{
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
    return c;
}
static inline unsigned __sad(int a, int b, unsigned c) { return c + (unsigned) (a > b ? a - b : b - a); }
static inline unsigned __usad(unsigned a, unsigned b, unsigned c) { return c + (a > b ? a - b : b - a); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned) (((unsigned long long) a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b)
{
    return (unsigned long long) (((unsigned __int128) a * b) >> 64);
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s)
{
    s &= 31;
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s)
{
    s &= 31;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
template <typename T> static inline void __stcs(T *p, T v) { *p = v; }
using std::max;
using std::min;
static inline int max(int a, unsigned b) { return a > (int) b ? a : (int) b; }

/* atomics */
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAnd(unsigned *p, unsigned v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicExch(unsigned long long *p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicMax(int *p, int v)
{
    int o = *p;
    while (o < v && !__atomic_compare_exchange_n(p, &o, v, 0, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return o;
}
static inline unsigned atomicMax(unsigned *p, unsigned v)
{
    unsigned o = *p;
    while (o < v && !__atomic_compare_exchange_n(p, &o, v, 0, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return o;
}
static inline int atomicMin(int *p, int v)
{
    int o = *p;
    while (o > v && !__atomic_compare_exchange_n(p, &o, v, 0, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return o;
}
static inline unsigned atomicMin(unsigned *p, unsigned v)
{
    unsigned o = *p;
    while (o > v && !__atomic_compare_exchange_n(p, &o, v, 0, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return o;
}
static inline unsigned atomicCAS(unsigned *p, unsigned cmp, unsigned v)
{
    __atomic_compare_exchange_n(p, &cmp, v, 0, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
static inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v)
{
    __atomic_compare_exchange_n(p, &cmp, v, 0, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}

/* ---- runtime API subset ------------------------------------------------- */
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef struct emuEvent { double t; } *cudaEvent_t;
#define cudaSuccess 0
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost, cudaMemcpyDefault };
#define cudaStreamNonBlocking 1
#define cudaEventDisableTiming 2
#define cudaHostAllocDefault 0
#define cudaFuncAttributeMaxDynamicSharedMemorySize 8
static inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **) p, n); }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
template <typename T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMallocHost((void **) p, n); }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMallocHost(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = 0)
{
    for (size_t y = 0; y < h; y++) memcpy((char *) d + y * dp, (const char *) s + y * sp, w);
    return 0;
}
static inline cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind k) { return cudaMemcpy2DAsync(d, dp, s, sp, w, h, k); }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = 0; return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = 0; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer; void *hostPointer; };
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) { a->type = cudaMemoryTypeHost; a->device = 0; a->devicePointer = (void *) p; a->hostPointer = (void *) p; return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
double emu_now_ms();
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emuEvent{0}; return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = emu_now_ms(); return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float) (b->t - a->t); return 0; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }

#define DSV_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define DSV_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(emu::g_dyn_smem)

#endif /* DSV_CPU_EMU */
