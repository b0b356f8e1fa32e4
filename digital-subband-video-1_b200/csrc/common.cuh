/*
 * common.cuh -- shared device/host helpers for libdsv1_b200 (sm_100a).
 *
 * Integer semantics follow the reference exactly:
 * C truncating division on negatives for the LL scaling and the /4 of the Haar
 * inverse, round-half-away-from-zero for round2/4/8, arithmetic >> on negatives.
 */
#pragma once

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#ifndef DSV_CPU_EMU
#include <cuda_runtime.h>
/* every launch goes through here: when the calling thread has a KernelTimes collector active (an engine step,
 * ktime.cu) the launch is bracketed by two CUDA events on its stream, keyed by the kernel's name */
#define DSV_LAUNCH(kernel, grid, block, smem, stream, ...)               \
    do {                                                                 \
        static const int kt_slot_ = dsv::kt_slot(#kernel);               \
        dsv::KtScope kt_scope_(kt_slot_, (stream));                      \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);      \
    } while (0)
#define DSV_DYN_SMEM(type, name)                                   \
    extern __shared__ __align__(16) unsigned char dsv_dyn_smem_[]; \
    type *name = reinterpret_cast<type *>(dsv_dyn_smem_)
#endif

#define DSV_HD __host__ __device__ __forceinline__
#define DSV_D __device__ __forceinline__

/* There is no CPU fallback to continue on, but a CUDA failure (out of memory, lost device) must not take the caller's
 * process down either: it is logged and thrown as dsv::CudaError, and every C entry point turns that into the
 * reference's own error return (dsv_enc: no packets + DSV_ERROR, dsv_dec: DSV_DEC_ERROR, dsvb_ / dsvk_: negative / NULL). */
namespace dsv {
struct CudaError {
    int code;
};
struct Unsupported {
};
[[noreturn]] void cuda_fail(int code, const char *msg, const char *file, int line, const char *expr);
} // namespace dsv
#define CUDA_CHECK(expr)                                                                            \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            dsv::cuda_fail((int) e_, cudaGetErrorString(e_), __FILE__, __LINE__, #expr);            \
        }                                                                                           \
    } while (0)
#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())
/* first and last line of a C entry point's body: nothing thrown below crosses the C ABI */
#define DSV_API_BEGIN try {
#define DSV_API_END(ret)                                                                            \
    }                                                                                               \
    catch (const dsv::CudaError &)                                                                  \
    {                                                                                               \
        return ret;                                                                                 \
    }                                                                                               \
    catch (const dsv::Unsupported &)                                                                \
    {                                                                                               \
        return ret;                                                                                 \
    }                                                                                               \
    catch (const std::bad_alloc &)                                                                  \
    {                                                                                               \
        fprintf(stderr, "[dsv1_b200] out of host memory\n");                                        \
        return ret;                                                                                 \
    }

namespace dsv {

/* ---- per-kernel live timing (ktime.cu) ---------------------------------------------------------- */
#define KT_MAX_SLOTS 64
int kt_slot(const char *kernel_name);      /* registers the name on first use; -1 when the table is full */
int kt_count();
const char *kt_name(int slot);
struct KernelTimes {
    struct Rec {
        int slot;
        cudaEvent_t e0, e1;
    };
    /* two generations: a decoder step's records are read when its parity comes round again */
    Rec *recs[2] = {nullptr, nullptr};
    int n[2] = {0, 0}, cap[2] = {0, 0};
    int gen = 0;
    double ms[KT_MAX_SLOTS] = {0};
    unsigned long long launches[KT_MAX_SLOTS] = {0};
    void open(int generation) { gen = generation; }
    Rec *next();
    void collect(int generation); /* the events of that generation must have completed */
    void reset();
    void destroy();
};
extern thread_local KernelTimes *kt_current;
struct KtActivate { /* RAII: launches of this thread are attributed to `t` while the object lives */
    KernelTimes *prev;
    explicit KtActivate(KernelTimes *t) : prev(kt_current) { kt_current = t; }
    ~KtActivate() { kt_current = prev; }
};
#ifndef DSV_CPU_EMU
struct KtScope {
    KernelTimes::Rec *r = nullptr;
    cudaStream_t st;
    KtScope(int slot, cudaStream_t s) : st(s)
    {
        if (kt_current && slot >= 0) {
            r = kt_current->next();
            r->slot = slot;
            cudaEventRecord(r->e0, st);
        }
    }
    ~KtScope()
    {
        if (r) {
            cudaEventRecord(r->e1, st);
        }
    }
};
#endif

DSV_HD int imin(int a, int b) { return a < b ? a : b; }
DSV_HD int imax(int a, int b) { return a > b ? a : b; }
DSV_HD int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
DSV_HD int iabs(int v) { return v < 0 ? -v : v; }
DSV_HD int ceil_shift(int v, int s) { return (v + (1 << s) - 1) >> s; }
DSV_HD int ceil_div(int a, int b) { return (a + b - 1) / b; }
DSV_HD uint8_t clamp_u8(int v) { return (uint8_t) (v < 0 ? 0 : (v > 255 ? 255 : v)); }

/* smallest L with 2^L >= n (hzcc.c:437-447) */
DSV_HD int lb2(unsigned n)
{
    int l = 0;
    while ((1u << l) < n) {
        l++;
    }
    return l;
}

/* per-level base quantiser (hzcc.c:77-92) */
DSV_HD int get_quant(int q, int isP, int level)
{
    if (isP) {
        q = q * 3 / 2;
    }
    if (level == 1) {
        q = q * 2 / 3;
    } else if (level == 2) {
        q = q * 3 / 2;
    }
    return q < 16 ? 16 : q;
}

/* round half away from zero, divide by 2^S (sbt.c:63-88) */
template <int S> DSV_HD int rnd_shift(int v)
{
    /* v < 0: -((-v + half) >> S) == ceil((v - half) / 2^S) == (v + half - 1) >> S; branch-free with the sign word */
    return (v + (1 << (S - 1)) + (v >> 31)) >> S;
}

/* LL scaling, C truncation (sbt.c:20-21) */
DSV_HD int ll_down(int v) { return v * 4 / 5; }
DSV_HD int ll_up(int v) { return v * 5 / 4; }
/* C `/ 4` (truncate toward zero) without a divide */
DSV_HD int div4_trunc(int v) { return v / 4; } /* the compiler's sign word + LEA.HI + shift: 3 instructions */

/* m such that (i * m) >> 20 == i / d for 0 <= i < 1024 and 1 <= d <= 64: floor(2^20 / d) + 1.  On the device the
 * floor comes from a correctly rounded float reciprocal (error < 0.125 / d, below the smallest non-zero fractional
 * part 1 / d of 2^20 / d, so the truncation is exact) instead of a ~25-instruction integer divide. */
DSV_HD int magic20(int d)
{
#ifdef __CUDA_ARCH__
    return (int) (1048576.0f * __frcp_rn((float) d)) + 1;
#else
    return (1 << 20) / d + 1;
#endif
}

/*
 * Exact unsigned division by a runtime-constant divisor d for numerators < 2^31:
 *   n / d == (n * mul) >> sh,  sh = 31 + ceil(log2 d),  mul = floor(2^sh / d) + 1
 * (round-up method; for powers of two mul = 2^31 exactly).  Built on the host once
 * per frame for the handful of quantisers a frame uses.
 */
struct FastDiv {
    unsigned long long mul;
    int sh;
    int d;
};
static inline FastDiv make_fastdiv(int d)
{
    FastDiv f;
    int l = lb2((unsigned) d);
    f.d = d;
    f.sh = 31 + l;
    if ((1 << l) == d) {
        f.mul = 1ull << 31;
    } else {
        f.mul = ((1ull << f.sh) / (unsigned long long) d) + 1;
    }
    return f;
}
DSV_HD unsigned fastdiv(unsigned n, const FastDiv &f) { return (unsigned) (((unsigned long long) n * f.mul) >> f.sh); }

/* dead-zone quantiser / reconstruction (hzcc.c:94-128) */
DSV_HD int dz_quant(int v, int q, const FastDiv &two_q)
{
    unsigned m = (unsigned) iabs(v) * 2u;
    if (m <= (unsigned) q) {
        return 0;
    }
    int r = (int) fastdiv(m + 1u, two_q);
    return v < 0 ? -r : r;
}
DSV_HD int dz_dequant(int v, int q)
{
    int m = (iabs(v) * (2 * q) + q) >> 1;
    return v < 0 ? -m : m;
}
/* top level: shift of the magnitude (hzcc.c:114-135) */
DSV_HD int p2_quant(int v, int s) { return v < 0 ? -((-v) >> s) : (v >> s); }
DSV_HD int p2_dequant(int v, int s) { return (int) ((unsigned) v << s); }

/* 4 bytes at an arbitrary address: two aligned loads + funnel shift (any address space); reads up to 3 bytes
 * past p + 3, which every frame allocation covers with its guard band */
DSV_D unsigned ld4u(const uint8_t *p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const unsigned *q = reinterpret_cast<const unsigned *>(a & ~(uintptr_t) 3);
    const unsigned sh = (unsigned) (a & 3) * 8;
    return __funnelshift_r(q[0], q[1], sh); /* sh == 0 yields q[0]; q[1] is always readable (guard band / padding) */
}
/* asynchronous global -> shared copies (cp.async, SASS LDGSTS): the data never occupies a register while in flight,
 * so a thread can keep many rows outstanding.  dst: shared memory; src: global memory, aligned to the copy size. */
DSV_D void cp_async4(void *dst, const void *src)
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned) __cvta_generic_to_shared(dst)), "l"(__cvta_generic_to_global(src)) : "memory");
#else
    *reinterpret_cast<uint32_t *>(dst) = *reinterpret_cast<const uint32_t *>(src);
#endif
}
DSV_D void cp_async8(void *dst, const void *src)
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned) __cvta_generic_to_shared(dst)), "l"(__cvta_generic_to_global(src)) : "memory");
#else
    *reinterpret_cast<uint64_t *>(dst) = *reinterpret_cast<const uint64_t *>(src);
#endif
}
DSV_D void cp_async_commit()
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N> DSV_D void cp_async_wait() /* at most N of this thread's committed groups still pending */
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

DSV_HD int byte_of(unsigned w, int i) { return (int) ((w >> (8 * i)) & 0xff); }
/* four ints -> four saturated bytes, a in the lowest byte: two cvt.pack.sat instructions on the device */
DSV_HD unsigned pack_u8x4(int a, int b, int c, int d)
{
#if defined(__CUDA_ARCH__)
    unsigned hi, r;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(hi));
    return r;
#else
    return (unsigned) clamp_u8(a) | ((unsigned) clamp_u8(b) << 8) | ((unsigned) clamp_u8(c) << 16) | ((unsigned) clamp_u8(d) << 24);
#endif
}
/* per-byte (a + b + 1) >> 1 */
/* the half-pel kernel -a + 9b + 9c - d (bmc.c:124-174, hme.c:302-348) over the four bytes of w, a in the lowest
 * byte: one dp4a with the taps as signed bytes */
DSV_HD int hp_taps_u8x4(unsigned w)
{
#if defined(__CUDA_ARCH__)
    int r;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0xff0909ffu), "r"(0));
    return r;
#else
    return 9 * (int) (((w >> 8) & 0xffu) + ((w >> 16) & 0xffu)) - (int) ((w & 0xffu) + (w >> 24));
#endif
}
/* c + lo16(v) * t0 + hi16(v) * t1: v holds two signed 16-bit values, t0 / t1 are the two low signed bytes of taps */
DSV_HD int dp2a_lo_s16(unsigned v, unsigned taps, int c)
{
#if defined(__CUDA_ARCH__)
    int r;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(taps), "r"(c));
    return r;
#else
    return c + (int) (int16_t) (v & 0xffffu) * (int) (int8_t) (taps & 0xffu) + (int) (int16_t) (v >> 16) * (int) (int8_t) ((taps >> 8) & 0xffu);
#endif
}
DSV_HD unsigned avg_up_u8x4(unsigned a, unsigned b) { return (a | b) - (((a ^ b) >> 1) & 0x7f7f7f7fu); }

/* Reference frame geometry (frame.c:63-120): 64-sample border, stride rounded up to 16 */
#define DSV_BORDER 64
DSV_HD int frame_stride(int w) { return (w + 2 * DSV_BORDER + 15) & ~15; }

} // namespace dsv
