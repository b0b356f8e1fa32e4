/*
 * ktime.cu -- live per-kernel timing for the measurement contract (bench.py "kernels" / "roofline"): every
 * DSV_LAUNCH made while an engine step is active is bracketed by two CUDA events on the launching stream; the
 * elapsed times are accumulated per kernel name when the step has left the GPU.  No reference counterpart.
 */
#include <mutex>

#include "common.cuh"

namespace dsv {

void cuda_fail(int code, const char *msg, const char *file, int line, const char *expr)
{
    fprintf(stderr, "[dsv1_b200] CUDA error %d (%s) at %s:%d: %s\n", code, msg, file, line, expr);
#ifndef DSV_CPU_EMU
    cudaGetLastError(); /* clear the sticky-less error state so that later calls report their own */
#endif
    throw CudaError{code};
}

static std::mutex kt_mutex;
static const char *kt_names[KT_MAX_SLOTS];
static int kt_n = 0;

int kt_slot(const char *kernel_name)
{
    std::lock_guard<std::mutex> lk(kt_mutex);
    for (int i = 0; i < kt_n; i++) {
        if (strcmp(kt_names[i], kernel_name) == 0) {
            return i;
        }
    }
    if (kt_n >= KT_MAX_SLOTS) {
        return -1;
    }
    kt_names[kt_n] = kernel_name;
    return kt_n++;
}

int kt_count()
{
    std::lock_guard<std::mutex> lk(kt_mutex);
    return kt_n;
}

const char *kt_name(int slot)
{
    std::lock_guard<std::mutex> lk(kt_mutex);
    return slot >= 0 && slot < kt_n ? kt_names[slot] : "";
}

thread_local KernelTimes *kt_current = nullptr;

KernelTimes::Rec *KernelTimes::next()
{
    const int g = gen;
    if (n[g] == cap[g]) {
        const int ncap = cap[g] ? cap[g] * 2 : 64;
        Rec *nr = (Rec *) realloc(recs[g], sizeof(Rec) * (size_t) ncap);
        if (!nr) {
            throw std::bad_alloc();
        }
        for (int i = cap[g]; i < ncap; i++) {
            CUDA_CHECK(cudaEventCreate(&nr[i].e0));
            CUDA_CHECK(cudaEventCreate(&nr[i].e1));
        }
        recs[g] = nr;
        cap[g] = ncap;
    }
    return &recs[g][n[g]++];
}

void KernelTimes::collect(int g)
{
    for (int i = 0; i < n[g]; i++) {
        float t = 0;
        if (cudaEventElapsedTime(&t, recs[g][i].e0, recs[g][i].e1) == cudaSuccess) {
            ms[recs[g][i].slot] += t;
            launches[recs[g][i].slot]++;
        }
    }
    n[g] = 0;
}

void KernelTimes::reset()
{
    memset(ms, 0, sizeof(ms));
    memset(launches, 0, sizeof(launches));
}

void KernelTimes::destroy()
{
    for (int g = 0; g < 2; g++) {
        for (int i = 0; i < cap[g]; i++) {
            cudaEventDestroy(recs[g][i].e0);
            cudaEventDestroy(recs[g][i].e1);
        }
        free(recs[g]);
        recs[g] = nullptr;
        cap[g] = n[g] = 0;
    }
}

} // namespace dsv
