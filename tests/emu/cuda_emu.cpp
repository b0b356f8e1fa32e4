/* tests/emu/cuda_emu.cpp -- TEST TOOLING ONLY: block runner for cuda_emu.h */
#include "cuda_emu.h"

#include <chrono>
#include <condition_variable>
#include <mutex>

namespace emu {

BlockCtx *g_blk = nullptr;
thread_local uint3_e t_threadIdx, t_blockIdx;
thread_local unsigned t_lin;
dim3 g_blockDim, g_gridDim;
unsigned char *g_dyn_smem = nullptr;

namespace {
struct Pool {
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_go, cv_done;
    unsigned long long epoch = 0;
    unsigned active = 0, remaining = 0;
    uint3_e blk{0, 0, 0};
    const std::function<void()> *body = nullptr;
    bool quit = false;

    void worker(unsigned id)
    {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv_go.wait(lk, [&] { return quit || epoch != seen; });
                if (quit) {
                    return;
                }
                seen = epoch;
                if (id >= active) {
                    continue;
                }
            }
            t_lin = id;
            t_threadIdx.x = id % g_blockDim.x;
            t_threadIdx.y = (id / g_blockDim.x) % g_blockDim.y;
            t_threadIdx.z = id / (g_blockDim.x * g_blockDim.y);
            t_blockIdx = blk;
            (*body)();
            g_blk->wbar[id >> 5]->arrive_and_drop();
            g_blk->bar->arrive_and_drop();
            {
                std::unique_lock<std::mutex> lk(m);
                if (--remaining == 0) {
                    cv_done.notify_all();
                }
            }
        }
    }
    void ensure(unsigned n)
    {
        while (workers.size() < n) {
            unsigned id = (unsigned) workers.size();
            workers.emplace_back([this, id] { worker(id); });
        }
    }
    void run_block(unsigned n, uint3_e b, const std::function<void()> &f)
    {
        std::unique_lock<std::mutex> lk(m);
        active = n;
        remaining = n;
        blk = b;
        body = &f;
        epoch++;
        cv_go.notify_all();
        cv_done.wait(lk, [&] { return remaining == 0; });
    }
    ~Pool()
    {
        {
            std::unique_lock<std::mutex> lk(m);
            quit = true;
            cv_go.notify_all();
        }
        for (auto &t : workers) {
            t.join();
        }
    }
};
Pool &pool()
{
    static Pool p;
    return p;
}
std::mutex g_launch_mutex;
} // namespace

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body)
{
    std::lock_guard<std::mutex> guard(g_launch_mutex);
    unsigned n = block.x * block.y * block.z;
    std::vector<unsigned char> dyn(smem + 64);
    g_dyn_smem = dyn.data();
    g_blockDim = block;
    g_gridDim = grid;
    pool().ensure(n);
    for (unsigned bz = 0; bz < grid.z; bz++) {
        for (unsigned by = 0; by < grid.y; by++) {
            for (unsigned bx = 0; bx < grid.x; bx++) {
                BlockCtx ctx;
                ctx.nthreads = n;
                ctx.bar.reset(new std::barrier<>(n));
                for (unsigned w = 0; w < (n + 31) / 32; w++) {
                    ctx.wbar.emplace_back(new std::barrier<>(std::min(32u, n - w * 32)));
                }
                ctx.xchg.assign(n, 0);
                g_blk = &ctx;
                pool().run_block(n, uint3_e{bx, by, bz}, body);
                g_blk = nullptr;
            }
        }
    }
    g_dyn_smem = nullptr;
}

} // namespace emu

double emu_now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
