// Latency of small control-plane operations while a bulk DMA saturates PCIe in one direction.
// nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o pcie_ctl pcie_ctl.cu ; ./pcie_ctl
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <thread>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__global__ void k_nop(int *p) { if (p && threadIdx.x == 9999) *p = 1; }
__global__ void k_copy(int4 *dst, const int4 *src, int n) { for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i]; }
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
    const size_t big = 256u << 20;
    char *h_big, *d_big; CK(cudaMallocHost(&h_big, big)); CK(cudaMalloc(&d_big, big));
    const int small = 64 << 10; // 64 KB of descriptors
    int4 *h_s, *d_s; CK(cudaMallocHost(&h_s, small)); CK(cudaMalloc(&d_s, small)); memset(h_s, 1, small);
    cudaStream_t st, bg; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&bg, cudaStreamNonBlocking));
    cudaEvent_t ev; CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (int mode = 0; mode < 3; mode++) {
        std::atomic<bool> stop(false);
        std::thread t([&] {
            while (!stop && mode) {
                if (mode == 1) cudaMemcpyAsync(d_big, h_big, big, cudaMemcpyHostToDevice, bg); else cudaMemcpyAsync(h_big, d_big, big, cudaMemcpyDeviceToHost, bg);
                cudaStreamSynchronize(bg);
            }
        });
        std::this_thread::sleep_for(std::chrono::milliseconds(50));
        const char *names[] = {"empty kernel + stream sync", "empty kernel + event sync", "zero-copy read 64 KB (SM) + sync", "zero-copy write 64 KB (SM) + sync",
                               "DMA H2D 64 KB + sync", "DMA D2H 64 KB + sync", "zero-copy read 256 B + sync", "zero-copy write 256 B + sync"};
        for (int op = 0; op < 8; op++) {
            double tot = 0, mx = 0; const int reps = 200;
            for (int r = 0; r < reps + 10; r++) {
                double t0 = now_us();
                switch (op) {
                    case 0: k_nop<<<1, 32, 0, st>>>(nullptr); cudaStreamSynchronize(st); break;
                    case 1: k_nop<<<1, 32, 0, st>>>(nullptr); cudaEventRecord(ev, st); cudaEventSynchronize(ev); break;
                    case 2: k_copy<<<16, 256, 0, st>>>(d_s, h_s, small / 16); cudaStreamSynchronize(st); break;
                    case 3: k_copy<<<16, 256, 0, st>>>(h_s, d_s, small / 16); cudaStreamSynchronize(st); break;
                    case 4: cudaMemcpyAsync(d_s, h_s, small, cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st); break;
                    case 5: cudaMemcpyAsync(h_s, d_s, small, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); break;
                    case 6: k_copy<<<1, 32, 0, st>>>(d_s, h_s, 16); cudaStreamSynchronize(st); break;
                    case 7: k_copy<<<1, 32, 0, st>>>(h_s, d_s, 16); cudaStreamSynchronize(st); break;
                }
                double dt = now_us() - t0;
                if (r >= 10) { tot += dt; mx = dt > mx ? dt : mx; }
            }
            printf("background %-4s  %-36s avg %8.1f us  max %8.1f us\n", mode == 0 ? "none" : mode == 1 ? "H2D" : "D2H", names[op], tot / reps, mx);
        }
        stop = true; t.join();
    }
    return 0;
}
