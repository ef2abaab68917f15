// uvm_prefetch_probe.cu - how long does bringing managed arrays back to the GPU take after the host has read them?
// Reproduces the sequence of the reference's main.cpp on this library: arrays written by a kernel (the parser), read by a
// host loop (main.cpp:46-52), then used by kernels again.  Strategy (argv[1]):
//   0  cudaMemPrefetchAsync + stream sync per array (what hostmem.cpp does)
//   1  the same after cudaMemAdviseSetPreferredLocation(device)
//   2  no prefetch: the kernel faults the pages in
//   3  prefetch in 2 MB pieces
//   4  cudaMemcpyAsync into a device buffer instead (no migration at all)
// nvcc -O2 -o uvm_probe uvm_prefetch_probe.cu
#include <chrono>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
__global__ void fill(int* p, size_t n) { for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) p[i] = (int)i; }
__global__ void sum(const int* p, size_t n, unsigned long long* out) { unsigned long long s = 0; for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) s += p[i]; atomicAdd(out, s); }

int main(int argc, char** argv)
{
    const int strategy = argc > 1 ? atoi(argv[1]) : 0;
    const int rounds = argc > 2 ? atoi(argv[2]) : 3;
    const size_t sizes[5] = {21u << 20, 21u << 20, 42u << 20, 8u << 20, 8u << 20};
    double t0 = now();
    cudaFree(0);
    printf("strategy %d: context %.1f ms\n", strategy, now() - t0);
    int* a[5];
    int* dev[5];
    unsigned long long* out;
    cudaMalloc(&out, 8);
    for (int k = 0; k < 5; ++k) {
        cudaMallocManaged(&a[k], sizes[k]);
        cudaMalloc(&dev[k], sizes[k]);
        fill<<<296, 256>>>(a[k], sizes[k] / 4);
    }
    cudaDeviceSynchronize();
    for (int r = 0; r < rounds; ++r) {
        t0 = now();
        volatile long long h = 0;
        for (int rep = 0; rep < (argc > 3 ? atoi(argv[3]) : 10); ++rep)
            for (int k = 0; k < 5; ++k)
                for (size_t i = 0; i < sizes[k] / 4; i += 1) h += a[k][i];
        const double t_host = now() - t0;
        double tk[5];
        for (int k = 0; k < 5; ++k) {
            t0 = now();
            if (strategy == 0 || strategy == 1) {
                if (strategy == 1) cudaMemAdvise(a[k], sizes[k], cudaMemAdviseSetPreferredLocation, 0);
                cudaMemPrefetchAsync(a[k], sizes[k], 0, 0);
                cudaStreamSynchronize(0);
            } else if (strategy == 3) {
                for (size_t o = 0; o < sizes[k]; o += 2u << 20) cudaMemPrefetchAsync((char*)a[k] + o, 2u << 20, 0, 0);
                cudaStreamSynchronize(0);
            } else if (strategy == 4) {
                cudaMemcpyAsync(dev[k], a[k], sizes[k], cudaMemcpyDefault, 0);
                cudaStreamSynchronize(0);
            }
            tk[k] = now() - t0;
        }
        t0 = now();
        for (int k = 0; k < 5; ++k) sum<<<296, 256>>>(strategy == 4 ? dev[k] : a[k], sizes[k] / 4, out);
        cudaDeviceSynchronize();
        const double t_kern = now() - t0;
        printf("  round %d: host loop %.0f ms | to GPU: %.2f %.2f %.2f %.2f %.2f ms | kernels %.2f ms\n", r, t_host, tk[0], tk[1], tk[2], tk[3], tk[4], t_kern);
    }
    return 0;
}
