// runtime.cu -- device selection, memory, error text, launch counter.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace thsp {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void note_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void forget_launches(unsigned n) { g_launches.fetch_sub(n, std::memory_order_relaxed); }

int ensure_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("no usable CUDA device (%s); libthsparse_cuda has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return 1;
    }
    return 0;
}

static constexpr int kMaxDev = 16;
int sm_count()
{
    static int cached[kMaxDev] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

static constexpr int kSlots = 8;
struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
    uint64_t uses = 0;   // handed out this many times: lets a two-step call notice that somebody used the slot in between
};
static Scratch g_scratch[kMaxDev][kSlots];
static std::mutex g_scratch_mu;

void* scratch(size_t bytes, int slot)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev || slot < 0 || slot >= kSlots) return nullptr;
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    Scratch& s = g_scratch[dev][slot];
    ++s.uses;
    if (s.cap < bytes) {
        if (s.p) {
            // growing frees the old buffer, which earlier launches may still use: wait for them.  Not possible while this
            // thread captures a stream (the wait is refused and a pointer baked into a graph would dangle): fail loudly.
            const cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                cudaGetLastError();
                set_error("scratch slot %d cannot grow from %zu to %zu bytes here: %s (stream capture active?)", slot, s.cap, bytes,
                          cudaGetErrorString(e));
                return nullptr;
            }
            cudaFree(s.p);
            s.p = nullptr;
        }
        size_t cap = bytes < 4096 ? 4096 : bytes;
        if (cudaMalloc(&s.p, cap) != cudaSuccess) {
            s.p = nullptr;
            s.cap = 0;
            set_error("scratch allocation of %zu bytes failed", cap);
            return nullptr;
        }
        s.cap = cap;
    }
    return s.p;
}

uint64_t scratch_uses(int slot)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev || slot < 0 || slot >= kSlots) return 0;
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    return g_scratch[dev][slot].uses;
}

}  // namespace thsp

using namespace thsp;

extern "C" {

const char* thsp_version(void) { return "thsparse-b200 0.1 (sm_100a)"; }
const char* thsp_last_error(void) { return g_err; }

int thsp_device_count(int* count)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    if (count) *count = n;
    return 0;
}
int thsp_set_device(int device)
{
    THSP_CUDA(cudaSetDevice(device));
    return 0;
}
int thsp_get_device(int* device)
{
    THSP_CUDA(cudaGetDevice(device));
    return 0;
}
int thsp_sm_count(int* count)
{
    if (ensure_device()) return 1;
    *count = sm_count();
    return 0;
}
int thsp_malloc(void** ptr, size_t bytes)
{
    if (ensure_device()) return 1;
    THSP_CUDA(cudaMalloc(ptr, bytes ? bytes : 16));
    return 0;
}
int thsp_malloc_managed(void** ptr, size_t bytes)
{
    if (ensure_device()) return 1;
    THSP_CUDA(cudaMallocManaged(ptr, bytes ? bytes : 16, cudaMemAttachGlobal));
    return 0;
}
int thsp_malloc_host(void** ptr, size_t bytes)
{
    if (ensure_device()) return 1;
    THSP_CUDA(cudaMallocHost(ptr, bytes ? bytes : 16));
    return 0;
}
int thsp_host_register(void* ptr, size_t bytes)
{
    if (ensure_device()) return 1;
    THSP_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return 0;
}
int thsp_host_unregister(void* ptr)
{
    if (ptr) THSP_CUDA(cudaHostUnregister(ptr));
    return 0;
}
int thsp_free(void* ptr)
{
    if (ptr) THSP_CUDA(cudaFree(ptr));
    return 0;
}
int thsp_free_host(void* ptr)
{
    if (ptr) THSP_CUDA(cudaFreeHost(ptr));
    return 0;
}
int thsp_pointer_kind(const void* ptr)
{
    cudaPointerAttributes at;
    memset(&at, 0, sizeof(at));
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
        cudaGetLastError();
        return -1;   // the runtime could not say (sticky error, shutting down): NOT "plain host memory"
    }
    switch (at.type) {
        case cudaMemoryTypeDevice: return 1;
        case cudaMemoryTypeManaged: return 2;
        case cudaMemoryTypeHost: return 3;
        default: return 0;
    }
}
int thsp_memcpy_h2d(void* dst, const void* src, size_t bytes, thsp_stream_t s)
{
    THSP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(s)));
    return 0;
}
int thsp_memcpy_d2h(void* dst, const void* src, size_t bytes, thsp_stream_t s)
{
    THSP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(s)));
    return 0;
}
int thsp_memcpy_d2d(void* dst, const void* src, size_t bytes, thsp_stream_t s)
{
    THSP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(s)));
    return 0;
}
int thsp_memset(void* dst, int byte, size_t bytes, thsp_stream_t s)
{
    THSP_CUDA(cudaMemsetAsync(dst, byte, bytes, as_stream(s)));
    return 0;
}
int thsp_prefetch(const void* p, size_t bytes, int to_device, thsp_stream_t s)
{
    int dev = 0;
    THSP_CUDA(cudaGetDevice(&dev));
    THSP_CUDA(cudaMemPrefetchAsync(p, bytes, to_device ? dev : cudaCpuDeviceId, as_stream(s)));
    return 0;
}
int thsp_advise_read_mostly(const void* managed_ptr, size_t bytes, int on)
{
    int dev = 0;
    THSP_CUDA(cudaGetDevice(&dev));
    if (bytes) THSP_CUDA(cudaMemAdvise(managed_ptr, bytes, on ? cudaMemAdviseSetReadMostly : cudaMemAdviseUnsetReadMostly, dev));
    return 0;
}
int thsp_stream_sync(thsp_stream_t s)
{
    THSP_CUDA(cudaStreamSynchronize(as_stream(s)));
    return 0;
}
int thsp_device_sync(void)
{
    THSP_CUDA(cudaDeviceSynchronize());
    return 0;
}
uint64_t thsp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// Scratch grows with the largest call seen (a COO->CSR of 268 M entries keeps 9.7 GB of sort buffers);
// a caller that is done converting can hand it back.
int thsp_scratch_release(void)
{
    int dev = 0;
    THSP_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDev) return 0;
    THSP_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    for (int i = 0; i < kSlots; ++i) {
        Scratch& sc = g_scratch[dev][i];
        if (sc.p) cudaFree(sc.p);
        sc.p = nullptr;
        sc.cap = 0;
        ++sc.uses;   // whoever remembered a pointer into this slot must notice (thsp_coo2ell_prepare)
    }
    return 0;
}

}  // extern "C"
