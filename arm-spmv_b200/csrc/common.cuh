// common.cuh -- shared device/host helpers for libthsparse_cuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "thsp.h"

namespace thsp {

// ---- error plumbing: C ABI returns codes, the message is kept per thread ---------------
void set_error(const char* fmt, ...);
void note_launch(unsigned n = 1);
void forget_launches(unsigned n);   // kernels that were captured into a graph, not launched

#define THSP_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            ::thsp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return 1;                                                                                \
        }                                                                                            \
    } while (0)

#define THSP_LAUNCH_CHECK()                                                                           \
    do {                                                                                             \
        ::thsp::note_launch();                                                                       \
        cudaError_t e_ = cudaGetLastError();                                                         \
        if (e_ != cudaSuccess) {                                                                     \
            ::thsp::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return 1;                                                                                \
        }                                                                                            \
    } while (0)

#define THSP_REQUIRE(cond, msg)                                            \
    do {                                                                   \
        if (!(cond)) {                                                     \
            ::thsp::set_error("%s:%d: %s", __FILE__, __LINE__, msg);       \
            return 2;                                                      \
        }                                                                  \
    } while (0)

static inline cudaStream_t as_stream(thsp_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();            // SMs of the current device (cached per device)
int ensure_device();       // 0 if a CUDA device is usable, else sets error

// Per-device scratch that lives for the process (partials of reductions, scan block sums...).
// Grows monotonically; never shared between streams concurrently by this library's own calls.
void* scratch(size_t bytes, int slot);
uint64_t scratch_uses(int slot);   // how often the slot has been handed out on the current device

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- device helpers -----------------------------------------------------------------
#ifdef __CUDACC__

// Unfused IEEE arithmetic: the reference is g++ -O2 on x86-64 (no FMA contraction), so
// a*b+c there is two roundings.  Using these everywhere keeps elementwise results and the
// in-order CSR/ELL/DIA sums bit-identical to it.
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }

// Streaming loads: read-only path, do not allocate in L1 (keeps L1 for the x gathers).
__device__ __forceinline__ double ld_stream(const double* p)
{
    double v;
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream(const float* p)
{
    float v;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_stream(const int* p)
{
    int v;
    asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_stream2(const double* p)
{
    double2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int2 ld_stream2(const int* p)
{
    int2 v;
    asm("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ld_stream4(const int* p)
{
    int4 v;
    asm("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream4(const float* p)
{
    float4 v;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
// Streaming loads with an L2 evict-first policy (createpolicy): a stream that is read once must
// not push the gathered vector out of L2.
__device__ __forceinline__ double ld_stream_ef(const double* p, uint64_t pol)
{
    double v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_stream_ef(const float* p, uint64_t pol)
{
    float v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ int ld_stream_ef(const int* p, uint64_t pol)
{
    int v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ int4 ld_stream4_ef(const int* p, uint64_t pol)
{
    int4 v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
        : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
        : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 ld_stream4_ef(const float* p, uint64_t pol)
{
    float4 v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
        : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double2 ld_stream2_ef(const double* p, uint64_t pol)
{
    double2 v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
// 256-bit loads (sm_100: LDG.E.256): a lane takes a whole 32-byte sector in one request.
struct int8v { int v[8]; };
struct double4v { double v[4]; };
struct float8v { float v[8]; };
__device__ __forceinline__ int8v ld_stream8_ef(const int* p, uint64_t pol)
{
    int8v r;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float8v ld_stream8_ef(const float* p, uint64_t pol)
{
    float8v r;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
        : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
        : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double4v ld_stream4_ef(const double* p, uint64_t pol)
{
    double4v r;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
        : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
        : "l"(p), "l"(pol));
    return r;
}
// Eight consecutive elements of a stream into registers: one 256-bit load (two for doubles) when the
// address is 32-byte aligned and all eight exist, guarded scalar loads otherwise.  Used by the kernels
// whose lanes own consecutive entries (merge-path CSR, COO).
template <bool kVec>
__device__ __forceinline__ void load_block8(const int* p, int limit, int* c, uint64_t pol)
{
    // limit = entries readable from p; kVec: p is 32-byte aligned
    if (kVec && limit >= 8) {
        const int8v a = ld_stream8_ef(p, pol);
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] = a.v[k];
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] = k < limit ? ld_stream_ef(p + k, pol) : 0;
    }
}
template <bool kVec>
__device__ __forceinline__ void load_block8(const double* p, int limit, double* v, uint64_t pol)
{
    if (kVec && limit >= 8) {
        const double4v a = ld_stream4_ef(p, pol), b = ld_stream4_ef(p + 4, pol);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = a.v[k];
            v[k + 4] = b.v[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = k < limit ? ld_stream_ef(p + k, pol) : 0.0;
    }
}
template <bool kVec>
__device__ __forceinline__ void load_block8(const float* p, int limit, float* v, uint64_t pol)
{
    if (kVec && limit >= 8) {
        const float8v a = ld_stream8_ef(p, pol);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = a.v[k];
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = k < limit ? ld_stream_ef(p + k, pol) : 0.f;
    }
}

// x gathers: read-only path, allocate in L1 (neighbouring rows hit the same lines).
template <typename T>
__device__ __forceinline__ T ld_gather(const T* p) { return __ldg(p); }
// Random gathers with no reuse inside an SM: do not allocate in L1.
__device__ __forceinline__ double ld_gather_na(const double* p) { return ld_stream(p); }
__device__ __forceinline__ float ld_gather_na(const float* p) { return ld_stream(p); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- mbarrier + 1-D bulk async copy (TMA, SASS: UBLKCP) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#endif  // __CUDACC__

}  // namespace thsp
