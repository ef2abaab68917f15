// vec_ops.cu -- vector kernels behind vec_vec.h / vector.h of the reference.
//
// Replaces: vec_dot, vec_axpby (src/vec_vec.cpp:15-94) and Vector::{Fill,Scale,Shift,Copy,
// AddScaled,Add2Scaled}, checkVector (src/vector.cpp:59-171).
// Elementwise kernels use unfused mul/add (common.cuh) and the reference's branch structure,
// so their outputs are bit-identical to the reference.  All are pure streaming kernels
// (HBM-bound): 256-thread CTAs, 4 independent elements in flight per thread, grid sized to a
// multiple of the SM count.
#include <algorithm>

#include "common.cuh"
#include "tree_sum.cuh"

namespace thsp {

static constexpr int kEwThreads = 256;
static constexpr int kEwUnroll = 4;

template <class F>
__global__ void __launch_bounds__(kEwThreads) ew_kernel(int64_t n, F f)
{
    const int64_t stride = (int64_t)gridDim.x * kEwThreads;
    int64_t i = (int64_t)blockIdx.x * kEwThreads + threadIdx.x;
    for (; i + (kEwUnroll - 1) * stride < n; i += kEwUnroll * stride) {
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) f(i + u * stride);
    }
    for (; i < n; i += stride) f(i);
}

static inline int ew_grid(int64_t n)
{
    int64_t want = (n + (int64_t)kEwThreads * kEwUnroll - 1) / ((int64_t)kEwThreads * kEwUnroll);
    int64_t cap = (int64_t)sm_count() * 8;  // 8 x 256 threads = full occupancy
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

template <class F>
static int launch_ew(int64_t n, cudaStream_t s, F f)
{
    if (n <= 0) return 0;
    ew_kernel<<<ew_grid(n), kEwThreads, 0, s>>>(n, f);
    THSP_LAUNCH_CHECK();
    return 0;
}

// ---- reductions: per-thread serial partial -> warp tree -> CTA tree -> fixed-order final pass.
// Summation order (documented for parity): element i goes to thread
// (i mod G*256); each thread adds its elements in ascending i; partials are combined by a
// butterfly over lanes, then over warps, then the G CTA partials are added by CTA 0 of the
// second launch in the same tree shape.  Deterministic for a given n and SM count.
static constexpr int kRedThreads = 256;

template <typename T>
__device__ __forceinline__ T block_sum(T v)
{
    __shared__ T warp_part[kRedThreads / 32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = 0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < kRedThreads / 32 ? warp_part[threadIdx.x] : T(0);
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kRedThreads) dot_partial_kernel(int64_t n, const double* __restrict__ x,
                                                                  const double* __restrict__ y, double* __restrict__ part)
{
    const int64_t stride = (int64_t)gridDim.x * kRedThreads;
    double acc = 0.0;
    int64_t i = (int64_t)blockIdx.x * kRedThreads + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        double a0 = ld_stream(x + i), b0 = ld_stream(y + i);
        double a1 = ld_stream(x + i + stride), b1 = ld_stream(y + i + stride);
        double a2 = ld_stream(x + i + 2 * stride), b2 = ld_stream(y + i + 2 * stride);
        double a3 = ld_stream(x + i + 3 * stride), b3 = ld_stream(y + i + 3 * stride);
        acc = add_rn(acc, mul_rn(a0, b0));
        acc = add_rn(acc, mul_rn(a1, b1));
        acc = add_rn(acc, mul_rn(a2, b2));
        acc = add_rn(acc, mul_rn(a3, b3));
    }
    for (; i < n; i += stride) acc = add_rn(acc, mul_rn(ld_stream(x + i), ld_stream(y + i)));
    double r = block_sum(acc);
    if (threadIdx.x == 0) part[blockIdx.x] = r;
}

__global__ void __launch_bounds__(kRedThreads) sum_final_kernel(int nparts, const double* __restrict__ part,
                                                                double* __restrict__ out)
{
    double acc = 0.0;
    for (int i = threadIdx.x; i < nparts; i += kRedThreads) acc += part[i];
    double r = block_sum(acc);
    if (threadIdx.x == 0) *out = r;
}

static int dot_to_device(int64_t n, const double* x, const double* y, double* out_dev, cudaStream_t s)
{
    int grid = ew_grid(n > 0 ? n : 1);
    double* part = static_cast<double*>(scratch(sizeof(double) * (size_t)grid, 0));
    if (!part) return 1;
    dot_partial_kernel<<<grid, kRedThreads, 0, s>>>(n, x, y, part);
    THSP_LAUNCH_CHECK();
    sum_final_kernel<<<1, kRedThreads, 0, s>>>(grid, part, out_dev);
    THSP_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(kRedThreads) maxdiff_partial_kernel(int64_t n, const double* __restrict__ x,
                                                                      const double* __restrict__ y, int* __restrict__ bad)
{
    const int64_t stride = (int64_t)gridDim.x * kRedThreads;
    int flag = 0;
    for (int64_t i = (int64_t)blockIdx.x * kRedThreads + threadIdx.x; i < n; i += stride)
        if (fabs(x[i] - y[i]) > 1e-6) flag = 1;
    if (__syncthreads_or(flag) && threadIdx.x == 0) atomicOr(bad, 1);
}

// dst_k[offset+i] = src[i] * (1/sqrt(sumsq)) for each peer replica.
struct PeerList {
    double* p[8];
};
__global__ void __launch_bounds__(kEwThreads) scale_broadcast_kernel(int64_t n, const double* __restrict__ src,
                                                                     const double* __restrict__ sumsq, PeerList peers,
                                                                     int npeers, int64_t offset)
{
    const double inv = 1.0 / sqrt(*sumsq);
    const int64_t stride = (int64_t)gridDim.x * kEwThreads;
    for (int64_t i = (int64_t)blockIdx.x * kEwThreads + threadIdx.x; i < n; i += stride) {
        double v = mul_rn(inv, src[i]);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < npeers) peers.p[k][offset + i] = v;
    }
}

// ---- canonical sum of squares (tree_sum.cuh): tile partials, then the index-bit tree ----------------------------
// tile_ss[t] = butterfly sum of y_i^2 over rows [32 t, 32 t + 32); a warp per tile.  The CSR stream kernel writes the
// same numbers from its epilogue (csr_spmv.cu); this kernel serves the other kernels and other callers.
__global__ void __launch_bounds__(256) tile_sumsq_kernel(int64_t n, const double* __restrict__ y, double* __restrict__ tile_ss)
{
    const int lane = threadIdx.x & 31;
    const int64_t ntiles = (n + 31) >> 5;
    const int64_t gw = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5, GW = ((int64_t)gridDim.x * 256) >> 5;
    for (int64_t t = gw; t < ntiles; t += GW) {
        const int64_t i = t * 32 + lane;
        const double v = i < n ? ld_stream(y + i) : 0.0;
        const double q = warp_butterfly_sum(mul_rn(v, v));
        if (lane == 0) tile_ss[t] = q;
    }
}
__global__ void __launch_bounds__(kTreeThreads) tree_blocks_kernel(int64_t m, const double* __restrict__ vals, double* __restrict__ out)
{
    const int64_t b0 = (int64_t)blockIdx.x * kTreeBlock;
    const double r = block_tree_sum(vals + b0, (int)min((int64_t)kTreeBlock, m - b0));
    if (threadIdx.x == 0) out[blockIdx.x] = r;
}
__global__ void __launch_bounds__(kTreeThreads) tree_finish_kernel(int m, double* a, double* b, double* __restrict__ out)
{
    const double r = block_tree_finish(a, m, a, b);
    if (threadIdx.x == 0) *out = r;
}

// diag[i] = sum of the entries (i, i) of a CSR matrix (0 when the row has none), one thread per row.
__global__ void __launch_bounds__(256) csr_diagonal_kernel(int nrow, const int* __restrict__ rp, const int* __restrict__ ci,
                                                           const double* __restrict__ va, double* __restrict__ diag)
{
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= nrow) return;
    double d = 0.0;
    for (int p = rp[r]; p < rp[r + 1]; ++p)
        if (ci[p] == r) d = add_rn(d, va[p]);
    diag[r] = d;
}

}  // namespace thsp

using namespace thsp;

extern "C" {

int thsp_csr_diagonal_f64(int nrow, const int* row_ptr, const int* col_ind, const double* val, double* diag, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    if (nrow <= 0) return 0;
    csr_diagonal_kernel<<<div_up(nrow, 256), 256, 0, as_stream(stream)>>>(nrow, row_ptr, col_ind, val, diag);
    THSP_LAUNCH_CHECK();
    return 0;
}

// x[i] += omega * r[i] / diag[i]: the update of a (damped) Jacobi sweep, r = b - A x
int thsp_jacobi_update_f64(int64_t n, double omega, const double* diag, const double* r, double* x, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    return launch_ew(n, as_stream(stream), [=] __device__(int64_t i) { x[i] = add_rn(x[i], mul_rn(omega, __ddiv_rn(r[i], diag[i]))); });
}

int thsp_dot_dev_f64(int64_t n, const double* x, const double* y, double* result_dev, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    return dot_to_device(n, x, y, result_dev, as_stream(stream));
}

int thsp_dot_f64(int64_t n, const double* x, const double* y, double* result_host, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    double* out = static_cast<double*>(scratch(sizeof(double), 1));
    if (!out) return 1;
    if (dot_to_device(n, x, y, out, as_stream(stream))) return 1;
    THSP_CUDA(cudaMemcpyAsync(result_host, out, sizeof(double), cudaMemcpyDeviceToHost, as_stream(stream)));
    THSP_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return 0;
}

int thsp_sumsq_dev_f64(int64_t n, const double* y, double* out_dev, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    return dot_to_device(n, y, y, out_dev, as_stream(stream));
}

int thsp_tile_sumsq_f64(int64_t n, const double* y, double* tile_ss, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    if (n <= 0) return 0;
    const int64_t ntiles = (n + 31) / 32;
    const int grid = (int)std::min<int64_t>((int64_t)sm_count() * 8, (ntiles + 7) / 8);
    tile_sumsq_kernel<<<grid, 256, 0, as_stream(stream)>>>(n, y, tile_ss);
    THSP_LAUNCH_CHECK();
    return 0;
}

int thsp_tree_sum_f64(int64_t m, const double* vals, double* out_dev, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    cudaStream_t s = as_stream(stream);
    if (m <= 0) {
        THSP_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double), s));
        return 0;
    }
    const int64_t nb = (m + kTreeBlock - 1) / kTreeBlock;
    THSP_REQUIRE(nb <= (int64_t)1 << 30, "too many values");
    if (nb == 1) {
        tree_blocks_kernel<<<1, kTreeThreads, 0, s>>>(m, vals, out_dev);
        THSP_LAUNCH_CHECK();
        return 0;
    }
    const int64_t nb2 = (nb + kTreeBlock - 1) / kTreeBlock;
    double* a = static_cast<double*>(scratch(sizeof(double) * (size_t)(nb + nb2 + 2), 0));
    if (!a) return 1;
    tree_blocks_kernel<<<(int)nb, kTreeThreads, 0, s>>>(m, vals, a);
    THSP_LAUNCH_CHECK();
    tree_finish_kernel<<<1, kTreeThreads, 0, s>>>((int)nb, a, a + nb, out_dev);
    THSP_LAUNCH_CHECK();
    return 0;
}

// out[i] = x[i] * *scale (what the stream kernel's scaled-x variant does on the fly, as a pass: for plans that run
// another kernel)
int thsp_scale_by_dev_f64(int64_t n, const double* x, const double* scale_dev, double* out, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    return launch_ew(n, as_stream(stream), [=] __device__(int64_t i) { out[i] = mul_rn(x[i], __ldg(scale_dev)); });
}
// *inv = 1 / sqrt(*sumsq): the factor of vec_axpby(1/sqrt(s), y, 0, y), left on the device for the next product
int thsp_inv_sqrt_dev_f64(const double* sumsq_dev, double* inv_dev, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    return launch_ew(1, as_stream(stream), [=] __device__(int64_t) { *inv_dev = __ddiv_rn(1.0, __dsqrt_rn(*sumsq_dev)); });
}
int thsp_axpby_f64(int64_t n, double alpha, const double* x, double beta, const double* y, double* w, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    cudaStream_t s = as_stream(stream);
    // Branch order and per-branch expression follow src/vec_vec.cpp:38-93.
    if (alpha == 0) return launch_ew(n, s, [=] __device__(int64_t i) { w[i] = mul_rn(beta, y[i]); });
    if (beta == 0) return launch_ew(n, s, [=] __device__(int64_t i) { w[i] = mul_rn(alpha, x[i]); });
    if (alpha == 1) return launch_ew(n, s, [=] __device__(int64_t i) { w[i] = add_rn(mul_rn(beta, y[i]), x[i]); });
    if (alpha == -1) return launch_ew(n, s, [=] __device__(int64_t i) { w[i] = add_rn(mul_rn(beta, y[i]), -x[i]); });
    if (beta == 1) return launch_ew(n, s, [=] __device__(int64_t i) { w[i] = add_rn(mul_rn(alpha, x[i]), y[i]); });
    if (beta == -1) return launch_ew(n, s, [=] __device__(int64_t i) { w[i] = add_rn(mul_rn(alpha, x[i]), -y[i]); });
    return launch_ew(n, s, [=] __device__(int64_t i) { w[i] = add_rn(mul_rn(alpha, x[i]), mul_rn(beta, y[i])); });
}

int thsp_fill_f64(int64_t n, double a, double* v, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    return launch_ew(n, as_stream(stream), [=] __device__(int64_t i) { v[i] = a; });
}
int thsp_scale_f64(int64_t n, double a, double* v, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    return launch_ew(n, as_stream(stream), [=] __device__(int64_t i) { v[i] = mul_rn(v[i], a); });
}
int thsp_shift_f64(int64_t n, double a, double* v, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    return launch_ew(n, as_stream(stream), [=] __device__(int64_t i) { v[i] = add_rn(v[i], a); });
}
int thsp_copy_f64(int64_t n, const double* x, double* v, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    return launch_ew(n, as_stream(stream), [=] __device__(int64_t i) { v[i] = x[i]; });
}
int thsp_add_scaled_f64(int64_t n, double a, const double* x, double* v, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    cudaStream_t s = as_stream(stream);
    // src/vector.cpp:98-128
    if (a == 0) return 0;
    if (a == 1) return launch_ew(n, s, [=] __device__(int64_t i) { v[i] = add_rn(v[i], x[i]); });
    if (a == -1) return launch_ew(n, s, [=] __device__(int64_t i) { v[i] = add_rn(v[i], -x[i]); });
    return launch_ew(n, s, [=] __device__(int64_t i) { v[i] = add_rn(v[i], mul_rn(a, x[i])); });
}
int thsp_add2_scaled_f64(int64_t n, double a, const double* x, double b, const double* y, double* v, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    cudaStream_t s = as_stream(stream);
    // src/vector.cpp:130-159: the right-hand side (a*x + b*y) is formed first, then added to v.
    if (a == 0) return thsp_add_scaled_f64(n, b, y, v, stream);
    if (b == 0) return thsp_add_scaled_f64(n, a, x, v, stream);
    if (a == 1)
        return launch_ew(n, s, [=] __device__(int64_t i) { v[i] = add_rn(v[i], add_rn(x[i], mul_rn(b, y[i]))); });
    if (b == 1)
        return launch_ew(n, s, [=] __device__(int64_t i) { v[i] = add_rn(v[i], add_rn(mul_rn(a, x[i]), y[i])); });
    return launch_ew(n, s, [=] __device__(int64_t i) { v[i] = add_rn(v[i], add_rn(mul_rn(a, x[i]), mul_rn(b, y[i]))); });
}

int thsp_check_vector_f64(int64_t nx, const double* x, int64_t ny, const double* y, int* ok_host, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    if (nx != ny) {
        *ok_host = 0;
        return 0;
    }
    cudaStream_t s = as_stream(stream);
    int* bad = static_cast<int*>(scratch(sizeof(int), 1));
    if (!bad) return 1;
    THSP_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
    if (nx > 0) {
        maxdiff_partial_kernel<<<ew_grid(nx), kRedThreads, 0, s>>>(nx, x, y, bad);
        THSP_LAUNCH_CHECK();
    }
    int h = 0;
    THSP_CUDA(cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    THSP_CUDA(cudaStreamSynchronize(s));
    *ok_host = h ? 0 : 1;
    return 0;
}

int thsp_scale_broadcast_f64(int64_t n, const double* src, const double* sumsq_dev, double* const* peer_dst, int npeers,
                             int64_t offset, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    THSP_REQUIRE(npeers >= 1 && npeers <= 8, "npeers must be in 1..8");
    PeerList pl;
    for (int k = 0; k < 8; ++k) pl.p[k] = k < npeers ? peer_dst[k] : nullptr;
    if (n <= 0) return 0;
    scale_broadcast_kernel<<<ew_grid(n), kEwThreads, 0, as_stream(stream)>>>(n, src, sumsq_dev, pl, npeers, offset);
    THSP_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
