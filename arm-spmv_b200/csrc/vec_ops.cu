// vec_ops.cu -- vector kernels behind vec_vec.h / vector.h of the reference.
//
// Replaces: vec_dot, vec_axpby (src/vec_vec.cpp:15-94) and Vector::{Fill,Scale,Shift,Copy,
// AddScaled,Add2Scaled}, checkVector (src/vector.cpp:59-171).
// Elementwise kernels use unfused mul/add (common.cuh) and the reference's branch structure,
// so their outputs are bit-identical to the reference.  All are pure streaming kernels
// (HBM-bound): 256-thread CTAs, four consecutive elements per thread moved with 16-byte accesses (map_kernel).
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

// ---- elementwise maps out[i] = g(a[i], b[i], c[i]) with 128-bit accesses -------------------------------------------------
// A thread owns FOUR consecutive elements and moves them as one 32-byte access per array (two 16-byte ones when the arrays
// are only 16-byte aligned); a CTA owns 1024 consecutive elements.  The first form of these kernels (ew_kernel
// above: one element per access, four accesses 2.4 MB apart per thread) reached the HBM roofline only with two input
// streams - 6.6 TB/s for w = a x + b y, but 4.0 TB/s for w = a x and 4.9 TB/s for the in-place v += a x on 134 M
// doubles (scripts/vec_bench.py, profiles/r02_vec_ops.txt); with 16-byte accesses all of them run at 6.5-6.9 TB/s.
// out may alias an input (in-place updates): a thread reads its four elements before it writes them.
static constexpr int kMapThreads = 256;
static constexpr int kMapPer = 4;

// four consecutive doubles as ONE 32-byte access (sm_100: LDG.E.256 / STG.E.256 - a whole sector per thread, so no store ever
// covers half a sector), as two 16-byte accesses, or one by one
__device__ __forceinline__ void load4(const double* p, int vec, double (&v)[4])
{
    if (vec == 2) {
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
    } else {
        const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
}
__device__ __forceinline__ void store4(double* p, int vec, const double (&v)[4])
{
    if (vec == 2) {
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
    } else {
        *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
    }
}

template <int NIN, class G>
__global__ void __launch_bounds__(kMapThreads) map_kernel(int64_t n, double* out, const double* a, const double* b, const double* c, G g,
                                                          int vec)
{
    const int64_t stride = (int64_t)gridDim.x * kMapThreads * kMapPer;
    for (int64_t base = ((int64_t)blockIdx.x * kMapThreads + threadIdx.x) * kMapPer; base < n; base += stride) {
        double va[kMapPer] = {0, 0, 0, 0}, vb[kMapPer] = {0, 0, 0, 0}, vc[kMapPer] = {0, 0, 0, 0}, r[kMapPer];
        if (vec && base + kMapPer <= n) {
            if (NIN >= 1) load4(a + base, vec, va);
            if (NIN >= 2) load4(b + base, vec, vb);
            if (NIN >= 3) load4(c + base, vec, vc);
#pragma unroll
            for (int k = 0; k < kMapPer; ++k) r[k] = g(va[k], vb[k], vc[k]);
            store4(out + base, vec, r);
        } else {
#pragma unroll
            for (int k = 0; k < kMapPer; ++k)
                if (base + k < n) {
                    const double xa = NIN >= 1 ? a[base + k] : 0.0, xb = NIN >= 2 ? b[base + k] : 0.0, xc = NIN >= 3 ? c[base + k] : 0.0;
                    out[base + k] = g(xa, xb, xc);
                }
        }
    }
}

template <int NIN, class G>
static int launch_map(int64_t n, cudaStream_t s, double* out, const double* a, const double* b, const double* c, G g)
{
    if (n <= 0) return 0;
    const uintptr_t bits = (uintptr_t)out | (NIN >= 1 ? (uintptr_t)a : 0) | (NIN >= 2 ? (uintptr_t)b : 0) | (NIN >= 3 ? (uintptr_t)c : 0);
    const int64_t ctas = (n + (int64_t)kMapThreads * kMapPer - 1) / ((int64_t)kMapThreads * kMapPer);
    const int grid = (int)std::min<int64_t>(ctas, (int64_t)sm_count() * 2048);   // one pass for anything below 300 M elements
    map_kernel<NIN, G><<<grid, kMapThreads, 0, s>>>(n, out, a, b, c, g, (bits & 31) == 0 ? 2 : ((bits & 15) == 0 ? 1 : 0));
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
    constexpr int U = 4;   // four tiles per trip: four loads in flight per lane
    for (int64_t t0 = gw; t0 < ntiles; t0 += U * GW) {
        double v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = (t0 + u * GW) * 32 + lane;
            v[u] = (t0 + u * GW < ntiles && i < n) ? ld_stream(y + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double q = warp_butterfly_sum(mul_rn(v[u], v[u]));
            if (lane == 0 && t0 + u * GW < ntiles) tile_ss[t0 + u * GW] = q;
        }
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
    return launch_map<3>(n, as_stream(stream), x, x, r, diag,
                         [=] __device__(double xi, double ri, double di) { return add_rn(xi, mul_rn(omega, __ddiv_rn(ri, di))); });
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
    return launch_map<1>(n, as_stream(stream), out, x, nullptr, nullptr, [=] __device__(double xi, double, double) { return mul_rn(xi, __ldg(scale_dev)); });
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
    if (alpha == 0) return launch_map<1>(n, s, w, y, nullptr, nullptr, [=] __device__(double yi, double, double) { return mul_rn(beta, yi); });
    if (beta == 0) return launch_map<1>(n, s, w, x, nullptr, nullptr, [=] __device__(double xi, double, double) { return mul_rn(alpha, xi); });
    if (alpha == 1) return launch_map<2>(n, s, w, x, y, nullptr, [=] __device__(double xi, double yi, double) { return add_rn(mul_rn(beta, yi), xi); });
    if (alpha == -1) return launch_map<2>(n, s, w, x, y, nullptr, [=] __device__(double xi, double yi, double) { return add_rn(mul_rn(beta, yi), -xi); });
    if (beta == 1) return launch_map<2>(n, s, w, x, y, nullptr, [=] __device__(double xi, double yi, double) { return add_rn(mul_rn(alpha, xi), yi); });
    if (beta == -1) return launch_map<2>(n, s, w, x, y, nullptr, [=] __device__(double xi, double yi, double) { return add_rn(mul_rn(alpha, xi), -yi); });
    return launch_map<2>(n, s, w, x, y, nullptr, [=] __device__(double xi, double yi, double) { return add_rn(mul_rn(alpha, xi), mul_rn(beta, yi)); });
}

int thsp_fill_f64(int64_t n, double a, double* v, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    return launch_map<0>(n, as_stream(stream), v, nullptr, nullptr, nullptr, [=] __device__(double, double, double) { return a; });
}
int thsp_scale_f64(int64_t n, double a, double* v, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    return launch_map<1>(n, as_stream(stream), v, v, nullptr, nullptr, [=] __device__(double vi, double, double) { return mul_rn(vi, a); });
}
int thsp_shift_f64(int64_t n, double a, double* v, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    return launch_map<1>(n, as_stream(stream), v, v, nullptr, nullptr, [=] __device__(double vi, double, double) { return add_rn(vi, a); });
}
int thsp_copy_f64(int64_t n, const double* x, double* v, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    return launch_map<1>(n, as_stream(stream), v, x, nullptr, nullptr, [=] __device__(double xi, double, double) { return xi; });
}
int thsp_add_scaled_f64(int64_t n, double a, const double* x, double* v, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    cudaStream_t s = as_stream(stream);
    // src/vector.cpp:98-128
    if (a == 0) return 0;
    if (a == 1) return launch_map<2>(n, s, v, v, x, nullptr, [=] __device__(double vi, double xi, double) { return add_rn(vi, xi); });
    if (a == -1) return launch_map<2>(n, s, v, v, x, nullptr, [=] __device__(double vi, double xi, double) { return add_rn(vi, -xi); });
    return launch_map<2>(n, s, v, v, x, nullptr, [=] __device__(double vi, double xi, double) { return add_rn(vi, mul_rn(a, xi)); });
}
int thsp_add2_scaled_f64(int64_t n, double a, const double* x, double b, const double* y, double* v, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    cudaStream_t s = as_stream(stream);
    // src/vector.cpp:130-159: the right-hand side (a*x + b*y) is formed first, then added to v.
    if (a == 0) return thsp_add_scaled_f64(n, b, y, v, stream);
    if (b == 0) return thsp_add_scaled_f64(n, a, x, v, stream);
    if (a == 1)
        return launch_map<3>(n, s, v, v, x, y, [=] __device__(double vi, double xi, double yi) { return add_rn(vi, add_rn(xi, mul_rn(b, yi))); });
    if (b == 1)
        return launch_map<3>(n, s, v, v, x, y, [=] __device__(double vi, double xi, double yi) { return add_rn(vi, add_rn(mul_rn(a, xi), yi)); });
    return launch_map<3>(n, s, v, v, x, y, [=] __device__(double vi, double xi, double yi) { return add_rn(vi, add_rn(mul_rn(a, xi), mul_rn(b, yi))); });
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
