// tree_sum.cuh -- the canonical sum used for ||y||^2 in the iterated loop (sm_100a).
//
// vec_dot (src/vec_vec.cpp:15-29) leaves the order of its OpenMP reduction open.  The power iteration needs more than
// "some deterministic order": y = A x keeps the reference's per-row order on any number of GPUs, so the whole loop is
// bit-identical for 1, 2, 4 or 8 GPUs exactly when sum y_i^2 is - and that needs an order that does not depend on
// who owns which rows.  The canonical order:
//   * rows are taken in TILES of 32 consecutive rows; a tile's partial is the xor-butterfly (offsets 16, 8, 4, 2, 1)
//     of the 32 squares - what a warp of the CSR stream kernel holds when it finishes a tile, so the SpMV writes the
//     partials in its epilogue and no kernel reads y again;
//   * partials are combined by the binary tree over their INDEX BITS (element i, a multiple of 2^(L+1), takes in
//     element i + 2^L at level L), missing elements counting as +0.0 - adding +0.0 to a sum of squares changes no bit.
// A rank whose first tile index is a multiple of the power of two that covers its tile count computes a subtree of
// the global tree; the per-rank results are combined by the same tree over rank numbers.  Row blocks of
// src/mat_vec.cpp:233 (nrow / G rows each) are aligned like that whenever nrow is a multiple of 32 G and G a power of two.
#pragma once
#include "common.cuh"

namespace thsp {

static constexpr int kTreeThreads = 256;
static constexpr int kTreePerThread = 16;
static constexpr int kTreeBlock = kTreeThreads * kTreePerThread;   // 4096 values per CTA

// xor-butterfly over the warp, unfused adds: every lane ends with the same bits
__device__ __forceinline__ double warp_butterfly_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = add_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Index-bit tree over up to 4096 values at p (cnt of them exist, the rest count as +0.0), by a 256-thread CTA.
// The result is returned to every thread.
__device__ __forceinline__ double block_tree_sum(const double* __restrict__ p, int cnt)
{
    __shared__ double s_w[kTreeThreads / 32];
    __shared__ double s_out;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    double v[kTreePerThread];
    const int base = t * kTreePerThread;
#pragma unroll
    for (int k = 0; k < kTreePerThread; ++k) v[k] = base + k < cnt ? __ldcg(p + base + k) : 0.0;
#pragma unroll
    for (int s = 1; s < kTreePerThread; s <<= 1)
#pragma unroll
        for (int k = 0; k < kTreePerThread; k += 2 * s) v[k] = add_rn(v[k], v[k + s]);
    double r = v[0];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) r = add_rn(r, __shfl_down_sync(0xffffffffu, r, o));   // valid in lanes that are multiples of 2o
    if (lane == 0) s_w[w] = r;
    __syncthreads();
    if (w == 0) {
        double q = lane < kTreeThreads / 32 ? s_w[lane] : 0.0;
#pragma unroll
        for (int o = 1; o < kTreeThreads / 32; o <<= 1) q = add_rn(q, __shfl_down_sync(0xffffffffu, q, o));
        if (lane == 0) s_out = q;
    }
    __syncthreads();
    const double out = s_out;
    __syncthreads();
    return out;
}

// The rest of the tree over m block results, by ONE CTA: rounds of 4096 until one value is left.  a / b are two
// scratch areas of ceil(m / 4096) doubles or more; vals may be a.  Returned to every thread.
__device__ __forceinline__ double block_tree_finish(const double* vals, int m, double* a, double* b)
{
    const double* src = vals;
    double* dst = (vals == a) ? b : a;
    while (true) {
        const int nb = (m + kTreeBlock - 1) / kTreeBlock;
        double last = 0.0;
        for (int c = 0; c < nb; ++c) {
            last = block_tree_sum(src + (size_t)c * kTreeBlock, min(kTreeBlock, m - c * kTreeBlock));
            if (nb > 1 && threadIdx.x == 0) dst[c] = last;
        }
        if (nb == 1) return last;
        __threadfence_block();
        __syncthreads();
        m = nb;
        src = dst;
        dst = (dst == a) ? b : a;
    }
}

}  // namespace thsp
