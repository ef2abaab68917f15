// convert.cu -- format conversions on the GPU, index arrays bit-identical to the reference.
//
// Replaces the converting constructors CSRMatrix(const COOMatrix&) (src/matrix.cpp:115-154),
// CSCMatrix(const COOMatrix&) (:295-325), ELLMatrix(const COOMatrix&) (:450-500) and
// DIAMatrix(const CSRMatrix&) (:673-726).
//
// The reference's "histogram, running sum, backward fill with pre-decrement" is a STABLE
// counting sort of the entries by row (column for CSC): inside a bucket entries keep their COO
// order and duplicates survive.  On the GPU:
//   1. one pass over the keys builds the bucket histogram (one atomic per run of equal
//      adjacent keys) and notes whether the keys are already non-decreasing;
//   2. a three-phase exclusive scan of the histogram is row_ptr / col_ptr;
//   3. already sorted  -> the permutation is the identity (stencil generators, sorted .mtx);
//      otherwise       -> stable LSD radix sort of (key, original index), 8 bits per pass,
//                         ceil(log2(nbuckets)/8) passes, ranks inside a CTA from
//                         __match_any_sync so equal digits keep their order;
//   4. one gather writes the payload (ELL: slot = position - row_ptr[row], column-major).
// The packed `diagonal` (row==col entries in COO order) is a stable stream compaction.
// Everything here is integer/byte work bound by HBM traffic.
#include <algorithm>

#include "common.cuh"

namespace thsp {

// =========================================================== exclusive scan (int32) =======
static constexpr int kScanThreads = 256;
static constexpr int kScanItems = 4;
static constexpr int kScanTile = kScanThreads * kScanItems;  // 1024

__device__ __forceinline__ int block_exclusive_scan(int v, int* total)
{
    // exclusive scan of one int per thread across a 256-thread CTA
    __shared__ int warp_tot[kScanThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0) {
        int t = lane < kScanThreads / 32 ? warp_tot[lane] : 0;
        int ti = t;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, ti, d);
            if (lane >= d) ti += u;
        }
        if (lane < kScanThreads / 32) warp_tot[lane] = ti - t;  // exclusive warp offsets
        if (lane == kScanThreads / 32 - 1) *total = ti;
    }
    __syncthreads();
    int r = warp_tot[w] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(int n, const int* __restrict__ in, int* __restrict__ bsum)
{
    __shared__ int tot;
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) s += in[base + i];
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

// out[i] = boff[block] + exclusive prefix inside the tile; in may alias out.
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(int n, const int* in, int* out, const int* __restrict__ boff,
                                                                  int nblocks)
{
    __shared__ int tot;
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = base + i < n ? in[base + i] : 0;
        s += v[i];
    }
    int pre = block_exclusive_scan(s, &tot) + (boff ? boff[blockIdx.x] : 0);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = pre;
        pre += v[i];
    }
    if (blockIdx.x == nblocks - 1 && threadIdx.x == 0) out[n] = boff ? boff[nblocks] : tot;
}

static int scan_rec(int n, const int* in, int* out, int* tmp, cudaStream_t s)
{
    const int nb = div_up(n, kScanTile);
    if (nb <= 1) {
        scan_apply_kernel<<<1, kScanThreads, 0, s>>>(n, in, out, nullptr, 1);
        THSP_LAUNCH_CHECK();
        return 0;
    }
    int* bsum = tmp;
    int* boff = tmp + nb;
    scan_reduce_kernel<<<nb, kScanThreads, 0, s>>>(n, in, bsum);
    THSP_LAUNCH_CHECK();
    if (scan_rec(nb, bsum, boff, tmp + 2 * nb + 1, s)) return 1;
    scan_apply_kernel<<<nb, kScanThreads, 0, s>>>(n, in, out, boff, nb);
    THSP_LAUNCH_CHECK();
    return 0;
}

// out has n+1 entries; in may alias out.  scratch slot 5.
int exclusive_scan(int n, const int* in, int* out, cudaStream_t s)
{
    if (n <= 0) {
        THSP_CUDA(cudaMemsetAsync(out, 0, sizeof(int), s));
        return 0;
    }
    size_t need = 0;
    for (int m = n; m > kScanTile;) {
        int nb = div_up(m, kScanTile);
        need += 2 * (size_t)nb + 1;
        m = nb;
    }
    int* tmp = static_cast<int*>(scratch((need + 4) * sizeof(int), 5));
    if (!tmp) return 1;
    return scan_rec(n, in, out, tmp, s);
}

// ============================================================ histogram + sortedness =======
__global__ void __launch_bounds__(256) hist_kernel(int n, const int* __restrict__ key, int* __restrict__ cnt,
                                                   int* __restrict__ unsorted)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int stride = gridDim.x * 256;
    const int rounds = (n + stride - 1) / stride;
    int bad = 0;
    for (int it = 0; it < rounds; ++it) {
        const int k = it * stride + blockIdx.x * 256 + threadIdx.x;
        const int mine = k < n ? ld_stream(key + k) : -1;
        const int up = __shfl_up_sync(full, mine, 1);  // every lane takes part: never inside a short-circuit
        const int prev = lane == 0 ? ((k > 0 && k < n) ? ld_stream(key + k - 1) : -1) : up;
        if (k < n && k > 0 && prev > mine) bad = 1;
        const bool head = (lane == 0) || (up != mine);
        const unsigned heads = __ballot_sync(full, head);
        if (head && mine >= 0) {
            const unsigned above = lane == 31 ? 0u : (heads >> (lane + 1)) << (lane + 1);
            const int end = above ? __ffs(above) - 1 : 32;
            atomicAdd(cnt + mine, end - lane);
        }
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(unsorted, 1);
}

// ================================================================== LSD radix sort ========
// One pass = digit histogram per CTA tile -> exclusive scan over (digit-major, tile-minor) counts
// -> stable scatter.  A tile is 4096 consecutive elements; warp w owns elements
// [w*512, (w+1)*512) of it and walks them in 16 rounds of 32 (coalesced loads).
// Stable rank of an element = (elements with the same digit earlier in the tile): inside a warp
// it comes from __match_any_sync against a warp-private running counter in shared memory (no CTA
// barrier per round), across warps from one prefix over the eight warp counters per digit.
// Elements are then placed in shared memory in sorted order and written out so that consecutive
// threads write consecutive addresses of a digit run.
static constexpr int kRadixThreads = 256;
static constexpr int kRadixWarps = kRadixThreads / 32;
static constexpr int kRadixRounds = 16;
static constexpr int kRadixTile = kRadixThreads * kRadixRounds;  // 4096 keys per CTA

__global__ void __launch_bounds__(kRadixThreads) radix_hist_kernel(int n, const int* __restrict__ key, int shift,
                                                                   int* __restrict__ counts, int nblk)
{
    __shared__ int h[256];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    h[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * kRadixTile;
#pragma unroll 4
    for (int r = 0; r < kRadixRounds; ++r) {
        const int k = base + r * kRadixThreads + threadIdx.x;
        const int d = k < n ? ((ld_stream(key + k) >> shift) & 255) : 256 + lane;
        const unsigned peers = __match_any_sync(full, d);          // one atomic per distinct digit in the warp
        if (k < n && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&h[d], __popc(peers));
    }
    __syncthreads();
    counts[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// idx_in == nullptr means the identity (first pass).
__global__ void __launch_bounds__(kRadixThreads) radix_scatter_kernel(int n, const int* __restrict__ key_in,
                                                                      const int* __restrict__ idx_in, int shift,
                                                                      const int* __restrict__ offsets, int nblk,
                                                                      int* __restrict__ key_out, int* __restrict__ idx_out)
{
    __shared__ int wcnt[kRadixWarps][256];   // per-warp digit counters, later exclusive warp offsets
    __shared__ int tile_off[256];            // first position of each digit inside the sorted tile
    __shared__ int gbase[256];               // where this tile's run of each digit starts in the output
    __shared__ int s_key[kRadixTile];
    __shared__ int s_idx[kRadixTile];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tile = blockIdx.x * kRadixTile;
    const int tile_n = min(kRadixTile, n - tile);
#pragma unroll
    for (int i = 0; i < kRadixWarps; ++i) wcnt[i][threadIdx.x] = 0;
    gbase[threadIdx.x] = offsets[threadIdx.x * nblk + blockIdx.x];
    __syncthreads();

    int kv[kRadixRounds], iv[kRadixRounds], rk[kRadixRounds];  // key, index, rank inside the warp's segment
#pragma unroll
    for (int r = 0; r < kRadixRounds; ++r) {
        const int e = w * (32 * kRadixRounds) + r * 32 + lane;   // position inside the tile
        const bool ok = e < tile_n;
        kv[r] = ok ? key_in[tile + e] : 0;
        iv[r] = ok ? (idx_in ? idx_in[tile + e] : tile + e) : 0;
    }
#pragma unroll
    for (int r = 0; r < kRadixRounds; ++r) {
        const int e = w * (32 * kRadixRounds) + r * 32 + lane;
        const bool ok = e < tile_n;
        const int d = ok ? ((kv[r] >> shift) & 255) : 256 + lane;  // invalid lanes match nobody
        const unsigned peers = __match_any_sync(full, d);
        const int before = __popc(peers & ((1u << lane) - 1u));
        int base = 0;
        if (ok) base = wcnt[w][d];
        __syncwarp();
        if (ok && before == 0) wcnt[w][d] = base + __popc(peers);
        __syncwarp();
        rk[r] = base + before;
    }
    __syncthreads();
    {   // thread d: exclusive prefix of digit d over the warps, and the digit's total in the tile
        const int d = threadIdx.x;
        int run = 0;
#pragma unroll
        for (int i = 0; i < kRadixWarps; ++i) {
            const int c = wcnt[i][d];
            wcnt[i][d] = run;
            run += c;
        }
        int tot;
        const int excl = block_exclusive_scan(run, &tot);   // exclusive scan over digits
        tile_off[d] = excl;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRadixRounds; ++r) {
        const int e = w * (32 * kRadixRounds) + r * 32 + lane;
        if (e < tile_n) {
            const int d = (kv[r] >> shift) & 255;
            const int pos = tile_off[d] + wcnt[w][d] + rk[r];
            s_key[pos] = kv[r];
            s_idx[pos] = iv[r];
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < tile_n; t += kRadixThreads) {
        const int k = s_key[t];
        const int d = (k >> shift) & 255;
        const int out = gbase[d] + (t - tile_off[d]);
        key_out[out] = k;
        idx_out[out] = s_idx[t];
    }
}

// Sort (key, index) stably by key in [0, nbuckets).  Returns device pointers to the sorted keys
// and the permutation (scratch slots 6/7, valid until the next conversion call on this device).
static int stable_sort_by_key(int n, int nbuckets, const int* key, const int** sorted_key, const int** perm, cudaStream_t s)
{
    int bits = 1;
    while (bits < 31 && (1 << bits) < nbuckets) ++bits;
    const int passes = (bits + 7) / 8;
    const int nblk = div_up(n, kRadixTile);
    int* buf = static_cast<int*>(scratch(sizeof(int) * 4 * (size_t)n, 6));
    int* counts = static_cast<int*>(scratch(sizeof(int) * (256 * (size_t)nblk + 1), 7));
    if (!buf || !counts) return 1;
    int* kbuf[2] = {buf, buf + (size_t)n};
    int* ibuf[2] = {buf + 2 * (size_t)n, buf + 3 * (size_t)n};
    const int* kin = key;
    const int* iin = nullptr;
    for (int p = 0; p < passes; ++p) {
        radix_hist_kernel<<<nblk, kRadixThreads, 0, s>>>(n, kin, 8 * p, counts, nblk);
        THSP_LAUNCH_CHECK();
        if (exclusive_scan(256 * nblk, counts, counts, s)) return 1;
        radix_scatter_kernel<<<nblk, kRadixThreads, 0, s>>>(n, kin, iin, 8 * p, counts, nblk, kbuf[p & 1], ibuf[p & 1]);
        THSP_LAUNCH_CHECK();
        kin = kbuf[p & 1];
        iin = ibuf[p & 1];
    }
    *sorted_key = kin;
    *perm = iin;
    return 0;
}

// =============================================================== payload placement ========
// perm == nullptr: identity.
__global__ void __launch_bounds__(256) gather_kernel(int n, const int* __restrict__ perm, const int* __restrict__ other,
                                                     const double* __restrict__ val, int* __restrict__ out_other,
                                                     double* __restrict__ out_val)
{
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= n) return;
    const int k = perm ? perm[p] : p;
    out_other[p] = other[k];
    out_val[p] = val[k];
}

__global__ void __launch_bounds__(256) ell_place_kernel(int n, int nrow, const int* __restrict__ sorted_row,
                                                        const int* __restrict__ perm, const int* __restrict__ row_ptr,
                                                        const int* __restrict__ col, const double* __restrict__ val,
                                                        int* __restrict__ out_col, double* __restrict__ out_val)
{
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= n) return;
    const int k = perm ? perm[p] : p;
    const int r = sorted_row[p];
    const size_t at = (size_t)(p - row_ptr[r]) * nrow + r;
    out_col[at] = col[k];
    out_val[at] = val[k];
}

__global__ void __launch_bounds__(256) max_len_kernel(int nrow, const int* __restrict__ cnt, int* __restrict__ out)
{
    int m = 0;
    for (int r = blockIdx.x * 256 + threadIdx.x; r < nrow; r += gridDim.x * 256) m = max(m, cnt[r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// ---- packed diagonal: stable compaction of entries with row == col --------------------------
__global__ void __launch_bounds__(kScanThreads) diag_count_kernel(int n, const int* __restrict__ ri, const int* __restrict__ ci,
                                                                  int* __restrict__ bcnt)
{
    __shared__ int tot;
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) s += (ri[base + i] == ci[base + i]);
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) bcnt[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(kScanThreads) diag_scatter_kernel(int n, const int* __restrict__ ri, const int* __restrict__ ci,
                                                                    const double* __restrict__ val, const int* __restrict__ boff,
                                                                    int cap, double* __restrict__ diag)
{
    __shared__ int tot;
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    bool f[kScanItems];
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        f[i] = base + i < n && ri[base + i] == ci[base + i];
        s += f[i];
    }
    int pos = block_exclusive_scan(s, &tot) + boff[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (f[i]) {
            if (pos < cap) diag[pos] = val[base + i];
            ++pos;
        }
}

static int pack_diagonal(int nnz, const int* ri, const int* ci, const double* val, int cap, double* diag, int* ndiag_host,
                         cudaStream_t s)
{
    if (nnz <= 0) {
        if (ndiag_host) *ndiag_host = 0;
        return 0;
    }
    const int nb = div_up(nnz, kScanTile);
    int* bcnt = static_cast<int*>(scratch(sizeof(int) * (2 * (size_t)nb + 2), 4));
    if (!bcnt) return 1;
    int* boff = bcnt + nb;
    diag_count_kernel<<<nb, kScanThreads, 0, s>>>(nnz, ri, ci, bcnt);
    THSP_LAUNCH_CHECK();
    if (exclusive_scan(nb, bcnt, boff, s)) return 1;
    if (diag) {
        diag_scatter_kernel<<<nb, kScanThreads, 0, s>>>(nnz, ri, ci, val, boff, cap, diag);
        THSP_LAUNCH_CHECK();
    }
    if (ndiag_host) {
        THSP_CUDA(cudaMemcpyAsync(ndiag_host, boff + nb, sizeof(int), cudaMemcpyDeviceToHost, s));
        THSP_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

// Common front half of COO->CSR/CSC/ELL: ptr = exclusive scan of the key histogram, plus the
// stable order.  *sorted_key/*perm are nullptr-permutation (identity) when already sorted.
static int bucket_order(int nbuckets, int nnz, const int* key, int* ptr, const int** sorted_key, const int** perm,
                        cudaStream_t s)
{
    THSP_CUDA(cudaMemsetAsync(ptr, 0, sizeof(int) * ((size_t)nbuckets + 1), s));
    *sorted_key = key;
    *perm = nullptr;
    if (nnz <= 0) return 0;
    int* flag = static_cast<int*>(scratch(sizeof(int), 1));
    if (!flag) return 1;
    THSP_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
    hist_kernel<<<std::min(div_up(nnz, 256), sm_count() * 16), 256, 0, s>>>(nnz, key, ptr, flag);
    THSP_LAUNCH_CHECK();
    if (exclusive_scan(nbuckets, ptr, ptr, s)) return 1;
    int unsorted = 0;
    THSP_CUDA(cudaMemcpyAsync(&unsorted, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    THSP_CUDA(cudaStreamSynchronize(s));
    if (unsorted) return stable_sort_by_key(nnz, nbuckets, key, sorted_key, perm, s);
    return 0;
}

// ============================================================================ DIA ==========
__global__ void __launch_bounds__(256) dia_mark_kernel(int nrow, int span, const int* __restrict__ rp, const int* __restrict__ ci,
                                                       int* __restrict__ seen)
{
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= nrow) return;
    for (int p = rp[r]; p < rp[r + 1]; ++p) {
        const int m = nrow - r + ci[p];
        if (m < span) seen[m] = 1;  // m == span is the corner diagonal the reference drops
    }
}
__global__ void __launch_bounds__(256) dia_offsets_kernel(int span, int nrow, const int* __restrict__ seen,
                                                          const int* __restrict__ pos, int cap, int* __restrict__ offsets)
{
    const int m = blockIdx.x * 256 + threadIdx.x;
    if (m < span && seen[m] && pos[m] < cap) offsets[pos[m]] = m - nrow;
}
__global__ void __launch_bounds__(256) dia_slot_kernel(int ndiags, int nrow, const int* __restrict__ offsets, int* __restrict__ slot)
{
    const int d = blockIdx.x * 256 + threadIdx.x;
    if (d < ndiags) slot[offsets[d] + nrow] = d;
}
__global__ void __launch_bounds__(256) dia_fill_kernel(int nrow, int span, int ndiags, const int* __restrict__ rp,
                                                       const int* __restrict__ ci, const double* __restrict__ val,
                                                       const int* __restrict__ slot, double* __restrict__ values)
{
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= nrow) return;
    // one thread per row, entries in stored order: a duplicate (i,j) overwrites the earlier one
    for (int p = rp[r]; p < rp[r + 1]; ++p) {
        const int m = nrow - r + ci[p];
        if (m < span && slot[m] >= 0) values[(size_t)r * ndiags + slot[m]] = val[p];
    }
}

}  // namespace thsp

using namespace thsp;

extern "C" {

int thsp_exclusive_scan_i32(int n, const int* counts, int* out, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    return exclusive_scan(n, counts, out, as_stream(stream));
}

int thsp_coo2csr(int nrow, int ncol, int nnz, const int* row_ind, const int* col_ind, const double* val, int* row_ptr,
                 int* out_col_ind, double* out_val, double* diagonal, int* ndiag, thsp_stream_t stream)
{
    (void)ncol;
    if (ensure_device()) return 1;
    cudaStream_t s = as_stream(stream);
    const int *sk, *perm;
    if (bucket_order(nrow, nnz, row_ind, row_ptr, &sk, &perm, s)) return 1;
    if (nnz > 0) {
        gather_kernel<<<div_up(nnz, 256), 256, 0, s>>>(nnz, perm, col_ind, val, out_col_ind, out_val);
        THSP_LAUNCH_CHECK();
    }
    if (diagonal || ndiag) return pack_diagonal(nnz, row_ind, col_ind, val, nrow, diagonal, ndiag, s);
    return 0;
}

int thsp_coo2csc(int nrow, int ncol, int nnz, const int* row_ind, const int* col_ind, const double* val, int* col_ptr,
                 int* out_row_ind, double* out_val, thsp_stream_t stream)
{
    (void)nrow;
    if (ensure_device()) return 1;
    cudaStream_t s = as_stream(stream);
    const int *sk, *perm;
    if (bucket_order(ncol, nnz, col_ind, col_ptr, &sk, &perm, s)) return 1;
    if (nnz > 0) {
        gather_kernel<<<div_up(nnz, 256), 256, 0, s>>>(nnz, perm, row_ind, val, out_row_ind, out_val);
        THSP_LAUNCH_CHECK();
    }
    return 0;
}

int thsp_coo2ell_width(int nrow, int nnz, const int* row_ind, int* width, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    cudaStream_t s = as_stream(stream);
    *width = 0;
    if (nrow <= 0 || nnz <= 0) return 0;
    int* cnt = static_cast<int*>(scratch(sizeof(int) * ((size_t)nrow + 2), 2));
    int* flag = static_cast<int*>(scratch(sizeof(int), 1));
    if (!cnt || !flag) return 1;
    int* mx = cnt + nrow;
    THSP_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)nrow + 2), s));
    hist_kernel<<<std::min(div_up(nnz, 256), sm_count() * 16), 256, 0, s>>>(nnz, row_ind, cnt, flag);
    THSP_LAUNCH_CHECK();
    max_len_kernel<<<std::min(div_up(nrow, 256), sm_count() * 8), 256, 0, s>>>(nrow, cnt, mx);
    THSP_LAUNCH_CHECK();
    THSP_CUDA(cudaMemcpyAsync(width, mx, sizeof(int), cudaMemcpyDeviceToHost, s));
    THSP_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int thsp_coo2ell(int nrow, int ncol, int nnz, const int* row_ind, const int* col_ind, const double* val, int width,
                 int* out_col_ind, double* out_val, double* diagonal, int* ndiag, thsp_stream_t stream)
{
    (void)ncol;
    if (ensure_device()) return 1;
    cudaStream_t s = as_stream(stream);
    const size_t total = (size_t)nrow * (size_t)width;
    if (total) {
        THSP_CUDA(cudaMemsetAsync(out_col_ind, 0, sizeof(int) * total, s));   // padding: column 0
        THSP_CUDA(cudaMemsetAsync(out_val, 0, sizeof(double) * total, s));    // padding: +0.0
    }
    int* rp = static_cast<int*>(scratch(sizeof(int) * ((size_t)nrow + 2), 2));
    if (!rp) return 1;
    const int *sk, *perm;
    if (bucket_order(nrow, nnz, row_ind, rp, &sk, &perm, s)) return 1;
    if (nnz > 0) {
        ell_place_kernel<<<div_up(nnz, 256), 256, 0, s>>>(nnz, nrow, sk, perm, rp, col_ind, val, out_col_ind, out_val);
        THSP_LAUNCH_CHECK();
    }
    if (diagonal || ndiag) return pack_diagonal(nnz, row_ind, col_ind, val, nrow, diagonal, ndiag, s);
    return 0;
}

int thsp_csr2dia_offsets(int nrow, int ncol, const int* row_ptr, const int* col_ind, int* ndiags, int* offsets,
                         int offsets_capacity, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    cudaStream_t s = as_stream(stream);
    const int span = nrow + ncol - 1;
    *ndiags = 0;
    if (span <= 0 || nrow <= 0) return 0;
    int* seen = static_cast<int*>(scratch(sizeof(int) * (2 * (size_t)span + 4), 2));
    if (!seen) return 1;
    int* pos = seen + span + 1;
    THSP_CUDA(cudaMemsetAsync(seen, 0, sizeof(int) * ((size_t)span + 1), s));
    dia_mark_kernel<<<div_up(nrow, 256), 256, 0, s>>>(nrow, span, row_ptr, col_ind, seen);
    THSP_LAUNCH_CHECK();
    if (exclusive_scan(span, seen, pos, s)) return 1;
    if (offsets) {
        dia_offsets_kernel<<<div_up(span, 256), 256, 0, s>>>(span, nrow, seen, pos, offsets_capacity, offsets);
        THSP_LAUNCH_CHECK();
    }
    THSP_CUDA(cudaMemcpyAsync(ndiags, pos + span, sizeof(int), cudaMemcpyDeviceToHost, s));
    THSP_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int thsp_csr2dia_fill(int nrow, int ncol, const int* row_ptr, const int* col_ind, const double* val, int ndiags,
                      const int* offsets, double* values, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    cudaStream_t s = as_stream(stream);
    const int span = nrow + ncol - 1;
    if (nrow <= 0 || ndiags <= 0) return 0;
    int* slot = static_cast<int*>(scratch(sizeof(int) * ((size_t)span + 2), 2));
    if (!slot) return 1;
    THSP_CUDA(cudaMemsetAsync(slot, 0xff, sizeof(int) * ((size_t)span + 1), s));
    THSP_CUDA(cudaMemsetAsync(values, 0, sizeof(double) * (size_t)nrow * (size_t)ndiags, s));
    dia_slot_kernel<<<div_up(ndiags, 256), 256, 0, s>>>(ndiags, nrow, offsets, slot);
    THSP_LAUNCH_CHECK();
    dia_fill_kernel<<<div_up(nrow, 256), 256, 0, s>>>(nrow, span, ndiags, row_ptr, col_ind, val, slot, values);
    THSP_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
