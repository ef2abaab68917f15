// exchange.cu -- the vector half of one power-iteration step fused with its exchange over NVLink.
//
// The reference has no iterated loop; composed from its own calls a step is
//     y = A x ; s = vec_dot(y, y) ; x = vec_axpby(1/sqrt(s), y, 0, y)        (src/vec_vec.cpp:15-53)
// Row-partitioned over the GPUs of one box (src/mat_vec.cpp:230-297 made multi-GPU) the dot needs the
// partial sums of all ranks and the next SpMV needs the pieces of x that other ranks own.  Both
// exchanges are a few bytes to a few MB: calling a collective library for them costs more in
// launch + rendezvous latency than the transfer itself.  Here the kernels that PRODUCE the data
// store it straight into the peers' memory (peer pointers into a symmetric allocation, NVLink),
// followed by a release flag; the kernels that CONSUME it acquire the flag.  One stream, no host
// round trip, no collective call in the loop:
//
//   thsp_xchg_sumsq_publish_f64   per-CTA partial sums of y^2 (fixed tree) -> the last CTA adds them
//                                 in CTA order, stores the rank's partial into slot [parity][rank]
//                                 of EVERY rank's control block, then the iteration number into
//                                 the matching flag.
//   thsp_xchg_scale_push_f64      one warp waits until the flags of all ranks show this iteration
//                                 (lane r on rank r) and adds the partials in rank order - every
//                                 rank gets the same bits; then x_own = y / sqrt(sum) goes into the
//                                 local replica and, for the index ranges other ranks read, into
//                                 their replicas as well; when all CTAs are done the last one
//                                 raises the "halo from <rank>" flag on those ranks.
//   thsp_xchg_wait                one thread spins until the halo flags of the given source ranks
//                                 show the iteration; launched in front of the rows that read
//                                 remote parts of x.
//
// Control block (uint64 words, symmetric memory, zero-initialised): [0,32) partial sums as double
// bits [parity][rank], [32,64) their flags, [64,80) halo flags [source rank], [80] timeout marker.
// Two parities are enough: a rank can only be one publish ahead of the slowest reader, because
// its next scale waits for that reader's next partial.  Spins give up after ~15 s and set word 80
// instead of hanging the GPU.
#include <algorithm>

#include "common.cuh"
#include "tree_sum.cuh"

namespace thsp {

static constexpr int kXThreads = 256;
static constexpr int kXMaxRanks = 16;
static constexpr int kXPart = 0, kXPartFlag = 32, kXHalo = 64, kXErr = 80;
static constexpr long long kXSpinLimit = 30000000000LL;   // cycles (~15 s)

__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint64_t* p, uint64_t v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool spin_until(const uint64_t* flag, uint64_t want, uint64_t* err)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < want) {
        if (clock64() - t0 > kXSpinLimit) {
            *err = 1;
            return false;
        }
        __nanosleep(64);
    }
    return true;
}

__device__ __forceinline__ double block_sum_x(double v)
{
    __shared__ double ws[kXThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = add_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < kXThreads / 32 ? ws[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) r = add_rn(r, __shfl_xor_sync(0xffffffffu, r, o));
    }
    __syncthreads();
    return r;   // valid in thread 0
}

struct XPeers {
    uint64_t* ctrl[kXMaxRanks];
};

__global__ void __launch_bounds__(kXThreads) xchg_sumsq_publish_kernel(int64_t n, const double* __restrict__ y,
                                                                       double* __restrict__ part, unsigned* __restrict__ ticket,
                                                                       uint64_t iter, int world, int rank, XPeers peers)
{
    __shared__ bool last;
    const int64_t stride = (int64_t)gridDim.x * kXThreads;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kXThreads + threadIdx.x; i < n; i += stride) {
        const double v = y[i];
        acc = add_rn(acc, mul_rn(v, v));
    }
    const double r = block_sum_x(acc);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = r;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double tot = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kXThreads) tot = add_rn(tot, __ldcg(part + i));   // fixed order per thread
    tot = block_sum_x(tot);
    if (threadIdx.x < 32) {
        // lane p talks to rank p: value, one system-scope fence, then the flag (a fence orders the
        // lane's earlier store before its later one for every observer - no release needed per flag)
        tot = __shfl_sync(0xffffffffu, tot, 0);
        const int par = (int)(iter & 1);
        const int p = threadIdx.x;
        if (p == 0) *ticket = 0;
        if (p < world) st_relaxed_sys(peers.ctrl[p] + kXPart + par * kXMaxRanks + rank, (uint64_t)__double_as_longlong(tot));
        __threadfence_system();
        if (p < world) st_relaxed_sys(peers.ctrl[p] + kXPartFlag + par * kXMaxRanks + rank, iter);
    }
}

// One warp: lane r waits for rank r's partial of this iteration; the partials are then added in
// rank order (every rank gets the same bits) and the sum lands in *sumsq_out for the scale kernel.
__global__ void xchg_reduce_kernel(uint64_t* ctrl, uint64_t iter, int world, double* __restrict__ sumsq_out)
{
    const int r = threadIdx.x;
    const int par = (int)(iter & 1);
    double v = 0.0;
    bool ok = true;
    if (r < world) {
        ok = spin_until(ctrl + kXPartFlag + par * kXMaxRanks + r, iter, ctrl + kXErr);
        v = __longlong_as_double((long long)ld_acquire_sys(ctrl + kXPart + par * kXMaxRanks + r));
    }
    double tot = 0.0;
    for (int k = 0; k < world; ++k) tot = add_rn(tot, __shfl_sync(0xffffffffu, v, k));
    // a peer that never published: poison the sum, so that the step cannot go on with stale partials unnoticed
    if (!__all_sync(0xffffffffu, ok)) tot = __longlong_as_double(0x7ff8000000000000LL);
    if (r == 0) *sumsq_out = tot;
}

struct XDests {
    double* x[kXMaxRanks];        // the destination rank's replica of x
    uint64_t* ctrl[kXMaxRanks];   // its control block
    int64_t lo[kXMaxRanks], hi[kXMaxRanks];   // global index range of x it reads from this rank
    int n;
};

__global__ void __launch_bounds__(kXThreads) xchg_scale_push_kernel(int64_t n, const double* __restrict__ y,
                                                                    unsigned* __restrict__ ticket, uint64_t iter, int rank,
                                                                    double* __restrict__ x_local, int64_t offset, XDests dst,
                                                                    const double* __restrict__ sumsq)
{
    __shared__ bool last;
    const double inv = 1.0 / sqrt(*sumsq);
    const int64_t stride = (int64_t)gridDim.x * kXThreads;
    bool remote = false;
    for (int64_t i = (int64_t)blockIdx.x * kXThreads + threadIdx.x; i < n; i += stride) {
        const double v = mul_rn(inv, y[i]);   // vec_axpby's beta == 0 branch: w = alpha * x
        const int64_t g = offset + i;
        x_local[g] = v;
        for (int d = 0; d < dst.n; ++d)
            if (g >= dst.lo[d] && g < dst.hi[d]) {
                dst.x[d][g] = v;
                remote = true;
            }
    }
    if (dst.n == 0) return;
    if (remote) __threadfence_system();   // this thread's remote stores are visible before the ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();           // ... and ordered before the ticket in the thread that takes it
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x < 32) {
        if (threadIdx.x == 0) *ticket = 0;
        __threadfence_system();
        if ((int)threadIdx.x < dst.n) st_relaxed_sys(dst.ctrl[threadIdx.x] + kXHalo + rank, iter);
    }
}

// ---- the whole vector half of a step in ONE kernel -----------------------------------------------------------------
// Input: the per-tile sums of squares the SpMV left behind (csr_spmv.cu epilogue, tree_sum.cuh).  All CTAs are resident
// at once (cooperative launch), because they wait for each other and for the peers:
//   1. CTAs reduce blocks of 4096 tile partials with the index-bit tree; the last one to finish completes the tree over
//      the block results and stores the rank's partial + flag into every rank's control block (lane p -> rank p);
//   2. a warp of EVERY CTA waits for the flags of all ranks (lane r on rank r) and combines the partials with the same
//      tree over rank numbers: same bits on every rank and - for aligned row blocks - as on one GPU;
//   3. x_own = y / sqrt(sum) into the local replica and, for the ranges other ranks read, into their replicas; the last
//      CTA to finish raises the "pieces from <rank>" flags.
// A peer that never shows up (~15 s) poisons the sum with NaN and no flag is raised: the failure is loud, not silent.
__global__ void __launch_bounds__(kXThreads) xchg_norm_scale_push_kernel(int ntiles, const double* __restrict__ tile_ss, double* part,
                                                                         unsigned* __restrict__ tickets, int64_t n,
                                                                         const double* __restrict__ y, uint64_t iter, int world, int rank,
                                                                         XPeers peers, double* __restrict__ x_local, int64_t offset,
                                                                         XDests dst, double* __restrict__ sumsq_out)
{
    __shared__ bool last;
    __shared__ double s_tot;
    static_assert(kXThreads == kTreeThreads, "block_tree_sum is written for this CTA size");
    const int nb = (ntiles + kTreeBlock - 1) / kTreeBlock;
    double* part_b = part + nb;
    const int par = (int)(iter & 1);
    // ---- 1. this rank's partial
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const double r = block_tree_sum(tile_ss + (size_t)b * kTreeBlock, min(kTreeBlock, ntiles - b * kTreeBlock));
        if (threadIdx.x == 0) part[b] = r;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(tickets, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        const double mine = nb > 0 ? block_tree_finish(part, nb, part, part_b) : 0.0;
        if (threadIdx.x < 32) {
            const int p = threadIdx.x;
            if (p == 0) tickets[0] = 0;
            if (p < world) st_relaxed_sys(peers.ctrl[p] + kXPart + par * kXMaxRanks + rank, (uint64_t)__double_as_longlong(mine));
            __threadfence_system();
            if (p < world) st_relaxed_sys(peers.ctrl[p] + kXPartFlag + par * kXMaxRanks + rank, iter);
        }
    }
    // ---- 2. everybody's partials, combined by the tree over rank numbers
    if (threadIdx.x < 32) {
        uint64_t* ctrl = peers.ctrl[rank];
        const int r = threadIdx.x;
        double v = 0.0;
        bool ok = true;
        if (r < world) {
            ok = spin_until(ctrl + kXPartFlag + par * kXMaxRanks + r, iter, ctrl + kXErr);
            v = __longlong_as_double((long long)ld_acquire_sys(ctrl + kXPart + par * kXMaxRanks + r));
        }
#pragma unroll
        for (int o = 1; o < kXMaxRanks; o <<= 1) v = add_rn(v, __shfl_down_sync(0xffffffffu, v, o));   // lanes >= world hold +0.0
        if (!__all_sync(0xffffffffu, ok)) v = __longlong_as_double(0x7ff8000000000000LL);
        if (r == 0) s_tot = v;
    }
    __syncthreads();
    const double tot = s_tot;
    if (blockIdx.x == 0 && threadIdx.x == 0) *sumsq_out = tot;
    if (tot != tot) return;   // poisoned: leave x alone and raise nothing
    // ---- 3. normalise, store locally and into the neighbours' replicas
    const double inv = 1.0 / sqrt(tot);
    const int64_t stride = (int64_t)gridDim.x * kXThreads;
    bool remote = false;
    constexpr int U = 4;   // independent loads in flight per thread
    for (int64_t i0 = (int64_t)blockIdx.x * kXThreads + threadIdx.x; i0 < n; i0 += U * stride) {
        double v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = i0 + u * stride < n ? ld_stream(y + i0 + u * stride) : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i >= n) break;
            const double w = mul_rn(inv, v[u]);   // vec_axpby's beta == 0 branch: w = alpha * x
            const int64_t g = offset + i;
            x_local[g] = w;
            for (int d = 0; d < dst.n; ++d)
                if (g >= dst.lo[d] && g < dst.hi[d]) {
                    dst.x[d][g] = w;
                    remote = true;
                }
        }
    }
    if (dst.n == 0) return;
    if (remote) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(tickets + 1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x < 32) {
        if (threadIdx.x == 0) tickets[1] = 0;
        __threadfence_system();
        if ((int)threadIdx.x < dst.n) st_relaxed_sys(dst.ctrl[threadIdx.x] + kXHalo + rank, iter);
    }
}

// ---- the same step with the normalisation DEFERRED into the next product ---------------------------------------------
// x = y / ||y|| is never written: the next SpMV multiplies every gathered y_j by 1/||y|| itself (csr_stream_kernel,
// kScale - the same rounding the normalising pass would have done, so y keeps its bits).  What is left of the vector half:
//   1. the tree over the tile partials, published to every rank (as above);
//   2. the pieces of the RAW y that other ranks read, copied into their replicas of the vector (16-byte stores over
//      NVLink where the piece is aligned), then the "pieces from <rank>" flags - none of this waits for the norm;
//   3. the CTA that published waits for the partials of all ranks, combines them (tree over rank numbers) and leaves the
//      sum and 1/sqrt(sum) on the device for the next product.
// Nothing here reads or writes the n-element vector except the pushed pieces: the 2 x 8 n bytes of the normalising pass
// are gone, and only one CTA waits for the peers (no cooperative launch).
__global__ void __launch_bounds__(kXThreads) xchg_norm_push_kernel(int ntiles, const double* __restrict__ tile_ss, double* part,
                                                                   unsigned* __restrict__ tickets, const double* __restrict__ y,
                                                                   uint64_t iter, int world, int rank, XPeers peers, int64_t offset,
                                                                   XDests dst, double* __restrict__ sumsq_out, double* __restrict__ inv_out)
{
    __shared__ bool last_tree, last_push;
    static_assert(kXThreads == kTreeThreads, "block_tree_sum is written for this CTA size");
    const int nb = (ntiles + kTreeBlock - 1) / kTreeBlock;
    double* part_b = part + nb;
    const int par = (int)(iter & 1);
    // ---- 1. this rank's partial
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const double r = block_tree_sum(tile_ss + (size_t)b * kTreeBlock, min(kTreeBlock, ntiles - b * kTreeBlock));
        if (threadIdx.x == 0) part[b] = r;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        last_tree = atomicAdd(tickets, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last_tree) {
        __threadfence();
        const double mine = nb > 0 ? block_tree_finish(part, nb, part, part_b) : 0.0;
        if (threadIdx.x < 32) {
            const int p = threadIdx.x;
            if (p == 0) tickets[0] = 0;
            if (p < world) st_relaxed_sys(peers.ctrl[p] + kXPart + par * kXMaxRanks + rank, (uint64_t)__double_as_longlong(mine));
            __threadfence_system();
            if (p < world) st_relaxed_sys(peers.ctrl[p] + kXPartFlag + par * kXMaxRanks + rank, iter);
        }
    }
    // ---- 2. the pieces of y the neighbours read, as they are
    if (dst.n > 0) {
        bool remote = false;
        const int64_t t = (int64_t)blockIdx.x * kXThreads + threadIdx.x, stride = (int64_t)gridDim.x * kXThreads;
        for (int d = 0; d < dst.n; ++d) {
            const int64_t lo = dst.lo[d], len = dst.hi[d] - dst.lo[d];
            if (len <= 0) continue;
            const double* src = y + (lo - offset);
            double* out = dst.x[d] + lo;
            if (((((uintptr_t)src) | ((uintptr_t)out)) & 15) == 0) {
                const int64_t n2 = len >> 1;
                for (int64_t i = t; i < n2; i += stride) {
                    reinterpret_cast<double2*>(out)[i] = ld_stream2(src + 2 * i);
                    remote = true;
                }
                if ((len & 1) && t == 0) {
                    out[len - 1] = src[len - 1];
                    remote = true;
                }
            } else {
                for (int64_t i = t; i < len; i += stride) {
                    out[i] = ld_stream(src + i);
                    remote = true;
                }
            }
        }
        if (remote) __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            last_push = atomicAdd(tickets + 1, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (last_push && threadIdx.x < 32) {
            if (threadIdx.x == 0) tickets[1] = 0;
            __threadfence_system();
            if ((int)threadIdx.x < dst.n) st_relaxed_sys(dst.ctrl[threadIdx.x] + kXHalo + rank, iter);
        }
    }
    // ---- 3. everybody's partials, combined by the tree over rank numbers: the norm for the next product
    if (last_tree && threadIdx.x < 32) {
        uint64_t* ctrl = peers.ctrl[rank];
        const int r = threadIdx.x;
        double v = 0.0;
        bool ok = true;
        if (r < world) {
            ok = spin_until(ctrl + kXPartFlag + par * kXMaxRanks + r, iter, ctrl + kXErr);
            v = __longlong_as_double((long long)ld_acquire_sys(ctrl + kXPart + par * kXMaxRanks + r));
        }
#pragma unroll
        for (int o = 1; o < kXMaxRanks; o <<= 1) v = add_rn(v, __shfl_down_sync(0xffffffffu, v, o));   // lanes >= world hold +0.0
        if (!__all_sync(0xffffffffu, ok)) v = __longlong_as_double(0x7ff8000000000000LL);   // a peer never showed up: loud
        if (r == 0) {
            *sumsq_out = v;
            *inv_out = __ddiv_rn(1.0, __dsqrt_rn(v));
        }
    }
}

__global__ void xchg_wait_kernel(uint64_t* ctrl, uint64_t iter, unsigned src_mask)
{
    if (threadIdx.x != 0) return;
    for (int r = 0; r < kXMaxRanks; ++r)
        if (src_mask & (1u << r)) spin_until(ctrl + kXHalo + r, iter, ctrl + kXErr);
}

}  // namespace thsp

using namespace thsp;

extern "C" {

int thsp_xchg_ctrl_bytes(void) { return 128 * (int)sizeof(uint64_t); }

// work: 2 unsigned tickets (zero-initialised) followed by space for the per-CTA partial sums
// (thsp_xchg_work_bytes()), private to this rank.
int thsp_xchg_work_bytes(void) { return 64 + 8 * 4096; }

int thsp_xchg_sumsq_publish_f64(int64_t n, const double* y, uint64_t iter, int world, int rank, void* const* peer_ctrl, void* work,
                                thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    THSP_REQUIRE(world >= 1 && world <= kXMaxRanks && rank >= 0 && rank < world, "bad world / rank");
    XPeers pe;
    for (int p = 0; p < kXMaxRanks; ++p) pe.ctrl[p] = p < world ? static_cast<uint64_t*>(peer_ctrl[p]) : nullptr;
    int grid = (int)std::min<int64_t>(4096, std::max<int64_t>(1, (n + kXThreads * 8 - 1) / (kXThreads * 8)));
    grid = std::min(grid, sm_count() * 8);
    unsigned* ticket = static_cast<unsigned*>(work);
    double* part = reinterpret_cast<double*>(static_cast<char*>(work) + 64);
    xchg_sumsq_publish_kernel<<<grid, kXThreads, 0, as_stream(stream)>>>(n, y, part, ticket, iter, world, rank, pe);
    THSP_LAUNCH_CHECK();
    return 0;
}

int thsp_xchg_scale_push_f64(int64_t n, const double* y, uint64_t iter, int world, int rank, void* ctrl_local, void* work,
                             double* x_local, int64_t offset, int ndest, double* const* dest_x, void* const* dest_ctrl,
                             const int64_t* dest_lo, const int64_t* dest_hi, double* sumsq_out, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    THSP_REQUIRE(world >= 1 && world <= kXMaxRanks && rank >= 0 && rank < world, "bad world / rank");
    THSP_REQUIRE(ndest >= 0 && ndest < kXMaxRanks, "too many destinations");
    THSP_REQUIRE(sumsq_out != nullptr, "sumsq_out is the device scalar the sum is handed over in");
    XDests d;
    d.n = ndest;
    for (int k = 0; k < kXMaxRanks; ++k) {
        d.x[k] = k < ndest ? dest_x[k] : nullptr;
        d.ctrl[k] = k < ndest ? static_cast<uint64_t*>(dest_ctrl[k]) : nullptr;
        d.lo[k] = k < ndest ? dest_lo[k] : 0;
        d.hi[k] = k < ndest ? dest_hi[k] : 0;
    }
    cudaStream_t s = as_stream(stream);
    xchg_reduce_kernel<<<1, 32, 0, s>>>(static_cast<uint64_t*>(ctrl_local), iter, world, sumsq_out);
    THSP_LAUNCH_CHECK();
    int grid = (int)std::min<int64_t>(sm_count() * 8, std::max<int64_t>(1, (n + kXThreads * 4 - 1) / (kXThreads * 4)));
    unsigned* ticket = static_cast<unsigned*>(work) + 1;
    xchg_scale_push_kernel<<<grid, kXThreads, 0, s>>>(n, y, ticket, iter, rank, x_local, offset, d, sumsq_out);
    THSP_LAUNCH_CHECK();
    return 0;
}

int thsp_xchg_norm_scale_push_f64(int64_t n, const double* y, const double* tile_ss, uint64_t iter, int world, int rank,
                                  void* const* peer_ctrl, void* work, double* x_local, int64_t offset, int ndest,
                                  double* const* dest_x, void* const* dest_ctrl, const int64_t* dest_lo, const int64_t* dest_hi,
                                  double* sumsq_out, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    THSP_REQUIRE(world >= 1 && world <= kXMaxRanks && rank >= 0 && rank < world, "bad world / rank");
    THSP_REQUIRE(ndest >= 0 && ndest < kXMaxRanks, "too many destinations");
    THSP_REQUIRE(sumsq_out != nullptr && tile_ss != nullptr, "tile_ss / sumsq_out missing");
    const int64_t ntiles64 = (n + 31) / 32;
    THSP_REQUIRE(ntiles64 <= (int64_t)4095 * kTreeBlock, "slice too long for the work buffer");
    XPeers pe;
    for (int p = 0; p < kXMaxRanks; ++p) pe.ctrl[p] = p < world ? static_cast<uint64_t*>(peer_ctrl[p]) : nullptr;
    XDests d;
    d.n = ndest;
    for (int k = 0; k < kXMaxRanks; ++k) {
        d.x[k] = k < ndest ? dest_x[k] : nullptr;
        d.ctrl[k] = k < ndest ? static_cast<uint64_t*>(dest_ctrl[k]) : nullptr;
        d.lo[k] = k < ndest ? dest_lo[k] : 0;
        d.hi[k] = k < ndest ? dest_hi[k] : 0;
    }
    // all CTAs must be resident together: they wait for each other's block results and for the peers
    static int per_sm[16] = {};
    int dev = 0;
    THSP_CUDA(cudaGetDevice(&dev));
    if (!per_sm[dev & 15]) {
        int k = 0;
        THSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, xchg_norm_scale_push_kernel, kXThreads, 0));
        per_sm[dev & 15] = std::max(1, std::min(k, 4));
    }
    int grid = (int)std::min<int64_t>((int64_t)sm_count() * per_sm[dev & 15], std::max<int64_t>(1, (n + kXThreads * 4 - 1) / (kXThreads * 4)));
    int ntiles = (int)ntiles64;
    unsigned* tickets = static_cast<unsigned*>(work);
    double* part = reinterpret_cast<double*>(static_cast<char*>(work) + 64);
    void* args[] = {&ntiles, (void*)&tile_ss, &part, &tickets, &n, (void*)&y, &iter, &world, &rank, &pe, &x_local, &offset, &d, &sumsq_out};
    THSP_CUDA(cudaLaunchCooperativeKernel((const void*)xchg_norm_scale_push_kernel, dim3(grid), dim3(kXThreads), args, 0, as_stream(stream)));
    THSP_LAUNCH_CHECK();
    return 0;
}

int thsp_xchg_norm_push_f64(int64_t n, const double* y, const double* tile_ss, uint64_t iter, int world, int rank,
                            void* const* peer_ctrl, void* work, int64_t offset, int ndest, double* const* dest_x,
                            void* const* dest_ctrl, const int64_t* dest_lo, const int64_t* dest_hi, double* sumsq_out,
                            double* inv_out, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    THSP_REQUIRE(world >= 1 && world <= kXMaxRanks && rank >= 0 && rank < world, "bad world / rank");
    THSP_REQUIRE(ndest >= 0 && ndest < kXMaxRanks, "too many destinations");
    THSP_REQUIRE(sumsq_out != nullptr && inv_out != nullptr && tile_ss != nullptr, "tile_ss / sumsq_out / inv_out missing");
    const int64_t ntiles64 = (n + 31) / 32;
    THSP_REQUIRE(ntiles64 <= (int64_t)4095 * kTreeBlock, "slice too long for the work buffer");
    XPeers pe;
    for (int p = 0; p < kXMaxRanks; ++p) pe.ctrl[p] = p < world ? static_cast<uint64_t*>(peer_ctrl[p]) : nullptr;
    XDests d;
    d.n = ndest;
    int64_t pushed = 0;
    for (int k = 0; k < kXMaxRanks; ++k) {
        d.x[k] = k < ndest ? dest_x[k] : nullptr;
        d.ctrl[k] = k < ndest ? static_cast<uint64_t*>(dest_ctrl[k]) : nullptr;
        d.lo[k] = k < ndest ? dest_lo[k] : 0;
        d.hi[k] = k < ndest ? dest_hi[k] : 0;
        if (k < ndest) {
            THSP_REQUIRE(dest_lo[k] >= offset && dest_hi[k] <= offset + n, "a rank pushes pieces of its own slice only");
            pushed = std::max(pushed, dest_hi[k] - dest_lo[k]);
        }
    }
    int ntiles = (int)ntiles64;
    const int nb = (ntiles + kTreeBlock - 1) / kTreeBlock;
    // enough CTAs for the blocks of the tree and for 16-byte stores of the longest piece, at most two per SM
    const int64_t want = std::max<int64_t>(nb, (pushed / 2 + kXThreads - 1) / kXThreads);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count() * 2));
    unsigned* tickets = static_cast<unsigned*>(work);
    double* part = reinterpret_cast<double*>(static_cast<char*>(work) + 64);
    xchg_norm_push_kernel<<<grid, kXThreads, 0, as_stream(stream)>>>(ntiles, tile_ss, part, tickets, y, iter, world, rank, pe, offset, d,
                                                                    sumsq_out, inv_out);
    THSP_LAUNCH_CHECK();
    return 0;
}

int thsp_xchg_wait(void* ctrl_local, uint64_t iter, unsigned src_mask, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    if (!src_mask) return 0;
    xchg_wait_kernel<<<1, 32, 0, as_stream(stream)>>>(static_cast<uint64_t*>(ctrl_local), iter, src_mask);
    THSP_LAUNCH_CHECK();
    return 0;
}

// 1 if a spin of this rank has timed out since the control block was zeroed
int thsp_xchg_timed_out(const void* ctrl_local, int* flag_host, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    uint64_t v = 0;
    THSP_CUDA(cudaMemcpyAsync(&v, static_cast<const uint64_t*>(ctrl_local) + kXErr, sizeof(v), cudaMemcpyDeviceToHost, as_stream(stream)));
    THSP_CUDA(cudaStreamSynchronize(as_stream(stream)));
    *flag_host = v ? 1 : 0;
    return 0;
}

}  // extern "C"
