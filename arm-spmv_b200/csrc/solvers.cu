// solvers.cu -- the callers the reference prepared for and never wrote, behind the C ABI.
//
// The reference keeps a `diagonal` array per CSR / ELL matrix "for SymGS" (include/matrix.h:36,81; filled at
// src/matrix.cpp:146-153, 491-499) and ships the vector kernels of a Krylov loop (vec_dot / vec_axpby,
// src/vec_vec.cpp:15-94; Vector::AddScaled / Add2Scaled, src/vector.cpp:96-159) without a caller.  Here:
//
//  SymGS  one symmetric Gauss-Seidel sweep x <- SGS(A, r; x) that consumes the diagonal array: per row
//             t = sum_j a_ij x_j (from 0, stored order, the diagonal included) ;  s = r_i - t ;  s += x_i d_i ;  x_i = s / d_i
//         forward over the rows, then backward.  Rows are grouped by COLOUR (no two rows of a colour touch each other), a
//         colour is one kernel launch with a thread per row adding its row left to right with unfused arithmetic; colours
//         ascending, then descending.  Inside a colour the rows are independent, so the result does not depend on the
//         order in which the GPU happens to run them: the CPU checker's twin (oracle/oracle.c) walks the same colours serially and
//         gets the same bits.  The colouring is a deterministic speculative greedy one (rounds of: uncoloured rows take
//         the smallest colour no coloured neighbour has, conflicts between rows coloured in the same round send the one
//         with the lower hash back), at most 64 colours; structurally non-symmetric matrices are handled (an edge seen
//         from one side only still separates its two rows).
//  CG     preconditioned conjugate gradients on a CSR plan (the SpMV is thsp_csr_plan_spmv_f64), preconditioner: none,
//         Jacobi (z = r / d) or one SymGS sweep from z = 0.  The vector updates are the reference's own forms
//         (AddScaled: v += a x; vec_axpby with alpha == 1: w = beta y + x), the scalars stay on the device, and every
//         dot product is taken in the canonical order of tree_sum.cuh - so the checker's CG, composed from its own serial SpMV,
//         the same forms and the same order, reproduces every iterate bit for bit when the SpMV kernel keeps the
//         reference's order (stream / scalar).
// Everything is HBM-bound: a SymGS sweep reads the matrix twice (forward + backward), a CG iteration once.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tree_sum.cuh"

namespace thsp {

int exclusive_scan(int n, const int* in, int* out, cudaStream_t s);   // convert.cu
int stream_gs_color(int nrow, int nnz_total, double mean_len, const int* rp, const int* col, const double* val, const int* rows,
                    const double* r, const double* diag, double* x, cudaStream_t s);   // csr_spmv.cu

__host__ __device__ inline uint64_t color_prio(uint64_t v)
{
    v += 0x9E3779B97F4A7C15ull;
    v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull;
    v = (v ^ (v >> 27)) * 0x94D049BB133111EBull;
    return v ^ (v >> 31);
}

// ---- colouring ---------------------------------------------------------------------------------------------------
// assign: every uncoloured row takes the smallest colour none of its already coloured out-neighbours has
__global__ void __launch_bounds__(256) color_assign_kernel(int nrow, const int* __restrict__ rp, const int* __restrict__ ci,
                                                           const int* __restrict__ cin, int* __restrict__ cout, int* __restrict__ overflow,
                                                           const unsigned long long* __restrict__ forbid)
{
    const int v = blockIdx.x * 256 + threadIdx.x;
    if (v >= nrow) return;
    const int c = cin[v];
    if (c >= 0) {
        cout[v] = c;
        return;
    }
    unsigned long long used = forbid[v];   // colours of rows that point at v without v pointing back (learnt from lost conflicts)
    for (int p = rp[v]; p < rp[v + 1]; ++p) {
        const int u = ci[p];
        if (u == v || u < 0 || u >= nrow) continue;
        const int cu = cin[u];
        if (cu >= 0) used |= 1ull << cu;
    }
    const int pick = __ffsll((long long)~used) - 1;   // 64 colours at most
    if (pick < 0) {
        *overflow = 1;
        cout[v] = 0;
        return;
    }
    cout[v] = pick;
}
// resolve: an edge v -> u whose ends got the same colour sends one of them back: the one coloured in this round
// (if only one was), else the one with the lower hash.  All writers write the same value: no race that matters.
__global__ void __launch_bounds__(256) color_resolve_kernel(int nrow, const int* __restrict__ rp, const int* __restrict__ ci,
                                                            const int* __restrict__ cin, const int* __restrict__ cout, int* __restrict__ redo,
                                                            unsigned long long* __restrict__ forbid)
{
    const int v = blockIdx.x * 256 + threadIdx.x;
    if (v >= nrow) return;
    const int cv = cout[v];
    const bool vnew = cin[v] < 0;
    const uint64_t pv = color_prio((uint64_t)v);
    for (int p = rp[v]; p < rp[v + 1]; ++p) {
        const int u = ci[p];
        if (u == v || u < 0 || u >= nrow) continue;
        if (cout[u] != cv) continue;
        const bool unew = cin[u] < 0;
        if (!unew && !vnew) continue;   // cannot happen for colours given in earlier rounds, kept for safety
        bool u_loses;
        if (unew && !vnew) u_loses = true;
        else if (!unew && vnew) u_loses = false;
        else {
            const uint64_t pu = color_prio((uint64_t)u);
            u_loses = pu < pv || (pu == pv && u < v);
        }
        redo[u_loses ? u : v] = 1;
        // A row coloured in THIS round that clashes with a row coloured EARLIER cannot have seen it (it avoids the colours
        // of the coloured rows it points at): the entry exists from the other side only.  Left alone it would pick the same
        // colour again next round, for ever - so it remembers what to avoid.  Clashes between two rows of the same round
        // need no memory: next round the winner is an earlier row (seen and avoided, or caught here).  A set union: the
        // order of the atomics does not matter.  Forbidding after EVERY lost clash was measured too: 61 colours instead
        // of 19 on the 27-point stencil (profiles/r02_solvers.txt).
        if (unew != vnew) atomicOr(forbid + (unew ? u : v), 1ull << cv);
    }
}
// ---- first attempt: the greedy colouring in NATURAL row order, level by level ------------------------------------------
// colour(v) = smallest colour no lower-numbered neighbour has - what a serial greedy pass over rows 0, 1, 2, ... produces.
// A row can be coloured as soon as all its lower neighbours are: every row counts its lower neighbours (`pending`), rows
// with none form the first worklist; a kernel colours the rows of the current worklist and, for each higher neighbour,
// counts one lower neighbour off - whoever reaches zero goes onto the next worklist.  Every entry is touched twice in
// all, one small launch per level of the dependency graph (7 n levels on an n^3 27-point stencil).
// Why bother: on matrices from regular grids this is the parity colouring - 8 colours for a 27-point stencil where the
// hash-ordered colouring below needs 19 - and the rows of a colour are every other grid point, so that the rows of a
// 32-row tile of the permuted matrix still gather from x lines they share (profiles/r02_solvers.txt: the sweep went from
// 6.2 to 2.4 ms).  The first version of this iterated the rule as a fixed point over ALL rows until nothing changed:
// the same colouring, but rows keep changing until the wave reaches them - 761 rounds x 0.7 ms = 0.53 s on 256^3.
// Deep graphs (a tridiagonal matrix has as many levels as rows) stop at a level limit; rows never reached - that, or
// neighbour counts that do not match from both sides in a non-symmetric pattern - stay uncoloured and are finished by the
// hash-ordered rounds below.
__global__ void __launch_bounds__(256) level_init_kernel(int nrow, const int* __restrict__ rp, const int* __restrict__ ci,
                                                         int* __restrict__ color, int* __restrict__ pending, int* __restrict__ list,
                                                         int* __restrict__ count)
{
    const int v = blockIdx.x * 256 + threadIdx.x;
    if (v >= nrow) return;
    int lower = 0;
    for (int p = rp[v]; p < rp[v + 1]; ++p) {
        const int u = ci[p];
        lower += (u >= 0 && u < v);
    }
    color[v] = -1;
    pending[v] = lower;
    if (lower == 0) list[atomicAdd(count, 1)] = v;
}
__global__ void __launch_bounds__(256) level_round_kernel(int nrow, const int* __restrict__ rp, const int* __restrict__ ci, int* color,
                                                          int* __restrict__ pending, const int* __restrict__ list_in,
                                                          int* __restrict__ list_out, int* __restrict__ counts, int round,
                                                          int* __restrict__ overflow)
{
    // counts[round % 3] rows wait in list_in; newcomers go to list_out / counts[(round + 1) % 3]; the third counter
    // (the previous round's input) is reset here for the round after next
    const int n_in = counts[round % 3];
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[(round + 2) % 3] = 0;
    int* n_out = counts + (round + 1) % 3;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n_in; i += gridDim.x * 256) {
        const int v = list_in[i];
        unsigned long long used = 0ull;
        const int e0 = rp[v], e1 = rp[v + 1];
        for (int p = e0; p < e1; ++p) {
            const int u = ci[p];
            if (u >= 0 && u < v) {
                const int cu = color[u];
                if (cu >= 0) used |= 1ull << cu;
            }
        }
        int c = __ffsll((long long)~used) - 1;
        if (c < 0) {
            *overflow = 1;   // more than 64 colours: give this attempt up
            c = 63;
        }
        color[v] = c;
        for (int p = e0; p < e1; ++p) {
            const int w = ci[p];
            if (w > v && w < nrow && atomicSub(pending + w, 1) == 1) list_out[atomicAdd(n_out, 1)] = w;
        }
    }
}
// after the natural-order attempt: an entry (v, u) whose ends share a colour sends the higher-numbered one back (it is
// the one that failed to avoid the other: it cannot see it, or the iteration was cut short) and tells it what to avoid
__global__ void __launch_bounds__(256) color_verify_kernel(int nrow, const int* __restrict__ rp, const int* __restrict__ ci,
                                                           const int* __restrict__ color, int* __restrict__ redo,
                                                           unsigned long long* __restrict__ forbid)
{
    const int v = blockIdx.x * 256 + threadIdx.x;
    if (v >= nrow) return;
    const int cv = color[v];
    if (cv < 0) return;
    for (int p = rp[v]; p < rp[v + 1]; ++p) {
        const int u = ci[p];
        if (u == v || u < 0 || u >= nrow || color[u] != cv) continue;
        const int loser = u > v ? u : v;
        redo[loser] = 1;
        atomicOr(forbid + loser, 1ull << cv);
    }
}

__global__ void __launch_bounds__(256) color_apply_kernel(int nrow, const int* __restrict__ cout, int* __restrict__ redo,
                                                          int* __restrict__ cin, int* __restrict__ left)
{
    const int v = blockIdx.x * 256 + threadIdx.x;
    int mine = 0;
    if (v < nrow) {
        if (redo[v]) {
            cin[v] = -1;
            redo[v] = 0;
            mine = 1;
        } else {
            cin[v] = cout[v];
        }
    }
    const int any = __syncthreads_count(mine);
    if (threadIdx.x == 0 && any) atomicAdd(left, any);
}
__global__ void __launch_bounds__(256) color_flag_kernel(int nrow, const int* __restrict__ color, int c, int* __restrict__ flag)
{
    const int v = blockIdx.x * 256 + threadIdx.x;
    if (v < nrow) flag[v] = color[v] == c ? 1 : 0;
}
__global__ void __launch_bounds__(256) color_place_kernel(int nrow, const int* __restrict__ color, int c, const int* __restrict__ pos,
                                                          int base, int* __restrict__ perm)
{
    const int v = blockIdx.x * 256 + threadIdx.x;
    if (v < nrow && color[v] == c) perm[base + pos[v]] = v;
}

// ---- the matrix permuted by colour (rows of a colour contiguous, columns unchanged): what lets a colour be STREAMED ----
__global__ void __launch_bounds__(256) perm_len_kernel(int nrow, const int* __restrict__ perm, const int* __restrict__ rp,
                                                       int* __restrict__ len, int* __restrict__ maxlen)
{
    const int k = blockIdx.x * 256 + threadIdx.x;
    int l = 0;
    if (k < nrow) {
        const int i = perm[k];
        l = rp[i + 1] - rp[i];
        len[k] = l;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l = max(l, __shfl_xor_sync(0xffffffffu, l, o));
    if ((threadIdx.x & 31) == 0 && l > 0) atomicMax(maxlen, l);
}
// a warp per permuted row copies the row's entries (coalesced on both sides)
__global__ void __launch_bounds__(256) perm_copy_kernel(int nrow, const int* __restrict__ perm, const int* __restrict__ rp,
                                                        const int* __restrict__ ci, const double* __restrict__ va,
                                                        const int* __restrict__ prp, int* __restrict__ pci, double* __restrict__ pva)
{
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5, GW = ((int64_t)gridDim.x * 256) >> 5;
    for (int64_t k = gw; k < nrow; k += GW) {
        const int i = perm[k];
        const int src = rp[i], n = rp[i + 1] - src, dst = prp[k];
        for (int e = lane; e < n; e += 32) {
            pci[dst + e] = ld_stream(ci + src + e);
            pva[dst + e] = ld_stream(va + src + e);
        }
    }
}

// ---- one colour of a sweep: a thread per row, the row's entries left to right ---------------------------------------
__global__ void __launch_bounds__(256) symgs_color_kernel(int count, const int* __restrict__ rows, const int* __restrict__ rp,
                                                          const int* __restrict__ ci, const double* __restrict__ va,
                                                          const double* __restrict__ diag, const double* __restrict__ r, double* x)
{
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= count) return;
    const int i = rows[k];
    double t = 0.0;
    const int e1 = __ldg(rp + i + 1);
    for (int p = __ldg(rp + i); p < e1; ++p) t = add_rn(t, mul_rn(__ldg(va + p), x[__ldg(ci + p)]));   // x of other colours only changes between launches
    const double d = diag[i];
    double s = add_rn(r[i], -t);
    s = add_rn(s, mul_rn(x[i], d));
    x[i] = __ddiv_rn(s, d);
}

// ---- CG pieces ------------------------------------------------------------------------------------------------------
// tile[t] = butterfly sum of a_i * b_i over rows [32 t, 32 t + 32): the dot product's half of tree_sum.cuh
__global__ void __launch_bounds__(256) tile_dot_kernel(int64_t n, const double* __restrict__ a, const double* __restrict__ b,
                                                       double* __restrict__ tile)
{
    const int lane = threadIdx.x & 31;
    const int64_t ntiles = (n + 31) >> 5;
    const int64_t gw = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5, GW = ((int64_t)gridDim.x * 256) >> 5;
    // four tiles per trip: eight loads in flight per lane (one tile per trip left the kernel waiting on latency: 3.6 TB/s)
    constexpr int U = 4;
    for (int64_t t0 = gw; t0 < ntiles; t0 += U * GW) {
        double va[U], vb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = (t0 + u * GW) * 32 + lane;
            const bool in = t0 + u * GW < ntiles && i < n;
            va[u] = in ? ld_stream(a + i) : 0.0;
            vb[u] = in ? ld_stream(b + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double q = warp_butterfly_sum(mul_rn(va[u], vb[u]));
            if (lane == 0 && t0 + u * GW < ntiles) tile[t0 + u * GW] = q;
        }
    }
}
// s[2] = s[0] / s[1]; s[3] = -s[2]         (alpha = rz / pAp)
__global__ void cg_alpha_kernel(double* s)
{
    const double a = __ddiv_rn(s[0], s[1]);
    s[2] = a;
    s[3] = -a;
}
// s[5] = s[4] / s[0]; s[0] = s[4]          (beta = rz_new / rz ; rz = rz_new)
__global__ void cg_beta_kernel(double* s)
{
    s[5] = __ddiv_rn(s[4], s[0]);
    s[0] = s[4];
}
// Vector::AddScaled (src/vector.cpp:98-128) with the scalar on the device: v += a x in the general form (a = +-1 give
// the same bits as the reference's special branches, a = 0 leaves v alone as there)
// Four consecutive elements per thread, 16-byte accesses when both arrays allow it (vec_ops.cu map_kernel: one element per
// access reaches 4-5 TB/s on these two-stream and in-place shapes, 16-byte accesses 6.5-6.9 TB/s).
template <class G>
__device__ __forceinline__ void map2_inplace(int64_t n, const double* __restrict__ x, double* __restrict__ v, G g)
{
    const bool vec = ((((uintptr_t)x) | ((uintptr_t)v)) & 15) == 0;
    const int64_t stride = (int64_t)gridDim.x * 256 * 4;
    for (int64_t base = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; base < n; base += stride) {
        if (vec && base + 4 <= n) {
            const double2 x0 = *reinterpret_cast<const double2*>(x + base), x1 = *reinterpret_cast<const double2*>(x + base + 2);
            const double2 v0 = *reinterpret_cast<const double2*>(v + base), v1 = *reinterpret_cast<const double2*>(v + base + 2);
            *reinterpret_cast<double2*>(v + base) = make_double2(g(v0.x, x0.x), g(v0.y, x0.y));
            *reinterpret_cast<double2*>(v + base + 2) = make_double2(g(v1.x, x1.x), g(v1.y, x1.y));
        } else {
            for (int k = 0; k < 4 && base + k < n; ++k) v[base + k] = g(v[base + k], x[base + k]);
        }
    }
}
__global__ void __launch_bounds__(256) add_scaled_dev_kernel(int64_t n, const double* __restrict__ a_dev, const double* __restrict__ x,
                                                             double* __restrict__ v)
{
    const double a = *a_dev;
    if (a == 0.0) return;
    map2_inplace(n, x, v, [a](double vi, double xi) { return add_rn(vi, mul_rn(a, xi)); });
}
// vec_axpby(1, z, beta, p, p): the alpha == 1 branch (src/vec_vec.cpp:54-61): w = beta y + x
__global__ void __launch_bounds__(256) xpby_dev_kernel(int64_t n, const double* __restrict__ z, const double* __restrict__ beta_dev,
                                                       double* __restrict__ p)
{
    const double b = *beta_dev;
    map2_inplace(n, z, p, [b](double pi, double zi) { return add_rn(mul_rn(b, pi), zi); });
}
__global__ void __launch_bounds__(256) jacobi_apply_kernel(int64_t n, const double* __restrict__ d, const double* __restrict__ r,
                                                           double* __restrict__ z)
{
    const bool vec = ((((uintptr_t)d) | ((uintptr_t)r) | ((uintptr_t)z)) & 15) == 0;
    const int64_t stride = (int64_t)gridDim.x * 256 * 4;
    for (int64_t base = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; base < n; base += stride) {
        if (vec && base + 4 <= n) {
            const double2 r0 = *reinterpret_cast<const double2*>(r + base), r1 = *reinterpret_cast<const double2*>(r + base + 2);
            const double2 d0 = *reinterpret_cast<const double2*>(d + base), d1 = *reinterpret_cast<const double2*>(d + base + 2);
            *reinterpret_cast<double2*>(z + base) = make_double2(__ddiv_rn(r0.x, d0.x), __ddiv_rn(r0.y, d0.y));
            *reinterpret_cast<double2*>(z + base + 2) = make_double2(__ddiv_rn(r1.x, d1.x), __ddiv_rn(r1.y, d1.y));
        } else {
            for (int k = 0; k < 4 && base + k < n; ++k) z[base + k] = __ddiv_rn(r[base + k], d[base + k]);
        }
    }
}

static inline int ew_blocks(int64_t n) { return (int)std::min<int64_t>((int64_t)sm_count() * 2048, std::max<int64_t>(1, (n + 1023) / 1024)); }

}  // namespace thsp

using namespace thsp;

struct thsp_symgs_plan {
    int nrow = 0, ncolors = 0, rounds = 0, natural_rounds = 0;
    int* perm = nullptr;            // device: rows grouped by colour, ascending inside a colour
    int* color = nullptr;           // device: colour of every row
    std::vector<int> color_ptr;     // host: [ncolors + 1]
    // The matrix permuted by colour (a snapshot taken when the plan is made: rows of a colour are one contiguous CSR
    // slab, which the TMA stream kernel of csr_spmv.cu sweeps with its Gauss-Seidel epilogue).  nullptr = a thread per
    // row on the caller's arrays.
    int nnz = 0;
    double mean_len = 0.0;
    int *prp = nullptr, *pci = nullptr;
    double* pva = nullptr;
};

extern "C" {

static int symgs_permute(thsp_symgs_plan* p, const int* row_ptr, const int* col_ind, const double* val, cudaStream_t s)
{
    const int nrow = p->nrow;
    int nnz = 0;
    THSP_CUDA(cudaMemcpyAsync(&nnz, row_ptr + nrow, sizeof(int), cudaMemcpyDeviceToHost, s));
    THSP_CUDA(cudaStreamSynchronize(s));
    if (nnz <= 0) return 0;
    int *len = nullptr, *mx = nullptr;
    THSP_CUDA(cudaMalloc(&len, sizeof(int) * ((size_t)nrow + 2)));
    mx = len + nrow + 1;
    THSP_CUDA(cudaMemsetAsync(mx, 0, sizeof(int), s));
    perm_len_kernel<<<div_up(nrow, 256), 256, 0, s>>>(nrow, p->perm, row_ptr, len, mx);
    THSP_LAUNCH_CHECK();
    int maxlen = 0;
    THSP_CUDA(cudaMemcpyAsync(&maxlen, mx, sizeof(int), cudaMemcpyDeviceToHost, s));
    THSP_CUDA(cudaStreamSynchronize(s));
    const double mean = (double)nnz / nrow;
    if (mean < 4.0 || maxlen > 2048 || nrow < 4096) {   // the stream kernel's own conditions (thsp_csr_plan_create)
        cudaFree(len);
        return 0;
    }
    THSP_CUDA(cudaMalloc(&p->prp, sizeof(int) * ((size_t)nrow + 1)));
    THSP_CUDA(cudaMalloc(&p->pci, sizeof(int) * (size_t)nnz));
    THSP_CUDA(cudaMalloc(&p->pva, sizeof(double) * (size_t)nnz));
    if (exclusive_scan(nrow, len, p->prp, s)) return 1;
    perm_copy_kernel<<<std::min(div_up((int64_t)nrow * 32, 256), sm_count() * 16), 256, 0, s>>>(nrow, p->perm, row_ptr, col_ind, val, p->prp,
                                                                                             p->pci, p->pva);
    THSP_LAUNCH_CHECK();
    THSP_CUDA(cudaStreamSynchronize(s));
    cudaFree(len);
    p->nnz = nnz;
    p->mean_len = mean;
    return 0;
}

int thsp_symgs_plan_create(thsp_symgs_plan** out, int nrow, const int* row_ptr, const int* col_ind, const double* val,
                           thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    THSP_REQUIRE(out != nullptr && nrow >= 0, "bad arguments");
    cudaStream_t s = as_stream(stream);
    thsp_symgs_plan* p = new thsp_symgs_plan();
    p->nrow = nrow;
    p->color_ptr.assign(1, 0);
    *out = p;
    if (nrow == 0) return 0;
    int *cin = nullptr, *cout = nullptr, *redo = nullptr, *small = nullptr;
    unsigned long long* forbid = nullptr;
    THSP_CUDA(cudaMalloc(&forbid, sizeof(unsigned long long) * (size_t)nrow));
    THSP_CUDA(cudaMemsetAsync(forbid, 0, sizeof(unsigned long long) * (size_t)nrow, s));
    THSP_CUDA(cudaMalloc(&cin, sizeof(int) * (size_t)nrow));
    THSP_CUDA(cudaMalloc(&cout, sizeof(int) * ((size_t)nrow + 1)));
    THSP_CUDA(cudaMalloc(&redo, sizeof(int) * ((size_t)nrow + 1)));
    THSP_CUDA(cudaMalloc(&small, sizeof(int) * 4));
    THSP_CUDA(cudaMemsetAsync(redo, 0, sizeof(int) * ((size_t)nrow + 1), s));
    const int grid = div_up(nrow, 256);
    int rc = 0;
    {   // natural-order greedy, level by level (level_round_kernel); what it does not reach stays at -1
        static const int env_cap = getenv("THSP_COLOR_NATURAL_ROUNDS") ? atoi(getenv("THSP_COLOR_NATURAL_ROUNDS")) : -1;
        const int cap = env_cap >= 0 ? env_cap : 16384;
        int *pending = nullptr, *lists = nullptr, *cnt = nullptr;
        THSP_CUDA(cudaMalloc(&pending, sizeof(int) * (size_t)nrow));
        THSP_CUDA(cudaMalloc(&lists, sizeof(int) * 2 * (size_t)nrow));
        THSP_CUDA(cudaMalloc(&cnt, sizeof(int) * 8));
        THSP_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * 8, s));
        bool gave_up = cap == 0;
        if (!gave_up) {
            level_init_kernel<<<grid, 256, 0, s>>>(nrow, row_ptr, col_ind, cin, pending, lists, cnt);
            THSP_LAUNCH_CHECK();
        } else {
            THSP_CUDA(cudaMemsetAsync(cin, 0xff, sizeof(int) * (size_t)nrow, s));
        }
        const int lgrid = std::min(grid, sm_count() * 4);
        for (int round = 0; round < cap && !gave_up; ++round) {
            int* lin = lists + (size_t)(round & 1) * nrow;
            int* lout = lists + (size_t)((round + 1) & 1) * nrow;
            level_round_kernel<<<lgrid, 256, 0, s>>>(nrow, row_ptr, col_ind, cin, pending, lin, lout, cnt, round, cnt + 4);
            THSP_LAUNCH_CHECK();
            p->natural_rounds = round + 1;
            if ((round & 31) == 31 || round + 1 == cap) {   // look at the counters every 32nd level only
                int h[5];
                THSP_CUDA(cudaMemcpyAsync(h, cnt, sizeof(h), cudaMemcpyDeviceToHost, s));
                THSP_CUDA(cudaStreamSynchronize(s));
                if (h[4]) gave_up = true;
                if (h[(round + 1) % 3] == 0) break;   // nobody is waiting for the next level
            }
        }
        cudaFree(pending);
        cudaFree(lists);
        cudaFree(cnt);
        if (gave_up) {
            THSP_CUDA(cudaMemsetAsync(cin, 0xff, sizeof(int) * (size_t)nrow, s));
            p->natural_rounds = 0;
        } else {
            color_verify_kernel<<<grid, 256, 0, s>>>(nrow, row_ptr, col_ind, cin, redo, forbid);
            THSP_LAUNCH_CHECK();
            THSP_CUDA(cudaMemcpyAsync(cout, cin, sizeof(int) * (size_t)nrow, cudaMemcpyDeviceToDevice, s));
            THSP_CUDA(cudaMemsetAsync(small, 0, sizeof(int) * 4, s));
            color_apply_kernel<<<grid, 256, 0, s>>>(nrow, cout, redo, cin, small);
            THSP_LAUNCH_CHECK();
        }
    }
    for (int round = 0; round < 200; ++round) {
        THSP_CUDA(cudaMemsetAsync(small, 0, sizeof(int) * 4, s));
        color_assign_kernel<<<grid, 256, 0, s>>>(nrow, row_ptr, col_ind, cin, cout, small + 1, forbid);
        THSP_LAUNCH_CHECK();
        color_resolve_kernel<<<grid, 256, 0, s>>>(nrow, row_ptr, col_ind, cin, cout, redo, forbid);
        THSP_LAUNCH_CHECK();
        color_apply_kernel<<<grid, 256, 0, s>>>(nrow, cout, redo, cin, small);
        THSP_LAUNCH_CHECK();
        int h[2] = {0, 0};
        THSP_CUDA(cudaMemcpyAsync(h, small, sizeof(h), cudaMemcpyDeviceToHost, s));
        THSP_CUDA(cudaStreamSynchronize(s));
        p->rounds = round + 1;
        if (h[1]) {
            set_error("SymGS colouring needs more than 64 colours (a row with more than 63 coloured neighbours)");
            rc = 2;
            break;
        }
        if (h[0] == 0) break;
        if (round == 199) {
            set_error("SymGS colouring did not settle in 200 rounds");
            rc = 2;
        }
    }
    if (!rc) {
        // rows of colour c, ascending: flag, scan, place - one colour at a time (a plan is built once per matrix)
        THSP_CUDA(cudaMalloc(&p->perm, sizeof(int) * (size_t)nrow));
        int base = 0;
        for (int c = 0; c < 64 && base < nrow && !rc; ++c) {
            color_flag_kernel<<<grid, 256, 0, s>>>(nrow, cin, c, redo);
            THSP_LAUNCH_CHECK();
            if (exclusive_scan(nrow, redo, cout, s)) { rc = 1; break; }
            int cnt = 0;
            THSP_CUDA(cudaMemcpyAsync(&cnt, cout + nrow, sizeof(int), cudaMemcpyDeviceToHost, s));
            THSP_CUDA(cudaStreamSynchronize(s));
            if (cnt > 0) {
                color_place_kernel<<<grid, 256, 0, s>>>(nrow, cin, c, cout, base, p->perm);
                THSP_LAUNCH_CHECK();
            }
            base += cnt;
            p->color_ptr.push_back(base);
            p->ncolors = c + 1;
        }
        THSP_CUDA(cudaStreamSynchronize(s));
        while (p->ncolors > 0 && p->color_ptr[p->ncolors] == p->color_ptr[p->ncolors - 1]) {   // trailing empty colours
            p->color_ptr.pop_back();
            --p->ncolors;
        }
        p->color = cin;
        cin = nullptr;
        if (val && symgs_permute(p, row_ptr, col_ind, val, s)) rc = 1;
    }
    if (cin) cudaFree(cin);
    cudaFree(forbid);
    cudaFree(cout);
    cudaFree(redo);
    cudaFree(small);
    if (rc) {
        thsp_symgs_plan_destroy(p);
        *out = nullptr;
    }
    return rc;
}

int thsp_symgs_plan_destroy(thsp_symgs_plan* p)
{
    if (!p) return 0;
    if (p->perm) cudaFree(p->perm);
    if (p->color) cudaFree(p->color);
    if (p->prp) cudaFree(p->prp);
    if (p->pci) cudaFree(p->pci);
    if (p->pva) cudaFree(p->pva);
    delete p;
    return 0;
}

int thsp_symgs_plan_streams(const thsp_symgs_plan* p, int* streams)
{
    THSP_REQUIRE(p != nullptr && streams != nullptr, "null plan");
    *streams = p->pva ? 1 : 0;
    return 0;
}

int thsp_symgs_plan_info(const thsp_symgs_plan* p, int* ncolors, int* rounds, int* color_ptr_host, int capacity, const int** perm_dev,
                         const int** color_dev)
{
    THSP_REQUIRE(p != nullptr, "null plan");
    if (ncolors) *ncolors = p->ncolors;
    if (rounds) *rounds = p->natural_rounds + p->rounds;
    if (color_ptr_host)
        for (int c = 0; c <= p->ncolors && c < capacity; ++c) color_ptr_host[c] = p->color_ptr[c];
    if (perm_dev) *perm_dev = p->perm;
    if (color_dev) *color_dev = p->color;
    return 0;
}

static int symgs_sweep(const thsp_symgs_plan* p, const int* rp, const int* ci, const double* va, const double* diag,
                       const double* r, double* x, cudaStream_t s)
{
    for (int pass = 0; pass < 2; ++pass)
        for (int k = 0; k < p->ncolors; ++k) {
            const int c = pass == 0 ? k : p->ncolors - 1 - k;
            const int off = p->color_ptr[c], cnt = p->color_ptr[c + 1] - off;
            if (cnt <= 0) continue;
            if (p->pva && cnt >= 2048) {   // the colour's rows are a contiguous slab of the permuted copy: stream it
                if (stream_gs_color(cnt, p->nnz, p->mean_len, p->prp + off, p->pci, p->pva, p->perm + off, r, diag, x, s)) return 1;
                continue;
            }
            symgs_color_kernel<<<div_up(cnt, 256), 256, 0, s>>>(cnt, p->perm + off, rp, ci, va, diag, r, x);
            THSP_LAUNCH_CHECK();
        }
    return 0;
}

int thsp_symgs_f64(const thsp_symgs_plan* plan, int nrow, const int* row_ptr, const int* col_ind, const double* val,
                   const double* diagonal, const double* r, double* x, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    THSP_REQUIRE(plan != nullptr && plan->nrow == nrow, "plan is null or was made for another matrix");
    THSP_REQUIRE(diagonal != nullptr, "SymGS divides by the diagonal array (include/matrix.h:36)");
    return symgs_sweep(plan, row_ptr, col_ind, val, diagonal, r, x, as_stream(stream));
}

// dot(a, b) in the canonical order into *out (device): tile partials, then the tree
static int dot_canonical(int64_t n, const double* a, const double* b, double* tiles, double* out, thsp_stream_t stream)
{
    const int64_t ntiles = (n + 31) / 32;
    const int grid = (int)std::min<int64_t>((int64_t)sm_count() * 8, std::max<int64_t>(1, (ntiles + 7) / 8));
    tile_dot_kernel<<<grid, 256, 0, as_stream(stream)>>>(n, a, b, tiles);
    THSP_LAUNCH_CHECK();
    return thsp_tree_sum_f64(ntiles, tiles, out, stream);
}

int thsp_dot_canonical_dev_f64(int64_t n, const double* x, const double* y, double* tile_scratch, double* out_dev, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    THSP_REQUIRE(tile_scratch != nullptr && out_dev != nullptr, "tile_scratch (ceil(n/32) doubles) and out_dev are required");
    if (n <= 0) return thsp_tree_sum_f64(0, tile_scratch, out_dev, stream);
    return dot_canonical(n, x, y, tile_scratch, out_dev, stream);
}

int64_t thsp_cg_work_doubles(int64_t n) { return 4 * n + (n + 31) / 32 + 16; }

int thsp_cg_f64(const thsp_csr_plan* A, int n, int precond, const thsp_symgs_plan* M, const int* row_ptr, const int* col_ind,
                const double* val, const double* diagonal, const double* b, double* x, int maxit, double tol, double* work,
                int* iters_out, double* relres_out, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    THSP_REQUIRE(A != nullptr && n > 0 && work != nullptr, "plan / size / work missing");
    THSP_REQUIRE(precond >= 0 && precond <= 2, "precond: 0 none, 1 Jacobi, 2 SymGS");
    THSP_REQUIRE(precond == 0 || diagonal != nullptr, "the preconditioners divide by the diagonal array");
    THSP_REQUIRE(precond != 2 || (M != nullptr && row_ptr && col_ind && val), "SymGS needs its plan and the matrix arrays");
    cudaStream_t s = as_stream(stream);
    const int64_t N = n;
    double* r = work;
    double* p = r + N;
    double* Ap = p + N;
    double* z = precond ? Ap + N : r;                 // no preconditioner: z is r
    double* tiles = work + 4 * N;
    double* sc = tiles + (N + 31) / 32;               // [0] rz  [1] pAp  [2] alpha  [3] -alpha  [4] rz_new  [5] beta  [6] rr  [7] bb
    const int eb = ew_blocks(N);
    auto precondition = [&]() -> int {
        if (precond == 1) {
            jacobi_apply_kernel<<<eb, 256, 0, s>>>(N, diagonal, r, z);
            THSP_LAUNCH_CHECK();
        } else if (precond == 2) {
            THSP_CUDA(cudaMemsetAsync(z, 0, sizeof(double) * (size_t)N, s));
            if (symgs_sweep(M, row_ptr, col_ind, val, diagonal, r, z, s)) return 1;
        }
        return 0;
    };
    double h[2] = {0.0, 0.0};
    // r = b - A x  (y = A x, then vec_axpby(1, b, -1, y): the alpha == 1 branch gives  -1*y + b)
    if (thsp_csr_plan_spmv_f64(A, x, r, 0, stream)) return 1;
    if (thsp_axpby_f64(N, 1.0, b, -1.0, r, r, stream)) return 1;
    if (dot_canonical(N, b, b, tiles, sc + 7, stream)) return 1;
    if (dot_canonical(N, r, r, tiles, sc + 6, stream)) return 1;
    if (precondition()) return 1;
    THSP_CUDA(cudaMemcpyAsync(p, z, sizeof(double) * (size_t)N, cudaMemcpyDeviceToDevice, s));
    if (precond) {
        if (dot_canonical(N, r, z, tiles, sc + 0, stream)) return 1;
    } else {
        THSP_CUDA(cudaMemcpyAsync(sc + 0, sc + 6, sizeof(double), cudaMemcpyDeviceToDevice, s));
    }
    THSP_CUDA(cudaMemcpyAsync(h, sc + 6, sizeof(h), cudaMemcpyDeviceToHost, s));
    THSP_CUDA(cudaStreamSynchronize(s));
    const double bnorm = h[1] > 0.0 ? sqrt(h[1]) : 1.0;
    double rel = sqrt(h[0]) / bnorm;
    int it = 0;
    while (it < maxit && rel > tol) {
        if (thsp_csr_plan_spmv_f64(A, p, Ap, 0, stream)) return 1;
        if (dot_canonical(N, p, Ap, tiles, sc + 1, stream)) return 1;
        cg_alpha_kernel<<<1, 1, 0, s>>>(sc);
        THSP_LAUNCH_CHECK();
        add_scaled_dev_kernel<<<eb, 256, 0, s>>>(N, sc + 2, p, x);     // x += alpha p
        THSP_LAUNCH_CHECK();
        add_scaled_dev_kernel<<<eb, 256, 0, s>>>(N, sc + 3, Ap, r);    // r += (-alpha) A p
        THSP_LAUNCH_CHECK();
        if (dot_canonical(N, r, r, tiles, sc + 6, stream)) return 1;
        if (precondition()) return 1;
        if (precond) {
            if (dot_canonical(N, r, z, tiles, sc + 4, stream)) return 1;
        } else {
            THSP_CUDA(cudaMemcpyAsync(sc + 4, sc + 6, sizeof(double), cudaMemcpyDeviceToDevice, s));
        }
        cg_beta_kernel<<<1, 1, 0, s>>>(sc);
        THSP_LAUNCH_CHECK();
        xpby_dev_kernel<<<eb, 256, 0, s>>>(N, z, sc + 5, p);           // p = beta p + z
        THSP_LAUNCH_CHECK();
        THSP_CUDA(cudaMemcpyAsync(h, sc + 6, sizeof(double), cudaMemcpyDeviceToHost, s));
        THSP_CUDA(cudaStreamSynchronize(s));
        rel = sqrt(h[0]) / bnorm;
        ++it;
    }
    if (iters_out) *iters_out = it;
    if (relres_out) *relres_out = rel;
    return 0;
}

}  // extern "C"
