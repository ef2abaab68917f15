// generate.cu -- device-side synthetic inputs and small utilities.
//
// The big configurations cannot come from a Matrix Market file (256^3 would be 11 GB of text,
// 512^3 does not fit the reference's int32 structs at all), so matrices are generated straight
// into device memory.  oracle/oracle.c holds CPU twins that produce identical arrays, which is
// how parity tests get the same matrix on both sides.
#include <algorithm>

#include "common.cuh"

namespace thsp {

__host__ __device__ inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ inline double u01(uint64_t bits) { return (double)(bits >> 11) * (1.0 / 9007199254740992.0); }

// ---- 27-point stencil on an n^3 grid ---------------------------------------------------
// 1-D neighbour count f(i) = 3 - [i==0] - [i==n-1]; prefix F(i) = sum_{i'<i} f(i').
// Entries before row (z,y,x) in lexicographic order: F(z) T^2 + f(z) (F(y) T + f(y) F(x)), T = 3n-2.
__host__ __device__ inline int64_t st_f(int i, int n) { return 3 - (i == 0) - (i == n - 1); }
__host__ __device__ inline int64_t st_F(int i, int n) { return 3 * (int64_t)i - (i >= 1) - (i >= n); }
__host__ __device__ inline int64_t stencil_prefix(int64_t row, int n)
{
    const int64_t T = 3 * (int64_t)n - 2;
    const int64_t nn = (int64_t)n * n;
    if (row >= nn * n) return T * T * T;
    const int x = (int)(row % n), y = (int)((row / n) % n), z = (int)(row / nn);
    return st_F(z, n) * T * T + st_f(z, n) * (st_F(y, n) * T + st_f(y, n) * st_F(x, n));
}

__global__ void __launch_bounds__(256) stencil_csr_kernel(int n, int64_t r0, int64_t r1, int64_t base, int* __restrict__ row_ptr,
                                                          int* __restrict__ col, double* __restrict__ val)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t r = r0 + i;
    if (r > r1) return;
    int64_t p = stencil_prefix(r, n) - base;
    row_ptr[i] = (int)p;
    if (r == r1 || col == nullptr) return;
    const int x = (int)(r % n), y = (int)((r / n) % n), z = (int)(r / ((int64_t)n * n));
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx, yy = y + dy, zz = z + dz;
                if (xx < 0 || yy < 0 || zz < 0 || xx >= n || yy >= n || zz >= n) continue;
                const int64_t c = ((int64_t)zz * n + yy) * n + xx;
                col[p] = (int)c;
                val[p] = (c == r) ? 26.0 : -1.0;
                ++p;
            }
}

// Column-major ELL slab of width 27; slot = rank among the in-grid neighbours; padding (0, 0.0).
__global__ void __launch_bounds__(256) stencil_ell_kernel(int n, int64_t nrow, int* __restrict__ col, double* __restrict__ val)
{
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= nrow) return;
    const int x = (int)(r % n), y = (int)((r / n) % n), z = (int)(r / ((int64_t)n * n));
    int slot = 0;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx, yy = y + dy, zz = z + dz;
                if (xx < 0 || yy < 0 || zz < 0 || xx >= n || yy >= n || zz >= n) continue;
                const int64_t c = ((int64_t)zz * n + yy) * n + xx;
                col[(size_t)slot * nrow + r] = (int)c;
                val[(size_t)slot * nrow + r] = (c == r) ? 26.0 : -1.0;
                ++slot;
            }
    for (; slot < 27; ++slot) {
        col[(size_t)slot * nrow + r] = 0;
        val[(size_t)slot * nrow + r] = 0.0;
    }
}

__global__ void __launch_bounds__(256) stencil_coo_kernel(int n, int64_t nrow, int* __restrict__ row, int* __restrict__ col,
                                                          double* __restrict__ val)
{
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= nrow) return;
    int64_t p = stencil_prefix(r, n);
    const int x = (int)(r % n), y = (int)((r / n) % n), z = (int)(r / ((int64_t)n * n));
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx, yy = y + dy, zz = z + dz;
                if (xx < 0 || yy < 0 || zz < 0 || xx >= n || yy >= n || zz >= n) continue;
                const int64_t c = ((int64_t)zz * n + yy) * n + xx;
                row[p] = (int)r;
                col[p] = (int)c;
                val[p] = (c == r) ? 26.0 : -1.0;
                ++p;
            }
}

// ---- 5-point Laplacian on n x n, COO, entries in (N,W,C,E,S) order ------------------------
// entries before grid row i: 5 n i - 2 i - (n... ) computed directly: row i has n points, each
// 5 minus the missing neighbours.  prefix(i,j) = sum over earlier points.
__host__ __device__ inline int64_t lap5_prefix(int i, int j, int n)
{
    // full rows before i: each row has 5n - 2 (W/E ends) entries, minus n for the first row (no N)
    // and minus n for the last row (no S).
    int64_t p = (int64_t)i * (5 * (int64_t)n - 2);
    if (i > 0) p -= n;               // row 0 lacks N
    // (the last row is never "before" any row)
    // points before j in row i: 5 each, minus 1 for j'=0 (no W), minus N/S absences
    int64_t per = 5 - (i == 0) - (i == n - 1);
    p += (int64_t)j * per - (j > 0 ? 1 : 0);
    return p;
}
__global__ void __launch_bounds__(256) lap5_coo_kernel(int n, int* __restrict__ row, int* __restrict__ col, double* __restrict__ val)
{
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= (int64_t)n * n) return;
    const int i = (int)(t / n), j = (int)(t % n);
    const int r = (int)t;
    int64_t p = lap5_prefix(i, j, n);
    if (i > 0)     { row[p] = r; col[p] = r - n; val[p] = -1.0; ++p; }
    if (j > 0)     { row[p] = r; col[p] = r - 1; val[p] = -1.0; ++p; }
                   { row[p] = r; col[p] = r;     val[p] = 4.0;  ++p; }
    if (j < n - 1) { row[p] = r; col[p] = r + 1; val[p] = -1.0; ++p; }
    if (i < n - 1) { row[p] = r; col[p] = r + n; val[p] = -1.0; ++p; }
}

__global__ void __launch_bounds__(256) uniform_coo_kernel(int nrow, int ncol, int64_t nnz, uint64_t seed, int* __restrict__ row,
                                                          int* __restrict__ col, double* __restrict__ val)
{
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (k >= nnz) return;
    const uint64_t h = mix64(seed * 0xD1342543DE82EF95ull + (uint64_t)k);
    const uint64_t g = mix64(h);
    row[k] = (int)(((h >> 32) * (uint64_t)nrow) >> 32);
    col[k] = (int)(((h & 0xFFFFFFFFull) * (uint64_t)ncol) >> 32);
    val[k] = u01(g);
}

__global__ void __launch_bounds__(256) rmat_coo_kernel(int scale, int64_t nnz, uint64_t seed, int* __restrict__ row,
                                                       int* __restrict__ col, double* __restrict__ val)
{
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (k >= nnz) return;
    uint64_t s = mix64(seed * 0xD1342543DE82EF95ull + (uint64_t)k);
    int r = 0, c = 0;
    for (int lvl = 0; lvl < scale; ++lvl) {
        s = mix64(s);
        const double u = u01(s);
        int rb, cb;
        if (u < 0.57) { rb = 0; cb = 0; }
        else if (u < 0.76) { rb = 0; cb = 1; }
        else if (u < 0.95) { rb = 1; cb = 0; }
        else { rb = 1; cb = 1; }
        r = (r << 1) | rb;
        c = (c << 1) | cb;
    }
    row[k] = r;
    col[k] = c;
    val[k] = u01(mix64(s));
}

__global__ void __launch_bounds__(256) vector_kernel(int64_t n, uint64_t seed, double* __restrict__ v)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) v[i] = u01(mix64(seed * 0xD1342543DE82EF95ull + (uint64_t)i));
}
__global__ void __launch_bounds__(256) to_f32_kernel(int64_t n, const double* __restrict__ s, float* __restrict__ d)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) d[i] = (float)s[i];
}
__global__ void __launch_bounds__(256) flush_kernel(size_t n16, int4* __restrict__ p, int tag)
{
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += stride) p[i] = make_int4(tag, tag, tag, tag);
}
// Order-independent 64-bit fingerprint of a piece of a vector: sum over i of mix64(bits(v_i) ^ mix64(first + i)) modulo
// 2^64.  Pieces of one vector add up to the fingerprint of the whole, whoever holds them.
__global__ void __launch_bounds__(256) hash_kernel(int64_t n, const double* __restrict__ v, uint64_t first, unsigned long long* __restrict__ out)
{
    unsigned long long h = 0;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride)
        h += mix64((uint64_t)__double_as_longlong(v[i]) ^ mix64(first + (uint64_t)i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0 && h) atomicAdd(out, h);
}
__global__ void __launch_bounds__(256) slice_row_ptr_kernel(const int* __restrict__ rp, int start, int count, int* __restrict__ sub)
{
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j <= count) sub[j] = rp[start + j] - rp[start];
}

}  // namespace thsp

using namespace thsp;

extern "C" {

int64_t thsp_stencil27_nnz(int n, int64_t row_begin, int64_t row_end)
{
    return stencil_prefix(row_end, n) - stencil_prefix(row_begin, n);
}
int64_t thsp_lap5_nnz(int n) { return 5 * (int64_t)n * n - 4 * (int64_t)n; }

int thsp_gen_stencil27_csr(int n, int64_t row_begin, int64_t row_end, int* row_ptr, int* col_ind, double* val,
                           thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    const int64_t N = (int64_t)n * n * n;
    THSP_REQUIRE(n >= 1 && row_begin >= 0 && row_begin <= row_end && row_end <= N, "bad stencil row range");
    const int64_t sub = thsp_stencil27_nnz(n, row_begin, row_end);
    THSP_REQUIRE(sub <= 0x7fffffffLL, "row block holds more than 2^31-1 entries; use more blocks");
    const int64_t rows = row_end - row_begin;
    stencil_csr_kernel<<<div_up(rows + 1, 256), 256, 0, as_stream(stream)>>>(n, row_begin, row_end,
                                                                            stencil_prefix(row_begin, n), row_ptr, col_ind, val);
    THSP_LAUNCH_CHECK();
    return 0;
}
int thsp_gen_stencil27_ell(int n, int* col_ind, double* val, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    const int64_t N = (int64_t)n * n * n;
    THSP_REQUIRE(N <= 0x7fffffffLL, "grid too large for int32 rows");
    stencil_ell_kernel<<<div_up(N, 256), 256, 0, as_stream(stream)>>>(n, N, col_ind, val);
    THSP_LAUNCH_CHECK();
    return 0;
}
int thsp_gen_stencil27_coo(int n, int* row_ind, int* col_ind, double* val, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    const int64_t N = (int64_t)n * n * n;
    THSP_REQUIRE(thsp_stencil27_nnz(n, 0, N) <= 0x7fffffffLL, "more than 2^31-1 entries");
    stencil_coo_kernel<<<div_up(N, 256), 256, 0, as_stream(stream)>>>(n, N, row_ind, col_ind, val);
    THSP_LAUNCH_CHECK();
    return 0;
}
int thsp_gen_lap5_coo(int n, int* row_ind, int* col_ind, double* val, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    THSP_REQUIRE(n >= 2, "lap5 needs n >= 2");
    lap5_coo_kernel<<<div_up((int64_t)n * n, 256), 256, 0, as_stream(stream)>>>(n, row_ind, col_ind, val);
    THSP_LAUNCH_CHECK();
    return 0;
}
int thsp_gen_uniform_coo(int nrow, int ncol, int64_t nnz, uint64_t seed, int* row_ind, int* col_ind, double* val,
                         thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    if (nnz <= 0) return 0;
    uniform_coo_kernel<<<div_up(nnz, 256), 256, 0, as_stream(stream)>>>(nrow, ncol, nnz, seed, row_ind, col_ind, val);
    THSP_LAUNCH_CHECK();
    return 0;
}
int thsp_gen_rmat_coo(int scale, int64_t nnz, uint64_t seed, int* row_ind, int* col_ind, double* val, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    THSP_REQUIRE(scale >= 1 && scale <= 30, "rmat scale must be in 1..30");
    if (nnz <= 0) return 0;
    rmat_coo_kernel<<<div_up(nnz, 256), 256, 0, as_stream(stream)>>>(scale, nnz, seed, row_ind, col_ind, val);
    THSP_LAUNCH_CHECK();
    return 0;
}
int thsp_gen_vector_f64(int64_t n, uint64_t seed, double* v, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    if (n <= 0) return 0;
    vector_kernel<<<div_up(n, 256), 256, 0, as_stream(stream)>>>(n, seed, v);
    THSP_LAUNCH_CHECK();
    return 0;
}
int thsp_f64_to_f32(int64_t n, const double* src, float* dst, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    if (n <= 0) return 0;
    to_f32_kernel<<<div_up(n, 256), 256, 0, as_stream(stream)>>>(n, src, dst);
    THSP_LAUNCH_CHECK();
    return 0;
}
int thsp_hash_f64(int64_t n, const double* v, uint64_t first_index, uint64_t* hash_host, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    cudaStream_t s = as_stream(stream);
    unsigned long long* d = static_cast<unsigned long long*>(scratch(sizeof(unsigned long long), 1));
    if (!d) return 1;
    THSP_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), s));
    if (n > 0) {
        hash_kernel<<<(int)std::min<int64_t>((int64_t)sm_count() * 8, (n + 255) / 256), 256, 0, s>>>(n, v, first_index, d);
        THSP_LAUNCH_CHECK();
    }
    THSP_CUDA(cudaMemcpyAsync(hash_host, d, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    THSP_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int thsp_flush_l2(void* scratch_buf, size_t bytes, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    static int tag = 0;
    flush_kernel<<<sm_count() * 8, 256, 0, as_stream(stream)>>>(bytes / 16, static_cast<int4*>(scratch_buf), ++tag);
    THSP_LAUNCH_CHECK();
    return 0;
}

int thsp_partition_rows(int64_t nrow, int nparts, int part, int64_t* start, int64_t* count)
{
    THSP_REQUIRE(nparts >= 1 && part >= 0 && part < nparts, "bad partition index");
    const int64_t per = nrow / nparts;  // src/mat_vec.cpp:233
    *start = (int64_t)part * per;
    *count = (part == nparts - 1) ? nrow - *start : per;  // :245-246 last block takes the remainder
    return 0;
}
int thsp_csr_slice_row_ptr(const int* row_ptr, int start, int count, int* sub_row_ptr, thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    slice_row_ptr_kernel<<<div_up(count + 1, 256), 256, 0, as_stream(stream)>>>(row_ptr, start, count, sub_row_ptr);
    THSP_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
