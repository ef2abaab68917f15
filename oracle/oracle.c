/*
 * oracle.c -- CPU restatement of the arm-spmv hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker for the CUDA library, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  Nothing under arm-spmv_b200/ links, imports or calls it.
 *
 * Parity status: PINNED.  Every function below is compared (tests/test_oracle.py)
 * against outputs of the unmodified reference sources compiled from
 * /root/reference by oracle/Makefile (oracle/_ref/libref.so) -- both live in the
 * build container and through the committed fixtures in tests/golden/ -- and
 * against the hand-checked known-answer vector of SURVEY.md appendix A.1.
 *
 * All arithmetic is written as separate IEEE multiply and add (this file is
 * compiled with -ffp-contract=off) because the reference is built by
 * `g++ -O2` on x86-64 without -march, which never fuses.
 * Loops are serial: that is the reference's order with OMP_NUM_THREADS=1, the
 * only order in which its COO/CSC atomics are deterministic.
 *
 * Each function cites the reference lines it follows (paths under /root/reference).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ SpMV -- */

/* src/mat_vec.cpp:18-42  COOMatirxMatVector: one y[row] += v*x[col] per stored entry, k ascending. */
ORACLE_API void oracle_coo_spmv(int nnz, const int *ri, const int *ci, const double *v,
                                const double *x, double *y)
{
    for (int k = 0; k < nnz; ++k) {
        double p = v[k] * x[ci[k]];
        y[ri[k]] = y[ri[k]] + p;
    }
}

/* src/mat_vec.cpp:44-67  CSRMatrixMatVector: per row, s starts at 0.0, entries added
 * left to right in stored order, then a single y[i] += s. */
ORACLE_API void oracle_csr_spmv(int nrow, const int *rp, const int *ci, const double *v,
                                const double *x, double *y)
{
    for (int r = 0; r < nrow; ++r) {
        double s = 0.0;
        for (int p = rp[r]; p < rp[r + 1]; ++p) {
            double t = v[p] * x[ci[p]];
            s = s + t;
        }
        y[r] = y[r] + s;
    }
}

/* fp32 extension (no reference counterpart, SURVEY.md 8(c) "fp32"): same order in float. */
ORACLE_API void oracle_csr_spmv_f32(int nrow, const int *rp, const int *ci, const float *v,
                                    const float *x, float *y)
{
    for (int r = 0; r < nrow; ++r) {
        float s = 0.0f;
        for (int p = rp[r]; p < rp[r + 1]; ++p) {
            float t = v[p] * x[ci[p]];
            s = s + t;
        }
        y[r] = y[r] + s;
    }
}

/* src/mat_vec.cpp:69-95  CSCMatrixMatVector: columns ascending, entries of a column in
 * stored order, y[row] += v*x[col] each. */
ORACLE_API void oracle_csc_spmv(int ncol, const int *cp, const int *ri, const double *v,
                                const double *x, double *y)
{
    for (int c = 0; c < ncol; ++c) {
        for (int p = cp[c]; p < cp[c + 1]; ++p) {
            double t = v[p] * x[c];
            y[ri[p]] = y[ri[p]] + t;
        }
    }
}

/* src/mat_vec.cpp:97-121  ELLMatrixMatVector: slot-major sweep over the column-major slab,
 * accumulating straight into y (y's old value is the first addend).  Padding slots hold
 * (col 0, 0.0) and are multiplied like any other. */
ORACLE_API void oracle_ell_spmv(int nrow, int width, const int *ci, const double *v,
                                const double *x, double *y)
{
    for (int s = 0; s < width; ++s) {
        const int *cs = ci + (size_t)s * nrow;
        const double *vs = v + (size_t)s * nrow;
        for (int r = 0; r < nrow; ++r) {
            double t = vs[r] * x[cs[r]];
            y[r] = y[r] + t;
        }
    }
}

/* fp32 extension of the same sweep (no reference counterpart, SURVEY.md 8(c) "fp32"): same order in float. */
ORACLE_API void oracle_ell_spmv_f32(int nrow, int width, const int *ci, const float *v,
                                    const float *x, float *y)
{
    for (int s = 0; s < width; ++s) {
        const int *cs = ci + (size_t)s * nrow;
        const float *vs = v + (size_t)s * nrow;
        for (int r = 0; r < nrow; ++r) {
            float t = vs[r] * x[cs[r]];
            y[r] = y[r] + t;
        }
    }
}

/* src/mat_vec.cpp:123-146  DIAMatrixMatVector: row-major diagonals, d ascending, accumulate
 * into y; the column guard compares against nrow (not ncol) exactly as the reference does. */
ORACLE_API void oracle_dia_spmv(int nrow, int ndiags, const int *off, const double *v,
                                const double *x, double *y)
{
    for (int r = 0; r < nrow; ++r) {
        for (int d = 0; d < ndiags; ++d) {
            int c = r + off[d];
            if (c >= 0 && c < nrow) {
                double t = v[(size_t)r * ndiags + d] * x[c];
                y[r] = y[r] + t;
            }
        }
    }
}

/* ----------------------------------------------------------- conversions -- */

/* Shared by CSR (key=row) and CSC (key=col): histogram at index key, running sum so that
 * ptr[i] is the END of bucket i, then a backward pass that pre-decrements the bucket cursor.
 * The net effect is a stable counting sort.  src/matrix.cpp:125-144 and :305-324. */
static void stable_bucket(int nbuckets, int nnz, const int *key, const int *other,
                          const double *val, int *ptr, int *other_out, double *val_out)
{
    for (int i = 0; i <= nbuckets; ++i) ptr[i] = 0;
    for (int k = 0; k < nnz; ++k) ptr[key[k]] += 1;
    for (int i = 0; i < nbuckets; ++i) ptr[i + 1] += ptr[i];
    for (int k = nnz - 1; k >= 0; --k) {
        int slot = --ptr[key[k]];
        other_out[slot] = other[k];
        val_out[slot] = val[k];
    }
}

/* Packed diagonal: values of entries with row==col, in COO order; returns how many.
 * src/matrix.cpp:146-153 (CSR) and :491-499 (ELL).  The reference writes past diagonal[nrow-1]
 * when duplicates make the count exceed nrow; the oracle stops at `cap`. */
static int pack_diagonal(int nnz, const int *ri, const int *ci, const double *v, double *diag, int cap)
{
    int n = 0;
    for (int k = 0; k < nnz; ++k)
        if (ri[k] == ci[k]) {
            if (n < cap) diag[n] = v[k];
            ++n;
        }
    return n;
}

/* src/matrix.cpp:115-154  CSRMatrix::CSRMatrix(const COOMatrix&). */
ORACLE_API int oracle_coo2csr(int nrow, int ncol, int nnz, const int *ri, const int *ci, const double *v,
                              int *row_ptr, int *col_ind, double *values, double *diagonal)
{
    (void)ncol;
    stable_bucket(nrow, nnz, ri, ci, v, row_ptr, col_ind, values);
    return diagonal ? pack_diagonal(nnz, ri, ci, v, diagonal, nrow) : 0;
}

/* src/matrix.cpp:295-325  CSCMatrix::CSCMatrix(const COOMatrix&). */
ORACLE_API void oracle_coo2csc(int nrow, int ncol, int nnz, const int *ri, const int *ci, const double *v,
                               int *col_ptr, int *row_ind, double *values)
{
    (void)nrow;
    stable_bucket(ncol, nnz, ci, ri, v, col_ptr, row_ind, values);
}

/* src/matrix.cpp:456-469  width of the ELL slab = longest row (duplicates counted). */
ORACLE_API int oracle_coo2ell_width(int nrow, int nnz, const int *ri)
{
    int *cnt = (int *)calloc((size_t)(nrow > 0 ? nrow : 1), sizeof(int));
    int w = 0;
    for (int k = 0; k < nnz; ++k) cnt[ri[k]] += 1;
    for (int r = 0; r < nrow; ++r) if (cnt[r] > w) w = cnt[r];
    free(cnt);
    return w;
}

/* src/matrix.cpp:471-499  zero-filled column-major slab, backward fill with slot = --count[row]
 * (so slots keep COO order), then the packed diagonal. */
ORACLE_API int oracle_coo2ell(int nrow, int ncol, int nnz, const int *ri, const int *ci, const double *v,
                              int width, int *col_ind, double *values, double *diagonal)
{
    (void)ncol;
    int *cnt = (int *)calloc((size_t)(nrow > 0 ? nrow : 1), sizeof(int));
    size_t total = (size_t)nrow * (size_t)width;
    for (size_t i = 0; i < total; ++i) { col_ind[i] = 0; values[i] = 0.0; }
    for (int k = 0; k < nnz; ++k) cnt[ri[k]] += 1;
    for (int k = nnz - 1; k >= 0; --k) {
        int r = ri[k];
        int slot = --cnt[r];
        col_ind[(size_t)slot * nrow + r] = ci[k];
        values[(size_t)slot * nrow + r] = v[k];
    }
    free(cnt);
    return diagonal ? pack_diagonal(nnz, ri, ci, v, diagonal, nrow) : 0;
}

/* src/matrix.cpp:673-693  which diagonals are occupied.  The reference indexes its map with
 * (nrow - i + j) in [1, nrow+ncol-1] but sizes and scans it as [0, nrow+ncol-1), so the
 * top-right corner diagonal (0, ncol-1) is written out of bounds and never emitted
 * (SURVEY.md A.3).  The oracle reproduces the visible result: that diagonal is dropped from
 * `offsets` (and counted nowhere). Returns ndiags; offsets may be NULL to just count. */
ORACLE_API int oracle_csr2dia_offsets(int nrow, int ncol, const int *rp, const int *ci, int *offsets)
{
    int span = nrow + ncol - 1;
    if (span <= 0) return 0;
    unsigned char *seen = (unsigned char *)calloc((size_t)span + 1, 1);
    for (int r = 0; r < nrow; ++r)
        for (int p = rp[r]; p < rp[r + 1]; ++p) seen[nrow - r + ci[p]] = 1;
    int nd = 0;
    for (int m = 0; m < span; ++m)
        if (seen[m]) {
            if (offsets) offsets[nd] = m - nrow;
            ++nd;
        }
    free(seen);
    return nd;
}

/* src/matrix.cpp:695-725  row-major fill, values[i*ndiags+d]; a duplicate (i,j) overwrites
 * the earlier one (last stored entry wins).  Entries on the dropped corner diagonal are
 * skipped (the reference reads an out-of-bounds map slot there). */
ORACLE_API void oracle_csr2dia_fill(int nrow, int ncol, const int *rp, const int *ci, const double *v,
                                    int ndiags, const int *offsets, double *values)
{
    int span = nrow + ncol - 1;
    int *slot = (int *)malloc(((size_t)(span > 0 ? span : 0) + 1) * sizeof(int));
    for (int m = 0; m <= span; ++m) slot[m] = -1;
    for (int d = 0; d < ndiags; ++d) slot[offsets[d] + nrow] = d;
    for (size_t i = 0; i < (size_t)nrow * (size_t)ndiags; ++i) values[i] = 0.0;
    for (int r = 0; r < nrow; ++r)
        for (int p = rp[r]; p < rp[r + 1]; ++p) {
            int m = nrow - r + ci[p];
            if (m < span && slot[m] >= 0) values[(size_t)r * ndiags + slot[m]] = v[p];
        }
    free(slot);
}

/* --------------------------------------------------------- vector kernels -- */

/* src/vec_vec.cpp:15-29  vec_dot, serial order (OpenMP's reduction order is unspecified). */
ORACLE_API double oracle_dot(int n, const double *x, const double *y)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        double t = x[i] * y[i];
        s = s + t;
    }
    return s;
}

/* src/vec_vec.cpp:31-94  vec_axpby: w = alpha*x + beta*y with the reference's seven-way
 * dispatch on alpha/beta; the branch decides which multiplies exist, hence the rounding. */
ORACLE_API void oracle_axpby(int n, double alpha, const double *x, double beta, const double *y, double *w)
{
    if (alpha == 0)       for (int i = 0; i < n; ++i) w[i] = beta * y[i];
    else if (beta == 0)   for (int i = 0; i < n; ++i) w[i] = alpha * x[i];
    else if (alpha == 1)  for (int i = 0; i < n; ++i) { double t = beta * y[i];  w[i] = t + x[i]; }
    else if (alpha == -1) for (int i = 0; i < n; ++i) { double t = beta * y[i];  w[i] = t - x[i]; }
    else if (beta == 1)   for (int i = 0; i < n; ++i) { double t = alpha * x[i]; w[i] = t + y[i]; }
    else if (beta == -1)  for (int i = 0; i < n; ++i) { double t = alpha * x[i]; w[i] = t - y[i]; }
    else for (int i = 0; i < n; ++i) { double a = alpha * x[i]; double b = beta * y[i]; w[i] = a + b; }
}

/* src/vector.cpp:59-63 Fill, :78-86 Scale, :88-96 Shift, :71-76 Copy. */
ORACLE_API void oracle_fill(int n, double a, double *v)  { for (int i = 0; i < n; ++i) v[i] = a; }
ORACLE_API void oracle_scale(int n, double a, double *v) { for (int i = 0; i < n; ++i) v[i] = v[i] * a; }
ORACLE_API void oracle_shift(int n, double a, double *v) { for (int i = 0; i < n; ++i) v[i] = v[i] + a; }
ORACLE_API void oracle_copy(int n, const double *x, double *v) { for (int i = 0; i < n; ++i) v[i] = x[i]; }

/* src/vector.cpp:98-128  AddScaled: v += a*x with a in {0,1,-1} special-cased. */
ORACLE_API void oracle_add_scaled(int n, double a, const double *x, double *v)
{
    if (a == 0) return;
    if (a == 1)       for (int i = 0; i < n; ++i) v[i] = v[i] + x[i];
    else if (a == -1) for (int i = 0; i < n; ++i) v[i] = v[i] - x[i];
    else for (int i = 0; i < n; ++i) { double t = a * x[i]; v[i] = v[i] + t; }
}

/* src/vector.cpp:130-159  Add2Scaled: v += a*x + b*y; (a*x + b*y) is formed first. */
ORACLE_API void oracle_add2_scaled(int n, double a, const double *x, double b, const double *y, double *v)
{
    if (a == 0)      oracle_add_scaled(n, b, y, v);
    else if (b == 0) oracle_add_scaled(n, a, x, v);
    else if (a == 1) for (int i = 0; i < n; ++i) { double t = b * y[i]; double u = x[i] + t; v[i] = v[i] + u; }
    else if (b == 1) for (int i = 0; i < n; ++i) { double t = a * x[i]; double u = t + y[i]; v[i] = v[i] + u; }
    else for (int i = 0; i < n; ++i) { double t = a * x[i]; double s = b * y[i]; double u = t + s; v[i] = v[i] + u; }
}

/* src/vector.cpp:161-171  checkVector: same length and every |x-y| <= 1e-6. */
ORACLE_API int oracle_check_vector(int nx, const double *x, int ny, const double *y)
{
    if (nx != ny) return 0;
    for (int i = 0; i < nx; ++i) if (fabs(x[i] - y[i]) > 1e-6) return 0;
    return 1;
}

/* src/vector.cpp:65-69  FillRandom: (double)rand()/RAND_MAX, sequential glibc stream. */
ORACLE_API void oracle_fill_random(int n, double *v)
{
    for (int i = 0; i < n; ++i) v[i] = (double)rand() / RAND_MAX;
}
ORACLE_API void oracle_srand(unsigned s) { srand(s); }

/* ------------------------------------------------------------- partition -- */

/* src/mat_vec.cpp:233,244-246  equal ROW blocks, the last block takes the remainder. */
ORACLE_API void oracle_partition(int n, int nparts, int part, int *start, int *count)
{
    int per = n / nparts;
    *start = part * per;
    *count = (part == nparts - 1) ? (n - *start) : per;
}

/* src/mat_vec.cpp:248-263  the block's row_ptr rebased so that it starts at 0. */
ORACLE_API int oracle_csr_slice(const int *rp, int start, int count, int *sub_rp)
{
    int base = rp[start];
    for (int j = 0; j <= count; ++j) sub_rp[j] = rp[start + j] - base;
    return rp[start + count] - base;
}

/* ------------------------------------------------------ synthetic inputs -- */
/* SURVEY.md 8(d) "Synthetic inputs".  These mirror the device generators in
 * arm-spmv_b200/csrc/generate.cu so that parity can be checked on the same matrices. */

static inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t bits) { return (double)(bits >> 11) * (1.0 / 9007199254740992.0); }

/* x_i = uniform[0,1) from a counter hash of (seed, i). */
ORACLE_API void oracle_gen_vector(int64_t n, uint64_t seed, double *v)
{
    for (int64_t i = 0; i < n; ++i) v[i] = u01(mix64(seed * 0xD1342543DE82EF95ull + (uint64_t)i));
}

/* 27-point stencil on an n^3 grid, row=(z*n+y)*n+x, neighbours in lexicographic (dz,dy,dx)
 * order (ascending columns), 26 on the diagonal and -1 elsewhere.  Rows [r0,r1), row_ptr
 * rebased to 0.  Returns the number of entries written. */
ORACLE_API int64_t oracle_gen_stencil27_csr(int n, int64_t r0, int64_t r1, int *row_ptr, int *col, double *val)
{
    int64_t p = 0;
    for (int64_t r = r0; r < r1; ++r) {
        int x = (int)(r % n), y = (int)((r / n) % n), z = (int)(r / ((int64_t)n * n));
        row_ptr[r - r0] = (int)p;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    int xx = x + dx, yy = y + dy, zz = z + dz;
                    if (xx < 0 || yy < 0 || zz < 0 || xx >= n || yy >= n || zz >= n) continue;
                    int64_t c = ((int64_t)zz * n + yy) * n + xx;
                    if (col) { col[p] = (int)c; val[p] = (c == r) ? 26.0 : -1.0; }
                    ++p;
                }
    }
    row_ptr[r1 - r0] = (int)p;
    return p;
}

/* 5-point Laplacian on an n x n grid as COO, entries row-major in (N,W,C,E,S) order. */
ORACLE_API int oracle_gen_lap5_coo(int n, int *ri, int *ci, double *v)
{
    int p = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            int r = i * n + j;
            if (i > 0)     { if (ri) { ri[p] = r; ci[p] = r - n; v[p] = -1.0; } ++p; }
            if (j > 0)     { if (ri) { ri[p] = r; ci[p] = r - 1; v[p] = -1.0; } ++p; }
                           { if (ri) { ri[p] = r; ci[p] = r;     v[p] = 4.0;  } ++p; }
            if (j < n - 1) { if (ri) { ri[p] = r; ci[p] = r + 1; v[p] = -1.0; } ++p; }
            if (i < n - 1) { if (ri) { ri[p] = r; ci[p] = r + n; v[p] = -1.0; } ++p; }
        }
    return p;
}

/* Uniform random COO: row, col ~ U[0,nrow) x U[0,ncol), value U[0,1), entry k from hash(seed,k). */
ORACLE_API void oracle_gen_uniform_coo(int nrow, int ncol, int64_t nnz, uint64_t seed, int *ri, int *ci, double *v)
{
    for (int64_t k = 0; k < nnz; ++k) {
        uint64_t h = mix64(seed * 0xD1342543DE82EF95ull + (uint64_t)k);
        uint64_t g = mix64(h);
        ri[k] = (int)((h >> 32) * (uint64_t)nrow >> 32);
        ci[k] = (int)((h & 0xFFFFFFFFull) * (uint64_t)ncol >> 32);
        v[k] = u01(g);
    }
}

/* R-MAT (a,b,c,d)=(0.57,0.19,0.19,0.05): `scale` quadrant draws per edge, duplicates kept. */
ORACLE_API void oracle_gen_rmat_coo(int scale, int64_t nnz, uint64_t seed, int *ri, int *ci, double *v)
{
    for (int64_t k = 0; k < nnz; ++k) {
        uint64_t s = mix64(seed * 0xD1342543DE82EF95ull + (uint64_t)k);
        int r = 0, c = 0;
        for (int lvl = 0; lvl < scale; ++lvl) {
            s = mix64(s);
            double u = u01(s);
            int rb = 0, cb = 0;
            if (u < 0.57) { rb = 0; cb = 0; }
            else if (u < 0.76) { rb = 0; cb = 1; }
            else if (u < 0.95) { rb = 1; cb = 0; }
            else { rb = 1; cb = 1; }
            r = (r << 1) | rb;
            c = (c << 1) | cb;
        }
        ri[k] = r; ci[k] = c;
        v[k] = u01(mix64(s));
    }
}

/* ------------------------------------------- canonical sum of squares + fingerprint -- */
/* No reference counterpart: vec_dot (src/vec_vec.cpp:15-29) leaves its reduction order to OpenMP.  These restate the
 * order the CUDA library fixes for the iterated loop (arm-spmv_b200/csrc/tree_sum.cuh) so that tests can demand the
 * same bits: tiles of 32 rows reduced by an xor-butterfly, then the binary tree over index bits, absent elements
 * counting as +0.0. */
ORACLE_API void oracle_tile_sumsq(int64_t n, const double *y, double *tile_ss)
{
    int64_t ntiles = (n + 31) / 32;
    for (int64_t t = 0; t < ntiles; ++t) {
        double v[32], w[32];
        for (int l = 0; l < 32; ++l) {
            int64_t i = t * 32 + l;
            double a = i < n ? y[i] : 0.0;
            v[l] = a * a;
        }
        for (int o = 16; o > 0; o >>= 1) {
            for (int l = 0; l < 32; ++l) w[l] = v[l] + v[l ^ o];
            for (int l = 0; l < 32; ++l) v[l] = w[l];
        }
        tile_ss[t] = v[0];
    }
}

/* index-bit tree: at level L element i (a multiple of 2^(L+1)) takes in element i + 2^L */
ORACLE_API double oracle_tree_sum(int64_t m, const double *vals)
{
    if (m <= 0) return 0.0;
    double *a = (double *)malloc(sizeof(double) * (size_t)m);
    memcpy(a, vals, sizeof(double) * (size_t)m);
    for (int64_t s = 1; s < m; s <<= 1)
        for (int64_t i = 0; i + s < m; i += 2 * s) a[i] = a[i] + a[i + s];
    double r = a[0];
    free(a);
    return r;
}

/* order-independent fingerprint of v[0..n) placed at global index `first` (twin of thsp_hash_f64) */
ORACLE_API uint64_t oracle_hash_f64(int64_t n, const double *v, uint64_t first)
{
    uint64_t h = 0;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t b;
        memcpy(&b, v + i, 8);
        h += mix64(b ^ mix64(first + (uint64_t)i));
    }
    return h;
}

/* ------------------------------------------------ SymGS and CG (no reference counterpart) -- */
/* The reference prepares for both (`diagonal` "for SymGS", include/matrix.h:36,81; vec_dot / vec_axpby / AddScaled without
 * a caller) and implements neither.  These restate what arm-spmv_b200/csrc/solvers.cu computes, serially, so that the
 * tests can demand the same bits. */

/* One row of a Gauss-Seidel sweep: t = sum_j a_ij x_j over ALL stored entries of the row (from 0.0, in stored order, the
 * diagonal included - the row sum of CSRMatrixMatVector, src/mat_vec.cpp:58-62), s = r_i - t, then s += x_i d_i and
 * x_i = s / d_i. */
static void symgs_row(int i, const int *rp, const int *ci, const double *va, const double *diag, const double *r, double *x)
{
    double acc = 0.0;
    for (int p = rp[i]; p < rp[i + 1]; ++p) {
        double t = va[p] * x[ci[p]];
        acc = acc + t;
    }
    double s = r[i] - acc;
    double u = x[i] * diag[i];
    s = s + u;
    x[i] = s / diag[i];
}

/* Multicolour symmetric sweep: colours ascending then descending, rows of a colour in the order of perm (they do not
 * touch each other, so any order gives the same bits).  color_ptr has ncolors + 1 entries. */
ORACLE_API void oracle_symgs(int ncolors, const int *color_ptr, const int *perm, const int *rp, const int *ci, const double *va,
                             const double *diag, const double *r, double *x)
{
    for (int c = 0; c < ncolors; ++c)
        for (int k = color_ptr[c]; k < color_ptr[c + 1]; ++k) symgs_row(perm[k], rp, ci, va, diag, r, x);
    for (int c = ncolors - 1; c >= 0; --c)
        for (int k = color_ptr[c]; k < color_ptr[c + 1]; ++k) symgs_row(perm[k], rp, ci, va, diag, r, x);
}

/* The textbook sequential sweep (rows 0..n-1, then n-1..0): what the multicolour sweep is a reordering of. */
ORACLE_API void oracle_symgs_sequential(int n, const int *rp, const int *ci, const double *va, const double *diag, const double *r,
                                        double *x)
{
    for (int i = 0; i < n; ++i) symgs_row(i, rp, ci, va, diag, r, x);
    for (int i = n - 1; i >= 0; --i) symgs_row(i, rp, ci, va, diag, r, x);
}

/* dot product in the canonical order: products per tile of 32 reduced by the xor-butterfly, tiles by the index-bit tree */
static double dot_canonical(int64_t n, const double *a, const double *b)
{
    int64_t ntiles = (n + 31) / 32;
    double *tile = (double *)malloc(sizeof(double) * (size_t)(ntiles > 0 ? ntiles : 1));
    for (int64_t t = 0; t < ntiles; ++t) {
        double v[32], w[32];
        for (int l = 0; l < 32; ++l) {
            int64_t i = t * 32 + l;
            v[l] = i < n ? a[i] * b[i] : 0.0;
        }
        for (int o = 16; o > 0; o >>= 1) {
            for (int l = 0; l < 32; ++l) w[l] = v[l] + v[l ^ o];
            for (int l = 0; l < 32; ++l) v[l] = w[l];
        }
        tile[t] = v[0];
    }
    double r = oracle_tree_sum(ntiles, tile);
    free(tile);
    return r;
}
ORACLE_API double oracle_dot_canonical(int64_t n, const double *a, const double *b) { return dot_canonical(n, a, b); }

/* Preconditioned CG exactly as thsp_cg_f64 runs it: r = b - A x through CSRMatrixMatVector + vec_axpby(1, b, -1, .),
 * x += alpha p and r += (-alpha) A p through AddScaled, p = beta p + z through vec_axpby(1, z, beta, p), dots canonical.
 * precond 0 none, 1 Jacobi, 2 one multicolour SymGS sweep from z = 0.  Returns the iteration count; *relres = ||r||/||b||. */
ORACLE_API int oracle_cg(int n, const int *rp, const int *ci, const double *va, const double *diag, int precond, int ncolors,
                         const int *color_ptr, const int *perm, const double *b, double *x, int maxit, double tol, double *relres)
{
    double *r = (double *)calloc((size_t)n, sizeof(double)), *p = (double *)calloc((size_t)n, sizeof(double));
    double *Ap = (double *)calloc((size_t)n, sizeof(double)), *zb = (double *)calloc((size_t)n, sizeof(double));
    double *z = precond ? zb : r;
    oracle_csr_spmv(n, rp, ci, va, x, r);                 /* r = 0 + A x */
    oracle_axpby(n, 1.0, b, -1.0, r, r);                  /* r = -1*r + b */
    double bb = dot_canonical(n, b, b), rr = dot_canonical(n, r, r);
#define PRECONDITION()                                                                                      \
    do {                                                                                                    \
        if (precond == 1) for (int i = 0; i < n; ++i) z[i] = r[i] / diag[i];                                \
        else if (precond == 2) { memset(z, 0, sizeof(double) * (size_t)n); oracle_symgs(ncolors, color_ptr, perm, rp, ci, va, diag, r, z); } \
    } while (0)
    PRECONDITION();
    memcpy(p, z, sizeof(double) * (size_t)n);
    double rz = precond ? dot_canonical(n, r, z) : rr;
    double bnorm = bb > 0.0 ? sqrt(bb) : 1.0;
    double rel = sqrt(rr) / bnorm;
    int it = 0;
    while (it < maxit && rel > tol) {
        memset(Ap, 0, sizeof(double) * (size_t)n);
        oracle_csr_spmv(n, rp, ci, va, p, Ap);
        double pAp = dot_canonical(n, p, Ap);
        double alpha = rz / pAp;
        oracle_add_scaled(n, alpha, p, x);
        oracle_add_scaled(n, -alpha, Ap, r);
        rr = dot_canonical(n, r, r);
        PRECONDITION();
        double rz_new = precond ? dot_canonical(n, r, z) : rr;
        double beta = rz_new / rz;
        rz = rz_new;
        oracle_axpby(n, 1.0, z, beta, p, p);
        rel = sqrt(rr) / bnorm;
        ++it;
    }
#undef PRECONDITION
    if (relres) *relres = rel;
    free(r); free(p); free(Ap); free(zb);
    return it;
}
