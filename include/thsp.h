/*
 * thsp.h -- C ABI of the B200-native SpMV library (libthsparse_cuda.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  The
 * reference (ChuheHong/arm-spmv) has no FFI of its own -- its API is C++ free functions and
 * classes (include/mat_vec.h, matrix.h, vec_vec.h, vector.h) -- so each entry point below
 * names the reference function or loop it replaces (paths relative to the reference root).
 * The C++ classes in this directory (matrix.h, vector.h, mat_vec.h, vec_vec.h, data_io.h)
 * are thin g++-compiled callers of these functions; tests/ and bench.py bind the same
 * functions through ctypes.
 *
 * Conventions
 *  - Every array argument is a DEVICE-accessible pointer (cudaMalloc, cudaMallocManaged or a
 *    torch CUDA tensor's data_ptr) unless the name ends in _host.
 *  - Indices are int32, values fp64 unless suffixed _f32 (the reference is double-only;
 *    fp32 is an extension asked for by the north star).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *    asynchronous on that stream unless documented otherwise; the C++ classes synchronise
 *    before returning because the reference's API is synchronous (main.cpp:56-59 brackets
 *    calls with mytimer()).
 *  - Return value: 0 on success, non-zero on failure; thsp_last_error() describes the last
 *    failure on the calling thread.  There is no CPU fallback anywhere: without a CUDA device
 *    every compute entry point fails.
 */
#ifndef THSP_H
#define THSP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define THSP_API __attribute__((visibility("default")))
#else
#define THSP_API
#endif

typedef void* thsp_stream_t;

/* ------------------------------------------------------------------ runtime -------- */
THSP_API const char* thsp_version(void);
THSP_API const char* thsp_last_error(void);
THSP_API int thsp_device_count(int* count);
THSP_API int thsp_set_device(int device);
THSP_API int thsp_get_device(int* device);
THSP_API int thsp_sm_count(int* count);
/* Memory.  `managed` allocations are host-dereferenceable (the reference's classes expose raw
 * pointers that main.cpp:48-51 reads on the host); see INTEGRATION.md "ownership". */
THSP_API int thsp_malloc(void** ptr, size_t bytes);
THSP_API int thsp_malloc_managed(void** ptr, size_t bytes);
THSP_API int thsp_malloc_host(void** ptr, size_t bytes); /* pinned */
/* Page-lock host memory the caller already owns (a shared-memory segment holding x for all the processes of a box):
 * copies from it run at pinned-memory speed and may be captured into the host-buffer SpMV's CUDA graph. */
THSP_API int thsp_host_register(void* ptr, size_t bytes);
THSP_API int thsp_host_unregister(void* ptr);
THSP_API int thsp_free(void* ptr);
THSP_API int thsp_free_host(void* ptr);
/* 0 = plain host, 1 = device, 2 = managed, 3 = pinned host, -1 = the CUDA runtime could not classify the pointer
 * (a sticky error, or the runtime is shutting down) - callers must not treat that as host memory */
THSP_API int thsp_pointer_kind(const void* ptr);
THSP_API int thsp_memcpy_h2d(void* dst, const void* src_host, size_t bytes, thsp_stream_t stream);
THSP_API int thsp_memcpy_d2h(void* dst_host, const void* src, size_t bytes, thsp_stream_t stream);
THSP_API int thsp_memcpy_d2d(void* dst, const void* src, size_t bytes, thsp_stream_t stream);
THSP_API int thsp_memset(void* dst, int byte, size_t bytes, thsp_stream_t stream);
THSP_API int thsp_prefetch(const void* managed_ptr, size_t bytes, int to_device, thsp_stream_t stream);
/* cudaMemAdviseSetReadMostly (on = 1) / Unset (on = 0) on a managed range that is COMPLETE: host reads then leave the copy
 * in HBM valid (the reference's main.cpp:46-52 reads the COO arrays on the host between the reader and the first product).
 * Not for arrays a kernel still has to write: writes to read-mostly pages are extremely slow. */
THSP_API int thsp_advise_read_mostly(const void* managed_ptr, size_t bytes, int on);
THSP_API int thsp_stream_sync(thsp_stream_t stream);
THSP_API int thsp_device_sync(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
THSP_API uint64_t thsp_launch_count(void);
/* Scratch (partials of reductions, merge-path carries, the products of the two-phase COO path, the buffers of the
 * conversions) is kept per DEVICE, not per stream: calls that need it must not run at the same time on different streams
 * or host threads of one device - like the reference, whose conversions and products are not re-entrant either
 * (SURVEY.md 8b).  A buffer that has to grow waits for the device first and refuses to grow inside a stream capture. */
/* Frees the library's scratch buffers on the current device (they grow to the largest call seen:
 * sort buffers of a conversion, staged text of the reader).  The next call re-allocates. */
THSP_API int thsp_scratch_release(void);

/* --------------------------------------------------------------------- SpMV -------- */
/* Kernel ids for CSR.  AUTO picks from the row-length statistics gathered by the plan. */
enum {
    THSP_CSR_AUTO = 0,
    THSP_CSR_SCALAR = 1,    /* one thread per row, loads straight from global              */
    THSP_CSR_VECTOR = 2,    /* L lanes per row (L = 2..32, `lanes` argument), shuffle tree   */
    THSP_CSR_STREAM = 3,    /* TMA bulk ring -> shared memory, one thread per row, in order  */
    THSP_CSR_MERGE = 4      /* nnz-balanced tiles with segmented reduction + carry fix-up    */
};

/* CSRMatrixMatVector (src/mat_vec.cpp:44-67): y[i] (+)= sum_j val[j]*x[col[j]].
 * accumulate=1 is the reference's y += A x; accumulate=0 writes y = A x (saves the y read and
 * the caller's Fill(0)).  Stateless form: chooses a kernel from nnz/nrow alone. */
THSP_API int thsp_csr_spmv_f64(int nrow, int ncol, int nnz, const int* row_ptr, const int* col_ind, const double* val,
                               const double* x, double* y, int accumulate, thsp_stream_t stream);
THSP_API int thsp_csr_spmv_f32(int nrow, int ncol, int nnz, const int* row_ptr, const int* col_ind, const float* val,
                               const float* x, float* y, int accumulate, thsp_stream_t stream);
/* Forced-kernel form (tests and tuning).  `lanes` is used by THSP_CSR_VECTOR only. */
THSP_API int thsp_csr_spmv_kernel_f64(int kernel, int lanes, int nrow, int ncol, int nnz, const int* row_ptr,
                                      const int* col_ind, const double* val, const double* x, double* y,
                                      int accumulate, thsp_stream_t stream);
THSP_API int thsp_csr_spmv_kernel_f32(int kernel, int lanes, int nrow, int ncol, int nnz, const int* row_ptr,
                                      const int* col_ind, const float* val, const float* x, float* y,
                                      int accumulate, thsp_stream_t stream);

/* Plan: row-length histogram -> kernel + lane count, plus the merge kernel's tile table.
 * Replaces nothing in the reference (its loop is static-scheduled, src/mat_vec.cpp:54-57);
 * it is where the north star's "lane count picked from the row-length histogram" lives. */
typedef struct thsp_csr_plan thsp_csr_plan;
THSP_API int thsp_csr_plan_create(thsp_csr_plan** plan, int nrow, int ncol, int nnz, const int* row_ptr,
                                  const int* col_ind, const void* val, int value_bytes /* 8 or 4 */,
                                  thsp_stream_t stream);
THSP_API int thsp_csr_plan_destroy(thsp_csr_plan* plan);
/* A plan remembers the arrays' addresses, the entry count and the kernel chosen from the row lengths - nothing else
 * derived from their contents.  If a caller rewrites row_ptr in place so that row_ptr[nrow] changes, the kernels whose
 * launch shape depends on the entry count (stream, merge) notice it on the device, write nothing and raise a flag:
 * after synchronising the stream, *stale = 1 says "destroy this plan, create a new one, repeat the product"; the flag is
 * cleared by the call.  The C++ classes do exactly that (csrc/host/mat_vec.cpp). */
THSP_API int thsp_csr_plan_stale(const thsp_csr_plan* plan, int* stale);
THSP_API int thsp_csr_plan_kernel(const thsp_csr_plan* plan, int* kernel, int* lanes);
THSP_API int thsp_csr_plan_set_kernel(thsp_csr_plan* plan, int kernel, int lanes);
/* Tuning knobs of the STREAM kernel (0 keeps the current value): warps per CTA, ring depth per
 * warp, entries per stage (multiple of 4), number of persistent CTAs. */
THSP_API int thsp_csr_plan_set_stream_config(thsp_csr_plan* plan, int warps, int stages, int chunk, int ctas);
/* Replace the heuristic choice by a measurement: times every applicable kernel on this matrix
 * (scratch vectors, ~20 SpMVs) and keeps the fastest.  Synchronous. */
THSP_API int thsp_csr_plan_autotune(thsp_csr_plan* plan, thsp_stream_t stream);
/* histogram[b] = number of rows whose length l satisfies: b=0: l==0; b>=1: 2^(b-1) <= l < 2^b  (32 bins) */
THSP_API int thsp_csr_plan_histogram(const thsp_csr_plan* plan, int64_t* histogram32, int* max_row_len);
THSP_API int thsp_csr_plan_spmv_f64(const thsp_csr_plan* plan, const double* x, double* y, int accumulate,
                                    thsp_stream_t stream);
THSP_API int thsp_csr_plan_spmv_f32(const thsp_csr_plan* plan, const float* x, float* y, int accumulate,
                                    thsp_stream_t stream);
/* The product, and in the same pass the sum of squares of each TILE of 32 consecutive rows of the result:
 * tile_ss[t] = y[32t]^2 + ... + y[32t+31]^2 added as an xor-butterfly (offsets 16, 8, 4, 2, 1), ceil(nrow/32) doubles.
 * With thsp_tree_sum_f64 this is vec_dot(y, y) (src/vec_vec.cpp:15-29) in an order that does not depend on how the
 * rows are spread over GPUs (csrc/tree_sum.cuh); the stream kernel writes the partials from its epilogue, so y is not
 * read again - the other kernels are followed by thsp_tile_sumsq_f64. */
THSP_API int thsp_csr_plan_spmv_sumsq_f64(const thsp_csr_plan* plan, const double* x, double* y, int accumulate,
                                          double* tile_ss, thsp_stream_t stream);
/* y = A (s x) with s = *xscale read from the device when the kernel starts: every gathered x_j is multiplied by s
 * (one rounding, as a pass x <- s x would have done - same bits of y) before it meets its matrix entry.  Lets an iterated
 * loop keep its vector unnormalised: x = vec_axpby(1/nrm, y, 0, y) (src/vec_vec.cpp:46-53) is never written.  Needs a
 * plan that runs the stream kernel (error otherwise); tile_ss as above, or NULL. */
THSP_API int thsp_csr_plan_spmv_scaled_f64(const thsp_csr_plan* plan, const double* x, const double* xscale, double* y,
                                           int accumulate, double* tile_ss, thsp_stream_t stream);
/* Same, with HOST x and y (pinned or pageable): H2D of x, kernel, D2H of y, then synchronises.
 * This is the call bench.py times for its end-to-end number.  Two forms: row chunks pipelined over three streams (any
 * kernel, pageable buffers, y += A x), and - stream kernel, y = A x, page-locked x and y, >= 2 M rows - the "flow" form:
 * one upload of x, one persistent launch that multiplies right behind the arriving x and stores y straight into y_host
 * (x_dev_scratch is pre-filled with a NaN pattern that marks values not yet arrived; an x that contains that very
 * pattern, 0x7FF85EEDC0DEF00D, makes the form give up after ~4 s and the chunked form run instead).  THSP_HOST_FLOW=0
 * selects the chunked form always.  y_dev_scratch is not touched by the flow form. */
THSP_API int thsp_csr_plan_spmv_host_f64(const thsp_csr_plan* plan, const double* x_host, double* y_host,
                                         double* x_dev_scratch, double* y_dev_scratch, int accumulate,
                                         thsp_stream_t stream);

/* ELLMatrixMatVector (src/mat_vec.cpp:97-121): column-major slab col[i + k*nrow]; accumulates
 * slot by slot into y exactly in the reference's order (y's old value is the first addend). */
THSP_API int thsp_ell_spmv_f64(int nrow, int ncol, int width, const int* col_ind, const double* val, const double* x,
                               double* y, thsp_stream_t stream);
THSP_API int thsp_ell_spmv_f32(int nrow, int ncol, int width, const int* col_ind, const float* val, const float* x,
                               float* y, thsp_stream_t stream);
/* COOMatirxMatVector [sic] (src/mat_vec.cpp:18-42): y[row[k]] += val[k]*x[col[k]]. */
THSP_API int thsp_coo_spmv_f64(int nrow, int ncol, int nnz, const int* row_ind, const int* col_ind, const double* val,
                               const double* x, double* y, thsp_stream_t stream);
/* The same with the path named (tests, benches): 0 = one kernel (products and scatter fused), 1 = slab by slab, all
 * products first (only x touched at random), then the scatter (only y) - what thsp_coo_spmv_f64 picks for large
 * matrices whose rows and columns both jump at random; -1 = let the library decide as thsp_coo_spmv_f64 does. */
THSP_API int thsp_coo_spmv_path_f64(int path, int nrow, int ncol, int nnz, const int* row_ind, const int* col_ind,
                                    const double* val, const double* x, double* y, thsp_stream_t stream);
/* CSCMatrixMatVector (src/mat_vec.cpp:69-95): column scatter y[row[j]] += val[j]*x[c]. */
THSP_API int thsp_csc_spmv_f64(int nrow, int ncol, int nnz, const int* col_ptr, const int* row_ind, const double* val,
                               const double* x, double* y, thsp_stream_t stream);
/* DIAMatrixMatVector (src/mat_vec.cpp:123-146): row-major values[i*ndiags+d], guard j<nrow. */
THSP_API int thsp_dia_spmv_f64(int nrow, int ncol, int ndiags, const int* offsets, const double* values,
                               const double* x, double* y, thsp_stream_t stream);
/* Row block [row_begin, row_begin+row_count) of the same product: `values` and `y` point at the
 * block's first row, x is the whole vector, nrow the whole matrix' row count (the column guard).
 * DIAMatrixMatVectorNumaThread (src/mat_vec.cpp:580-606) with a global, not block-local, guard. */
THSP_API int thsp_dia_spmv_rows_f64(int row_begin, int row_count, int nrow, int ndiags, const int* offsets,
                                    const double* values, const double* x, double* y, thsp_stream_t stream);

/* -------------------------------------------------------------- conversions -------- */
/* All are stable: within a row (column) entries keep their COO order and duplicates are kept,
 * so every output array equals the reference constructor's bit for bit. */
/* CSRMatrix::CSRMatrix(const COOMatrix&) (src/matrix.cpp:115-154).  diagonal may be NULL;
 * *ndiag (host, may be NULL) receives the number of row==col entries; at most nrow are stored. */
THSP_API int thsp_coo2csr(int nrow, int ncol, int nnz, const int* row_ind, const int* col_ind, const double* val,
                          int* row_ptr, int* out_col_ind, double* out_val, double* diagonal, int* ndiag,
                          thsp_stream_t stream);
/* CSCMatrix::CSCMatrix(const COOMatrix&) (src/matrix.cpp:295-325). */
THSP_API int thsp_coo2csc(int nrow, int ncol, int nnz, const int* row_ind, const int* col_ind, const double* val,
                          int* col_ptr, int* out_row_ind, double* out_val, thsp_stream_t stream);
/* Which way the last thsp_coo2csr / thsp_coo2csc of this process went (for tests and traces): 0 = keys already
 * ordered, entries copied through; 1 = stable radix sort; 2 = entries ordered by the OTHER index within a band that
 * fits L2 (a row-by-row stencil or banded matrix on its way to CSC): per-bucket cursors and a per-bucket sort by entry
 * number, no radix sort; 3 = the same tried and given up for the radix sort after the first pass (a bucket longer than
 * 64 entries, an index out of range, the order breaking or the band widening later in the arrays).  The output arrays
 * are the same bits whichever way it went. */
THSP_API int thsp_coo_last_path(void);
/* ELLMatrix::ELLMatrix(const COOMatrix&) (src/matrix.cpp:450-500) in two steps because the
 * caller must allocate nrow*width slots: width = longest row, then the fill. Synchronous. */
THSP_API int thsp_coo2ell_width(int nrow, int nnz, const int* row_ind, int* width, thsp_stream_t stream);
/* The same width, obtained by doing the expensive half of the conversion (the row sort) and keeping it in the
 * library's scratch: a thsp_coo2ell with the same arrays, sizes and width that FOLLOWS it on this device - no other
 * conversion, Matrix Market parse or merge-path SpMV in between, entries unchanged - only writes the slab (one pass
 * over the entries instead of a histogram plus the whole conversion).  Anything else in between is detected and
 * thsp_coo2ell converts from scratch as usual.  Synchronous; not re-entrant (like the reference). */
THSP_API int thsp_coo2ell_prepare(int nrow, int ncol, int nnz, const int* row_ind, const int* col_ind, const double* val,
                                  int* width, thsp_stream_t stream);
THSP_API int thsp_coo2ell(int nrow, int ncol, int nnz, const int* row_ind, const int* col_ind, const double* val,
                          int width, int* out_col_ind, double* out_val, double* diagonal, int* ndiag,
                          thsp_stream_t stream);
/* DIAMatrix::DIAMatrix(const CSRMatrix&) (src/matrix.cpp:673-726): count/emit ascending offsets,
 * then the row-major fill (last duplicate wins).  offsets == NULL just counts. Synchronous. */
THSP_API int thsp_csr2dia_offsets(int nrow, int ncol, const int* row_ptr, const int* col_ind, int* ndiags,
                                  int* offsets, int offsets_capacity, thsp_stream_t stream);
THSP_API int thsp_csr2dia_fill(int nrow, int ncol, const int* row_ptr, const int* col_ind, const double* val,
                               int ndiags, const int* offsets, double* values, thsp_stream_t stream);
/* Called by the reader once the sizes are known (COOMatrixRead, src/data_io.cpp:45-105): grows the library's scratch
 * to what conversions of an nrow x ncol matrix with nnz entries need and runs each conversion once on a 96-entry matrix,
 * so that the constructors that follow - main.cpp:38-41 calls each exactly once - run at their steady-state speed instead
 * of paying for allocations and first-launch kernel loading.  Optional: everything works without it.  Synchronous. */
THSP_API int thsp_prepare_conversions(int nrow, int ncol, int nnz, thsp_stream_t stream);
/* row_ptr scan on its own: exclusive prefix sum of n int32 counts into out[0..n] (out[n]=total). */
THSP_API int thsp_exclusive_scan_i32(int n, const int* counts, int* out, thsp_stream_t stream);

/* ------------------------------------------------------------------ vectors -------- */
/* vec_dot (src/vec_vec.cpp:15-29).  Deterministic two-level tree; *result_host is written
 * after a stream synchronise.  The _dev form leaves the scalar on the device (no sync). */
THSP_API int thsp_dot_f64(int64_t n, const double* x, const double* y, double* result_host, thsp_stream_t stream);
THSP_API int thsp_dot_dev_f64(int64_t n, const double* x, const double* y, double* result_dev, thsp_stream_t stream);
/* vec_dot(y, y) in the canonical order of csrc/tree_sum.cuh, in two halves: per-tile partials (32 rows, butterfly), then
 * the binary tree over the index bits of the m partials (absent ones count as +0.0) into *out_dev.  The tree half also
 * combines the per-rank results of a partitioned vector (m = number of ranks). */
THSP_API int thsp_tile_sumsq_f64(int64_t n, const double* y, double* tile_ss, thsp_stream_t stream);
/* out[i] = x[i] * *scale_dev; *inv_dev = 1 / sqrt(*sumsq_dev) - the two halves of x = vec_axpby(1/sqrt(s), y, 0, y)
 * with the scalar staying on the device (iterated loops that defer the normalisation into the next product) */
THSP_API int thsp_scale_by_dev_f64(int64_t n, const double* x, const double* scale_dev, double* out, thsp_stream_t stream);
THSP_API int thsp_inv_sqrt_dev_f64(const double* sumsq_dev, double* inv_dev, thsp_stream_t stream);
THSP_API int thsp_tree_sum_f64(int64_t m, const double* vals, double* out_dev, thsp_stream_t stream);
/* Order-independent 64-bit fingerprint of v[0..n) sitting at global index first_index of a longer vector: pieces add up
 * (mod 2^64) to the fingerprint of the whole.  bench.py compares row blocks on N GPUs with the one-GPU run. Synchronous. */
THSP_API int thsp_hash_f64(int64_t n, const double* v, uint64_t first_index, uint64_t* hash_host, thsp_stream_t stream);
/* vec_axpby (src/vec_vec.cpp:31-94): same seven branches, unfused multiply/add -> bit-exact. */
THSP_API int thsp_axpby_f64(int64_t n, double alpha, const double* x, double beta, const double* y, double* w,
                            thsp_stream_t stream);
/* Vector::Fill/Scale/Shift/Copy/AddScaled/Add2Scaled (src/vector.cpp:59-159), bit-exact. */
THSP_API int thsp_fill_f64(int64_t n, double a, double* v, thsp_stream_t stream);
THSP_API int thsp_scale_f64(int64_t n, double a, double* v, thsp_stream_t stream);
THSP_API int thsp_shift_f64(int64_t n, double a, double* v, thsp_stream_t stream);
THSP_API int thsp_copy_f64(int64_t n, const double* x, double* v, thsp_stream_t stream);
THSP_API int thsp_add_scaled_f64(int64_t n, double a, const double* x, double* v, thsp_stream_t stream);
THSP_API int thsp_add2_scaled_f64(int64_t n, double a, const double* x, double b, const double* y, double* v,
                                  thsp_stream_t stream);
/* checkVector (src/vector.cpp:161-171): *ok_host = 1 iff sizes match and max|x-y| <= 1e-6. Synchronous. */
THSP_API int thsp_check_vector_f64(int64_t nx, const double* x, int64_t ny, const double* y, int* ok_host,
                                   thsp_stream_t stream);

/* ------------------------------------------------ callers above the path */
/* The reference keeps a `diagonal` array "for SymGS" (include/matrix.h:36,81; src/matrix.cpp:146-153, 491-499) and the
 * vector kernels of a Krylov loop (src/vec_vec.cpp:15-94, src/vector.cpp:96-159) without a caller.  csrc/solvers.cu. */
/* diag[i] = sum of the stored entries (i, i) of row i, 0 if none.  Equals the reference's packed `diagonal` whenever that
 * is usable at all: a row-sorted COO with one diagonal entry per row (the packing is by COO order, src/matrix.cpp:146-153). */
THSP_API int thsp_csr_diagonal_f64(int nrow, const int* row_ptr, const int* col_ind, const double* val, double* diag,
                                   thsp_stream_t stream);
/* x[i] += omega * r[i] / diag[i] */
THSP_API int thsp_jacobi_update_f64(int64_t n, double omega, const double* diag, const double* r, double* x,
                                    thsp_stream_t stream);
/* Symmetric Gauss-Seidel.  The plan colours the rows (deterministic greedy colouring, <= 64 colours, any sparsity
 * pattern) and groups them by colour; thsp_symgs_f64 does one forward and one backward sweep
 *     t = sum_j a_ij x_j (from 0, stored order) ; s = r_i - t ; s += x_i d_i ; x_i = s / d_i     (d = `diagonal`, one value per row)
 * colour by colour, a thread per row in stored order, unfused arithmetic: the same bits as a serial walk over the same
 * colours (the checker's twin in oracle/oracle.c).  thsp_symgs_plan_info: colour count, colouring rounds, colour offsets (host, ncolors + 1 ints) and the
 * device arrays perm (rows grouped by colour) / color (colour of each row). */
typedef struct thsp_symgs_plan thsp_symgs_plan;
/* val != NULL: the plan also keeps a copy of the matrix permuted by colour (a snapshot: make a new plan after changing the
 * matrix), which lets each colour be swept by the TMA stream kernel of the CSR path instead of a thread per row - same
 * arithmetic, same bits, several times the bandwidth on matrices the stream kernel takes (mean row length >= 4, longest
 * row <= 2048).  *streams of thsp_symgs_plan_streams tells whether that copy exists. */
THSP_API int thsp_symgs_plan_create(thsp_symgs_plan** plan, int nrow, const int* row_ptr, const int* col_ind, const double* val,
                                    thsp_stream_t stream);
THSP_API int thsp_symgs_plan_streams(const thsp_symgs_plan* plan, int* streams);
THSP_API int thsp_symgs_plan_destroy(thsp_symgs_plan* plan);
THSP_API int thsp_symgs_plan_info(const thsp_symgs_plan* plan, int* ncolors, int* rounds, int* color_ptr_host, int capacity,
                                  const int** perm_dev, const int** color_dev);
THSP_API int thsp_symgs_f64(const thsp_symgs_plan* plan, int nrow, const int* row_ptr, const int* col_ind, const double* val,
                            const double* diagonal, const double* r, double* x, thsp_stream_t stream);
/* vec_dot (src/vec_vec.cpp:15-29) in the canonical order (32-element butterflies, index-bit tree: csrc/tree_sum.cuh),
 * result on the device; tile_scratch = ceil(n/32) doubles. */
THSP_API int thsp_dot_canonical_dev_f64(int64_t n, const double* x, const double* y, double* tile_scratch, double* out_dev,
                                        thsp_stream_t stream);
/* Preconditioned conjugate gradients for a symmetric positive definite CSR matrix: A = its plan (the SpMV), precond 0 none /
 * 1 Jacobi (z = r / diagonal) / 2 one SymGS sweep from z = 0 (needs M and the matrix arrays).  Vector updates in the
 * reference's forms (AddScaled, vec_axpby alpha == 1), dots in the canonical order, scalars on the device; the host reads
 * ||r||^2 once per iteration.  work = thsp_cg_work_doubles(n) doubles of device memory.  Stops at ||r|| / ||b|| <= tol or
 * maxit; *iters / *relres (host) report where.  Synchronous. */
THSP_API int64_t thsp_cg_work_doubles(int64_t n);
THSP_API int thsp_cg_f64(const thsp_csr_plan* A, int n, int precond, const thsp_symgs_plan* M, const int* row_ptr, const int* col_ind,
                         const double* val, const double* diagonal, const double* b, double* x, int maxit, double tol, double* work,
                         int* iters, double* relres, thsp_stream_t stream);

/* --------------------------------------------- row-block partition (multi-GPU) ------ */
/* *MatVectorNuma (src/mat_vec.cpp:230-268): equal row blocks, last takes the remainder. Host-only. */
THSP_API int thsp_partition_rows(int64_t nrow, int nparts, int part, int64_t* start, int64_t* count);
/* sub_row_ptr[j] = row_ptr[start+j] - row_ptr[start], j = 0..count (src/mat_vec.cpp:260-263). */
THSP_API int thsp_csr_slice_row_ptr(const int* row_ptr, int start, int count, int* sub_row_ptr, thsp_stream_t stream);
/* Power-iteration tail fused with the x refresh: dst_k[offset+i] = src[i] * (1/sqrt(*sumsq_dev))
 * for every peer replica k (peer pointers are device pointers mapped over NVLink, or just the
 * local replica when npeers == 1).  Replaces nothing in the reference (its NUMA loop never
 * refreshes x); composes vec_axpby's beta==0 branch (src/vec_vec.cpp:46-53). */
THSP_API int thsp_scale_broadcast_f64(int64_t n, const double* src, const double* sumsq_dev, double* const* peer_dst,
                                      int npeers, int64_t offset, thsp_stream_t stream);
/* sum of squares of y into *out_dev (device scalar), deterministic; = vec_dot(y,y). */
THSP_API int thsp_sumsq_dev_f64(int64_t n, const double* y, double* out_dev, thsp_stream_t stream);

/* ------------------------------------------------ Matrix Market entries on the GPU --- */
/* The entry loop of COOMatrixRead (src/data_io.cpp:83-88: fscanf("%d %d %lg\n") per entry, indices
 * made 0-based).  text_host[0, len) = the bytes of the file after the size line (host memory);
 * row_ind / col_ind / val = device or managed arrays of nnz entries.  The values are the correctly
 * rounded doubles strtod would return.  *status = 0: parsed here.  *status = 1: the text holds
 * something the fast conversions do not cover (inf/nan, hex floats, more than 19 significant
 * digits, malformed or missing tokens, 4 GB or more) - outputs undefined, run the scanf loop. */
THSP_API int thsp_mtx_parse_coo(const char* text_host, size_t len, int nnz, int* row_ind, int* col_ind, double* val,
                                int* status, thsp_stream_t stream);

/* ------------------------- power-iteration step fused with its exchange (NVLink) ------ */
/* The vector half of y = A x; s = vec_dot(y,y); x = vec_axpby(1/sqrt(s), y, 0, y) (src/vec_vec.cpp:15-53)
 * for row blocks on several GPUs (src/mat_vec.cpp:230-297), with no collective call: producers store
 * into the peers' memory and raise a flag, consumers wait on the flag (csrc/exchange.cu).
 * ctrl = thsp_xchg_ctrl_bytes() of zeroed memory per rank that every rank can address (peer
 * pointers); work = thsp_xchg_work_bytes() of zeroed private device memory. iter counts from 1. */
THSP_API int thsp_xchg_ctrl_bytes(void);
THSP_API int thsp_xchg_work_bytes(void);
/* partial sum of y_i^2 of this rank (fixed order) -> slot [iter&1][rank] of every rank's ctrl */
THSP_API int thsp_xchg_sumsq_publish_f64(int64_t n, const double* y, uint64_t iter, int world, int rank,
                                         void* const* peer_ctrl, void* work, thsp_stream_t stream);
/* waits for all partials of `iter`, adds them in rank order, x[offset+i] = y[i]/sqrt(sum) into the
 * local replica and into dest_x[d] where dest_lo[d] <= offset+i < dest_hi[d]; then raises the
 * "halo from `rank`" flag in dest_ctrl[d].  *sumsq_out (device scalar, required) = the sum. */
THSP_API int thsp_xchg_scale_push_f64(int64_t n, const double* y, uint64_t iter, int world, int rank, void* ctrl_local,
                                      void* work, double* x_local, int64_t offset, int ndest, double* const* dest_x,
                                      void* const* dest_ctrl, const int64_t* dest_lo, const int64_t* dest_hi,
                                      double* sumsq_out, thsp_stream_t stream);
/* Both of the above in one kernel, fed by the per-tile sums of squares of thsp_csr_plan_spmv_sumsq_f64 (tile_ss, ceil(n/32)
 * doubles for this rank's n rows): tree over the tiles, partial published to every rank, partials of all ranks combined
 * by the same tree over rank numbers (csrc/tree_sum.cuh: same bits on 1, 2, 4, 8 GPUs for aligned row blocks), then the
 * normalise + push + flags of thsp_xchg_scale_push_f64.  peer_ctrl[r] = control block of rank r (own one included).
 * A peer that never publishes poisons *sumsq_out with NaN; x is then left untouched. */
THSP_API int thsp_xchg_norm_scale_push_f64(int64_t n, const double* y, const double* tile_ss, uint64_t iter, int world, int rank,
                                           void* const* peer_ctrl, void* work, double* x_local, int64_t offset, int ndest,
                                           double* const* dest_x, void* const* dest_ctrl, const int64_t* dest_lo,
                                           const int64_t* dest_hi, double* sumsq_out, thsp_stream_t stream);
/* The step with the normalisation deferred into the next product (thsp_csr_plan_spmv_scaled_f64): y points at this rank's
 * n rows (global rows offset .. offset+n), which ARE its slice of the next input vector.  Tree over tile_ss, partial
 * published; the pieces [dest_lo[d], dest_hi[d]) of the raw y copied into dest_x[d] (the peers' replicas of the vector
 * the NEXT product reads) and their flags raised; partials of all ranks combined: *sumsq_out = sum, *inv_out =
 * 1/sqrt(sum) (device scalars).  The n-element vector is neither read nor written except for the pushed pieces. */
THSP_API int thsp_xchg_norm_push_f64(int64_t n, const double* y, const double* tile_ss, uint64_t iter, int world, int rank,
                                     void* const* peer_ctrl, void* work, int64_t offset, int ndest, double* const* dest_x,
                                     void* const* dest_ctrl, const int64_t* dest_lo, const int64_t* dest_hi,
                                     double* sumsq_out, double* inv_out, thsp_stream_t stream);
/* stream-ordered wait until the ranks in src_mask have raised their halo flag for `iter` */
THSP_API int thsp_xchg_wait(void* ctrl_local, uint64_t iter, unsigned src_mask, thsp_stream_t stream);
/* *flag_host = 1 if a wait of this rank gave up (~15 s) instead of hanging the GPU */
THSP_API int thsp_xchg_timed_out(const void* ctrl_local, int* flag_host, thsp_stream_t stream);

/* --------------------------------------------------------- synthetic inputs --------- */
/* SURVEY.md 8(d).  Device-side generators (the big configs cannot go through a .mtx file);
 * oracle/oracle.c carries CPU twins that produce identical arrays. */
/* 27-point stencil on n^3, rows [row_begin,row_end), row_ptr rebased to 0.  Call with
 * col_ind == NULL to fill row_ptr only (sub_nnz = row_ptr[row_end-row_begin] fits int32). */
THSP_API int thsp_gen_stencil27_csr(int n, int64_t row_begin, int64_t row_end, int* row_ptr, int* col_ind,
                                    double* val, thsp_stream_t stream);
THSP_API int64_t thsp_stencil27_nnz(int n, int64_t row_begin, int64_t row_end);
/* Same matrix straight into the reference's column-major ELL slab (width 27). */
THSP_API int thsp_gen_stencil27_ell(int n, int* col_ind, double* val, thsp_stream_t stream);
/* Same matrix as COO in row-major order (for the conversion benchmarks). */
THSP_API int thsp_gen_stencil27_coo(int n, int* row_ind, int* col_ind, double* val, thsp_stream_t stream);
THSP_API int thsp_gen_lap5_coo(int n, int* row_ind, int* col_ind, double* val, thsp_stream_t stream);
THSP_API int64_t thsp_lap5_nnz(int n);
THSP_API int thsp_gen_uniform_coo(int nrow, int ncol, int64_t nnz, uint64_t seed, int* row_ind, int* col_ind,
                                  double* val, thsp_stream_t stream);
THSP_API int thsp_gen_rmat_coo(int scale, int64_t nnz, uint64_t seed, int* row_ind, int* col_ind, double* val,
                               thsp_stream_t stream);
THSP_API int thsp_gen_vector_f64(int64_t n, uint64_t seed, double* v, thsp_stream_t stream);
THSP_API int thsp_f64_to_f32(int64_t n, const double* src, float* dst, thsp_stream_t stream);
/* Write `bytes` of junk through a buffer larger than L2 so the next timed launch starts cold. */
THSP_API int thsp_flush_l2(void* scratch, size_t bytes, thsp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* THSP_H */
