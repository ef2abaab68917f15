// mat_vec.cpp -- SpMV entry points of the arm-spmv API (replaces src/mat_vec.cpp).
//
//  *MatVector      one GPU, y += A x through the kernels of thsp.h, synchronous.
//  *MatVectorNuma  the reference's NUMA placement (src/mat_vec.cpp:148-484) mapped onto GPUs:
//                  row blocks (column blocks for CSC, entry blocks for COO) on
//                  G = min(nthreads, #GPUs) devices, x replicated, 50 timed repeats with all
//                  devices running concurrently, the reference's "### <FMT> NUMA GFLOPS" line,
//                  and - unlike the reference - the accumulated result written back to y.
#include "mat_vec.h"

#include <cuda_runtime_api.h>
#include <math.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "hostmem.h"
#include "mytime.h"
#include "numa_node.h"

using namespace thsp_host;

// ------------------------------------------------------------------------------------------
// Matrix arrays: library-owned (managed) memory is used in place; arrays adopted from the caller
// (src/matrix.cpp:12-15,88-91) get a device mirror that is uploaded once and found again on later calls
// (hostmem.h mirror()).  x and y change between calls: host vectors are staged through pooled device buffers.
namespace {
struct FirstCalls {   // THSP_TRACE=1: the first three calls of a product are timed (one-time costs against the steady state)
    int n = 0;
    bool more() { return n++ < 3; }
};
}  // namespace
#define TRACE_FIRST(name)            \
    static FirstCalls calls_;        \
    Trace tr_(calls_.more() ? name : nullptr)

void COOMatirxMatVector(const COOMatrix& A, const Vector& x, Vector& y)
{
    TRACE_FIRST("COOMatirxMatVector");
    if (A.nnz <= 0) return;
    const int* ri = mirror(A.row_ind, A.nnz);
    const int* ci = mirror(A.col_ind, A.nnz);
    const double* va = mirror(A.values, A.nnz);
    if (kind(A.values) == 2 && first_gpu_use(A.values)) {
        prefetch_traced(A.row_ind, sizeof(int) * (size_t)A.nnz);
        prefetch_traced(A.col_ind, sizeof(int) * (size_t)A.nnz);
        prefetch_traced(A.values, sizeof(double) * (size_t)A.nnz);
    }
    View<double> xv(x.values, A.ncol, false), yv(y.values, A.nrow, true);
    ok(thsp_coo_spmv_f64(A.nrow, A.ncol, A.nnz, ri, ci, va, xv, yv, nullptr), "COO SpMV");
    sync();
    yv.commit();
}

void CSRMatrixMatVector(const CSRMatrix& A, const Vector& x, Vector& y)
{
    TRACE_FIRST("CSRMatrixMatVector");
    if (A.nrow <= 0) return;
    View<double> xv(x.values, A.ncol, false), yv(y.values, A.nrow, true);
    for (int attempt = 0; attempt < 2; ++attempt) {
        // The plan remembers the kernel chosen from the row-length histogram and the entry count; creating it also
        // brings a managed matrix into HBM / uploads the mirror of an adopted one.  Plans are dropped when the matrix
        // is freed or reassigned (CSRMatrix::Free) and when a kernel finds that the entry count has changed.
        int nnz = 0;
        const int* rp = mirror(A.row_ptr, (size_t)A.nrow + 1);
        thsp_csr_plan* plan = nullptr;
        const int* ci = nullptr;
        const double* va = nullptr;
        const bool own = rp == A.row_ptr;   // device-usable in place
        if (own) plan = csr_plan(A.nrow, A.ncol, -1, A.row_ptr, A.col_ind, A.values, &nnz);
        if (!plan) {
            nnz = own ? peek_int(A.row_ptr + A.nrow) : A.row_ptr[A.nrow];
            ci = mirror(A.col_ind, (size_t)nnz);
            va = mirror(A.values, (size_t)nnz);
            if (own && kind(A.values) == 2) {
                if (first_gpu_use(A.row_ptr)) prefetch_traced(A.row_ptr, sizeof(int) * ((size_t)A.nrow + 1));
                if (nnz && first_gpu_use(A.col_ind)) prefetch_traced(A.col_ind, sizeof(int) * (size_t)nnz);
                if (nnz && first_gpu_use(A.values)) prefetch_traced(A.values, sizeof(double) * (size_t)nnz);
            }
            plan = csr_plan(A.nrow, A.ncol, nnz, rp, ci, va);   // keyed by the DEVICE addresses: found again through the mirrors
        }
        ok(thsp_csr_plan_spmv_f64(plan, xv, yv, 1, nullptr), "CSR SpMV");
        sync();
        int stale = 0;
        ok(thsp_csr_plan_stale(plan, &stale), "CSR plan check");
        if (!stale) break;
        if (attempt == 1) die("CSR SpMV (row_ptr changes while the product runs)");
        forget_plans(rp);   // row_ptr[nrow] is not what the plan was made for: nothing was written, plan again
        if (!own) invalidate(A.row_ptr), invalidate(A.col_ind), invalidate(A.values);
    }
    yv.commit();
}

void CSCMatrixMatVector(const CSCMatrix& A, const Vector& x, Vector& y)
{
    TRACE_FIRST("CSCMatrixMatVector");
    if (A.ncol <= 0) return;
    const int* cp = mirror(A.col_ptr, (size_t)A.ncol + 1);
    const int nnz = cp == A.col_ptr ? peek_int(A.col_ptr + A.ncol) : A.col_ptr[A.ncol];
    const int* ri = mirror(A.row_ind, (size_t)nnz);
    const double* va = mirror(A.values, (size_t)nnz);
    if (kind(A.values) == 2 && first_gpu_use(A.values)) {
        prefetch_traced(A.col_ptr, sizeof(int) * ((size_t)A.ncol + 1));
        prefetch_traced(A.row_ind, sizeof(int) * (size_t)nnz);
        prefetch_traced(A.values, sizeof(double) * (size_t)nnz);
    }
    View<double> xv(x.values, A.ncol, false), yv(y.values, A.nrow, true);
    ok(thsp_csc_spmv_f64(A.nrow, A.ncol, nnz, cp, ri, va, xv, yv, nullptr), "CSC SpMV");
    sync();
    yv.commit();
}

void ELLMatrixMatVector(const ELLMatrix& A, const Vector& x, Vector& y)
{
    TRACE_FIRST("ELLMatrixMatVector");
    const size_t total = (size_t)A.nrow * (size_t)A.nonzeros_in_row;
    if (total == 0) return;
    const int* ci = mirror(A.col_ind, total);
    const double* va = mirror(A.values, total);
    if (kind(A.values) == 2 && first_gpu_use(A.values)) {
        prefetch_traced(A.col_ind, sizeof(int) * total);
        prefetch_traced(A.values, sizeof(double) * total);
    }
    View<double> xv(x.values, A.ncol, false), yv(y.values, A.nrow, true);
    ok(thsp_ell_spmv_f64(A.nrow, A.ncol, A.nonzeros_in_row, ci, va, xv, yv, nullptr), "ELL SpMV");
    sync();
    yv.commit();
}

void DIAMatrixMatVector(const DIAMatrix& A, const Vector& x, Vector& y)
{
    TRACE_FIRST("DIAMatrixMatVector");
    const size_t total = (size_t)A.nrow * (size_t)A.ndiags;
    if (total == 0) return;
    const int* off = mirror(A.offsets, (size_t)A.ndiags);
    const double* va = mirror(A.values, total);
    if (kind(A.values) == 2 && first_gpu_use(A.values)) prefetch_traced(A.values, sizeof(double) * total);
    // The reference guards columns with j < nrow (src/mat_vec.cpp:140), so it reads x[0..nrow).
    View<double> xv(x.values, x.size, false), yv(y.values, A.nrow, true);
    ok(thsp_dia_spmv_f64(A.nrow, A.ncol, A.ndiags, off, va, xv, yv, nullptr), "DIA SpMV");
    sync();
    yv.commit();
}

// ============================================================ partitioned ("Numa") path ====
namespace {

const int kRepeats = 50;  // NTESTS in the reference (src/mat_vec.cpp:201,270,339,400,455)
thread_local bool g_driver_syncs = false;   // set while a *Numa driver on THIS thread launches on all devices and syncs once

int gpu_count_for(int nthreads)
{
    int n = 0;
    ok(thsp_device_count(&n), "device count");
    if (n <= 0) die("no CUDA device");
    if (nthreads < 1) nthreads = 1;
    return nthreads < n ? nthreads : n;
}

template <class T>
T* dev_alloc(size_t n)
{
    void* p = nullptr;
    ok(thsp_malloc(&p, (n ? n : 1) * sizeof(T)), "device allocation");
    return static_cast<T*>(p);
}

void use(int dev) { ok(thsp_set_device(dev), "set device"); }

void sync_all(int G)
{
    for (int d = 0; d < G; ++d) {
        use(d);
        ok(thsp_device_sync(), "device synchronise");
    }
}

// The reference's timing expression, kept verbatim in meaning (main.cpp:60, mat_vec.cpp:214):
// milliseconds per repeat plus a 1e-6 relative term; GFLOP/s = 2 nnz / t_ms / 1e6.
void report(const char* fmt, double nnz, double t_begin, double t_end)
{
    const double dt = t_end - t_begin;
    const double t_avg = (dt * 1000.0 + dt / 1000.0) / kRepeats;
    printf("### %s NUMA GFLOPS = %.5f\n", fmt, 2.0 * nnz / t_avg / pow(10, 6));
}

template <class Node, class Body>
void timed_repeats(std::vector<Node>& p, int G, Body body, const char* fmt, double nnz)
{
    g_driver_syncs = true;
    sync_all(G);
    const double t0 = mytimer();
    for (int k = 0; k < kRepeats; ++k)
        for (int i = 0; i < G; ++i) body((void*)&p[i]);  // asynchronous launches, one per GPU
    sync_all(G);
    const double t1 = mytimer();
    g_driver_syncs = false;
    report(fmt, nnz, t0, t1);
}

// y(dev 0, length n) += Y(dev d): private full-length results of the column/entry partitions.
void reduce_into(double* acc_dev0, const double* Yd, size_t n, double* tmp_dev0)
{
    copy_bytes(tmp_dev0, Yd, n * sizeof(double));
    ok(thsp_add_scaled_f64((int64_t)n, 1.0, tmp_dev0, acc_dev0, nullptr), "partial-result reduction");
    ok(thsp_device_sync(), "device synchronise");
}

}  // namespace

void* CSRMatrixMatVectorNumaThread(void* args)
{
    NumaNode4CSR* pn = static_cast<NumaNode4CSR*>(args);
    use(pn->alloc);
    // the block's plan (kernel chosen from ITS row lengths, as CSRMatrixMatVector does for the whole matrix) is made by the
    // driver below before the clock starts; a caller-built node gets its plan here on the first call
    thsp_csr_plan* plan = csr_plan(pn->rows_per_node, 0, pn->nnz, pn->sub_row_ptr, pn->sub_col_ind, pn->sub_values);
    ok(thsp_csr_plan_spmv_f64(plan, pn->X, pn->Y, 1, nullptr), "CSR block SpMV");
    if (!g_driver_syncs) ok(thsp_device_sync(), "device synchronise");
    return nullptr;
}

void* ELLMatrixMatVectorNumaThread(void* args)
{
    NumaNode4ELL* pn = static_cast<NumaNode4ELL*>(args);
    use(pn->alloc);
    ok(thsp_ell_spmv_f64(pn->rows_per_node, 0, pn->nonzeros_in_row, pn->sub_col_ind, pn->sub_values, pn->X, pn->Y, nullptr),
       "ELL block SpMV");
    if (!g_driver_syncs) ok(thsp_device_sync(), "device synchronise");
    return nullptr;
}

void* COOMatrixMatVectorNumaThread(void* args)
{
    NumaNode4COO* pn = static_cast<NumaNode4COO*>(args);
    use(pn->alloc);
    // row indices stay global; Y - start_row makes y[row - start_row] of the reference (:500)
    // (nrow = one past the last global row of the block: the kernel's row bound is on global indices)
    ok(thsp_coo_spmv_f64(pn->start_row + pn->rows_per_node, 0, pn->nnz, pn->sub_row_ind, pn->sub_col_ind, pn->sub_values, pn->X,
                         pn->Y - pn->start_row, nullptr),
       "COO block SpMV");
    if (!g_driver_syncs) ok(thsp_device_sync(), "device synchronise");
    return nullptr;
}

namespace {
// What the reference's node structs have no field for (include/numa_node.h is kept field-compatible): the row count of
// the whole matrix, per node, filled by the *Numa drivers below.  Keyed by the node's address and locked, so that the
// exported thread entries stay usable on their own and from several host threads.
std::mutex g_extra_mu;
std::unordered_map<const void*, int> g_rows_total;
void set_rows_total(const void* node, int n)
{
    std::lock_guard<std::mutex> lk(g_extra_mu);
    g_rows_total[node] = n;
}
void drop_rows_total(const void* node)
{
    std::lock_guard<std::mutex> lk(g_extra_mu);
    g_rows_total.erase(node);
}
int rows_total(const void* node, int fallback)
{
    std::lock_guard<std::mutex> lk(g_extra_mu);
    auto it = g_rows_total.find(node);
    return it == g_rows_total.end() ? fallback : it->second;
}
}  // namespace

void* CSCMatrixMatVectorNumaThread(void* args)
{
    NumaNode4CSC* pn = static_cast<NumaNode4CSC*>(args);
    use(pn->alloc);
    // A node built by the caller (the reference's thread body needs no row count, src/mat_vec.cpp:553-560): the kernel
    // only uses the count to clamp its shared-memory window, so "unknown" = no clamp.
    const int nrow = rows_total(pn, 0x7fffffff);
    ok(thsp_csc_spmv_f64(nrow, pn->cols_per_node, pn->nnz, pn->sub_col_ptr, pn->sub_row_ind, pn->sub_values, pn->X, pn->Y, nullptr),
       "CSC block SpMV");
    if (!g_driver_syncs) ok(thsp_device_sync(), "device synchronise");
    return nullptr;
}

void* DIAMatrixMatVectorNumaThread(void* args)
{
    NumaNode4DIA* pn = static_cast<NumaNode4DIA*>(args);
    use(pn->alloc);
    // caller-built node: the reference's own (block-local) column guard, src/mat_vec.cpp:597-598
    const int total = rows_total(pn, pn->start_row + pn->rows_per_node);
    ok(thsp_dia_spmv_rows_f64(pn->start_row, pn->rows_per_node, total, pn->ndiags, pn->offsets, pn->values, pn->X, pn->Y, nullptr),
       "DIA block SpMV");
    if (!g_driver_syncs) ok(thsp_device_sync(), "device synchronise");
    return nullptr;
}

// ---- CSR: row blocks, row_ptr rebased (src/mat_vec.cpp:230-297) --------------------------------
void CSRMatrixMatVectorNuma(const CSRMatrix& A, const Vector& x, Vector& y, int nthreads)
{
    Trace tr("CSRMatrixMatVectorNuma");
    const int G = gpu_count_for(nthreads);
    int dev0 = 0;
    thsp_get_device(&dev0);
    std::vector<NumaNode4CSR> p(G);
    std::vector<int> rp_host((size_t)A.nrow + 1);
    copy_bytes(rp_host.data(), A.row_ptr, sizeof(int) * ((size_t)A.nrow + 1));
    for (int i = 0; i < G; ++i) {
        int64_t start = 0, count = 0;
        ok(thsp_partition_rows(A.nrow, G, i, &start, &count), "row partition");
        NumaNode4CSR& b = p[i];
        b.alloc = i;
        b.core_ind = i;
        b.start_row = (int)start;
        b.rows_per_node = (int)count;
        const int e0 = rp_host[start], e1 = rp_host[start + count];
        b.nnz = e1 - e0;
        use(i);
        b.sub_row_ptr = dev_alloc<int>((size_t)count + 1);
        b.sub_col_ind = dev_alloc<int>(b.nnz);
        b.sub_values = dev_alloc<double>(b.nnz);
        b.X = dev_alloc<double>(x.size);
        b.Y = dev_alloc<double>(count);
        std::vector<int> sub((size_t)count + 1);
        for (int64_t j = 0; j <= count; ++j) sub[j] = rp_host[start + j] - e0;  // rebase (:260-263)
        copy_bytes(b.sub_row_ptr, sub.data(), sizeof(int) * sub.size());
        if (b.nnz) {
            copy_bytes(b.sub_col_ind, A.col_ind + e0, sizeof(int) * (size_t)b.nnz);
            copy_bytes(b.sub_values, A.values + e0, sizeof(double) * (size_t)b.nnz);
        }
        copy_bytes(b.X, x.values, sizeof(double) * (size_t)x.size);
        ok(thsp_memset(b.Y, 0, sizeof(double) * (size_t)count, nullptr), "memset");
        (void)csr_plan(b.rows_per_node, 0, b.nnz, b.sub_row_ptr, b.sub_col_ind, b.sub_values);   // row histogram -> kernel, off the clock
    }
    timed_repeats(p, G, CSRMatrixMatVectorNumaThread, "CSR", (double)rp_host[A.nrow]);
    for (int i = 0; i < G; ++i) {
        use(i);
        forget_plans(p[i].sub_row_ptr);   // the arrays go away: so does the plan keyed by their addresses
        copy_bytes(y.values + p[i].start_row, p[i].Y, sizeof(double) * (size_t)p[i].rows_per_node);
        thsp_free(p[i].sub_row_ptr); thsp_free(p[i].sub_col_ind); thsp_free(p[i].sub_values); thsp_free(p[i].X); thsp_free(p[i].Y);
    }
    use(dev0);
}

// ---- ELL: row blocks of the column-major slab (src/mat_vec.cpp:368-426) ------------------------
void ELLMatrixMatVectorNuma(const ELLMatrix& A, const Vector& x, Vector& y, int nthreads)
{
    Trace tr("ELLMatrixMatVectorNuma");
    const int G = gpu_count_for(nthreads);
    int dev0 = 0;
    thsp_get_device(&dev0);
    const int K = A.nonzeros_in_row;
    std::vector<NumaNode4ELL> p(G);
    std::vector<int> starts(G);
    for (int i = 0; i < G; ++i) {
        int64_t start = 0, count = 0;
        ok(thsp_partition_rows(A.nrow, G, i, &start, &count), "row partition");
        NumaNode4ELL& b = p[i];
        b.alloc = i;
        b.core_ind = i;
        b.rows_per_node = (int)count;
        b.nonzeros_in_row = K;
        starts[i] = (int)start;
        use(i);
        b.sub_col_ind = dev_alloc<int>((size_t)count * K);
        b.sub_values = dev_alloc<double>((size_t)count * K);
        b.X = dev_alloc<double>(x.size);
        b.Y = dev_alloc<double>(count);
        if (count && K) {
            // rows [start, start+count) of every slot: a strided (2-D) copy out of the slab
            if (cudaMemcpy2D(b.sub_col_ind, (size_t)count * sizeof(int), A.col_ind + start, (size_t)A.nrow * sizeof(int),
                             (size_t)count * sizeof(int), K, cudaMemcpyDefault) != cudaSuccess ||
                cudaMemcpy2D(b.sub_values, (size_t)count * sizeof(double), A.values + start, (size_t)A.nrow * sizeof(double),
                             (size_t)count * sizeof(double), K, cudaMemcpyDefault) != cudaSuccess)
                die("ELL block copy");
        }
        copy_bytes(b.X, x.values, sizeof(double) * (size_t)x.size);
        ok(thsp_memset(b.Y, 0, sizeof(double) * (size_t)count, nullptr), "memset");
    }
    timed_repeats(p, G, ELLMatrixMatVectorNumaThread, "ELL", (double)A.nrow * (double)K);  // the reference counts padded slots (:415)
    for (int i = 0; i < G; ++i) {
        use(i);
        copy_bytes(y.values + starts[i], p[i].Y, sizeof(double) * (size_t)p[i].rows_per_node);
        thsp_free(p[i].sub_col_ind); thsp_free(p[i].sub_values); thsp_free(p[i].X); thsp_free(p[i].Y);
    }
    use(dev0);
}

// ---- COO: equal runs of entries, private full-length y, reduced at the end ------------------
// (the reference splits by row range and silently assumes a row-sorted COO, :170-183; splitting
//  the entry stream works for any order)
void COOMatrixMatVectorNuma(const COOMatrix& A, const Vector& x, Vector& y, int nthreads)
{
    Trace tr("COOMatrixMatVectorNuma");
    const int G = gpu_count_for(nthreads);
    int dev0 = 0;
    thsp_get_device(&dev0);
    std::vector<NumaNode4COO> p(G);
    for (int i = 0; i < G; ++i) {
        int64_t start = 0, count = 0;
        ok(thsp_partition_rows(A.nnz, G, i, &start, &count), "entry partition");
        NumaNode4COO& b = p[i];
        b.alloc = i;
        b.core_ind = i;
        b.nnz = (int)count;
        b.start_row = 0;
        b.rows_per_node = A.nrow;
        use(i);
        b.sub_row_ind = dev_alloc<int>(count);
        b.sub_col_ind = dev_alloc<int>(count);
        b.sub_values = dev_alloc<double>(count);
        b.X = dev_alloc<double>(x.size);
        b.Y = dev_alloc<double>(A.nrow);
        if (count) {
            copy_bytes(b.sub_row_ind, A.row_ind + start, sizeof(int) * (size_t)count);
            copy_bytes(b.sub_col_ind, A.col_ind + start, sizeof(int) * (size_t)count);
            copy_bytes(b.sub_values, A.values + start, sizeof(double) * (size_t)count);
        }
        copy_bytes(b.X, x.values, sizeof(double) * (size_t)x.size);
        ok(thsp_memset(b.Y, 0, sizeof(double) * (size_t)A.nrow, nullptr), "memset");
    }
    timed_repeats(p, G, COOMatrixMatVectorNumaThread, "COO", (double)A.nnz);
    use(0);
    double* tmp = G > 1 ? dev_alloc<double>(A.nrow) : nullptr;
    for (int i = 1; i < G; ++i) reduce_into(p[0].Y, p[i].Y, A.nrow, tmp);
    copy_bytes(y.values, p[0].Y, sizeof(double) * (size_t)A.nrow);
    if (tmp) thsp_free(tmp);
    for (int i = 0; i < G; ++i) {
        use(i);
        thsp_free(p[i].sub_row_ind); thsp_free(p[i].sub_col_ind); thsp_free(p[i].sub_values); thsp_free(p[i].X); thsp_free(p[i].Y);
    }
    use(dev0);
}

// ---- CSC: column blocks, private full-length y (src/mat_vec.cpp:299-366), reduced at the end ----
void CSCMatrixMatVectorNuma(const CSCMatrix& A, const Vector& x, Vector& y, int nthreads)
{
    Trace tr("CSCMatrixMatVectorNuma");
    const int G = gpu_count_for(nthreads);
    int dev0 = 0;
    thsp_get_device(&dev0);
    std::vector<NumaNode4CSC> p(G);
    std::vector<int> cp_host((size_t)A.ncol + 1);
    copy_bytes(cp_host.data(), A.col_ptr, sizeof(int) * ((size_t)A.ncol + 1));
    for (int i = 0; i < G; ++i) {
        int64_t start = 0, count = 0;
        ok(thsp_partition_rows(A.ncol, G, i, &start, &count), "column partition");
        NumaNode4CSC& b = p[i];
        set_rows_total(&b, A.nrow);
        b.alloc = i;
        b.core_ind = i;
        b.start_col = (int)start;
        b.cols_per_node = (int)count;
        const int e0 = cp_host[start], e1 = cp_host[start + count];
        b.nnz = e1 - e0;
        use(i);
        b.sub_col_ptr = dev_alloc<int>((size_t)count + 1);
        b.sub_row_ind = dev_alloc<int>(b.nnz);
        b.sub_values = dev_alloc<double>(b.nnz);
        b.X = dev_alloc<double>(count);
        b.Y = dev_alloc<double>(A.nrow);
        std::vector<int> sub((size_t)count + 1);
        for (int64_t j = 0; j <= count; ++j) sub[j] = cp_host[start + j] - e0;
        copy_bytes(b.sub_col_ptr, sub.data(), sizeof(int) * sub.size());
        if (b.nnz) {
            copy_bytes(b.sub_row_ind, A.row_ind + e0, sizeof(int) * (size_t)b.nnz);
            copy_bytes(b.sub_values, A.values + e0, sizeof(double) * (size_t)b.nnz);
        }
        if (count) copy_bytes(b.X, x.values + start, sizeof(double) * (size_t)count);
        ok(thsp_memset(b.Y, 0, sizeof(double) * (size_t)A.nrow, nullptr), "memset");
    }
    timed_repeats(p, G, CSCMatrixMatVectorNumaThread, "CSC", (double)cp_host[A.ncol]);
    use(0);
    double* tmp = G > 1 ? dev_alloc<double>(A.nrow) : nullptr;
    for (int i = 1; i < G; ++i) reduce_into(p[0].Y, p[i].Y, A.nrow, tmp);
    copy_bytes(y.values, p[0].Y, sizeof(double) * (size_t)A.nrow);
    if (tmp) thsp_free(tmp);
    for (int i = 0; i < G; ++i) {
        use(i);
        thsp_free(p[i].sub_col_ptr); thsp_free(p[i].sub_row_ind); thsp_free(p[i].sub_values); thsp_free(p[i].X); thsp_free(p[i].Y);
        drop_rows_total(&p[i]);
    }
    use(dev0);
}

// ---- DIA: row blocks of the row-major diagonals (src/mat_vec.cpp:428-484) ----------------------
void DIAMatrixMatVectorNuma(const DIAMatrix& A, const Vector& x, Vector& y, int nthreads)
{
    Trace tr("DIAMatrixMatVectorNuma");
    const int G = gpu_count_for(nthreads);
    int dev0 = 0;
    thsp_get_device(&dev0);
    std::vector<NumaNode4DIA> p(G);
    for (int i = 0; i < G; ++i) {
        int64_t start = 0, count = 0;
        ok(thsp_partition_rows(A.nrow, G, i, &start, &count), "row partition");
        NumaNode4DIA& b = p[i];
        set_rows_total(&b, A.nrow);
        b.alloc = i;
        b.core_ind = i;
        b.start_row = (int)start;
        b.rows_per_node = (int)count;
        b.ndiags = A.ndiags;
        use(i);
        b.offsets = dev_alloc<int>(A.ndiags);
        b.values = dev_alloc<double>((size_t)count * A.ndiags);
        b.X = dev_alloc<double>(x.size);
        b.Y = dev_alloc<double>(count);
        if (A.ndiags) copy_bytes(b.offsets, A.offsets, sizeof(int) * (size_t)A.ndiags);
        if (count && A.ndiags) copy_bytes(b.values, A.values + (size_t)start * A.ndiags, sizeof(double) * (size_t)count * A.ndiags);
        copy_bytes(b.X, x.values, sizeof(double) * (size_t)x.size);
        ok(thsp_memset(b.Y, 0, sizeof(double) * (size_t)count, nullptr), "memset");
    }
    timed_repeats(p, G, DIAMatrixMatVectorNumaThread, "DIA", (double)A.nnz);
    for (int i = 0; i < G; ++i) {
        use(i);
        copy_bytes(y.values + p[i].start_row, p[i].Y, sizeof(double) * (size_t)p[i].rows_per_node);  // as the reference does (:472-477)
        thsp_free(p[i].offsets); thsp_free(p[i].values); thsp_free(p[i].X); thsp_free(p[i].Y);
        drop_rows_total(&p[i]);
    }
    use(dev0);
}
