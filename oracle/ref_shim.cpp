// ref_shim.cpp -- extern "C" doorway into the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).
//
// oracle/Makefile compiles /root/reference/src/*.cpp where they lie (with oracle/shim/numa.h
// standing in for the absent libnuma) and links them with this file into
// oracle/_ref/libref.so.  Nothing here re-implements the reference: every function builds
// the reference's own classes around caller arrays, calls the reference's own function and
// detaches the arrays again before the destructors (which delete[]) run.
//
// Used by: tests/golden/make_golden.py (fixture generation), tests/ (pinning the oracle),
// bench.py's cpu_baseline / --impl reference legs.  Never by the product.
#include <math.h>
#include <omp.h>
#include <pthread.h>
#include <string.h>
#include <sys/time.h>

#include <new>

#include "data_io.h"
#include "mat_vec.h"
#include "matrix.h"
#include "vec_vec.h"
#include "vector.h"

namespace {

double now_s()
{
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}

// Borrow caller memory inside a reference object; release() must run before the dtor.
struct VecView {
    Vector v;
    VecView(int n, const double* p) { v.size = n; v.values = const_cast<double*>(p); }
    ~VecView() { v.size = 0; v.values = nullptr; }
};
struct CooView {
    COOMatrix m;
    CooView(int nr, int nc, int nnz, const int* ri, const int* ci, const double* va)
    {
        m.nrow = nr; m.ncol = nc; m.nnz = nnz;
        m.row_ind = const_cast<int*>(ri); m.col_ind = const_cast<int*>(ci); m.values = const_cast<double*>(va);
    }
    ~CooView() { m.row_ind = nullptr; m.col_ind = nullptr; m.values = nullptr; }
};
struct CsrView {
    CSRMatrix m;
    CsrView(int nr, int nc, const int* rp, const int* ci, const double* va)
    {
        m.nrow = nr; m.ncol = nc;
        m.row_ptr = const_cast<int*>(rp); m.col_ind = const_cast<int*>(ci); m.values = const_cast<double*>(va);
        m.diagonal = nullptr;
    }
    ~CsrView() { m.row_ptr = nullptr; m.col_ind = nullptr; m.values = nullptr; m.diagonal = nullptr; }
};

// The reference's ELL constructor keeps an nrow-int VLA on the stack (src/matrix.cpp:457);
// run anything that may reach it on a thread with a roomy stack.
struct BigStackCall {
    void (*fn)(void*);
    void* arg;
};
void* big_stack_tramp(void* p)
{
    BigStackCall* c = static_cast<BigStackCall*>(p);
    c->fn(c->arg);
    return nullptr;
}
void on_big_stack(void (*fn)(void*), void* arg, size_t bytes)
{
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setstacksize(&at, bytes);
    BigStackCall c{fn, arg};
    pthread_t t;
    pthread_create(&t, &at, big_stack_tramp, &c);
    pthread_join(t, nullptr);
    pthread_attr_destroy(&at);
}

}  // namespace

extern "C" {

void ref_set_threads(int n) { omp_set_num_threads(n); }
int ref_max_threads(void) { return omp_get_max_threads(); }

// ---- SpMV: y += A x through the reference's own entry points (src/mat_vec.cpp:18-146)
void ref_coo_spmv(int nrow, int ncol, int nnz, const int* ri, const int* ci, const double* v, const double* x, double* y)
{
    CooView A(nrow, ncol, nnz, ri, ci, v);
    VecView X(ncol, x), Y(nrow, y);
    COOMatirxMatVector(A.m, X.v, Y.v);
}
void ref_csr_spmv(int nrow, int ncol, const int* rp, const int* ci, const double* v, const double* x, double* y)
{
    CsrView A(nrow, ncol, rp, ci, v);
    VecView X(ncol, x), Y(nrow, y);
    CSRMatrixMatVector(A.m, X.v, Y.v);
}
void ref_csc_spmv(int nrow, int ncol, const int* cp, const int* ri, const double* v, const double* x, double* y)
{
    CSCMatrix A;
    A.nrow = nrow; A.ncol = ncol;
    A.col_ptr = const_cast<int*>(cp); A.row_ind = const_cast<int*>(ri); A.values = const_cast<double*>(v);
    VecView X(ncol, x), Y(nrow, y);
    CSCMatrixMatVector(A, X.v, Y.v);
    A.col_ptr = nullptr; A.row_ind = nullptr; A.values = nullptr;
}
void ref_ell_spmv(int nrow, int ncol, int width, const int* ci, const double* v, const double* x, double* y)
{
    ELLMatrix A;
    A.nrow = nrow; A.ncol = ncol; A.nnz = 0; A.nonzeros_in_row = width;
    A.col_ind = const_cast<int*>(ci); A.values = const_cast<double*>(v); A.diagonal = nullptr;
    VecView X(ncol, x), Y(nrow, y);
    ELLMatrixMatVector(A, X.v, Y.v);
    A.col_ind = nullptr; A.values = nullptr; A.diagonal = nullptr;
}
void ref_dia_spmv(int nrow, int ncol, int ndiags, const int* off, const double* v, const double* x, double* y)
{
    DIAMatrix A;
    A.nrow = nrow; A.ncol = ncol; A.ndiags = ndiags; A.nnz = 0;
    A.offsets = const_cast<int*>(off); A.values = const_cast<double*>(v);
    VecView X(ncol, x), Y(nrow, y);
    DIAMatrixMatVector(A, X.v, Y.v);
    A.offsets = nullptr; A.values = nullptr;
}

// ---- conversions through the reference's converting constructors (src/matrix.cpp)
// `diagonal` receives min(#diag entries, nrow) values; the count is returned.
int ref_coo2csr(int nrow, int ncol, int nnz, const int* ri, const int* ci, const double* v, int* row_ptr, int* col_ind,
                double* values, double* diagonal)
{
    CooView A(nrow, ncol, nnz, ri, ci, v);
    int ndiag = 0;
    for (int k = 0; k < nnz; ++k) ndiag += (ri[k] == ci[k]);
    if (ndiag > nrow) return -1;  // the reference would overrun diagonal[]
    CSRMatrix B(A.m);
    memcpy(row_ptr, B.row_ptr, sizeof(int) * ((size_t)nrow + 1));
    memcpy(col_ind, B.col_ind, sizeof(int) * (size_t)nnz);
    memcpy(values, B.values, sizeof(double) * (size_t)nnz);
    if (diagonal) memcpy(diagonal, B.diagonal, sizeof(double) * (size_t)ndiag);
    return ndiag;
}
void ref_coo2csc(int nrow, int ncol, int nnz, const int* ri, const int* ci, const double* v, int* col_ptr, int* row_ind,
                 double* values)
{
    CooView A(nrow, ncol, nnz, ri, ci, v);
    CSCMatrix B(A.m);
    memcpy(col_ptr, B.col_ptr, sizeof(int) * ((size_t)ncol + 1));
    memcpy(row_ind, B.row_ind, sizeof(int) * (size_t)nnz);
    memcpy(values, B.values, sizeof(double) * (size_t)nnz);
}

struct EllJob {
    int nrow, ncol, nnz;
    const int *ri, *ci;
    const double* v;
    int cap_slots;  // capacity of the output slab in slots (nrow*width the caller expects), or 0 = width query
    int* col_ind;
    double* values;
    double* diagonal;
    int width, ndiag;
};
static void ell_job(void* p)
{
    EllJob* j = static_cast<EllJob*>(p);
    CooView A(j->nrow, j->ncol, j->nnz, j->ri, j->ci, j->v);
    ELLMatrix D(A.m);
    j->width = D.nonzeros_in_row;
    size_t total = (size_t)D.nrow * (size_t)D.nonzeros_in_row;
    if (j->col_ind && total <= (size_t)j->cap_slots) {
        memcpy(j->col_ind, D.col_ind, sizeof(int) * total);
        memcpy(j->values, D.values, sizeof(double) * total);
        if (j->diagonal) memcpy(j->diagonal, D.diagonal, sizeof(double) * (size_t)j->ndiag);
    }
}
// Returns the slab width K chosen by the reference; fills the outputs when cap_slots >= nrow*K.
int ref_coo2ell(int nrow, int ncol, int nnz, const int* ri, const int* ci, const double* v, int cap_slots, int* col_ind,
                double* values, double* diagonal)
{
    EllJob j{nrow, ncol, nnz, ri, ci, v, cap_slots, col_ind, values, diagonal, 0, 0};
    for (int k = 0; k < nnz; ++k) j.ndiag += (ri[k] == ci[k]);
    if (j.ndiag > nrow) return -1;
    on_big_stack(ell_job, &j, (size_t)nrow * sizeof(int) + (64u << 20));
    return j.width;
}

// DIA from CSR.  Pass values == NULL to learn ndiags first.  Refuses the (0, ncol-1) corner
// entry, on which the reference writes out of bounds.
int ref_csr2dia(int nrow, int ncol, const int* rp, const int* ci, const double* v, int cap_diags, int* offsets, double* values)
{
    for (int p = rp[0]; p < rp[1] && nrow > 0; ++p)
        if (ci[p] == ncol - 1) return -1;
    CsrView A(nrow, ncol, rp, ci, v);
    DIAMatrix E(A.m);
    int nd = E.ndiags;
    if (offsets && values && nd <= cap_diags) {
        memcpy(offsets, E.offsets, sizeof(int) * (size_t)nd);
        memcpy(values, E.values, sizeof(double) * (size_t)nd * (size_t)nrow);
    }
    // DIAMatrix mallocs and its Free() delete[]s: release with free() ourselves.
    free(E.offsets); free(E.values);
    E.offsets = nullptr; E.values = nullptr;
    return nd;
}

// ---- vector kernels (src/vec_vec.cpp, src/vector.cpp)
double ref_dot(int n, const double* x, const double* y)
{
    VecView X(n, x), Y(n, y);
    return vec_dot(X.v, Y.v);
}
void ref_axpby(int n, double alpha, const double* x, double beta, const double* y, double* w)
{
    VecView X(n, x), Y(n, y), W(n, w);
    vec_axpby(alpha, X.v, beta, Y.v, W.v);
}
void ref_fill(int n, double a, double* v) { VecView V(n, v); V.v.Fill(a); }
void ref_scale(int n, double a, double* v) { VecView V(n, v); V.v.Scale(a); }
void ref_shift(int n, double a, double* v) { VecView V(n, v); V.v.Shift(a); }
void ref_copy(int n, const double* x, double* v) { VecView X(n, x), V(n, v); V.v.Copy(X.v); }
void ref_add_scaled(int n, double a, const double* x, double* v) { VecView X(n, x), V(n, v); V.v.AddScaled(a, X.v); }
void ref_add2_scaled(int n, double a, const double* x, double b, const double* y, double* v)
{
    VecView X(n, x), Y(n, y), V(n, v);
    V.v.Add2Scaled(a, X.v, b, Y.v);
}
int ref_check_vector(int nx, const double* x, int ny, const double* y)
{
    VecView X(nx, x), Y(ny, y);
    return checkVector(X.v, Y.v) ? 1 : 0;
}
void ref_fill_random(int n, double* v) { VecView V(n, v); V.v.FillRandom(); }

// ---- Matrix Market reader (src/data_io.cpp:45-105): returns sizes, copies out when asked.
int ref_coo_read(const char* path, int* nrow, int* ncol, int* nnz, int cap, int* ri, int* ci, double* v)
{
    COOMatrix A;
    COOMatrixRead(path, A);
    *nrow = A.nrow; *ncol = A.ncol; *nnz = A.nnz;
    if (ri && A.nnz <= cap) {
        memcpy(ri, A.row_ind, sizeof(int) * (size_t)A.nnz);
        memcpy(ci, A.col_ind, sizeof(int) * (size_t)A.nnz);
        memcpy(v, A.values, sizeof(double) * (size_t)A.nnz);
    }
    return 0;
}

// ---- CPU baseline timing with the reference's own protocol (main.cpp:54-61): `reps`
// back-to-back calls between two clock reads, y not re-zeroed.  Returns seconds per call.
double ref_time_csr_spmv(int nrow, int ncol, const int* rp, const int* ci, const double* v, const double* x, double* y, int reps)
{
    CsrView A(nrow, ncol, rp, ci, v);
    VecView X(ncol, x), Y(nrow, y);
    CSRMatrixMatVector(A.m, X.v, Y.v);  // untimed: page-touch / thread-team start
    double t0 = now_s();
    for (int i = 0; i < reps; ++i) CSRMatrixMatVector(A.m, X.v, Y.v);
    return (now_s() - t0) / reps;
}
double ref_time_ell_spmv(int nrow, int ncol, int width, const int* ci, const double* v, const double* x, double* y, int reps)
{
    ELLMatrix A;
    A.nrow = nrow; A.ncol = ncol; A.nnz = 0; A.nonzeros_in_row = width;
    A.col_ind = const_cast<int*>(ci); A.values = const_cast<double*>(v); A.diagonal = nullptr;
    VecView X(ncol, x), Y(nrow, y);
    ELLMatrixMatVector(A, X.v, Y.v);
    double t0 = now_s();
    for (int i = 0; i < reps; ++i) ELLMatrixMatVector(A, X.v, Y.v);
    double dt = (now_s() - t0) / reps;
    A.col_ind = nullptr; A.values = nullptr; A.diagonal = nullptr;
    return dt;
}
// One power-iteration step composed from reference calls only:
// y=0; y+=A x; nrm=sqrt(dot(y,y)); x = (1/nrm) y.  Returns seconds per step, *nrm_out = last norm.
double ref_time_power_iteration(int n, const int* rp, const int* ci, const double* v, double* x, double* y, int steps, double* nrm_out)
{
    CsrView A(n, n, rp, ci, v);
    VecView X(n, x), Y(n, y);
    double nrm = 0.0;
    double t0 = now_s();
    for (int s = 0; s < steps; ++s) {
        Y.v.Fill(0.0);
        CSRMatrixMatVector(A.m, X.v, Y.v);
        nrm = sqrt(vec_dot(Y.v, Y.v));
        vec_axpby(1.0 / nrm, Y.v, 0.0, Y.v, X.v);
    }
    double dt = (now_s() - t0) / (steps > 0 ? steps : 1);
    if (nrm_out) *nrm_out = nrm;
    return dt;
}

}  // extern "C"
