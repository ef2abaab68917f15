// matrix.cpp -- COO/CSR/CSC/ELL/DIA containers of the arm-spmv API (replaces src/matrix.cpp).
//
// Containers only hold pointers; all conversion work is done by the GPU kernels behind thsp.h.
// Copies go through cudaMemcpy(Default) so they work for managed, device and host sources.
#include "matrix.h"

#include <string.h>

#include "hostmem.h"

using namespace thsp_host;

namespace {

// Device-usable arrays of a COO matrix: in place when the library owns them, else through the device mirror that the
// first conversion uploads and the next ones (CSR, CSC, ELL of the same matrix: main.cpp:38-41) find again.
struct CooArrays {
    const int *ri, *ci;
    const double* va;
    CooArrays(const COOMatrix& A) : ri(mirror(A.row_ind, A.nnz)), ci(mirror(A.col_ind, A.nnz)), va(mirror(A.values, A.nnz)) {}
};

void csr_from_coo(CSRMatrix& B, const COOMatrix& A)
{
    Trace tr("CSRMatrix(COOMatrix)");
    B.nrow = A.nrow;
    B.ncol = A.ncol;
    B.row_ptr = alloc_matrix<int>((size_t)A.nrow + 1);
    B.col_ind = alloc_matrix<int>(A.nnz);
    B.values = alloc_matrix<double>(A.nnz);
    B.diagonal = alloc_matrix<double>(A.nrow);
    CooArrays in(A);
    ok(thsp_coo2csr(A.nrow, A.ncol, A.nnz, in.ri, in.ci, in.va, B.row_ptr, B.col_ind, B.values, B.diagonal, nullptr, nullptr),
       "COO -> CSR");
    sync();
    // The arrays were written by kernels: they ARE in HBM, the first product need not prefetch them; and the plan
    // (row-length histogram -> kernel) is made here, once, instead of inside the first timed product (main.cpp:66-71).
    first_gpu_use(B.row_ptr), first_gpu_use(B.col_ind), first_gpu_use(B.values);
    if (A.nrow > 0) csr_plan(A.nrow, A.ncol, A.nnz, B.row_ptr, B.col_ind, B.values);
}

void csc_from_coo(CSCMatrix& C, const COOMatrix& A)
{
    Trace tr("CSCMatrix(COOMatrix)");
    C.nrow = A.nrow;
    C.ncol = A.ncol;
    C.col_ptr = alloc_matrix<int>((size_t)A.ncol + 1);
    C.row_ind = alloc_matrix<int>(A.nnz);
    C.values = alloc_matrix<double>(A.nnz);
    CooArrays in(A);
    ok(thsp_coo2csc(A.nrow, A.ncol, A.nnz, in.ri, in.ci, in.va, C.col_ptr, C.row_ind, C.values, nullptr), "COO -> CSC");
    sync();
    first_gpu_use(C.values);   // written by kernels: already in HBM (CSCMatrixMatVector prefetches on first use otherwise)
}

void ell_from_coo(ELLMatrix& D, const COOMatrix& A)
{
    Trace tr("ELLMatrix(COOMatrix)");
    D.nrow = A.nrow;
    D.ncol = A.ncol;
    D.nnz = A.nnz;
    CooArrays in(A);
    int width = 0;
    ok(thsp_coo2ell_prepare(A.nrow, A.ncol, A.nnz, in.ri, in.ci, in.va, &width, nullptr), "ELL width");
    D.nonzeros_in_row = width;
    const size_t total = (size_t)A.nrow * (size_t)width;
    D.col_ind = alloc_matrix<int>(total);
    D.values = alloc_matrix<double>(total);
    D.diagonal = alloc_matrix<double>(A.nrow);
    ok(thsp_coo2ell(A.nrow, A.ncol, A.nnz, in.ri, in.ci, in.va, width, D.col_ind, D.values, D.diagonal, nullptr, nullptr),
       "COO -> ELL");
    sync();
    first_gpu_use(D.values);
}

void dia_from_csr(DIAMatrix& E, const CSRMatrix& A)
{
    Trace tr("DIAMatrix(CSRMatrix)");
    const int nnz = A.nrow > 0 ? peek_int(A.row_ptr + A.nrow) : 0;
    E.nnz = nnz;
    E.nrow = A.nrow;
    E.ncol = A.ncol;
    const int* rp = mirror(A.row_ptr, (size_t)A.nrow + 1);
    const int* ci = mirror(A.col_ind, (size_t)nnz);
    const double* va = mirror(A.values, (size_t)nnz);
    int nd = 0;
    ok(thsp_csr2dia_offsets(A.nrow, A.ncol, rp, ci, &nd, nullptr, 0, nullptr), "CSR -> DIA (count)");
    E.ndiags = nd;
    E.offsets = alloc_matrix<int>(nd);
    E.values = alloc_matrix<double>((size_t)A.nrow * (size_t)nd);
    ok(thsp_csr2dia_offsets(A.nrow, A.ncol, rp, ci, &nd, E.offsets, nd, nullptr), "CSR -> DIA (offsets)");
    ok(thsp_csr2dia_fill(A.nrow, A.ncol, rp, ci, va, nd, E.offsets, E.values, nullptr), "CSR -> DIA (fill)");
    sync();
    first_gpu_use(E.values);
}

}  // namespace

// ------------------------------------------------------------------------------ COO ------
COOMatrix::COOMatrix() : nrow(0), ncol(0), nnz(0), row_ind(nullptr), col_ind(nullptr), values(nullptr) {}

COOMatrix::COOMatrix(int n, int m, int nz, int* ri, int* ci, double* va) : nrow(n), ncol(m), nnz(nz), row_ind(ri), col_ind(ci), values(va) {}

COOMatrix::COOMatrix(const COOMatrix& A)
    : nrow(A.nrow), ncol(A.ncol), nnz(A.nnz), row_ind(alloc_matrix<int>(A.nnz)), col_ind(alloc_matrix<int>(A.nnz)), values(alloc_matrix<double>(A.nnz))
{
    copy(row_ind, A.row_ind, (size_t)nnz);
    copy(col_ind, A.col_ind, (size_t)nnz);
    copy(values, A.values, (size_t)nnz);
}

COOMatrix::~COOMatrix() { Free(); }

COOMatrix& COOMatrix::operator=(const COOMatrix& A)
{
    if (this == &A) return *this;
    Free();
    nrow = A.nrow;
    ncol = A.ncol;
    nnz = A.nnz;
    row_ind = alloc_matrix<int>(nnz);
    col_ind = alloc_matrix<int>(nnz);
    values = alloc_matrix<double>(nnz);
    copy(row_ind, A.row_ind, (size_t)nnz);
    copy(col_ind, A.col_ind, (size_t)nnz);
    copy(values, A.values, (size_t)nnz);
    return *this;
}

void COOMatrix::Free()
{
    release(row_ind);
    release(col_ind);
    release(values);
    nrow = ncol = nnz = 0;
}

// ------------------------------------------------------------------------------ CSR ------
CSRMatrix::CSRMatrix() : nrow(0), ncol(0), row_ptr(nullptr), col_ind(nullptr), values(nullptr), diagonal(nullptr) {}

CSRMatrix::CSRMatrix(int n, int m, int* rp, int* ci, double* va, double* dg)
    : nrow(n), ncol(m), row_ptr(rp), col_ind(ci), values(va), diagonal(dg)
{
}

CSRMatrix::CSRMatrix(const CSRMatrix& A) : nrow(0), ncol(0), row_ptr(nullptr), col_ind(nullptr), values(nullptr), diagonal(nullptr)
{
    *this = A;
}

CSRMatrix::CSRMatrix(const COOMatrix& A) : nrow(0), ncol(0), row_ptr(nullptr), col_ind(nullptr), values(nullptr), diagonal(nullptr)
{
    csr_from_coo(*this, A);
}

CSRMatrix::~CSRMatrix() { Free(); }

CSRMatrix& CSRMatrix::operator=(const CSRMatrix& A)
{
    if (this == &A) return *this;
    Free();
    nrow = A.nrow;
    ncol = A.ncol;
    const int nnz = (A.row_ptr && A.nrow >= 0) ? peek_int(A.row_ptr + A.nrow) : 0;
    row_ptr = alloc_matrix<int>((size_t)nrow + 1);
    col_ind = alloc_matrix<int>(nnz);
    values = alloc_matrix<double>(nnz);
    diagonal = alloc_matrix<double>(nrow);
    copy(row_ptr, A.row_ptr, (size_t)nrow + 1);
    copy(col_ind, A.col_ind, (size_t)nnz);
    copy(values, A.values, (size_t)nnz);
    if (A.diagonal) copy(diagonal, A.diagonal, (size_t)nrow);
    return *this;
}

CSRMatrix& CSRMatrix::operator=(const COOMatrix& A)
{
    Free();
    csr_from_coo(*this, A);
    return *this;
}

void CSRMatrix::Free()
{
    forget_plans(row_ptr);
    release(row_ptr);
    release(col_ind);
    release(values);
    release(diagonal);
    nrow = ncol = 0;
}

// ------------------------------------------------------------------------------ CSC ------
CSCMatrix::CSCMatrix() : nrow(0), ncol(0), row_ind(nullptr), col_ptr(nullptr), values(nullptr) {}

CSCMatrix::CSCMatrix(int n, int m, int* ri, int* cp, double* va) : nrow(n), ncol(m), row_ind(ri), col_ptr(cp), values(va) {}

CSCMatrix::CSCMatrix(const CSCMatrix& A) : nrow(0), ncol(0), row_ind(nullptr), col_ptr(nullptr), values(nullptr) { *this = A; }

CSCMatrix::CSCMatrix(const COOMatrix& A) : nrow(0), ncol(0), row_ind(nullptr), col_ptr(nullptr), values(nullptr) { csc_from_coo(*this, A); }

CSCMatrix::~CSCMatrix() { Free(); }

CSCMatrix& CSCMatrix::operator=(const CSCMatrix& A)
{
    if (this == &A) return *this;
    Free();
    nrow = A.nrow;
    ncol = A.ncol;
    const int nnz = A.col_ptr ? peek_int(A.col_ptr + A.ncol) : 0;
    col_ptr = alloc_matrix<int>((size_t)ncol + 1);
    row_ind = alloc_matrix<int>(nnz);
    values = alloc_matrix<double>(nnz);
    copy(col_ptr, A.col_ptr, (size_t)ncol + 1);
    copy(row_ind, A.row_ind, (size_t)nnz);
    copy(values, A.values, (size_t)nnz);
    return *this;
}

CSCMatrix& CSCMatrix::operator=(const COOMatrix& A)
{
    Free();
    csc_from_coo(*this, A);
    return *this;
}

void CSCMatrix::Free()
{
    release(col_ptr);
    release(row_ind);
    release(values);
    nrow = ncol = 0;
}

// ------------------------------------------------------------------------------ ELL ------
ELLMatrix::ELLMatrix() : nrow(0), ncol(0), nnz(0), nonzeros_in_row(0), col_ind(nullptr), values(nullptr), diagonal(nullptr) {}

ELLMatrix::ELLMatrix(int n, int m, int nz, int width, int* ci, double* va, double* dg)
    : nrow(n), ncol(m), nnz(nz), nonzeros_in_row(width), col_ind(ci), values(va), diagonal(dg)
{
}

ELLMatrix::ELLMatrix(const ELLMatrix& A) : nrow(0), ncol(0), nnz(0), nonzeros_in_row(0), col_ind(nullptr), values(nullptr), diagonal(nullptr)
{
    *this = A;
}

ELLMatrix::ELLMatrix(const COOMatrix& A) : nrow(0), ncol(0), nnz(0), nonzeros_in_row(0), col_ind(nullptr), values(nullptr), diagonal(nullptr)
{
    ell_from_coo(*this, A);
}

ELLMatrix::~ELLMatrix() { Free(); }

ELLMatrix& ELLMatrix::operator=(const ELLMatrix& A)
{
    if (this == &A) return *this;
    Free();
    nrow = A.nrow;
    ncol = A.ncol;
    nnz = A.nnz;
    nonzeros_in_row = A.nonzeros_in_row;
    const size_t total = (size_t)nrow * (size_t)nonzeros_in_row;
    col_ind = alloc_matrix<int>(total);
    values = alloc_matrix<double>(total);
    diagonal = alloc_matrix<double>(nrow);
    copy(col_ind, A.col_ind, total);
    copy(values, A.values, total);
    if (A.diagonal) copy(diagonal, A.diagonal, (size_t)nrow);
    return *this;
}

ELLMatrix& ELLMatrix::operator=(const COOMatrix& A)
{
    Free();
    ell_from_coo(*this, A);
    return *this;
}

void ELLMatrix::Free()
{
    release(col_ind);
    release(values);
    release(diagonal);
    nrow = ncol = nnz = nonzeros_in_row = 0;
}

// ---------------------------------------------------------------------------- Block ------
// The reference defines only these three members (src/matrix.cpp:619-632).
BlockMatrix::BlockMatrix() : nrow(0), ncol(0), nnz(0), nblocks(0), block_size(nullptr), row_ind(nullptr), col_ind(nullptr), values(nullptr) {}

BlockMatrix::BlockMatrix(int n, int m, int nz, int nb, int* bs, int* ri, int* ci, double** va)
    : nrow(n), ncol(m), nnz(nz), nblocks(nb), block_size(bs), row_ind(ri), col_ind(ci), values(va)
{
}

BlockMatrix::BlockMatrix(const COOMatrix& A)
    : nrow(A.nrow), ncol(A.ncol), nnz(A.nnz), nblocks(0), block_size(nullptr), row_ind(nullptr), col_ind(nullptr), values(nullptr)
{
}

// ------------------------------------------------------------------------------ DIA ------
DIAMatrix::DIAMatrix() : nnz(0), nrow(0), ncol(0), ndiags(0), offsets(nullptr), values(nullptr) {}

DIAMatrix::DIAMatrix(int n, int m, int nd, int* off, double* va) : nnz(0), nrow(n), ncol(m), ndiags(nd), offsets(off), values(va) {}

DIAMatrix::DIAMatrix(const DIAMatrix& A) : nnz(0), nrow(0), ncol(0), ndiags(0), offsets(nullptr), values(nullptr) { *this = A; }

DIAMatrix::DIAMatrix(const CSRMatrix& A) : nnz(0), nrow(0), ncol(0), ndiags(0), offsets(nullptr), values(nullptr) { dia_from_csr(*this, A); }

DIAMatrix::~DIAMatrix() { Free(); }

DIAMatrix& DIAMatrix::operator=(const DIAMatrix& A)
{
    if (this == &A) return *this;
    Free();
    nnz = A.nnz;
    nrow = A.nrow;
    ncol = A.ncol;
    ndiags = A.ndiags;
    offsets = alloc_matrix<int>(ndiags);
    values = alloc_matrix<double>((size_t)nrow * (size_t)ndiags);
    copy(offsets, A.offsets, (size_t)ndiags);
    copy(values, A.values, (size_t)nrow * (size_t)ndiags);
    return *this;
}

DIAMatrix& DIAMatrix::operator=(const CSRMatrix& A)
{
    Free();
    dia_from_csr(*this, A);
    return *this;
}

void DIAMatrix::Free()
{
    release(offsets);
    release(values);
}
