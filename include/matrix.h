// matrix.h -- sparse matrix containers of the arm-spmv API, backed by the B200 library.
//
// Source-compatible with the reference's include/matrix.h:7-138: every public field, every
// constructor / method signature is kept, so main.cpp and other callers recompile unchanged.
// What changed underneath:
//   * arrays allocated by the library (converting constructors, copies, COOMatrixRead) live in
//     CUDA managed memory: host code may still index them, kernels stream them from HBM;
//   * pointer-taking constructors adopt caller memory exactly like the reference
//     (src/matrix.cpp:12-15,88-91); Free()/destructors release managed storage with cudaFree
//     and adopted host storage with delete[];
//   * COO -> CSR / CSC / ELL and CSR -> DIA run on the GPU (thsp_coo2csr, thsp_coo2csc,
//     thsp_coo2ell, thsp_csr2dia_* in thsp.h) and produce the same arrays bit for bit;
//   * operator=(const COOMatrix&) gives the converting constructor's result (the reference's
//     assignment versions leave row_ptr[nrow] / col_ptr[ncol] uninitialised).
#ifndef MATRIX_H
#define MATRIX_H

#include <stdio.h>
#include <stdlib.h>

// Coordinate triples, any order, duplicates allowed.
class COOMatrix {
public:
    int nrow;
    int ncol;
    int nnz;

    int*    row_ind;
    int*    col_ind;
    double* values;

    COOMatrix();
    COOMatrix(int n, int m, int nnz, int* row_ind, int* col_ind, double* values);
    COOMatrix(const COOMatrix& A);
    ~COOMatrix();
    COOMatrix& operator=(const COOMatrix& A);

    void Free();
};

// Compressed rows.  `diagonal` holds the row==col values packed in COO order (first nrow at most).
class CSRMatrix {
public:
    int nrow;
    int ncol;

    int*    row_ptr;
    int*    col_ind;
    double* values;
    double* diagonal;

    CSRMatrix();
    CSRMatrix(int n, int m, int* row_ptr, int* col_ind, double* values, double* diagonal);
    CSRMatrix(const CSRMatrix& A);
    CSRMatrix(const COOMatrix& A);   // stable by row: entries keep their COO order inside a row
    ~CSRMatrix();
    CSRMatrix& operator=(const CSRMatrix& A);
    CSRMatrix& operator=(const COOMatrix& A);

    void Free();
};

// Compressed columns.
class CSCMatrix {
public:
    int nrow;
    int ncol;

    int*    row_ind;
    int*    col_ptr;
    double* values;

    CSCMatrix();
    CSCMatrix(int n, int m, int* row_ind, int* col_ptr, double* values);
    CSCMatrix(const CSCMatrix& A);
    CSCMatrix(const COOMatrix& A);
    ~CSCMatrix();
    CSCMatrix& operator=(const CSCMatrix& A);
    CSCMatrix& operator=(const COOMatrix& A);

    void Free();
};

// ELLPACK, COLUMN-major slab: slot k of row i is element [i + k*nrow]; padding is (col 0, 0.0).
class ELLMatrix {
public:
    int nrow;
    int ncol;
    int nnz;
    int nonzeros_in_row;

    int*    col_ind;
    double* values;
    double* diagonal;

    ELLMatrix();
    ELLMatrix(int n, int m, int nnz, int nonzeros_in_row, int* col_ind, double* values, double* diagonal);
    ELLMatrix(const ELLMatrix& A);
    ELLMatrix(const COOMatrix& A);
    ~ELLMatrix();
    ELLMatrix& operator=(const ELLMatrix& A);
    ELLMatrix& operator=(const COOMatrix& A);

    void Free();
};

// Declared by the reference (include/matrix.h:95-115) but only partly defined there
// (src/matrix.cpp:619-632: no copy constructor, destructor, assignment or Free), so nothing can
// use it.  Kept declaration-compatible; the three members the reference defines exist here too.
class BlockMatrix {
public:
    int nrow;
    int ncol;
    int nnz;
    int nblocks;

    int*     block_size;
    int*     row_ind;
    int*     col_ind;
    double** values;

    BlockMatrix();
    BlockMatrix(int n, int m, int nnz, int nblocks, int* block_size, int* row_ind, int* col_ind, double** values);
    BlockMatrix(const BlockMatrix& A);
    BlockMatrix(const COOMatrix& A);
    ~BlockMatrix();
    BlockMatrix& operator=(const BlockMatrix& A);
    BlockMatrix& operator=(const COOMatrix& A);

    void Free();
};

// Diagonals, ROW-major: values[i*ndiags + d] is the entry of row i on diagonal offsets[d].
class DIAMatrix {
public:
    int nnz;
    int nrow;
    int ncol;
    int ndiags;

    int*    offsets;
    double* values;

    DIAMatrix();
    DIAMatrix(int n, int m, int ndiags, int* offsets, double* values);
    DIAMatrix(const DIAMatrix& A);
    DIAMatrix(const CSRMatrix& A);
    ~DIAMatrix();
    DIAMatrix& operator=(const DIAMatrix& A);
    DIAMatrix& operator=(const CSRMatrix& A);

    void Free();
};

#endif  // MATRIX_H
