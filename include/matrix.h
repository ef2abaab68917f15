// matrix.h -- sparse matrix containers of the arm-spmv API, backed by the B200 library.
//
// Source-compatible with the reference's include/matrix.h:7-138: every public field (name, type,
// order) and every constructor / method signature is kept, so main.cpp and other callers
// recompile unchanged.  What changed underneath:
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

// ---------------------------------------------------------------------------------------------
// COO: coordinate triples in any order, duplicates allowed (they add up in a product).
class COOMatrix {
public:
    int nrow, ncol, nnz;              // shape and number of stored triples
    int *row_ind, *col_ind;           // [nnz] each, 0-based
    double* values;                   // [nnz]

    ~COOMatrix();
    COOMatrix();
    COOMatrix(const COOMatrix& other);                                                   // deep copy
    COOMatrix(int rows, int cols, int entries, int* rows_of, int* cols_of, double* vals);  // adopts the arrays
    COOMatrix& operator=(const COOMatrix& other);
    void Free();
};

// ---------------------------------------------------------------------------------------------
// CSR: entries of row i are [row_ptr[i], row_ptr[i+1]).  `diagonal` holds the row==col values
// packed in COO order (at most nrow of them) - the reference keeps it "for SymGS".
class CSRMatrix {
public:
    int nrow, ncol;
    int *row_ptr, *col_ind;           // [nrow+1], [nnz]
    double *values, *diagonal;        // [nnz], [nrow]

    ~CSRMatrix();
    CSRMatrix();
    CSRMatrix(const COOMatrix& coo);  // stable by row: inside a row entries keep their COO order
    CSRMatrix(const CSRMatrix& other);
    CSRMatrix(int rows, int cols, int* ptr, int* cols_of, double* vals, double* diag);    // adopts the arrays
    CSRMatrix& operator=(const COOMatrix& coo);
    CSRMatrix& operator=(const CSRMatrix& other);
    void Free();
};

// ---------------------------------------------------------------------------------------------
// CSC: entries of column j are [col_ptr[j], col_ptr[j+1]).
class CSCMatrix {
public:
    int nrow, ncol;
    int *row_ind, *col_ptr;           // [nnz], [ncol+1]
    double* values;                   // [nnz]

    ~CSCMatrix();
    CSCMatrix();
    CSCMatrix(const COOMatrix& coo);  // stable by column
    CSCMatrix(const CSCMatrix& other);
    CSCMatrix(int rows, int cols, int* rows_of, int* ptr, double* vals);                  // adopts the arrays
    CSCMatrix& operator=(const COOMatrix& coo);
    CSCMatrix& operator=(const CSCMatrix& other);
    void Free();
};

// ---------------------------------------------------------------------------------------------
// ELLPACK, COLUMN-major slab: slot k of row i is element [i + k*nrow]; padding is (col 0, 0.0).
class ELLMatrix {
public:
    int nrow, ncol, nnz;
    int nonzeros_in_row;              // slab width = longest row
    int* col_ind;                     // [nrow * nonzeros_in_row]
    double *values, *diagonal;        // [nrow * nonzeros_in_row], [nrow]

    ~ELLMatrix();
    ELLMatrix();
    ELLMatrix(const COOMatrix& coo);
    ELLMatrix(const ELLMatrix& other);
    ELLMatrix(int rows, int cols, int entries, int width, int* cols_of, double* vals, double* diag);   // adopts the arrays
    ELLMatrix& operator=(const COOMatrix& coo);
    ELLMatrix& operator=(const ELLMatrix& other);
    void Free();
};

// ---------------------------------------------------------------------------------------------
// Declared by the reference (include/matrix.h:95-115) but only partly defined there
// (src/matrix.cpp:619-632: no copy constructor, destructor, assignment or Free), so nothing can
// use it.  Kept declaration-compatible; the three members the reference defines exist here too.
class BlockMatrix {
public:
    int nrow, ncol, nnz, nblocks;
    int *block_size, *row_ind, *col_ind;
    double** values;

    ~BlockMatrix();
    BlockMatrix();
    BlockMatrix(const COOMatrix& coo);
    BlockMatrix(const BlockMatrix& other);
    BlockMatrix(int rows, int cols, int entries, int blocks, int* sizes, int* rows_of, int* cols_of, double** vals);
    BlockMatrix& operator=(const COOMatrix& coo);
    BlockMatrix& operator=(const BlockMatrix& other);
    void Free();
};

// ---------------------------------------------------------------------------------------------
// DIA, ROW-major: values[i*ndiags + d] is the entry of row i on diagonal offsets[d] (ascending).
class DIAMatrix {
public:
    int nnz, nrow, ncol;
    int ndiags;                       // number of occupied diagonals
    int* offsets;                     // [ndiags], column minus row
    double* values;                   // [nrow * ndiags]

    ~DIAMatrix();
    DIAMatrix();
    DIAMatrix(const CSRMatrix& csr);  // a duplicate (i,j) overwrites the earlier one
    DIAMatrix(const DIAMatrix& other);
    DIAMatrix(int rows, int cols, int diagonals, int* offs, double* vals);                // adopts the arrays
    DIAMatrix& operator=(const CSRMatrix& csr);
    DIAMatrix& operator=(const DIAMatrix& other);
    void Free();
};

#endif  // MATRIX_H
