// mat_vec.h -- SpMV entry points of the arm-spmv API (reference include/mat_vec.h:7-23).
//
// Every function computes y += A*x and returns when y is complete (the reference's callers
// bracket these calls with mytimer(), main.cpp:56-59).  Underneath they launch the sm_100a
// kernels declared in thsp.h; there is no CPU code path.
#ifndef MAT_VEC_H
#define MAT_VEC_H

#include "matrix.h"
#include "vector.h"

// "Matirx" is the reference's spelling and part of its link interface.
void COOMatirxMatVector(const COOMatrix& A, const Vector& x, Vector& y);
void CSRMatrixMatVector(const CSRMatrix& A, const Vector& x, Vector& y);
void CSCMatrixMatVector(const CSCMatrix& A, const Vector& x, Vector& y);
void ELLMatrixMatVector(const ELLMatrix& A, const Vector& x, Vector& y);
void DIAMatrixMatVector(const DIAMatrix& A, const Vector& x, Vector& y);

// Partitioned variants.  The reference splits rows (columns for CSC) into `nthreads` equal
// blocks placed on NUMA nodes, runs 50 repeats and prints "### <FMT> NUMA GFLOPS = ...".
// Here the blocks are placed on min(nthreads, #GPUs) B200s with x replicated; the same line
// is printed.  Unlike the reference (which drops the result for four of the five formats,
// SURVEY.md A.3) the accumulated blocks are written back into y.
void COOMatrixMatVectorNuma(const COOMatrix& A, const Vector& x, Vector& y, int nthreads);
void CSRMatrixMatVectorNuma(const CSRMatrix& A, const Vector& x, Vector& y, int nthreads);
void CSCMatrixMatVectorNuma(const CSCMatrix& A, const Vector& x, Vector& y, int nthreads);
void ELLMatrixMatVectorNuma(const ELLMatrix& A, const Vector& x, Vector& y, int nthreads);
void DIAMatrixMatVectorNuma(const DIAMatrix& A, const Vector& x, Vector& y, int nthreads);

// pthread bodies of the reference's partitioned variants; `args` is a NumaNode4* (numa_node.h)
// whose arrays are device pointers here.  Each runs one y_block += A_block * x on its GPU.
void* COOMatrixMatVectorNumaThread(void* args);
void* CSRMatrixMatVectorNumaThread(void* args);
void* CSCMatrixMatVectorNumaThread(void* args);
void* ELLMatrixMatVectorNumaThread(void* args);
void* DIAMatrixMatVectorNumaThread(void* args);

#endif  // MAT_VEC_H
