// mat_vec.h -- SpMV entry points of the arm-spmv API (reference include/mat_vec.h:7-23).
//
// Every function computes y += A*x and returns when y is complete (the reference's callers
// bracket these calls with mytimer(), main.cpp:56-59).  Underneath they launch the sm_100a
// kernels declared in thsp.h; there is no CPU code path.
#ifndef MAT_VEC_H
#define MAT_VEC_H

#include "matrix.h"
#include "vector.h"

// ---- one GPU --------------------------------------------------------------------------------
void CSRMatrixMatVector(const CSRMatrix& A, const Vector& x, Vector& y);   // TMA-stream / vector / merge kernel, chosen per matrix
void ELLMatrixMatVector(const ELLMatrix& A, const Vector& x, Vector& y);
void DIAMatrixMatVector(const DIAMatrix& A, const Vector& x, Vector& y);
void CSCMatrixMatVector(const CSCMatrix& A, const Vector& x, Vector& y);
void COOMatirxMatVector(const COOMatrix& A, const Vector& x, Vector& y);   // "Matirx": the reference's spelling, part of its link interface

// ---- partitioned --------------------------------------------------------------------------
// The reference splits rows (columns for CSC) into `nthreads` equal blocks placed on NUMA nodes,
// runs 50 repeats and prints "### <FMT> NUMA GFLOPS = ...".  Here the blocks are placed on
// min(nthreads, #GPUs) B200s with x replicated; the same line is printed.  Unlike the reference
// (which drops the result for four of the five formats) the accumulated blocks
// are written back into y.
void CSRMatrixMatVectorNuma(const CSRMatrix& A, const Vector& x, Vector& y, int nthreads);
void ELLMatrixMatVectorNuma(const ELLMatrix& A, const Vector& x, Vector& y, int nthreads);
void DIAMatrixMatVectorNuma(const DIAMatrix& A, const Vector& x, Vector& y, int nthreads);
void CSCMatrixMatVectorNuma(const CSCMatrix& A, const Vector& x, Vector& y, int nthreads);
void COOMatrixMatVectorNuma(const COOMatrix& A, const Vector& x, Vector& y, int nthreads);

// Bodies of the reference's pthreads; `block` points at a NumaNode4CSR / ...ELL / ...DIA / ...CSC /
// ...COO (numa_node.h) whose arrays are device pointers here.  Each runs one
// y_block += A_block * x on the block's GPU and returns NULL.
void* CSRMatrixMatVectorNumaThread(void* block);
void* ELLMatrixMatVectorNumaThread(void* block);
void* DIAMatrixMatVectorNumaThread(void* block);
void* CSCMatrixMatVectorNumaThread(void* block);
void* COOMatrixMatVectorNumaThread(void* block);

#endif  // MAT_VEC_H
