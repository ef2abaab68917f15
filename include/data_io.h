// data_io.h -- file readers/writers of the arm-spmv API (reference include/data_io.h:9-15).
// Text parsing stays on the CPU; the arrays land in CUDA managed memory ready for the GPU.
#ifndef DATA_IO_H
#define DATA_IO_H

#include <stdio.h>

#include "matrix.h"
#include "vector.h"

// "<n>" then one "%20.16g" value per line (src/data_io.cpp:10-40)
void VectorRead(const char* filename, Vector& x);
void VectorWrite(const char* filename, const Vector& x);

// Matrix Market coordinate file -> COO (1-based -> 0-based); CSR/CSC/ELL = read COO + convert.
void COOMatrixRead(const char* filename, COOMatrix& A);
void CSRMatrixRead(const char* filename, CSRMatrix& A);
void CSCMatrixRead(const char* filename, CSCMatrix& A);
void ELLMatrixRead(const char* filename, ELLMatrix& A);

#endif  // DATA_IO_H
