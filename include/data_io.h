// data_io.h -- file readers/writers of the arm-spmv API (reference include/data_io.h:9-15).
// Text parsing stays on the CPU; the arrays land in CUDA managed memory ready for the GPU.
#ifndef DATA_IO_H
#define DATA_IO_H

#include <stdio.h>

#include "matrix.h"
#include "vector.h"

// Matrix Market "coordinate real general" file -> COO (file indices are 1-based, ours 0-based).
// Prints the reference's progress lines and "### ROW=.., COL=.., NNZ=.."; exits on a bad file.
void COOMatrixRead(const char* path, COOMatrix& A);
// Read as COO, then convert on the GPU.
void CSRMatrixRead(const char* path, CSRMatrix& A);
void CSCMatrixRead(const char* path, CSCMatrix& A);
void ELLMatrixRead(const char* path, ELLMatrix& A);

// Plain text vectors: "<n>" then one "%20.16g" value per line (src/data_io.cpp:10-40).
void VectorWrite(const char* path, const Vector& x);
void VectorRead(const char* path, Vector& x);

#endif  // DATA_IO_H
