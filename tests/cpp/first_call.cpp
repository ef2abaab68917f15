// first_call.cpp -- times consecutive calls through the C++ API to separate one-time costs
// (managed-memory migration, lazy kernel loading) from the steady state.  Diagnostic only.
#include <stdio.h>
#include <stdlib.h>
#include "data_io.h"
#include "mat_vec.h"
#include "mytime.h"
int main(int argc, char** argv)
{
    if (argc < 2) return 1;
    COOMatrix A;
    COOMatrixRead(argv[1], A);
    Vector x, y;
    x.Resize(A.ncol); y.Resize(A.nrow); x.FillRandom();
    double s = 0; for (int i = 0; i < A.nnz; ++i) s += A.values[i] * x.values[A.col_ind[i]];   // host touches everything, like main.cpp
    mytimer();
    double t0 = mytimer(); y.Fill(0); double t1 = mytimer(); printf("first Fill       %.3f ms\n", (t1 - t0) * 1e3);
    for (int k = 0; k < 4; ++k) { t0 = mytimer(); COOMatirxMatVector(A, x, y); t1 = mytimer(); printf("COO call %d       %.3f ms\n", k, (t1 - t0) * 1e3); }
    t0 = mytimer(); CSRMatrix B(A); t1 = mytimer(); printf("COO->CSR         %.3f ms\n", (t1 - t0) * 1e3);
    t0 = mytimer(); CSRMatrix B2(A); t1 = mytimer(); printf("COO->CSR again   %.3f ms\n", (t1 - t0) * 1e3);
    for (int k = 0; k < 4; ++k) { t0 = mytimer(); CSRMatrixMatVector(B, x, y); t1 = mytimer(); printf("CSR call %d       %.3f ms\n", k, (t1 - t0) * 1e3); }
    t0 = mytimer(); ELLMatrix D(A); t1 = mytimer(); printf("COO->ELL         %.3f ms\n", (t1 - t0) * 1e3);
    for (int k = 0; k < 3; ++k) { t0 = mytimer(); ELLMatrixMatVector(D, x, y); t1 = mytimer(); printf("ELL call %d       %.3f ms\n", k, (t1 - t0) * 1e3); }
    printf("%g\n", s);
    return 0;
}
