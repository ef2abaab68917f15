// read_time.cpp -- wall time of COOMatrixRead (include/data_io.h) on one file, linked against bin/TH_sparse.a
// the way the reference's driver is.  Usage: read_time <file.mtx> [repeats]
#include <stdio.h>
#include <stdlib.h>

#include "data_io.h"
#include "matrix.h"
#include "mytime.h"

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    const int reps = argc > 2 ? atoi(argv[2]) : 3;
    mytimer();
    for (int r = 0; r < reps; ++r) {
        COOMatrix A;
        const double t0 = mytimer();
        COOMatrixRead(argv[1], A);
        const double t1 = mytimer();
        printf("### COOMatrixRead %d: %.1f ms, nnz %d, last entry (%d, %d, %.17g)\n", r, 1e3 * (t1 - t0), A.nnz, A.row_ind[A.nnz - 1],
               A.col_ind[A.nnz - 1], A.values[A.nnz - 1]);
    }
    return 0;
}
