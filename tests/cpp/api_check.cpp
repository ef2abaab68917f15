// api_check.cpp -- exercises the PUBLIC C++ API of the arm-spmv headers and dumps every result.
//
// The same source is compiled twice: against this repo's include/ + bin/TH_sparse.a (GPU) and
// against the unmodified reference headers + objects (oracle/_ref/api_check_ref, CPU).  The
// parity test (tests/test_gpu_dropin.py) compares the two dumps.  Only calls that exist in the
// reference's headers are used.
//
//   api_check <matrix.mtx> <out-dir> [numa-threads]
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "data_io.h"
#include "mat_vec.h"
#include "matrix.h"
#include "mytime.h"
#include "vec_vec.h"
#include "vector.h"

static void dump(const std::string& path, const void* p, size_t bytes)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); exit(2); }
    if (bytes) fwrite(p, 1, bytes, f);
    fclose(f);
}
static void dump_vec(const std::string& dir, const char* name, const Vector& v) { dump(dir + "/" + name + ".f64", v.values, sizeof(double) * v.size); }

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: api_check <matrix.mtx> <out-dir> [numa-threads]\n"); return 1; }
    const std::string dir = argv[2];
    const int numa_threads = argc > 3 ? atoi(argv[3]) : 0;

    COOMatrix A;
    COOMatrixRead(argv[1], A);
    const int nrow = A.nrow, ncol = A.ncol;
    Vector x, y;
    x.Resize(ncol);
    y.Resize(nrow);
    x.FillRandom();
    dump_vec(dir, "x", x);

    // --- conversions
    CSRMatrix B(A);
    CSCMatrix C(A);
    ELLMatrix D(A);
    DIAMatrix E(B);
    dump(dir + "/csr_row_ptr.i32", B.row_ptr, sizeof(int) * (nrow + 1));
    dump(dir + "/csr_col_ind.i32", B.col_ind, sizeof(int) * A.nnz);
    dump(dir + "/csr_values.f64", B.values, sizeof(double) * A.nnz);
    dump(dir + "/csc_col_ptr.i32", C.col_ptr, sizeof(int) * (ncol + 1));
    dump(dir + "/csc_row_ind.i32", C.row_ind, sizeof(int) * A.nnz);
    dump(dir + "/csc_values.f64", C.values, sizeof(double) * A.nnz);
    dump(dir + "/ell_col_ind.i32", D.col_ind, sizeof(int) * (size_t)nrow * D.nonzeros_in_row);
    dump(dir + "/ell_values.f64", D.values, sizeof(double) * (size_t)nrow * D.nonzeros_in_row);
    dump(dir + "/dia_offsets.i32", E.offsets, sizeof(int) * E.ndiags);
    dump(dir + "/dia_values.f64", E.values, sizeof(double) * (size_t)nrow * E.ndiags);
    int meta[6] = {nrow, ncol, A.nnz, D.nonzeros_in_row, E.ndiags, E.nnz};
    dump(dir + "/meta.i32", meta, sizeof(meta));

    // --- y += A x, three times without re-zeroing (main.cpp's usage pattern)
    y.Fill(0); for (int k = 0; k < 3; ++k) COOMatirxMatVector(A, x, y); dump_vec(dir, "y_coo", y);
    y.Fill(0); for (int k = 0; k < 3; ++k) CSRMatrixMatVector(B, x, y); dump_vec(dir, "y_csr", y);
    y.Fill(0); for (int k = 0; k < 3; ++k) CSCMatrixMatVector(C, x, y); dump_vec(dir, "y_csc", y);
    y.Fill(0); for (int k = 0; k < 3; ++k) ELLMatrixMatVector(D, x, y); dump_vec(dir, "y_ell", y);
    y.Fill(0); for (int k = 0; k < 3; ++k) DIAMatrixMatVector(E, x, y); dump_vec(dir, "y_dia", y);

    // --- copies and assignment keep contents
    CSRMatrix B2(B);
    CSRMatrix B3;
    B3 = A;
    y.Fill(0); CSRMatrixMatVector(B2, x, y); dump_vec(dir, "y_csr_copy", y);
    y.Fill(0); CSRMatrixMatVector(B3, x, y); dump_vec(dir, "y_csr_assign", y);

    // --- vector kernels
    Vector w, z(y);
    w.Resize(nrow);
    double scal[4];
    scal[0] = vec_dot(y, y);
    vec_axpby(0.5, y, -2.0, z, w);   dump_vec(dir, "w_axpby", w);
    vec_axpby(1.0 / sqrt(scal[0]), y, 0.0, y, w); dump_vec(dir, "w_normalised", w);   // power-iteration step
    z.Scale(1.5); z.Shift(-0.25); z.AddScaled(0.3, y); z.Add2Scaled(0.1, y, -1.0, w); dump_vec(dir, "z_chain", z);
    scal[1] = vec_dot(w, w);
    scal[2] = checkVector(y, y) ? 1.0 : 0.0;
    scal[3] = checkVector(y, z) ? 1.0 : 0.0;
    dump(dir + "/scalars.f64", scal, sizeof(scal));

    // --- partitioned variants (results are only written back by this repo's build; see mat_vec.h)
    if (numa_threads > 0) {
        y.Fill(0); CSRMatrixMatVectorNuma(B, x, y, numa_threads); dump_vec(dir, "y_csr_numa", y);
        y.Fill(0); ELLMatrixMatVectorNuma(D, x, y, numa_threads); dump_vec(dir, "y_ell_numa", y);
        y.Fill(0); COOMatrixMatVectorNuma(A, x, y, numa_threads); dump_vec(dir, "y_coo_numa", y);
        y.Fill(0); CSCMatrixMatVectorNuma(C, x, y, numa_threads); dump_vec(dir, "y_csc_numa", y);
        y.Fill(0); DIAMatrixMatVectorNuma(E, x, y, numa_threads); dump_vec(dir, "y_dia_numa", y);
    }
    printf("api_check done: %d x %d, nnz %d, K %d, ndiags %d, t=%.3f s\n", nrow, ncol, A.nnz, D.nonzeros_in_row, E.ndiags, mytimer());
    return 0;
}
