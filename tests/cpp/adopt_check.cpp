// adopt_check.cpp -- arrays ADOPTED by the pointer-taking constructors (src/matrix.cpp:12-15,88-91; src/vector.cpp:12)
// through the C++ drop-in API: plain new[] memory handed to CSRMatrix / Vector, products repeated, arrays refilled and
// row_ptr rewritten in place between calls.  Every product is compared with the loop of src/mat_vec.cpp:44-67 run right
// here on the host; the device mirrors are counted through THSP_TRACE=1 by the test that runs this (tests/test_gpu_dropin.py).
//   adopt_check <n>      tridiagonal-plus matrix with n rows
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mat_vec.h"
#include "matrix.h"
#include "vec_vec.h"
#include "vector.h"

static int check(const char* what, int n, const int* rp, const int* ci, const double* va, const double* x, const double* y0, const double* y)
{
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int p = rp[i]; p < rp[i + 1]; ++p) {
            double t = va[p] * x[ci[p]];
            s = s + t;
        }
        double want = y0[i] + s;
        if (memcmp(&want, &y[i], sizeof(double)) != 0 && !(fabs(want - y[i]) <= 1e-12 * (fabs(want) + 1.0))) ++bad;
    }
    printf("%-44s %s\n", what, bad ? "MISMATCH" : "ok");
    return bad;
}

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 5000;
    const int cap = 5 * n;
    int* rp = new int[n + 1];
    int* ci = new int[cap];
    double* va = new double[cap];
    int nnz = 0;
    for (int i = 0; i < n; ++i) {
        rp[i] = nnz;
        for (int d = -2; d <= 2; ++d)
            if (i + d >= 0 && i + d < n && (d != 2 || i % 3 == 0)) {
                ci[nnz] = i + d;
                va[nnz] = 1.0 / (1 + abs(d)) + 1e-3 * i;
                ++nnz;
            }
    }
    rp[n] = nnz;
    double* xs = new double[n];
    double* ys = new double[n];
    double* y0 = new double[n];
    for (int i = 0; i < n; ++i) { xs[i] = 0.5 + (i % 7) * 0.125; ys[i] = 0.25 * (i % 5); y0[i] = ys[i]; }
    int bad = 0;
    {
        CSRMatrix A(n, n, rp, ci, va, nullptr);   // adopts rp / ci / va
        Vector x(n, xs), y(n, ys);                // adopt xs / ys
        fprintf(stderr, "[adopt] call 1\n");
        CSRMatrixMatVector(A, x, y);
        bad += check("first product (mirrors uploaded)", n, rp, ci, va, xs, y0, y.values);
        memcpy(y0, y.values, sizeof(double) * n);
        fprintf(stderr, "[adopt] call 2\n");
        CSRMatrixMatVector(A, x, y);
        bad += check("second product (mirrors found again)", n, rp, ci, va, xs, y0, y.values);
        // x changes between calls: staged every time
        for (int i = 0; i < n; ++i) xs[i] = 1.0 - 1e-4 * i;
        memcpy(y0, y.values, sizeof(double) * n);
        fprintf(stderr, "[adopt] call 3\n");
        CSRMatrixMatVector(A, x, y);
        bad += check("x rewritten by the caller", n, rp, ci, va, xs, y0, y.values);
        // the caller refills the values in place: the sampled fingerprint sees it
        for (int p = 0; p < nnz; ++p) va[p] = -va[p] + 0.01 * (p % 11);
        memcpy(y0, y.values, sizeof(double) * n);
        fprintf(stderr, "[adopt] call 4\n");
        CSRMatrixMatVector(A, x, y);
        bad += check("values refilled in place", n, rp, ci, va, xs, y0, y.values);
        // ... and rewrites the structure in place: last third of the rows emptied, row_ptr[n] shrinks
        const int cut = 2 * n / 3;
        for (int i = cut + 1; i <= n; ++i) rp[i] = rp[cut];
        memcpy(y0, y.values, sizeof(double) * n);
        fprintf(stderr, "[adopt] call 5\n");
        CSRMatrixMatVector(A, x, y);
        bad += check("row_ptr rewritten in place (fewer entries)", n, rp, ci, va, xs, y0, y.values);
        // the other formats through the same adopted arrays: COO
        int* ri = new int[nnz];
        int* cj = new int[nnz];
        double* vv = new double[nnz];
        const int m = rp[n];
        for (int i = 0; i < n; ++i)
            for (int p = rp[i]; p < rp[i + 1]; ++p) { ri[p] = i; cj[p] = ci[p]; vv[p] = va[p]; }
        COOMatrix C(n, n, m, ri, cj, vv);
        memcpy(y0, y.values, sizeof(double) * n);
        COOMatirxMatVector(C, x, y);
        bad += check("COO with adopted arrays", n, rp, ci, va, xs, y0, y.values);
        CSRMatrix B(C);   // conversion from an adopted COO (mirror shared with the product above)
        memcpy(y0, y.values, sizeof(double) * n);
        CSRMatrixMatVector(B, x, y);
        bad += check("CSR converted from the adopted COO", n, rp, ci, va, xs, y0, y.values);
        // destructors release: adopted arrays with delete[], library arrays with cudaFree
    }
    {
        // A LIBRARY-owned matrix (managed memory, host-writable through the public pointers) with rows long enough for
        // the TMA stream kernel, whose bulk copies are bounded by the plan's entry count: the host rewrites row_ptr in
        // place after the plan exists; the kernel must notice (thsp_csr_plan_stale), write nothing, and the product must
        // be repeated with a fresh plan.
        const int w = 11, m = n;
        int cnt = 0;
        for (int i = 0; i < m; ++i)
            for (int d = -5; d <= 5; ++d) cnt += (i + d >= 0 && i + d < m);
        int* ri = new int[cnt];
        int* cj = new int[cnt];
        double* vv = new double[cnt];
        int k = 0;
        for (int i = 0; i < m; ++i)
            for (int d = -5; d <= 5; ++d)
                if (i + d >= 0 && i + d < m) { ri[k] = i; cj[k] = i + d; vv[k] = 1.0 + 0.001 * ((i + 3 * d) % 17); ++k; }
        (void)w;
        COOMatrix C(m, m, cnt, ri, cj, vv);
        CSRMatrix B(C);
        Vector x, y;
        x.Resize(m); y.Resize(m);
        for (int i = 0; i < m; ++i) { x.values[i] = 0.5 + 0.001 * (i % 13); y.values[i] = 0.0; }
        double* yb = new double[m];
        memset(yb, 0, sizeof(double) * m);
        CSRMatrixMatVector(B, x, y);
        bad += check("library-owned CSR, stream-sized rows", m, B.row_ptr, B.col_ind, B.values, x.values, yb, y.values);
        const int cut = m / 2;
        for (int i = cut + 1; i <= m; ++i) B.row_ptr[i] = B.row_ptr[cut];   // host write into managed memory
        memcpy(yb, y.values, sizeof(double) * m);
        fprintf(stderr, "[adopt] managed rewrite\n");
        CSRMatrixMatVector(B, x, y);
        bad += check("row_ptr of a library-owned CSR rewritten", m, B.row_ptr, B.col_ind, B.values, x.values, yb, y.values);
        delete[] yb;
    }
    delete[] y0;
    printf("%s\n", bad ? "FAILED" : "ALL OK");
    return bad ? 1 : 0;
}
