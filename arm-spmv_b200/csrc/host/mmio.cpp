// mmio.cpp -- independent implementation of the NIST mmio interface declared in mmio.h.
#include "mmio.h"

#include <ctype.h>
#include <stdlib.h>
#include <string.h>

char MM_MTX_STR[20] = "matrix";
char MM_ARRAY_STR[20] = "array";
char MM_DENSE_STR[20] = "array";
char MM_COORDINATE_STR[20] = "coordinate";
char MM_SPARSE_STR[20] = "coordinate";
char MM_COMPLEX_STR[20] = "complex";
char MM_REAL_STR[20] = "real";
char MM_INT_STR[20] = "integer";
char MM_GENERAL_STR[20] = "general";
char MM_SYMM_STR[20] = "symmetric";
char MM_HERM_STR[20] = "hermitian";
char MM_SKEW_STR[20] = "skew-symmetric";
char MM_PATTERN_STR[20] = "pattern";

namespace {

void lower(char* s)
{
    for (; *s; ++s) *s = (char)tolower((unsigned char)*s);
}

// Next line that is neither a '%' comment nor blank.  Returns 0, or an MM_ error.
int next_data_line(FILE* f, char* line)
{
    for (;;) {
        if (!fgets(line, MM_MAX_LINE_LENGTH, f)) return MM_PREMATURE_EOF;
        const char* p = line;
        while (*p == ' ' || *p == '\t') ++p;
        if (*p == '%' || *p == '\n' || *p == '\r' || *p == '\0') continue;
        return 0;
    }
}

const char* field_name(const MM_typecode t)
{
    if (mm_is_real(t)) return MM_REAL_STR;
    if (mm_is_complex(t)) return MM_COMPLEX_STR;
    if (mm_is_pattern(t)) return MM_PATTERN_STR;
    if (mm_is_integer(t)) return MM_INT_STR;
    return nullptr;
}
const char* symmetry_name(const MM_typecode t)
{
    if (mm_is_general(t)) return MM_GENERAL_STR;
    if (mm_is_symmetric(t)) return MM_SYMM_STR;
    if (mm_is_hermitian(t)) return MM_HERM_STR;
    if (mm_is_skew(t)) return MM_SKEW_STR;
    return nullptr;
}

}  // namespace

int mm_is_valid(MM_typecode t)
{
    if (!mm_is_matrix(t)) return 0;
    if (mm_is_dense(t) && mm_is_pattern(t)) return 0;
    if (mm_is_real(t) && mm_is_hermitian(t)) return 0;
    if (mm_is_pattern(t) && (mm_is_hermitian(t) || mm_is_skew(t))) return 0;
    return 1;
}

char* mm_typecode_to_str(MM_typecode t)
{
    const char* obj = mm_is_matrix(t) ? MM_MTX_STR : nullptr;
    const char* fmt = mm_is_sparse(t) ? MM_SPARSE_STR : (mm_is_dense(t) ? MM_DENSE_STR : nullptr);
    const char* fld = field_name(t);
    const char* sym = symmetry_name(t);
    if (!obj || !fmt || !fld || !sym) return nullptr;
    char buf[MM_MAX_LINE_LENGTH];
    snprintf(buf, sizeof(buf), "%s %s %s %s", obj, fmt, fld, sym);
    char* out = (char*)malloc(strlen(buf) + 1);
    if (out) strcpy(out, buf);
    return out;
}

int mm_read_banner(FILE* f, MM_typecode* matcode)
{
    char line[MM_MAX_LINE_LENGTH];
    char banner[MM_MAX_TOKEN_LENGTH], obj[MM_MAX_TOKEN_LENGTH], fmt[MM_MAX_TOKEN_LENGTH], fld[MM_MAX_TOKEN_LENGTH],
        sym[MM_MAX_TOKEN_LENGTH];
    mm_clear_typecode(matcode);
    if (!fgets(line, MM_MAX_LINE_LENGTH, f)) return MM_PREMATURE_EOF;
    if (sscanf(line, "%63s %63s %63s %63s %63s", banner, obj, fmt, fld, sym) != 5) return MM_PREMATURE_EOF;
    lower(obj); lower(fmt); lower(fld); lower(sym);
    if (strncmp(banner, MatrixMarketBanner, strlen(MatrixMarketBanner)) != 0) return MM_NO_HEADER;
    if (strcmp(obj, MM_MTX_STR) != 0) return MM_UNSUPPORTED_TYPE;
    mm_set_matrix(matcode);
    if (strcmp(fmt, MM_SPARSE_STR) == 0) mm_set_sparse(matcode);
    else if (strcmp(fmt, MM_DENSE_STR) == 0) mm_set_dense(matcode);
    else return MM_UNSUPPORTED_TYPE;
    if (strcmp(fld, MM_REAL_STR) == 0) mm_set_real(matcode);
    else if (strcmp(fld, MM_COMPLEX_STR) == 0) mm_set_complex(matcode);
    else if (strcmp(fld, MM_PATTERN_STR) == 0) mm_set_pattern(matcode);
    else if (strcmp(fld, MM_INT_STR) == 0) mm_set_integer(matcode);
    else return MM_UNSUPPORTED_TYPE;
    if (strcmp(sym, MM_GENERAL_STR) == 0) mm_set_general(matcode);
    else if (strcmp(sym, MM_SYMM_STR) == 0) mm_set_symmetric(matcode);
    else if (strcmp(sym, MM_HERM_STR) == 0) mm_set_hermitian(matcode);
    else if (strcmp(sym, MM_SKEW_STR) == 0) mm_set_skew(matcode);
    else return MM_UNSUPPORTED_TYPE;
    return 0;
}

int mm_read_mtx_crd_size(FILE* f, int* M, int* N, int* nz)
{
    char line[MM_MAX_LINE_LENGTH];
    *M = *N = *nz = 0;
    for (;;) {
        int rc = next_data_line(f, line);
        if (rc) return rc;
        if (sscanf(line, "%d %d %d", M, N, nz) == 3) return 0;
    }
}

int mm_read_mtx_array_size(FILE* f, int* M, int* N)
{
    char line[MM_MAX_LINE_LENGTH];
    *M = *N = 0;
    for (;;) {
        int rc = next_data_line(f, line);
        if (rc) return rc;
        if (sscanf(line, "%d %d", M, N) == 2) return 0;
    }
}

int mm_write_banner(FILE* f, MM_typecode matcode)
{
    char* s = mm_typecode_to_str(matcode);
    if (!s) return MM_COULD_NOT_WRITE_FILE;
    int n = fprintf(f, "%s %s\n", MatrixMarketBanner, s);
    free(s);
    return n < 0 ? MM_COULD_NOT_WRITE_FILE : 0;
}

int mm_write_mtx_crd_size(FILE* f, int M, int N, int nz) { return fprintf(f, "%d %d %d\n", M, N, nz) < 0 ? MM_COULD_NOT_WRITE_FILE : 0; }

int mm_write_mtx_array_size(FILE* f, int M, int N) { return fprintf(f, "%d %d\n", M, N) < 0 ? MM_COULD_NOT_WRITE_FILE : 0; }

int mm_read_mtx_crd_entry(FILE* f, int* I, int* J, double* real, double* img, MM_typecode t)
{
    if (mm_is_complex(t)) return fscanf(f, "%d %d %lg %lg", I, J, real, img) == 4 ? 0 : MM_PREMATURE_EOF;
    if (mm_is_real(t) || mm_is_integer(t)) return fscanf(f, "%d %d %lg\n", I, J, real) == 3 ? 0 : MM_PREMATURE_EOF;
    if (mm_is_pattern(t)) return fscanf(f, "%d %d", I, J) == 2 ? 0 : MM_PREMATURE_EOF;
    return MM_UNSUPPORTED_TYPE;
}

int mm_read_mtx_crd_data(FILE* f, int M, int N, int nz, int I[], int J[], double val[], MM_typecode t)
{
    (void)M; (void)N;
    for (int k = 0; k < nz; ++k) {
        double im = 0.0;
        double* re = mm_is_pattern(t) ? nullptr : (mm_is_complex(t) ? &val[2 * k] : &val[k]);
        double dummy = 0.0;
        int rc = mm_read_mtx_crd_entry(f, &I[k], &J[k], re ? re : &dummy, mm_is_complex(t) ? &val[2 * k + 1] : &im, t);
        if (rc) return rc;
    }
    return 0;
}

int mm_write_mtx_crd(char fname[], int M, int N, int nz, int I[], int J[], double val[], MM_typecode t)
{
    FILE* f = strcmp(fname, "stdout") == 0 ? stdout : fopen(fname, "w");
    if (!f) return MM_COULD_NOT_WRITE_FILE;
    int rc = mm_write_banner(f, t);
    if (!rc) rc = mm_write_mtx_crd_size(f, M, N, nz);
    for (int k = 0; k < nz && !rc; ++k) {
        if (mm_is_pattern(t)) fprintf(f, "%d %d\n", I[k], J[k]);
        else if (mm_is_complex(t)) fprintf(f, "%d %d %20.16g %20.16g\n", I[k], J[k], val[2 * k], val[2 * k + 1]);
        else if (mm_is_real(t) || mm_is_integer(t)) fprintf(f, "%d %d %20.16g\n", I[k], J[k], val[k]);
        else rc = MM_UNSUPPORTED_TYPE;
    }
    if (f != stdout) fclose(f);
    return rc;
}

int mm_read_unsymmetric_sparse(const char* fname, int* M_, int* N_, int* nz_, double** val_, int** I_, int** J_)
{
    FILE* f = fopen(fname, "r");
    if (!f) return -1;
    MM_typecode t;
    if (mm_read_banner(f, &t) != 0 || !(mm_is_real(t) && mm_is_matrix(t) && mm_is_sparse(t))) {
        fclose(f);
        return -1;
    }
    int M, N, nz;
    if (mm_read_mtx_crd_size(f, &M, &N, &nz) != 0) {
        fclose(f);
        return -1;
    }
    int* I = (int*)malloc(sizeof(int) * (size_t)(nz > 0 ? nz : 1));
    int* J = (int*)malloc(sizeof(int) * (size_t)(nz > 0 ? nz : 1));
    double* v = (double*)malloc(sizeof(double) * (size_t)(nz > 0 ? nz : 1));
    for (int k = 0; k < nz; ++k) {
        if (fscanf(f, "%d %d %lg\n", &I[k], &J[k], &v[k]) != 3) break;
        --I[k];
        --J[k];
    }
    fclose(f);
    *M_ = M; *N_ = N; *nz_ = nz; *val_ = v; *I_ = I; *J_ = J;
    return 0;
}

// Two more entry points the reference's archive exports (nm bin/TH_sparse.a) although its header
// does not declare them; kept so that anything linking against them still links.
char* mm_strdup(const char* s)
{
    char* d = (char*)malloc(strlen(s) + 1);
    return d ? strcpy(d, s) : nullptr;
}

int mm_read_mtx_crd(char* fname, int* M, int* N, int* nz, int** I, int** J, double** val, MM_typecode* matcode)
{
    FILE* f = strcmp(fname, "stdin") == 0 ? stdin : fopen(fname, "r");
    if (!f) return MM_COULD_NOT_READ_FILE;
    int rc = mm_read_banner(f, matcode);
    if (!rc && !(mm_is_valid(*matcode) && mm_is_sparse(*matcode) && mm_is_matrix(*matcode))) rc = MM_UNSUPPORTED_TYPE;
    if (!rc) rc = mm_read_mtx_crd_size(f, M, N, nz);
    if (!rc) {
        const size_t n = (size_t)(*nz > 0 ? *nz : 1);
        *I = (int*)malloc(n * sizeof(int));
        *J = (int*)malloc(n * sizeof(int));
        *val = mm_is_pattern(*matcode) ? nullptr : (double*)malloc(n * sizeof(double) * (mm_is_complex(*matcode) ? 2 : 1));
        rc = mm_read_mtx_crd_data(f, *M, *N, *nz, *I, *J, *val, *matcode);
    }
    if (f != stdin) fclose(f);
    return rc;
}
