// mmio.h -- Matrix Market banner / size-line helpers with the NIST "mmio" C interface.
//
// The reference vendors NIST's public-domain mmio library (include/mmio.h, src/mmio.cpp) and
// calls mm_read_banner, mm_read_mtx_crd_size, the mm_is_* predicates and mm_typecode_to_str from
// COOMatrixRead (src/data_io.cpp:60-74).  This is an independent implementation of the same
// interface (same names, signatures, type-code layout and error codes), so code written
// against mmio.h keeps compiling and linking.  File parsing stays on the CPU: it is not part
// of the accelerated path.
#ifndef MMIO_H
#define MMIO_H

#include <stdio.h>

#define MM_MAX_LINE_LENGTH 1025
#define MatrixMarketBanner "%%MatrixMarket"
#define MM_MAX_TOKEN_LENGTH 64

// typecode[0] object   'M' matrix
// typecode[1] format   'C' coordinate (sparse) | 'A' array (dense)
// typecode[2] field    'R' real | 'C' complex | 'P' pattern | 'I' integer
// typecode[3] symmetry 'G' general | 'S' symmetric | 'K' skew-symmetric | 'H' hermitian
typedef char MM_typecode[4];

#define MM_COULD_NOT_READ_FILE 11
#define MM_PREMATURE_EOF 12
#define MM_NOT_MTX 13
#define MM_NO_HEADER 14
#define MM_UNSUPPORTED_TYPE 15
#define MM_LINE_TOO_LONG 16
#define MM_COULD_NOT_WRITE_FILE 17

// queries
#define mm_is_matrix(t) ((t)[0] == 'M')
#define mm_is_sparse(t) ((t)[1] == 'C')
#define mm_is_coordinate(t) ((t)[1] == 'C')
#define mm_is_dense(t) ((t)[1] == 'A')
#define mm_is_array(t) ((t)[1] == 'A')
#define mm_is_complex(t) ((t)[2] == 'C')
#define mm_is_real(t) ((t)[2] == 'R')
#define mm_is_pattern(t) ((t)[2] == 'P')
#define mm_is_integer(t) ((t)[2] == 'I')
#define mm_is_symmetric(t) ((t)[3] == 'S')
#define mm_is_general(t) ((t)[3] == 'G')
#define mm_is_skew(t) ((t)[3] == 'K')
#define mm_is_hermitian(t) ((t)[3] == 'H')

// setters take a pointer to the typecode, as in NIST's header
#define mm_set_matrix(t) ((*(t))[0] = 'M')
#define mm_set_coordinate(t) ((*(t))[1] = 'C')
#define mm_set_array(t) ((*(t))[1] = 'A')
#define mm_set_dense(t) mm_set_array(t)
#define mm_set_sparse(t) mm_set_coordinate(t)
#define mm_set_complex(t) ((*(t))[2] = 'C')
#define mm_set_real(t) ((*(t))[2] = 'R')
#define mm_set_pattern(t) ((*(t))[2] = 'P')
#define mm_set_integer(t) ((*(t))[2] = 'I')
#define mm_set_symmetric(t) ((*(t))[3] = 'S')
#define mm_set_general(t) ((*(t))[3] = 'G')
#define mm_set_skew(t) ((*(t))[3] = 'K')
#define mm_set_hermitian(t) ((*(t))[3] = 'H')
#define mm_clear_typecode(t) ((*(t))[0] = (*(t))[1] = (*(t))[2] = ' ', (*(t))[3] = 'G')
#define mm_initialize_typecode(t) mm_clear_typecode(t)

extern char MM_MTX_STR[20];
extern char MM_ARRAY_STR[20];
extern char MM_DENSE_STR[20];
extern char MM_COORDINATE_STR[20];
extern char MM_SPARSE_STR[20];
extern char MM_COMPLEX_STR[20];
extern char MM_REAL_STR[20];
extern char MM_INT_STR[20];
extern char MM_GENERAL_STR[20];
extern char MM_SYMM_STR[20];
extern char MM_HERM_STR[20];
extern char MM_SKEW_STR[20];
extern char MM_PATTERN_STR[20];

int   mm_is_valid(MM_typecode matcode);
char* mm_typecode_to_str(MM_typecode matcode);  // malloc'ed; caller frees

int mm_read_banner(FILE* f, MM_typecode* matcode);
int mm_read_mtx_crd_size(FILE* f, int* M, int* N, int* nz);
int mm_read_mtx_array_size(FILE* f, int* M, int* N);

int mm_write_banner(FILE* f, MM_typecode matcode);
int mm_write_mtx_crd_size(FILE* f, int M, int N, int nz);
int mm_write_mtx_array_size(FILE* f, int M, int N);

int mm_read_mtx_crd_entry(FILE* f, int* I, int* J, double* real, double* img, MM_typecode matcode);
int mm_read_mtx_crd_data(FILE* f, int M, int N, int nz, int I[], int J[], double val[], MM_typecode matcode);
int mm_write_mtx_crd(char fname[], int M, int N, int nz, int I[], int J[], double val[], MM_typecode matcode);
int mm_read_unsymmetric_sparse(const char* fname, int* M_, int* N_, int* nz_, double** val_, int** I_, int** J_);

#endif  // MMIO_H
