// numa_node.h -- per-block descriptors of the partitioned SpMV (reference include/numa_node.h).
//
// Field-compatible with the reference.  In the reference `alloc` is a NUMA node and the arrays
// sit in numa_alloc_onnode memory; here `alloc` is a CUDA device ordinal and the arrays are
// cudaMalloc'ed on that device.  Row pointers of a block are rebased to start at 0
// (src/mat_vec.cpp:260-263), which is also what keeps them int32 when the whole matrix has
// more than 2^31 entries.
#ifndef NUMA_NODE_H
#define NUMA_NODE_H

class NumaNode4COO {
public:
    int     alloc;          // device holding this block
    int     core_ind;       // block index
    int     nnz;
    int     start_row;
    int     rows_per_node;
    int*    sub_row_ind;
    int*    sub_col_ind;
    double* sub_values;
    double* X;
    double* Y;
};

class NumaNode4CSR {
public:
    int     alloc;
    int     nnz;
    int     core_ind;
    int     start_row;
    int     rows_per_node;
    int*    sub_row_ptr;    // rebased: sub_row_ptr[0] == 0
    int*    sub_col_ind;
    double* sub_values;
    double* X;              // full-length replica of x
    double* Y;              // this block's rows of y
};

class NumaNode4CSC {
public:
    int     alloc;
    int     nnz;
    int     core_ind;
    int     start_col;
    int     cols_per_node;
    int*    sub_col_ptr;    // rebased
    int*    sub_row_ind;
    double* sub_values;
    double* X;              // this block's columns of x
    double* Y;              // full-length private y
};

class NumaNode4ELL {
public:
    int     alloc;
    int     core_ind;
    int     rows_per_node;
    int     nonzeros_in_row;
    int*    sub_col_ind;    // column-major slab of the block: [i + k*rows_per_node]
    double* sub_values;
    double* X;
    double* Y;
};

struct NumaNode4DIA {
public:
    int     alloc;
    int     core_ind;
    int     start_row;
    int     rows_per_node;
    int     ndiags;
    int*    offsets;
    double* values;
    double* X;
    double* Y;
};

#endif  // NUMA_NODE_H
