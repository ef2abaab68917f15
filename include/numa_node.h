// numa_node.h -- per-block descriptors of the partitioned SpMV (reference include/numa_node.h).
//
// Field-compatible with the reference (same names and types).  There `alloc` is a NUMA node and
// the arrays sit in numa_alloc_onnode memory; here `alloc` is a CUDA device ordinal and the
// arrays are cudaMalloc'ed on that device.  Row pointers of a block are rebased to start at 0
// (src/mat_vec.cpp:260-263), which is also what keeps them int32 when the whole matrix has more
// than 2^31 entries.
#ifndef NUMA_NODE_H
#define NUMA_NODE_H

// A run of COO entries.  Row indices stay global; the kernel subtracts start_row.
class NumaNode4COO {
public:
    int alloc, core_ind;                         // device holding the block, block index
    int nnz, start_row, rows_per_node;
    int *sub_row_ind, *sub_col_ind;              // [nnz] each, on device `alloc`
    double *sub_values, *X, *Y;                  // [nnz], replica of x, this block's y
};

// A block of consecutive rows.
class NumaNode4CSR {
public:
    int alloc, nnz, core_ind;
    int start_row, rows_per_node;
    int *sub_row_ptr, *sub_col_ind;              // [rows_per_node+1] with sub_row_ptr[0] == 0, [nnz]
    double *sub_values, *X, *Y;                  // [nnz], full-length replica of x, rows of y owned by the block
};

// A block of consecutive columns; Y is a full-length private result to be reduced.
class NumaNode4CSC {
public:
    int alloc, nnz, core_ind;
    int start_col, cols_per_node;
    int *sub_col_ptr, *sub_row_ind;              // rebased column pointers, global row indices
    double *sub_values, *X, *Y;                  // [nnz], the block's columns of x, full-length y
};

// A block of consecutive rows of the column-major slab: element [i + k*rows_per_node].
class NumaNode4ELL {
public:
    int alloc, core_ind;
    int rows_per_node, nonzeros_in_row;
    int* sub_col_ind;
    double *sub_values, *X, *Y;
};

// A block of consecutive rows of the row-major diagonals.
struct NumaNode4DIA {
public:
    int alloc, core_ind;
    int start_row, rows_per_node, ndiags;
    int* offsets;
    double *values, *X, *Y;
};

#endif  // NUMA_NODE_H
