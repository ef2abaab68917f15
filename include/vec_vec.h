// vec_vec.h -- the two vector-vector kernels of the arm-spmv API (reference include/vec_vec.h:6-7),
// executed on the GPU through thsp_dot_f64 / thsp_axpby_f64 (thsp.h).
#ifndef VEC_VEC_H
#define VEC_VEC_H

#include "vector.h"

// w = alpha*x + beta*y over w.size entries.  The reference dispatches on alpha, beta in {0, 1, -1}
// (seven branches, src/vec_vec.cpp:38-93); the same branches with unfused mul/add run on the GPU, so
// results are bit-identical.  `w` is const in the reference's signature although it is written.
void vec_axpby(double alpha, const Vector& x, double beta, const Vector& y, const Vector& w);

// sum_i x_i*y_i over x.size entries; a fixed two-level reduction tree, deterministic.
double vec_dot(const Vector& x, const Vector& y);

#endif  // VEC_VEC_H
