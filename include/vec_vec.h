// vec_vec.h -- vector-vector kernels of the arm-spmv API (reference include/vec_vec.h:6-7).
#ifndef VEC_VEC_H
#define VEC_VEC_H

#include "vector.h"

// sum_i x_i*y_i over x.size entries; deterministic two-level tree on the GPU (thsp_dot_f64).
double vec_dot(const Vector& x, const Vector& y);
// w = alpha*x + beta*y over w.size entries, with the reference's seven-way dispatch on
// alpha/beta in {0, 1, -1} (src/vec_vec.cpp:38-93) so results are bit-identical (thsp_axpby_f64).
void   vec_axpby(double alpha, const Vector& x, double beta, const Vector& y, const Vector& w);

#endif  // VEC_VEC_H
