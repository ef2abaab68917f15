// vec_vec.cpp -- vec_dot / vec_axpby of the arm-spmv API (replaces src/vec_vec.cpp).
#include "vec_vec.h"

#include "hostmem.h"

using namespace thsp_host;

double vec_dot(const Vector& x, const Vector& y)
{
    const int n = x.size;  // the reference takes the length from x (src/vec_vec.cpp:17)
    View<double> xv(x.values, n, false);
    View<double> yv(y.values, n, false);
    double r = 0.0;
    ok(thsp_dot_f64(n, xv, yv, &r, nullptr), "vec_dot");
    return r;
}

void vec_axpby(double alpha, const Vector& x, double beta, const Vector& y, const Vector& w)
{
    const int n = w.size;  // ... and from w here (src/vec_vec.cpp:33)
    View<double> xv(x.values, n, false);
    View<double> yv(y.values, n, false);
    View<double> wv(w.values, n, true, false);
    ok(thsp_axpby_f64(n, alpha, xv, beta, yv, wv, nullptr), "vec_axpby");
    wv.commit();
    sync();
}
