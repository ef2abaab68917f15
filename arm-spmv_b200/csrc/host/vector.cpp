// vector.cpp -- Vector of the arm-spmv API on top of the C ABI (replaces src/vector.cpp).
#include "vector.h"

#include <math.h>
#include <stdlib.h>

#include "hostmem.h"

using namespace thsp_host;

Vector::Vector() : size(0), values(nullptr) {}

Vector::Vector(int n, double* v) : size(n), values(v) {}

Vector::Vector(const Vector& x) : size(x.size), values(alloc<double>(x.size)) { copy(values, x.values, (size_t)size); }

Vector::~Vector() { release(values); }

Vector& Vector::operator=(double a)
{
    Fill(a);
    return *this;
}

Vector& Vector::operator=(const Vector& x)
{
    if (this == &x) return *this;
    Resize(x.size);
    copy(values, x.values, (size_t)size);
    return *this;
}

void Vector::Free()
{
    release(values);
    size = 0;
}

void Vector::Resize(int n)
{
    release(values);
    size = n;
    values = alloc<double>(n);
}

void Vector::Fill(double a) const
{
    View<double> v(values, size, true, false);  // overwritten: no need to migrate old contents
    ok(thsp_fill_f64(size, a, v, nullptr), "Vector::Fill");
    v.commit();
    sync();
}

void Vector::FillRandom() const
{
    Trace tr("Vector::FillRandom");
    // Deliberately on the host: the reference's values are the process-wide glibc rand()
    // stream (src/vector.cpp:65-69); drawing it here keeps x identical to the reference's.
    for (int i = 0; i < size; ++i) values[i] = (double)rand() / RAND_MAX;
}

void Vector::Copy(const Vector& x) const { copy(values, x.values, (size_t)size); }

void Vector::Scale(double a) const
{
    View<double> v(values, size, true);
    ok(thsp_scale_f64(size, a, v, nullptr), "Vector::Scale");
    v.commit();
    sync();
}

void Vector::Shift(double a) const
{
    View<double> v(values, size, true);
    ok(thsp_shift_f64(size, a, v, nullptr), "Vector::Shift");
    v.commit();
    sync();
}

void Vector::AddScaled(double a, const Vector& x) const
{
    View<double> v(values, size, true);
    View<double> xv(x.values, size, false);
    ok(thsp_add_scaled_f64(size, a, xv, v, nullptr), "Vector::AddScaled");
    v.commit();
    sync();
}

void Vector::Add2Scaled(double a, const Vector& x, double b, const Vector& y) const
{
    View<double> v(values, size, true);
    View<double> xv(x.values, size, false);
    View<double> yv(y.values, size, false);
    ok(thsp_add2_scaled_f64(size, a, xv, b, yv, v, nullptr), "Vector::Add2Scaled");
    v.commit();
    sync();
}

bool checkVector(const Vector& x, const Vector& y)
{
    if (x.size != y.size) return false;
    View<double> xv(x.values, x.size, false);
    View<double> yv(y.values, y.size, false);
    int same = 0;
    ok(thsp_check_vector_f64(x.size, xv, y.size, yv, &same, nullptr), "checkVector");
    return same != 0;
}
