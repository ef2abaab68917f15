// vector.h -- dense fp64 vector of the arm-spmv API, backed by the B200 library.
//
// Source-compatible with the reference's include/vector.h:4-26: same public fields, same
// method signatures (including the `const` on mutating methods), so main.cpp compiles
// unchanged.  Differences are behind the interface:
//   * storage the library allocates (ctor copies, Resize, operator=) is CUDA managed memory,
//     so `values[i]` still works on the host (main.cpp:48-51) while kernels read it in HBM;
//   * a pointer handed to Vector(int, double*) is adopted as in the reference and released
//     with delete[]; managed storage is released with cudaFree.  The destructor tells them
//     apart by asking the CUDA runtime what kind of pointer it holds;
//   * Fill/Scale/Shift/Copy/AddScaled/Add2Scaled run as sm_100a kernels (thsp_*_f64 in thsp.h)
//     and return after the stream has drained, like the synchronous originals.
#ifndef VECTOR_H
#define VECTOR_H

class Vector {
public:
    int     size;
    double* values;

    Vector();
    Vector(int n, double* values);   // adopts `values` (reference: src/vector.cpp:12)
    Vector(const Vector& x);         // deep copy
    ~Vector();

    Vector& operator=(double a);
    Vector& operator=(const Vector& x);

    void Free();
    void Resize(int n);              // contents are not preserved (src/vector.cpp:51-57)
    void Fill(double a) const;
    void FillRandom() const;         // host glibc rand()/RAND_MAX sequence, as in the reference
    void Copy(const Vector& x) const;
    void Scale(double a) const;
    void Shift(double a) const;
    void AddScaled(double a, const Vector& x) const;
    void Add2Scaled(double a, const Vector& x, double b, const Vector& y) const;
};

// true iff sizes match and every |x_i - y_i| <= 1e-6 (src/vector.cpp:161-171)
bool checkVector(const Vector& x, const Vector& y);

#endif  // VECTOR_H
