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
    int size;            // number of entries
    double* values;      // [size]

    ~Vector();
    Vector();
    Vector(const Vector& other);              // deep copy
    Vector(int length, double* storage);      // adopts `storage` (reference: src/vector.cpp:12)

    Vector& operator=(const Vector& other);   // resize + copy
    Vector& operator=(double a);              // same as Fill(a)

    void Resize(int length);                  // contents are not preserved (src/vector.cpp:51-57)
    void Free();

    // elementwise, on the GPU; `const` as in the reference although they write through `values`
    void Fill(double a) const;                                                          // v = a
    void Copy(const Vector& x) const;                                                   // v = x
    void Scale(double a) const;                                                         // v *= a
    void Shift(double a) const;                                                         // v += a
    void AddScaled(double a, const Vector& x) const;                                    // v += a x
    void Add2Scaled(double a, const Vector& x, double b, const Vector& y) const;        // v += a x + b y
    void FillRandom() const;                  // host glibc rand()/RAND_MAX sequence, as in the reference
};

// true iff sizes match and every |x_i - y_i| <= 1e-6 (src/vector.cpp:161-171)
bool checkVector(const Vector& x, const Vector& y);

#endif  // VECTOR_H
