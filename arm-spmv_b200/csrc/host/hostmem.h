// hostmem.h -- glue between the reference's raw-pointer classes and the C ABI (thsp.h).
//
// Ownership model (INTEGRATION.md "Ownership"):
//   * arrays the library allocates are CUDA managed memory (host-dereferenceable, GPU-resident after a prefetch).
//     Every such allocation is RECORDED here; release() frees by that record, never by asking the CUDA runtime what
//     a pointer looks like - a query that fails (sticky error, runtime already unloaded at exit) must not turn a
//     managed array into a delete[];
//   * arrays a caller hands to an adopting constructor (src/matrix.cpp:12-15,88-91; src/vector.cpp:12) are whatever
//     the caller allocated - normally new[] memory.  Matrix arrays of that kind get a DEVICE MIRROR that is uploaded
//     once and reused by later calls (dropped by Free() / the destructor / thsp_host::invalidate); vectors are staged
//     through pooled device buffers on every call, because their contents change between calls.
#pragma once
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>

#include "thsp.h"

namespace thsp_host {

[[noreturn]] void die(const char* what);
inline void ok(int rc, const char* what)
{
    if (rc) die(what);
}

void* alloc_managed_bytes(size_t bytes);   // recorded as library-owned
template <class T>
inline T* alloc(size_t n)
{
    return static_cast<T*>(alloc_managed_bytes((n ? n : 1) * sizeof(T)));
}

// Matrix arrays.  (Tried: cudaMemAdviseSetReadMostly, so that the host loop of main.cpp:46-52 would leave the copy in HBM
// valid.  Measured on B200 / driver 580: kernels WRITING such arrays - every conversion, the GPU parser - crawl, COO->CSC of
// 5.2 M entries 4.8 -> 883 ms, the whole driver run 2.3 -> 4.6 s.  Plain managed memory it is - except for the COO arrays of
// COOMatrixRead, which are marked AFTER the parser has written them: data_io.cpp.)
template <class T>
inline T* alloc_matrix(size_t n)
{
    return alloc<T>(n);
}

// 0 plain host, 1 device, 2 managed, 3 pinned host, -1 the CUDA runtime could not say.  Library-owned arrays are
// answered from the record; other pointers are asked once and remembered until they are released or invalidated.
int kind(const void* p);

// Managed arrays are prefetched to the GPU the first time a kernel is about to read them and
// not again: cudaMemPrefetchAsync on resident pages still walks the range (measured ~1 ms per
// 20 MB), which would dominate a 13 us SpMV.  If host code writes into such an array later, its
// pages migrate back on the CPU fault and return to the GPU on the next kernel's page faults.
bool first_gpu_use(const void* p);
void prefetch_traced(const void* p, size_t bytes);   // prefetch to the GPU; THSP_TRACE=1 prints how long it took
void forget_gpu_use(const void* p);

// Release an array held by one of the API classes: library-owned -> cudaFree by the record; adopted -> by what the
// runtime says it is (delete[] for ordinary host memory); if the runtime cannot say, the array is leaked, not freed.
void release_bytes(void* p);
template <class T>
inline void release(T*& p)
{
    if (!p) return;
    release_bytes(p);
    p = nullptr;
}

void sync();

// Pooled device staging buffers for host vectors (x in, y in/out): no cudaMalloc / cudaFree per call.
void* stage_acquire(size_t bytes);
void stage_release(void* dev, size_t bytes);

// Device mirror of a host array that does not change between calls (the arrays of an adopted matrix): uploaded on
// first use, found again by (address, size) and a fingerprint of 64 sampled words; thsp_host::invalidate(p) or
// releasing p drops it.  Returns p itself for device / managed memory.
const void* mirror_bytes(const void* p, size_t bytes);
template <class T>
inline const T* mirror(const T* p, size_t n)
{
    return static_cast<const T*>(mirror_bytes(p, n * sizeof(T)));
}
// Forget everything remembered about p (plan, mirror, pointer kind): call after rewriting an adopted array in place.
void invalidate(const void* p);

// A device-usable view of `n` elements at `p`.  Managed/device memory is used in place (managed
// is prefetched when `prefetch` is set); plain or pinned host memory is copied to a pooled
// device buffer, and copied back on commit() when the view is writable.
template <class T>
class View {
public:
    View(const T* p, size_t n, bool writable, bool prefetch = true) : host_(const_cast<T*>(p)), n_(n), writable_(writable)
    {
        const int k = kind(p);
        if (k == 1 || k == 2 || n == 0) {
            dev_ = host_;
            if (k == 2 && n && prefetch && first_gpu_use(p)) prefetch_traced(p, n * sizeof(T));
        } else {
            if (k < 0) die("pointer classification (CUDA runtime unusable)");
            dev_ = static_cast<T*>(stage_acquire(n * sizeof(T)));
            staged_ = true;
            ok(thsp_memcpy_h2d(dev_, p, n * sizeof(T), nullptr), "staging copy");
        }
    }
    ~View()
    {
        if (staged_) {
            if (writable_ && !committed_) commit();
            stage_release(dev_, n_ * sizeof(T));
        }
    }
    View(const View&) = delete;
    View& operator=(const View&) = delete;
    T* get() const { return dev_; }
    operator T*() const { return dev_; }
    void commit()
    {
        if (staged_ && writable_) {
            ok(thsp_memcpy_d2h(host_, dev_, n_ * sizeof(T), nullptr), "staging copy back");
            ok(thsp_stream_sync(nullptr), "sync");
        }
        committed_ = true;
    }

private:
    T* host_;
    T* dev_ = nullptr;
    size_t n_;
    bool writable_, staged_ = false, committed_ = false;
};

// Copy n elements between any two kinds of memory (host, managed, device).
void copy_bytes(void* dst, const void* src, size_t bytes);
template <class T>
inline void copy(T* dst, const T* src, size_t n)
{
    if (n) copy_bytes(dst, src, n * sizeof(T));
}

// One int from device-accessible memory without migrating its page to the host.
int peek_int(const int* p);

// CSR plan cache (kernel choice from the row-length histogram), keyed by the matrix arrays (device addresses).
// nnz < 0: look the matrix up by its arrays alone (a hit also returns the entry count through
// *nnz_out, saving the device read of row_ptr[nrow]); returns nullptr on a miss.
// A plan holds nothing derived from the CONTENTS of the arrays except the entry count and the kernel choice: the
// kernels read the entry count from row_ptr[nrow] themselves and refuse to run when it differs from the plan's
// (thsp_csr_plan_stale); CSRMatrixMatVector then rebuilds the plan and repeats the product.
thsp_csr_plan* csr_plan(int nrow, int ncol, int nnz, const int* row_ptr, const int* col_ind, const double* val, int* nnz_out = nullptr);
void forget_plans(const void* any_array);

// THSP_TRACE=1: wall time of a scope on stderr (where a run of the reference's driver spends its seconds)
class Trace {
public:
    explicit Trace(const char* what);
    ~Trace();

private:
    const char* what_;
    double t0_;
};

}  // namespace thsp_host
