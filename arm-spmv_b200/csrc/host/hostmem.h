// hostmem.h -- glue between the reference's raw-pointer classes and the C ABI (thsp.h).
//
// Ownership model (INTEGRATION.md "Ownership"):
//   * arrays the library allocates are CUDA managed memory (host-dereferenceable, GPU-resident
//     after a prefetch);
//   * arrays a caller hands to an adopting constructor are ordinary new[] memory; they are
//     staged through a temporary device buffer on every call and released with delete[].
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

void* alloc_managed_bytes(size_t bytes);
template <class T>
inline T* alloc(size_t n)
{
    return static_cast<T*>(alloc_managed_bytes((n ? n : 1) * sizeof(T)));
}

// 0 plain host, 1 device, 2 managed, 3 pinned host
inline int kind(const void* p) { return p ? thsp_pointer_kind(p) : 0; }

// Managed arrays are prefetched to the GPU the first time a kernel is about to read them and
// not again: cudaMemPrefetchAsync on resident pages still walks the range (measured ~1 ms per
// 20 MB), which would dominate a 13 us SpMV.  If host code writes into such an array later, its
// pages migrate back on the CPU fault and return to the GPU on the next kernel's page faults.
bool first_gpu_use(const void* p);
void prefetch_traced(const void* p, size_t bytes);   // prefetch to the GPU; THSP_TRACE=1 prints how long it took
void forget_gpu_use(const void* p);

// Release an array owned by one of the API classes, whichever way it was obtained.
template <class T>
inline void release(T*& p)
{
    if (!p) return;
    const int k = kind(p);
    forget_gpu_use(p);
    if (k == 1 || k == 2) ok(thsp_free(p), "cudaFree");
    else if (k == 3) ok(thsp_free_host(p), "cudaFreeHost");
    else delete[] p;
    p = nullptr;
}

void sync();

// A device-usable view of `n` elements at `p`.  Managed/device memory is used in place (managed
// is prefetched when `prefetch` is set); plain or pinned host memory is copied to a temporary
// device buffer, and copied back on commit() when the view is writable.
template <class T>
class View {
public:
    View(const T* p, size_t n, bool writable, bool prefetch = true) : host_(const_cast<T*>(p)), n_(n), writable_(writable)
    {
        const int k = kind(p);
        if (k == 1 || k == 2 || n == 0) {
            dev_ = host_;
            if (k == 2 && n && first_gpu_use(p) && prefetch) prefetch_traced(p, n * sizeof(T));
        } else {
            void* d = nullptr;
            ok(thsp_malloc(&d, n * sizeof(T)), "staging allocation");
            dev_ = static_cast<T*>(d);
            staged_ = true;
            ok(thsp_memcpy_h2d(dev_, p, n * sizeof(T), nullptr), "staging copy");
        }
    }
    ~View()
    {
        if (staged_) {
            if (writable_ && !committed_) commit();
            thsp_free(dev_);
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

// CSR plan cache (kernel choice from the row-length histogram), keyed by the matrix arrays.
// nnz < 0: look the matrix up by its arrays alone (a hit also returns the entry count through
// *nnz_out, saving the device read of row_ptr[nrow]); returns nullptr on a miss.
thsp_csr_plan* csr_plan(int nrow, int ncol, int nnz, const int* row_ptr, const int* col_ind, const double* val, int* nnz_out = nullptr);
void forget_plans(const void* any_array);

}  // namespace thsp_host
