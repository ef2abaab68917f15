#include "hostmem.h"

#include <cuda_runtime_api.h>
#include <string.h>
#include <time.h>

#include <unordered_set>
#include <vector>

namespace thsp_host {

void die(const char* what)
{
    // Same convention as the reference's reader (src/data_io.cpp:53-75): message, exit(1).
    fprintf(stderr, "*** TH_sparse (B200): %s failed: %s ***\n", what, thsp_last_error());
    exit(1);
}

void* alloc_managed_bytes(size_t bytes)
{
    void* p = nullptr;
    ok(thsp_malloc_managed(&p, bytes), "managed allocation");
    return p;
}

void sync() { ok(thsp_stream_sync(nullptr), "stream synchronise"); }

static std::unordered_set<const void*>& seen_on_gpu()
{
    static std::unordered_set<const void*> s;
    return s;
}
void prefetch_traced(const void* p, size_t bytes)
{
    static const bool trace = getenv("THSP_TRACE") && getenv("THSP_TRACE")[0] == '1';
    if (!trace) {
        // Synchronise after each first-use prefetch: queuing several prefetches and a kernel behind
        // them without waiting took 613 ms for 92 MB on the B200 box (driver 580), 2.8 ms with it.
        ok(thsp_prefetch(p, bytes, 1, nullptr), "prefetch");
        ok(thsp_stream_sync(nullptr), "sync");
        return;
    }
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    ok(thsp_prefetch(p, bytes, 1, nullptr), "prefetch");
    ok(thsp_stream_sync(nullptr), "sync");
    clock_gettime(CLOCK_MONOTONIC, &b);
    fprintf(stderr, "[thsp] prefetch %p %.1f MB: %.3f ms\n", p, bytes / 1e6, (b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) * 1e-6);
}
bool first_gpu_use(const void* p) { return seen_on_gpu().insert(p).second; }
void forget_gpu_use(const void* p) { seen_on_gpu().erase(p); }

void copy_bytes(void* dst, const void* src, size_t bytes)
{
    if (cudaMemcpy(dst, src, bytes, cudaMemcpyDefault) != cudaSuccess) {
        fprintf(stderr, "*** TH_sparse (B200): cudaMemcpy failed: %s ***\n", cudaGetErrorString(cudaGetLastError()));
        exit(1);
    }
}

int peek_int(const int* p)
{
    const int k = kind(p);
    if (k == 0 || k == 3) return *p;
    int v = 0;
    copy_bytes(&v, p, sizeof(int));
    return v;
}

namespace {
struct PlanEntry {
    const int* row_ptr;
    const int* col_ind;
    const double* val;
    int nrow, ncol, nnz;
    thsp_csr_plan* plan;
};
std::vector<PlanEntry>& plans()
{
    static std::vector<PlanEntry> v;
    return v;
}
}  // namespace

thsp_csr_plan* csr_plan(int nrow, int ncol, int nnz, const int* row_ptr, const int* col_ind, const double* val, int* nnz_out)
{
    auto& v = plans();
    for (auto& e : v)
        if (e.row_ptr == row_ptr && e.col_ind == col_ind && e.val == val && e.nrow == nrow && e.ncol == ncol && (nnz < 0 || e.nnz == nnz)) {
            if (nnz_out) *nnz_out = e.nnz;
            return e.plan;
        }
    if (nnz < 0) return nullptr;
    if (v.size() >= 16) {  // small cache: drop the oldest
        thsp_csr_plan_destroy(v.front().plan);
        v.erase(v.begin());
    }
    thsp_csr_plan* p = nullptr;
    ok(thsp_csr_plan_create(&p, nrow, ncol, nnz, row_ptr, col_ind, val, 8, nullptr), "CSR plan");
    if (getenv("THSP_AUTOTUNE") && getenv("THSP_AUTOTUNE")[0] == '1') ok(thsp_csr_plan_autotune(p, nullptr), "CSR autotune");
    v.push_back(PlanEntry{row_ptr, col_ind, val, nrow, ncol, nnz, p});
    if (nnz_out) *nnz_out = nnz;
    return p;
}

void forget_plans(const void* a)
{
    auto& v = plans();
    for (size_t i = 0; i < v.size();) {
        if (v[i].row_ptr == a || v[i].col_ind == a || v[i].val == a) {
            thsp_csr_plan_destroy(v[i].plan);
            v.erase(v.begin() + i);
        } else {
            ++i;
        }
    }
}

}  // namespace thsp_host
