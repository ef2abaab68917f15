#include "hostmem.h"

#include <cuda_runtime_api.h>
#include <string.h>

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

thsp_csr_plan* csr_plan(int nrow, int ncol, int nnz, const int* row_ptr, const int* col_ind, const double* val)
{
    auto& v = plans();
    for (auto& e : v)
        if (e.row_ptr == row_ptr && e.col_ind == col_ind && e.val == val && e.nrow == nrow && e.ncol == ncol && e.nnz == nnz)
            return e.plan;
    if (v.size() >= 16) {  // small cache: drop the oldest
        thsp_csr_plan_destroy(v.front().plan);
        v.erase(v.begin());
    }
    thsp_csr_plan* p = nullptr;
    ok(thsp_csr_plan_create(&p, nrow, ncol, nnz, row_ptr, col_ind, val, 8, nullptr), "CSR plan");
    v.push_back(PlanEntry{row_ptr, col_ind, val, nrow, ncol, nnz, p});
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
