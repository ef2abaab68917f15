#include "hostmem.h"

#include <cuda_runtime_api.h>
#include <stdint.h>
#include <string.h>
#include <time.h>

#include <mutex>
#include <new>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace thsp_host {

void die(const char* what)
{
    // Same convention as the reference's reader (src/data_io.cpp:53-75): message, exit(1).
    fprintf(stderr, "*** TH_sparse (B200): %s failed: %s ***\n", what, thsp_last_error());
    exit(1);
}

namespace {
// Everything below is process-wide bookkeeping; the API is not re-entrant by design, the lock only
// keeps the tables intact when a caller does use several host threads.
std::recursive_mutex& mu()
{
    static std::recursive_mutex* m = new std::recursive_mutex();   // never destroyed: objects with static storage release late
    return *m;
}
struct Tables {
    std::unordered_map<const void*, size_t> owned;        // library allocations (managed memory): address -> bytes
    std::unordered_map<const void*, int> foreign_kind;    // what the runtime said about pointers we did not allocate
    std::unordered_set<const void*> seen_on_gpu;
    struct Mirror {
        size_t bytes;
        void* dev;
        uint64_t print;
    };
    std::unordered_map<const void*, Mirror> mirrors;
    struct Stage {
        void* dev;
        size_t cap;
        bool busy;
    };
    std::vector<Stage> stages;
};
Tables& tb()
{
    static Tables* t = new Tables();   // leaked on purpose, like the mutex
    return *t;
}

double now_ms()
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
}
bool trace_on()
{
    static const bool on = getenv("THSP_TRACE") && getenv("THSP_TRACE")[0] == '1';
    return on;
}

// 64 words spread over the array + its size: cheap to recompute on every call (a few cache misses), changes whenever
// a caller refills an adopted array; single-element edits need thsp_host::invalidate.
uint64_t fingerprint(const void* p, size_t bytes)
{
    uint64_t h = 0x9E3779B97F4A7C15ull ^ bytes;
    const size_t words = bytes / 8;
    const unsigned char* b = static_cast<const unsigned char*>(p);
    auto mix = [&](uint64_t v) {
        h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    };
    if (words == 0) {
        for (size_t i = 0; i < bytes; ++i) mix(b[i]);
        return h;
    }
    const size_t n = words < 64 ? words : 64;
    for (size_t k = 0; k < n; ++k) {
        const size_t w = n > 1 ? k * (words - 1) / (n - 1) : 0;
        uint64_t v;
        memcpy(&v, b + w * 8, 8);
        mix(v);
    }
    return h;
}
}  // namespace

void* alloc_managed_bytes(size_t bytes)
{
    void* p = nullptr;
    ok(thsp_malloc_managed(&p, bytes), "managed allocation");
    // A fresh managed array has no pages anywhere; the kernel that fills it (every converting constructor) would fault
    // them in one by one.  Asking for them in HBM now moves nothing and halves the constructors' only call in main.cpp
    // (5-point Laplacian: CSR 5.9 -> 3.0 ms, CSC 4.9 -> 1.5, ELL 3.8 -> 1.2, DIA 2.2 -> 0.9; five runs of five alike - the
    // stall of cudaMemPrefetchAsync described at prefetch_traced() concerns pages the host has touched).  An array the
    // host fills instead (Vector::FillRandom) simply takes its pages back on first touch.  THSP_POPULATE=0 turns it off.
    static const bool populate = !(getenv("THSP_POPULATE") && getenv("THSP_POPULATE")[0] == '0');
    if (populate && bytes >= (1u << 20) && thsp_prefetch(p, bytes, 1, nullptr) != 0) { /* advisory only */ }
    std::lock_guard<std::recursive_mutex> lk(mu());
    tb().owned[p] = bytes;
    return p;
}

int kind(const void* p)
{
    if (!p) return 0;
    std::lock_guard<std::recursive_mutex> lk(mu());
    Tables& t = tb();
    if (t.owned.count(p)) return 2;
    auto it = t.foreign_kind.find(p);
    if (it != t.foreign_kind.end()) return it->second;
    const int k = thsp_pointer_kind(p);
    if (k >= 0) {
        if (t.foreign_kind.size() > 4096) t.foreign_kind.clear();
        t.foreign_kind[p] = k;
    }
    return k;
}

void release_bytes(void* p)
{
    if (!p) return;
    std::lock_guard<std::recursive_mutex> lk(mu());
    Tables& t = tb();
    t.seen_on_gpu.erase(p);
    auto m = t.mirrors.find(p);
    if (m != t.mirrors.end()) {
        thsp_free(m->second.dev);
        t.mirrors.erase(m);
    }
    auto o = t.owned.find(p);
    if (o != t.owned.end()) {
        t.owned.erase(o);
        thsp_free(p);   // a failure here means the runtime is already going down (process exit): nothing left to do
        return;
    }
    t.foreign_kind.erase(p);
    const int k = thsp_pointer_kind(p);   // asked afresh: the caller may have re-registered or re-allocated the address
    if (k == 0) ::operator delete[](p);   // the reference's contract for adopted arrays: new[] memory (src/matrix.cpp:31-39)
    else if (k == 1 || k == 2) thsp_free(p);
    else if (k == 3) thsp_free_host(p);
    // k < 0: the runtime cannot classify the pointer any more - leaking beats handing CUDA memory to delete[]
}

void sync() { ok(thsp_stream_sync(nullptr), "stream synchronise"); }

void prefetch_traced(const void* p, size_t bytes)
{
    // Managed pages the host has touched come back to the GPU here, once per array, in front of the first kernel that
    // reads them.  cudaMemPrefetchAsync does that at ~35 GB/s - and on the B200 box (driver 580) one such call in every
    // 3 to 15 takes 60-940 ms instead of 0.6 ms (profiles/r02_uvm_prefetch_probe.txt: any round, any array, with or
    // without a warm-up; queuing several prefetches without waiting showed the same 613 ms in round 1).  In the reference's
    // main.cpp that call sits INSIDE the timed COO loop (the host reads the COO arrays at :46-52 just before) and turned
    // "### COO CPU GFLOPS" from 110 into 1 in half of the runs.  A kernel that simply READS the array lets the GPU's page
    // faults pull the pages over: ~8 GB/s (2.7 ms for 21 MB), and the same in 8 of 8 runs.  THSP_PREFETCH=async brings
    // cudaMemPrefetchAsync back.
    static const bool async = getenv("THSP_PREFETCH") && !strcmp(getenv("THSP_PREFETCH"), "async");
    const double t0 = trace_on() ? now_ms() : 0.0;
    if (!async && bytes >= 16) {
        uint64_t h = 0;
        ok(thsp_hash_f64((int64_t)(bytes / 8), static_cast<const double*>(p), 0, &h, nullptr), "page touch");
    } else {
        ok(thsp_prefetch(p, bytes, 1, nullptr), "prefetch");
    }
    ok(thsp_stream_sync(nullptr), "sync");
    if (trace_on()) fprintf(stderr, "[thsp] %s %p %.1f MB: %.3f ms\n", async ? "prefetch" : "fault-in", p, bytes / 1e6, now_ms() - t0);
}
bool first_gpu_use(const void* p)
{
    std::lock_guard<std::recursive_mutex> lk(mu());
    return tb().seen_on_gpu.insert(p).second;
}
void forget_gpu_use(const void* p)
{
    std::lock_guard<std::recursive_mutex> lk(mu());
    tb().seen_on_gpu.erase(p);
}

void* stage_acquire(size_t bytes)
{
    std::lock_guard<std::recursive_mutex> lk(mu());
    Tables& t = tb();
    Tables::Stage* best = nullptr;
    for (auto& s : t.stages)
        if (!s.busy && s.cap >= bytes && s.cap <= 4 * bytes + 4096 && (!best || s.cap < best->cap)) best = &s;
    if (best) {
        best->busy = true;
        return best->dev;
    }
    if (t.stages.size() >= 24) {   // drop the idle buffers before growing further
        for (size_t i = 0; i < t.stages.size();)
            if (!t.stages[i].busy) {
                thsp_free(t.stages[i].dev);
                t.stages.erase(t.stages.begin() + i);
            } else {
                ++i;
            }
    }
    const size_t cap = (bytes + 255) & ~(size_t)255;
    void* d = nullptr;
    ok(thsp_malloc(&d, cap ? cap : 256), "staging allocation");
    t.stages.push_back(Tables::Stage{d, cap, true});
    return d;
}
void stage_release(void* dev, size_t)
{
    std::lock_guard<std::recursive_mutex> lk(mu());
    for (auto& s : tb().stages)
        if (s.dev == dev) {
            s.busy = false;
            return;
        }
}

const void* mirror_bytes(const void* p, size_t bytes)
{
    if (!p || bytes == 0) return p;
    const int k = kind(p);
    if (k == 1 || k == 2) return p;
    if (k < 0) die("pointer classification (CUDA runtime unusable)");
    std::lock_guard<std::recursive_mutex> lk(mu());
    Tables& t = tb();
    const uint64_t fp = fingerprint(p, bytes);
    auto it = t.mirrors.find(p);
    if (it != t.mirrors.end() && it->second.bytes == bytes && it->second.print == fp) return it->second.dev;
    const double t0 = trace_on() ? now_ms() : 0.0;
    void* d = nullptr;
    if (it != t.mirrors.end() && it->second.bytes == bytes) {
        d = it->second.dev;   // same array refilled: upload again into the same buffer
    } else {
        if (it != t.mirrors.end()) {
            thsp_free(it->second.dev);
            t.mirrors.erase(it);
        }
        ok(thsp_malloc(&d, bytes), "device mirror allocation");
    }
    copy_bytes(d, p, bytes);
    t.mirrors[p] = Tables::Mirror{bytes, d, fp};
    if (trace_on()) fprintf(stderr, "[thsp] device mirror of %p, %.1f MB: %.3f ms\n", p, bytes / 1e6, now_ms() - t0);
    return d;
}

void invalidate(const void* p)
{
    if (!p) return;
    std::lock_guard<std::recursive_mutex> lk(mu());
    Tables& t = tb();
    auto m = t.mirrors.find(p);
    if (m != t.mirrors.end()) {
        forget_plans(m->second.dev);
        thsp_free(m->second.dev);
        t.mirrors.erase(m);
    }
    t.foreign_kind.erase(p);
    forget_plans(p);
}

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
    static std::vector<PlanEntry>* v = new std::vector<PlanEntry>();
    return *v;
}
}  // namespace

thsp_csr_plan* csr_plan(int nrow, int ncol, int nnz, const int* row_ptr, const int* col_ind, const double* val, int* nnz_out)
{
    std::lock_guard<std::recursive_mutex> lk(mu());
    auto& v = plans();
    for (auto& e : v)
        if (e.row_ptr == row_ptr && e.col_ind == col_ind && e.val == val && e.nrow == nrow && e.ncol == ncol && (nnz < 0 || e.nnz == nnz)) {
            if (nnz_out) *nnz_out = e.nnz;
            return e.plan;
        }
    if (nnz < 0) return nullptr;
    forget_plans(row_ptr);   // same arrays, another entry count: the old plan is stale
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
    if (!a) return;
    std::lock_guard<std::recursive_mutex> lk(mu());
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

static double trace_origin()
{
    static const double t = now_ms();
    return t;
}
Trace::Trace(const char* what) : what_(what), t0_(0.0)
{
    if (what_ && trace_on()) {
        trace_origin();
        t0_ = now_ms();
    }
}
Trace::~Trace()
{
    if (what_ && trace_on()) fprintf(stderr, "[thsp] at %9.1f ms  %-28s %10.3f ms\n", t0_ - trace_origin(), what_, now_ms() - t0_);
}

}  // namespace thsp_host
