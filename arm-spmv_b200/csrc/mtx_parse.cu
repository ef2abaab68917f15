// mtx_parse.cu -- the entry section of a Matrix Market coordinate file parsed on the GPU.
//
// Replaces the loop of COOMatrixRead (src/data_io.cpp:83-88):
//     for (i = 0; i < nz; i++) { fscanf(f, "%d %d %lg\n", &I, &J, &val); I--; J--; ... }
// which is the whole wall time of BASELINE configs[0] once the SpMV takes microseconds (5.2 M
// lines through fscanf, serial).  fscanf does not know about lines: it reads whitespace-separated
// tokens, three per entry.  So does this parser:
//   1. every byte that starts a token (non-space after a space) is counted per 4 KB block;
//   2. exclusive scan of the block counts (convert.cu) -> index of each block's first token;
//   3. the byte offset of token k is written to tok[k] for k < 3 nnz;
//   4. one thread per entry converts its three tokens (mtx_number.h: %d, %d, and %lg by
//      Eisel-Lemire - correctly rounded, the same bits as strtod) and stores row-1, col-1, value.
// Anything the fast conversions do not cover (inf/nan, hex floats, > 19 significant digits,
// malformed or missing tokens) sets *status = 1 and the caller runs the reference's scanf loop
// instead - the file's meaning is never guessed.  The banner and size line stay with mmio on
// the CPU (src/mmio.cpp).
#include <algorithm>

#include "common.cuh"
#include "mtx_number.h"

namespace thsp {

int exclusive_scan(int n, const int* in, int* out, cudaStream_t s);   // convert.cu

static constexpr int kTokThreads = 256;
static constexpr int kTokBytes = 16;                       // bytes per thread
static constexpr int kTokBlock = kTokThreads * kTokBytes;  // 4 KB of text per CTA

__device__ __forceinline__ unsigned token_starts(const char* __restrict__ text, size_t len, size_t base)
{
    // bit i set: byte base+i starts a token
    unsigned m = 0;
    if (base >= len) return 0;
    bool prev_space = base == 0 ? true : thsp_num::is_space((unsigned char)text[base - 1]);
    unsigned char c[kTokBytes];
    if (base + kTokBytes <= len && ((((uintptr_t)text) + base) & 15) == 0) {
        const uint4 v = *reinterpret_cast<const uint4*>(text + base);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < kTokBytes; ++i) c[i] = (unsigned char)(w[i >> 2] >> (8 * (i & 3)));
    } else {
#pragma unroll
        for (int i = 0; i < kTokBytes; ++i) c[i] = base + i < len ? (unsigned char)text[base + i] : (unsigned char)' ';
    }
#pragma unroll
    for (int i = 0; i < kTokBytes; ++i) {
        const bool sp = thsp_num::is_space(c[i]);
        if (!sp && prev_space) m |= 1u << i;
        prev_space = sp;
    }
    return m;
}

__device__ __forceinline__ int block_excl_scan_256(int v, int* total)
{
    __shared__ int wt[kTokThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) wt[w] = inc;
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < kTokThreads / 32; ++i) {
        if (i < w) before += wt[i];
        tot += wt[i];
    }
    __syncthreads();
    *total = tot;
    return before + inc - v;
}

__global__ void __launch_bounds__(kTokThreads) mtx_count_kernel(const char* __restrict__ text, size_t len, int* __restrict__ bcnt)
{
    const size_t base = (size_t)blockIdx.x * kTokBlock + (size_t)threadIdx.x * kTokBytes;
    int tot;
    block_excl_scan_256(__popc(token_starts(text, len, base)), &tot);
    if (threadIdx.x == 0) bcnt[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kTokThreads) mtx_offsets_kernel(const char* __restrict__ text, size_t len,
                                                                  const int* __restrict__ boff, int64_t ntok_max,
                                                                  unsigned* __restrict__ tok)
{
    const size_t base = (size_t)blockIdx.x * kTokBlock + (size_t)threadIdx.x * kTokBytes;
    unsigned m = token_starts(text, len, base);
    int tot;
    int64_t k = (int64_t)boff[blockIdx.x] + block_excl_scan_256(__popc(m), &tot);
    while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        if (k < ntok_max) tok[k] = (unsigned)(base + i);
        ++k;
    }
}

__device__ __forceinline__ const char* token_end(const char* p, const char* end)
{
    while (p < end && !thsp_num::is_space((unsigned char)*p)) ++p;
    return p;
}

__global__ void __launch_bounds__(256) mtx_parse_kernel(const char* __restrict__ text, size_t len, const unsigned* __restrict__ tok,
                                                        int nnz, int* __restrict__ ri, int* __restrict__ ci, double* __restrict__ va,
                                                        int* __restrict__ bad)
{
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= nnz) return;
    const char* end = text + len;
    const char* a = text + tok[3 * (size_t)e];
    const char* b = text + tok[3 * (size_t)e + 1];
    const char* c = text + tok[3 * (size_t)e + 2];
    int i = 0, j = 0;
    double v = 0.0;
    const bool ok = thsp_num::parse_int_token(a, token_end(a, end), &i) && thsp_num::parse_int_token(b, token_end(b, end), &j) &&
                    thsp_num::parse_double_token(c, token_end(c, end), &v);
    if (!ok) {
        *bad = 1;   // benign race: every writer stores 1
        return;
    }
    ri[e] = i - 1;   // the file is 1-based (src/data_io.cpp:86-87)
    ci[e] = j - 1;
    va[e] = v;
}

}  // namespace thsp

using namespace thsp;

extern "C" {

int thsp_mtx_parse_coo(const char* text_host, size_t len, int nnz, int* row_ind, int* col_ind, double* val, int* status,
                       thsp_stream_t stream)
{
    if (ensure_device()) return 1;
    THSP_REQUIRE(status != nullptr, "status is required");
    *status = 1;
    if (nnz <= 0) {
        *status = 0;
        return 0;
    }
    if (len == 0 || len >= ((size_t)1 << 32)) return 0;   // token offsets are 32-bit
    cudaStream_t s = as_stream(stream);
    const int nblk = div_up((int64_t)len, kTokBlock);
    const int64_t ntok = 3 * (int64_t)nnz;
    char* text = static_cast<char*>(scratch(len + 16, 6));
    unsigned* tok = static_cast<unsigned*>(scratch(sizeof(unsigned) * (size_t)ntok + sizeof(int) * (2 * (size_t)nblk + 4), 7));
    if (!text || !tok) return 1;
    int* bcnt = reinterpret_cast<int*>(tok + ntok);
    int* boff = bcnt + nblk + 1;
    int* bad = boff + nblk + 1;
    THSP_CUDA(cudaMemcpyAsync(text, text_host, len, cudaMemcpyHostToDevice, s));
    THSP_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
    mtx_count_kernel<<<nblk, kTokThreads, 0, s>>>(text, len, bcnt);
    THSP_LAUNCH_CHECK();
    if (exclusive_scan(nblk, bcnt, boff, s)) return 1;
    int total = 0;
    THSP_CUDA(cudaMemcpyAsync(&total, boff + nblk, sizeof(int), cudaMemcpyDeviceToHost, s));
    THSP_CUDA(cudaStreamSynchronize(s));
    if ((int64_t)total < ntok) return 0;   // the file ends early: let the scanf loop report it the reference's way
    mtx_offsets_kernel<<<nblk, kTokThreads, 0, s>>>(text, len, boff, ntok, tok);
    THSP_LAUNCH_CHECK();
    mtx_parse_kernel<<<div_up(nnz, 256), 256, 0, s>>>(text, len, tok, nnz, row_ind, col_ind, val, bad);
    THSP_LAUNCH_CHECK();
    int h_bad = 0;
    THSP_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    THSP_CUDA(cudaStreamSynchronize(s));
    *status = h_bad ? 1 : 0;
    return 0;
}

}  // extern "C"
