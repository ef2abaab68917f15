// mtx_number.h -- text -> int / binary64 for the Matrix Market entry parser, host + device.
//
// COOMatrixRead reads every entry with fscanf("%d %d %lg\n") (src/data_io.cpp:83-88); %lg is
// strtod, i.e. the correctly rounded binary64 of the decimal string.  decimal_to_double() below is
// the Eisel-Lemire conversion ("Number Parsing at a Gigabyte per Second", Lemire 2021, with the
// no-fallback result of Mushtak & Lemire 2023): the decimal significand w (up to 19 digits, exact
// in 64 bits) times a 128-bit truncated power of five gives enough bits to round correctly for
// every (w, q).  Strings it does not take - more than 19 significant digits, inf/nan, hex floats,
// anything that is not [+-]digits[.digits][e[+-]digits] - are reported, and the caller falls back
// to the reference's own scanf loop for that file.  The same functions compile for the host,
// where tests/ checks them against strtod.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define THSP_HD __host__ __device__ __forceinline__
#else
#define THSP_HD inline
#endif

namespace thsp_num {

static const uint64_t h_pow5[2 * 651] = {
#include "pow5_table.inc"
};
#if defined(__CUDACC__)
__device__ static const uint64_t d_pow5[2 * 651] = {
#include "pow5_table.inc"
};
#endif
#if defined(__CUDA_ARCH__)
#define THSP_POW5 d_pow5
#else
#define THSP_POW5 h_pow5
#endif

THSP_HD int clz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}
THSP_HD void mul64(uint64_t a, uint64_t b, uint64_t* hi, uint64_t* lo)
{
#if defined(__CUDA_ARCH__)
    *lo = a * b;
    *hi = __umul64hi(a, b);
#else
    const unsigned __int128 p = (unsigned __int128)a * b;
    *lo = (uint64_t)p;
    *hi = (uint64_t)(p >> 64);
#endif
}
THSP_HD double bits_to_double(uint64_t b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    union { uint64_t u; double d; } c;
    c.u = b;
    return c.d;
#endif
}

// binary64 nearest to w * 10^q (w != 0 handled too), round half to even.
THSP_HD double decimal_to_double(uint64_t w, int64_t q, bool negative)
{
    const uint64_t sign = negative ? 0x8000000000000000ULL : 0ULL;
    if (w == 0 || q < -342) return bits_to_double(sign);
    if (q > 308) return bits_to_double(sign | 0x7FF0000000000000ULL);
    const int lz = clz64(w);
    w <<= lz;
    // w * 5^q, 128-bit table entry, refined with the low half when the first product cannot decide
    const uint64_t* t = THSP_POW5 + 2 * (q + 342);
    uint64_t hi, lo;
    mul64(w, t[0], &hi, &lo);
    const uint64_t precision_mask = 0xFFFFFFFFFFFFFFFFULL >> 55;   // 52 explicit bits + 3
    if ((hi & precision_mask) == precision_mask) {
        uint64_t hi2, lo2;
        mul64(w, t[1], &hi2, &lo2);
        lo += hi2;
        if (hi2 > lo) ++hi;
    }
    const int upperbit = (int)(hi >> 63);
    const int shift = upperbit + 64 - 52 - 3;
    uint64_t mantissa = hi >> shift;
    int64_t power2 = (((152170 + 65536) * q) >> 16) + 63 + upperbit - lz + 1023;
    if (power2 <= 0) {   // subnormal
        if (-power2 + 1 >= 64) return bits_to_double(sign);
        mantissa >>= -power2 + 1;
        mantissa += (mantissa & 1);
        mantissa >>= 1;
        power2 = (mantissa < (1ULL << 52)) ? 0 : 1;
        return bits_to_double(sign | ((uint64_t)power2 << 52) | (mantissa & ~(1ULL << 52)));
    }
    // exactly half way between two doubles: round to even
    if (lo <= 1 && q >= -4 && q <= 23 && (mantissa & 3) == 1) {
        if ((mantissa << shift) == hi) mantissa &= ~1ULL;
    }
    mantissa += (mantissa & 1);
    mantissa >>= 1;
    if (mantissa >= (2ULL << 52)) {
        mantissa = 1ULL << 52;
        ++power2;
    }
    mantissa &= ~(1ULL << 52);
    if (power2 >= 0x7FF) return bits_to_double(sign | 0x7FF0000000000000ULL);
    return bits_to_double(sign | ((uint64_t)power2 << 52) | mantissa);
}

THSP_HD bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }   // isspace in the C locale
THSP_HD bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }

// One whitespace-delimited token [p, end) as %d.  false: not [+-]digits or outside int range.
THSP_HD bool parse_int_token(const char* p, const char* end, int* out)
{
    bool neg = false;
    if (p < end && (*p == '+' || *p == '-')) neg = *p++ == '-';
    if (p == end) return false;
    int64_t v = 0;
    for (; p < end; ++p) {
        if (!is_digit((unsigned char)*p)) return false;
        v = v * 10 + (*p - '0');
        if (v > 2147483648LL) return false;
    }
    if (neg) v = -v;
    if (v > 2147483647LL) return false;
    *out = (int)v;
    return true;
}

// One token as %lg.  false: a form decimal_to_double() does not cover (caller falls back to strtod).
THSP_HD bool parse_double_token(const char* p, const char* end, double* out)
{
    bool neg = false;
    if (p < end && (*p == '+' || *p == '-')) neg = *p++ == '-';
    uint64_t w = 0;
    int digits = 0;          // significant digits taken into w
    int64_t q = 0;           // decimal exponent of w
    bool any = false, dropped_nonzero = false;
    for (; p < end && is_digit((unsigned char)*p); ++p) {
        any = true;
        const int d = *p - '0';
        if (digits == 0 && d == 0) continue;   // leading zeros
        if (digits < 19) {
            w = w * 10 + d;
            ++digits;
        } else {
            ++q;   // digit beyond the 19th: scales the value, must be zero to stay exact
            if (d) dropped_nonzero = true;
        }
    }
    if (p < end && *p == '.') {
        ++p;
        for (; p < end && is_digit((unsigned char)*p); ++p) {
            any = true;
            const int d = *p - '0';
            if (digits == 0 && d == 0) {
                --q;
                continue;
            }
            if (digits < 19) {
                w = w * 10 + d;
                ++digits;
                --q;
            } else if (d) {
                dropped_nonzero = true;
            }
        }
    }
    if (!any) return false;
    if (p < end && (*p == 'e' || *p == 'E')) {
        ++p;
        bool eneg = false;
        if (p < end && (*p == '+' || *p == '-')) eneg = *p++ == '-';
        if (p == end) return false;
        int64_t e = 0;
        for (; p < end; ++p) {
            if (!is_digit((unsigned char)*p)) return false;
            if (e < 100000) e = e * 10 + (*p - '0');
        }
        q += eneg ? -e : e;
    }
    if (p != end) return false;            // trailing characters: not a plain decimal number
    if (dropped_nonzero) return false;     // more than 19 significant digits: leave it to strtod
    *out = decimal_to_double(w, q, neg);
    return true;
}

}  // namespace thsp_num
