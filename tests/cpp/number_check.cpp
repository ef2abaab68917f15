// number_check.cpp -- the GPU Matrix Market parser's number conversion (csrc/mtx_number.h, compiled here
// for the host) against strtod / strtol, the functions behind the reference's fscanf("%d %d %lg")
// (src/data_io.cpp:85).  Usage: number_check <cases> <seed>; prints "ok <cases> <fallbacks>" or the first mismatch.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <random>

#include "mtx_number.h"

static uint64_t bits(double d)
{
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
}

static int check(const char* s, long* fallbacks)
{
    double mine = 0.0;
    const bool ok = thsp_num::parse_double_token(s, s + strlen(s), &mine);
    char* endp = nullptr;
    const double ref = strtod(s, &endp);
    if (!ok) {
        ++*fallbacks;
        return 0;
    }
    if (*endp != 0 || bits(mine) != bits(ref)) {
        printf("MISMATCH '%s': mine %.17g (%016llx) strtod %.17g (%016llx)\n", s, mine, (unsigned long long)bits(mine), ref,
               (unsigned long long)bits(ref));
        return 1;
    }
    return 0;
}

int main(int argc, char** argv)
{
    const long cases = argc > 1 ? atol(argv[1]) : 1000000;
    std::mt19937_64 rng(argc > 2 ? atol(argv[2]) : 1);
    long fallbacks = 0, done = 0;
    char buf[256];
    // fixed known-hard inputs: halfway cases, subnormals, limits, signed zeros
    const char* fixed[] = {"0", "-0", "0.0", "-0.0e10", "1", "-1", "4", "26", "1e0", "1e22", "1e23", "9007199254740993", "9007199254740992",
                           "9007199254740991", "4.9e-324", "4.9406564584124654e-324", "2.4703282292062327e-324", "2.4703282292062328e-324",
                           "2.2250738585072014e-308", "2.2250738585072011e-308", "2.2250738585072009e-308", "1.7976931348623157e308",
                           "1.7976931348623158e308", "1.7976931348623159e308", "1e309", "1e-400", "123456789012345678", "1234567890123456789",
                           "0.1", "0.2", "0.3", "0.30000000000000004", "5e-324", "1.0000000000000002", "1.00000000000000011102230246251565",
                           "8.5", ".5", "5.", "+.5e+1", "1E5", "1e-5", "000123.456000", "1.5e-310", "3.1415926535897932", "2.718281828459045",
                           "100000000000000000000", "100000000000000000001", "0.000000000000000000000000000001", "1e+0005"};
    for (const char* s : fixed)
        if (check(s, &fallbacks)) return 1;
    std::uniform_real_distribution<double> u01(0.0, 1.0);
    for (; done < cases; ++done) {
        const int kind = (int)(rng() % 8);
        if (kind == 0) {   // %.17g of a uniform(0,1) value: what the synthetic .mtx files hold
            snprintf(buf, sizeof buf, "%.17g", u01(rng));
        } else if (kind == 1) {   // %.17g / %.16g / %g of a random bit pattern (all exponents, subnormals)
            uint64_t b = rng();
            double d;
            memcpy(&d, &b, 8);
            if (d != d || d - d != 0.0) d = 1.0;
            const char* fmts[] = {"%.17g", "%.16g", "%.15g", "%g", "%.17e", "%.3f"};
            const char* f = fmts[rng() % 6];
            if (f[2] == '3' && (d > 1e15 || d < -1e15)) f = "%.17g";
            snprintf(buf, sizeof buf, f, d);
        } else if (kind == 2) {   // random digit strings with a random exponent
            const int nd = 1 + (int)(rng() % 19);
            int p = 0;
            if (rng() & 1) buf[p++] = '-';
            const int dot = (int)(rng() % (nd + 1));
            for (int i = 0; i < nd; ++i) {
                if (i == dot) buf[p++] = '.';
                buf[p++] = (char)('0' + rng() % 10);
            }
            const int e = (int)(rng() % 700) - 350;
            snprintf(buf + p, sizeof buf - p, "e%d", e);
        } else if (kind == 3) {   // halfway between two doubles, exactly (integers 2^53 .. 2^64 and .5 ulp decimals)
            const uint64_t m = (1ULL << 53) + (rng() >> 11) * 2 + 1;   // odd: halfway when shifted
            snprintf(buf, sizeof buf, "%llu", (unsigned long long)m);
        } else if (kind == 4) {   // small integers and simple decimals, as in hand-written matrices
            snprintf(buf, sizeof buf, "%d.%02d", (int)(rng() % 2000) - 1000, (int)(rng() % 100));
        } else if (kind == 5) {   // near the subnormal boundary
            const double d = 2.2250738585072014e-308 * (0.5 + u01(rng));
            snprintf(buf, sizeof buf, "%.17g", d);
        } else if (kind == 6) {   // 19 significant digits
            snprintf(buf, sizeof buf, "%llu.%09llue%d", (unsigned long long)(1 + rng() % 9), (unsigned long long)(rng() % 1000000000ULL) ,
                     (int)(rng() % 600) - 300);
            // append 9 more digits to the fraction: rewrite as d.ffffffffffffffffffeX
            char t[256];
            char* ep = strchr(buf, 'e');
            const int e = atoi(ep + 1);
            *ep = 0;
            snprintf(t, sizeof t, "%s%09llue%d", buf, (unsigned long long)(rng() % 1000000000ULL), e);
            strcpy(buf, t);
        } else {   // near the overflow boundary
            const double d = 1.7976931348623157e308 * (0.9 + 0.1 * u01(rng));
            snprintf(buf, sizeof buf, "%.17g", d);
        }
        if (check(buf, &fallbacks)) return 1;
    }
    // %d
    for (long i = 0; i < 200000; ++i) {
        const long v = (long)(rng() % 4294967296ULL) - 2147483648L;
        snprintf(buf, sizeof buf, (rng() & 1) && v >= 0 ? "+%ld" : "%ld", v);
        int mine = 0;
        if (!thsp_num::parse_int_token(buf, buf + strlen(buf), &mine) || mine != (int)strtol(buf, nullptr, 10)) {
            printf("MISMATCH int '%s' -> %d\n", buf, mine);
            return 1;
        }
    }
    printf("ok %ld %ld\n", done, fallbacks);
    return 0;
}
