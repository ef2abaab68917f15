// mytime.cpp -- seconds since the first call; the first call returns 0 (reference src/mytime.cpp).
#include "mytime.h"

#include <time.h>

double mytimer(void)
{
    static bool started = false;
    static struct timespec t0;
    struct timespec t;
    clock_gettime(CLOCK_REALTIME, &t);
    if (!started) {
        started = true;
        t0 = t;
        return 0.0;
    }
    return (double)(t.tv_sec - t0.tv_sec) + 1e-9 * (double)(t.tv_nsec - t0.tv_nsec);
}
