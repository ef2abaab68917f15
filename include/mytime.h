// mytime.h -- wall-clock timer of the arm-spmv API (reference include/mytime.h:4).
#ifndef MYTIME_H
#define MYTIME_H

// Seconds since the first call; the first call itself returns 0.0 (src/mytime.cpp:6-18).
double mytimer(void);

#endif  // MYTIME_H
