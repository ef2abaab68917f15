// mytime.h -- wall-clock timer of the arm-spmv API (reference include/mytime.h:4, src/mytime.cpp:6-18).
// Returns seconds elapsed since the first call; that first call itself returns 0.0.
// All SpMV entry points are synchronous, so bracketing them with mytimer() times finished GPU work.
#ifndef MYTIME_H
#define MYTIME_H

double mytimer(void);

#endif  // MYTIME_H
