// Stub so that g++ can compile openkite_b200/csrc/kite_model.cuh for the HOST in the CPU test-suite.
// Test infrastructure only: lets `-m "not gpu"` tests check the device model's arithmetic (analytic
// Jacobians, RK4 staging) against the oracle without a GPU.  The product never includes this file.
#pragma once
#include <cmath>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
#include <cstring>
using std::fma; using std::sqrt; using std::asin; using std::atan2; using std::exp; using std::fabs; using std::fmin; using std::fmax;
