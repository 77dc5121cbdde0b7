// oracle/ref_sharedmath_rappids.h -- TEST INFRASTRUCTURE: forced include (-include) for the
// "sharedmath" build of the unmodified reference RAPPIDS planner sources.
//
// On top of oracle/ref_sharedmath.h (cos / acos -> agf_math.h) it redirects the two uses the planner
// makes of pow(): squares, pow(x, 2) (RapidTrajectoryGenerator.cpp:100-110, SingleAxisTrajectory.cpp:166-172),
// and the cube root of the cubic solver, pow(x, 1./3) (Common/Common/Math/RootFinder.hpp:81).
#pragma once
#include "ref_sharedmath.h"
#ifdef __cplusplus
static inline double agf_ref_pow(double x, double y) {
  if (y == 2.0) return x * x;
  if (y == 1. / 3) return agf_cbrt_pos(x);
  fprintf(stderr, "agf_ref_pow: unexpected exponent %g\n", y);
  abort();
}
#define pow agf_ref_pow
#endif
