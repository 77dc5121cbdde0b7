// oracle/ref_sharedmath_mocap.h -- TEST INFRASTRUCTURE: forced include for the shared-math build of the reference's
// Components/Offboard/MocapStateEstimator.cpp.  Everything ref_sharedmath.h redirects, plus exp(): the estimator's
// exp(-dtInt / tau) (MocapStateEstimator.cpp:96,164) has a data-dependent argument, so CPU and GPU need one routine.
#pragma once
#include "ref_sharedmath.h"
#ifdef __cplusplus
#define exp agf_exp
#endif
