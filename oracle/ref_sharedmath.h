// oracle/ref_sharedmath.h -- TEST INFRASTRUCTURE: forced include (-include) for the
// "sharedmath" build of the unmodified reference sources.
//
// Redirects the libm calls the hot path makes (Common/Common/Math/Rotation.hpp:262-316 and the
// acosf call sites KalmanFilter6DOF.cpp:81,96,132 / QuadcopterLogic.cpp:429,491 /
// QuadcopterAttitudeController.hpp:50) to the deterministic routines of
// agri-fly_b200/csrc/agf_math.h, so that a CPU run and a GPU run execute the same arithmetic
// (hard part 1).  Every standard header the reference includes is pulled in
// FIRST, so the macros below only ever rewrite the reference's own call sites.
// sqrt/sqrtf/fabs/fabsf stay with the platform: they are exact (correctly rounded) everywhere.
// exp() (Motor.cpp:56, LowPassFilterFirstOrder.hpp:31) stays glibc on both sides: its argument is
// a per-run constant and the product evaluates it on the host with the same glibc.
#pragma once
#ifdef __cplusplus
#include <assert.h>
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <fstream>
#include <iostream>
#include <limits>
#include <memory>
#include <mutex>
#include <queue>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../agri-fly_b200/csrc/agf_math.h"

// errno-emulating wrappers: the reference detects acosf domain errors through errno
static inline float agf_ref_acosf(float x) {
  if (x > 1.0f || x < -1.0f) errno = EDOM;
  return agf_acosf(x);
}
static inline float agf_ref_asinf(float x) {
  if (x > 1.0f || x < -1.0f) errno = EDOM;
  return agf_asinf(x);
}
static inline double agf_ref_acos(double x) {
  if (x > 1.0 || x < -1.0) errno = EDOM;
  return agf_acos(x);
}
static inline double agf_ref_asin(double x) {
  if (x > 1.0 || x < -1.0) errno = EDOM;
  return agf_asin(x);
}

#define sinf agf_sinf
#define cosf agf_cosf
#define acosf agf_ref_acosf
#define asinf agf_ref_asinf
#define atan2f agf_atan2f
#define sin agf_sin
#define cos agf_cos
#define acos agf_ref_acos
#define asin agf_ref_asin
#define atan2 agf_atan2
#endif
