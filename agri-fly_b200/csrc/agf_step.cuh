// agf_step.cuh -- the per-vehicle simulation step as sm_100a device code.
//
// One vehicle per thread, the whole vehicle state in registers across all ticks of a launch;
// state and per-vehicle parameters live in HBM as structure-of-arrays of 16-byte quads
// (float4 / double2 / uint4), so every load/store instruction of a warp is a run of fully
// coalesced 128-bit accesses; parameters shared by the population ride in the kernel-parameter
// constant bank (broadcast reads).
//
// What is computed (reference file:line, paths under Components/Components/ unless noted):
//   motors         Simulation/Motor.cpp:39-84
//   rigid body     Simulation/Quadcopter_T.cpp:86-156
//   IMU synthesis  Simulation/Quadcopter_T.cpp:159-183
//   IMU intake     Logic/QuadcopterLogic.hpp:32-59  (+ Common/Common/Math/LowPassFilterSecondOrder.hpp:51-63)
//   estimator      Logic/KalmanFilter6DOF.cpp:33-309
//   state machine  Logic/QuadcopterLogic.cpp:164-391
//   controllers    Logic/QuadcopterLogic.cpp:393-588, Logic/Quadcopter{Position,Attitude,AngularVelocity}Controller.hpp,
//                  Logic/QuadcopterMixer.hpp:63-99
//   UWB exchange   Simulation/Quadcopter_T.cpp:191-199, Simulation/UWBNetwork.cpp:22-89
//   time base      Common/Common/Time/Timer.hpp:27-53 (integer microseconds)
//
// Template axes:  P = plant real (double = the reference's mixed precision, float = FP32 mode);
// PARITY = bit-comparable arithmetic (TU compiled with -fmad=false, agf_math.h libm) vs fast;
// UWB = the vehicle ranges against anchors, so the 9x9 covariance is live (81 more registers);
// HK = full housekeeping (telemetry warnings, battery/temperature filters, rate monitors,
// propeller calibration).
//
// The EKF covariance algebra exploits the structural zeros/ones of the transition matrix f and of
// the measurement row H but keeps the reference's sequential-k summation order for the remaining
// terms, so it is bit-identical to the dense products of KalmanFilter6DOF.cpp:232,267-269,296
// for finite inputs (adding an exact +-0 never changes a sum); tests/test_parity_gpu.py checks it.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include <stdlib.h>
#include <string.h>
#include <mutex>

#include "agf_math.h"
#include "agf_types.h"
#include "agf_launch.h"

namespace agf {

#define AGF_DEV __host__ __device__ __forceinline__
// Rarely taken paths (estimator resets, start-up alignment, the external-acceleration controller, optional noise
// sources) are real calls: the per-tick loop then spans ~half the bytes, which matters because it does not fit
// the SM's instruction caches (profiles/: "no_instruction" was the largest single stall reason).
#define AGF_COLD __host__ __device__ __noinline__
#if defined(__CUDA_ARCH__)
#define AGF_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#define AGF_UNLIKELY(x) (x)
#endif

// ---------------------------------------------------------------------------------------------
// libm policy
// ---------------------------------------------------------------------------------------------
template<bool PARITY>
struct Mf;
template<>
struct Mf<true> {
  static AGF_DEV float sin(float x) { return agf_sinf(x); }
  static AGF_DEV float cos(float x) { return agf_cosf(x); }
  static AGF_DEV float asin(float x) { return agf_asinf(x); }
  static AGF_DEV float acos(float x) { return agf_acosf(x); }
  static AGF_DEV float atan2(float y, float x) { return agf_atan2f(y, x); }
  static AGF_DEV double sin(double x) { return agf_sin(x); }
  static AGF_DEV double cos(double x) { return agf_cos(x); }
  static AGF_DEV double acos(double x) { return agf_acos(x); }
  static AGF_DEV double asin(double x) { return agf_asin(x); }
  static AGF_DEV double exp(double x) { return agf_exp(x); }
};
// Fast variants: the step only ever takes sin/cos of HALF rotation angles per tick (|x| << 1), so a
// short odd/even polynomial (error < 1e-9 relative for |x| <= 0.5) serves the hot path and the
// general libm routine is an out-of-line fallback that keeps the code small.
static __device__ __noinline__ float slow_sinf(float x) { return ::sinf(x); }
static __device__ __noinline__ float slow_cosf(float x) { return ::cosf(x); }
static __device__ __noinline__ double slow_sin(double x) { return ::sin(x); }
static __device__ __noinline__ double slow_cos(double x) { return ::cos(x); }
template<>
struct Mf<false> {
#if defined(__CUDA_ARCH__)
  static AGF_DEV float sin(float x) {
    if (::fabsf(x) > 0.5f) return slow_sinf(x);
    const float z = x * x;
    float p = ::fmaf(z, 2.75573192e-6f, -1.98412698e-4f);
    p = ::fmaf(z, p, 8.33333333e-3f);
    p = ::fmaf(z, p, -1.66666667e-1f);
    return ::fmaf(x * z, p, x);
  }
  static AGF_DEV float cos(float x) {
    if (::fabsf(x) > 0.5f) return slow_cosf(x);
    const float z = x * x;
    float p = ::fmaf(z, 2.48015873e-5f, -1.38888889e-3f);
    p = ::fmaf(z, p, 4.16666667e-2f);
    p = ::fmaf(z, p, -0.5f);
    return ::fmaf(z, p, 1.0f);
  }
  static AGF_DEV double sin(double x) {
    if (::fabs(x) > 0.5) return slow_sin(x);
    const double z = x * x;
    double p = ::fma(z, 1.6059043836821613e-10, -2.5052108385441720e-8);
    p = ::fma(z, p, 2.7557319223985893e-6);
    p = ::fma(z, p, -1.9841269841269841e-4);
    p = ::fma(z, p, 8.3333333333333333e-3);
    p = ::fma(z, p, -1.6666666666666667e-1);
    return ::fma(x * z, p, x);
  }
  static AGF_DEV double cos(double x) {
    if (::fabs(x) > 0.5) return slow_cos(x);
    const double z = x * x;
    double p = ::fma(z, -1.1470745597729725e-11, 2.0876756987868099e-9);
    p = ::fma(z, p, -2.7557319223985888e-7);
    p = ::fma(z, p, 2.4801587301587302e-5);
    p = ::fma(z, p, -1.3888888888888889e-3);
    p = ::fma(z, p, 4.1666666666666664e-2);
    p = ::fma(z, p, -0.5);
    return ::fma(z, p, 1.0);
  }
#else
  static AGF_DEV float sin(float x) { return ::sinf(x); }
  static AGF_DEV float cos(float x) { return ::cosf(x); }
  static AGF_DEV double sin(double x) { return ::sin(x); }
  static AGF_DEV double cos(double x) { return ::cos(x); }
#endif
  static AGF_DEV float asin(float x) { return ::asinf(x); }
  static AGF_DEV float acos(float x) { return ::acosf(x); }
  static AGF_DEV float atan2(float y, float x) { return ::atan2f(y, x); }
  static AGF_DEV double acos(double x) { return ::acos(x); }
  static AGF_DEV double asin(double x) { return ::asin(x); }
  static AGF_DEV double exp(double x) { return ::exp(x); }
};
// division: IEEE in the parity variant, reciprocal-multiply in the fast ones
template<bool PARITY> AGF_DEV float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
  if (!PARITY) return __fdividef(a, b);
#endif
  return a / b;
}
template<bool PARITY> AGF_DEV double fdiv(double a, double b) { return a / b; }
AGF_DEV float rsqrtf_(float x) {  // 1/sqrt(x)
#if defined(__CUDA_ARCH__)
  return ::rsqrtf(x);
#else
  return 1.0f / ::sqrtf(x);
#endif
}
AGF_DEV float rsqrt_(float x) { return ::sqrtf(x); }
AGF_DEV double rsqrt_(double x) { return ::sqrt(x); }
AGF_DEV float pmin_(float a, float b) { return ::fminf(a, b); }
AGF_DEV double pmin_(double a, double b) { return ::fmin(a, b); }
AGF_DEV float pmax_(float a, float b) { return ::fmaxf(a, b); }
AGF_DEV double pmax_(double a, double b) { return ::fmax(a, b); }
AGF_DEV float rabs_(float x) { return ::fabsf(x); }
AGF_DEV double rabs_(double x) { return ::fabs(x); }
// Packed FP32 (sm_100: fma.rn.f32x2 / mul.f32x2 / add.f32x2 -> FFMA2 / FMUL2 / FADD2, two lanes per issued instruction; a
// scalar operand is broadcast for free).  Used by the fast variants where two lanes share their coefficients.
AGF_DEV float2 f2_fma(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__)
  return __ffma2_rn(a, b, c);
#else
  return make_float2(::fmaf(a.x, b.x, c.x), ::fmaf(a.y, b.y, c.y));
#endif
}
AGF_DEV float2 f2_mul(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return __fmul2_rn(a, b);
#else
  return make_float2(a.x * b.x, a.y * b.y);
#endif
}
AGF_DEV float2 f2_add(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
AGF_DEV float2 f2_fma(float2 a, float s, float2 c) { return f2_fma(a, make_float2(s, s), c); }
AGF_DEV float2 f2_mul(float2 a, float s) { return f2_mul(a, make_float2(s, s)); }

// ---------------------------------------------------------------------------------------------
// Vec3 / Rotation with the reference's operation order
// (Common/Common/Math/Vec3.hpp, Common/Common/Math/Rotation.hpp)
// ---------------------------------------------------------------------------------------------
template<typename R>
struct V3 {
  R x, y, z;
  AGF_DEV V3() {}
  AGF_DEV V3(R a, R b, R c) : x(a), y(b), z(c) {}
};
template<typename R> AGF_DEV V3<R> operator+(const V3<R>& a, const V3<R>& b) { return V3<R>(a.x + b.x, a.y + b.y, a.z + b.z); }
template<typename R> AGF_DEV V3<R> operator-(const V3<R>& a, const V3<R>& b) { return V3<R>(a.x - b.x, a.y - b.y, a.z - b.z); }
template<typename R> AGF_DEV V3<R> operator*(R s, const V3<R>& v) { return V3<R>(s * v.x, s * v.y, s * v.z); }
template<typename R> AGF_DEV V3<R> operator*(const V3<R>& v, R s) { return V3<R>(s * v.x, s * v.y, s * v.z); }
template<typename R> AGF_DEV V3<R> operator/(const V3<R>& v, R s) { return V3<R>(v.x / s, v.y / s, v.z / s); }
template<typename R> AGF_DEV R dot(const V3<R>& a, const V3<R>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template<typename R> AGF_DEV V3<R> cross(const V3<R>& a, const V3<R>& b) {
  return V3<R>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template<typename R> AGF_DEV R norm(const V3<R>& a) { return rsqrt_(dot(a, a)); }
template<bool PARITY, typename R> AGF_DEV V3<R> vdiv(const V3<R>& v, R s) {
  if (PARITY) return V3<R>(v.x / s, v.y / s, v.z / s);
  const R inv = fdiv<false>(R(1), s);
  return V3<R>(v.x * inv, v.y * inv, v.z * inv);
}

// Matrix<Real,3,3> * Vec3 (Vec3.hpp:202-210): accumulate from 0 in column order
template<typename R>
AGF_DEV V3<R> matvec(const R* m, const V3<R>& v) {
  R o0 = R(0), o1 = R(0), o2 = R(0);
  o0 += m[0] * v.x; o0 += m[1] * v.y; o0 += m[2] * v.z;
  o1 += m[3] * v.x; o1 += m[4] * v.y; o1 += m[5] * v.z;
  o2 += m[6] * v.x; o2 += m[7] * v.y; o2 += m[8] * v.z;
  return V3<R>(o0, o1, o2);
}

template<typename R>
struct Q4 {
  R w, x, y, z;
  AGF_DEV Q4() {}
  AGF_DEV Q4(R a, R b, R c, R d) : w(a), x(b), y(c), z(d) {}
};
template<typename R> AGF_DEV Q4<R> qinv(const Q4<R>& q) { return Q4<R>(q.w, -q.x, -q.y, -q.z); }
// Rotation::operator* (Rotation.hpp:124-131): a * r1
template<typename R>
AGF_DEV Q4<R> qmul(const Q4<R>& a, const Q4<R>& r1) {
  R c0 = r1.w * a.w - r1.x * a.x - r1.y * a.y - r1.z * a.z;
  R c1 = r1.x * a.w + r1.w * a.x + r1.z * a.y - r1.y * a.z;
  R c2 = r1.y * a.w - r1.z * a.x + r1.w * a.y + r1.x * a.z;
  R c3 = r1.z * a.w + r1.y * a.x - r1.x * a.y + r1.w * a.z;
  return Q4<R>(c0, c1, c2, c3);
}
// GetRotationMatrix (Rotation.hpp:196-217)
template<typename R>
AGF_DEV void qmatrix(const Q4<R>& q, R* Rm) {
  const R r0 = q.w * q.w, r1 = q.x * q.x, r2 = q.y * q.y, r3 = q.z * q.z;
  Rm[0] = r0 + r1 - r2 - r3;
  Rm[1] = 2 * q.x * q.y - 2 * q.w * q.z;
  Rm[2] = 2 * q.x * q.z + 2 * q.w * q.y;
  Rm[3] = 2 * q.x * q.y + 2 * q.w * q.z;
  Rm[4] = r0 - r1 + r2 - r3;
  Rm[5] = 2 * q.y * q.z - 2 * q.w * q.x;
  Rm[6] = 2 * q.x * q.z - 2 * q.w * q.y;
  Rm[7] = 2 * q.y * q.z + 2 * q.w * q.x;
  Rm[8] = r0 - r1 - r2 + r3;
}
// Rotate (Rotation.hpp:236-245)
template<typename R>
AGF_DEV V3<R> qrot(const Q4<R>& q, const V3<R>& in) {
  R Rm[9];
  qmatrix(q, Rm);
  return V3<R>(Rm[0] * in.x + Rm[1] * in.y + Rm[2] * in.z, Rm[3] * in.x + Rm[4] * in.y + Rm[5] * in.z,
               Rm[6] * in.x + Rm[7] * in.y + Rm[8] * in.z);
}
// third row of the rotation matrix applied to e3: (att * (0,0,1)).z, with the exact-zero products kept out
template<typename R>
AGF_DEV R qrot_e3_z(const Q4<R>& q) {
  // Rm[6]*0 + Rm[7]*0 + Rm[8]*1 == Rm[8] for finite Rm
  return q.w * q.w - q.x * q.x - q.y * q.y + q.z * q.z;
}
// FromAxisAngle (Rotation.hpp:92-97)
template<bool PARITY, typename R>
AGF_DEV Q4<R> q_from_axis_angle(const V3<R>& u, R angle) {
  const R h = angle * R(0.5);
  const R s = Mf<PARITY>::sin(h);
  return Q4<R>(Mf<PARITY>::cos(h), s * u.x, s * u.y, s * u.z);
}
// FromRotationVector (Rotation.hpp:84-89); returns false (identity) below MIN_ANGLE
template<bool PARITY, typename R>
AGF_DEV bool q_from_rotvec(const V3<R>& rv, Q4<R>& out) {
  const R theta = norm(rv);
  if (theta < R(4.84813681e-6)) return false;
  out = q_from_axis_angle<PARITY>(vdiv<PARITY>(rv, theta), theta);
  return true;
}
// Fast variants: FromRotationVector as two even polynomials of the half angle h, q = (cos h, rv * sin(h)/(2h)),
// evaluated in z = h^2 = |rv|^2/4 -- no square root, no division, no range test for sin and cos separately.
// |h| <= 0.5 covers every per-tick rotation (|w| < 500 rad/s at 2 ms); larger angles take the general routine.
template<typename R> struct RotvecPoly;
template<> struct RotvecPoly<float> {
  static AGF_DEV float cosp(float z) {  // cos(h), error < 3e-10 for z <= 0.25
    float p = ::fmaf(z, 2.48015873e-5f, -1.38888889e-3f);
    p = ::fmaf(z, p, 4.16666667e-2f);
    p = ::fmaf(z, p, -0.5f);
    return ::fmaf(z, p, 1.0f);
  }
  static AGF_DEV float half_sinc(float z) {  // sin(h)/(2h)
    float p = ::fmaf(z, 1.37786596e-6f, -9.92063492e-5f);
    p = ::fmaf(z, p, 4.16666667e-3f);
    p = ::fmaf(z, p, -8.33333333e-2f);
    return ::fmaf(z, p, 0.5f);
  }
};
template<> struct RotvecPoly<double> {
  static AGF_DEV double cosp(double z) {
    double p = ::fma(z, 4.7794773323873853e-14, -1.1470745597729725e-11);
    p = ::fma(z, p, 2.0876756987868099e-9);
    p = ::fma(z, p, -2.7557319223985888e-7);
    p = ::fma(z, p, 2.4801587301587302e-5);
    p = ::fma(z, p, -1.3888888888888889e-3);
    p = ::fma(z, p, 4.1666666666666664e-2);
    p = ::fma(z, p, -0.5);
    return ::fma(z, p, 1.0);
  }
  static AGF_DEV double half_sinc(double z) {
    double p = ::fma(z, 1.4056851100590741e-15, -3.8238541122497209e-13);
    p = ::fma(z, p, 8.0295219184108065e-11);
    p = ::fma(z, p, -1.2526054192720860e-8);
    p = ::fma(z, p, 1.3778659611992946e-6);
    p = ::fma(z, p, -9.9206349206349206e-5);
    p = ::fma(z, p, 4.1666666666666666e-3);
    p = ::fma(z, p, -8.3333333333333329e-2);
    return ::fma(z, p, 0.5);
  }
};
template<typename R>
static AGF_COLD Q4<R> q_from_rotvec_general(V3<R> rv) {  // angles beyond the polynomial's range
  Q4<R> d(R(1), R(0), R(0), R(0));
  q_from_rotvec<false>(rv, d);
  return d;
}
template<bool PARITY, typename R>
AGF_DEV Q4<R> q_apply_rotvec(const Q4<R>& a, const V3<R>& rv) {  // a * FromRotationVector(rv)
  if constexpr (PARITY) {
    Q4<R> d;
    if (!q_from_rotvec<PARITY>(rv, d)) return a;  // a * Identity == a exactly
    return qmul(a, d);
  } else {
    const R z = R(0.25) * dot(rv, rv);
    Q4<R> d;
    if (AGF_UNLIKELY(z > R(0.25))) {
      d = q_from_rotvec_general<R>(rv);
    } else {
      // below MIN_ANGLE the reference returns the identity (Rotation.hpp:84-88): theta^2 < MIN_ANGLE^2 <=> z < MIN_ANGLE^2/4
      const bool tiny = z < R(5.8761074e-12);
      const R k = tiny ? R(0) : RotvecPoly<R>::half_sinc(z);
      d = Q4<R>(tiny ? R(1) : RotvecPoly<R>::cosp(z), k * rv.x, k * rv.y, k * rv.z);
    }
    return qmul(a, d);
  }
}
// FromEulerYPR (Rotation.hpp:99-110)
template<bool PARITY>
AGF_DEV Q4<float> q_from_euler_ypr(float y, float p, float r) {
  typedef Mf<PARITY> M;
  const float h = 0.5f;
  Q4<float> o;
  o.w = M::cos(h * y) * M::cos(h * p) * M::cos(h * r) + M::sin(h * y) * M::sin(h * p) * M::sin(h * r);
  o.x = M::cos(h * y) * M::cos(h * p) * M::sin(h * r) - M::sin(h * y) * M::sin(h * p) * M::cos(h * r);
  o.y = M::cos(h * y) * M::sin(h * p) * M::cos(h * r) + M::sin(h * y) * M::cos(h * p) * M::sin(h * r);
  o.z = M::sin(h * y) * M::cos(h * p) * M::cos(h * r) - M::cos(h * y) * M::sin(h * p) * M::sin(h * r);
  return o;
}
// ToVectorPartOfQuaternion / ToRotationVector (Rotation.hpp:144-161)
template<bool PARITY>
AGF_DEV V3<float> q_to_rotvec(const Q4<float>& q) {
  V3<float> n = q.w > 0 ? V3<float>(q.x, q.y, q.z) : V3<float>(-q.x, -q.y, -q.z);
  const float nn = norm(n);
  const float angle = Mf<PARITY>::asin(nn) * 2;
  if (angle < 4.84813681e-6f) return V3<float>(0, 0, 0);
  return n * fdiv<PARITY>(angle, nn);
}
// acosf with the reference's errno fallback (KalmanFilter6DOF.cpp:95-103)
template<bool PARITY>
AGF_DEV float acos_guarded(float c) {
  float a = Mf<PARITY>::acos(c);
  if (c > 1.0f || c < -1.0f) a = c < 0 ? 3.14159274f : 0.0f;  // float(M_PI)
  return a;
}

// ---------------------------------------------------------------------------------------------
// LowPassFilterSecondOrder::Apply (LowPassFilterSecondOrder.hpp:51-63); st = {xm0, xm1, ym0, ym1}
// ---------------------------------------------------------------------------------------------
AGF_DEV float lpf2(const Lpf2Coef& c, float* st, int stride, float in) {
  float out = c.b2 * in;
  out = out + (c.b0 * st[0] + c.b1 * st[stride]);
  out = out + ((-c.a1) * st[2 * stride] - c.a2 * st[3 * stride]);
  st[0] = st[stride];
  st[stride] = in;
  st[2 * stride] = st[3 * stride];
  st[3 * stride] = out;
  return out;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator and Box-Muller normals (replaces the per-object
// std::default_random_engine of Quadcopter_T.hpp:122-123 and the global mt19937 of UWBNetwork.cpp:4)
// ---------------------------------------------------------------------------------------------
AGF_DEV float bits_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
AGF_DEV uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return uint32_t((uint64_t(a) * uint64_t(b)) >> 32);
#endif
}
// Round keys key + r * (0x9E3779B9, 0xBB67AE85) are the same for every draw of a batch: the host tabulates them
// (philox_round_keys, StepShared::philox_rk) and the rounds read them from the constant bank.
#ifndef AGF_PHILOX_ROUNDS
#define AGF_PHILOX_ROUNDS 10
#endif
AGF_HDI void philox_round_keys(uint64_t seed, uint32_t* rk /* [2 * AGF_PHILOX_ROUNDS] */) {
  uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
  for (int r = 0; r < AGF_PHILOX_ROUNDS; r++) {
    rk[2 * r] = k0;
    rk[2 * r + 1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
template<int ROUNDS>
AGF_DEV uint4 philox4x32(uint4 ctr, const uint32_t* rk) {
#pragma unroll
  for (int r = 0; r < ROUNDS; r++) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ rk[2 * r], lo1, hi0 ^ ctr.w ^ rk[2 * r + 1], lo0);
  }
  return ctr;
}
// Box-Muller from two 21-bit uniforms: u1 in (0,1] (tails reach 5.4 sigma), angle = 2 pi u2.  The uniforms are
// assembled in the mantissa of a float in [1,2) -- no integer-to-float conversion (XU pipe) needed.
AGF_DEV void box_muller21(uint32_t f1, uint32_t f2, float& n0, float& n1) {
#if defined(__CUDA_ARCH__)
  const float a = __uint_as_float(0x3f800000u | (f1 << 2));  // 1 + u, u = f1 / 2^21
  const float b = __uint_as_float(0x3f800000u | (f2 << 2));
  const float u1 = 2.0f - a;                                 // (0, 1]
  const float r = ::sqrtf(-1.38629436f * __log2f(u1));       // sqrt(-2 ln u1)
  float sn, cs;
  __sincosf(6.28318530718f * (b - 1.0f), &sn, &cs);
#else
  const float u1 = 1.0f - float(f1) * (1.0f / 2097152.0f);
  const float r = ::sqrtf(-2.0f * ::logf(u1));
  const float ang = 6.28318530718f * (float(f2) * (1.0f / 2097152.0f));
  const float sn = ::sinf(ang), cs = ::cosf(ang);
#endif
  n0 = r * cs;
  n1 = r * sn;
}
// 6 standard normals for (vehicle, cycle, stream): ONE Philox4x32-10 block, its 128 bits cut into six 21-bit fields
AGF_DEV void normals6(const uint32_t* rk, uint64_t vehicle, uint32_t cycle, uint32_t stream, float* n) {
  const uint4 r = philox4x32<AGF_PHILOX_ROUNDS>(make_uint4(uint32_t(vehicle), uint32_t(vehicle >> 32), cycle, stream), rk);
  const uint32_t m = 0x1FFFFFu;
  box_muller21(r.x & m, ((r.x >> 21) | (r.y << 11)) & m, n[0], n[1]);
  box_muller21((r.y >> 10) & m, r.z & m, n[2], n[3]);
  box_muller21(((r.z >> 21) | (r.w << 11)) & m, (r.w >> 10) & m, n[4], n[5]);
}
struct Normals6 {
  float n[6];
};
static AGF_COLD Normals6 normals6_cold(uint64_t seed, uint64_t vehicle, uint32_t cycle, uint32_t stream) {
  Normals6 o;
  uint32_t rk[2 * AGF_PHILOX_ROUNDS];
  philox_round_keys(seed, rk);
  normals6(rk, vehicle, cycle, stream, o.n);
  return o;
}

// One completed UWB range (UWBNetwork.cpp:62-75): the uniform that decides "outlier?", the normal of the additive range
// noise and the normal of an outlier, from ONE Philox block of stream 2, counter (vehicle, tick).
struct UwbDraw {
  float u, n_noise, n_outlier;
};
static AGF_COLD UwbDraw uwb_draw_cold(uint64_t seed, uint64_t vehicle, uint32_t tick) {
  uint32_t rk[2 * AGF_PHILOX_ROUNDS];
  philox_round_keys(seed, rk);
  const uint4 r = philox4x32<AGF_PHILOX_ROUNDS>(make_uint4(uint32_t(vehicle), uint32_t(vehicle >> 32), tick, 2u), rk);
  const uint32_t m = 0x1FFFFFu;
  UwbDraw d;
  box_muller21(r.x & m, ((r.x >> 21) | (r.y << 11)) & m, d.n_noise, d.n_outlier);
  d.u = float(r.w >> 8) * (1.0f / 16777216.0f);  // [0, 1)
  return d;
}

// ---------------------------------------------------------------------------------------------
// register-resident vehicle state
// ---------------------------------------------------------------------------------------------
template<typename P, bool PARITY, bool UWB, bool HK>
struct VState {
  // plant
  P pos[3], vel[3], att[4], w[3], ms[4];
  // FP32 fast variants: compensation terms of the position / velocity sums (compensated integration in tick())
  P cpos[(!PARITY && sizeof(P) == 4) ? 3 : 1], cvel[(!PARITY && sizeof(P) == 4) ? 3 : 1];
  // logic
  float cmd[4];      // _desMotorSpeeds == _motorSpeedCommands after every logic run
  float dforce[PARITY ? 4 : 1];   // _desMotorForcesForTelemetry (fast variants: derived from cmd at the end of a launch)
  float radio_f[4];  // floats[0..3] of the last radio message (the only ones any controller reads)
  // IMU low-pass states [xm0 xm1 ym0 ym1] x 3 components, component-major [4*c + k]: in registers in the
  // parity variant, in the thread's shared-memory scratch in the fast variants (Scratch below)
  float gyro_lp[PARITY ? 12 : 1], acc_lp[PARITY ? 12 : 1];
  float kpos[3], kvel[3], kw[3], katt[4], kcorr[3];
  float uwb_range;    // range held by the vehicle's radio (UWBRadio::_meas.range)
  float logic_range;  // range handed to the logic (QuadcopterLogic::_uwbRangeMeas.range)
  uint32_t bits;      // see B_* below
  uint32_t cnt;       // numMeasRejectedSequentially (bits 0-7) | telemetry warnings (bits 24-31)
  uint32_t cycle;
  uint32_t kfcnt;     // numResets (lo 16) | numMeasRejected (hi 16)
  uint32_t uwb_count;
  uint32_t uwbw;      // anchor indices: network responder | radio meas responder | logic meas target | next target
  uint32_t age_radio, age_uwb, age_est_reset;
  // housekeeping (HK)
  // battery-voltage and temperature low-pass states: parity variants only.  Both filters have a CONSTANT input and are
  // initialised at their fixed point (QuadcopterLogic.cpp:138-139; 25 C), so the fast variants take the fixed point -- the
  // filtered voltage is the supply voltage -- instead of iterating two second-order filters per tick towards where they
  // already are (the float recursion only dithers in the last bits); their stored states are left untouched.
  float temp_lp[PARITY ? 4 : 1], batt_lp[PARITY ? 4 : 1];
  float batt_vfilt, mon_cmd_lpdt, mon_loop_lpdt;
  float pc_accum[PARITY ? 4 : 1], pc_corr[PARITY ? 4 : 1];  // fast variants: in the scratch (SQ_PC_*)
  uint32_t pc_count, age_mon_cmd, age_mon_loop;
  // estimator covariance (UWB): full 9x9 in registers in the parity variant (the reference's predict step
  // does not keep it exactly symmetric); packed upper triangle in shared-memory scratch in the fast variants
  float cov[(UWB && PARITY) ? 81 : 1];
  // radio true position latched at the last logic run (UWBRadio::_uwbTruePosition)
  P rpos[UWB ? 3 : 1];
  // offboard-loop command queue payloads (parity variant; the fast variants keep them in the scratch)
  float offq[PARITY ? 4 * AGF_OFFQ : 1];
};

// bit layout of VState::bits
enum : uint32_t {
  B_FS_SHIFT = 0, B_FS_MASK = 0x7u,            // flight state
  B_PANIC_SHIFT = 3, B_PANIC_MASK = 0x7u,      // first panic reason
  B_RTYPE_SHIFT = 6, B_RTYPE_MASK = 0x7u,      // radio message type
  B_RFLAGS_SHIFT = 9, B_RFLAGS_MASK = 0xFFu,   // radio message flags
  B_RADIO_NEW = 1u << 17,
  B_IMU_INIT = 1u << 18,
  B_UWB_INIT = 1u << 19,
  B_UWB_NEW = 1u << 20,        // logic's _uwbRangeMeas.isNew
  B_KF_RESET_SEEN = 1u << 21,  // _numResets != _lastCheckNumResets
  B_PC_RUNNING = 1u << 22,     // propeller calibration running
  B_RADIO_MEAS_NEW = 1u << 23  // UWBRadio::_meas.haveNew
};
// byte lanes of VState::uwbw
enum : uint32_t { W_NET_RESP = 0, W_RADIO_RESP = 8, W_LOGIC_TARGET = 16, W_NEXT_TARGET = 24 };
AGF_DEV uint32_t bget(uint32_t w, uint32_t shift, uint32_t mask) { return (w >> shift) & mask; }
AGF_DEV uint32_t bset(uint32_t w, uint32_t shift, uint32_t mask, uint32_t v) {
  return (w & ~(mask << shift)) | ((v & mask) << shift);
}

// Thread-private scratch in shared memory (fast variants): quads laid out [quad][thread] so that a
// warp's 128-bit accesses are conflict free.  Holds what is touched once per tick but would otherwise
// pin ~70 registers for the whole tick: the six IMU low-pass states and the packed EKF covariance.
struct Scratch {
  float4* q;   // already offset by the thread index
};
// threads per block of the variants that use the scratch: a compile-time stride turns every scratch address
// into base register + immediate and lets the compiler tell the quads apart (no false LDS/STS dependencies)
enum { SQ_STRIDE = AGF_BLOCK_THREADS };
// after the filters / covariance: three quads of propeller-calibration state (per-motor thrust correction, its derived
// 1 / (correction * kF), the calibration accumulators) -- read once per tick or rarer, so not worth nine registers;
// then AGF_OFFQ quads (offboard queue) when the loop is on
enum { SQ_LPF = 0, SQ_COV = 6, SQ_PC_NOUWB = 6, SQ_PC_UWB = 18, SQ_QUADS_NOUWB = 9, SQ_QUADS_UWB = 21 };
enum { SQ_PC_CORR = 0, SQ_PC_INV = 1, SQ_PC_ACCUM = 2 };
// Scratch accesses are volatile 128-bit shared-memory instructions: with the compile-time stride the compiler
// could otherwise forward a tick's stores to the next tick's loads, i.e. keep the whole scratch in registers
// across the loop -- the opposite of what the scratch is for (measured: 3x the local-memory spill traffic).
AGF_DEV float4 sq_load(const Scratch& sc, int quad) {
#if defined(__CUDA_ARCH__)
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(uint32_t(__cvta_generic_to_shared(sc.q)) + uint32_t(quad * SQ_STRIDE * sizeof(float4))));
  return v;
#else
  return sc.q[quad * SQ_STRIDE];
#endif
}
AGF_DEV void sq_store(const Scratch& sc, int quad, const float4& v) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(uint32_t(__cvta_generic_to_shared(sc.q)) + uint32_t(quad * SQ_STRIDE * sizeof(float4))),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
#else
  sc.q[quad * SQ_STRIDE] = v;
#endif
}
// Packed symmetric 9x9 covariance of the fast variants: 48 floats = 12 scratch quads, arranged for the packed FP32 path
// (FFMA2: two lanes per instruction, operands in even-aligned register pairs).  Blocks p (position), v (velocity),
// a (attitude error).  Rows 0 and 1 of a 3x3 block are interleaved, one pair per column:
//   pairs 0-2  (Ppv[0][c], Ppv[1][c])    pairs 3-5  (Ppa[0][c], Ppa[1][c])    pairs 6-8   (Pva[0][c], Pva[1][c])
//   pairs 9-11 (Pvv[0][c], Pvv[1][c])    pairs 12-14 (Paa[0][c], Paa[1][c])   (pair k = floats 2k, 2k+1)
// then row 2 of Ppv, Ppa, Pva (floats 30-38), Pvv[2][2], Paa[2][2] (39, 40) and the upper triangle of Ppp (41-46).
// Row 2 of the symmetric blocks Pvv / Paa is read from the column-2 pair; their (1,0) entry exists twice, as the
// upper half of the column-0 pair and as the lower half of the column-1 pair (CI_VV10 mirrors (3,4), CI_AA10 mirrors (6,7)).
enum { CP_PV = 0, CP_PA = 3, CP_VA = 6, CP_VV = 9, CP_AA = 12, CI_PV2 = 30, CI_PA2 = 33, CI_VA2 = 36, CI_VV22 = 39, CI_AA22 = 40,
       CI_PP = 41, CI_VV10 = 2 * CP_VV + 1, CI_AA10 = 2 * CP_AA + 1 };
AGF_DEV constexpr int SI_rect(int pair0, int row2, int r, int c) { return r < 2 ? 2 * (pair0 + c) + r : row2 + c; }
AGF_DEV constexpr int SI_sym(int pair0, int s22, int r, int c) {  // r <= c
  return c < 2 ? 2 * (pair0 + c) + r : (r < 2 ? 2 * (pair0 + 2) + r : s22);
}
AGF_DEV constexpr int SI(int i, int j) {  // any (i, j); the canonical place of a symmetric block's entry is its upper one
  if (i > j) { const int t = i; i = j; j = t; }
  const int bi = i / 3, bj = j / 3, r = i % 3, c = j % 3;
  if (bi == 0 && bj == 0) return CI_PP + (r == 0 ? c : (r == 1 ? 2 + c : 5));
  if (bi == 0 && bj == 1) return SI_rect(CP_PV, CI_PV2, r, c);
  if (bi == 0) return SI_rect(CP_PA, CI_PA2, r, c);
  if (bi == 1 && bj == 2) return SI_rect(CP_VA, CI_VA2, r, c);
  if (bi == 1) return SI_sym(CP_VV, CI_VV22, r, c);
  return SI_sym(CP_AA, CI_AA22, r, c);
}
AGF_DEV void cov_fill_mirrors(float* P) {
  P[CI_VV10] = P[SI(3, 4)];
  P[CI_AA10] = P[SI(6, 7)];
}
AGF_DEV void cov_load(const Scratch& sc, float* P) {
#pragma unroll
  for (int q = 0; q < 12; q++) {
    const float4 v = sq_load(sc, SQ_COV + q);
    P[4 * q] = v.x; P[4 * q + 1] = v.y; P[4 * q + 2] = v.z; P[4 * q + 3] = v.w;
  }
}
AGF_DEV void cov_store(const Scratch& sc, const float* P) {
#pragma unroll
  for (int q = 0; q < 12; q++) sq_store(sc, SQ_COV + q, make_float4(P[4 * q], P[4 * q + 1], P[4 * q + 2], P[4 * q + 3]));
}
// The three components of one sensor through the same filter (fast variants).  Scratch quads q0 .. q0+2:
//   {xm0.x, xm0.y, xm1.x, xm1.y}, {ym0.x, ym0.y, ym1.x, ym1.y} (x and y as packed pairs), {xm0, xm1, ym0, ym1} of z;
// one multiply-add chain per output, the x/y pair on the packed FP32 path.
AGF_DEV V3<float> lpf2_scratch3(const Lpf2Coef& c, const Scratch& sc, int q0, const V3<float>& in) {
  const float4 X = sq_load(sc, q0), Y = sq_load(sc, q0 + 1), Z = sq_load(sc, q0 + 2);
  const float2 inxy = make_float2(in.x, in.y);
  float2 o = f2_mul(inxy, c.b2);
  o = f2_fma(make_float2(X.x, X.y), c.b0, o);
  o = f2_fma(make_float2(X.z, X.w), c.b1, o);
  o = f2_fma(make_float2(Y.x, Y.y), -c.a1, o);
  o = f2_fma(make_float2(Y.z, Y.w), -c.a2, o);
  sq_store(sc, q0, make_float4(X.z, X.w, inxy.x, inxy.y));
  sq_store(sc, q0 + 1, make_float4(Y.z, Y.w, o.x, o.y));
  const float oz = ::fmaf(-c.a2, Z.w, ::fmaf(-c.a1, Z.z, ::fmaf(c.b1, Z.y, ::fmaf(c.b0, Z.x, c.b2 * in.z))));
  sq_store(sc, q0 + 2, make_float4(Z.y, in.z, Z.w, oz));
  return V3<float>(o.x, o.y, oz);
}
// HBM keeps the filter states component-major ([4 * c + k], k = xm0 xm1 ym0 ym1): conversion to / from the scratch layout
AGF_DEV void lpf_scratch_put(const Scratch& sc, int q0, const float* f) {
  sq_store(sc, q0, make_float4(f[0], f[4], f[1], f[5]));
  sq_store(sc, q0 + 1, make_float4(f[2], f[6], f[3], f[7]));
  sq_store(sc, q0 + 2, make_float4(f[8], f[9], f[10], f[11]));
}
AGF_DEV void lpf_scratch_get(const Scratch& sc, int q0, float* f) {
  const float4 X = sq_load(sc, q0), Y = sq_load(sc, q0 + 1), Z = sq_load(sc, q0 + 2);
  f[0] = X.x; f[4] = X.y; f[1] = X.z; f[5] = X.w;
  f[2] = Y.x; f[6] = Y.y; f[3] = Y.z; f[7] = Y.w;
  f[8] = Z.x; f[9] = Z.y; f[10] = Z.z; f[11] = Z.w;
}

// State is read once per work item and may have been written by another CTA of the same launch (balanced
// schedule below): load through L2 only (ld.global.cg), never from a possibly stale L1 line.
template<typename T> AGF_DEV T ldro_(const T* p) {  // read-only for the whole launch: L1-cached, one line serves every warp of the SM
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
template<typename T> AGF_DEV T ldcg_(const T* p) {
#if defined(__CUDA_ARCH__)
  return __ldcg(p);
#else
  return *p;
#endif
}

// --- flat (de)serialisation order; the host get/set kernels use the same tables (agf_types.h) ---
template<typename P, bool PARITY, bool UWB, bool HK>
AGF_DEV void state_load(VState<P, PARITY, UWB, HK>& s, const StateArrays<P>& a, size_t n, size_t i, const Scratch& sc, float mix_kf) {
  typedef typename VecOf<P>::type PV;
  constexpr int VP = VecOf<P>::lanes;
  constexpr bool COMP = !PARITY && sizeof(P) == 4;  // compensated FP32 integration: two more triples of plant scalars
  P rp[NP_PAD];
#pragma unroll
  for (int q = 0; q < (COMP ? NP_PAD : NP_REF) / VP; q++) {
    PV v = ldcg_(&a.sp[size_t(q) * n + i]);
    VecOf<P>::unpack(v, &rp[q * VP]);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) { s.pos[k] = rp[SP_POS + k]; s.vel[k] = rp[SP_VEL + k]; s.w[k] = rp[SP_W + k]; }
  if constexpr (COMP) {
#pragma unroll
    for (int k = 0; k < 3; k++) { s.cpos[k] = rp[SP_CPOS + k]; s.cvel[k] = rp[SP_CVEL + k]; }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) { s.att[k] = rp[SP_ATT + k]; s.ms[k] = rp[SP_MS + k]; }
  if constexpr (UWB) {
#pragma unroll
    for (int k = 0; k < 3; k++) s.rpos[k] = rp[SP_RPOS + k];
  }
  float rf[NF_PAD];
#pragma unroll
  for (int q = 0; q < NF_PAD / 4; q++) {
    if (!HK && q >= NF_CORE / 4) break;
    float4 v = ldcg_(&a.sf[size_t(q) * n + i]);
    rf[4 * q] = v.x; rf[4 * q + 1] = v.y; rf[4 * q + 2] = v.z; rf[4 * q + 3] = v.w;
  }
#pragma unroll
  for (int k = 0; k < 4; k++) { s.cmd[k] = rf[SF_CMD + k]; s.radio_f[k] = rf[SF_RADIO + k]; s.katt[k] = rf[SF_KATT + k]; }
  if constexpr (PARITY) {
#pragma unroll
    for (int k = 0; k < 4; k++) s.dforce[k] = rf[SF_DFORCE + k];
  }
  if constexpr (PARITY) {
#pragma unroll
    for (int k = 0; k < 12; k++) { s.gyro_lp[k] = rf[SF_GYRO_LP + k]; s.acc_lp[k] = rf[SF_ACC_LP + k]; }
  } else {
    lpf_scratch_put(sc, SQ_LPF, &rf[SF_GYRO_LP]);
    lpf_scratch_put(sc, SQ_LPF + 3, &rf[SF_ACC_LP]);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) { s.kpos[k] = rf[SF_KPOS + k]; s.kvel[k] = rf[SF_KVEL + k]; s.kw[k] = rf[SF_KW + k]; s.kcorr[k] = rf[SF_KCORR + k]; }
  s.uwb_range = rf[SF_UWB_RANGE];
  s.logic_range = rf[SF_LOGIC_RANGE];
  if constexpr (HK) {
#pragma unroll
    if constexpr (PARITY) {
#pragma unroll
      for (int k = 0; k < 4; k++) { s.pc_accum[k] = rf[SF_PC_ACCUM + k]; s.pc_corr[k] = rf[SF_PC_CORR + k]; }
#pragma unroll
      for (int k = 0; k < 4; k++) { s.temp_lp[k] = rf[SF_TEMP_LP + k]; s.batt_lp[k] = rf[SF_BATT_LP + k]; }
    } else {
      const int q0 = UWB ? SQ_PC_UWB : SQ_PC_NOUWB;
      sq_store(sc, q0 + SQ_PC_CORR, make_float4(rf[SF_PC_CORR], rf[SF_PC_CORR + 1], rf[SF_PC_CORR + 2], rf[SF_PC_CORR + 3]));
      sq_store(sc, q0 + SQ_PC_INV, make_float4(1.0f / (rf[SF_PC_CORR] * mix_kf), 1.0f / (rf[SF_PC_CORR + 1] * mix_kf),
                                               1.0f / (rf[SF_PC_CORR + 2] * mix_kf), 1.0f / (rf[SF_PC_CORR + 3] * mix_kf)));
      sq_store(sc, q0 + SQ_PC_ACCUM, make_float4(rf[SF_PC_ACCUM], rf[SF_PC_ACCUM + 1], rf[SF_PC_ACCUM + 2], rf[SF_PC_ACCUM + 3]));
    }
    s.batt_vfilt = rf[SF_BATT_VFILT]; s.mon_cmd_lpdt = rf[SF_MON_CMD]; s.mon_loop_lpdt = rf[SF_MON_LOOP];
  }
  uint32_t ru[NU_PAD];
#pragma unroll
  for (int q = 0; q < NU_PAD / 4; q++) {
    if (!HK && q >= NU_CORE / 4) break;
    uint4 v = ldcg_(&a.su[size_t(q) * n + i]);
    ru[4 * q] = v.x; ru[4 * q + 1] = v.y; ru[4 * q + 2] = v.z; ru[4 * q + 3] = v.w;
  }
  s.bits = ru[SU_BITS]; s.cnt = ru[SU_CNT]; s.cycle = ru[SU_CYCLE]; s.kfcnt = ru[SU_KFCNT];
  s.uwb_count = ru[SU_UWB_COUNT]; s.age_radio = ru[SU_AGE_RADIO]; s.age_uwb = ru[SU_AGE_UWB];
  s.uwbw = ru[SU_UWBW];
  if (HK) { s.age_est_reset = ru[SU_AGE_EST_RESET]; s.pc_count = ru[SU_PC_COUNT]; s.age_mon_cmd = ru[SU_AGE_MON_CMD]; s.age_mon_loop = ru[SU_AGE_MON_LOOP]; }
  if (a.sq) {
#pragma unroll
    for (int q = 0; q < AGF_OFFQ; q++) {
      const float4 v = ldcg_(&a.sq[size_t(q) * n + i]);
      if constexpr (PARITY) {
        s.offq[4 * q] = v.x; s.offq[4 * q + 1] = v.y; s.offq[4 * q + 2] = v.z; s.offq[4 * q + 3] = v.w;
      } else {
        sq_store(sc, (UWB ? SQ_QUADS_UWB : SQ_QUADS_NOUWB) + q, v);
      }
    }
  }
  if constexpr (UWB) {
    float full[NC_PAD];
#pragma unroll
    for (int q = 0; q < NC_PAD / 4; q++) {
      float4 v = ldcg_(&a.sc[size_t(q) * n + i]);
      full[4 * q + 0] = v.x; full[4 * q + 1] = v.y; full[4 * q + 2] = v.z; full[4 * q + 3] = v.w;
    }
    if constexpr (PARITY) {
#pragma unroll
      for (int k = 0; k < 81; k++) s.cov[k] = full[k];
    } else {
      float Ps[48];
#pragma unroll
      Ps[47] = 0.0f;
#pragma unroll
      for (int r = 0; r < 9; r++)
#pragma unroll
        for (int c = r; c < 9; c++) Ps[SI(r, c)] = full[9 * r + c];
      cov_fill_mirrors(Ps);
      cov_store(sc, Ps);
    }
  }
}

template<typename P, bool PARITY, bool UWB, bool HK>
AGF_DEV void state_store(const VState<P, PARITY, UWB, HK>& s, const StateArrays<P>& a, size_t n, size_t i, const Scratch& sc, float mix_kf) {
  typedef typename VecOf<P>::type PV;
  constexpr int VP = VecOf<P>::lanes;
  P rp[NP_PAD];
#pragma unroll
  for (int k = 0; k < NP_PAD; k++) rp[k] = P(0);
#pragma unroll
  for (int k = 0; k < 3; k++) { rp[SP_POS + k] = s.pos[k]; rp[SP_VEL + k] = s.vel[k]; rp[SP_W + k] = s.w[k]; }
#pragma unroll
  for (int k = 0; k < 4; k++) { rp[SP_ATT + k] = s.att[k]; rp[SP_MS + k] = s.ms[k]; }
  if constexpr (UWB) {
#pragma unroll
    for (int k = 0; k < 3; k++) rp[SP_RPOS + k] = s.rpos[k];
  }
  constexpr bool COMP = !PARITY && sizeof(P) == 4;
  if constexpr (COMP) {
#pragma unroll
    for (int k = 0; k < 3; k++) { rp[SP_CPOS + k] = s.cpos[k]; rp[SP_CVEL + k] = s.cvel[k]; }
  }
#pragma unroll
  for (int q = 0; q < (COMP ? NP_PAD : NP_REF) / VP; q++) {
    const bool radio_quad = q * VP >= SP_RPOS && q * VP < SP_CVEL;
    if (!UWB && radio_quad) continue;  // the radio's latched position only exists with ranging
    a.sp[size_t(q) * n + i] = VecOf<P>::pack(&rp[q * VP]);
  }
  float rf[NF_PAD];
#pragma unroll
  for (int k = 0; k < NF_PAD; k++) rf[k] = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    rf[SF_CMD + k] = s.cmd[k]; rf[SF_RADIO + k] = s.radio_f[k]; rf[SF_KATT + k] = s.katt[k];
    // fast variants: the commanded force is not carried, it is the mixer's thrust<->speed map inverted (times the
    // calibration correction, below, with housekeeping)
    if constexpr (PARITY) rf[SF_DFORCE + k] = s.dforce[k];
    else rf[SF_DFORCE + k] = mix_kf * s.cmd[k] * s.cmd[k];
  }
  if constexpr (!PARITY && HK) {
    const float4 c4 = sq_load(sc, (UWB ? SQ_PC_UWB : SQ_PC_NOUWB) + SQ_PC_CORR);
    rf[SF_DFORCE] *= c4.x; rf[SF_DFORCE + 1] *= c4.y; rf[SF_DFORCE + 2] *= c4.z; rf[SF_DFORCE + 3] *= c4.w;
  }
  if constexpr (PARITY) {
#pragma unroll
    for (int k = 0; k < 12; k++) { rf[SF_GYRO_LP + k] = s.gyro_lp[k]; rf[SF_ACC_LP + k] = s.acc_lp[k]; }
  } else {
    lpf_scratch_get(sc, SQ_LPF, &rf[SF_GYRO_LP]);
    lpf_scratch_get(sc, SQ_LPF + 3, &rf[SF_ACC_LP]);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) { rf[SF_KPOS + k] = s.kpos[k]; rf[SF_KVEL + k] = s.kvel[k]; rf[SF_KW + k] = s.kw[k]; rf[SF_KCORR + k] = s.kcorr[k]; }
  rf[SF_UWB_RANGE] = s.uwb_range;
  rf[SF_LOGIC_RANGE] = s.logic_range;
  if constexpr (HK) {
#pragma unroll
    if constexpr (PARITY) {
#pragma unroll
      for (int k = 0; k < 4; k++) { rf[SF_PC_ACCUM + k] = s.pc_accum[k]; rf[SF_PC_CORR + k] = s.pc_corr[k]; }
#pragma unroll
      for (int k = 0; k < 4; k++) { rf[SF_TEMP_LP + k] = s.temp_lp[k]; rf[SF_BATT_LP + k] = s.batt_lp[k]; }
    } else {
      const int q0 = UWB ? SQ_PC_UWB : SQ_PC_NOUWB;
      const float4 c4 = sq_load(sc, q0 + SQ_PC_CORR), a4 = sq_load(sc, q0 + SQ_PC_ACCUM);
      rf[SF_PC_CORR] = c4.x; rf[SF_PC_CORR + 1] = c4.y; rf[SF_PC_CORR + 2] = c4.z; rf[SF_PC_CORR + 3] = c4.w;
      rf[SF_PC_ACCUM] = a4.x; rf[SF_PC_ACCUM + 1] = a4.y; rf[SF_PC_ACCUM + 2] = a4.z; rf[SF_PC_ACCUM + 3] = a4.w;
    }
    rf[SF_BATT_VFILT] = s.batt_vfilt; rf[SF_MON_CMD] = s.mon_cmd_lpdt; rf[SF_MON_LOOP] = s.mon_loop_lpdt;
  }
#pragma unroll
  for (int q = 0; q < NF_PAD / 4; q++) {
    if (!HK && q >= NF_CORE / 4) break;
    if (!PARITY && (q == SF_TEMP_LP / 4 || q == SF_BATT_LP / 4)) continue;  // fixed-point filters: stored states stay as they are
    a.sf[size_t(q) * n + i] = make_float4(rf[4 * q], rf[4 * q + 1], rf[4 * q + 2], rf[4 * q + 3]);
  }
  uint32_t ru[NU_PAD];
#pragma unroll
  for (int k = 0; k < NU_PAD; k++) ru[k] = 0;
  ru[SU_BITS] = s.bits; ru[SU_CNT] = s.cnt; ru[SU_CYCLE] = s.cycle; ru[SU_KFCNT] = s.kfcnt;
  ru[SU_UWB_COUNT] = s.uwb_count; ru[SU_AGE_RADIO] = s.age_radio; ru[SU_AGE_UWB] = s.age_uwb;
  ru[SU_UWBW] = s.uwbw;
  if (HK) { ru[SU_AGE_EST_RESET] = s.age_est_reset; ru[SU_PC_COUNT] = s.pc_count; ru[SU_AGE_MON_CMD] = s.age_mon_cmd; ru[SU_AGE_MON_LOOP] = s.age_mon_loop; }
  if (HK) {
#pragma unroll
    for (int q = 0; q < NU_PAD / 4; q++)
      a.su[size_t(q) * n + i] = make_uint4(ru[4 * q], ru[4 * q + 1], ru[4 * q + 2], ru[4 * q + 3]);
  } else {
    // the housekeeping words are not maintained by this variant: keep what is stored
#pragma unroll
    for (int q = 0; q < NU_CORE / 4; q++)
      a.su[size_t(q) * n + i] = make_uint4(ru[4 * q], ru[4 * q + 1], ru[4 * q + 2], ru[4 * q + 3]);
  }
  if (a.sq) {
#pragma unroll
    for (int q = 0; q < AGF_OFFQ; q++) {
      if constexpr (PARITY) {
        a.sq[size_t(q) * n + i] = make_float4(s.offq[4 * q], s.offq[4 * q + 1], s.offq[4 * q + 2], s.offq[4 * q + 3]);
      } else {
        a.sq[size_t(q) * n + i] = sq_load(sc, (UWB ? SQ_QUADS_UWB : SQ_QUADS_NOUWB) + q);
      }
    }
  }
  if constexpr (UWB) {
    float full[NC_PAD];
#pragma unroll
    for (int k = 81; k < NC_PAD; k++) full[k] = 0.0f;
    if constexpr (PARITY) {
#pragma unroll
      for (int k = 0; k < 81; k++) full[k] = s.cov[k];
    } else {
      float Ps[48];
      cov_load(sc, Ps);
#pragma unroll
      for (int r = 0; r < 9; r++)
#pragma unroll
        for (int c = 0; c < 9; c++) full[9 * r + c] = Ps[SI(r, c)];
    }
#pragma unroll
    for (int q = 0; q < NC_PAD / 4; q++)
      a.sc[size_t(q) * n + i] = make_float4(full[4 * q], full[4 * q + 1], full[4 * q + 2], full[4 * q + 3]);
  }
}

// ---------------------------------------------------------------------------------------------
// estimator: KalmanFilter6DOF
// ---------------------------------------------------------------------------------------------
#define AGF_COV(i, j) s.cov[9 * (i) + (j)]

// The estimator members the rarely taken paths touch, packed so they can cross a real call (fast variants)
struct KfCore {
  float kpos[3], kvel[3], kw[3], kcorr[3], katt[4];
  uint32_t bits, kfcnt, cnt;
};
template<typename ST> AGF_DEV void kf_core_get(KfCore& c, const ST& s) {
#pragma unroll
  for (int k = 0; k < 3; k++) { c.kpos[k] = s.kpos[k]; c.kvel[k] = s.kvel[k]; c.kw[k] = s.kw[k]; c.kcorr[k] = s.kcorr[k]; }
#pragma unroll
  for (int k = 0; k < 4; k++) c.katt[k] = s.katt[k];
  c.bits = s.bits; c.kfcnt = s.kfcnt; c.cnt = s.cnt;
}
template<typename ST> AGF_DEV void kf_core_put(ST& s, const KfCore& c) {
#pragma unroll
  for (int k = 0; k < 3; k++) { s.kpos[k] = c.kpos[k]; s.kvel[k] = c.kvel[k]; s.kw[k] = c.kw[k]; s.kcorr[k] = c.kcorr[k]; }
#pragma unroll
  for (int k = 0; k < 4; k++) s.katt[k] = c.katt[k];
  s.bits = c.bits; s.kfcnt = c.kfcnt; s.cnt = c.cnt;
}

// ST is the register-resident VState, or a KfCore in the out-of-line copies of the fast variants (whose
// covariance lives in the shared-memory scratch, not in ST)
template<bool PARITY, bool UWB, typename ST>
AGF_DEV void kf_reset(ST& s, const Scratch& sc) {  // KalmanFilter6DOF.cpp:33-68
  s.kfcnt = (s.kfcnt & 0xFFFF0000u) | min((s.kfcnt & 0xFFFFu) + 1u, 0xFFFFu);  // saturating, see kf_range_rejected
  s.bits &= ~(B_IMU_INIT | B_UWB_INIT);
  s.bits |= B_KF_RESET_SEEN;
#pragma unroll
  for (int k = 0; k < 3; k++) { s.kpos[k] = 0; s.kvel[k] = 0; s.kw[k] = 0; s.kcorr[k] = 0; }
  s.katt[0] = 1; s.katt[1] = 0; s.katt[2] = 0; s.katt[3] = 0;
  if constexpr (UWB) {
    const float sp = 3.0f, sv = 3.0f;
    const float sperp = 10.0f * 3.14159274f / 180.0f, sabout = 30.0f * 3.14159274f / 180.0f;
    if constexpr (PARITY) {
#pragma unroll
      for (int k = 0; k < 81; k++) s.cov[k] = 0;
#pragma unroll
      for (int k = 0; k < 3; k++) { AGF_COV(k, k) = sp * sp; AGF_COV(3 + k, 3 + k) = sv * sv; }
      AGF_COV(6, 6) = sperp * sperp;
      AGF_COV(7, 7) = sperp * sperp;
      AGF_COV(8, 8) = sabout * sabout;
    } else {
      float Ps[48];
#pragma unroll
      for (int k = 0; k < 48; k++) Ps[k] = 0.0f;
#pragma unroll
      for (int k = 0; k < 3; k++) { Ps[SI(k, k)] = sp * sp; Ps[SI(3 + k, 3 + k)] = sv * sv; }
      Ps[SI(6, 6)] = sperp * sperp;
      Ps[SI(7, 7)] = sperp * sperp;
      Ps[SI(8, 8)] = sabout * sabout;
      cov_store(sc, Ps);
    }
  }
}

// gravity alignment shared by the first Predict and the complementary filter (:77-108, :123-145)
template<bool PARITY>
AGF_DEV void gravity_axis_angle(const Q4<float>& att, const V3<float>& acc, V3<float>& ax, float& ang) {
  const V3<float> expAcc = qrot(qinv(att), V3<float>(0, 0, 1));
  const float n = norm(acc);  // GetUnitVector, Vec3.hpp:126-129
  const V3<float> accUnit = vdiv<PARITY>(acc, n);
  const float cosErr = dot(expAcc, accUnit);
  V3<float> rotAx = cross(accUnit, expAcc);
  const float rn = norm(rotAx);
  if (rn > 1e-6f) {
    rotAx = vdiv<PARITY>(rotAx, rn);
  } else {
    rotAx = V3<float>(1, 0, 0);
  }
  ax = rotAx;
  ang = acos_guarded<PARITY>(cosErr);
}

// The two start-up branches of Predict: first call = reset + gravity alignment (:71-108); until the first UWB
// range (always, without UWB) = complementary filter (:114-147).
template<bool PARITY, bool UWB, typename ST>
AGF_DEV void kf_predict_startup(ST& s, const Scratch& sc, const V3<float>& gyro, const V3<float>& acc, float dt) {
  if (!(s.bits & B_IMU_INIT)) {  // :71-108
    kf_reset<PARITY, UWB>(s, sc);
    s.bits |= B_IMU_INIT;
    Q4<float> att(s.katt[0], s.katt[1], s.katt[2], s.katt[3]);
    V3<float> ax;
    float ang;
    gravity_axis_angle<PARITY>(att, acc, ax, ang);
    att = qmul(att, q_from_axis_angle<PARITY>(ax, ang));
    s.katt[0] = att.w; s.katt[1] = att.x; s.katt[2] = att.y; s.katt[3] = att.z;
    return;
  }
  Q4<float> att(s.katt[0], s.katt[1], s.katt[2], s.katt[3]);
  s.kw[0] = gyro.x; s.kw[1] = gyro.y; s.kw[2] = gyro.z;
  att = q_apply_rotvec<PARITY>(att, gyro * dt);
  V3<float> ax;
  float ang;
  gravity_axis_angle<PARITY>(att, acc, ax, ang);
  const float corr = (dt / 4.0f) * ang;
  att = qmul(att, q_from_axis_angle<PARITY>(ax, corr));
  s.katt[0] = att.w; s.katt[1] = att.x; s.katt[2] = att.y; s.katt[3] = att.z;
}
static AGF_COLD void kf_predict_startup_cold(KfCore* c, Scratch sc, V3<float> gyro, V3<float> acc, float dt) {
  kf_predict_startup<false, true>(*c, sc, gyro, acc, dt);
}
// a rejected range: counters, reset after 5 in a row (:272-284)
template<bool PARITY, bool UWB, typename ST>
AGF_DEV void kf_range_rejected(ST& s, const Scratch& sc) {
  // The reference's counters are unsigned ints and Reset() does not clear the in-a-row count, so a filter that keeps
  // rejecting resets on EVERY further rejection (a panicked vehicle lying far from its anchors does this for thousands
  // of ranges).  The packed fields saturate instead of wrapping: the decision `seq >= 5` stays the reference's for any run
  // length (found by the full-size C2 parity test: a wrap at 256 silently skipped four resets).
  const uint32_t rej = min((s.kfcnt >> 16) + 1u, 0xFFFFu);
  s.kfcnt = (s.kfcnt & 0xFFFFu) | (rej << 16);
  const uint32_t seq = min((s.cnt & 0xFFu) + 1u, 0xFFu);
  s.cnt = (s.cnt & ~0xFFu) | seq;
  if (seq >= 5u) kf_reset<PARITY, UWB>(s, sc);
}
static AGF_COLD void kf_range_rejected_cold(KfCore* c, Scratch sc) { kf_range_rejected<false, true>(*c, sc); }

template<bool PARITY, typename P, bool UWB, bool HK>
AGF_DEV void kf_predict(VState<P, PARITY, UWB, HK>& s, const Scratch& sc, const V3<float>& gyro, const V3<float>& acc, float dt) {
  if constexpr (!PARITY && UWB) {  // start-up is rare once ranging runs: out of line
    if (AGF_UNLIKELY((s.bits & (B_IMU_INIT | B_UWB_INIT)) != (B_IMU_INIT | B_UWB_INIT))) {
      KfCore c;
      kf_core_get(c, s);
      kf_predict_startup_cold(&c, sc, gyro, acc, dt);
      kf_core_put(s, c);
      return;
    }
  } else {
    if (!UWB || (s.bits & (B_IMU_INIT | B_UWB_INIT)) != (B_IMU_INIT | B_UWB_INIT)) {
      kf_predict_startup<PARITY, UWB>(s, sc, gyro, acc, dt);
      return;
    }
  }
  Q4<float> att(s.katt[0], s.katt[1], s.katt[2], s.katt[3]);
  if constexpr (UWB) {  // :149-241
    const V3<float> p0(s.kpos[0], s.kpos[1], s.kpos[2]), v0(s.kvel[0], s.kvel[1], s.kvel[2]);
    float Rm[9];
    qmatrix(att, Rm);
    const V3<float> a_w = V3<float>(Rm[0] * acc.x + Rm[1] * acc.y + Rm[2] * acc.z, Rm[3] * acc.x + Rm[4] * acc.y + Rm[5] * acc.z,
                                    Rm[6] * acc.x + Rm[7] * acc.y + Rm[8] * acc.z) + V3<float>(0, 0, -9.81f);
    const V3<float> p1 = p0 + v0 * dt, v1 = v0 + a_w * dt;
    s.kpos[0] = p1.x; s.kpos[1] = p1.y; s.kpos[2] = p1.z;
    s.kvel[0] = v1.x; s.kvel[1] = v1.y; s.kvel[2] = v1.z;
    const Q4<float> att1 = q_apply_rotvec<PARITY>(att, gyro * dt);
    s.katt[0] = att1.w; s.katt[1] = att1.x; s.katt[2] = att1.y; s.katt[3] = att1.z;
    s.kw[0] = gyro.x; s.kw[1] = gyro.y; s.kw[2] = gyro.z;

    // vel <- att block of f, :184-209   A[r][c]
    float A[3][3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      if constexpr (PARITY) {
        A[r][0] = dt * (+acc.y * Rm[3 * r + 2] - acc.z * Rm[3 * r + 1]);
        A[r][1] = dt * (-acc.x * Rm[3 * r + 2] + acc.z * Rm[3 * r + 0]);
        A[r][2] = dt * (+acc.x * Rm[3 * r + 1] - acc.y * Rm[3 * r + 0]);
      } else {  // dt folded into the measured acceleration once
        const float ax = dt * acc.x, ay = dt * acc.y, az = dt * acc.z;
        A[r][0] = ::fmaf(ay, Rm[3 * r + 2], -(az * Rm[3 * r + 1]));
        A[r][1] = ::fmaf(az, Rm[3 * r + 0], -(ax * Rm[3 * r + 2]));
        A[r][2] = ::fmaf(ax, Rm[3 * r + 1], -(ay * Rm[3 * r + 0]));
      }
    }
    // att <- att block, :212-228   S[r][c]
    const float gx = dt * gyro.x + s.kcorr[0] / 2.0f;
    const float gy = dt * gyro.y + s.kcorr[1] / 2.0f;
    const float gz = dt * gyro.z + s.kcorr[2] / 2.0f;
    const float S[3][3] = {{1.0f, +gz, -gy}, {-gz, 1.0f, +gx}, {+gy, -gx, 1.0f}};
    s.kcorr[0] = 0; s.kcorr[1] = 0; s.kcorr[2] = 0;
    const float qa = 5.0f * 5.0f * dt * dt, qg = 0.1f * 0.1f * dt * dt;  // process noise :234-239
    if constexpr (!PARITY) {
      // Fast variants: P is kept exactly symmetric (packed, 47 values, see SI) and propagated block-wise,
      // f = [[I, dt I, 0], [0, I, A], [0, 0, S]] = F2 * F1 with F1 the position/velocity coupling.  Same algebra as the
      // dense f P f^T of KalmanFilter6DOF.cpp:232 up to rounding, in about a third of its multiply-adds: rows 0 and 1 of
      // every 3x3 block ride in the two lanes of the packed FP32 instructions (FFMA2: pair x broadcast scalar + pair),
      // row 2 is scalar.  Right-multiplications (X M^T) take the pair from X and the scalar from M; left-multiplications
      // (M X) take the pair from a column of M and the scalar from X, so nothing is ever transposed or shuffled.
      float Pm[48];
      cov_load(sc, Pm);
#define PR(k) make_float2(Pm[2 * (k)], Pm[2 * (k) + 1])
#define PRSET(k, v) { const float2 v_ = (v); Pm[2 * (k)] = v_.x; Pm[2 * (k) + 1] = v_.y; }
#define PVS(r, c) Pm[SI_rect(CP_PV, CI_PV2, (r), (c))]
#define PAS(r, c) Pm[SI_rect(CP_PA, CI_PA2, (r), (c))]
#define VAS(r, c) Pm[SI_rect(CP_VA, CI_VA2, (r), (c))]
#define VVS(r, c) Pm[(r) <= (c) ? SI_sym(CP_VV, CI_VV22, (r), (c)) : SI_sym(CP_VV, CI_VV22, (c), (r))]
#define AAS(r, c) Pm[(r) <= (c) ? SI_sym(CP_AA, CI_AA22, (r), (c)) : SI_sym(CP_AA, CI_AA22, (c), (r))]
#define PPS(r, c) Pm[SI((r), (c))]
      float2 Acol[3], Scol[3];  // columns of A and S, rows 0 and 1
#pragma unroll
      for (int k = 0; k < 3; k++) { Acol[k] = make_float2(A[0][k], A[1][k]); Scol[k] = make_float2(S[0][k], S[1][k]); }
      // F1: Ppv += dt Pvv ; Ppp += dt (Ppv_new + Ppv_old^T) ; Ppa += dt Pva
      float2 Vn[3];
      float v2n[3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        Vn[c] = f2_fma(PR(CP_VV + c), dt, PR(CP_PV + c));
        v2n[c] = ::fmaf(dt, VVS(2, c), PVS(2, c));
      }
#define VN(r, c) ((r) == 0 ? Vn[c].x : ((r) == 1 ? Vn[c].y : v2n[c]))
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = i; j < 3; j++) PPS(i, j) = ::fmaf(dt, VN(i, j) + PVS(j, i), PPS(i, j));
#pragma unroll
      for (int c = 0; c < 3; c++) {
        PRSET(CP_PA + c, f2_fma(PR(CP_VA + c), dt, PR(CP_PA + c)));
        PAS(2, c) = ::fmaf(dt, VAS(2, c), PAS(2, c));
      }
      // F2: T = Pva + A Paa
      float2 Tp[3];
      float t2[3];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        Tp[j] = f2_fma(Acol[2], AAS(2, j), f2_fma(Acol[1], AAS(1, j), f2_fma(Acol[0], AAS(0, j), PR(CP_VA + j))));
        t2[j] = ::fmaf(A[2][2], AAS(2, j), ::fmaf(A[2][1], AAS(1, j), ::fmaf(A[2][0], AAS(0, j), VAS(2, j))));
      }
      // Pvv += A Pva^T + T A^T   (rows 0, 1 for every column; (2,2) scalar; (2,0), (2,1) are the column-2 pair)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        float2 acc = PR(CP_VV + j);
#pragma unroll
        for (int k = 0; k < 3; k++) acc = f2_fma(Acol[k], VAS(j, k), acc);
#pragma unroll
        for (int k = 0; k < 3; k++) acc = f2_fma(Tp[k], A[j][k], acc);
        PRSET(CP_VV + j, acc);
      }
      {
        float acc = Pm[CI_VV22];
#pragma unroll
        for (int k = 0; k < 3; k++) acc = ::fmaf(A[2][k], VAS(2, k), acc);
#pragma unroll
        for (int k = 0; k < 3; k++) acc = ::fmaf(t2[k], A[2][k], acc);
        Pm[CI_VV22] = acc + qa;
      }
      Pm[CI_VV10] = Pm[SI(3, 4)];  // keep P exactly symmetric
      Pm[SI(3, 3)] += qa;
      Pm[SI(4, 4)] += qa;
      // Ppv = Ppv_new + Ppa A^T
#pragma unroll
      for (int j = 0; j < 3; j++) {
        float2 acc = Vn[j];
        float a2 = v2n[j];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          acc = f2_fma(PR(CP_PA + k), A[j][k], acc);
          a2 = ::fmaf(PAS(2, k), A[j][k], a2);
        }
        PRSET(CP_PV + j, acc);
        PVS(2, j) = a2;
      }
#undef VN
      // Ppa = Ppa S^T ; Pva = T S^T ; W = Paa S^T   (S has a unit diagonal: two terms onto the diagonal one)
      float2 An[3], Wp[3];
      float a2n[3], w2[3];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        An[j] = f2_fma(PR(CP_PA + j2), S[j][j2], f2_fma(PR(CP_PA + j1), S[j][j1], PR(CP_PA + j)));
        a2n[j] = ::fmaf(PAS(2, j2), S[j][j2], ::fmaf(PAS(2, j1), S[j][j1], PAS(2, j)));
        Wp[j] = f2_fma(PR(CP_AA + j2), S[j][j2], f2_fma(PR(CP_AA + j1), S[j][j1], PR(CP_AA + j)));
        w2[j] = ::fmaf(AAS(2, j2), S[j][j2], ::fmaf(AAS(2, j1), S[j][j1], AAS(2, j)));
      }
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        PRSET(CP_PA + j, An[j]);
        PAS(2, j) = a2n[j];
        PRSET(CP_VA + j, f2_fma(Tp[j2], S[j][j2], f2_fma(Tp[j1], S[j][j1], Tp[j])));
        VAS(2, j) = ::fmaf(t2[j2], S[j][j2], ::fmaf(t2[j1], S[j][j1], t2[j]));
      }
      // Paa = S W   (rows 0, 1: pair from the columns of S, scalar from W; (2,2) scalar)
#define WS(r, c) ((r) == 0 ? Wp[c].x : ((r) == 1 ? Wp[c].y : w2[c]))
#pragma unroll
      for (int j = 0; j < 3; j++)
        PRSET(CP_AA + j, f2_fma(Scol[2], WS(2, j), f2_fma(Scol[1], WS(1, j), f2_mul(Scol[0], WS(0, j)))));
      Pm[CI_AA22] = ::fmaf(S[2][1], WS(1, 2), ::fmaf(S[2][0], WS(0, 2), w2[2])) + qg;
#undef WS
      Pm[CI_AA10] = Pm[SI(6, 7)];
      Pm[SI(6, 6)] += qg;
      Pm[SI(7, 7)] += qg;
#undef PR
#undef PRSET
#undef PVS
#undef PAS
#undef VAS
#undef VVS
#undef AAS
#undef PPS
      cov_store(sc, Pm);
    } else {

    // FP = f * P, in place, row blocks in an order that only reads not-yet-overwritten rows
#pragma unroll
    for (int j = 0; j < 9; j++) {
#pragma unroll
      for (int r = 0; r < 3; r++) AGF_COV(r, j) = AGF_COV(r, j) + dt * AGF_COV(3 + r, j);
      const float p6 = AGF_COV(6, j), p7 = AGF_COV(7, j), p8 = AGF_COV(8, j);
#pragma unroll
      for (int r = 0; r < 3; r++)
        AGF_COV(3 + r, j) = ((AGF_COV(3 + r, j) + A[r][0] * p6) + A[r][1] * p7) + A[r][2] * p8;
#pragma unroll
      for (int r = 0; r < 3; r++) AGF_COV(6 + r, j) = (S[r][0] * p6 + S[r][1] * p7) + S[r][2] * p8;
    }
    // P' = FP * f^T, in place, column blocks likewise
#pragma unroll
    for (int i = 0; i < 9; i++) {
#pragma unroll
      for (int r = 0; r < 3; r++) AGF_COV(i, r) = AGF_COV(i, r) + AGF_COV(i, 3 + r) * dt;
      const float p6 = AGF_COV(i, 6), p7 = AGF_COV(i, 7), p8 = AGF_COV(i, 8);
#pragma unroll
      for (int r = 0; r < 3; r++)
        AGF_COV(i, 3 + r) = ((AGF_COV(i, 3 + r) + p6 * A[r][0]) + p7 * A[r][1]) + p8 * A[r][2];
#pragma unroll
      for (int r = 0; r < 3; r++) AGF_COV(i, 6 + r) = (p6 * S[r][0] + p7 * S[r][1]) + p8 * S[r][2];
    }
#pragma unroll
    for (int k = 0; k < 3; k++) { AGF_COV(3 + k, 3 + k) += qa; AGF_COV(6 + k, 6 + k) += qg; }
    }
  }
}

template<bool PARITY, typename P, bool UWB, bool HK>
AGF_DEV void kf_update_range(VState<P, PARITY, UWB, HK>& s, const Scratch& sc, const V3<float>& target, float range) {  // :243-301
  if constexpr (UWB) {
  if (!(s.bits & B_IMU_INIT)) return;
  if (!(range == range)) return;
  s.bits |= B_UWB_INIT;
  const V3<float> d = V3<float>(s.kpos[0], s.kpos[1], s.kpos[2]) - target;
  const float expRange = norm(d);
  const V3<float> H = vdiv<PARITY>(d, expRange);
  const float innov = range - expRange;
  if constexpr (!PARITY) {
    // fast variants: symmetric P, P <- P - (P H^T)(P H^T)^T / S  (== (I - L H) P for symmetric P)
    float Pm[48];
    cov_load(sc, Pm);
#define PS(i, j) Pm[SI((i), (j))]
    float PHt[9];
#pragma unroll
    for (int i = 0; i < 9; i++) PHt[i] = PS(i, 0) * H.x + PS(i, 1) * H.y + PS(i, 2) * H.z;
    const float innovCov = (H.x * PHt[0] + H.y * PHt[1] + H.z * PHt[2]) + 0.14f * 0.14f;
    const float inv = fdiv<false>(1.0f, innovCov);
    if (AGF_UNLIKELY(innov * innov * inv > 3.0f * 3.0f)) {
      KfCore c;
      kf_core_get(c, s);
      kf_range_rejected_cold(&c, sc);
      kf_core_put(s, c);
      return;
    }
    s.cnt &= ~0xFFu;
    float L[9];
#pragma unroll
    for (int i = 0; i < 9; i++) L[i] = PHt[i] * inv;
#pragma unroll
    for (int k = 0; k < 3; k++) { s.kpos[k] += L[k] * innov; s.kvel[k] += L[3 + k] * innov; s.kcorr[k] = L[6 + k] * innov; }
    Q4<float> att(s.katt[0], s.katt[1], s.katt[2], s.katt[3]);
    att = q_apply_rotvec<PARITY>(att, V3<float>(s.kcorr[0], s.kcorr[1], s.kcorr[2]));
    s.katt[0] = att.w; s.katt[1] = att.x; s.katt[2] = att.y; s.katt[3] = att.z;
    // rows 0 and 1 of every block on the packed FP32 path (pair of gains x broadcast PHt + pair), row 2 and Ppp scalar
    const float2 nLp = make_float2(-L[0], -L[1]), nLv = make_float2(-L[3], -L[4]), nLa = make_float2(-L[6], -L[7]);
#define PR(k) make_float2(Pm[2 * (k)], Pm[2 * (k) + 1])
#define PRSET(k, v) { const float2 v_ = (v); Pm[2 * (k)] = v_.x; Pm[2 * (k) + 1] = v_.y; }
#pragma unroll
    for (int c = 0; c < 3; c++) {
      PRSET(CP_PV + c, f2_fma(nLp, PHt[3 + c], PR(CP_PV + c)));
      PRSET(CP_PA + c, f2_fma(nLp, PHt[6 + c], PR(CP_PA + c)));
      PRSET(CP_VA + c, f2_fma(nLv, PHt[6 + c], PR(CP_VA + c)));
      PRSET(CP_VV + c, f2_fma(nLv, PHt[3 + c], PR(CP_VV + c)));
      PRSET(CP_AA + c, f2_fma(nLa, PHt[6 + c], PR(CP_AA + c)));
      Pm[CI_PV2 + c] = ::fmaf(-L[2], PHt[3 + c], Pm[CI_PV2 + c]);
      Pm[CI_PA2 + c] = ::fmaf(-L[2], PHt[6 + c], Pm[CI_PA2 + c]);
      Pm[CI_VA2 + c] = ::fmaf(-L[5], PHt[6 + c], Pm[CI_VA2 + c]);
    }
#undef PR
#undef PRSET
    Pm[CI_VV22] = ::fmaf(-L[5], PHt[5], Pm[CI_VV22]);
    Pm[CI_AA22] = ::fmaf(-L[8], PHt[8], Pm[CI_AA22]);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = i; j < 3; j++) PS(i, j) = ::fmaf(-L[i], PHt[j], PS(i, j));
    cov_fill_mirrors(Pm);  // the two lanes of (1,0) / (0,1) may round differently: keep P exactly symmetric
#undef PS
    cov_store(sc, Pm);
  } else {
  float PHt[9];
#pragma unroll
  for (int i = 0; i < 9; i++) PHt[i] = (AGF_COV(i, 0) * H.x + AGF_COV(i, 1) * H.y) + AGF_COV(i, 2) * H.z;
  const float hph = (H.x * PHt[0] + H.y * PHt[1]) + H.z * PHt[2];
  const float innovCov = hph + 0.14f * 0.14f;
  const float inv = 1 / innovCov;
  float L[9];
#pragma unroll
  for (int i = 0; i < 9; i++) L[i] = PHt[i] * inv;
  const float d2 = innov * innov / innovCov;
  if (d2 > 3.0f * 3.0f) {
    kf_range_rejected<PARITY, UWB>(s, sc);
    return;
  }
  s.cnt &= ~0xFFu;
  float dx[9];
#pragma unroll
  for (int i = 0; i < 9; i++) dx[i] = L[i] * innov;
#pragma unroll
  for (int k = 0; k < 3; k++) { s.kpos[k] = s.kpos[k] + dx[k]; s.kvel[k] = s.kvel[k] + dx[3 + k]; s.kcorr[k] = dx[6 + k]; }
  Q4<float> att(s.katt[0], s.katt[1], s.katt[2], s.katt[3]);
  att = q_apply_rotvec<PARITY>(att, V3<float>(dx[6], dx[7], dx[8]));
  s.katt[0] = att.w; s.katt[1] = att.x; s.katt[2] = att.y; s.katt[3] = att.z;
  // P <- (I - L H) P ; rows 3..8 first (they read the old rows 0..2), then rows 0..2
  const float h[3] = {H.x, H.y, H.z};
#pragma unroll
  for (int j = 0; j < 9; j++) {
    const float p0 = AGF_COV(0, j), p1 = AGF_COV(1, j), p2 = AGF_COV(2, j);
#pragma unroll
    for (int i = 3; i < 9; i++)
      AGF_COV(i, j) = (((0.0f - L[i] * h[0]) * p0 + (0.0f - L[i] * h[1]) * p1) + (0.0f - L[i] * h[2]) * p2) + AGF_COV(i, j);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float a0 = (i == 0 ? 1.0f : 0.0f) - L[i] * h[0];
      const float a1 = (i == 1 ? 1.0f : 0.0f) - L[i] * h[1];
      const float a2 = (i == 2 ? 1.0f : 0.0f) - L[i] * h[2];
      AGF_COV(i, j) = (a0 * p0 + a1 * p1) + a2 * p2;
    }
  }
  // MakeCovarianceSymmetric :303-309
#pragma unroll
  for (int i = 0; i < 9; i++)
#pragma unroll
    for (int j = i + 1; j < 9; j++) AGF_COV(i, j) = AGF_COV(j, i);
  }
  }
}

// ---------------------------------------------------------------------------------------------
// controllers and mixer
// ---------------------------------------------------------------------------------------------
// QuadcopterAngularVelocityController::GetDesiredTorques (:25-37); inertia = diag(ixx, ixx, izz) as
// QuadcopterConstants builds it (the zero off-diagonal products are exact and dropped)
template<bool PARITY>
AGF_DEV V3<float> ctl_torques(const LogicParams& k, const V3<float>& des, const V3<float>& est) {
  const V3<float> err = des - est;
  const V3<float> acc = PARITY ? V3<float>(err.x / k.tc_w_xy, err.y / k.tc_w_xy, err.z / k.tc_w_z)
                               : V3<float>(err.x * k.inv_tc_w_xy, err.y * k.inv_tc_w_xy, err.z * k.inv_tc_w_z);
  const V3<float> Iw(k.ixx * est.x, k.ixx * est.y, k.izz * est.z);
  const V3<float> nonlin = cross(est, Iw);
  return V3<float>(k.ixx * acc.x, k.ixx * acc.y, k.izz * acc.z) + nonlin;
}

// QuadcopterAttitudeController::GetDesiredAngularVelocity (:35-68)
template<bool PARITY>
AGF_DEV V3<float> ctl_att_core(float tc_att_xy, float tc_att_z, float k3_fast, float k12_fast, const Q4<float>& desAtt, const Q4<float>& estAtt) {
  const Q4<float> err = qmul(qinv(desAtt), estAtt);
  const V3<float> rotVec = q_to_rotvec<PARITY>(err);
  V3<float> e3b, redAx;
  float c;
  if (PARITY) {
    e3b = qrot(qinv(err), V3<float>(0, 0, 1));
    redAx = cross(e3b, V3<float>(0, 0, 1));
    c = dot(e3b, V3<float>(0, 0, 1));
  } else {  // third column of R(err^-1), cross/dot with e3 written out
    const Q4<float> qi = qinv(err);
    e3b = V3<float>(2 * qi.x * qi.z + 2 * qi.w * qi.y, 2 * qi.y * qi.z - 2 * qi.w * qi.x, qi.w * qi.w - qi.x * qi.x - qi.y * qi.y + qi.z * qi.z);
    redAx = V3<float>(e3b.y, -e3b.x, 0.0f);
    c = e3b.z;
  }
  float redAn;
  if constexpr (PARITY) {
    if (c >= 1.0f) {
      redAn = 0;
    } else if (c <= -1.0f) {
      redAn = 3.14159274f;
    } else {
      redAn = Mf<PARITY>::acos(c);
    }
  } else {
    redAn = Mf<PARITY>::acos(::fminf(::fmaxf(c, -1.0f), 1.0f));  // acosf(+-1) = 0 / pi exactly
  }
  const float n = norm(redAx);
  if (n < 1e-12f) {
    redAx = V3<float>(0, 0, 0);
  } else {
    redAx = vdiv<PARITY>(redAx, n);
  }
  const float k3 = PARITY ? 1.0f / tc_att_z : k3_fast, k12 = PARITY ? 1.0f / tc_att_xy : k12_fast;
  return (-k3) * rotVec - ((k12 - k3) * redAn) * redAx;
}

template<bool PARITY>
AGF_DEV V3<float> ctl_att(const LogicParams& k, const Q4<float>& desAtt, const Q4<float>& estAtt) {
  return ctl_att_core<PARITY>(k.tc_att_xy, k.tc_att_z, k.k3_att, k.k12_att, desAtt, estAtt);
}

// QuadcopterMixer::GetMotorForces + PropellerSpeedsFromThrust (:63-99)
template<typename P, bool PARITY, bool UWB, bool HK>
AGF_DEV void ctl_mix(VState<P, PARITY, UWB, HK>& s, const Scratch& sc, const LogicParams& k, float totF, const V3<float>& t) {
  const float desF = totF > k.max_cmd_total ? k.max_cmd_total : totF;
  float f[4];
  if (PARITY) {
    f[0] = (-t.x / k.mix_d - t.y / k.mix_d - t.z / k.mix_kt + desF) / 4.0f;
    f[1] = (-t.x / k.mix_d + t.y / k.mix_d + t.z / k.mix_kt + desF) / 4.0f;
    f[2] = (+t.x / k.mix_d + t.y / k.mix_d - t.z / k.mix_kt + desF) / 4.0f;
    f[3] = (+t.x / k.mix_d - t.y / k.mix_d + t.z / k.mix_kt + desF) / 4.0f;
  } else {
    const float tx = t.x * k.inv_mix_d, ty = t.y * k.inv_mix_d, tz = t.z * k.inv_mix_kt;
    f[0] = (-tx - ty - tz + desF) * 0.25f;
    f[1] = (-tx + ty + tz + desF) * 0.25f;
    f[2] = (+tx + ty - tz + desF) * 0.25f;
    f[3] = (+tx - ty + tz + desF) * 0.25f;
  }
  float inv_ck[4] = {k.inv_mix_kf, k.inv_mix_kf, k.inv_mix_kf, k.inv_mix_kf};  // 1 / (calibration correction * kF), fast variants
  if constexpr (!PARITY && HK) {
    const float4 v = sq_load(sc, (UWB ? SQ_PC_UWB : SQ_PC_NOUWB) + SQ_PC_INV);
    inv_ck[0] = v.x; inv_ck[1] = v.y; inv_ck[2] = v.z; inv_ck[3] = v.w;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if constexpr (PARITY) {
      if (f[i] < k.min_thrust) {
        f[i] = k.min_thrust;
      } else if (f[i] > k.max_thrust) {
        f[i] = k.max_thrust;
      }
    } else {
      f[i] = ::fminf(::fmaxf(f[i], k.min_thrust), k.max_thrust);
    }
    if constexpr (PARITY) s.dforce[i] = f[i];
    if constexpr (PARITY) {
      const float corr = HK ? s.pc_corr[i] : 1.0f;
      s.cmd[i] = f[i] <= 0 ? 0.0f : ::sqrtf(fdiv<PARITY>(f[i], corr * k.mix_kf));
    } else {
      s.cmd[i] = ::sqrtf(::fmaxf(f[i], 0.0f) * inv_ck[i]);  // sqrt(0) = 0: no branch for f <= 0
    }
  }
}

// thrust direction -> attitude (QuadcopterLogic.cpp:423-445)
template<bool PARITY>
AGF_DEV Q4<float> att_from_thrust_dir(const V3<float>& dir) {
  if constexpr (PARITY) {
    const V3<float> e3(0, 0, 1);
    const float cosAngle = dot(dir, e3);
    const float angle = acos_guarded<PARITY>(cosAngle);
    const V3<float> rotAx = cross(e3, dir);
    const float n = norm(rotAx);
    Q4<float> out(1, 0, 0, 0);
    if (!(n < 1e-6f)) {
      Q4<float> d;
      if (q_from_rotvec<PARITY>(rotAx * fdiv<PARITY>(angle, n), d)) out = d;
    }
    return out;
  } else {
    // Same rotation in closed form: axis = e3 x dir = (-dy, dx, 0)/sin(a), cos(a) = dz, so with c = cos(a/2) =
    // sqrt((1 + dz)/2):  q = (c, -dy/(2c), dx/(2c), 0).  No acos, no sin/cos; identity where the reference's
    // guards (|e3 x dir| < 1e-6, angle < MIN_ANGLE) return it.
    const float n2 = dir.x * dir.x + dir.y * dir.y;
    const float c2 = 0.5f * (1.0f + dir.z);
    const bool ident = (n2 < 1e-12f) || !(c2 > 0.0f);
    const float ic = rsqrtf_(ident ? 1.0f : c2);
    const float k = 0.5f * ic;
    return ident ? Q4<float>(1, 0, 0, 0) : Q4<float>(c2 * ic, -dir.y * k, dir.x * k, 0.0f);
  }
}

// ---------------------------------------------------------------------------------------------
// Offboard loop: Offboard::QuadcopterController::Run (Offboard/QuadcopterController.cpp:11-74),
// then RadioMessageDecoded::CreateRatesCommand + the decode on the vehicle side (RadioTypes.hpp:73-116,158-171,
// 210-219).  Returns the four floats the onboard logic will read: thrust [m/s^2], body rates [rad/s].
// ---------------------------------------------------------------------------------------------
AGF_DEV float radio_quantise(float v, float limit) {
  int out;
  if ((v > -limit) && (v < limit)) {
    out = int(v * 32768 / limit + 0.5f) + 32768;
  } else if (v > -limit) {
    out = 65536 - 1;
  } else {
    out = 0;  // min value, and NaN
  }
  out &= 0xFFFF;  // two bytes on the wire
  return limit * (out - 32768) / float(32768);
}
// attitude whose z axis is `dir` (shortest rotation from e3), then yawed: QuadcopterController.cpp:46-71 / :107-127
template<bool PARITY>
AGF_DEV Q4<float> offboard_att_from_dir(const V3<float>& dir, float yaw) {
  const V3<float> e3(0, 0, 1);
  const float cosAngle = dot(dir, e3);
  float angle;
  if (cosAngle >= (1 - 1e-12f)) {
    angle = 0;
  } else if (cosAngle <= -(1 - 1e-12f)) {
    angle = 3.14159274f;
  } else {
    angle = Mf<PARITY>::acos(cosAngle);
  }
  const V3<float> rotAx = cross(e3, dir);
  const float n = norm(rotAx);
  Q4<float> att(1, 0, 0, 0);
  if (!(n < 1e-6f)) {
    Q4<float> d;
    if (q_from_rotvec<PARITY>(rotAx * fdiv<PARITY>(angle, n), d)) att = d;
  }
  Q4<float> yawq(1, 0, 0, 0);
  {
    Q4<float> d;
    if (q_from_rotvec<PARITY>(V3<float>(0, 0, yaw), d)) yawq = d;
  }
  return qmul(att, yawq);
}
AGF_DEV V3<float> v3f(const double v[3]) { return V3<float>(float(v[0]), float(v[1]), float(v[2])); }
// GetDesAcceleration (QuadcopterPositionController.hpp:22-27)
AGF_DEV V3<float> offboard_des_acc(const OffboardParams& c, const V3<float>& estPos, const V3<float>& estVel, const V3<float>& desPos,
                                   const V3<float>& desVel, const V3<float>& desAcc) {
  const V3<float> dv = desVel - estVel;
  return (((desPos - estPos) * c.nat_freq) * c.nat_freq + ((V3<float>(2 * dv.x, 2 * dv.y, 2 * dv.z) * c.nat_freq) * c.damping)) + desAcc;
}
AGF_DEV float4 offboard_rates_packet(double thrust, const V3<float>& w) {
  return make_float4(radio_quantise(float(thrust), 35.0f), radio_quantise(w.x, 35.0f), radio_quantise(w.y, 35.0f),
                     radio_quantise(w.z, 35.0f));
}
// QuadcopterController::Run (Offboard/QuadcopterController.cpp:11-74)
template<bool PARITY, typename P>
AGF_DEV float4 offboard_command(const OffboardParams& c, const V3<P>& curPos, const V3<P>& curVel, const Q4<P>& curAtt,
                                const double des[3], const double desVel[3], const double desAcc[3], double yaw, double& thrustOut,
                                V3<double>& wOut) {
  const V3<float> e3(0, 0, 1);
  const V3<float> estPos = V3<float>(float(curPos.x), float(curPos.y), float(curPos.z));
  const V3<float> estVel = V3<float>(float(curVel.x), float(curVel.y), float(curVel.z));
  const Q4<float> attf = Q4<float>(float(curAtt.w), float(curAtt.x), float(curAtt.y), float(curAtt.z));
  const V3<float> cmdAcc = offboard_des_acc(c, estPos, estVel, v3f(des), v3f(desVel), v3f(desAcc));
  V3<float> cmdProperAcc = cmdAcc + V3<float>(0, 0, 9.81f);
  {
    const float n0 = norm(cmdProperAcc);
    if (double(n0) > c.max_proper) cmdProperAcc = cmdProperAcc * float(c.max_proper / double(norm(cmdProperAcc)));
  }
  if (double(cmdProperAcc.z) < c.min_vert) cmdProperAcc.z = float(c.min_vert);
  const float normCmdProperAcc = norm(cmdProperAcc);
  const V3<float> cmdThrustDir = vdiv<PARITY>(cmdProperAcc, normCmdProperAcc);
  double outCmdThrust = double(normCmdProperAcc * dot(qrot(attf, V3<float>(0, 0, 1)), cmdThrustDir));
  if (outCmdThrust < c.min_proper) outCmdThrust = c.min_proper;
  const Q4<float> cmdAttYawed = offboard_att_from_dir<PARITY>(cmdThrustDir, float(yaw));
  const V3<float> w = ctl_att_core<PARITY>(c.tc_att_xy, c.tc_att_z, c.k3_att, c.k12_att, cmdAttYawed, attf);
  thrustOut = outCmdThrust;
  wOut = V3<double>(double(w.x), double(w.y), double(w.z));
  return offboard_rates_packet(outCmdThrust, w);
}
// QuadcopterController::RunTracking (Offboard/QuadcopterController.cpp:76-131)
template<bool PARITY, typename P>
AGF_DEV float4 offboard_tracking(const OffboardParams& c, const V3<P>& curPos, const V3<P>& curVel, const Q4<P>& curAtt,
                                 const double refPos[3], const double refVel[3], const double refAcc[3], double yaw,
                                 double refThrust, const double refAngVel[3], double& thrustOut, V3<double>& wOut) {
  const V3<float> estPos = V3<float>(float(curPos.x), float(curPos.y), float(curPos.z));
  const V3<float> estVel = V3<float>(float(curVel.x), float(curVel.y), float(curVel.z));
  const Q4<float> attf = Q4<float>(float(curAtt.w), float(curAtt.x), float(curAtt.y), float(curAtt.z));
  const V3<float> accErr = offboard_des_acc(c, estPos, estVel, v3f(refPos), v3f(refVel), V3<float>(0.0f, 0.0f, 0.0f));
  // double = double + float
  const double outCmdThrust = refThrust + double(dot(accErr, qrot(attf, V3<float>(0, 0, 1))));
  // (refAcc + accErr + Vec3f(0,0,9.81f)): Vec3d + Vec3f converts the right operand to Vec3d; GetNorm2() in double,
  // stored to float; the division is Vec3d / double(float)
  const V3<double> sum = (V3<double>(refAcc[0], refAcc[1], refAcc[2]) + V3<double>(double(accErr.x), double(accErr.y), double(accErr.z))) +
                         V3<double>(0.0, 0.0, double(9.81f));
  const float normRefProperAcc = float(norm(sum));
  const V3<double> dird = sum / double(normRefProperAcc);
  const V3<float> refThrustDir(float(dird.x), float(dird.y), float(dird.z));
  const Q4<float> refAttYawed = offboard_att_from_dir<PARITY>(refThrustDir, float(yaw));
  const V3<float> angVelErr = ctl_att_core<PARITY>(c.tc_att_xy, c.tc_att_z, c.k3_att, c.k12_att, refAttYawed, attf);
  // outCmdAngVel = refAngVel + Vec3d(Vec3f(...)); CreateRatesCommand takes Vec3f(cmdAngVel)
  wOut = V3<double>(refAngVel[0] + double(angVelErr.x), refAngVel[1] + double(angVelErr.y), refAngVel[2] + double(angVelErr.z));
  thrustOut = outCmdThrust;
  const V3<float> w(float(wOut.x), float(wOut.y), float(wOut.z));
  return offboard_rates_packet(outCmdThrust, w);
}

// ---- reference generators (agrifly_b200.h "offboard loop: reference generators") --------------
// queue payload markers: x = +inf "no command this round" (wait stage), x = NaN "idle command"
AGF_DEV float4 offboard_no_command() { return make_float4(1.0f / 0.0f, 0.0f, 0.0f, 0.0f); }
AGF_DEV float4 offboard_idle_command() { return make_float4(0.0f / 0.0f, 0.0f, 0.0f, 0.0f); }
AGF_DEV float4 offboard_kill_command() { return make_float4(-1.0f / 0.0f, 0.0f, 0.0f, 0.0f); }  // x = -inf: emergency kill
// SingleAxisTrajectory::GetPosition / GetVelocity / GetAcceleration (SingleAxisTrajectory.hpp:57-63); q: p0 v0 a0 alpha beta gamma
AGF_DEV double sat_pos(const double* q, double t) {
  return q[0] + q[1] * t + (1 / 2.0) * q[2] * t * t + (1 / 6.0) * q[5] * t * t * t + (1 / 24.0) * q[4] * t * t * t * t +
         (1 / 120.0) * q[3] * t * t * t * t * t;
}
AGF_DEV double sat_vel(const double* q, double t) {
  return q[1] + q[2] * t + (1 / 2.0) * q[5] * t * t + (1 / 6.0) * q[4] * t * t * t + (1 / 24.0) * q[3] * t * t * t * t;
}
AGF_DEV double sat_acc(const double* q, double t) { return q[2] + q[5] * t + (1 / 2.0) * q[4] * t * t + (1 / 6.0) * q[3] * t * t * t; }
// (GetAcceleration(t) - _grav): thrust vector of the primitive (RapidTrajectoryGenerator.hpp:187-194)
AGF_DEV V3<double> rtg_thrust_vec(const double* tr, double t) {
  return V3<double>(sat_acc(tr, t) - tr[18], sat_acc(tr + 6, t) - tr[19], sat_acc(tr + 12, t) - tr[20]);
}
AGF_DEV V3<double> unit_vector_d(const V3<double>& v) {  // Vec3.hpp:126-129: the norm is stored in a float
  const float n = float(norm(v));
  return v / double(n);
}
// RapidTrajectoryGenerator::GetOmega (RapidTrajectoryGenerator.cpp:264-286)
template<bool PARITY>
AGF_DEV V3<double> rtg_omega(const double* tr, double t, double timeStep) {
  const V3<double> n0 = unit_vector_d(rtg_thrust_vec(tr, t));
  const V3<double> n1 = unit_vector_d(rtg_thrust_vec(tr, t + timeStep));
  const V3<double> crossProd = cross(n0, n1);
  if (norm(crossProd) <= 1e-6) return V3<double>(0, 0, 0);
  const V3<double> n = unit_vector_d(crossProd);
  const double d = dot(n0, n1);
  if (d > 1.0 || d < -1.0 || d != d) return V3<double>(0, 0, 0);  // errno set by acos (domain error)
  const double angle = Mf<PARITY>::acos(d) / timeStep;
  return angle * n;
}

// ---------------------------------------------------------------------------------------------
// Offboard::MocapStateEstimator + PredictionPipe per vehicle (agrifly_b200.h "offboard loop: state estimator";
// Components/Offboard/MocapStateEstimator.cpp, PredictionPipe.hpp).  State blocked by warp in HBM (agf_types.h est_index), touched on
// mocap packets (every 2-3 ticks) and at command generation only; everything here is double as in the reference.
// ---------------------------------------------------------------------------------------------
// The per-vehicle arrays of the offboard loop are read through L2 (ld.global.cg), like the vehicle state: with the
// balanced schedule another CTA may have written them earlier in the same launch, and L1 is not coherent.  (Reading the
// estimator's state through L1 instead is safe by ownership -- one thread writes a vehicle's lines, hand-overs between CTAs
// are release / acquire ordered -- and was measured: SLOWER, 1.39e10 -> 1.29e10 vehicle-steps/s; 23 KB of estimator
// state per warp evict the loop's own lines.)
AGF_DEV double ldg2(const double* p) { return ldcg_(p); }
// E: the type the estimator COMPUTES in.  double = the reference's (parity variants, and the fast variants with an FP64
// plant); float in the FP32 fast variants ("FP32 mode": the plant it estimates is float as well) -- the state is stored in
// double either way, and every time (estimate time, activation times, integration steps) stays double.
template<typename E>
struct EstCoreT {
  V3<E> pos, vel, w;
  Q4<E> att;
};
typedef EstCoreT<double> EstCore;
template<typename E>
AGF_DEV void est_load(const double* st, size_t n, EstCoreT<E>& e) {
  e.pos = V3<E>(E(ldg2(st + (E_POS + 0) * n)), E(ldg2(st + (E_POS + 1) * n)), E(ldg2(st + (E_POS + 2) * n)));
  e.vel = V3<E>(E(ldg2(st + (E_VEL + 0) * n)), E(ldg2(st + (E_VEL + 1) * n)), E(ldg2(st + (E_VEL + 2) * n)));
  e.w = V3<E>(E(ldg2(st + (E_W + 0) * n)), E(ldg2(st + (E_W + 1) * n)), E(ldg2(st + (E_W + 2) * n)));
  e.att = Q4<E>(E(ldg2(st + (E_ATT + 0) * n)), E(ldg2(st + (E_ATT + 1) * n)), E(ldg2(st + (E_ATT + 2) * n)), E(ldg2(st + (E_ATT + 3) * n)));
}
template<typename E>
AGF_DEV void est_store(double* st, size_t n, const EstCoreT<E>& e) {
  st[(E_POS + 0) * n] = e.pos.x; st[(E_POS + 1) * n] = e.pos.y; st[(E_POS + 2) * n] = e.pos.z;
  st[(E_VEL + 0) * n] = e.vel.x; st[(E_VEL + 1) * n] = e.vel.y; st[(E_VEL + 2) * n] = e.vel.z;
  st[(E_W + 0) * n] = e.w.x; st[(E_W + 1) * n] = e.w.y; st[(E_W + 2) * n] = e.w.z;
  st[(E_ATT + 0) * n] = e.att.w; st[(E_ATT + 1) * n] = e.att.x; st[(E_ATT + 2) * n] = e.att.y; st[(E_ATT + 3) * n] = e.att.z;
}
template<typename E>
struct EstMsgT {
  V3<E> acc, w;
  bool ballistic;
};
// The activation times of all message slots, fetched in one go: the scans below then run on registers instead of a
// chain of dependent L2 round trips (the estimator state is read through L2, see ldg2).  A free slot reads E_SLOT_FREE.
struct EstPipe {
  double ta[AGF_OFFEST_PIPE];
};
AGF_DEV void est_pipe_load(const double* st, size_t n, EstPipe& p) {
#pragma unroll
  for (int k = 0; k < AGF_OFFEST_PIPE; k++) p.ta[k] = ldg2(st + size_t(E_PIPE + E_MSG * k) * n);
}
AGF_DEV int est_pipe_count(const EstPipe& p) {
  int c = 0;
#pragma unroll
  for (int k = 0; k < AGF_OFFEST_PIPE; k++) c += p.ta[k] < E_SLOT_FREE ? 1 : 0;
  return c;
}
// PredictionPipe::GetActiveMessage (PredictionPipe.hpp:32-53) + the "no messages" default of its callers.  The reference
// walks from the newest message to the first one whose activation time has passed; activation times increase with the
// order of arrival, so that message is the active one with the largest time, and the `tLast` it has when it stops is
// the smallest time among the not-yet-active ones (agf_types.h: the slots are an unordered set).
template<typename E>
AGF_DEV void est_fetch(const double* st, size_t n, const EstPipe& pipe, double t, EstMsgT<E>& m, double& timeRemaining) {
  double tLast = 1e10;
  int hit = -1;
  double taHit = -1e300;
#pragma unroll
  for (int k = 0; k < AGF_OFFEST_PIPE; k++) {
    const double ta = pipe.ta[k];
    if (ta < E_SLOT_FREE) {
      if ((t + 1e-6) >= ta) {
        if (ta > taHit) {
          taHit = ta;
          hit = k;
        }
      } else if (ta < tLast) {
        tLast = ta;
      }
    }
  }
  if (hit >= 0) {
    const double* q = st + size_t(E_PIPE + E_MSG * hit) * n;
    m.acc = V3<E>(E(ldg2(q + 1 * n)), E(ldg2(q + 2 * n)), E(ldg2(q + 3 * n)));
    m.w = V3<E>(E(ldg2(q + 4 * n)), E(ldg2(q + 5 * n)), E(ldg2(q + 6 * n)));
    m.ballistic = ldg2(q + 7 * n) != 0.0;
    timeRemaining = tLast - taHit;
    return;
  }
  m.acc = V3<E>(0, 0, 0);
  m.w = V3<E>(0, 0, 0);
  m.ballistic = true;
  timeRemaining = 1e10;
}
template<bool PARITY>
AGF_DEV Q4<double> est_rotvec_q(const V3<double>& rv) {  // Rotationd::FromRotationVector
  Q4<double> d;
  return q_from_rotvec<PARITY>(rv, d) ? d : Q4<double>(1, 0, 0, 0);
}
template<bool PARITY>
AGF_DEV V3<double> est_q_to_rotvec(const Q4<double>& q) {  // Rotation.hpp:144-161
  const V3<double> nv = q.w > 0 ? V3<double>(q.x, q.y, q.z) : V3<double>(-q.x, -q.y, -q.z);
  const double nn = norm(nv);
  const double angle = Mf<PARITY>::asin(nn) * 2;
  if (angle < 4.84813681e-6) return V3<double>(0, 0, 0);
  return nv * (angle / nn);
}
// Fast variants of the estimator's pieces.  The estimator stays in double, but what the reference computes through libm
// (sin / cos / asin / acos / exp / sqrt of doubles, ~50-150 instructions each on the GPU) is evaluated where the values
// live: per-step rotation increments and measurement errors are small angles (polynomials in the squared angle, general
// routine out of line beyond their range), the 6-sigma gates compare squares, the first-order lag of the angular
// velocity takes a float exponential.  Tolerance: tests/test_parity_gpu.py::test_offboard_estimator_fast_variants.
template<bool PARITY, typename E>
AGF_DEV Q4<E> est_apply_rotvec(const Q4<E>& a, const V3<E>& rv) {  // a * Rotationd::FromRotationVector(rv)
  if constexpr (PARITY) return qmul(a, est_rotvec_q<true>(rv));
  else return q_apply_rotvec<false>(a, rv);
}
static AGF_COLD V3<double> est_q_to_rotvec_general(Q4<double> q) { return est_q_to_rotvec<false>(q); }
template<typename E>
AGF_DEV V3<E> est_q_to_rotvec_fast(const Q4<E>& q) {
  const V3<E> nv = q.w > 0 ? V3<E>(q.x, q.y, q.z) : V3<E>(-q.x, -q.y, -q.z);
  const E x2 = dot(nv, nv);  // sin^2(angle / 2)
  if (AGF_UNLIKELY(x2 > E(0.01))) {
    const V3<double> r = est_q_to_rotvec_general(Q4<double>(double(q.w), double(q.x), double(q.y), double(q.z)));
    return V3<E>(E(r.x), E(r.y), E(r.z));
  }
  if (x2 < E(5.8761074e-12)) return V3<E>(0, 0, 0);  // angle < MIN_ANGLE (Rotation.hpp:150)
  // angle / |nv| = 2 asin(x) / x = 2 (1 + x^2/6 + 3x^4/40 + 15x^6/336 + 105x^8/3456 + 945x^10/42240 + ...), |error| < 2e-14 here
  E p = x2 * E(0.017352764423076924) + E(0.022372159090909092);
  p = x2 * p + E(0.030381944444444444);
  p = x2 * p + E(0.044642857142857144);
  p = x2 * p + E(0.075);
  p = x2 * p + E(0.16666666666666666);
  p = x2 * p + E(1.0);
  return nv * (E(2.0) * p);
}
template<bool PARITY, typename E>
AGF_DEV E est_lag_factor(const EstParams& ep, double dtInt) {  // exp(-dt / tau) of the angular-velocity model (:95, :156)
  if constexpr (PARITY) {
    return Mf<true>::exp(-dtInt / ep.tc_angvel);
  } else {
#if defined(__CUDA_ARCH__)
    return E(__expf(float(-dtInt * ep.inv_tc_angvel)));
#else
    return E(::expf(float(-dtInt * ep.inv_tc_angvel)));
#endif
  }
}
// A V A^T + Q with A = [[1, dt], [0, 1]] (:171-186), written out (fast variants)
template<typename E>
AGF_DEV void est_propagate_var_fast(E* v, E dtInt, E proc) {
  const E d2 = dtInt * dtInt;
  v[0] = (v[0] + dtInt * ((v[1] + v[2]) + dtInt * v[3])) + d2 * d2 * proc * E(0.25);
  v[1] = v[1] + dtInt * v[3];
  v[2] = v[2] + dtInt * v[3];
  v[3] = v[3] + d2 * proc;
}

// MocapStateEstimator::GetPrediction (MocapStateEstimator.cpp:61-118)
template<bool PARITY, typename E>
AGF_DEV void mocap_predict(const EstParams& ep, size_t i, uint64_t now_us, double dt, EstCoreT<E>& o, EstPipe& pipe) {
  constexpr size_t n = E_LANES;
  const double* st = ep.state + est_index(i);
  const double tEnd = dt + double(now_us - ep.t0_us) * 1e-6;
  const double tStart = double(uint64_t(ldg2(st + E_TEST * n))) * 1e-6;
  EstCoreT<E> m;
  est_pipe_load(st, n, pipe);
  est_load(st, n, m);
  o = m;
  double t = tStart;
  while ((t + 1e-6) < tEnd) {
    EstMsgT<E> cmd;
    double predictionTime;
    est_fetch(st, n, pipe, t, cmd, predictionTime);
    double dtInt = tEnd - t;
    if (dtInt > (predictionTime + 1e-6)) dtInt = predictionTime;
    const E h = E(dtInt);
    const V3<E> newPos = (o.pos + m.vel * h) + ((cmd.acc * h) * h) / E(2.0);  // sic: _vel (:90)
    const V3<E> newVel = o.vel + cmd.acc * h;
    const Q4<E> newAtt = est_apply_rotvec<PARITY>(o.att, m.w * h);  // sic: _angVel (:92)
    E discrete = est_lag_factor<PARITY, E>(ep, dtInt);
    if (cmd.ballistic) discrete = 1;
    const V3<E> newW = discrete * o.w + (E(1) - discrete) * cmd.w;
    o.pos = newPos; o.vel = newVel; o.att = newAtt; o.w = newW;
    t += dtInt;
  }
}
template<typename E>
AGF_DEV void est_reset_variance(E* vp, E* va) {  // :52-60
  vp[0] = 25.0; vp[3] = 25.0; vp[1] = vp[2] = 0.0;
  va[0] = 1.0; va[3] = 400; va[1] = va[2] = 0.0;
}
// C = A * B for 2x2 row-major, coefficient-wise, sequential in k (the oracle's Eigen contract)
AGF_DEV void est_mm2(const double* a, const double* b, double* c) {
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int q = 0; q < 2; q++) {
      double acc = a[2 * r] * b[q];
      acc += a[2 * r + 1] * b[2 + q];
      c[2 * r + q] = acc;
    }
}
AGF_DEV void est_propagate_var(double* v, double dtInt, double proc) {  // A V A^T + Q (:171-186)
  const double A[4] = {1, dtInt, 0, 1}, At[4] = {1, 0, dtInt, 1};
  double t1[4], t2[4];
  est_mm2(A, v, t1);
  est_mm2(t1, At, t2);
  v[0] = t2[0] + dtInt * dtInt * dtInt * dtInt * proc / 4;
  v[1] = t2[1] + 0.0;
  v[2] = t2[2] + 0.0;
  v[3] = t2[3] + dtInt * dtInt * proc;
}
// MocapStateEstimator::UpdateWithMeasurement (:120-265)
template<bool PARITY, typename E>
AGF_DEV void mocap_update(const EstParams& ep, size_t i, uint64_t now_us, const V3<E>& measPos, const Q4<E>& measAtt) {
  constexpr size_t n = E_LANES;
  double* st = ep.state + est_index(i);
  EstCoreT<E> e;
  E vp[4], va[4];
  // everything the update reads, requested before the first use (one L2 round trip instead of one per group)
  const double inited = ldg2(st + E_INIT * n);
  EstPipe pipe;
  est_pipe_load(st, n, pipe);
  est_load(st, n, e);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    vp[k] = E(ldg2(st + (E_VP + k) * n));
    va[k] = E(ldg2(st + (E_VA + k) * n));
  }
  uint64_t est_us = uint64_t(ldg2(st + E_TEST * n));
  double nrej = ldg2(st + E_NREJ * n), nrejc = ldg2(st + E_NREJC * n);
  if (inited == 0.0) {
    e.pos = measPos;
    e.vel = V3<E>(0, 0, 0);
    e.att = measAtt;
    e.w = V3<E>(0, 0, 0);
    est_store(st, n, e);
    est_reset_variance(vp, va);
    for (int k = 0; k < 4; k++) {
      st[(E_VP + k) * n] = vp[k];
      st[(E_VA + k) * n] = va[k];
    }
    st[E_INIT * n] = 1.0;
    st[E_LASTGOOD * n] = double(now_us);
    return;
  }
  const double t0 = double(est_us) * 1e-6;
  const double tEnd = double(now_us - ep.t0_us) * 1e-6;
  if (tEnd > t0) {
    for (;;) {
      const double tNow = double(est_us) * 1e-6;
      if ((tNow + 1e-6) >= tEnd) break;
      EstMsgT<E> p;
      double predictionTime;
      est_fetch(st, n, pipe, tNow, p, predictionTime);
      double dtInt = tEnd - tNow;
      if (dtInt > (predictionTime + 1e-6)) dtInt = predictionTime;
      const EstCoreT<E> c = e;
      const E h = E(dtInt);
      e.pos = c.pos + c.vel * h;
      e.vel = c.vel + p.acc * h;
      e.att = est_apply_rotvec<PARITY>(c.att, c.w * h);
      E discrete = est_lag_factor<PARITY, E>(ep, dtInt);
      if (p.ballistic) discrete = 1;
      e.w = discrete * c.w + (E(1) - discrete) * p.w;
      est_us += uint64_t(0.5 + dtInt * 1e6);
      if constexpr (PARITY) {
        est_propagate_var(vp, dtInt, ep.proc_pos);
        est_propagate_var(va, dtInt, ep.proc_att);
      } else {
        est_propagate_var_fast(vp, h, E(ep.proc_pos));
        est_propagate_var_fast(va, h, E(ep.proc_att));
      }
    }
  }
  E innovP = vp[0] + E(ep.meas_pos * ep.meas_pos);
  E innovA = va[0] + E(ep.meas_att * ep.meas_att);
  bool reject;
  if constexpr (PARITY) {
    const double distP = norm(measPos - e.pos) / ::sqrt(3 * innovP);
    const Q4<double> dq = qmul(qinv(measAtt), e.att);
    const double distA = (Mf<PARITY>::acos(::fabs(dq.w)) * 2.0) / ::sqrt(innovA);  // Rotation::GetAngle
    reject = (distP > ep.reject) || (distA > ep.reject);
  } else {
    // the same two gates without square roots and acos: |d|^2 > r^2 * 3 S_p ; angle/2 > r sqrt(S_a)/2 <=> |dq.w| < cos(..)
    const V3<E> d = measPos - e.pos;
    const E dqw = rabs_(measAtt.w * e.att.w + measAtt.x * e.att.x + measAtt.y * e.att.y + measAtt.z * e.att.z);
    const float half = 0.5f * float(ep.reject) * ::sqrtf(float(innovA));
    reject = (dot(d, d) > E(ep.reject * ep.reject) * (E(3) * innovP)) || (half < 1.5707963f && float(dqw) < ::cosf(half));
  }
  if (reject && nrejc < 10.0) {
    nrej += 1.0;
    nrejc += 1.0;
  } else {
    if (nrejc >= 10.0) {  // forced acceptance: Reset() (:37-50), then the update against the zeroed state
      e.pos = V3<E>(0, 0, 0);
      e.vel = V3<E>(0, 0, 0);
      e.att = Q4<E>(1, 0, 0, 0);
      e.w = V3<E>(0, 0, 0);
      est_reset_variance(vp, va);
      est_us = now_us - ep.t0_us;
      st[E_INIT * n] = 0.0;
      innovP = vp[0] + E(ep.meas_pos * ep.meas_pos);
      innovA = va[0] + E(ep.meas_att * ep.meas_att);
    }
    nrejc = 0.0;
    st[E_LASTGOOD * n] = double(now_us);
    const E sP = 1 / innovP, sA = 1 / innovA;
    E gP[2], gA[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {  // V * H^T * (1 / S), H = [1 0]
      E a = vp[2 * r] * E(1.0);
      a += vp[2 * r + 1] * E(0.0);
      gP[r] = a * sP;
      E b = va[2 * r] * E(1.0);
      b += va[2 * r + 1] * E(0.0);
      gA[r] = b * sA;
    }
    const V3<E> errP = measPos - e.pos;
    e.pos = e.pos + gP[0] * errP;
    e.vel = e.vel + gP[1] * errP;
    const Q4<E> qerr = qmul(qinv(e.att), measAtt);
    V3<E> errA;
    if constexpr (PARITY) errA = est_q_to_rotvec<true>(qerr);
    else errA = est_q_to_rotvec_fast(qerr);
    e.att = est_apply_rotvec<PARITY>(e.att, gA[0] * errA);
    e.w = e.w + gA[1] * errA;
    if constexpr (PARITY) {
      const double Ip[4] = {1.0 - gP[0] * 1.0, 0.0 - gP[0] * 0.0, 0.0 - gP[1] * 1.0, 1.0 - gP[1] * 0.0};
      const double Ia[4] = {1.0 - gA[0] * 1.0, 0.0 - gA[0] * 0.0, 0.0 - gA[1] * 1.0, 1.0 - gA[1] * 0.0};
      double np_[4], na_[4];
      est_mm2(Ip, vp, np_);
      est_mm2(Ia, va, na_);
      for (int k = 0; k < 4; k++) {
        vp[k] = np_[k];
        va[k] = na_[k];
      }
    } else {  // (I - K H) V with H = [1 0], written out
      const E p0 = vp[0], p1 = vp[1], a0 = va[0], a1 = va[1];
      vp[0] = (E(1) - gP[0]) * p0; vp[1] = (E(1) - gP[0]) * p1; vp[2] = vp[2] - gP[1] * p0; vp[3] = vp[3] - gP[1] * p1;
      va[0] = (E(1) - gA[0]) * a0; va[1] = (E(1) - gA[0]) * a1; va[2] = va[2] - gA[1] * a0; va[3] = va[3] - gA[1] * a1;
    }
  }
  {  // symmetry (:257-261)
    const E p01 = (vp[1] + vp[2]) * E(0.5), p10 = (vp[2] + vp[1]) * E(0.5), a01 = (va[1] + va[2]) * E(0.5), a10 = (va[2] + va[1]) * E(0.5);
    vp[0] = (vp[0] + vp[0]) * E(0.5); vp[3] = (vp[3] + vp[3]) * E(0.5); vp[1] = p01; vp[2] = p10;
    va[0] = (va[0] + va[0]) * E(0.5); va[3] = (va[3] + va[3]) * E(0.5); va[1] = a01; va[2] = a10;
  }
  est_store(st, n, e);
  for (int k = 0; k < 4; k++) {
    st[(E_VP + k) * n] = vp[k];
    st[(E_VA + k) * n] = va[k];
  }
  st[E_TEST * n] = double(est_us);
  st[E_NREJ * n] = nrej;
  st[E_NREJC * n] = nrejc;
  {  // PredictionPipe::ClearExpiredMessages(estimate time) (PredictionPipe.hpp:55-68): the front message goes while the
    // second one is already active, i.e. every active message but the newest is freed (agf_types.h)
    const double cur = double(est_us) * 1e-6;
    double newest = -1e300;
#pragma unroll
    for (int k = 0; k < AGF_OFFEST_PIPE; k++)
      if (pipe.ta[k] <= cur && pipe.ta[k] > newest) newest = pipe.ta[k];
    int drop = 0;
#pragma unroll
    for (int k = 0; k < AGF_OFFEST_PIPE; k++)
      if (pipe.ta[k] < newest) {
        st[size_t(E_PIPE + E_MSG * k) * n] = E_SLOT_FREE;
        drop++;
      }
    if (drop > 0) st[E_NPIPE * n] = double(est_pipe_count(pipe) - drop);
  }
}
// MocapStateEstimator::SetPredictedValues -> PredictionPipe::AddMessage (hpp:74-80, PredictionPipe.hpp:25-30)
template<typename E>
AGF_DEV void mocap_set_predicted(const EstParams& ep, size_t i, uint64_t now_us, const EstPipe& pipe, const V3<E>& w,
                                    const V3<E>& acc) {
  constexpr size_t n = E_LANES;
  double* st = ep.state + est_index(i);
  // `pipe`: the slots as mocap_predict read them for this command (nothing touches them in between)
  int cnt = 0, slot = -1, oldest = 0;
  double t_oldest = E_SLOT_FREE;
#pragma unroll
  for (int k = AGF_OFFEST_PIPE - 1; k >= 0; k--) {
    if (pipe.ta[k] < E_SLOT_FREE) cnt++;
    else slot = k;
    if (pipe.ta[k] <= t_oldest) {
      t_oldest = pipe.ta[k];
      oldest = k;
    }
  }
  if (slot < 0) {  // full: cannot happen while measurements arrive; keep the newest messages
    slot = oldest;
    cnt--;
  }
  double* q = st + size_t(E_PIPE + E_MSG * slot) * n;
  q[0] = double(now_us - ep.t0_us) * 1e-6 + ep.delay;
  q[1 * n] = acc.x; q[2 * n] = acc.y; q[3 * n] = acc.z;
  q[4 * n] = w.x; q[5 * n] = w.y; q[6 * n] = w.z;
  q[7 * n] = 0.0;
  st[E_NPIPE * n] = double(cnt + 1);
}
// compute type of the fast variants' estimator: float with an FP32 plant, double with an FP64 plant (see EstCoreT)
template<typename P> struct EstOf { typedef double type; };
template<> struct EstOf<float> { typedef float type; };
template<typename P>
static AGF_COLD void mocap_update_cold(const EstParams* ep, size_t i, uint64_t now_us, V3<P> p, Q4<P> a) {
  typedef typename EstOf<P>::type E;
  mocap_update<false, E>(*ep, i, now_us, V3<E>(E(p.x), E(p.y), E(p.z)), Q4<E>(E(a.w), E(a.x), E(a.y), E(a.z)));
}

// One round of the offboard main loop for vehicle i at clock reading t_gen: desired state from the configured
// generator, then the controller; returns the queue payload.
template<bool PARITY>
AGF_DEV float4 offboard_generate_core(const OffboardParams& c, size_t i, size_t n, uint64_t t_gen, const V3<double>& cp,
                                      const V3<double>& cv, const Q4<double>& ca, int& predicted, double& thrustOut, V3<double>& wOut) {
  typedef double P;
  const double zero3[3] = {0.0, 0.0, 0.0};
  predicted = 2;  // 0: no SetPredictedValues, 1: SetPredictedValues(0, 0), 2: from the command
  double des[3];
  if (c.ref_kind == AGF_OFFREF_TARGETS) {
    uint32_t ti = 0;
    for (uint32_t j = 1; j < c.n_targets; j++)
      if (c.targets[j].time_us <= t_gen) ti = j;
    des[0] = c.targets[ti].pos[0]; des[1] = c.targets[ti].pos[1]; des[2] = c.targets[ti].pos[2];
  } else {
    des[0] = c.desired[0]; des[1] = c.desired[1]; des[2] = c.desired[2];
  }
  if (c.offsets) {
    des[0] = des[0] + c.offsets[i];
    des[1] = des[1] + c.offsets[n + i];
    des[2] = des[2] + c.offsets[2 * n + i];
  }
  if (c.ref_kind == AGF_OFFREF_TARGETS) return offboard_command<PARITY, P>(c, cp, cv, ca, des, zero3, zero3, double(c.yaw), thrustOut, wOut);
  if (c.ref_kind == AGF_OFFREF_TRAJECTORY) {
    if (!(t_gen > c.start_us)) return offboard_command<PARITY, P>(c, cp, cv, ca, des, zero3, zero3, c.desired_yaw, thrustOut, wOut);
    double tr[AGF_OFFTRAJ_DOUBLES];
    for (int k = 0; k < AGF_OFFTRAJ_DOUBLES; k++) tr[k] = ldg2(c.traj + size_t(k) * n + i);
    double traj_t = double(t_gen - c.start_us) * 1e-6;
    const double tEnd = tr[21];
    double tp[3], tv[3], ta[3];
    if (traj_t < tEnd) {
      traj_t += 0.04;
      for (int a = 0; a < 3; a++) {
        tp[a] = sat_pos(tr + 6 * a, traj_t); tv[a] = sat_vel(tr + 6 * a, traj_t); ta[a] = sat_acc(tr + 6 * a, traj_t);
      }
    } else {
      for (int a = 0; a < 3; a++) {
        tp[a] = sat_pos(tr + 6 * a, tEnd); tv[a] = 0; ta[a] = 0;
      }
    }
    if (tp[2] < 0) {
      tp[2] = 0;
      if (tv[2] < 0) tv[2] = 0;
      if (ta[2] < 0) ta[2] = 0;
    }
    const Q4<double> trajAtt(tr[22], tr[23], tr[24], tr[25]);
    const V3<double> rp = qrot(trajAtt, V3<double>(tp[0], tp[1], tp[2])) + V3<double>(tr[26], tr[27], tr[28]);
    const V3<double> rv = qrot(trajAtt, V3<double>(tv[0], tv[1], tv[2]));
    const V3<double> ra = qrot(trajAtt, V3<double>(ta[0], ta[1], ta[2]));
    const double refThrust = norm(rtg_thrust_vec(tr, traj_t));
    const Q4<double> cad(double(ca.w), double(ca.x), double(ca.y), double(ca.z));
    const V3<double> rw = qrot(qmul(qinv(cad), trajAtt), rtg_omega<PARITY>(tr, traj_t, 0.02));
    const double refPos[3] = {rp.x, rp.y, rp.z}, refVel[3] = {rv.x, rv.y, rv.z}, refAcc[3] = {ra.x, ra.y, ra.z},
                 refW[3] = {rw.x, rw.y, rw.z};
    return offboard_tracking<PARITY, P>(c, cp, cv, ca, refPos, refVel, refAcc, c.desired_yaw, refThrust, refW, thrustOut, wOut);
  }
  // AGF_OFFREF_STAGES: ExampleVehicleStateMachine::Run (ExampleVehicleStateMachine.cpp:93-370)
  double* st = c.state + i;  // field k at st[k * n]
  const bool shouldStart = t_gen >= c.start_us, shouldStop = t_gen >= c.stop_us;
  int stage = int(ldg2(st));
  const bool stageChange = stage != int(ldg2(st + n));
  st[n] = double(stage);
  const uint64_t stageStart = stageChange ? t_gen : uint64_t(ldg2(st + 2 * n));
  if (stageChange) st[2 * n] = double(t_gen);
  const double ts = double(t_gen - stageStart) * 1e-6;  // _stageTimer->GetSeconds<double>()
  const int stage_in = stage;
  float4 out;
  double cmdYaw = ldg2(st + 15 * n);
  // SafetyNet::UpdateWithEstimator + GetIsSafe (SafetyNet.hpp:70-106), on every Run()
  bool safe = true;
  if (c.safety_net) {
    const double since = c.est.kind == AGF_OFFEST_MOCAP ? double(t_gen - uint64_t(ldg2(c.est.state + est_index(i) + size_t(E_LASTGOOD) * E_LANES))) * 1e-6 : 0.0;
    const bool notSeen = since > c.not_seen_timeout;
    bool unsafePos = false;
    const double ep[3] = {cp.x, cp.y, cp.z};
    for (int a = 0; a < 3; a++) {
      if (ep[a] < c.safe_min[a]) unsafePos = true;
      if (ep[a] > c.safe_max[a]) unsafePos = true;
    }
    bool upsideDownAndLow = false;
    if (cp.z < c.min_normal_height) {
      if (qrot(ca, V3<double>(0, 0, 1)).z < 0) upsideDownAndLow = true;
    }
    safe = !(notSeen || unsafePos || upsideDownAndLow);
  }
  switch (stage) {
    case AGF_STAGE_WAIT_FOR_START:
      if (shouldStart) stage = AGF_STAGE_SPOOL_UP;
      out = offboard_no_command();
      predicted = 0;
      break;
    case AGF_STAGE_SPOOL_UP:
      if (!safe) stage = AGF_STAGE_EMERGENCY;
      out = offboard_rates_packet(9.81 * 0.25, V3<float>(0.0f, 0.0f, 0.0f));
      predicted = 1;
      if (ts > 0.5) stage = AGF_STAGE_TAKEOFF;
      break;
    case AGF_STAGE_TAKEOFF: {
      if (stageChange) {
        st[3 * n] = double(cp.x); st[4 * n] = double(cp.y); st[5 * n] = double(cp.z);
      }
      if (!safe) stage = AGF_STAGE_EMERGENCY;
      double frac = ts / 2.0;
      if (frac >= 1.0) {
        stage = AGF_STAGE_FLIGHT;
        frac = 1.0;
      }
      double cmdPos[3];
      for (int a = 0; a < 3; a++) cmdPos[a] = (1 - frac) * (stageChange ? (a == 0 ? double(cp.x) : (a == 1 ? double(cp.y) : double(cp.z))) : ldg2(st + (3 + a) * n)) + frac * des[a];
      out = offboard_command<PARITY, P>(c, cp, cv, ca, cmdPos, zero3, zero3, cmdYaw, thrustOut, wOut);
    } break;
    case AGF_STAGE_FLIGHT: {
      if (!safe) stage = AGF_STAGE_EMERGENCY;
      double cmdPos[3] = {0, 0, 0}, cmdVel[3] = {0, 0, 0}, cmdAcc[3] = {0, 0, 0};
      const double t = ts;
      const double frac = (t / 2.0 < 1.0) ? t / 2.0 : 1.0;  // min(t / getIntoActionTime, 1.0)
      switch (c.traj_id) {
        case 0:
          cmdPos[0] = des[0]; cmdPos[1] = des[1]; cmdPos[2] = des[2];
          cmdYaw = 0;
          break;
        case 1: {
          const double radius = 1.0, angSpeed = 0.5;
          const double cs = Mf<PARITY>::cos(angSpeed * t), sn = Mf<PARITY>::sin(angSpeed * t);
          cmdPos[0] = 0.0 + radius * cs; cmdPos[1] = -2.0 + radius * sn; cmdPos[2] = des[2] + radius * 0;
          cmdVel[0] = (radius * angSpeed) * -sn; cmdVel[1] = (radius * angSpeed) * cs; cmdVel[2] = (radius * angSpeed) * 0;
          cmdAcc[0] = (radius * (angSpeed * angSpeed)) * -cs; cmdAcc[1] = (radius * (angSpeed * angSpeed)) * -sn;
          cmdAcc[2] = (radius * (angSpeed * angSpeed)) * 0;
          cmdYaw = c.desired_yaw + angSpeed * t;
        } break;
        case 2: {
          const double amplitude = 1.0, angFreq = 2.0;
          const double cs = Mf<PARITY>::cos(angFreq * t), sn = Mf<PARITY>::sin(angFreq * t);
          cmdPos[0] = des[0] + amplitude * 0; cmdPos[1] = des[1] + amplitude * sn; cmdPos[2] = des[2] + amplitude * 0;
          cmdVel[0] = (amplitude * angFreq) * 0; cmdVel[1] = (amplitude * angFreq) * cs; cmdVel[2] = (amplitude * angFreq) * 0;
          cmdAcc[0] = (amplitude * (angFreq * angFreq)) * 0; cmdAcc[1] = (amplitude * (angFreq * angFreq)) * -sn;
          cmdAcc[2] = (amplitude * (angFreq * angFreq)) * 0;
          cmdYaw = c.desired_yaw;
        } break;
        case 3: {
          const double radius = 0.5, angSpeed = 1;
          const double cs = Mf<PARITY>::cos(angSpeed * t), sn = Mf<PARITY>::sin(angSpeed * t);
          cmdPos[0] = 0.0 + radius * cs; cmdPos[1] = 0.0 + radius * sn; cmdPos[2] = des[2] + radius * 0;
          cmdVel[0] = (radius * angSpeed) * -sn; cmdVel[1] = (radius * angSpeed) * cs; cmdVel[2] = (radius * angSpeed) * 0;
          cmdAcc[0] = (radius * (angSpeed * angSpeed)) * -cs; cmdAcc[1] = (radius * (angSpeed * angSpeed)) * -sn;
          cmdAcc[2] = (radius * (angSpeed * angSpeed)) * 0;
          cmdYaw = 0;
        } break;
        case 4: {
          const double radius = 0.5, angSpeed = 0.5;
          const double cs = Mf<PARITY>::cos(angSpeed * t), sn = Mf<PARITY>::sin(angSpeed * t);
          const double c4 = Mf<PARITY>::cos(angSpeed * t * 4), s4 = Mf<PARITY>::sin(angSpeed * t * 4);
          cmdPos[0] = 0.0 + radius * cs; cmdPos[1] = 0.0 + radius * sn; cmdPos[2] = des[2] + radius * c4;
          cmdVel[0] = (radius * angSpeed) * -sn; cmdVel[1] = (radius * angSpeed) * cs; cmdVel[2] = (radius * angSpeed) * -s4;
          cmdAcc[0] = (radius * (angSpeed * angSpeed)) * -cs; cmdAcc[1] = (radius * (angSpeed * angSpeed)) * -sn;
          cmdAcc[2] = (radius * (angSpeed * angSpeed)) * -c4;
          cmdYaw = angSpeed * t;
        } break;
        default:
          cmdPos[0] = des[0]; cmdPos[1] = des[1]; cmdPos[2] = des[2];
          cmdYaw = 0.2 * t;
          break;
      }
      double lp[3], lv[3], la[3];
      for (int a = 0; a < 3; a++) {
        lp[a] = (1 - frac) * des[a] + frac * cmdPos[a];
        lv[a] = frac * cmdVel[a];
        la[a] = frac * cmdAcc[a];
        st[(6 + a) * n] = lp[a]; st[(9 + a) * n] = lv[a]; st[(12 + a) * n] = la[a];
      }
      out = offboard_command<PARITY, P>(c, cp, cv, ca, lp, lv, la, cmdYaw, thrustOut, wOut);
      if (shouldStop) stage = AGF_STAGE_LANDING;
    } break;
    case AGF_STAGE_LANDING: {
      if (!safe) stage = AGF_STAGE_EMERGENCY;
      const double LANDING_SPEED = 0.5;
      const double frac = (ts / 2.0 < 1.0) ? ts / 2.0 : 1.0;
      double lp[3], lv[3], la[3], cmdPos[3], dp[3], dv[3], da[3];
      const double land[3] = {0.0, 0.0, -LANDING_SPEED};
      for (int a = 0; a < 3; a++) {
        lp[a] = ldg2(st + (6 + a) * n); lv[a] = ldg2(st + (9 + a) * n); la[a] = ldg2(st + (12 + a) * n);
        cmdPos[a] = lp[a] + ts * land[a];
      }
      if (cmdPos[2] < 0) stage = AGF_STAGE_COMPLETE;
      for (int a = 0; a < 3; a++) {
        dp[a] = (1 - frac) * lp[a] + frac * cmdPos[a];
        dv[a] = (1 - frac) * lv[a] + frac * land[a];
        da[a] = (1 - frac) * la[a] + frac * 0.0;
      }
      out = offboard_command<PARITY, P>(c, cp, cv, ca, dp, dv, da, cmdYaw, thrustOut, wOut);
    } break;
    case AGF_STAGE_COMPLETE:
      out = offboard_idle_command();
      predicted = 1;
      break;
    default:  // AGF_STAGE_EMERGENCY (:350-363): kill, nothing told to the estimator
      out = offboard_kill_command();
      predicted = 0;
      break;
  }
  if (stage != stage_in) st[0] = double(stage);
  st[15 * n] = cmdYaw;
  return out;
}
// the estimate the controller sees (true state or MocapStateEstimator::GetPrediction), the command, and what the
// estimator is told about it (main.cpp:468-469,652-654)
template<bool PARITY, typename P>
AGF_DEV float4 offboard_generate(const OffboardParams& c, size_t i, size_t n, uint64_t t_gen, const V3<P>& cp, const V3<P>& cv,
                                 const Q4<P>& ca) {
  typedef typename std::conditional<PARITY, double, typename EstOf<P>::type>::type E;  // what the estimator computes in
  EstCoreT<E> e;
  EstPipe pipe;
  if (c.est.kind == AGF_OFFEST_MOCAP) {
    mocap_predict<PARITY, E>(c.est, i, t_gen, c.est.delay, e, pipe);
  } else {
    e.pos = V3<E>(E(cp.x), E(cp.y), E(cp.z));
    e.vel = V3<E>(E(cv.x), E(cv.y), E(cv.z));
    e.att = Q4<E>(E(ca.w), E(ca.x), E(ca.y), E(ca.z));
  }
  int predicted;
  double thrust = 0.0;
  V3<double> w(0, 0, 0);
  const float4 out = offboard_generate_core<PARITY>(c, i, n, t_gen, V3<double>(double(e.pos.x), double(e.pos.y), double(e.pos.z)),
                                                    V3<double>(double(e.vel.x), double(e.vel.y), double(e.vel.z)),
                                                    Q4<double>(double(e.att.w), double(e.att.x), double(e.att.y), double(e.att.z)), predicted,
                                                    thrust, w);
  if (c.est.kind == AGF_OFFEST_MOCAP) {
    if (predicted == 1) mocap_set_predicted(c.est, i, t_gen, pipe, V3<E>(0, 0, 0), V3<E>(0, 0, 0));
    if (predicted == 2)
      mocap_set_predicted(c.est, i, t_gen, pipe, V3<E>(E(w.x), E(w.y), E(w.z)), qrot(e.att, V3<E>(0, 0, 1)) * E(thrust) - V3<E>(0, 0, E(9.81)));
  }
  return out;
}
template<typename P>
static AGF_COLD float4 offboard_generate_cold(const OffboardParams* c, size_t i, size_t n, uint64_t t_gen, V3<P> cp, V3<P> cv, Q4<P> ca) {
  return offboard_generate<false, P>(*c, i, n, t_gen, cp, cv, ca);
}

// RunControllerExternalAccelerationControl up to the desired body rates (QuadcopterLogic.cpp:459-517):
// returns (desW.x, desW.y, desW.z, thrust); thrust < 0 stands for "motors off" (desired z acceleration below -g/2)
template<bool PARITY>
AGF_DEV float4 ctl_accel_mode(const LogicParams& k, const float4& estAtt4, const float4& radio) {
  const Q4<float> estAtt(estAtt4.x, estAtt4.y, estAtt4.z, estAtt4.w);
  const V3<float> desAcc(radio.x, radio.y, radio.z);
  if (desAcc.z < -9.81f / 2) return make_float4(0.0f, 0.0f, 0.0f, -1.0f);
  const V3<float> proper = desAcc + V3<float>(0, 0, 9.81f);
  const float thrust = norm(proper);
  const V3<float> dir = vdiv<PARITY>(proper, thrust);
  const Q4<float> desAtt = att_from_thrust_dir<PARITY>(dir);
  // ToEulerYPR (Rotation.hpp:163-169); yaw is computed by the reference but unused
  const Q4<float>& q = estAtt;
  const float pch = -Mf<PARITY>::asin(2.0f * q.x * q.z - 2.0f * q.w * q.y);
  const float rll = Mf<PARITY>::atan2(2.0f * q.y * q.z + 2.0f * q.w * q.x, q.z * q.z - q.y * q.y - q.x * q.x + q.w * q.w);
  const Q4<float> noYaw = q_from_euler_ypr<PARITY>(0.0f, pch, rll);
  const V3<float> desW = ctl_att<PARITY>(k, desAtt, noYaw);
  return make_float4(desW.x, desW.y, radio.w, thrust);
}
static AGF_COLD float4 ctl_accel_mode_cold(const LogicParams* k, float4 estAtt, float4 radio) {
  return ctl_accel_mode<false>(*k, estAtt, radio);
}

// ---------------------------------------------------------------------------------------------
// QuadcopterLogic::Run (QuadcopterLogic.cpp:164-219) with its intake setters
// ---------------------------------------------------------------------------------------------
template<bool PARITY, typename P, bool UWB, bool HK>
AGF_DEV void logic_run(VState<P, PARITY, UWB, HK>& s, const Scratch& sc, const StepShared<P>& p, const V3<float>& gyroMeas,
                       const V3<float>& accMeas, float kf_dt) {
  const LogicParams& k = p.logic;
  // --- intake (QuadcopterLogic.hpp:32-59) ---
  if constexpr (HK) {
    if constexpr (PARITY) {
      s.batt_vfilt = lpf2(k.lp_batt, s.batt_lp, 1, k.batt_voltage);
      lpf2(k.lp_temp, s.temp_lp, 1, 25.0f);
    } else {
      s.batt_vfilt = k.batt_voltage;  // the filter's fixed point (VState)
    }
  }
  V3<float> g = gyroMeas, a = accMeas;
  if (!k.imu_identity) {
    g = matvec(k.R_imu, g);
    a = matvec(k.R_imu, a);
  }
  // gyro calibration bias is always (0,0,0) behind this API: rawMeas - bias == rawMeas
  V3<float> gf, af;
  if constexpr (PARITY) {
    gf = V3<float>(lpf2(k.lp_gyro, &s.gyro_lp[0], 1, g.x), lpf2(k.lp_gyro, &s.gyro_lp[4], 1, g.y), lpf2(k.lp_gyro, &s.gyro_lp[8], 1, g.z));
    af = V3<float>(lpf2(k.lp_acc, &s.acc_lp[0], 1, a.x), lpf2(k.lp_acc, &s.acc_lp[4], 1, a.y), lpf2(k.lp_acc, &s.acc_lp[8], 1, a.z));
  } else {
    gf = lpf2_scratch3(k.lp_gyro, sc, SQ_LPF, g);
    af = lpf2_scratch3(k.lp_acc, sc, SQ_LPF + 3, a);
  }

  uint32_t fs = bget(s.bits, B_FS_SHIFT, B_FS_MASK);
  if (fs == AGF_FS_UNINITIALIZED) return;
  s.cycle++;
  if (HK) {  // MonitorSimplePeriod::Update (QuadcopterLogic.hpp:390-394)
    const float mdt = float(s.age_mon_loop) * 1e-6f;
    s.mon_loop_lpdt = k.mon_loop_coef <= 0.0f ? mdt : k.mon_loop_coef * s.mon_loop_lpdt + (1 - k.mon_loop_coef) * mdt;
    s.age_mon_loop -= uint32_t(mdt * 1e6f);
  }
  // --- UpdateEstimator :221-273 ---
  kf_predict<PARITY>(s, sc, gf, af, kf_dt);
  if (UWB && (s.bits & B_UWB_NEW)) {
    s.bits &= ~B_UWB_NEW;
    s.uwb_count++;  // failure is never set by the simulated network (UWBNetwork.cpp:77)
    const uint32_t resp = (s.uwbw >> W_LOGIC_TARGET) & 0xFFu;  // anchor that produced the range
    uint32_t nxt = (s.uwbw >> W_NEXT_TARGET) & 0xFFu;
    nxt = (nxt + 1u) % p.n_anchors;
    s.uwbw = (s.uwbw & ~(0xFFu << W_NEXT_TARGET)) | (nxt << W_NEXT_TARGET);
    const V3<float> tp(p.anchors[resp].x, p.anchors[resp].y, p.anchors[resp].z);
    kf_update_range<PARITY>(s, sc, tp, s.logic_range);
  }
  // --- ParseIncomingCommunications :275-303 ---
  if (s.bits & B_RADIO_NEW) {
    s.bits &= ~B_RADIO_NEW;
    if (fs != AGF_FS_PANIC && fs != AGF_FS_KILLED) {
      const uint32_t ty = bget(s.bits, B_RTYPE_SHIFT, B_RTYPE_MASK);
      if (ty == AGF_RADIO_EMERGENCY_KILL) {
        fs = AGF_FS_KILLED;
        if (!bget(s.bits, B_PANIC_SHIFT, B_PANIC_MASK)) s.bits = bset(s.bits, B_PANIC_SHIFT, B_PANIC_MASK, AGF_PANIC_KILLED_EXTERNALLY);
      } else if (ty == AGF_RADIO_POSITION_CMD) {
        fs = AGF_FS_FULLY_AUTONOMOUS;
      } else if (ty == AGF_RADIO_EXTERNAL_ACCELERATION_CMD) {
        fs = AGF_FS_EXTERNAL_ACCELERATION_CONTROL;
      } else if (ty == AGF_RADIO_EXTERNAL_RATES_CMD) {
        fs = AGF_FS_EXTERNAL_RATES_CONTROL;
      } else if (ty == AGF_RADIO_IDLE_CMD) {
        fs = AGF_FS_IDLE;
      }
    }
  }
  const uint32_t rflags = bget(s.bits, B_RFLAGS_SHIFT, B_RFLAGS_MASK);
  // --- UpdateWarnings :305-342 ---
  if (HK) {
    uint32_t w = (s.cnt >> 24) & 0xFFu;
    if (s.batt_vfilt <= k.batt_warning) w |= AGF_WARN_LOW_BATT;
    if (::fabsf(s.mon_cmd_lpdt - 0.02f) > (0.1f * 0.02f)) w |= AGF_WARN_CMD_RATE;
    if (float(s.age_radio) * 1e-6f > 3 * 0.02f) w |= AGF_WARN_CMD_BATCH_DROP;
    if (::fabsf(s.mon_loop_lpdt - k.onboard_period) > (0.05f * k.onboard_period)) w |= AGF_WARN_ONBOARD_FREQ;
    if (s.bits & B_KF_RESET_SEEN) s.age_est_reset = 0;
    if (float(s.age_est_reset) * 1e-6f < 0.02f) w |= AGF_WARN_UWB_RESET;
    s.cnt = (s.cnt & 0x00FFFFFFu) | (w << 24);
  }
  s.bits &= ~B_KF_RESET_SEEN;
  // --- CheckPanicReasons :344-391 ---
  {
    const bool running = (s.cmd[0] > 0) | (s.cmd[1] > 0) | (s.cmd[2] > 0) | (s.cmd[3] > 0);
    uint32_t unsafe = 0;
    if (running) {
      const bool nochk = (rflags & AGF_RADIO_FLAG_DISABLE_ONBOARD_SAFETY) != 0;
      if (s.kpos[2] < -2.0f && !nochk) unsafe = AGF_PANIC_ONBOARD_ESTIMATE_CRAZY;
      if (s.age_uwb > 1500u * 1000u && fs == AGF_FS_FULLY_AUTONOMOUS) unsafe = AGF_PANIC_UWB_TIMEOUT;
      const Q4<float> ea(s.katt[0], s.katt[1], s.katt[2], s.katt[3]);
      if (qrot_e3_z(ea) < 0 && !nochk) unsafe = AGF_PANIC_UPSIDE_DOWN;
      if (s.age_radio > 1500u * 1000u) unsafe = AGF_PANIC_RADIO_CMD_TIMEOUT;
      if (HK && s.batt_vfilt <= k.batt_critical) unsafe = AGF_PANIC_LOW_BATTERY;
    }
    const bool critical = !(fs == AGF_FS_UNINITIALIZED || fs == AGF_FS_IDLE || fs == AGF_FS_PANIC || fs == AGF_FS_KILLED);
    if (unsafe && critical) {
      fs = AGF_FS_PANIC;
      s.bits = bset(s.bits, B_PANIC_SHIFT, B_PANIC_MASK, unsafe);
    }
  }
  s.bits = bset(s.bits, B_FS_SHIFT, B_FS_MASK, fs);

  // --- controllers :194-217 ---  Each flight mode produces (total thrust [m/s^2], desired body rates); the
  // rate controller and mixer that all three modes end in (:451-455, :519-524, :547-551) are written once.
  const V3<float> estW(s.kw[0], s.kw[1], s.kw[2]);
  bool powered = false;
  float thrust = 0.0f;
  V3<float> desW(0.0f, 0.0f, 0.0f);
  if (fs == AGF_FS_EXTERNAL_RATES_CONTROL) {  // :528-588
    desW = V3<float>(s.radio_f[1], s.radio_f[2], s.radio_f[3]);
    thrust = s.radio_f[0];
    powered = true;
  } else if (fs == AGF_FS_FULLY_AUTONOMOUS) {  // :393-457
    const V3<float> estPos(s.kpos[0], s.kpos[1], s.kpos[2]), estVel(s.kvel[0], s.kvel[1], s.kvel[2]);
    const Q4<float> estAtt(s.katt[0], s.katt[1], s.katt[2], s.katt[3]);
    const V3<float> desPos(s.radio_f[0], s.radio_f[1], s.radio_f[2]);
    // GetDesAcceleration (QuadcopterPositionController.hpp:22-27), desVel = desAcc = 0
    const V3<float> zero(0, 0, 0);
    const V3<float> dv = zero - estVel;
    const V3<float> desAcc = (((desPos - estPos) * k.nat_freq) * k.nat_freq +
                              ((V3<float>(2 * dv.x, 2 * dv.y, 2 * dv.z) * k.nat_freq) * k.damping)) + zero;
    const V3<float> proper = desAcc + V3<float>(0, 0, 9.81f);
    const float nProper = norm(proper);
    const V3<float> dir = vdiv<PARITY>(proper, nProper);
    const float corr = qrot_e3_z(estAtt);
    const float corrSat = corr < 1.00f ? 1.00f : corr;
    thrust = fdiv<PARITY>(nProper, corrSat);
    const Q4<float> desAtt = att_from_thrust_dir<PARITY>(dir);
    desW = ctl_att<PARITY>(k, desAtt, estAtt);
    powered = true;
  } else if (fs == AGF_FS_EXTERNAL_ACCELERATION_CONTROL) {  // :459-526
    float4 r;
    if constexpr (PARITY) {
      r = ctl_accel_mode<true>(k, make_float4(s.katt[0], s.katt[1], s.katt[2], s.katt[3]),
                               make_float4(s.radio_f[0], s.radio_f[1], s.radio_f[2], s.radio_f[3]));
    } else {
      r = ctl_accel_mode_cold(&k, make_float4(s.katt[0], s.katt[1], s.katt[2], s.katt[3]),
                              make_float4(s.radio_f[0], s.radio_f[1], s.radio_f[2], s.radio_f[3]));
    }
    if (r.w >= 0.0f) {  // thrust is a norm: negative marks "motors off" (:466-470)
      desW = V3<float>(r.x, r.y, r.z);
      thrust = r.w;
      powered = true;
    }
  }
  if (powered) {
    ctl_mix(s, sc, k, thrust * k.mass, ctl_torques<PARITY>(k, desW, estW));
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      s.cmd[i] = 0;
      if constexpr (PARITY) s.dforce[i] = 0;
    }
  }
  if constexpr (HK) {  // propeller calibration, rates mode only (:553-587)
    if (fs == AGF_FS_EXTERNAL_RATES_CONTROL) {
      if constexpr (PARITY) {
        if (rflags & AGF_RADIO_FLAG_CALIBRATE_MOTORS) {
          if (!(s.bits & B_PC_RUNNING)) {
            s.bits |= B_PC_RUNNING;
            s.pc_count = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) s.pc_accum[i] = 0;
          }
#pragma unroll
          for (int i = 0; i < 4; i++) s.pc_accum[i] += k.mix_kf * s.cmd[i] * s.cmd[i];
          s.pc_count++;
        } else if (s.bits & B_PC_RUNNING) {
          s.bits &= ~B_PC_RUNNING;
          if (s.pc_count >= 750u) {
            const float truePer = k.mass * 9.81f / 4.0f;
#pragma unroll
            for (int i = 0; i < 4; i++) {
              float f = (s.pc_count * truePer) / s.pc_accum[i];
              const float fmin = 0.7f, fmax = 1.0f / fmin;
              if (f > fmax) f = fmax;
              if (f < fmin) f = fmin;
              s.pc_corr[i] = f;
            }
          }
        }
      } else if (AGF_UNLIKELY((rflags & AGF_RADIO_FLAG_CALIBRATE_MOTORS) || (s.bits & B_PC_RUNNING))) {
        // the same on the scratch copies (accumulators, correction and its derived 1 / (correction * kF))
        const int q0 = UWB ? SQ_PC_UWB : SQ_PC_NOUWB;
        const float4 a4 = sq_load(sc, q0 + SQ_PC_ACCUM);
        float acc[4] = {a4.x, a4.y, a4.z, a4.w};
        if (rflags & AGF_RADIO_FLAG_CALIBRATE_MOTORS) {
          if (!(s.bits & B_PC_RUNNING)) {
            s.bits |= B_PC_RUNNING;
            s.pc_count = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) acc[i] = 0;
          }
#pragma unroll
          for (int i = 0; i < 4; i++) acc[i] += k.mix_kf * s.cmd[i] * s.cmd[i];
          s.pc_count++;
          sq_store(sc, q0 + SQ_PC_ACCUM, make_float4(acc[0], acc[1], acc[2], acc[3]));
        } else {
          s.bits &= ~B_PC_RUNNING;
          if (s.pc_count >= 750u) {
            const float truePer = k.mass * 9.81f / 4.0f;
            float c[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
              float f = (s.pc_count * truePer) / acc[i];
              const float fmin = 0.7f, fmax = 1.0f / fmin;
              if (f > fmax) f = fmax;
              if (f < fmin) f = fmin;
              c[i] = f;
            }
            sq_store(sc, q0 + SQ_PC_CORR, make_float4(c[0], c[1], c[2], c[3]));
            sq_store(sc, q0 + SQ_PC_INV, make_float4(1.0f / (c[0] * k.mix_kf), 1.0f / (c[1] * k.mix_kf), 1.0f / (c[2] * k.mix_kf),
                                                     1.0f / (c[3] * k.mix_kf)));
          }
        }
      }
    }
  }
}

// SetRadioMessage (QuadcopterLogic.hpp:110-116)
template<typename P, bool PARITY, bool UWB, bool HK>
AGF_DEV void radio_deliver(VState<P, PARITY, UWB, HK>& s, const LogicParams& k, uint32_t type, uint32_t flags, const float f[4]) {
  s.bits |= B_RADIO_NEW;
  s.bits = bset(s.bits, B_RTYPE_SHIFT, B_RTYPE_MASK, type);
  s.bits = bset(s.bits, B_RFLAGS_SHIFT, B_RFLAGS_MASK, flags);
#pragma unroll
  for (int i = 0; i < 4; i++) s.radio_f[i] = f[i];
  s.age_radio = 0;
  if (HK) {
    const float mdt = float(s.age_mon_cmd) * 1e-6f;
    s.mon_cmd_lpdt = k.mon_cmd_coef <= 0.0f ? mdt : k.mon_cmd_coef * s.mon_cmd_lpdt + (1 - k.mon_cmd_coef) * mdt;
    s.age_mon_cmd -= uint32_t(mdt * 1e6f);
  }
}

AGF_DEV uint32_t sat_add(uint32_t a, uint32_t d) { return a > 0xF0000000u ? a : a + d; }
// the same stopwatch for the fast variants: add, then clamp -- two instructions; it differs from sat_add only in the value a
// saturated counter holds (exactly the cap instead of up to cap + d), far above every threshold the logic compares it with
AGF_DEV uint32_t sat_add_fast(uint32_t a, uint32_t d) { return min(a + d, 0xF0000000u); }

// ---------------------------------------------------------------------------------------------
// one tick: [radio delivery] -> Quadcopter_T::Run -> UWBNetwork::Run -> clock advance
// ---------------------------------------------------------------------------------------------
// inertia products for the two plant-parameter carriers: full 3x3 (shared parameters, read straight from the
// kernel-parameter constant bank) and diagonal (per-vehicle sweep, registers)
template<typename P> AGF_DEV V3<P> inertia_mul(const PlantPV<P>& pv, const V3<P>& w) { return matvec(pv.I, w); }
template<typename P> AGF_DEV V3<P> inertia_inv_mul(const PlantPV<P>& pv, const V3<P>& x) { return matvec(pv.Iinv, x); }
template<typename P> AGF_DEV V3<P> inertia_mul(const PlantPVDiag<P>& pv, const V3<P>& w) {
  return V3<P>(pv.Id[0] * w.x, pv.Id[1] * w.y, pv.Id[2] * w.z);
}
template<typename P> AGF_DEV V3<P> inertia_inv_mul(const PlantPVDiag<P>& pv, const V3<P>& x) {
  return V3<P>(pv.Iinvd[0] * x.x, pv.Iinvd[1] * x.y, pv.Iinvd[2] * x.z);
}

// OFFB: the in-kernel offboard loop (command delivery, mocap packets, command generation) is compiled in.  It is a
// template axis because its call sites cost the hot loop registers even when never taken (spills in every fast
// variant, round 1: FP32 full mode 1.77e10 -> 1.26e10 vehicle-steps/s with the sites merely present).
template<typename P, bool PARITY, bool UWB, bool HK, bool OFFB, typename PVT>
AGF_DEV void tick(VState<P, PARITY, UWB, HK>& s, const Scratch& sc, const StepShared<P>& p, const PVT& pv, const TickPlan& plan,
                  uint64_t now_us, uint32_t dt_us, uint64_t abs_tick, uint64_t gidx, size_t i, size_t n) {
  if (OFFB && plan.off_deliver) {  // CommunicationsDelay::GetMessage -> SetCommandRadioMsg (main.cpp:737-739)
    float4 c;
    if constexpr (PARITY) {
      c = make_float4(s.offq[4 * plan.off_deliver_slot], s.offq[4 * plan.off_deliver_slot + 1], s.offq[4 * plan.off_deliver_slot + 2],
                      s.offq[4 * plan.off_deliver_slot + 3]);
    } else {
      c = sq_load(sc, (UWB ? SQ_QUADS_UWB : SQ_QUADS_NOUWB) + int(plan.off_deliver_slot));
    }
    if (c.x != c.x || c.x < -3.0e38f) {  // idle / kill command (CreateIdleCommand, CreateKillCommand): type and flags only
      const float cf[4] = {s.radio_f[0], s.radio_f[1], s.radio_f[2], s.radio_f[3]};
      radio_deliver(s, p.logic, c.x != c.x ? AGF_RADIO_IDLE_CMD : AGF_RADIO_EMERGENCY_KILL, p.off.flags, cf);
    } else if (c.x <= 3.0e38f) {  // +inf: nothing was sent that round
      const float cf[4] = {c.x, c.y, c.z, c.w};
      radio_deliver(s, p.logic, AGF_RADIO_EXTERNAL_RATES_CMD, p.off.flags, cf);
    }
  }
  if (plan.run_plant) {
    const P dt = sizeof(P) == 4 ? P(plan.plant_dt_f32) : P(double(plan.plant_dt_us) * 1e-6);
    const P inv_dt = PARITY ? P(0) : P(1) / dt;
    // ---- motors (Motor.cpp:39-84).  Axes are (0,0,+-1) and thrust is (0,0,f): the products with
    // the exact-zero components are dropped, the rest keeps the reference's order.
    P Fz = P(0), Tx = P(0), Ty = P(0), Tz = P(0);
    const P c = pv.motor_c;
#pragma unroll
    for (int m = 0; m < 4; m++) {
      const P sgn = (m & 1) ? P(-1) : P(1);  // spin axis z component
      P cmd = P(s.cmd[m]);
      if (cmd < 0) cmd = 0;
      const P old = s.ms[m];
      P sp = c * old + (1 - c) * cmd;
      if constexpr (PARITY) {
        if (sp > p.motor_max) {
          sp = p.motor_max;
        } else if (sp < p.motor_min) {
          sp = p.motor_min;
        }
      } else {
        sp = pmin_(pmax_(sp, p.motor_min), p.motor_max);
      }
      s.ms[m] = sp;
      const P fz = (pv.kF * sp) * rabs_(sp);
      const P aero = (((-pv.kTau) * sp) * rabs_(sp)) * sgn;
      P tz;
      if (PARITY) {
        const P angAcc = (sp - old) / dt;
        tz = (aero + P(0)) - ((angAcc * p.motor_J) * sgn);
      } else {
        tz = aero - (((sp - old) * inv_dt) * p.motor_J) * sgn;
      }
      const P tx = p.motor_pos[m][1] * fz;   // (r x T).x = ry*fz - rz*0
      const P ty = P(0) - p.motor_pos[m][0] * fz;  // (r x T).y = rz*0 - rx*fz
      Fz = Fz + fz; Tx = Tx + tx; Ty = Ty + ty; Tz = Tz + tz;
    }
    Q4<P> att(s.att[0], s.att[1], s.att[2], s.att[3]);
    V3<P> w(s.w[0], s.w[1], s.w[2]);
    V3<P> F(P(0), P(0), Fz), T(Tx, Ty, Tz);
    V3<P> extF(P(0), P(0), P(0));
    if (p.ext_force) {
      extF = V3<P>(p.ext_force[i], p.ext_force[n + i], p.ext_force[2 * n + i]);
      const V3<P> extT(p.ext_torque[i], p.ext_torque[n + i], p.ext_torque[2 * n + i]);
      T = T + qrot(qinv(att), extT);
    }
    // angular momentum: I*w + sum of rotor momenta (Quadcopter_T.cpp:113-117)
    V3<P> L = inertia_mul(pv, w);
#pragma unroll
    for (int m = 0; m < 4; m++) {
      const P sgn = (m & 1) ? P(-1) : P(1);
      L.z = L.z + (s.ms[m] * p.motor_J) * sgn;
    }
    const V3<P> angAcc = inertia_inv_mul(pv, T - cross(w, L));
    if (p.has_drag) {  // :123-128
      const V3<P> vb = qrot(qinv(att), V3<P>(s.vel[0], s.vel[1], s.vel[2]));
      F = F + V3<P>(p.drag[0] * (-vb.x), p.drag[1] * (-vb.y), p.drag[2] * (-vb.z));
    }
    V3<P> acc(P(0), P(0), P(-9.81));
    if (PARITY) {
      acc = acc + vdiv<true>(qrot(att, F) + extF, pv.mass);
    } else if (p.has_drag) {
      acc = acc + (qrot(att, F) + extF) * pv.inv_mass;
    } else {  // F = (0, 0, Fz): only the third column of the rotation matrix is needed
      const V3<P> c3(2 * att.x * att.z + 2 * att.w * att.y, 2 * att.y * att.z - 2 * att.w * att.x,
                     att.w * att.w - att.x * att.x - att.y * att.y + att.z * att.z);
      acc = acc + (c3 * F.z + extF) * pv.inv_mass;
    }
    const V3<P> pos(s.pos[0], s.pos[1], s.pos[2]), vel(s.vel[0], s.vel[1], s.vel[2]);
    V3<P> npos, nvel;
    Q4<P> natt = q_apply_rotvec<PARITY>(att, w * dt);
    if constexpr (!PARITY && sizeof(P) == 4) {
      // FP32 mode is not the reference's arithmetic; what it owes the reference is its trajectory (north star: 1e-4 over
      // 10 s).  Two places where plain FP32 loses that over 5 000 ticks, both measured against the reference population
      // (tests/test_fast_population_gpu.py, rates mode: worst vehicle 1.7e-4 without, < 5e-5 with):
      //  * the position and velocity sums: millimetre increments into metre-sized accumulators, thousands of times ->
      //    compensated (Kahan) summation, the running compensation is part of the stored state;
      //  * the attitude quaternion: the reference never normalises it, but in double it stays a unit quaternion to
      //    1e-16, while a float one is off by ~6e-8 from the first conversion on and scales the thrust it rotates
      //    -> one Newton step of 1/|q| per tick keeps |q| = 1 to second order.
      const P pp[3] = {pos.x, pos.y, pos.z}, vv[3] = {vel.x, vel.y, vel.z}, aa[3] = {acc.x, acc.y, acc.z};
      P np_[3], nv_[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const P inc = vv[k] * dt + ((P(0.5) * aa[k]) * dt) * dt;
        const P y = inc - s.cpos[k];
        const P t = pp[k] + y;
        s.cpos[k] = (t - pp[k]) - y;
        np_[k] = t;
        const P yv = aa[k] * dt - s.cvel[k];
        const P tv = vv[k] + yv;
        s.cvel[k] = (tv - vv[k]) - yv;
        nv_[k] = tv;
      }
      npos = V3<P>(np_[0], np_[1], np_[2]);
      nvel = V3<P>(nv_[0], nv_[1], nv_[2]);
      const P n2 = natt.w * natt.w + natt.x * natt.x + natt.y * natt.y + natt.z * natt.z;
      const P kn = P(1.5) - P(0.5) * n2;
      natt = Q4<P>(natt.w * kn, natt.x * kn, natt.y * kn, natt.z * kn);
    } else {
      npos = (pos + vel * dt) + ((P(0.5) * acc) * dt) * dt;
      nvel = vel + acc * dt;
    }
    V3<P> nw = w + angAcc * dt;
    if ((npos.z <= 0) && (nvel.z < 0)) {  // ground contact :146-151
      npos.z = 0;
      nvel.z = 0;
      acc.z = 0;
      nw = V3<P>(P(0), P(0), P(0));
      if constexpr (!PARITY && sizeof(P) == 4) { s.cpos[2] = 0; s.cvel[2] = 0; }
    }
    s.pos[0] = npos.x; s.pos[1] = npos.y; s.pos[2] = npos.z;
    s.vel[0] = nvel.x; s.vel[1] = nvel.y; s.vel[2] = nvel.z;
    s.att[0] = natt.w; s.att[1] = natt.x; s.att[2] = natt.y; s.att[3] = natt.z;
    s.w[0] = nw.x; s.w[1] = nw.y; s.w[2] = nw.z;

    if (plan.run_logic) {  // :159-199
      V3<float> g(float(nw.x), float(nw.y), float(nw.z));
      const V3<P> sf = qrot(qinv(natt), acc + V3<P>(P(0), P(0), P(9.81)));
      V3<float> a(float(sf.x), float(sf.y), float(sf.z));
      if (!p.logic.imu_identity) {
        g = matvec(p.logic.R_imu_inv, g);
        a = matvec(p.logic.R_imu_inv, a);
      }
      if (p.noise_on) {
        float nrm[6];
        normals6(p.philox_rk, gidx, s.cycle, 0u, nrm);
        g = g + V3<float>(nrm[0], nrm[1], nrm[2]) * p.sigma_gyro;
        a = a + V3<float>(nrm[3], nrm[4], nrm[5]) * p.sigma_acc;
        if (AGF_UNLIKELY(p.bias_on)) {  // constant per vehicle: counter (vehicle, 0xFFFFFFFF, 1)
          const Normals6 b = normals6_cold(p.seed, gidx, 0xFFFFFFFFu, 1u);
          g = g + V3<float>(b.n[0], b.n[1], b.n[2]) * p.bias_sigma_gyro;
          a = a + V3<float>(b.n[3], b.n[4], b.n[5]) * p.bias_sigma_acc;
        }
      }
      logic_run<PARITY>(s, sc, p, g, a, plan.kf_dt_f32);
      if constexpr (UWB) {  // radio exchange :191-199
        s.rpos[0] = s.pos[0]; s.rpos[1] = s.pos[1]; s.rpos[2] = s.pos[2];
        if (s.bits & B_RADIO_MEAS_NEW) {  // SetUWBMeasurement (QuadcopterLogic.hpp:61-69)
          s.bits &= ~B_RADIO_MEAS_NEW;
          s.age_uwb = 0;
          s.bits |= B_UWB_NEW;
          s.logic_range = s.uwb_range;
          const uint32_t rr = (s.uwbw >> W_RADIO_RESP) & 0xFFu;
          s.uwbw = (s.uwbw & ~(0xFFu << W_LOGIC_TARGET)) | (rr << W_LOGIC_TARGET);
        }
      }
    }
  }
  // ---- UWBNetwork::Run (UWBNetwork.cpp:22-89), private network: this vehicle + the anchors.
  // When it runs, and whether it starts or completes a transaction, depends on the clock only
  // (TickPlan); which anchor answers and the measured range are per vehicle.
  if constexpr (UWB) {
    if (plan.net_start) {  // :32-41 responder = the requester's next ranging target
      const uint32_t nxt = (s.uwbw >> W_NEXT_TARGET) & 0xFFu;
      s.uwbw = (s.uwbw & ~(0xFFu << W_NET_RESP)) | (nxt << W_NET_RESP);
    }
    if (plan.net_complete) {  // :47-82
      const uint32_t resp = (s.uwbw >> W_NET_RESP) & 0xFFu;
      const V3<P> d = V3<P>(s.rpos[0], s.rpos[1], s.rpos[2]) -
                      V3<P>(P(p.anchors[resp].x), P(p.anchors[resp].y), P(p.anchors[resp].z));
      P r = norm(d);
      if (AGF_UNLIKELY(p.uwb_noise_on)) {  // :62-72
        const UwbDraw dr = uwb_draw_cold(p.seed, gidx, uint32_t(abs_tick));
        if (dr.u < p.uwb_outlier_prob) {
          r = P(dr.n_outlier) * P(p.uwb_outlier_sigma);
        } else {
          r = r + P(dr.n_noise) * P(p.uwb_sigma);
        }
      } else {
        r = r + P(0);
      }
      s.uwb_range = float(r);
      s.uwbw = (s.uwbw & ~(0xFFu << W_RADIO_RESP)) | (resp << W_RADIO_RESP);
      s.bits |= B_RADIO_MEAS_NEW;
    }
  }
  if (OFFB && plan.mocap_update) {  // simulated mocap packet (main.cpp:451-457): the true pose after this tick's Run()
    const V3<P> mp(s.pos[0], s.pos[1], s.pos[2]);
    const Q4<P> ma(s.att[0], s.att[1], s.att[2], s.att[3]);
    if constexpr (PARITY) {
      mocap_update<true, double>(p.off.est, i, now_us + dt_us, V3<double>(double(mp.x), double(mp.y), double(mp.z)),
                         Q4<double>(double(ma.w), double(ma.x), double(ma.y), double(ma.z)));
    } else {
      mocap_update_cold<P>(&p.off.est, i, now_us + dt_us, mp, ma);
    }
  }
  if (OFFB && plan.off_generate) {  // offboard main loop (main.cpp:471-673), after the clock advance
    const uint64_t t_gen = now_us + dt_us;
    const V3<P> cp(s.pos[0], s.pos[1], s.pos[2]), cv(s.vel[0], s.vel[1], s.vel[2]);
    const Q4<P> ca(s.att[0], s.att[1], s.att[2], s.att[3]);
    if constexpr (PARITY) {
      const float4 c = offboard_generate<true, P>(p.off, i, n, t_gen, cp, cv, ca);
      s.offq[4 * plan.off_gen_slot] = c.x; s.offq[4 * plan.off_gen_slot + 1] = c.y;
      s.offq[4 * plan.off_gen_slot + 2] = c.z; s.offq[4 * plan.off_gen_slot + 3] = c.w;
    } else {
      const float4 c = offboard_generate_cold<P>(&p.off, i, n, t_gen, cp, cv, ca);
      sq_store(sc, (UWB ? SQ_QUADS_UWB : SQ_QUADS_NOUWB) + int(plan.off_gen_slot), c);
    }
  }
  if constexpr (PARITY) {
    s.age_radio = sat_add(s.age_radio, dt_us);
    s.age_uwb = sat_add(s.age_uwb, dt_us);
    if (HK) {
      s.age_est_reset = sat_add(s.age_est_reset, dt_us);
      s.age_mon_cmd = sat_add(s.age_mon_cmd, dt_us);
      s.age_mon_loop = sat_add(s.age_mon_loop, dt_us);
    }
  } else {
    s.age_radio = sat_add_fast(s.age_radio, dt_us);
    s.age_uwb = sat_add_fast(s.age_uwb, dt_us);
    if (HK) {
      s.age_est_reset = sat_add_fast(s.age_est_reset, dt_us);
      s.age_mon_cmd = sat_add_fast(s.age_mon_cmd, dt_us);
      s.age_mon_loop = sat_add_fast(s.age_mon_loop, dt_us);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// K1/K2/K3: the step kernel.  One vehicle per thread; many ticks per launch with the state in
// registers; optional trajectory log (a ring of records of 16-byte vectors [quad][vehicle], StepLaunch::log).
//
// Balanced schedule.  Every vehicle-tick costs the same, so a launch is nblocks x nticks equal units of
// work ("block-ticks").  A plain grid of nblocks CTAs runs them in ceil(nblocks / resident CTAs) waves and
// the last wave is partly empty (131072 vehicles: 1024 blocks over 592 resident CTAs = 2 waves for 1.73
// waves of work).  Instead the launch is one wave of G co-resident CTAs and CTA j owns the contiguous range
// [j*T/G, (j+1)*T/G) of the block-major sequence of T = nblocks*nticks block-ticks.  A range boundary that
// falls inside block b splits b's ticks between CTA j (ticks [0,t)) and CTA j+1 (ticks [t,nticks)): CTA j
// runs that head part FIRST, stores the state and publishes flags[b] = epoch; CTA j+1 runs its tail part
// LAST, after acquiring the flag.  Every CTA therefore finishes at the same time and nobody waits in practice.
// ---------------------------------------------------------------------------------------------
template<typename P>
AGF_DEV void plant_params_load(PlantPV<P>& pv, const StepLaunch<P>& L, size_t i) {
  pv = L.pv_shared;
  if (L.pv) {
    constexpr int VP = VecOf<P>::lanes;
    P r[NPV_PAD];
#pragma unroll
    for (int q = 0; q < NPV_PAD / VP; q++) VecOf<P>::unpack(L.pv[size_t(q) * L.n + i], &r[q * VP]);
    pv.mass = r[PV_MASS];
    pv.inv_mass = r[PV_INV_MASS];
#pragma unroll
    for (int k = 0; k < 9; k++) { pv.I[k] = P(0); pv.Iinv[k] = P(0); }
    pv.I[0] = r[PV_IXX]; pv.I[4] = r[PV_IYY]; pv.I[8] = r[PV_IZZ];
    pv.Iinv[0] = r[PV_IIXX]; pv.Iinv[4] = r[PV_IIYY]; pv.Iinv[8] = r[PV_IIZZ];
    pv.kF = r[PV_KF]; pv.kTau = r[PV_KTAU]; pv.motor_c = r[PV_MOTOR_C];
  }
}
template<typename P>
AGF_DEV void plant_params_load(PlantPVDiag<P>& pv, const StepLaunch<P>& L, size_t i) {
  constexpr int VP = VecOf<P>::lanes;
  P r[NPV_PAD];
#pragma unroll
  for (int q = 0; q < NPV_PAD / VP; q++) VecOf<P>::unpack(L.pv[size_t(q) * L.n + i], &r[q * VP]);
  pv.mass = r[PV_MASS];
  pv.inv_mass = r[PV_INV_MASS];
  pv.Id[0] = r[PV_IXX]; pv.Id[1] = r[PV_IYY]; pv.Id[2] = r[PV_IZZ];
  pv.Iinvd[0] = r[PV_IIXX]; pv.Iinvd[1] = r[PV_IIYY]; pv.Iinvd[2] = r[PV_IIZZ];
  pv.kF = r[PV_KF]; pv.kTau = r[PV_KTAU]; pv.motor_c = r[PV_MOTOR_C];
}

// Occupancy targets (blocks of AGF_BLOCK_THREADS per SM) of the fast variants; the parity variants
// take whatever registers they need.  Tuned with -Xptxas -v and ncu, see DESIGN.md.
#ifndef AGF_MINB_F32_UWB
#define AGF_MINB_F32_UWB 3
#endif
#ifndef AGF_MINB_F32_RATES
#define AGF_MINB_F32_RATES 4
#endif
#ifndef AGF_MINB_F64_UWB
#define AGF_MINB_F64_UWB 3
#endif
#ifndef AGF_MINB_F64_RATES
#define AGF_MINB_F64_RATES 4
#endif
// blocks per SM fewer for the kernels with the offboard loop compiled in (measured after the tick-plan table: truth-fed loop
// 2.13e10 with one fewer vs 1.85e10 with the same number; with the mocap estimator 7.7e9 vs 8.0e9 -- profiles/r1/)
#ifndef AGF_OFFB_MINB_DELTA
#define AGF_OFFB_MINB_DELTA 1
#endif
template<typename P, bool PARITY, bool UWB, bool OFFB = false>
constexpr int step_min_blocks() {
  return PARITY ? (UWB ? 1 : 3)  // parity full mode wants its 255 registers (3.7e9 vs 2.8e9 at 168); parity rates 168 (7.2e9 vs 6.5e9)
                : (sizeof(P) == 4 ? (UWB ? AGF_MINB_F32_UWB : AGF_MINB_F32_RATES) : (UWB ? AGF_MINB_F64_UWB : AGF_MINB_F64_RATES)) -
                      (OFFB ? AGF_OFFB_MINB_DELTA : 0);
}
template<bool PARITY, bool UWB>
constexpr size_t step_smem_bytes(int block, bool offboard) {
  return PARITY ? 0 : size_t(block) * ((UWB ? SQ_QUADS_UWB : SQ_QUADS_NOUWB) + (offboard ? AGF_OFFQ : 0)) * sizeof(float4);
}

// the ticks [t0, t1) of vehicle i; pv is the constant-bank struct or a register copy
template<typename P, bool PARITY, bool UWB, bool HK, bool OFFB, typename PVT>
AGF_DEV void step_ticks(const StepLaunch<P>& L, const PVT& pv, const Scratch& sc, size_t i, uint32_t t0, uint32_t t1) {
  VState<P, PARITY, UWB, HK> s;
  state_load(s, L.st, L.n, i, sc, L.sh.logic.mix_kf);
  // what each tick does (plant step, logic, ranging, offboard loop) depends on the clock only: the host evaluated the
  // stopwatch recurrence (agf_types.h timing_plan / timing_advance) for every tick of the launch; one 16-byte word per tick,
  // the same address for the whole grid, the next one requested a tick ahead
  uint4 pp = ldro_(L.plans + t0);
  // next scheduled radio delivery, as a tick offset into this launch (0xFFFFFFFF: none left)
  uint32_t si = L.sched_begin;
  while (si < L.sched_end && L.sched[si].tick < L.tick0 + t0) si++;
  uint32_t next_cmd = si < L.sched_end ? uint32_t(L.sched[si].tick - L.tick0) : 0xFFFFFFFFu;
  // trajectory log: ticks until the next record, this vehicle's place in that record, records left before the ring wraps
  // (all the index arithmetic of the ring happens here, once per work item; the loop only adds)
  constexpr int VP = VecOf<P>::lanes, LOGQ = (AGF_LOG_FIELDS - 1) / VP;
  typedef typename VecOf<P>::type LogVec;
  uint32_t log_in = 0, log_left = 0;
  LogVec* logp = nullptr;
  if (L.log) {
    log_in = L.log_stride - uint32_t((L.tick0 + t0) % L.log_stride);
    const uint32_t slot = (L.log_slot0 + (t0 + log_in - 1 - L.log_first_off) / L.log_stride) % L.log_capacity;
    logp = reinterpret_cast<LogVec*>(L.log + size_t(slot) * AGF_LOG_FIELDS * L.log_n) + i;
    log_left = L.log_capacity - slot;
  }
  const uint64_t gidx = L.first_global_index + i;
  for (uint32_t t = t0; t < t1; t++) {
    const uint64_t abs_tick = L.tick0 + t;
    // radio delivery before Run() (main.cpp:737-739 of the previous loop iteration)
    if (t == next_cmd) {
      const SchedEntryDev& e = L.sched[si];
      if (e.slot < 0) {
        radio_deliver(s, L.sh.logic, e.type, e.flags, e.f);
      } else {
        const float4 f = L.slots[e.slot].f[i];
        const uint32_t tf = L.slots[e.slot].tf[i];
        const float ff[4] = {f.x, f.y, f.z, f.w};
        radio_deliver(s, L.sh.logic, tf & 0xFFu, (tf >> 8) & 0xFFu, ff);
      }
      si++;
      next_cmd = si < L.sched_end ? uint32_t(L.sched[si].tick - L.tick0) : 0xFFFFFFFFu;
    }
    const TickPlan plan = unpack_plan(pp.x, pp.y, bits_float(pp.z), bits_float(pp.w));
    if (t + 1 < t1) pp = ldro_(L.plans + t + 1);
    tick<P, PARITY, UWB, HK, OFFB>(s, sc, L.sh, pv, plan, L.now0_us + uint64_t(t) * L.dt_us, L.dt_us, abs_tick, gidx, i, L.n);
    if (L.log && --log_in == 0) {
      log_in = L.log_stride;
      // record = LOGQ vectors [quad][vehicle] (16-byte stores, a warp writes 512 contiguous bytes) + the 17th value [vehicle]
      const P f[AGF_LOG_FIELDS - 1] = {s.pos[0], s.pos[1], s.pos[2], s.vel[0], s.vel[1], s.vel[2], s.att[0], s.att[1],
                                       s.att[2], s.att[3], s.w[0],   s.w[1],   s.w[2],   s.ms[0],  s.ms[1],  s.ms[2]};
#pragma unroll
      for (int q = 0; q < LOGQ; q++) logp[size_t(q) * L.log_n] = VecOf<P>::pack(&f[q * VP]);
      reinterpret_cast<P*>(logp + size_t(LOGQ) * L.log_n - i)[i] = s.ms[3];
      logp += (size_t(AGF_LOG_FIELDS) * L.log_n) / VP;  // log_n: the vehicle count padded to whole lines
      if (--log_left == 0) {
        log_left = L.log_capacity;
        logp -= (size_t(L.log_capacity) * AGF_LOG_FIELDS * L.log_n) / VP;
      }
    }
  }
  state_store(s, L.st, L.n, i, sc, L.sh.logic.mix_kf);
}

// one work item: ticks [t0, t1) of vehicle block b
template<typename P, bool PARITY, bool UWB, bool HK, bool PV, bool OFFB>
AGF_DEV void step_item(const StepLaunch<P>& L, const Scratch& sc, uint32_t b, uint32_t t0, uint32_t t1) {
  const size_t i = size_t(b) * blockDim.x + threadIdx.x;
  if (i >= L.n) return;
  if constexpr (PARITY) {  // generic carrier, shared or per-vehicle decided at run time
    PlantPV<P> pv;
    plant_params_load(pv, L, i);
    step_ticks<P, PARITY, UWB, HK, OFFB>(L, pv, sc, i, t0, t1);
  } else if constexpr (PV) {
    PlantPVDiag<P> pv;
    plant_params_load(pv, L, i);
    step_ticks<P, PARITY, UWB, HK, OFFB>(L, pv, sc, i, t0, t1);
  } else {
    step_ticks<P, PARITY, UWB, HK, OFFB>(L, L.pv_shared, sc, i, t0, t1);
  }
}

AGF_DEV void flag_publish(uint32_t* flag, uint32_t epoch) {
#if defined(__CUDA_ARCH__)
  __threadfence();   // this thread's state stores are visible device-wide ...
  __syncthreads();   // ... for every thread of the CTA, before one thread raises the flag
  if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
#endif
}
AGF_DEV void flag_wait(const uint32_t* flag, uint32_t epoch) {
#if defined(__CUDA_ARCH__)
  if (threadIdx.x == 0) {
    uint32_t v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      if (v != epoch) __nanosleep(200);
    } while (v != epoch);
  }
  __syncthreads();
#endif
}

// AGF_MAXNREG (tuning builds only): cap the registers per thread directly instead of through a blocks-per-SM target
#ifdef AGF_MAXNREG
#define AGF_STEP_BOUNDS __maxnreg__(AGF_MAXNREG)
#else
#define AGF_STEP_BOUNDS __launch_bounds__(AGF_BLOCK_THREADS, step_min_blocks<P, PARITY, UWB, OFFB>())
#endif
template<typename P, bool PARITY, bool UWB, bool HK, bool PV, bool OFFB>
__global__ void AGF_STEP_BOUNDS
step_kernel(const __grid_constant__ StepLaunch<P> L) {
  extern __shared__ float4 agf_scratch[];
  Scratch sc;
  sc.q = agf_scratch + threadIdx.x;
  // work items of this CTA, in order: [head part of block b_hi] [whole blocks] [tail part of block b_lo];
  // a single call site keeps one copy of the step in the kernel image
  uint32_t b_lo = blockIdx.x, t_lo = 0, b_hi = blockIdx.x + 1, t_hi = 0;
  if (L.balanced) {
    const uint64_t total = uint64_t(L.nblocks) * L.nticks;
    const uint64_t lo = (uint64_t(blockIdx.x) * total) / gridDim.x, hi = (uint64_t(blockIdx.x + 1) * total) / gridDim.x;
    b_lo = uint32_t(lo / L.nticks); t_lo = uint32_t(lo % L.nticks);
    b_hi = uint32_t(hi / L.nticks); t_hi = uint32_t(hi % L.nticks);
  }
  const uint32_t first_whole = t_lo ? b_lo + 1 : b_lo;
  const uint32_t n_head = t_hi ? 1u : 0u, n_whole = b_hi - first_whole, n_items = n_head + n_whole + (t_lo ? 1u : 0u);
  for (uint32_t k = 0; k < n_items; k++) {
    uint32_t b, t0 = 0, t1 = L.nticks;
    const bool head = k < n_head, tail = k >= n_head + n_whole;
    if (head) {  // first, then publish
      b = b_hi;
      t1 = t_hi;
    } else if (tail) {  // last, after its head part was stored by the neighbouring CTA
      b = b_lo;
      t0 = t_lo;
      flag_wait(L.flags + b_lo, L.epoch);
    } else {
      b = first_whole + (k - n_head);
    }
    step_item<P, PARITY, UWB, HK, PV, OFFB>(L, sc, b, t0, t1);
    if (head) flag_publish(L.flags + b_hi, L.epoch);
  }
}

// Host side of the schedule: one wave of co-resident CTAs when that is fewer than the vehicle blocks.  A CTA waits, at
// the end of its work, for what its lower-indexed neighbour does first; forward progress therefore needs the whole grid
// resident, which only a COOPERATIVE launch guarantees (cudaLaunchCooperativeKernel refuses a grid that does not fit and
// holds the launch back until it does, whatever else shares the GPU).  AGF_NO_BALANCED=1 in the environment, or a device
// without cooperative launch, falls back to the plain one-CTA-per-block grid.
template<typename P, typename K>
static cudaError_t launch_step_kernel(K kernel, StepLaunch<P> L, int block, size_t smem, cudaStream_t stream) {
  struct Cached { const void* fn; int block; size_t smem; int dev; int resident; int sms; int coop; };
  static Cached cache[32];
  static int ncached = 0;
  static std::mutex cache_mutex;  // one host thread per GPU may launch concurrently (agrifly_b200.h "Threading")
  static const bool no_balanced = [] { const char* e = getenv("AGF_NO_BALANCED"); return e && e[0] && e[0] != '0'; }();
  L.nblocks = uint32_t((L.n + block - 1) / block);
  L.balanced = 0;
  unsigned grid = L.nblocks;
  if (L.flags && L.nticks > 1 && !no_balanced) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int resident = -1, sms = 0, coop = 0;
    {
      std::lock_guard<std::mutex> lock(cache_mutex);
      for (int k = 0; k < ncached; k++)
        if (cache[k].fn == (const void*)kernel && cache[k].block == block && cache[k].smem == smem && cache[k].dev == dev) {
          resident = cache[k].resident;
          sms = cache[k].sms;
          coop = cache[k].coop;
        }
      if (resident < 0) {
        int per_sm = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem);
        if (e != cudaSuccess) return e;
        resident = sms * per_sm;
        if (ncached < 32) cache[ncached++] = Cached{(const void*)kernel, block, smem, dev, resident, sms, coop};
      }
    }
    if (coop && resident > 0 && L.nblocks > unsigned(resident)) {
      L.balanced = 1;
      grid = unsigned(resident);
    } else if (coop && sms > 0 && L.nblocks > unsigned(sms) && L.nblocks % unsigned(sms) != 0) {
      // fewer blocks than fit, but not a multiple of the SM count: the same number of CTAs on every SM, the block-ticks
      // dealt out evenly, instead of some SMs carrying one block more than the others for the whole launch
      L.balanced = 1;
      grid = (L.nblocks / unsigned(sms)) * unsigned(sms);
    }
  }
  if (L.balanced) {
    void* args[] = {(void*)&L};
    return cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(block), args, smem, stream);
  }
  kernel<<<grid, block, smem, stream>>>(L);
  return cudaGetLastError();
}

}  // namespace agf
