// agf_math.h -- deterministic elementary functions shared by host and device.
//
// Why this exists (hard part 1; DESIGN.md "Parity definition"):
// the reference's closed loop is sensitive to last-bit differences of the float
// libm calls made by its onboard logic (Common/Common/Math/Rotation.hpp:262-316:
// sinf cosf asinf acosf atan2f) and, much more weakly, of the double sin/cos of
// the plant's attitude update (Rotation.hpp:291-296).  glibc's and CUDA's libm
// differ in the last bit, so bit-level parity between a CPU run and a GPU run is
// only possible if both sides call the *same* routines.  Every routine below is
// built only from IEEE-754 binary64 add, subtract, multiply, divide and sqrt in
// a fixed order (no fused multiply-add, no vendor libm inside), so gcc on x86-64
// and nvcc on sm_100a produce bit-identical results by construction:
//   * on the device each operation is an explicit round-to-nearest intrinsic
//     (__dadd_rn, __dmul_rn, ...), which the compiler never contracts into FMA;
//   * on the host the file must be compiled with -ffp-contract=off (the oracle
//     Makefile and the test fixtures do so; x86-64 baseline has no FMA anyway).
// The float entry points evaluate in binary64 and round once to binary32.
//
// The polynomial / rational coefficients are the classical fdlibm ones
// (Sun Microsystems, freely redistributable); accuracy is < 1 ulp (binary64)
// for |x| < 1e5 and is checked against glibc in tests/test_agf_math.py.
//
// Users:  the CUDA step kernels (parity variant), the oracle port (oracle/port)
// and the "sharedmath" build of the unmodified reference (oracle/Makefile),
// which redirects the reference's libm calls here with a forced include.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define AGF_HD __host__ __device__ __forceinline__
#else
#define AGF_HD static inline
#endif
// Entry points (sin, cos, asin, acos, atan2, exp, cbrt).  AGF_MATH_OUTLINE (defined by the parity kernels' translation
// unit) makes them real device calls: the same arithmetic, one copy each instead of one per call site -- the parity step
// kernel was 32 000 instructions (0.5 MB) with them inlined and spent most of its time waiting for instruction fetches.
#if defined(__CUDACC__) && defined(AGF_MATH_OUTLINE)
#define AGF_HDX static __host__ __device__ __noinline__
#else
#define AGF_HDX AGF_HD
#endif

#if defined(__CUDA_ARCH__)
#define AGF_DADD(a, b) __dadd_rn((a), (b))
#define AGF_DSUB(a, b) __dsub_rn((a), (b))
#define AGF_DMUL(a, b) __dmul_rn((a), (b))
#define AGF_DDIV(a, b) __ddiv_rn((a), (b))
#define AGF_DSQRT(a) __dsqrt_rn((a))
#else
#define AGF_DADD(a, b) ((a) + (b))
#define AGF_DSUB(a, b) ((a) - (b))
#define AGF_DMUL(a, b) ((a) * (b))
#define AGF_DDIV(a, b) ((a) / (b))
#define AGF_DSQRT(a) sqrt((a))
#endif

// ---------------------------------------------------------------------------
// argument reduction: x = k*(pi/2) + r, |r| <= pi/4 (+ a hair), returns k mod 4
// three-term Cody-Waite; exact products for |k| < 2^20
// ---------------------------------------------------------------------------
AGF_HD int agf_rem_pio2(double x, double* r) {
  const double invpio2 = 6.36619772367581382433e-01;
  const double p1 = 1.57079632673412561417e+00;   // first 33 bits of pi/2
  const double p2 = 6.07710050630396597660e-11;   // next 33 bits
  const double p3 = 2.02226624871116645580e-21;   // next 33 bits
  const double p3t = 8.47842766036889956997e-32;  // tail
  const double magic = 6755399441055744.0;        // 1.5 * 2^52: rounds to nearest integer
  double t = AGF_DADD(AGF_DMUL(x, invpio2), magic);
  double fk = AGF_DSUB(t, magic);
  double y = AGF_DSUB(x, AGF_DMUL(fk, p1));
  y = AGF_DSUB(y, AGF_DMUL(fk, p2));
  y = AGF_DSUB(y, AGF_DMUL(fk, p3));
  y = AGF_DSUB(y, AGF_DMUL(fk, p3t));
  *r = y;
  long long k = (long long)fk;
  return (int)(k & 3);
}

// sin on |r| <= pi/4: r + r^3*(S1 + r^2*(S2 + ...))
AGF_HD double agf_ksin(double r) {
  const double S1 = -1.66666666666666324348e-01;
  const double S2 = 8.33333333332248946124e-03;
  const double S3 = -1.98412698298579493134e-04;
  const double S4 = 2.75573137070700676789e-06;
  const double S5 = -2.50507602534068634195e-08;
  const double S6 = 1.58969099521155010221e-10;
  double z = AGF_DMUL(r, r);
  double p = AGF_DADD(S5, AGF_DMUL(z, S6));
  p = AGF_DADD(S4, AGF_DMUL(z, p));
  p = AGF_DADD(S3, AGF_DMUL(z, p));
  p = AGF_DADD(S2, AGF_DMUL(z, p));
  p = AGF_DADD(S1, AGF_DMUL(z, p));
  double r3 = AGF_DMUL(z, r);
  return AGF_DADD(r, AGF_DMUL(r3, p));
}

// cos on |r| <= pi/4: 1 - r^2/2 + r^4*(C1 + r^2*(C2 + ...))
AGF_HD double agf_kcos(double r) {
  const double C1 = 4.16666666666666019037e-02;
  const double C2 = -1.38888888888741095749e-03;
  const double C3 = 2.48015872894767294178e-05;
  const double C4 = -2.75573143513906633035e-07;
  const double C5 = 2.08757232129817482790e-09;
  const double C6 = -1.13596475577881948265e-11;
  double z = AGF_DMUL(r, r);
  double p = AGF_DADD(C5, AGF_DMUL(z, C6));
  p = AGF_DADD(C4, AGF_DMUL(z, p));
  p = AGF_DADD(C3, AGF_DMUL(z, p));
  p = AGF_DADD(C2, AGF_DMUL(z, p));
  p = AGF_DADD(C1, AGF_DMUL(z, p));
  double hz = AGF_DMUL(0.5, z);
  double w = AGF_DSUB(1.0, hz);
  // 1 - hz is inexact; recover the rounding error (fdlibm trick) before adding the tail
  double e = AGF_DSUB(AGF_DSUB(1.0, w), hz);
  double tail = AGF_DADD(AGF_DMUL(AGF_DMUL(z, z), p), e);
  return AGF_DADD(w, tail);
}

AGF_HDX double agf_sin(double x) {
  if (!(x == x) || x - x != 0.0) return x - x;  // NaN, +-inf -> NaN
  double r;
  int q = agf_rem_pio2(x, &r);
  switch (q) {
    case 0: return agf_ksin(r);
    case 1: return agf_kcos(r);
    case 2: return -agf_ksin(r);
    default: return -agf_kcos(r);
  }
}

AGF_HDX double agf_cos(double x) {
  if (!(x == x) || x - x != 0.0) return x - x;
  double r;
  int q = agf_rem_pio2(x, &r);
  switch (q) {
    case 0: return agf_kcos(r);
    case 1: return -agf_ksin(r);
    case 2: return -agf_kcos(r);
    default: return agf_ksin(r);
  }
}

// ---------------------------------------------------------------------------
// asin / acos (fdlibm rational approximation of (asin(x)-x)/x^3 on [0, 0.5])
// ---------------------------------------------------------------------------
AGF_HD double agf_asin_R(double z) {
  const double pS0 = 1.66666666666666657415e-01;
  const double pS1 = -3.25565818622400915405e-01;
  const double pS2 = 2.01212532134862925881e-01;
  const double pS3 = -4.00555345006794114027e-02;
  const double pS4 = 7.91534994289814532176e-04;
  const double pS5 = 3.47933107596021167570e-05;
  const double qS1 = -2.40339491173441421878e+00;
  const double qS2 = 2.02094576023350569471e+00;
  const double qS3 = -6.88283971605453293030e-01;
  const double qS4 = 7.70381505559019352791e-02;
  double p = AGF_DADD(pS4, AGF_DMUL(z, pS5));
  p = AGF_DADD(pS3, AGF_DMUL(z, p));
  p = AGF_DADD(pS2, AGF_DMUL(z, p));
  p = AGF_DADD(pS1, AGF_DMUL(z, p));
  p = AGF_DADD(pS0, AGF_DMUL(z, p));
  p = AGF_DMUL(z, p);
  double q = AGF_DADD(qS3, AGF_DMUL(z, qS4));
  q = AGF_DADD(qS2, AGF_DMUL(z, q));
  q = AGF_DADD(qS1, AGF_DMUL(z, q));
  q = AGF_DADD(1.0, AGF_DMUL(z, q));
  return AGF_DDIV(p, q);
}

// returns NaN for |x| > 1 (callers that emulate errno test the argument themselves)
AGF_HDX double agf_asin(double x) {
  const double pio2_hi = 1.57079632679489655800e+00;
  const double pio2_lo = 6.12323399573676603587e-17;
  double ax = x < 0 ? -x : x;
  if (!(ax <= 1.0)) return (x - x) / (x - x);  // NaN in, or |x| > 1 -> NaN
  if (ax < 0.5) {
    double z = AGF_DMUL(x, x);
    return AGF_DADD(x, AGF_DMUL(x, agf_asin_R(z)));
  }
  // asin(x) = pi/2 - 2*asin(sqrt((1-|x|)/2))
  double z = AGF_DMUL(AGF_DSUB(1.0, ax), 0.5);
  double s = AGF_DSQRT(z);
  double rr = agf_asin_R(z);
  double t = AGF_DADD(s, AGF_DMUL(s, rr));  // asin(s)
  double res = AGF_DSUB(pio2_hi, AGF_DSUB(AGF_DMUL(2.0, t), pio2_lo));
  return x < 0 ? -res : res;
}

AGF_HDX double agf_acos(double x) {
  const double pio2_hi = 1.57079632679489655800e+00;
  const double pio2_lo = 6.12323399573676603587e-17;
  double ax = x < 0 ? -x : x;
  if (!(ax <= 1.0)) return (x - x) / (x - x);
  if (ax < 0.5) {
    double z = AGF_DMUL(x, x);
    double a = AGF_DADD(x, AGF_DMUL(x, agf_asin_R(z)));  // asin(x)
    return AGF_DSUB(pio2_hi, AGF_DSUB(a, pio2_lo));
  }
  double z = AGF_DMUL(AGF_DSUB(1.0, ax), 0.5);
  double s = AGF_DSQRT(z);
  double t = AGF_DADD(s, AGF_DMUL(s, agf_asin_R(z)));  // asin(sqrt((1-|x|)/2))
  if (x > 0) return AGF_DMUL(2.0, t);
  // x <= -0.5: pi - 2*t
  return AGF_DSUB(AGF_DMUL(2.0, pio2_hi), AGF_DSUB(AGF_DMUL(2.0, t), AGF_DMUL(2.0, pio2_lo)));
}

// ---------------------------------------------------------------------------
// atan / atan2 (fdlibm: reduce to [0, 7/16] with 4 break points, odd polynomial)
// ---------------------------------------------------------------------------
AGF_HD double agf_atan_poly(double x) {
  const double a0 = 3.33333333333329318027e-01;
  const double a1 = -1.99999999998764832476e-01;
  const double a2 = 1.42857142725034663711e-01;
  const double a3 = -1.11111104054623557880e-01;
  const double a4 = 9.09088713343650656196e-02;
  const double a5 = -7.69187620504482999495e-02;
  const double a6 = 6.66107313738753120669e-02;
  const double a7 = -5.83357013379057348645e-02;
  const double a8 = 4.97687799461593236017e-02;
  const double a9 = -3.65315727442169155270e-02;
  const double a10 = 1.62858201153657823623e-02;
  double z = AGF_DMUL(x, x);
  double p = AGF_DADD(a9, AGF_DMUL(z, a10));
  p = AGF_DADD(a8, AGF_DMUL(z, p));
  p = AGF_DADD(a7, AGF_DMUL(z, p));
  p = AGF_DADD(a6, AGF_DMUL(z, p));
  p = AGF_DADD(a5, AGF_DMUL(z, p));
  p = AGF_DADD(a4, AGF_DMUL(z, p));
  p = AGF_DADD(a3, AGF_DMUL(z, p));
  p = AGF_DADD(a2, AGF_DMUL(z, p));
  p = AGF_DADD(a1, AGF_DMUL(z, p));
  p = AGF_DADD(a0, AGF_DMUL(z, p));
  // atan(x) = x - x*z*p
  return AGF_DMUL(AGF_DMUL(x, z), p);
}

AGF_HD double agf_atan(double x) {
  const double hi0 = 4.63647609000806093515e-01, lo0 = 2.26987774529616870924e-17;  // atan(0.5)
  const double hi1 = 7.85398163397448278999e-01, lo1 = 3.06161699786838301793e-17;  // atan(1)
  const double hi2 = 9.82793723247329054082e-01, lo2 = 1.39033110312309984516e-17;  // atan(1.5)
  const double hi3 = 1.57079632679489655800e+00, lo3 = 6.12323399573676603587e-17;  // atan(inf)
  if (!(x == x)) return x;
  double ax = x < 0 ? -x : x;
  double res;
  if (ax < 0.4375) {
    res = AGF_DSUB(ax, agf_atan_poly(ax));
  } else {
    double t, hi, lo;
    if (ax < 0.6875) {
      t = AGF_DDIV(AGF_DSUB(AGF_DMUL(2.0, ax), 1.0), AGF_DADD(2.0, ax));
      hi = hi0; lo = lo0;
    } else if (ax < 1.1875) {
      t = AGF_DDIV(AGF_DSUB(ax, 1.0), AGF_DADD(ax, 1.0));
      hi = hi1; lo = lo1;
    } else if (ax < 2.4375) {
      t = AGF_DDIV(AGF_DSUB(ax, 1.5), AGF_DADD(1.0, AGF_DMUL(1.5, ax)));
      hi = hi2; lo = lo2;
    } else {
      t = AGF_DDIV(-1.0, ax);
      hi = hi3; lo = lo3;
    }
    // hi - ((t*z*p - lo) - t)
    res = AGF_DSUB(hi, AGF_DSUB(AGF_DSUB(agf_atan_poly(t), lo), t));
  }
  return x < 0 ? -res : res;
}

AGF_HDX double agf_atan2(double y, double x) {
  const double pi = 3.14159265358979311600e+00;
  const double pi_lo = 1.22464679914735317723e-16;
  const double pio2 = 1.57079632679489655800e+00;
  if (!(x == x) || !(y == y)) return x + y;
  if (y == 0.0) {
    // sign of zero: follow IEEE atan2 for the cases the path can hit
    if (x > 0 || (x == 0 && !signbit(x))) return y;  // +-0
    return signbit(y) ? -pi : pi;
  }
  if (x == 0.0) return y > 0 ? pio2 : -pio2;
  double ay = y < 0 ? -y : y;
  double ax = x < 0 ? -x : x;
  if (ax - ax != 0.0) {  // x infinite
    if (ay - ay != 0.0) {
      double v = x > 0 ? AGF_DMUL(0.5, pio2) : AGF_DMUL(1.5, pio2);
      return y > 0 ? v : -v;
    }
    double v = x > 0 ? 0.0 : pi;
    return y > 0 ? v : -v;
  }
  if (ay - ay != 0.0) return y > 0 ? pio2 : -pio2;
  double z = agf_atan(AGF_DDIV(ay, ax));  // in [0, pi/2]
  double res;
  if (x > 0) {
    res = z;
  } else {
    res = AGF_DSUB(pi, AGF_DSUB(z, pi_lo));
  }
  return y < 0 ? -res : res;
}

// ---------------------------------------------------------------------------
// binary32 entry points: evaluate in binary64, round once
// ---------------------------------------------------------------------------
AGF_HD float agf_sinf(float x) { return (float)agf_sin((double)x); }
AGF_HD float agf_cosf(float x) { return (float)agf_cos((double)x); }
AGF_HD float agf_asinf(float x) { return (float)agf_asin((double)x); }
AGF_HD float agf_acosf(float x) { return (float)agf_acos((double)x); }
AGF_HD float agf_atan2f(float y, float x) { return (float)agf_atan2((double)y, (double)x); }

// ---------------------------------------------------------------------------
// cube root of a non-negative number: the planner's cubic solver calls
// pow(|r| + sqrt(r^2 - q^3), 1./3) (Common/Common/Math/RootFinder.hpp:80).
// Exponent split by bit manipulation (exact), Newton iterations y <- (2y + m/y^2)/3
// on m in [1, 8) from a secant start (relative error 0.11 -> 1e-2 -> 1.5e-4 -> 2e-8 -> <1e-15).
// ---------------------------------------------------------------------------
AGF_HD long long agf_d2bits(double x) {
#if defined(__CUDA_ARCH__)
  return __double_as_longlong(x);
#else
  long long b;
  __builtin_memcpy(&b, &x, sizeof(b));
  return b;
#endif
}
AGF_HD double agf_bits2d(long long b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(b);
#else
  double x;
  __builtin_memcpy(&x, &b, sizeof(x));
  return x;
#endif
}
AGF_HDX double agf_cbrt_pos(double x) {
  if (!(x > 0.0) || x > 1.7976931348623157e308) return x;  // 0, NaN and +inf pass through
  int eadj = 0;
  if (x < 2.2250738585072014e-308) {  // subnormal: scale by 2^54 (exact)
    x = AGF_DMUL(x, 18014398509481984.0);
    eadj = -18;
  }
  long long b = agf_d2bits(x);
  int e = (int)((b >> 52) & 0x7ff) - 1023;
  int k = e >= 0 ? e / 3 : -((2 - e) / 3);  // floor(e / 3)
  int r = e - 3 * k;                         // 0, 1, 2
  double m = agf_bits2d((b & 0x000fffffffffffffLL) | ((long long)(1023 + r) << 52));  // [1, 8)
  double y = AGF_DADD(1.0, AGF_DDIV(AGF_DSUB(m, 1.0), 7.0));
#pragma unroll 1
  for (int it = 0; it < 6; it++)
    y = AGF_DDIV(AGF_DADD(AGF_DMUL(2.0, y), AGF_DDIV(m, AGF_DMUL(y, y))), 3.0);
  return AGF_DMUL(y, agf_bits2d((long long)(1023 + k + eadj) << 52));
}

// ---------------------------------------------------------------------------
// exp(x): the offboard state estimator's discrete angular-velocity lag exp(-dt / tau)
// (Components/Offboard/MocapStateEstimator.cpp:96,164) has a data-dependent argument, so it needs
// the same routine on both sides.  x = k ln2 + r with |r| <= ln2 / 2 (two-part ln2, the products with
// k are exact), exp(r) = 1 + 2 r / (2 - c(r)) - r ... in the rational form R(r^2), result scaled by 2^k.
// ---------------------------------------------------------------------------
AGF_HDX double agf_exp(double x) {
  if (x != x) return x;
  if (x > 709.0) return 1.0e308 * 10.0;  // overflow (+inf)
  if (x < -745.0) return 0.0;
  const double ln2hi = 6.93147180369123816490e-01, ln2lo = 1.90821492927058770002e-10;
  const double invln2 = 1.44269504088896338700e+00;
  const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  const double magic = 6755399441055744.0;  // 1.5 * 2^52
  const double fk = AGF_DSUB(AGF_DADD(AGF_DMUL(x, invln2), magic), magic);
  const int k = (int)fk;
  const double hi = AGF_DSUB(x, AGF_DMUL(fk, ln2hi));
  const double lo = AGF_DMUL(fk, ln2lo);
  const double r = AGF_DSUB(hi, lo);
  const double t = AGF_DMUL(r, r);
  double c = AGF_DADD(P4, AGF_DMUL(t, P5));
  c = AGF_DADD(P3, AGF_DMUL(t, c));
  c = AGF_DADD(P2, AGF_DMUL(t, c));
  c = AGF_DADD(P1, AGF_DMUL(t, c));
  c = AGF_DSUB(r, AGF_DMUL(t, c));
  const double y = AGF_DSUB(1.0, AGF_DSUB(AGF_DSUB(lo, AGF_DDIV(AGF_DMUL(r, c), AGF_DSUB(2.0, c))), hi));
  if (k >= -1021 && k <= 1023) return AGF_DMUL(y, agf_bits2d((long long)(1023 + k) << 52));
  // result near the subnormal range: two exact power-of-two factors
  return AGF_DMUL(AGF_DMUL(y, agf_bits2d((long long)(1023 + k + 1000) << 52)), agf_bits2d((long long)(1023 - 1000) << 52));
}
