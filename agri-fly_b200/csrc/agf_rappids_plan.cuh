// agf_rappids_plan.cuh -- K6: the batched RAPPIDS planner kernel (C5 / N3).
//
// FindLowestCostTrajectory (Components/Components/DepthImagePlanner/DepthImagePlanner.cpp:91-214) for every vehicle of a
// population, each on its own depth image, with the reference's sequential semantics kept exactly (candidates in order, cost
// pruning against the best collision-free candidate so far, pyramids generated on demand and reused).  Two kernels per call:
//   * rappids_candidates_kernel, one thread per candidate: motion primitive (TrajectoryGenerator/SingleAxisTrajectory.cpp:59-103),
//     cost, the recursive input-feasibility test (RapidTrajectoryGenerator.cpp:75-160) and the velocity test (:163-205);
//   * rappids_plan_kernel, one warp per vehicle (a whole CTA for the vehicles whose previous plan was long): the loop over
//     the candidates; the survivors are collision checked in candidate order (DepthImagePlanner.cpp:216-301) with the whole
//     warp on each: the four lateral faces of a pyramid on four lanes (:382-454), the pyramid list searched with one ballot
//     (:356-380), the pixel scans of InflatePyramid (:456-970) 32 pixels per step in the reference's scan order -- a ballot
//     finds the first pixel that changes the state, the update is applied warp-uniformly and the lanes behind it are
//     re-evaluated, so the result is the sequential one; stretches of the scan whose outcome does not depend on the order
//     (frames of the spiral expansion without a blocker, unblocked shrink spans) are taken in one step.
// Row walks read the image [H][W]; column walks read a transposed copy [W][H] so that both are coalesced.
//
// Compiled twice (agf_rappids_plan.cu): PARITY (-fmad=false, agf_math.h: bit-comparable with the oracle) and
// FAST (FMA contraction, CUDA libm).  No tensor cores: nothing here is a contraction.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdio.h>
#include <stdint.h>

#include "agf_math.h"

namespace agfr {

#define AGFR_DEV __device__ __forceinline__
// the per-candidate pieces (primitive generation, feasibility tests, pyramid geometry) run once per candidate or pyramid,
// lane parallel, and are a small share of the executed instructions but most of the kernel's code when inlined at every
// call site; as real calls the kernel is 122 KB instead of 155 KB and 1.2 % faster (AGFR_COLD_CALLS=0 inlines them again)
#ifndef AGFR_FRAME_UNROLL
#define AGFR_FRAME_UNROLL 1
#endif
#ifndef AGFR_PHASE_CLOCKS
#define AGFR_PHASE_CLOCKS 0  // tuning builds: per-phase clock64() totals of a few warps, printed at the end of the kernel
#endif
#ifndef AGFR_CAND_MIN_BLOCKS
#define AGFR_CAND_MIN_BLOCKS 8  // CTAs per SM the candidate pass aims at: 4 (110 registers) / 6 / 8 / 12 -> 8.0 / 7.5 / 7.1 / 7.2 ms
#endif
#ifndef AGFR_DIV_CALLS
#define AGFR_DIV_CALLS 1
#endif
#ifndef AGFR_SHRINK_FOLD
#define AGFR_SHRINK_FOLD 1
#endif
#ifndef AGFR_FRAME_JUMP
#define AGFR_FRAME_JUMP 1  // frame jumps of the spiral expansion compiled in (their length is PlanParams::frameJump)
#endif
#ifndef AGFR_COLD_CALLS
#define AGFR_COLD_CALLS 1
#endif
#if AGFR_COLD_CALLS
#define AGFR_PIECE __device__ __noinline__
#define AGFR_PIECE_FN static __device__ __noinline__
#else
#define AGFR_PIECE __device__ __forceinline__
#define AGFR_PIECE_FN __device__ __forceinline__
#endif
#define AGFR_FULL 0xffffffffu

constexpr int kBlock = 128;           // 4 warps = 4 vehicles per CTA
constexpr int kWarps = kBlock / 32;
constexpr int kMaxPyr = 32;
// CTAs per SM the register allocation aims at (round 1, fused kernel: 5 measured best; the planning pass alone, without the
// candidates' primitive in registers: 5 / 6 / 8 / 10 / 12 -> 74.8 / 73.2 / 70.5 / 69.6 / 71.6 ms, profiles/r2/rappids_variants_split.log)
#ifndef AGFR_MIN_BLOCKS
#define AGFR_MIN_BLOCKS 8  // the parity variant too: 3 / 5 / 6 / 8 -> 66.1 / 66.7 / 65.5 / 61.1 ms (rappids_variants_parity_occupancy.log)
#endif
constexpr int kFrameUnroll = AGFR_FRAME_UNROLL;
constexpr int kBuf = 2;               // _pyramidSearchPixelBuffer (DepthImagePlanner.cpp:60)

struct PlanParams {
  // images: four planes of one scene record per vehicle, vehicle v's planes at + v * vstride (agf_rappids.cu, Handle)
  const uint16_t* img;    // [H][W]
  const uint16_t* imgT;   // [W][H]
  const uint16_t* gminR;  // [H][GW] minimum of each group of 32 pixels of a row (pixels <= ignore count as 65535)
  const uint16_t* gminC;  // [W][GH] the same per column
  size_t vstride;         // elements between the records of consecutive vehicles
  const double* state;    // [n][12]: vel0, acc0, grav, cost vector
  const double* cands;    // [n][kcap][4]
  uint8_t* flags;         // [n][kcap]: the candidate pass leaves its verdict code here, the planning pass replaces it by the flags
  double* ccost;          // [n][kcap]: cost of every candidate (candidate pass -> planning pass)
  void* results;          // agf_rappids_result [n]
  double* pyramids;       // [n][kMaxPyr][17]
  double* prims;          // [n][9]: alpha, beta, gamma of the returned primitive per axis (SingleAxisTrajectory state), or null
  int* next;              // work counter
  const int* order;       // dispatch order: the k-th vehicle handed out is order[k] (null: index order)
  const int* nlong;       // *nlong vehicles at the head of `order` are planned by a whole CTA each (null: none)
  int* nextLong;          // work counter of those
  unsigned* work;         // [n] clock cycles this call spent on each vehicle (the next call's dispatch order)
  int n, k, kcap;
  int W, H, GW, GH;       // GW = ceil(W / 32), GH = ceil(H / 32)
  double scale, f, cx, cy, rPlan, minDist;
  double fminA, fmaxA, wmaxA, minSec, vmax;
  int maxPyr, costKind;
  int shrinkFold;         // fold the unblocked updates of a span of an edge region into one reduction (shrink_span)
  int frameJump;          // iterations of the spiral expansion taken in one step when their frame holds no blocker (< 2: off)
  int edgeOff, num;       // int(f * rTrue / minDist), int(f * rPlan / scale)   (DepthImagePlanner.cpp:460,608)
  int ignore;             // uint16(rTrue / scale)                               (:506)
};

struct ResultRec {  // == agf_rappids_result
  int32_t found, best_index, n_generated, n_cost_checks, n_collision_checks, n_velocity_checks, n_collision_free,
      n_pyramids;
  double best_cost;
  double best_coeffs[18];
  double best_tf;
  int32_t pyramid_cap_hit, reserved_;
};

template<bool PARITY>
struct Mth {
  static AGFR_DEV double cos(double x) { return PARITY ? agf_cos(x) : ::cos(x); }
  static AGFR_DEV double acos(double x) { return PARITY ? agf_acos(x) : ::acos(x); }
  static AGFR_DEV double cbrt(double x) { return PARITY ? agf_cbrt_pos(x) : ::cbrt(x); }
};

// ---------------------------------------------------------------------------------------------
// closed-form roots (Common/Common/Math/RootFinder.hpp:60-174; float 2*pi and float eps as there)
// ---------------------------------------------------------------------------------------------
// FP64 division and square root as real calls in the code the planning pass runs once per candidate or section: inline, the
// 47 divisions and 11 roots were a quarter of that code, all of it fetched from beyond the SM's 32 KB instruction cache every
// time (the planning pass is bound by instruction fetch, DESIGN K6).  Same operation, same result.
#if AGFR_DIV_CALLS
static __device__ __noinline__ double ddiv(double a, double b) { return a / b; }
static __device__ __noinline__ double dsqrt(double a) { return sqrt(a); }
#else
AGFR_DEV double ddiv(double a, double b) { return a / b; }
AGFR_DEV double dsqrt(double a) { return sqrt(a); }
#endif

template<bool PARITY>
__device__ __noinline__ unsigned cubic(double a, double b, double c, double* x) {
  const float piF = 3.141592653589793238463;
  const float twoPiF = 2 * piF;
  const float epsF = 1e-12;
  const double a2 = a * a;
  double q = ddiv(a2 - 3 * b, 9);
  const double r = ddiv(a * (2 * a2 - 9 * b) + 27 * c, 54);
  const double r2 = r * r;
  const double q3 = q * q * q;
  if (r2 < q3) {
    double t = ddiv(r, dsqrt(q3));
    if (t < -1) t = -1;
    if (t > 1) t = 1;
    t = Mth<PARITY>::acos(t);
    a = ddiv(a, 3);
    q = -2 * dsqrt(q);
    x[0] = q * Mth<PARITY>::cos(ddiv(t, 3)) - a;
    x[1] = q * Mth<PARITY>::cos(ddiv(t + double(twoPiF), 3.0)) - a;
    x[2] = q * Mth<PARITY>::cos(ddiv(t - double(twoPiF), 3.0)) - a;
    return 3;
  }
  double A = -Mth<PARITY>::cbrt(fabs(r) + dsqrt(r2 - q3));
  if (r < 0) A = -A;
  const double B = (fabs(A) < double(epsF) ? 0 : ddiv(q, A));
  a = ddiv(a, 3);
  x[0] = (A + B) - a;
  x[1] = -0.5 * (A + B) - a;
  x[2] = 0.5 * sqrt(3.0) * (A - B);  // constant-folded root
  if (fabs(x[2]) < double(epsF)) {
    x[2] = x[1];
    return 2;
  }
  return 1;
}

template<bool PARITY>
__device__ __noinline__ unsigned quartic(double a, double b, double c, double d, double* root) {
  const float epsF = 1e-12;
  const double a3 = -b;
  const double b3 = a * c - 4.0 * d;
  const double c3 = -a * a * d - c * c + 4.0 * b * d;
  int n = 0;
  double x3[3];
  const unsigned nz = cubic<PARITY>(a3, b3, c3, x3);
  double q1, q2, p1, p2, D, sqD, y;
  y = x3[0];
  if (nz != 1) {
    if (fabs(x3[1]) > fabs(y)) y = x3[1];
    if (fabs(x3[2]) > fabs(y)) y = x3[2];
  }
  D = y * y - 4 * d;
  if (fabs(D) < double(epsF)) {
    q1 = q2 = y * 0.5;
    D = a * a - 4.0 * (b - y);
    if (fabs(D) < double(epsF)) {
      p1 = p2 = a * 0.5;
    } else {
      sqD = dsqrt(D);
      p1 = (a + sqD) * 0.5;
      p2 = (a - sqD) * 0.5;
    }
  } else {
    sqD = dsqrt(D);
    q1 = (y + sqD) * 0.5;
    q2 = (y - sqD) * 0.5;
    p1 = ddiv(a * q1 - c, q1 - q2);
    p2 = ddiv(c - a * q2, q1 - q2);
  }
  D = p1 * p1 - 4 * q1;
  if (!(D < 0.0)) {
    sqD = dsqrt(D);
    root[n++] = (-p1 + sqD) * 0.5;
    root[n++] = (-p1 - sqD) * 0.5;
  }
  D = p2 * p2 - 4 * q2;
  if (!(D < 0.0)) {
    sqD = dsqrt(D);
    root[n++] = (-p2 + sqD) * 0.5;
    root[n++] = (-p2 - sqD) * 0.5;
  }
  return n;
}

// roots of c0 t^4 + .. + c4 (or the cubic when the leading coefficient vanishes): the dispatch used at
// DepthImagePlanner.cpp:318-325,412-419
template<bool PARITY>
AGFR_PIECE_FN unsigned poly_roots(const double* c, double* roots) {
  if (fabs(c[0]) > 1e-6) return quartic<PARITY>(ddiv(c[1], c[0]), ddiv(c[2], c[0]), ddiv(c[3], c[0]), ddiv(c[4], c[0]), roots);
  return cubic<PARITY>(ddiv(c[2], c[1]), ddiv(c[3], c[1]), ddiv(c[4], c[1]), roots);
}

// ---------------------------------------------------------------------------------------------
// motion primitive of one candidate (one lane): rest-to-rest-goal quintic per axis
// ---------------------------------------------------------------------------------------------
struct Axis {
  double v0, a0;        // initial velocity / acceleration (initial position is the focal point, 0)
  double al, be, ga;    // alpha, beta, gamma
  double pk0, pk1;      // acceleration extremum times (SingleAxisTrajectory.cpp:120-141)
  AGFR_DEV double jerk(double t) const { return ga + be * t + (1 / 2.0) * al * t * t; }
  AGFR_DEV double acc(double t) const { return a0 + ga * t + (1 / 2.0) * be * t * t + (1 / 6.0) * al * t * t * t; }
  AGFR_DEV double vel(double t) const {
    return v0 + a0 * t + (1 / 2.0) * ga * t * t + (1 / 6.0) * be * t * t * t + (1 / 24.0) * al * t * t * t * t;
  }
  AGFR_DEV double pos(double t) const {
    return 0.0 + v0 * t + (1 / 2.0) * a0 * t * t + (1 / 6.0) * ga * t * t * t + (1 / 24.0) * be * t * t * t * t +
           (1 / 120.0) * al * t * t * t * t * t;
  }
  // fully defined end state (pf, 0, 0) at Tf (SingleAxisTrajectory.cpp:59-78)
  AGFR_PIECE void generate(double pf, double Tf) {
    const double da = 0.0 - a0;
    const double dv = 0.0 - v0 - a0 * Tf;
    const double dp = pf - 0.0 - v0 * Tf - 0.5 * a0 * Tf * Tf;
    const double T2 = Tf * Tf, T3 = T2 * Tf, T4 = T3 * Tf, T5 = T4 * Tf;
    al = (60 * T2 * da - 360 * Tf * dv + 720 * dp) / T5;
    be = (-24 * T3 * da + 168 * T2 * dv - 360 * Tf * dp) / T5;
    ga = (3 * T4 * da - 24 * T3 * dv + 60 * T2 * dp) / T5;
    if (al != 0.0) {
      const double det = be * be - 2 * ga * al;
      if (det < 0) {
        pk0 = 0;
        pk1 = 0;
      } else {
        pk0 = (-be + sqrt(det)) / al;
        pk1 = (-be - sqrt(det)) / al;
      }
    } else {
      pk0 = (be != 0.0) ? -ga / be : 0;
      pk1 = 0;
    }
  }
  AGFR_PIECE void minmax_acc(double& lo, double& hi, double t1, double t2) const {
    const double e1 = acc(t1), e2 = acc(t2);
    lo = e2 < e1 ? e2 : e1;
    hi = e1 < e2 ? e2 : e1;
    if (pk0 > t1 && pk0 < t2) {
      const double e = acc(pk0);
      lo = e < lo ? e : lo;
      hi = hi < e ? e : hi;
    }
    if (pk1 > t1 && pk1 < t2) {
      const double e = acc(pk1);
      lo = e < lo ? e : lo;
      hi = hi < e ? e : hi;
    }
  }
  AGFR_PIECE double max_jerk_sq(double t1, double t2) const {
    const double j1 = jerk(t1), j2 = jerk(t2);
    const double s1 = j1 * j1, s2 = j2 * j2;
    double m = s1 < s2 ? s2 : s1;
    if (al != 0.0) {
      const double tm = -be / al;
      if (tm > t1 && tm < t2) {
        const double j = jerk(tm);
        const double s = j * j;
        m = s < m ? m : s;
      }
    }
    return m;
  }
};

enum { IN_FEASIBLE = 0, IN_INDETERMINABLE = 1, IN_THRUST_HIGH = 2, IN_THRUST_LOW = 3, IN_SPLIT = -1 };

struct Prim {
  Axis ax[3];
  double g[3];
  double tf;
  AGFR_DEV double thrust(double t) const {
    const double x = ax[0].acc(t) - g[0], y = ax[1].acc(t) - g[1], z = ax[2].acc(t) - g[2];
    return sqrt(x * x + y * y + z * z);
  }
  // one section of RapidTrajectoryGenerator::CheckInputFeasibilitySection (RapidTrajectoryGenerator.cpp:75-150)
  AGFR_DEV int section(const PlanParams& P, double t1, double t2) const {
    if (t2 - t1 < P.minSec) return IN_INDETERMINABLE;
    const double f1 = thrust(t1), f2 = thrust(t2);
    if ((f1 < f2 ? f2 : f1) > P.fmaxA) return IN_THRUST_HIGH;
    if ((f2 < f1 ? f2 : f1) < P.fminA) return IN_THRUST_LOW;
    double fminSqr = 0, fmaxSqr = 0, jmaxSqr = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      double amin, amax;
      ax[i].minmax_acc(amin, amax, t1, t2);
      const double v1 = amin - g[i];
      const double v2 = amax - g[i];
      const double s1 = v1 * v1, s2 = v2 * v2;
      if ((s1 < s2 ? s2 : s1) > P.fmaxA * P.fmaxA) return IN_THRUST_HIGH;
      const double f1a = fabs(v1), f2a = fabs(v2);
      if (v1 * v2 < 0) {
        fminSqr += 0;
      } else {
        const double m = f2a < f1a ? f2a : f1a;
        fminSqr += m * m;
      }
      const double M = f1a < f2a ? f2a : f1a;
      fmaxSqr += M * M;
      jmaxSqr += ax[i].max_jerk_sq(t1, t2);
    }
    const double fmin = sqrt(fminSqr);
    const double fmax = sqrt(fmaxSqr);
    double wBound;
    if (fminSqr > 1e-6)
      wBound = sqrt(jmaxSqr / fminSqr);
    else
      wBound = DBL_MAX;
    if (fmax < P.fminA) return IN_THRUST_LOW;
    if (fmin > P.fmaxA) return IN_THRUST_HIGH;
    if (fmin < P.fminA || fmax > P.fmaxA || wBound > P.wmaxA) return IN_SPLIT;
    return IN_FEASIBLE;
  }
  // the recursion (:133-147) is a depth-first walk over halved sections that stops at the first section that is
  // not feasible; the right halves still to visit are kept on a small stack
  __device__ __noinline__ int input_feasibility(const PlanParams& P) const {
    constexpr int kStack = 24;
    double hi[kStack];
    int sp = 0;
    double t1 = 0, t2 = tf;
    for (;;) {
      const int r = section(P, t1, t2);
      if (r == IN_SPLIT) {
        if (sp >= kStack) return IN_INDETERMINABLE;
        hi[sp++] = t2;            // right half [th, t2] later
        t2 = (t1 + t2) / 2;       // left half first
        continue;
      }
      if (r != IN_FEASIBLE) return r;
      if (sp == 0) return IN_FEASIBLE;
      t1 = t2;                    // the right sibling starts where this section ended (same value th)
      t2 = hi[--sp];
    }
  }
  // RapidTrajectoryGenerator::CheckVelocityFeasibility (:163-205): 0 feasible, 1 infeasible
  template<bool PARITY>
  __device__ __noinline__ int velocity_feasibility(double vmax) const {
#pragma unroll 1
    for (int dim = 0; dim < 3; dim++) {
      const double c0 = ax[dim].al / 6.0, c1 = ax[dim].be / 2.0, c2 = ax[dim].ga / 1.0, c3 = ax[dim].a0;
      double roots[5];
      unsigned n;
      if (fabs(c0) > 1e-6)
        n = cubic<PARITY>(c1 / c0, c2 / c0, c3 / c0, roots);
      else
        return 1;
      roots[n] = 0;
      roots[n + 1] = tf;
      for (unsigned i = 0; i < n + 2; i++) {
        const double r = roots[i];
        if (r < 0) continue;
        if (r > tf) continue;
        if (fabs(ax[0].vel(r)) >= vmax || fabs(ax[1].vel(r)) >= vmax || fabs(ax[2].vel(r)) >= vmax) return 1;
      }
    }
    return 0;
  }
};

// ---------------------------------------------------------------------------------------------
// warp-uniform pieces of the collision test
// ---------------------------------------------------------------------------------------------
struct Poly {
  double c[6][3];  // GetTrajectory(): t^5 first (RapidTrajectoryGenerator.hpp:232-241)
  AGFR_DEV double axis(int i, double t) const {
    return c[0][i] * t * t * t * t * t + c[1][i] * t * t * t * t + c[2][i] * t * t * t + c[3][i] * t * t + c[4][i] * t +
           c[5][i];
  }
};
struct Section {
  double t0, t1;
  bool inc;
};

struct WarpCtx {
  const uint16_t* img;
  const uint16_t* imgT;
  const uint16_t* gminR;
  const uint16_t* gminC;
  double* pdepth;  // shared: [kMaxPyr] base-plane depths, ascending
  int4* pedge;     // shared: [kMaxPyr] (right, top, left, bottom)
  int npyr;
  int capHit;  // a candidate needed a new pyramid when max_pyramids already existed (it was rejected, DepthImagePlanner.cpp:246-249)
  int lane;
#if AGFR_PHASE_CLOCKS
  unsigned long long* clk;  // [0] candidates [1] expansion [2] shrink [3] collision test outside inflate [4] initial rectangle
#endif
};

AGFR_DEV unsigned ld16(const uint16_t* p) { return (unsigned)__ldg(p); }
AGFR_DEV unsigned lanes_after(int src) { return 0xfffffffeu << src; }

// DepthImagePlanner::InflatePyramid (DepthImagePlanner.cpp:456-970), warp cooperative.
// The eight shrink regions share one scan routine; REGION selects geometry, trigger and update.
struct Shrink {
  int rS, lS, tS, bS;
};
enum { R_RIGHT = 0, R_LEFT, R_TOP, R_BOTTOM, R_TR, R_BR, R_TL, R_BL };

AGFR_DEV bool shrink_trigger(const int REGION, const Shrink& s, int num, int x, int y, int p) {
  switch (REGION) {
    case R_RIGHT: return num > (x - s.rS) * p;
    case R_LEFT: return (s.lS - x) * p < num;
    case R_TOP: return (s.tS - y) * p < num;
    case R_BOTTOM: return num > (y - s.bS) * p;
    case R_TR: return num > (x - s.rS) * p && (s.tS - y) * p < num;
    case R_BR: return num > (x - s.rS) * p && num > (y - s.bS) * p;
    case R_TL: return (s.lS - x) * p < num && (s.tS - y) * p < num;
    default: return (s.lS - x) * p < num && num > (y - s.bS) * p;
  }
}
// applies the update of one triggering pixel; false = "the pyramid cannot contain the sample point"
AGFR_DEV bool shrink_apply(const int REGION, Shrink& s, int num, int x, int y, int p, int x0, int y0) {
  const int q = num / p;
  const int rT = x - q, lT = x + q, tT = y + q, bT = y - q;
  if (REGION == R_RIGHT || REGION == R_LEFT) {
    const bool blocked = (REGION == R_RIGHT) ? (x0 > rT - kBuf) : (x0 < lT + kBuf);
    if (!blocked) {
      if (REGION == R_RIGHT) s.rS = rT; else s.lS = lT;
      return true;
    }
    const bool noTop = y0 < tT + kBuf, noBot = y0 > bT - kBuf;
    if (noTop && noBot) return false;
    if (noTop) {
      s.bS = bT;
    } else if (noBot) {
      s.tS = tT;
    } else {
      const int u = tT - s.tS, d = s.bS - bT;
      if (d > u) {
        s.tS = tT;
      } else if (REGION == R_RIGHT) {
        s.rS = bT;  // sic: the reference assigns the right edge here (DepthImagePlanner.cpp:648)
      } else {
        s.bS = bT;
      }
    }
    return true;
  }
  if (REGION == R_TOP || REGION == R_BOTTOM) {
    const bool blocked = (REGION == R_TOP) ? (y0 < tT + kBuf) : (y0 > bT - kBuf);
    if (!blocked) {
      if (REGION == R_TOP) s.tS = tT; else s.bS = bT;
      return true;
    }
    const bool noRight = x0 > rT - kBuf, noLeft = x0 < lT + kBuf;
    if (noRight && noLeft) return false;
    if (noRight) {
      s.lS = lT;
    } else if (noLeft) {
      s.rS = rT;
    } else {
      const int r = s.rS - rT, l = lT - s.lS;
      if (r > l) s.lS = lT; else s.rS = rT;
    }
    return true;
  }
  // corners: one horizontal (right/left) and one vertical (top/bottom) edge compete
  const bool isRight = (REGION == R_TR || REGION == R_BR), isTop = (REGION == R_TR || REGION == R_TL);
  const bool noH = isRight ? (x0 > rT - kBuf) : (x0 < lT + kBuf);
  const bool noV = isTop ? (y0 < tT + kBuf) : (y0 > bT - kBuf);
  if (noH && noV) return false;
  bool shrinkV;
  if (noH) {
    shrinkV = true;
  } else if (noV) {
    shrinkV = false;
  } else {
    const int hLoss = (isRight ? (s.rS - rT) : (lT - s.lS)) * (s.bS - s.tS);
    const int vLoss = (isTop ? (tT - s.tS) : (s.bS - bT)) * (s.rS - s.lS);
    shrinkV = hLoss > vLoss;
  }
  if (shrinkV) {
    if (isTop) s.tS = tT; else s.bS = bT;
  } else {
    if (isRight) s.rS = rT; else s.lS = lT;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// Group minima.  The scans below are waits on dependent pixel loads when every pixel is fetched (ncu, round 1: 64 % of
// the stall samples on the two pixel loads, 1.06 MB of DRAM reads per plan), and most of what they fetch cannot
// matter: a pixel only changes the state if it is nearer than the pyramid's depth.  gminR[y][g] / gminC[x][g] hold, for
// every image row / column, the minimum of each group of 32 pixels (pixels <= `ignore` count as 65535).  A scan first
// reads the minima of the groups its range touches (20 bytes per 320-pixel line, L1/L2 resident: 10 KB per image) and
// only groups whose minimum is below the bound are replayed 32 pixels per step in the reference's scan order.
// Skipping a group in which no pixel passes the order-independent part of the test is exact -- also for a group the
// range covers only partly, because its minimum bounds the sub-range's minimum from below.
// ---------------------------------------------------------------------------------------------
// Can any pixel of a group trigger an update (shrink_trigger) at all?  o: the line (x for column walks, y for row
// walks), [ia, ib]: the inner indices of the group's part of the range, gm: the group minimum (<= every seen pixel of
// it).  Each trigger has the form  f * p < num  with a factor f that does not depend on p; if the smallest factor of
// the group is >= 0 then f * p >= fmin * gm for all its pixels, so  fmin * gm >= num  rules every pixel out -- exactly,
// in integers.  The bounds only ever move inwards, which only makes factors larger, so testing with the current
// bounds is safe.
AGFR_DEV bool group_may_trigger(const int REGION, const Shrink& s, int num, int o, int ia, int ib, int gm) {
  const bool colWalk = (REGION == R_RIGHT || REGION == R_LEFT);
  bool may = true;
  if (REGION == R_RIGHT || REGION == R_TR || REGION == R_BR) {  // num > (x - rS) * p
    const int f = (colWalk ? o : ia) - s.rS;
    if (f >= 0 && f * gm >= num) may = false;
  }
  if (REGION == R_LEFT || REGION == R_TL || REGION == R_BL) {  // (lS - x) * p < num
    const int f = s.lS - (colWalk ? o : ib);
    if (f >= 0 && f * gm >= num) may = false;
  }
  if (REGION == R_TOP || REGION == R_TR || REGION == R_TL) {  // (tS - y) * p < num
    const int f = s.tS - o;
    if (f >= 0 && f * gm >= num) may = false;
  }
  if (REGION == R_BOTTOM || REGION == R_BR || REGION == R_BL) {  // num > (y - bS) * p
    const int f = o - s.bS;
    if (f >= 0 && f * gm >= num) may = false;
  }
  return may;
}

// the 32-pixels-per-step scan of inner indices i0 + dI*k, k < cnt, of one line (reference order)
AGFR_DEV bool shrink_span(const int REGION, const PlanParams& P, const WarpCtx& w, Shrink& s, int maxDepth, int x0, int y0,
                          int o, const uint16_t* line, int i0, int dI, int cnt) {
  const bool colWalk = (REGION == R_RIGHT || REGION == R_LEFT);  // outer x, inner y
  for (int base = 0; base < cnt; base += 32) {
    const int k = base + w.lane;
    const bool valid = k < cnt;
    const int i = i0 + dI * k;
    const int p = valid ? (int)ld16(line + i) : 0;
    const bool pre = valid && p > P.ignore && p < maxDepth;
    unsigned todo = __ballot_sync(AGFR_FULL, pre);
    const int x = colWalk ? o : i, y = colWalk ? i : o;
#if AGFR_SHRINK_FOLD
    // The four edge regions, when no triggering pixel of the span is "blocked" by the sample point: every update is an
    // assignment bound = f(pixel) that only ever moves the bound inwards, a pixel that no longer triggers once the bound has
    // moved would have moved it less, and whether a pixel is blocked does not depend on the bounds -- so the sequential result
    // is the extremum of f over the pixels that trigger with the bounds at the start of the span: one reduction instead of up
    // to 32 dependent updates (shrink_apply was 18 % of the planning pass's instructions).
    if (P.shrinkFold && REGION <= R_BOTTOM && todo) {
      const bool trig0 = pre && shrink_trigger(REGION, s, P.num, x, y, p);
      if (!__any_sync(AGFR_FULL, trig0)) continue;
      const int q = trig0 ? P.num / p : 0;
      const bool inward = (REGION == R_RIGHT || REGION == R_BOTTOM);  // bound decreases: x - q / y - q
      const int cand = REGION == R_RIGHT ? x - q : REGION == R_LEFT ? x + q : REGION == R_TOP ? y + q : y - q;
      const int ref0 = (REGION == R_RIGHT || REGION == R_LEFT) ? x0 : y0;
      const bool blocked = inward ? (ref0 > cand - kBuf) : (ref0 < cand + kBuf);
      if (!__any_sync(AGFR_FULL, trig0 && blocked)) {
        if (inward) {
          const int m = __reduce_min_sync(AGFR_FULL, trig0 ? cand : INT_MAX);
          if (REGION == R_RIGHT) s.rS = m; else s.bS = m;
        } else {
          const int m = __reduce_max_sync(AGFR_FULL, trig0 ? cand : INT_MIN);
          if (REGION == R_LEFT) s.lS = m; else s.tS = m;
        }
        continue;
      }
    }
#endif
    while (todo) {
      const bool trig = ((todo >> w.lane) & 1u) && shrink_trigger(REGION, s, P.num, x, y, p);
      const unsigned tm = __ballot_sync(AGFR_FULL, trig);
      if (!tm) break;
      const int src = __ffs(tm) - 1;
      const int sx = __shfl_sync(AGFR_FULL, x, src), sy = __shfl_sync(AGFR_FULL, y, src),
                sp = __shfl_sync(AGFR_FULL, p, src);
      if (!shrink_apply(REGION, s, P.num, sx, sy, sp, x0, y0)) return false;
      todo &= lanes_after(src);
    }
  }
  return true;
}

// scans one region in the reference's order.  outer: o = o0, o0+dO, .. (nOuter values); inner: i = i0 + dI*k, k < cnt
// One copy of this code for the eight regions (REGION is a run-time argument): inlined eight times the scan outgrew
// the instruction cache (ncu, round 1: 74 % of the stall samples were instruction fetches).
static __device__ __noinline__ bool shrink_region(const int REGION, const PlanParams& P, const WarpCtx& w, Shrink& s,
                                                  int maxDepth, int x0, int y0, int o0, int dO, int nOuter, int i0, int dI,
                                                  int cnt) {
  const bool colWalk = (REGION == R_RIGHT || REGION == R_LEFT);  // outer x, inner y
  if (cnt <= 0) return true;
  const int pitch = colWalk ? P.H : P.W, G = colWalk ? P.GH : P.GW;
  const int lo = dI > 0 ? i0 : i0 - (cnt - 1), hi = dI > 0 ? i0 + (cnt - 1) : i0;
  const uint16_t* plane = colWalk ? w.imgT : w.img;
  const uint16_t* gplane = colWalk ? w.gminC : w.gminR;
  const int g0 = lo >> 5, g1 = hi >> 5;
  // Several lines per step: a line's range touches ng <= 10 groups, so the 32 lanes take 32 / ng consecutive lines at once,
  // lane = line-in-step * ng + group-in-scan-order -- ascending lanes are the reference's scan order (line by line, then
  // along the line).  One load of group minima and one ballot then cover up to 32 lines (corner regions: 1-2 groups per
  // line) instead of one.  Groups flagged by the ballot are replayed pixel by pixel in lane order with the CURRENT bounds;
  // a flag computed with bounds that an earlier replay of the same step has since moved inwards is merely conservative
  // (group_may_trigger: inward bounds only rule more pixels out), so the result stays the sequential one.
  const int ng = g1 - g0 + 1;
  if (ng > 32) {  // images wider than 1024 pixels: one line per step, 32 groups at a time
    for (int oc = 0; oc < nOuter; oc++) {
      const int o = o0 + dO * oc;
      const uint16_t* gl = gplane + (size_t)o * G;
      for (int gb = 0; gb < ng; gb += 32) {
        const int k = gb + w.lane;
        const int g = dI > 0 ? g0 + k : g1 - k;
        const bool in = k < ng;
        const unsigned gm = in ? ld16(gl + g) : 65535u;
        unsigned need = __ballot_sync(AGFR_FULL, in && (int)gm < maxDepth &&
                                                     group_may_trigger(REGION, s, P.num, o, max(lo, g << 5), min(hi, (g << 5) + 31), (int)gm));
        while (need) {
          const int src = __ffs(need) - 1;
          need &= need - 1;
          const int gs = dI > 0 ? g0 + gb + src : g1 - gb - src;
          const int a = max(lo, gs << 5), b = min(hi, (gs << 5) + 31);
          if (!shrink_span(REGION, P, w, s, maxDepth, x0, y0, o, plane + (size_t)o * pitch, dI > 0 ? a : b, dI, b - a + 1))
            return false;
        }
      }
    }
    return true;
  }
  const int lps = 32 / ng;                      // lines per step (320-pixel rows: ng <= 10, at least 3)
  const int ll = w.lane / ng, k = w.lane - ll * ng;
  const int g = dI > 0 ? g0 + k : g1 - k;
  const int ga = max(lo, g << 5), gb_ = min(hi, (g << 5) + 31);
  for (int oc = 0; oc < nOuter; oc += lps) {
    const bool in = ll < lps && oc + ll < nOuter;
    const int o = o0 + dO * (oc + ll);
    const unsigned gm = in ? ld16(gplane + (size_t)o * G + g) : 65535u;
    unsigned need = __ballot_sync(AGFR_FULL, in && (int)gm < maxDepth && group_may_trigger(REGION, s, P.num, o, ga, gb_, (int)gm));
    while (need) {
      const int src = __ffs(need) - 1;
      need &= need - 1;
      const int os = __shfl_sync(AGFR_FULL, o, src), a = __shfl_sync(AGFR_FULL, ga, src), b = __shfl_sync(AGFR_FULL, gb_, src);
      if (!shrink_span(REGION, P, w, s, maxDepth, x0, y0, os, plane + (size_t)os * pitch, dI > 0 ? a : b, dI, b - a + 1))
        return false;
    }
  }
  return true;
}

// one line of the spiral expansion (DepthImagePlanner.cpp:521-597): returns true when a pixel nearer than the
// pyramid's minimum depth blocks the side; folds the depths seen before it into maxDepth.
// `plane` + idx * pitch is the pixel line, `gplane` + idx * G its group minima.
static __device__ __noinline__ bool expand_line(const PlanParams& P, const WarpCtx& w, const uint16_t* plane, const uint16_t* gplane,
                                                int pitch, int G, int idx, int a, int b, int minPyr, int& maxDepth) {
  const uint16_t* line = plane + (size_t)idx * pitch;
  const uint16_t* gl = gplane + (size_t)idx * G;
  unsigned mn = 65535u;
  bool blocked = false;
  const int g0 = a >> 5, g1 = b >> 5;
  for (int gb = g0; gb <= g1 && !blocked; gb += 32) {
    const int g = gb + w.lane;
    const bool in = g <= g1;
    const unsigned gm = in ? ld16(gl + g) : 65535u;
    const bool full = in && (g << 5) >= a && (g << 5) + 31 <= b;
    // a group is looked at pixel by pixel if it may hold a blocking pixel, or if it is covered partly and may lower maxDepth
    // (the minimum of the covered part is >= the group minimum, so gm >= maxDepth means it cannot)
    const bool cand = in && (int)gm < minPyr;
    unsigned scan = __ballot_sync(AGFR_FULL, cand || (in && !full && (int)gm < maxDepth));
    int blockLane = 32;
    while (scan) {
      const int src = __ffs(scan) - 1;
      scan &= scan - 1;
      const int gs = gb + src;
      const int ca = max(a, gs << 5), cb = min(b, (gs << 5) + 31);
      const int vv = ca + w.lane;
      const bool valid = vv <= cb;
      const int p = valid ? (int)ld16(line + vv) : 0;
      const bool sees = valid && p > P.ignore;
      const bool blk = sees && p < minPyr;
      const unsigned bm = __ballot_sync(AGFR_FULL, blk);
      const unsigned before = bm ? ((1u << (__ffs(bm) - 1)) - 1u) : AGFR_FULL;
      if (sees && !blk && ((before >> w.lane) & 1u)) mn = min(mn, (unsigned)p);
      if (bm) {
        blocked = true;
        blockLane = src;
        break;
      }
    }
    // whole groups without a blocking pixel, before the blocking group: their minimum is the minimum of their seen pixels
    if (full && !cand && w.lane < blockLane) mn = min(mn, gm);
  }
  mn = __reduce_min_sync(AGFR_FULL, mn);
  if ((int)mn < maxDepth) maxDepth = (int)mn;
  return blocked;
}

// ---------------------------------------------------------------------------------------------
// Frame jumps.  The spiral scans one pixel line per side and iteration, and every line is two dependent memory round trips
// (group minima, then the pixels of the partly covered end groups): the planning pass spent its time waiting on exactly
// those (ncu, round 2: expand_line 39 % of the stall samples).  K iterations in which no side is blocked scan, taken
// together, exactly the frame between the rectangle and the rectangle grown by K on every active side (line j of a side
// spans what the neighbouring sides reached before it; the union over j is the full frame, corners included), they leave
// every flag unchanged and fold every seen pixel of the frame into maxDepth -- in any order.  So a frame whose pixels
// hold no blocker (ignore < p < minPyr) is taken in ONE step: the group minima of a strip's interior groups and the
// pixels of its end groups are requested together.  A frame that holds a blocker -- or
// might (only its group minimum is known) -- is not jumped: the caller runs the line-by-line iterations there, so a refused
// jump costs time, never exactness.
// frame_pair: two opposite strips of the frame.  Returns false when they hold a blocking pixel; otherwise folds their seen
// depths into mn.  Needs 16-byte aligned pixel lines (W, H
// multiples of 8; the host switches jumps off otherwise).
// ---------------------------------------------------------------------------------------------
static __device__ __noinline__ bool frame_pair(const PlanParams& P, const WarpCtx& w, const uint16_t* plane, const uint16_t* gplane,
                                               int pitch, int G, int lineA, int lineB, bool useA, bool useB, int nl, int a, int b,
                                               int minPyr, unsigned& mnOut) {
  // Two opposite strips of the frame (rows above and below, or columns left and right): nl <= 8 lines each from lineA / lineB,
  // the same inner range [a, b].  Interior groups (always covered completely) are settled by their minima; the two end groups
  // -- the only ones a range can cover partly -- by their pixels: an end group of nl lines is a block of nl x 32 pixels, ONE
  // 16-byte load per lane (lane = 4 * line + quarter).  Every load of the pair is independent of every other: one memory round
  // trip.  A blocker is a seen pixel below minPyr, so the pair is free of blockers iff the minimum over its seen pixels is
  // not below minPyr: only that minimum is formed.
  const int g0 = a >> 5, g1 = b >> 5;
  const int ni = g1 - g0 - 1;            // interior groups per line
  const int per = ni > 0 ? nl * ni : 0;  // interior items per strip
  if (2 * per > 128) return false;       // wider images than four chunks cover: no jump
  const int row = w.lane >> 2, qtr = w.lane & 3;
  const bool inRow = row < nl, two = g1 > g0;
  const uint4 none = make_uint4(0, 0, 0, 0);  // 0 <= ignore: never seen
  const uint16_t* la = plane + (size_t)(lineA + row) * pitch + qtr * 8;
  const uint16_t* lb = plane + (size_t)(lineB + row) * pitch + qtr * 8;
  uint4 q[4];
  q[0] = (inRow && useA) ? __ldg(reinterpret_cast<const uint4*>(la + (g0 << 5))) : none;
  q[1] = (inRow && useA && two) ? __ldg(reinterpret_cast<const uint4*>(la + (g1 << 5))) : none;
  q[2] = (inRow && useB) ? __ldg(reinterpret_cast<const uint4*>(lb + (g0 << 5))) : none;
  q[3] = (inRow && useB && two) ? __ldg(reinterpret_cast<const uint4*>(lb + (g1 << 5))) : none;
  unsigned mn = 65535u;
  if (per > 0) {
    const unsigned rni = 65536u / (unsigned)ni + 1u;  // it / ni == (it * rni) >> 16 for it < 128, ni <= 32
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int it = w.lane + 32 * j;
      const bool second = it >= per;
      if (second) it -= per;
      if (it < per && (second ? useB : useA)) {
        const int l = (int)(((unsigned)it * rni) >> 16), k = it - l * ni;
        mn = min(mn, ld16(gplane + (size_t)((second ? lineB : lineA) + l) * G + g0 + 1 + k));
      }
    }
  }
  // pixels outside [a, b] and pixels not seen (<= ignore) count as 65535
  const int ign = P.ignore;
  // (a rolled loop: straight-line code for the 32 pixels of a lane was fetched once per call and the planning pass waited on
  // instruction fetch more than on anything else -- ncu, round 2: no-instruction 15.7 stall cycles per issue)
#pragma unroll kFrameUnroll
  for (int e = 0; e < 4; e++) {
    const int base = (((e & 1) ? g1 : g0) << 5) + qtr * 8;
    const int lo = a - base, hi = b - base;  // covered: lo <= i <= hi
    const unsigned wd[4] = {q[e].x, q[e].y, q[e].z, q[e].w};
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int p = (int)((i & 1) ? (wd[i >> 1] >> 16) : (wd[i >> 1] & 0xffffu));
      if (p > ign && i >= lo && i <= hi) mn = min(mn, (unsigned)p);
    }
  }
  mn = __reduce_min_sync(AGFR_FULL, mn);
  if ((int)mn < minPyr) return false;
  mnOut = min(mnOut, mn);
  return true;
}

static __device__ __noinline__ bool inflate(const PlanParams& P, const WarpCtx& w, int x0, int y0, double minimumDepth,
                                     double& outDepth, int4& outEdge) {
  const int W = P.W, H = P.H, edgeOff = P.edgeOff;
  if (x0 <= edgeOff + kBuf + 1 || x0 > W - edgeOff - kBuf - 1 || y0 <= edgeOff + kBuf + 1 ||
      y0 > H - edgeOff - kBuf - 1)
    return false;
  // uint16_t((minimumDepth + rPlan) / scale): x86 converts through int32 and keeps the low 16 bits
  const int minPyr = (int)(uint16_t)(int)((minimumDepth + P.rPlan) / P.scale);
  const double rD = P.f * P.rPlan / (P.scale * minPyr);
  // double -> int: out-of-range is INT_MIN on x86 (cvttsd2si), saturating here; both fail the size test below
  const int initR = (minPyr == 0) ? INT_MIN : (int)rD;
  if (initR == INT_MIN || initR > (1 << 20)) return false;
  if (2 * initR >= (W < H ? W : H) - 2 * edgeOff) return false;

  int left, top, right, bottom;
  if (y0 - initR < edgeOff) {
    top = edgeOff;
    bottom = top + 2 * initR;
  } else {
    bottom = min(H - edgeOff - 1, y0 + initR);
    top = bottom - 2 * initR;
  }
  if (x0 - initR < edgeOff) {
    left = edgeOff;
    right = left + 2 * initR;
  } else {
    right = min(W - edgeOff - 1, x0 + initR);
    left = right - 2 * initR;
  }
#if AGFR_PHASE_CLOCKS
  long long c0_ = clock64();
#endif
  // the initial rectangle must be free (:509-517; order independent)
  {
    bool bad = false;
    for (int y = top; y < bottom; y++)
      for (int xb = left; xb < right; xb += 32) {
        const int x = xb + w.lane;
        if (x < right) {
          const int p = (int)ld16(w.img + (size_t)y * W + x);
          bad |= (p <= minPyr && p > P.ignore);
        }
      }
    if (__any_sync(AGFR_FULL, bad)) return false;
  }
#if AGFR_PHASE_CLOCKS
  { long long c1_ = clock64(); w.clk[4] += c1_ - c0_; c0_ = c1_; }
#endif
  // spiral expansion (:519-599)
  int maxDepth = 65535;
  bool rf = true, tf = true, lf = true, bf = true;
#if AGFR_FRAME_JUMP
  int cool = 0;  // line-by-line iterations to run before the next jump is tried
#endif
  while (rf || tf || lf || bf) {
#if AGFR_FRAME_JUMP
    if (cool == 0) {
      int K = min(P.frameJump, 8);
      if (rf) K = min(K, W - edgeOff - 1 - right);
      if (tf) K = min(K, top - edgeOff);
      if (lf) K = min(K, left - edgeOff);
      if (bf) K = min(K, H - edgeOff - 1 - bottom);
      if (K >= 2) {
        const int r2 = right + (rf ? K : 0), t2 = top - (tf ? K : 0), l2 = left - (lf ? K : 0), b2 = bottom + (bf ? K : 0);
        unsigned mn = 65535u;
        // rows above and below span the grown width (corners included), columns left and right the old height
        bool ok = true;
        if (tf || bf) ok = frame_pair(P, w, w.img, w.gminR, W, P.GW, t2, bottom + 1, tf, bf, K, l2, r2, minPyr, mn);
        if (ok && (lf || rf)) ok = frame_pair(P, w, w.imgT, w.gminC, H, P.GH, l2, right + 1, lf, rf, K, top, bottom, minPyr, mn);
        if (ok) {
          if ((int)mn < maxDepth) maxDepth = (int)mn;
          right = r2;
          top = t2;
          left = l2;
          bottom = b2;
#if AGFR_PHASE_CLOCKS
          w.clk[5] += 1;
#endif
          continue;
        }
#if AGFR_PHASE_CLOCKS
        w.clk[6] += 1;
#endif
        cool = K;  // a blocker within K lines of some side: line by line until a side stops (or K iterations)
      }
    }
    if (cool > 0) cool--;
    const bool rf0 = rf, tf0 = tf, lf0 = lf, bf0 = bf;
#endif
#if AGFR_PHASE_CLOCKS
    w.clk[7] += 1;
#endif
    if (rf) {
      if (right < W - edgeOff - 1) {
        if (expand_line(P, w, w.imgT, w.gminC, H, P.GH, right + 1, top, bottom, minPyr, maxDepth)) {
          rf = false;
          right--;
        }
        right++;
      } else {
        rf = false;
      }
    }
    if (tf) {
      if (top > edgeOff) {
        if (expand_line(P, w, w.img, w.gminR, W, P.GW, top - 1, left, right, minPyr, maxDepth)) {
          tf = false;
          top++;
        }
        top--;
      } else {
        tf = false;
      }
    }
    if (lf) {
      if (left > edgeOff) {
        if (expand_line(P, w, w.imgT, w.gminC, H, P.GH, left - 1, top, bottom, minPyr, maxDepth)) {
          lf = false;
          left++;
        }
        left--;
      } else {
        lf = false;
      }
    }
    if (bf) {
      if (bottom < H - edgeOff - 1) {
        if (expand_line(P, w, w.img, w.gminR, W, P.GW, bottom + 1, left, right, minPyr, maxDepth)) {
          bf = false;
          bottom--;
        }
        bottom++;
      } else {
        bf = false;
      }
    }
#if AGFR_FRAME_JUMP
    if (rf != rf0 || tf != tf0 || lf != lf0 || bf != bf0) cool = 0;
#endif
  }
#if AGFR_PHASE_CLOCKS
  { long long c1_ = clock64(); w.clk[1] += c1_ - c0_; c0_ = c1_; }
#endif
  // shrink by the projected vehicle radius (:601-939)
  Shrink s;
  s.rS = W - 1 - edgeOff;
  s.lS = edgeOff;
  s.tS = edgeOff;
  s.bS = H - 1 - edgeOff;
  const int rows = bottom - top + 1, cols = right - left + 1;
  // the eight regions in the reference's order, through ONE out-of-line scan routine (the loop is kept rolled so that
  // the compiler does not clone the routine per region)
#pragma unroll 1
  for (int r = R_RIGHT; r <= R_BL; r++) {
    int o0, dO, nOuter, i0, dI, cnt;
    if (r == R_RIGHT || r == R_LEFT) {  // columns outwards, rows top..bottom
      o0 = (r == R_RIGHT) ? right : left;
      dO = (r == R_RIGHT) ? 1 : -1;
      nOuter = (r == R_RIGHT) ? W - right : left + 1;
      i0 = top; dI = 1; cnt = rows;
    } else {                            // rows outwards; columns left..right, or from a corner outwards
      const bool up = (r == R_TOP || r == R_TR || r == R_TL);
      o0 = up ? top : bottom;
      dO = up ? -1 : 1;
      nOuter = up ? top + 1 : H - bottom;
      if (r == R_TOP || r == R_BOTTOM) {
        i0 = left; dI = 1; cnt = cols;
      } else if (r == R_TR || r == R_BR) {
        i0 = right; dI = 1; cnt = W - right;
      } else {
        i0 = left; dI = -1; cnt = left + 1;
      }
    }
    if (!shrink_region(r, P, w, s, maxDepth, x0, y0, o0, dO, nOuter, i0, dI, cnt)) return false;
    if (r == R_LEFT && s.lS + kBuf > s.rS - kBuf) return false;
    if (r == R_BOTTOM && s.tS + kBuf > s.bS - kBuf) return false;
  }
#if AGFR_PHASE_CLOCKS
  w.clk[2] += clock64() - c0_;
#endif
  outDepth = maxDepth * P.scale - P.rPlan;
  outEdge = make_int4(s.rS, s.tS, s.lS, s.bS);
  return true;
}

// corner `k` of a pyramid (top right, top left, bottom left, bottom right; DepthImagePlanner.cpp:945-957)
AGFR_PIECE_FN void pyr_corner(const PlanParams& P, double depth, const int4& e, int k, double* c) {
  const int ex = (k == 0 || k == 3) ? e.x : e.z;  // right : left
  const int ey = (k < 2) ? e.y : e.w;             // top : bottom
  c[0] = depth * ddiv((double)ex - P.cx, P.f);
  c[1] = depth * ddiv((double)ey - P.cy, P.f);
  c[2] = depth * 1;
}
// unit normal of lateral face `f` (Pyramid.hpp:55-58; the norm is truncated to float, Vec3.hpp:126-129)
AGFR_PIECE_FN void pyr_normal(const PlanParams& P, double depth, const int4& e, int f, double* n) {
  double a[3], b[3];
  pyr_corner(P, depth, e, f, a);
  pyr_corner(P, depth, e, (f + 1) & 3, b);
  const double x = a[1] * b[2] - a[2] * b[1];
  const double y = a[2] * b[0] - a[0] * b[2];
  const double z = a[0] * b[1] - a[1] * b[0];
  const float nrm = (float)dsqrt(x * x + y * y + z * z);
  n[0] = ddiv(x, nrm);
  n[1] = ddiv(y, nrm);
  n[2] = ddiv(z, nrm);
}

AGFR_DEV Section make_section(const Poly& Q, double t0, double t1) {
  Section s;
  s.t0 = t0;
  s.t1 = t1;
  s.inc = Q.axis(2, t0) < Q.axis(2, t1);
  return s;
}
AGFR_DEV double deepest(const Poly& Q, const Section& s) { return Q.axis(2, s.inc ? s.t1 : s.t0); }

// DepthImagePlanner::IsCollisionFree (DepthImagePlanner.cpp:216-301) for one candidate; all lanes hold the same Q
template<bool PARITY>
__device__ __noinline__ bool collision_free(const PlanParams& P, WarpCtx& w, const Poly& Q, double tEnd) {
  // sections of monotonic depth (:303-354)
  Section st[8];
  int ns = 0;
  {
    double d[5];
#pragma unroll
    for (int i = 0; i < 5; i++) d[i] = (5 - i) * Q.c[i][2];
    double roots[6];
    roots[0] = 0.0;
    roots[1] = tEnd;
    const unsigned n = poly_roots<PARITY>(d, roots + 2);
    const int m = (int)n + 2;
    for (int i = 1; i < m; i++) {  // ascending
      const double kx = roots[i];
      int j = i - 1;
      while (j >= 0 && kx < roots[j]) {
        roots[j + 1] = roots[j];
        j--;
      }
      roots[j + 1] = kx;
    }
    for (unsigned i = 0; i < n + 1; i++) {
      if (roots[i] < 0.0) continue;
      if (fabs(roots[i] - roots[i + 1]) < 1e-6) continue;
      if (roots[i] >= tEnd) break;
      if (roots[i + 1] <= tEnd)
        st[ns++] = make_section(Q, roots[i], roots[i + 1]);
      else
        break;
    }
    // std::sort by deepest point == stable insertion sort at this size (ties are the rule: neighbours share a turning point)
    for (int i = 1; i < ns; i++) {
      const Section kx = st[i];
      const double kd = deepest(Q, kx);
      int j = i - 1;
      while (j >= 0 && kd < deepest(Q, st[j])) {
        st[j + 1] = st[j];
        j--;
      }
      st[j + 1] = kx;
    }
  }
  while (ns > 0) {
    const Section s = st[--ns];
    const double ts = s.inc ? s.t0 : s.t1, te = s.inc ? s.t1 : s.t0;
    const double startZ = Q.axis(2, ts);
    const double ex = Q.axis(0, te), ey = Q.axis(1, te), ez = Q.axis(2, te);
    if (startZ < P.minDist && ez < P.minDist) continue;
    const double px = ddiv(ex * P.f, ez) + P.cx;
    const double py = ddiv(ey * P.f, ez) + P.cy;
    // FindContainingPyramid (:356-380): first pyramid in depth order, not shallower than the point, that contains it
    double pdepth;
    int4 pedge;
    {
      bool ok = false;
      if (w.lane < w.npyr) {
        const double d = w.pdepth[w.lane];
        const int4 e = w.pedge[w.lane];
        ok = !(d < ez) && (e.z + kBuf < px) && (px < e.x - kBuf) && (e.y + kBuf < py) && (py < e.w - kBuf);
      }
      const unsigned hit = __ballot_sync(AGFR_FULL, ok);
      if (hit) {
        const int idx = __ffs(hit) - 1;
        pdepth = w.pdepth[idx];
        pedge = w.pedge[idx];
      } else {
        if (w.npyr >= P.maxPyr) {
          w.capHit = 1;
          return false;
        }
        // (int) of a double: the values reaching here are finite and small (the end point is deeper than minDist)
        if (!inflate(P, w, (int)px, (int)py, ez, pdepth, pedge)) return false;
        // insert in depth order (std::lower_bound on the new depth, :267-269)
        bool less = false;
        double d = 0;
        int4 e = make_int4(0, 0, 0, 0);
        if (w.lane < w.npyr) {
          d = w.pdepth[w.lane];
          e = w.pedge[w.lane];
          less = d < pdepth;
        }
        const int pos = __popc(__ballot_sync(AGFR_FULL, less));
        __syncwarp();
        if (w.lane < w.npyr && w.lane >= pos) {
          w.pdepth[w.lane + 1] = d;
          w.pedge[w.lane + 1] = e;
        }
        if (w.lane == 0) {
          w.pdepth[pos] = pdepth;
          w.pedge[pos] = pedge;
        }
        w.npyr++;
        __syncwarp();
      }
    }
    // FindDeepestCollisionTime (:382-454): lateral face (lane & 3); the deepest crossing over the four faces is
    // the latest (increasing depth) or the earliest (decreasing depth) one inside the section
    double nrm[3];
    pyr_normal(P, pdepth, pedge, w.lane & 3, nrm);
    double c[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int dim = 0; dim < 3; dim++)
#pragma unroll
      for (int k = 0; k < 5; k++) c[k] += nrm[dim] * Q.c[k][dim];
    double roots[4];
    const unsigned n = poly_roots<PARITY>(c, roots);
    double tc = s.inc ? s.t0 : s.t1;
    bool hit = false;
    for (unsigned i = 0; i < n; i++) {
      const double r = roots[i];
      if (s.inc) {
        if (!(r > s.t1) && r > s.t0 && r > tc) {
          tc = r;
          hit = true;
        }
      } else {
        if (!(r < s.t0) && r < s.t1 && r < tc) {
          tc = r;
          hit = true;
        }
      }
    }
#pragma unroll
    for (int off = 1; off <= 2; off <<= 1) {
      const double o = __shfl_xor_sync(AGFR_FULL, tc, off);
      tc = s.inc ? (o > tc ? o : tc) : (o < tc ? o : tc);
    }
    hit = __any_sync(AGFR_FULL, hit);
    if (hit) {
      if (s.inc)
        st[ns++] = make_section(Q, s.t0, tc);
      else
        st[ns++] = make_section(Q, tc, s.t1);
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// The kernels.  A planner call is two launches:
//   rappids_candidates_kernel  one THREAD per candidate: motion primitive, cost, and the recursive input-feasibility and
//                              velocity tests (RapidTrajectoryGenerator.cpp:75-205).  All of it is a pure function of the
//                              candidate and the vehicle's state, so it is evaluated for every candidate up front, all 32
//                              lanes busy, and leaves 9 bytes per candidate (cost, verdict code).
//   rappids_plan_kernel        one WARP per vehicle, vehicles handed out through a counter: the reference's sequential loop
//                              (DepthImagePlanner.cpp:91-214) over the precomputed costs / verdicts -- pruning against the best
//                              collision-free cost so far, the counters, and the collision test with its pyramids.
// Until round 2 both lived in one kernel (the tests ran speculatively for the lanes that beat the best cost at the start of
// a batch of 32).  Split, the planning kernel no longer carries the 100 KB of FP64 feasibility code next to InflatePyramid's
// scans (instruction fetch was a top stall) nor the primitive's registers across the collision test.
// ---------------------------------------------------------------------------------------------
enum { CODE_IN_MASK = 7, CODE_VEL_OK = 8 };  // verdict code: input-feasibility result | velocity test passed

template<bool PARITY>
__global__ void __launch_bounds__(128, AGFR_CAND_MIN_BLOCKS) rappids_candidates_kernel(const __grid_constant__ PlanParams P) {
  const size_t total = (size_t)P.n * (size_t)P.k;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t v = idx / (size_t)P.k;
    const int i = (int)(idx - v * (size_t)P.k);
    const double* st = P.state + v * 12;
    Prim pr;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      pr.ax[a].v0 = __ldg(st + a);
      pr.ax[a].a0 = __ldg(st + 3 + a);
      pr.g[a] = __ldg(st + 6 + a);
    }
    const double cv0 = __ldg(st + 9), cv1 = __ldg(st + 10), cv2 = __ldg(st + 11);
    const double4 c4 = *reinterpret_cast<const double4*>(P.cands + (v * P.kcap + i) * 4);
    const double T = c4.w;
    pr.tf = T;
    pr.ax[0].generate(c4.x, T);
    pr.ax[1].generate(c4.y, T);
    pr.ax[2].generate(c4.z, T);
    const double pe0 = pr.ax[0].pos(T), pe1 = pr.ax[1].pos(T), pe2 = pr.ax[2].pos(T);
    double cost;
    if (P.costKind == 0) {
      cost = -(cv0 * pe0 + cv1 * pe1 + cv2 * pe2) / T;
    } else {
      const double sg = sqrt((cv0 - 0) * (cv0 - 0) + (cv1 - 0) * (cv1 - 0) + (cv2 - 0) * (cv2 - 0));
      const double dx = cv0 - pe0, dy = cv1 - pe1, dz = cv2 - pe2;
      cost = -(sg - sqrt(dx * dx + dy * dy + dz * dz)) / T;
    }
    int code = pr.input_feasibility(P);
    if (code == IN_FEASIBLE && pr.template velocity_feasibility<PARITY>(P.vmax) == 0) code |= CODE_VEL_OK;
    P.ccost[v * P.kcap + i] = cost;
    P.flags[v * P.kcap + i] = (uint8_t)code;
  }
}

// One vehicle's planning loop.  coop == false: this warp alone, the reference's sequential loop.  coop == true: the four warps
// of the CTA on ONE vehicle (the vehicles whose previous plan was long -- a 65 536-plan launch is as long as its longest plan):
// every warp runs the same loop over the candidates with identical copies of the loop state; the next up to four candidates
// that reach the collision test are tested SPECULATIVELY, one per warp, each against its own copy of the pyramid list as it
// stood at the start of the round; the results are then committed in candidate order -- a result is valid as long as no
// earlier candidate of the round added a pyramid or became the best, the first one that does ends the round, its list becomes
// every warp's list and the candidates behind it are tested again.  A test is a deterministic function of (candidate, pyramid
// list), so what is committed is exactly what the sequential loop computes.
template<bool PARITY>
__device__ __noinline__ void plan_vehicle(const PlanParams& P, WarpCtx& w, const int v, const bool coop, const int wid,
                                          double (*s_depth)[kMaxPyr + 1], int4 (*s_edge)[kMaxPyr + 1], volatile int (*s_res)[4]
#if AGFR_PHASE_CLOCKS
                                          , unsigned long long* clk
#endif
) {
  const int lane = w.lane;
  const int nw = coop ? kWarps : 1, me = coop ? wid : 0;
  const long long work0 = clock64();
  w.img = P.img + (size_t)v * P.vstride;
  w.imgT = P.imgT + (size_t)v * P.vstride;
  w.gminR = P.gminR + (size_t)v * P.vstride;
  w.gminC = P.gminC + (size_t)v * P.vstride;
  w.npyr = 0;
  w.capHit = 0;
  const double* st = P.state + (size_t)v * 12;

  double best = DBL_MAX;
  int found = 0, bestIdx = -1, nCost = 0, nColl = 0, nVel = 0, nFree = 0, capHit = 0;

  for (int i0 = 0; i0 < P.k; i0 += 32) {
#if AGFR_PHASE_CLOCKS
    const long long cc0_ = clock64();
#endif
    const int i = i0 + lane;
    const bool valid = i < P.k;
    double cost = DBL_MAX;
    int code = 0;
    if (valid) {
      cost = P.ccost[(size_t)v * P.kcap + i];
      code = P.flags[(size_t)v * P.kcap + i];
    }
    if (coop) __syncthreads();  // every warp has read the verdict codes of this batch before warp 0 replaces them by the flags
    // the lanes that beat the best cost at the start of the batch (it can only get lower)
    unsigned pend = __ballot_sync(AGFR_FULL, valid && cost < best);
    unsigned flag = 0;
#if AGFR_PHASE_CLOCKS
    clk[0] += clock64() - cc0_;
#endif
    while (pend) {
      // --- the next up to nw candidates that reach the collision test, in candidate order
      int slot[kWarps];
      int ns = 0;
      {
        unsigned walk = pend;
        while (walk && ns < nw) {
          const int src = __ffs(walk) - 1;
          walk &= walk - 1;
          if (!(__shfl_sync(AGFR_FULL, cost, src) < best)) continue;
          const int cs = __shfl_sync(AGFR_FULL, code, src);
          if ((cs & CODE_IN_MASK) == IN_FEASIBLE && (cs & CODE_VEL_OK)) slot[ns++] = src;
        }
      }
      // --- test: warp `me` takes slot `me`
      const int npyr0 = w.npyr;
      int myFree = 0;
      if (me < ns) {
        // the survivor's primitive again (same routine, same inputs as the candidate pass), warp-uniform
        Poly Q;
        const double4 c4 = *reinterpret_cast<const double4*>(P.cands + ((size_t)v * P.kcap + i0 + slot[me]) * 4);
        const double goal[3] = {c4.x, c4.y, c4.z};
        const double Ts = c4.w;
#pragma unroll
        for (int a = 0; a < 3; a++) {
          Axis ax;
          ax.v0 = __ldg(st + a);
          ax.a0 = __ldg(st + 3 + a);
          ax.generate(goal[a], Ts);
          Q.c[0][a] = ddiv(ax.al, 120);
          Q.c[1][a] = ddiv(ax.be, 24);
          Q.c[2][a] = ddiv(ax.ga, 6);
          Q.c[3][a] = ax.acc(0.0) * 0.5;  // == / 2
          Q.c[4][a] = ax.vel(0.0);
          Q.c[5][a] = ax.pos(0.0);
        }
        w.capHit = 0;
#if AGFR_PHASE_CLOCKS
        const long long cf0_ = clock64();
        const unsigned long long in0_ = clk[1] + clk[2] + clk[4];
        myFree = collision_free<PARITY>(P, w, Q, Ts) ? 1 : 0;
        clk[3] += (clock64() - cf0_) - (clk[1] + clk[2] + clk[4] - in0_);
#else
        myFree = collision_free<PARITY>(P, w, Q, Ts) ? 1 : 0;
#endif
      }
      if (coop) {
        if (me < ns && lane == 0) {
          s_res[me][0] = myFree;
          s_res[me][1] = w.npyr;
          s_res[me][2] = w.capHit;
        }
        __syncthreads();
      }
      // --- commit in candidate order (every warp computes the same)
      int s = 0, adopt = -1;
      bool stop = false;
      while (pend && !stop) {
        const int src = __ffs(pend) - 1;
        const double csrc = __shfl_sync(AGFR_FULL, cost, src);
        if (!(csrc < best)) {
          pend &= pend - 1;
          continue;
        }
        const int cs = __shfl_sync(AGFR_FULL, code, src);
        const bool tested = (cs & CODE_IN_MASK) == IN_FEASIBLE && (cs & CODE_VEL_OK);
        if (tested && s >= ns) break;  // not tested in this round: the next one
        unsigned f = 1;  // LowCost
        nCost++;
        if ((cs & CODE_IN_MASK) == IN_FEASIBLE) {
          f |= 2;
          nColl++;
          if (cs & CODE_VEL_OK) {
            f |= 4;
            nVel++;
            const int rFree = coop ? s_res[s][0] : myFree;
            const int rNpyr = coop ? s_res[s][1] : w.npyr;
            capHit |= coop ? s_res[s][2] : w.capHit;
            if (rFree) {
              f |= 8;
              found = 1;
              best = csrc;
              nFree++;
              bestIdx = i0 + src;
              stop = true;  // the cost bound moved: what was tested behind it is void
            }
            if (rNpyr != npyr0) {
              adopt = s;
              stop = true;  // the pyramid list grew: what was tested behind it is void
            }
            s++;
          }
        }
        pend &= pend - 1;
        if (lane == src) flag = f;
      }
      if (coop && ns > 0) {
        // every warp continues with the committed list: that of the slot that changed it, else slot 0's (unchanged)
        const int from = adopt >= 0 ? adopt : 0;
        const int npyrT = s_res[from][1];
        if (me != from) {
          for (int q = lane; q < npyrT; q += 32) {
            s_depth[me][q] = s_depth[from][q];
            s_edge[me][q] = s_edge[from][q];
          }
          w.npyr = npyrT;
        }
        __syncthreads();
      }
    }
    if (valid && me == 0) P.flags[(size_t)v * P.kcap + i] = (uint8_t)flag;
  }
  if (me != 0) return;
  // results
  ResultRec* out = reinterpret_cast<ResultRec*>(P.results) + v;
  {  // the returned primitive -- polynomial coefficients, and the generator's own variables for the tracking loop --
     // regenerated from the winning candidate (same routine, same inputs) rather than carried through the candidate loop:
     // lane 3k+a holds coefficient k of axis a
    const int a = lane % 3, kq = lane / 3;
    double coef = 0, al = 0, be = 0, ga = 0, bestT = 0;
    if (found) {
      const double4 c4 = *reinterpret_cast<const double4*>(P.cands + ((size_t)v * P.kcap + bestIdx) * 4);
      bestT = c4.w;
      if (lane < 18) {
        const double goal = a == 0 ? c4.x : (a == 1 ? c4.y : c4.z);
        Axis ax;
        ax.v0 = __ldg(st + a);
        ax.a0 = __ldg(st + 3 + a);
        ax.generate(goal, c4.w);
        al = ax.al;
        be = ax.be;
        ga = ax.ga;
        coef = kq == 0 ? ddiv(al, 120) : kq == 1 ? ddiv(be, 24) : kq == 2 ? ddiv(ga, 6) : kq == 3 ? ax.acc(0.0) * 0.5 : kq == 4 ? ax.vel(0.0) : ax.pos(0.0);
      }
    }
    if (lane < 18) out->best_coeffs[lane] = coef;
    if (P.prims && lane < 3) {
      P.prims[(size_t)v * 9 + 3 * lane] = al;
      P.prims[(size_t)v * 9 + 3 * lane + 1] = be;
      P.prims[(size_t)v * 9 + 3 * lane + 2] = ga;
    }
    if (lane == 0) {
      out->found = found;
      out->best_index = bestIdx;
      out->n_generated = P.k;
      out->n_cost_checks = nCost;
      out->n_collision_checks = nColl;
      out->n_velocity_checks = nVel;
      out->n_collision_free = nFree;
      out->n_pyramids = w.npyr;
      out->best_cost = best;
      out->best_tf = bestT;
      out->pyramid_cap_hit = capHit;
      out->reserved_ = 0;
      const long long dt = clock64() - work0;
      P.work[v] = dt > 0xffffffffLL ? 0xffffffffu : (unsigned)dt;
    }
  }
  {
    double* rec = P.pyramids + ((size_t)v * kMaxPyr + lane) * 17;
    if (lane < w.npyr) {
      const double d = w.pdepth[lane];
      const int4 e = w.pedge[lane];
      rec[0] = d;
      rec[1] = e.x;
      rec[2] = e.y;
      rec[3] = e.z;
      rec[4] = e.w;
      for (int f = 0; f < 4; f++) pyr_normal(P, d, e, f, rec + 5 + 3 * f);
    } else {
      const double nanv = __longlong_as_double(0x7ff8000000000000LL);
      for (int q = 0; q < 17; q++) rec[q] = nanv;
    }
  }
  __syncwarp();
}

template<bool PARITY>
__global__ void __launch_bounds__(kBlock, AGFR_MIN_BLOCKS) rappids_plan_kernel(const __grid_constant__ PlanParams P) {
  __shared__ double s_depth[kWarps][kMaxPyr + 1];
  __shared__ int4 s_edge[kWarps][kMaxPyr + 1];
  __shared__ int s_res[kWarps][4];
  __shared__ int s_v;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  WarpCtx w;
  w.lane = lane;
  w.pdepth = s_depth[wid];
  w.pedge = s_edge[wid];
#if AGFR_PHASE_CLOCKS
  unsigned long long clk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  w.clk = clk;
  const long long cstart_ = clock64();
  int nplans_ = 0;
#define AGFR_CLK_ARG , clk
#else
#define AGFR_CLK_ARG
#endif
  // the vehicles whose previous plan was long (the head of the dispatch order): one CTA per vehicle
  int base = 0;
  if (P.nlong) {
    base = *P.nlong;
    for (;;) {
      if (threadIdx.x == 0) s_v = atomicAdd(P.nextLong, 1);
      __syncthreads();
      const int kq = s_v;
      __syncthreads();
      if (kq >= base) break;
      plan_vehicle<PARITY>(P, w, P.order[kq], true, wid, s_depth, s_edge, s_res AGFR_CLK_ARG);
    }
  }
  // all others: one warp per vehicle, vehicles handed out through a counter
  for (;;) {
    int v = 0;
    if (lane == 0) {
      v = atomicAdd(P.next, 1) + base;
      if (v < P.n && P.order) v = P.order[v];
    }
    v = __shfl_sync(AGFR_FULL, v, 0);
    if (v >= P.n) break;
#if AGFR_PHASE_CLOCKS
    nplans_++;
#endif
    plan_vehicle<PARITY>(P, w, v, false, wid, s_depth, s_edge, s_res AGFR_CLK_ARG);
  }
#if AGFR_PHASE_CLOCKS
  if (lane == 0 && wid == 0 && blockIdx.x % 97 == 0)
    printf("phase cycles block %d: plans %d total %lld | candidates %llu  initial-rect %llu  expansion %llu  shrink %llu  collision-test %llu | jumps taken %llu refused %llu, line-by-line iterations %llu\n", blockIdx.x, nplans_, clock64() - cstart_, clk[0], clk[4], clk[1], clk[2], clk[3], clk[5], clk[6], clk[7]);
#endif
}

}  // namespace agfr
