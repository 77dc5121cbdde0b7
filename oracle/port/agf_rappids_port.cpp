// oracle/port/agf_rappids_port.cpp -- TEST INFRASTRUCTURE: CPU restatement of the reference's RAPPIDS
// planner path (C5 / N3), behind oracle/rappids_api.h.
//
// Plain sequential C++ (no STL containers on the path, no Eigen/OpenCV), written from the reference's
// behaviour and pinned bit-for-bit against the unmodified reference sources (oracle/_ref/libagf_rappids_ref_*.so,
// tests/test_rappids_oracle.py) and the committed golden vectors (tests/golden/rappids_vectors.npz).
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load the resulting library; the product
// never does.  /root/reference paths cited below are relative to Components/Components/ unless noted.
//
//   primitive            TrajectoryGenerator/SingleAxisTrajectory.cpp:59-177, SingleAxisTrajectory.hpp:53-64
//   feasibility tests    TrajectoryGenerator/RapidTrajectoryGenerator.cpp:75-205
//   cubic / quartic      Common/Common/Math/RootFinder.hpp:60-174
//   polynomial helpers   Common/Common/Math/Trajectory.hpp:78-121, DepthImagePlanner/MonotonicTrajectory.hpp:33-59
//   pyramid              DepthImagePlanner/Pyramid.hpp:49-59, Common/Common/Math/Vec3.hpp:101-128
//   planner              DepthImagePlanner/DepthImagePlanner.cpp:91-214 (search), 216-303 (collision test),
//                        303-356 (monotonic sections), 356-380 (pyramid lookup), 382-454 (deepest collision),
//                        456-970 (pyramid inflation)
//   candidate sampler    DepthImagePlanner/DepthImagePlanner.hpp:334-393 (+ libstdc++ mt19937 / generate_canonical)
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <limits>
#include <thread>
#include <atomic>
#include <vector>

#include "../rappids_api.h"

#ifndef ORC_FLAVOUR
#define ORC_FLAVOUR "port-glibc"
#endif

#ifdef AGF_PORT_SHARED_MATH
#include "../../agri-fly_b200/csrc/agf_math.h"
static inline double m_cos(double x) { return agf_cos(x); }
static inline double m_acos(double x) { return agf_acos(x); }
static inline double m_cbrt(double x) { return agf_cbrt_pos(x); }
#else
static inline double m_cos(double x) { return cos(x); }
static inline double m_acos(double x) { return acos(x); }
static inline double m_cbrt(double x) { return pow(x, 1. / 3); }
#endif

namespace {

// ---------------------------------------------------------------------------------------------
// roots (RootFinder.hpp; note the float 2*pi and the float eps inside the double solver, :40-44)
// ---------------------------------------------------------------------------------------------
const float kPiF = 3.141592653589793238463;
const float k2PiF = 2 * kPiF;
const float kEpsF = 1e-12;

unsigned cubic(double a, double b, double c, double* x) {
  double a2 = a * a;
  double q = (a2 - 3 * b) / 9;
  double r = (a * (2 * a2 - 9 * b) + 27 * c) / 54;
  double r2 = r * r;
  double q3 = q * q * q;
  if (r2 < q3) {
    double t = r / sqrt(q3);
    if (t < -1) t = -1;
    if (t > 1) t = 1;
    t = m_acos(t);
    a /= 3;
    q = -2 * sqrt(q);
    x[0] = q * m_cos(t / 3) - a;
    x[1] = q * m_cos((t + double(k2PiF)) / double(3)) - a;
    x[2] = q * m_cos((t - double(k2PiF)) / double(3)) - a;
    return 3;
  }
  double A = -m_cbrt(fabs(r) + sqrt(r2 - q3));
  if (r < 0) A = -A;
  double B = (fabs(A) < double(kEpsF) ? 0 : q / A);
  a /= 3;
  x[0] = (A + B) - a;
  x[1] = double(-0.5) * (A + B) - a;
  x[2] = double(0.5) * sqrt(double(3.)) * (A - B);
  if (fabs(x[2]) < double(kEpsF)) {
    x[2] = x[1];
    return 2;
  }
  return 1;
}

unsigned quartic(double a, double b, double c, double d, double* root) {
  double a3 = -b;
  double b3 = a * c - double(4.) * d;
  double c3 = -a * a * d - c * c + double(4.) * b * d;
  int n = 0;
  double x3[3];
  unsigned nz = cubic(a3, b3, c3, x3);
  double q1, q2, p1, p2, D, sqD, y;
  y = x3[0];
  if (nz != 1) {
    if (fabs(x3[1]) > fabs(y)) y = x3[1];
    if (fabs(x3[2]) > fabs(y)) y = x3[2];
  }
  D = y * y - 4 * d;
  if (fabs(D) < double(kEpsF)) {
    q1 = q2 = y * double(0.5);
    D = a * a - double(4) * (b - y);
    if (fabs(D) < double(kEpsF)) {
      p1 = p2 = a * double(0.5);
    } else {
      sqD = sqrt(D);
      p1 = (a + sqD) * double(0.5);
      p2 = (a - sqD) * double(0.5);
    }
  } else {
    sqD = sqrt(D);
    q1 = (y + sqD) * double(0.5);
    q2 = (y - sqD) * double(0.5);
    p1 = (a * q1 - c) / (q1 - q2);
    p2 = (c - a * q2) / (q1 - q2);
  }
  D = p1 * p1 - 4 * q1;
  if (!(D < double(0.0))) {
    sqD = sqrt(D);
    root[n++] = (-p1 + sqD) * double(0.5);
    root[n++] = (-p1 - sqD) * double(0.5);
  }
  D = p2 * p2 - 4 * q2;
  if (!(D < double(0.0))) {
    sqD = sqrt(D);
    root[n++] = (-p2 + sqD) * double(0.5);
    root[n++] = (-p2 - sqD) * double(0.5);
  }
  return n;
}

// ascending sort of a handful of doubles (std::sort on values: any correct sort gives the same array)
void sort_small(double* v, int n) {
  for (int i = 1; i < n; i++) {
    double k = v[i];
    int j = i - 1;
    while (j >= 0 && k < v[j]) {
      v[j + 1] = v[j];
      j--;
    }
    v[j + 1] = k;
  }
}

// ---------------------------------------------------------------------------------------------
// motion primitive: per axis jerk parameters (alpha, beta, gamma) for a fully defined end state
// ---------------------------------------------------------------------------------------------
struct Axis {
  double p0, v0, a0, pf, vf, af;
  double al, be, ga;
  double peak[2];
  bool peakInit;
  double jerk(double t) const { return ga + be * t + (1 / 2.0) * al * t * t; }
  double acc(double t) const { return a0 + ga * t + (1 / 2.0) * be * t * t + (1 / 6.0) * al * t * t * t; }
  double vel(double t) const {
    return v0 + a0 * t + (1 / 2.0) * ga * t * t + (1 / 6.0) * be * t * t * t + (1 / 24.0) * al * t * t * t * t;
  }
  double pos(double t) const {
    return p0 + v0 * t + (1 / 2.0) * a0 * t * t + (1 / 6.0) * ga * t * t * t + (1 / 24.0) * be * t * t * t * t +
           (1 / 120.0) * al * t * t * t * t * t;
  }
  // SingleAxisTrajectory::GenerateTrajectory, branch "position, velocity and acceleration goal" (:75-78)
  void generate(double Tf) {
    double da = af - a0;
    double dv = vf - v0 - a0 * Tf;
    double dp = pf - p0 - v0 * Tf - 0.5 * a0 * Tf * Tf;
    const double T2 = Tf * Tf, T3 = T2 * Tf, T4 = T3 * Tf, T5 = T4 * Tf;
    al = (60 * T2 * da - 360 * Tf * dv + 720 * 1 * dp) / T5;
    be = (-24 * T3 * da + 168 * T2 * dv - 360 * Tf * dp) / T5;
    ga = (3 * T4 * da - 24 * T3 * dv + 60 * T2 * dp) / T5;
    peakInit = false;
  }
  // :118-154
  void minmax_acc(double& lo, double& hi, double t1, double t2) {
    if (!peakInit) {
      if (al) {
        double det = be * be - 2 * ga * al;
        if (det < 0) {
          peak[0] = 0;
          peak[1] = 0;
        } else {
          peak[0] = (-be + sqrt(det)) / al;
          peak[1] = (-be - sqrt(det)) / al;
        }
      } else {
        peak[0] = be ? -ga / be : 0;
        peak[1] = 0;
      }
      peakInit = true;
    }
    double e1 = acc(t1), e2 = acc(t2);
    lo = e2 < e1 ? e2 : e1;  // std::min(a, b) = (b < a) ? b : a
    hi = e1 < e2 ? e2 : e1;  // std::max(a, b) = (a < b) ? b : a
    for (int i = 0; i < 2; i++) {
      if (peak[i] <= t1) continue;
      if (peak[i] >= t2) continue;
      double e = acc(peak[i]);
      lo = e < lo ? e : lo;
      hi = hi < e ? e : hi;
    }
  }
  // :165-177
  double max_jerk_sq(double t1, double t2) const {
    double j1 = jerk(t1), j2 = jerk(t2);
    double s1 = j1 * j1, s2 = j2 * j2;
    double m = s1 < s2 ? s2 : s1;
    if (al) {
      double tm = -be / al;
      if (tm > t1 && tm < t2) {
        double j = jerk(tm);
        double s = j * j;
        m = s < m ? m : s;
      }
    }
    return m;
  }
};

enum { IN_FEASIBLE = 0, IN_INDETERMINABLE = 1, IN_THRUST_HIGH = 2, IN_THRUST_LOW = 3 };

struct Prim {
  Axis ax[3];
  double grav[3];
  double tf;
  void init(const double* v0, const double* a0, const double* g) {
    for (int i = 0; i < 3; i++) {
      ax[i].p0 = 0;
      ax[i].v0 = v0[i];
      ax[i].a0 = a0[i];
      ax[i].al = ax[i].be = ax[i].ga = 0;
      ax[i].peakInit = false;
      grav[i] = g[i];
    }
    tf = 0;
  }
  void generate(const double* goal, double T) {
    tf = T;
    for (int i = 0; i < 3; i++) {
      ax[i].pf = goal[i];
      ax[i].vf = 0;
      ax[i].af = 0;
      ax[i].generate(T);
    }
  }
  double thrust(double t) const {
    double x = ax[0].acc(t) - grav[0], y = ax[1].acc(t) - grav[1], z = ax[2].acc(t) - grav[2];
    return sqrt(x * x + y * y + z * z);
  }
  // one section of the recursive test (RapidTrajectoryGenerator.cpp:75-150): returns the verdict, or -1 when the
  // section must be split
  int section(double fminA, double fmaxA, double wmaxA, double t1, double t2, double minSec) {
    if (t2 - t1 < minSec) return IN_INDETERMINABLE;
    double f1 = thrust(t1), f2 = thrust(t2);
    if ((f1 < f2 ? f2 : f1) > fmaxA) return IN_THRUST_HIGH;
    if ((f2 < f1 ? f2 : f1) < fminA) return IN_THRUST_LOW;
    double fminSqr = 0, fmaxSqr = 0, jmaxSqr = 0;
    for (int i = 0; i < 3; i++) {
      double amin, amax;
      ax[i].minmax_acc(amin, amax, t1, t2);
      double v1 = amin - grav[i];
      double v2 = amax - grav[i];
      double s1 = v1 * v1, s2 = v2 * v2;
      if ((s1 < s2 ? s2 : s1) > fmaxA * fmaxA) return IN_THRUST_HIGH;
      double f1a = fabs(v1), f2a = fabs(v2);
      if (v1 * v2 < 0) {
        fminSqr += 0;
      } else {
        double m = f2a < f1a ? f2a : f1a;
        fminSqr += m * m;
      }
      double M = f1a < f2a ? f2a : f1a;
      fmaxSqr += M * M;
      jmaxSqr += ax[i].max_jerk_sq(t1, t2);
    }
    double fmin = sqrt(fminSqr);
    double fmax = sqrt(fmaxSqr);
    double wBound;
    if (fminSqr > 1e-6)
      wBound = sqrt(jmaxSqr / fminSqr);
    else
      wBound = std::numeric_limits<double>::max();
    if (fmax < fminA) return IN_THRUST_LOW;
    if (fmin > fmaxA) return IN_THRUST_HIGH;
    if (fmin < fminA || fmax > fmaxA || wBound > wmaxA) return -1;
    return IN_FEASIBLE;
  }
  // the recursion of :133-147 is a depth-first walk that stops at the first section that is not feasible
  int input_feasibility(double fminA, double fmaxA, double wmaxA, double minSec) {
    double s1[64], s2[64];
    int sp = 0;
    s1[0] = 0;
    s2[0] = tf;
    sp = 1;
    while (sp > 0) {
      sp--;
      double t1 = s1[sp], t2 = s2[sp];
      int r = section(fminA, fmaxA, wmaxA, t1, t2, minSec);
      if (r == -1) {
        double th = (t1 + t2) / 2;
        if (sp + 2 > 64) return IN_INDETERMINABLE;
        s1[sp] = th;  // second half is visited after the first
        s2[sp] = t2;
        sp++;
        s1[sp] = t1;
        s2[sp] = th;
        sp++;
      } else if (r != IN_FEASIBLE) {
        return r;
      }
    }
    return IN_FEASIBLE;
  }
  // :163-205 (0 = feasible, 1 = infeasible)
  int velocity_feasibility(double vmax) const {
    for (int dim = 0; dim < 3; dim++) {
      double c[4];
      c[0] = ax[dim].al / 6.0;
      c[1] = ax[dim].be / 2.0;
      c[2] = ax[dim].ga / 1.0;
      c[3] = ax[dim].a0;
      double roots[5];
      unsigned n = 0;
      if (fabs(c[0]) > 1e-6)
        n = cubic(c[1] / c[0], c[2] / c[0], c[3] / c[0], roots);
      else
        return 1;
      roots[n] = 0;
      roots[n + 1] = tf;
      for (unsigned i = 0; i < n + 2; i++) {
        if (roots[i] < 0) continue;
        if (roots[i] > tf) continue;
        double vx = ax[0].vel(roots[i]), vy = ax[1].vel(roots[i]), vz = ax[2].vel(roots[i]);
        if (fabs(vx) >= vmax || fabs(vy) >= vmax || fabs(vz) >= vmax) return 1;
      }
    }
    return 0;
  }
  // RapidTrajectoryGenerator::GetTrajectory (RapidTrajectoryGenerator.hpp:232-241): [6][3], t^5 first
  void coeffs(double c[6][3]) const {
    for (int a = 0; a < 3; a++) {
      c[0][a] = ax[a].al / 120;
      c[1][a] = ax[a].be / 24;
      c[2][a] = ax[a].ga / 6;
      c[3][a] = ax[a].acc(0) / 2;
      c[4][a] = ax[a].vel(0);
      c[5][a] = ax[a].pos(0);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// polynomial sections and pyramids
// ---------------------------------------------------------------------------------------------
struct Poly {
  double c[6][3];
  double axis(int i, double t) const {
    return c[0][i] * t * t * t * t * t + c[1][i] * t * t * t * t + c[2][i] * t * t * t + c[3][i] * t * t + c[4][i] * t +
           c[5][i];
  }
};
struct Section {
  double t0, t1;
  bool increasing;
};
struct Pyr {
  double depth;
  int right, top, left, bottom;
  double n[4][3];
};

void unit_cross(const double* a, const double* b, double* out) {
  double x = a[1] * b[2] - a[2] * b[1];
  double y = a[2] * b[0] - a[0] * b[2];
  double z = a[0] * b[1] - a[1] * b[0];
  float const nrm = sqrt(x * x + y * y + z * z);  // Vec3.hpp:127: the norm is truncated to float
  out[0] = x / nrm;
  out[1] = y / nrm;
  out[2] = z / nrm;
}

struct Planner {
  const orc_rappids_cfg* cfg;
  const uint16_t* img;
  int W, H;
  std::vector<Pyr> pyramids;  // ordered by base-plane depth
  int nGenerated, nCost, nColl, nVel, nFree;
  int maxPyr;
  static const int kBuf = 2;  // _pyramidSearchPixelBuffer

  void deproject(double x, double y, double depth, double* p) const {
    p[0] = depth * ((x - cfg->cx) / cfg->focal_length);
    p[1] = depth * ((y - cfg->cy) / cfg->focal_length);
    p[2] = depth * 1;
  }

  bool inflate(int x0, int y0, double minimumDepth, Pyr& out) const;
  bool find_pyramid(double px, double py, double depth, Pyr& out) const {
    size_t lo = 0, hi = pyramids.size();
    while (lo < hi) {  // first pyramid whose base plane is not shallower than `depth`
      size_t mid = (lo + hi) / 2;
      if (pyramids[mid].depth < depth)
        lo = mid + 1;
      else
        hi = mid;
    }
    for (size_t i = lo; i < pyramids.size(); i++) {
      const Pyr& p = pyramids[i];
      if (p.left + kBuf < px && px < p.right - kBuf && p.top + kBuf < py && py < p.bottom - kBuf) {
        out = p;
        return true;
      }
    }
    return false;
  }
  bool deepest_collision(const Poly& P, const Section& s, const Pyr& pyr, double& tOut) const {
    bool hit = false;
    tOut = s.increasing ? s.t0 : s.t1;
    for (int f = 0; f < 4; f++) {
      double c[5] = {0, 0, 0, 0, 0};
      for (int dim = 0; dim < 3; dim++)
        for (int k = 0; k < 5; k++) c[k] += pyr.n[f][dim] * P.c[k][dim];
      double roots[4];
      unsigned n;
      if (fabs(c[0]) > 1e-6)
        n = quartic(c[1] / c[0], c[2] / c[0], c[3] / c[0], c[4] / c[0], roots);
      else
        n = cubic(c[2] / c[1], c[3] / c[1], c[4] / c[1], roots);
      sort_small(roots, (int)n);
      if (s.increasing) {
        for (int i = (int)n - 1; i >= 0; i--) {
          if (roots[i] > s.t1) continue;
          if (roots[i] > s.t0) {
            if (roots[i] > tOut) {
              tOut = roots[i];
              hit = true;
              break;
            }
          } else {
            break;
          }
        }
      } else {
        for (int i = 0; i < (int)n; i++) {
          if (roots[i] < s.t0) continue;
          if (roots[i] < s.t1) {
            if (roots[i] < tOut) {
              tOut = roots[i];
              hit = true;
              break;
            }
          } else {
            break;
          }
        }
      }
    }
    return hit;
  }
  static Section make_section(const Poly& P, double t0, double t1) {
    Section s;
    s.t0 = t0;
    s.t1 = t1;
    s.increasing = P.axis(2, t0) < P.axis(2, t1);
    return s;
  }
  static double deepest(const Poly& P, const Section& s) { return P.axis(2, s.increasing ? s.t1 : s.t0); }

  bool collision_free(const Poly& P, double tStart, double tEnd) {
    // monotonic-depth sections (DepthImagePlanner.cpp:303-354)
    double d[5];
    for (int i = 0; i < 5; i++) d[i] = (5 - i) * P.c[i][2];
    double roots[6];
    roots[0] = tStart;
    roots[1] = tEnd;
    unsigned n;
    if (fabs(d[0]) > 1e-6)
      n = quartic(d[1] / d[0], d[2] / d[0], d[3] / d[0], d[4] / d[0], roots + 2);
    else
      n = cubic(d[2] / d[1], d[3] / d[1], d[4] / d[1], roots + 2);
    sort_small(roots, (int)n + 2);
    Section st[8];
    int ns = 0;
    for (unsigned i = 0; i < n + 1; i++) {
      if (roots[i] < tStart) continue;
      if (fabs(roots[i] - roots[i + 1]) < 1e-6) continue;
      if (roots[i] >= tEnd) break;
      if (roots[i + 1] <= tEnd)
        st[ns++] = make_section(P, roots[i], roots[i + 1]);
      else
        break;
    }
    // std::sort by deepest point; libstdc++ uses a (stable) insertion sort below 16 elements and ties are
    // the rule here (neighbouring sections share their turning point)
    for (int i = 1; i < ns; i++) {
      Section k = st[i];
      double kd = deepest(P, k);
      int j = i - 1;
      while (j >= 0 && kd < deepest(P, st[j])) {
        st[j + 1] = st[j];
        j--;
      }
      st[j + 1] = k;
    }
    // DepthImagePlanner.cpp:216-301
    while (ns > 0) {
      Section s = st[--ns];
      double ts = s.increasing ? s.t0 : s.t1, te = s.increasing ? s.t1 : s.t0;
      double startZ = P.axis(2, ts);
      double endP[3] = {P.axis(0, te), P.axis(1, te), P.axis(2, te)};
      if (startZ < cfg->min_checking_dist && endP[2] < cfg->min_checking_dist) continue;
      double px = endP[0] * cfg->focal_length / endP[2] + cfg->cx;
      double py = endP[1] * cfg->focal_length / endP[2] + cfg->cy;
      Pyr pyr;
      if (!find_pyramid(px, py, endP[2], pyr)) {
        if ((int64_t)pyramids.size() >= (int64_t)maxPyr) return false;
        if (!inflate((int)px, (int)py, endP[2], pyr)) return false;
        size_t lo = 0, hi = pyramids.size();
        while (lo < hi) {
          size_t mid = (lo + hi) / 2;
          if (pyramids[mid].depth < pyr.depth)
            lo = mid + 1;
          else
            hi = mid;
        }
        pyramids.insert(pyramids.begin() + lo, pyr);
      }
      double tc;
      if (deepest_collision(P, s, pyr, tc)) {
        if (s.increasing)
          st[ns++] = make_section(P, s.t0, tc);
        else
          st[ns++] = make_section(P, tc, s.t1);
      }
    }
    return true;
  }
};

// DepthImagePlanner::InflatePyramid (DepthImagePlanner.cpp:456-970), restated with one routine for the eight
// shrink regions instead of eight copies.  `A` is the axis that the region tries to shrink first.
struct Shrink {
  int right, left, top, bottom;  // the four "shrunk" edges
};

// every depth-image read of InflatePyramid goes through PX, which counts it: the algorithmic pixel traffic of a plan
// (orc_rappids_pixels_read, the numerator of K6's roofline in bench.py)
static std::atomic<uint64_t> g_pixels_read{0};
static thread_local uint64_t t_pixels_read = 0;
struct CountedImage {
  const uint16_t* p;
  uint16_t operator[](int i) const {
    t_pixels_read++;
    return p[i];
  }
};
bool Planner::inflate(int x0, int y0, double minimumDepth, Pyr& out) const {
  const CountedImage PX{img};
  const double f = cfg->focal_length, rTrue = cfg->true_radius, rPlan = cfg->planning_radius;
  const double scale = cfg->depth_scale;
  int edgeOff = f * rTrue / cfg->min_checking_dist;
  if (x0 <= edgeOff + kBuf + 1 || x0 > W - edgeOff - kBuf - 1 || y0 <= edgeOff + kBuf + 1 ||
      y0 > H - edgeOff - kBuf - 1)
    return false;
  uint16_t minPyrDepth = uint16_t((minimumDepth + rPlan) / scale);
  int initR = f * rPlan / (scale * minPyrDepth);
  if (2 * initR >= (W < H ? W : H) - 2 * edgeOff) return false;

  int left, top, right, bottom;
  if (y0 - initR < edgeOff) {
    top = edgeOff;
    bottom = top + 2 * initR;
  } else {
    int lim = H - edgeOff - 1;
    bottom = (y0 + initR < lim) ? y0 + initR : lim;
    top = bottom - 2 * initR;
  }
  if (x0 - initR < edgeOff) {
    left = edgeOff;
    right = left + 2 * initR;
  } else {
    int lim = W - edgeOff - 1;
    right = (x0 + initR < lim) ? x0 + initR : lim;
    left = right - 2 * initR;
  }
  uint16_t ignore = uint16_t(rTrue / scale);
  for (int y = top; y < bottom; y++)
    for (int x = left; x < right; x++) {
      uint16_t p = PX[y * W + x];
      if (p <= minPyrDepth && p > ignore) return false;
    }

  // spiral expansion
  uint16_t maxDepth = 65535;
  bool rf = true, tf = true, lf = true, bf = true;
  while (rf || tf || lf || bf) {
    if (rf) {
      if (right < W - edgeOff - 1) {
        for (int y = top; y <= bottom; y++) {
          uint16_t p = PX[y * W + right + 1];
          if (p > ignore) {
            if (p < minPyrDepth) {
              rf = false;
              right--;
              break;
            }
            if (p < maxDepth) maxDepth = p;
          }
        }
        right++;
      } else {
        rf = false;
      }
    }
    if (tf) {
      if (top > edgeOff) {
        for (int x = left; x <= right; x++) {
          uint16_t p = PX[(top - 1) * W + x];
          if (p > ignore) {
            if (p < minPyrDepth) {
              tf = false;
              top++;
              break;
            }
            if (p < maxDepth) maxDepth = p;
          }
        }
        top--;
      } else {
        tf = false;
      }
    }
    if (lf) {
      if (left > edgeOff) {
        for (int y = top; y <= bottom; y++) {
          uint16_t p = PX[y * W + left - 1];
          if (p > ignore) {
            if (p < minPyrDepth) {
              lf = false;
              left++;
              break;
            }
            if (p < maxDepth) maxDepth = p;
          }
        }
        left--;
      } else {
        lf = false;
      }
    }
    if (bf) {
      if (bottom < H - edgeOff - 1) {
        for (int x = left; x <= right; x++) {
          uint16_t p = PX[(bottom + 1) * W + x];
          if (p > ignore) {
            if (p < minPyrDepth) {
              bf = false;
              bottom--;
              break;
            }
            if (p < maxDepth) maxDepth = p;
          }
        }
        bottom++;
      } else {
        bf = false;
      }
    }
  }

  // shrink by the projected vehicle radius
  int rS = W - 1 - edgeOff, lS = edgeOff, tS = edgeOff, bS = H - 1 - edgeOff;
  int num = f * rPlan / scale;

  // --- right band: columns right..W-1 (outer), rows top..bottom (inner)
  for (int x = right; x < W; x++)
    for (int y = top; y <= bottom; y++) {
      int p = PX[y * W + x];
      if (p > ignore && p < maxDepth && num > (x - rS) * p) {
        int rT = x - int(num / p);
        if (x0 > rT - kBuf) {
          int tT = y + int(num / p), bT = y - int(num / p);
          if (y0 < tT + kBuf && y0 > bT - kBuf) return false;
          if (y0 < tT + kBuf) {
            bS = bT;
          } else if (y0 > bT - kBuf) {
            tS = tT;
          } else {
            int u = tT - tS, d = bS - bT;
            if (d > u)
              tS = tT;
            else
              rS = bT;  // sic: DepthImagePlanner.cpp:648 assigns the right edge here
          }
        } else {
          rS = rT;
        }
      }
    }
  // --- left band: columns left..0 (outer, descending), rows top..bottom
  for (int x = left; x >= 0; x--)
    for (int y = top; y <= bottom; y++) {
      int p = PX[y * W + x];
      if (p > ignore && p < maxDepth && (lS - x) * p < num) {
        int lT = x + int(num / p);
        if (x0 < lT + kBuf) {
          int tT = y + int(num / p), bT = y - int(num / p);
          if (y0 < tT + kBuf && y0 > bT - kBuf) return false;
          if (y0 < tT + kBuf) {
            bS = bT;
          } else if (y0 > bT - kBuf) {
            tS = tT;
          } else {
            int u = tT - tS, d = bS - bT;
            if (d > u)
              tS = tT;
            else
              bS = bT;
          }
        } else {
          lS = lT;
        }
      }
    }
  if (lS + kBuf > rS - kBuf) return false;
  // --- top band: rows top..0 (outer, descending), columns left..right
  for (int y = top; y >= 0; y--)
    for (int x = left; x <= right; x++) {
      int p = PX[y * W + x];
      if (p > ignore && p < maxDepth && (tS - y) * p < num) {
        int tT = y + int(num / p);
        if (y0 < tT + kBuf) {
          int rT = x - int(num / p), lT = x + int(num / p);
          if (x0 > rT - kBuf && x0 < lT + kBuf) return false;
          if (x0 > rT - kBuf) {
            lS = lT;
          } else if (x0 < lT + kBuf) {
            rS = rT;
          } else {
            int r = rS - rT, l = lT - lS;
            if (r > l)
              lS = lT;
            else
              rS = rT;
          }
        } else {
          tS = tT;
        }
      }
    }
  // --- bottom band: rows bottom..H-1, columns left..right
  for (int y = bottom; y < H; y++)
    for (int x = left; x <= right; x++) {
      int p = PX[y * W + x];
      if (p > ignore && p < maxDepth && num > (y - bS) * p) {
        int bT = y - int(num / p);
        if (y0 > bT - kBuf) {
          int rT = x - int(num / p), lT = x + int(num / p);
          if (x0 > rT - kBuf && x0 < lT + kBuf) return false;
          if (x0 > rT - kBuf) {
            lS = lT;
          } else if (x0 < lT + kBuf) {
            rS = rT;
          } else {
            int r = rS - rT, l = lT - lS;
            if (r > l)
              lS = lT;
            else
              rS = rT;
          }
        } else {
          bS = bT;
        }
      }
    }
  if (tS + kBuf > bS - kBuf) return false;
  // --- corners: top right, bottom right, top left, bottom left
  for (int y = top; y >= 0; y--)
    for (int x = right; x < W; x++) {
      int p = PX[y * W + x];
      if (p > ignore && p < maxDepth && num > (x - rS) * p && (tS - y) * p < num) {
        int rT = x - int(num / p), tT = y + int(num / p);
        if (x0 > rT - kBuf && y0 < tT + kBuf) return false;
        if (x0 > rT - kBuf) {
          tS = tT;
        } else if (y0 < tT + kBuf) {
          rS = rT;
        } else {
          int r = (rS - rT) * (bS - tS), u = (tT - tS) * (rS - lS);
          if (r > u)
            tS = tT;
          else
            rS = rT;
        }
      }
    }
  for (int y = bottom; y < H; y++)
    for (int x = right; x < W; x++) {
      int p = PX[y * W + x];
      if (p > ignore && p < maxDepth && num > (x - rS) * p && num > (y - bS) * p) {
        int rT = x - int(num / p), bT = y - int(num / p);
        if (x0 > rT - kBuf && y0 > bT - kBuf) return false;
        if (x0 > rT - kBuf) {
          bS = bT;
        } else if (y0 > bT - kBuf) {
          rS = rT;
        } else {
          int r = (rS - rT) * (bS - tS), d = (bS - bT) * (rS - lS);
          if (r > d)
            bS = bT;
          else
            rS = rT;
        }
      }
    }
  for (int y = top; y >= 0; y--)
    for (int x = left; x >= 0; x--) {
      int p = PX[y * W + x];
      if (p > ignore && p < maxDepth && (lS - x) * p < num && (tS - y) * p < num) {
        int lT = x + int(num / p), tT = y + int(num / p);
        if (x0 < lT + kBuf && y0 < tT + kBuf) return false;
        if (x0 < lT + kBuf) {
          tS = tT;
        } else if (y0 < tT + kBuf) {
          lS = lT;
        } else {
          int l = (lT - lS) * (bS - tS), u = (tT - tS) * (rS - lS);
          if (l > u)
            tS = tT;
          else
            lS = lT;
        }
      }
    }
  for (int y = bottom; y < H; y++)
    for (int x = left; x >= 0; x--) {
      int p = PX[y * W + x];
      if (p > ignore && p < maxDepth && (lS - x) * p < num && num > (y - bS) * p) {
        int lT = x + int(num / p), bT = y - int(num / p);
        if (x0 < lT + kBuf && y0 > bT - kBuf) return false;
        if (x0 < lT + kBuf) {
          bS = bT;
        } else if (y0 > bT - kBuf) {
          lS = lT;
        } else {
          int l = (lT - lS) * (bS - tS), d = (bS - bT) * (rS - lS);
          if (l > d)
            bS = bT;
          else
            lS = lT;
        }
      }
    }

  double depth = maxDepth * scale - rPlan;
  double c[4][3];
  deproject(double(rS), double(tS), depth, c[0]);
  deproject(double(lS), double(tS), depth, c[1]);
  deproject(double(lS), double(bS), depth, c[2]);
  deproject(double(rS), double(bS), depth, c[3]);
  out.depth = depth;
  out.right = rS;
  out.top = tS;
  out.left = lS;
  out.bottom = bS;
  unit_cross(c[0], c[1], out.n[0]);
  unit_cross(c[1], c[2], out.n[1]);
  unit_cross(c[2], c[3], out.n[2]);
  unit_cross(c[3], c[0], out.n[3]);
  return true;
}

// ---------------------------------------------------------------------------------------------
// candidate sampler: std::mt19937 + std::uniform_real_distribution<double> as libstdc++ implements them
// ---------------------------------------------------------------------------------------------
struct MT {
  uint32_t s[624];
  int idx;
  void seed(uint32_t v) {
    s[0] = v;
    for (int i = 1; i < 624; i++) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
  }
  uint32_t next() {
    if (idx >= 624) {
      for (int i = 0; i < 624; i++) {
        uint32_t y = (s[i] & 0x80000000u) | (s[(i + 1) % 624] & 0x7fffffffu);
        s[i] = s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = s[idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  double canonical() {  // generate_canonical<double, 53>: two draws, low word first
    double lo = (double)next();
    double hi = (double)next();
    double r = (lo + hi * 4294967296.0) / 18446744073709551616.0;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
  }
  double uniform(double a, double b) { return canonical() * (b - a) + a; }
};

int plan_one(const orc_rappids_cfg* cfg, const uint16_t* image, const double* vel0, const double* acc0,
             const double* grav, int n, const double* cands, int seed, const orc_rappids_sampler* sm,
             orc_rappids_out* out, uint8_t* results, double* cands_out, double* pyramids, int max_pyr_out) {
  Planner pl;
  pl.cfg = cfg;
  pl.img = image;
  pl.W = cfg->width;
  pl.H = cfg->height;
  pl.nGenerated = pl.nCost = pl.nColl = pl.nVel = pl.nFree = 0;
  pl.maxPyr = cfg->max_pyramids > 0 ? cfg->max_pyramids : std::numeric_limits<int>::max();
  Prim cand;
  cand.init(vel0, acc0, grav);
  MT mt;
  orc_rappids_sampler box;
  if (!cands) {
    mt.seed((uint32_t)seed);
    if (sm) {
      box = *sm;
      box.min_x = (int)sm->min_x;  // the custom constructor takes integer pixel bounds (DepthImagePlanner.hpp:367-370)
      box.max_x = (int)sm->max_x;
      box.min_y = (int)sm->min_y;
      box.max_y = (int)sm->max_y;
    } else {
      box.min_x = 0.1 * cfg->width;
      box.max_x = 0.9 * cfg->width;
      box.min_y = 0.1 * cfg->height;
      box.max_y = 0.9 * cfg->height;
      box.min_depth = 1.5;
      box.max_depth = 3.0;
      box.min_time = 2.0;
      box.max_time = 3.0;
    }
  }
  memset(out, 0, sizeof(*out));
  out->best_index = -1;
  double best = std::numeric_limits<double>::max();
  bool found = false;
  for (int i = 0; i < n; i++) {
    double goal[3], T;
    if (cands) {
      goal[0] = cands[4 * i];
      goal[1] = cands[4 * i + 1];
      goal[2] = cands[4 * i + 2];
      T = cands[4 * i + 3];
    } else {
      // g++ evaluates the three draws of DeprojectPixelToPoint(_pixelX(_gen), _pixelY(_gen), _depth(_gen), ..)
      // right to left (DepthImagePlanner.hpp:386-387; pinned against the reference build in the tests)
      double dep = mt.uniform(box.min_depth, box.max_depth);
      double py = mt.uniform(box.min_y, box.max_y);
      double px = mt.uniform(box.min_x, box.max_x);
      pl.deproject(px, py, dep, goal);
      T = mt.uniform(box.min_time, box.max_time);
    }
    if (cands_out) {
      cands_out[4 * i] = goal[0];
      cands_out[4 * i + 1] = goal[1];
      cands_out[4 * i + 2] = goal[2];
      cands_out[4 * i + 3] = T;
    }
    cand.generate(goal, T);
    pl.nGenerated++;
    // cost (DepthImagePlanner.hpp:436-441 / Rappids_Simulator main.cpp:95-109)
    double pe[3] = {cand.ax[0].pos(T), cand.ax[1].pos(T), cand.ax[2].pos(T)};
    double cost;
    if (cfg->cost_kind == 0) {
      cost = -(cfg->cost_vec[0] * pe[0] + cfg->cost_vec[1] * pe[1] + cfg->cost_vec[2] * pe[2]) / T;
    } else {
      double gx = cfg->cost_vec[0] - 0, gy = cfg->cost_vec[1] - 0, gz = cfg->cost_vec[2] - 0;
      double SG = sqrt(gx * gx + gy * gy + gz * gz);
      double dx = cfg->cost_vec[0] - pe[0], dy = cfg->cost_vec[1] - pe[1], dz = cfg->cost_vec[2] - pe[2];
      double PiG = sqrt(dx * dx + dy * dy + dz * dz);
      cost = -(SG - PiG) / T;
    }
    uint8_t res = 0;
    if (cost < best) {
      res |= 1;
      pl.nCost++;
      if (cand.input_feasibility(cfg->min_thrust, cfg->max_thrust, cfg->max_angvel, cfg->min_section_time) ==
          IN_FEASIBLE) {
        res |= 2;
        pl.nColl++;
        if (cand.velocity_feasibility(cfg->max_velocity) == 0) {
          res |= 4;
          pl.nVel++;
          Poly P;
          cand.coeffs(P.c);
          if (pl.collision_free(P, 0, T)) {
            res |= 8;
            found = true;
            best = cost;
            pl.nFree++;
            out->best_index = i;
            out->best_cost = cost;
            out->best_tf = T;
            memcpy(out->best_coeffs, P.c, sizeof(P.c));
          }
        }
      }
    }
    if (results) results[i] = res;
  }
  out->found = found ? 1 : 0;
  if (!found) out->best_cost = std::numeric_limits<double>::max();
  g_pixels_read += t_pixels_read;
  t_pixels_read = 0;
  out->n_generated = pl.nGenerated;
  out->n_cost_checks = pl.nCost;
  out->n_collision_checks = pl.nColl;
  out->n_velocity_checks = pl.nVel;
  out->n_collision_free = pl.nFree;
  out->n_pyramids = (int)pl.pyramids.size();
  if (pyramids) {
    for (int i = 0; i < (int)pl.pyramids.size() && i < max_pyr_out; i++) {
      double* p = pyramids + ORC_RAPPIDS_PYRAMID_DOUBLES * i;
      const Pyr& q = pl.pyramids[i];
      p[0] = q.depth;
      p[1] = q.right;
      p[2] = q.top;
      p[3] = q.left;
      p[4] = q.bottom;
      for (int f = 0; f < 4; f++)
        for (int a = 0; a < 3; a++) p[5 + 3 * f + a] = q.n[f][a];
    }
  }
  return 0;
}

}  // namespace

extern "C" {

const char* orc_rappids_flavour(void) { return ORC_FLAVOUR; }

int orc_rappids_plan(const orc_rappids_cfg* cfg, const uint16_t* image, const double vel0[3],
                     const double acc0[3], const double grav[3], int32_t n_candidates,
                     const double* candidates, int32_t seed, const orc_rappids_sampler* sampler,
                     orc_rappids_out* out, uint8_t* results, double* candidates_out, double* pyramids,
                     int32_t max_pyr_out) {
  return plan_one(cfg, image, vel0, acc0, grav, n_candidates, candidates, seed, sampler, out, results,
                  candidates_out, pyramids, max_pyr_out);
}

int orc_rappids_plan_many(const orc_rappids_cfg* cfg, int32_t n, const uint16_t* images, const double* vel0,
                          const double* acc0, const double* grav, int32_t k, const double* candidates,
                          orc_rappids_out* out, uint8_t* results, int32_t threads) {
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  const size_t npix = (size_t)cfg->width * cfg->height;
  for (int t = 0; t < threads; t++) {
    pool.emplace_back([=]() {
      for (int i = (int)((int64_t)n * t / threads); i < (int)((int64_t)n * (t + 1) / threads); i++)
        plan_one(cfg, images + npix * i, vel0 + 3 * i, acc0 + 3 * i, grav + 3 * i, k, candidates + (size_t)4 * k * i,
                 0, nullptr, out + i, results ? results + (size_t)k * i : nullptr, nullptr, nullptr, 0);
    });
  }
  for (auto& th : pool) th.join();
  return 0;
}

// DepthImagePlanner::IsCollisionFreeGroundTruth (DepthImagePlanner.cpp:1031-1097), restated
int orc_rappids_ground_truth(const orc_rappids_cfg* cfg, const uint16_t* image, const double vel0[3], const double acc0[3],
                             const double grav[3], int32_t n, const double* cands, uint8_t* free_out) {
  const int W = cfg->width, H = cfg->height;
  const double f = cfg->focal_length, cx = cfg->cx, cy = cfg->cy;
  const double timestep = 0.1;
  const uint16_t ignoreDist = uint16_t(cfg->true_radius / cfg->depth_scale);
  const int edge = f * cfg->true_radius / cfg->min_checking_dist;
  for (int i = 0; i < n; i++) {
    const double* c = cands + 4 * i;
    Prim p;
    p.init(vel0, acc0, grav);
    p.generate(c, c[3]);
    Poly P;
    p.coeffs(P.c);
    const double T = c[3];
    bool ok = true;
    for (double t = 0; t < T && ok; t += timestep) {  // field of view first (:1042-1057)
      const double x = P.axis(0, t), y = P.axis(1, t), z = P.axis(2, t);
      if (z < cfg->min_checking_dist) continue;
      const double px = x * f / z + cx, py = y * f / z + cy;
      if (px <= edge || px > W - edge || py <= edge || py > H - edge) ok = false;
    }
    for (double t = 0; t < T && ok; t += timestep) {  // every pixel's ray against the vehicle sphere (:1060-1094)
      const double tp[3] = {P.axis(0, t), P.axis(1, t), P.axis(2, t)};
      if (tp[2] < cfg->min_checking_dist) continue;
      const double n2 = tp[0] * tp[0] + tp[1] * tp[1] + tp[2] * tp[2];
      for (int yy = 0; yy < H && ok; yy++)
        for (int xx = 0; xx < W; xx++) {
          const uint16_t d = image[yy * W + xx];
          if (!(d > ignoreDist)) continue;
          const double v[3] = {(xx - cx) / f, (yy - cy) / f, 1.0};
          const float nrm = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);  // GetUnitVector: the norm is a float (Vec3.hpp:127)
          const double e[3] = {v[0] / nrm, v[1] / nrm, v[2] / nrm};
          const double te = tp[0] * e[0] + tp[1] * e[1] + tp[2] * e[2];
          const double under = pow(te, 2) - n2 + pow(cfg->planning_radius, 2);
          if (under >= 0) {
            const double second = (e[0] * tp[0] + e[1] * tp[1] + e[2] * tp[2]) + sqrt(under);
            const double depth = d * cfg->depth_scale;
            const double q[3] = {depth * ((xx - cx) / f), depth * ((yy - cy) / f), depth * 1};
            const double pd = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
            if (pd < second) {
              ok = false;
              break;
            }
          }
        }
    }
    free_out[i] = ok ? 1 : 0;
  }
  return 0;
}

// pixels read by InflatePyramid in all plans since the last call (port only)
uint64_t orc_rappids_pixels_read(void) { return g_pixels_read.exchange(0); }

int orc_rappids_solve_cubic(double a, double b, double c, double roots[3]) { return (int)cubic(a, b, c, roots); }
int orc_rappids_solve_quartic(double a, double b, double c, double d, double roots[4]) {
  return (int)quartic(a, b, c, d, roots);
}

int orc_rappids_primitive(const double vel0[3], const double acc0[3], const double grav[3], const double goal[3],
                          double T, double fmin, double fmax, double wmax, double min_section, double vmax,
                          double abg[9], int32_t* input_res, int32_t* vel_res) {
  Prim p;
  p.init(vel0, acc0, grav);
  p.generate(goal, T);
  for (int a = 0; a < 3; a++) {
    abg[3 * a + 0] = p.ax[a].al;
    abg[3 * a + 1] = p.ax[a].be;
    abg[3 * a + 2] = p.ax[a].ga;
  }
  *input_res = p.input_feasibility(fmin, fmax, wmax, min_section);
  *vel_res = p.velocity_feasibility(vmax);
  return 0;
}

}  // extern "C"
