// oracle/port/agf_port.cpp -- TEST INFRASTRUCTURE: CPU restatement of the reference hot path.
//
// An independent, literal, single-vehicle restatement of muellerlab/agri-fly's simulation step,
// written from the reference sources cited at each function (paths relative to the reference
// root).  It is the checker that travels to the GPU box as source (the reference itself does
// not); oracle/_ref (the unmodified reference compiled in the build container) pins it:
// tests/test_oracle_vs_ref.py requires bit-identical trajectories for both libm flavours.
// Never linked into, imported by, or executed from the product.
//
// Arithmetic rules followed throughout: same operand order and association as the reference
// expression, float stays float / double stays double, no FMA contraction (-ffp-contract=off),
// dense sequential-k matrix products (the contract of oracle/shim/Eigen/Dense).
// Time is the reference's integer-microsecond Timer (Common/Common/Time/Timer.hpp:19-63).
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <limits>
#include <thread>
#include <deque>
#include <vector>

#include "../oracle_api.h"

#ifdef AGF_PORT_SHARED_MATH
#include "../../agri-fly_b200/csrc/agf_math.h"
#endif

#ifndef ORC_FLAVOUR
#define ORC_FLAVOUR "port-glibc"
#endif

namespace port {

// ---------------------------------------------------------------------------------------------
// libm dispatch (Common/Common/Math/Rotation.hpp:262-316)
// ---------------------------------------------------------------------------------------------
#ifdef AGF_PORT_SHARED_MATH
inline float m_sin(float x) { return agf_sinf(x); }
inline float m_cos(float x) { return agf_cosf(x); }
inline float m_asin(float x) { return agf_asinf(x); }
inline float m_acos(float x) { return agf_acosf(x); }
inline float m_atan2(float y, float x) { return agf_atan2f(y, x); }
inline double m_sin(double x) { return agf_sin(x); }
inline double m_cos(double x) { return agf_cos(x); }
inline double m_asin(double x) { return agf_asin(x); }
inline double m_acos(double x) { return agf_acos(x); }
inline double m_atan2(double y, double x) { return agf_atan2(y, x); }
inline double m_exp(double x) { return agf_exp(x); }  // the offboard estimator's exp only; the plant's motor lag keeps glibc
#else
inline float m_sin(float x) { return sinf(x); }
inline float m_cos(float x) { return cosf(x); }
inline float m_asin(float x) { return asinf(x); }
inline float m_acos(float x) { return acosf(x); }
inline float m_atan2(float y, float x) { return atan2f(y, x); }
inline double m_sin(double x) { return sin(x); }
inline double m_cos(double x) { return cos(x); }
inline double m_asin(double x) { return asin(x); }
inline double m_acos(double x) { return acos(x); }
inline double m_atan2(double y, double x) { return atan2(y, x); }
inline double m_exp(double x) { return exp(x); }
#endif
inline float m_sqrt(float x) { return sqrtf(x); }
inline double m_sqrt(double x) { return sqrt(x); }

// The reference detects acosf domain errors through errno (KalmanFilter6DOF.cpp:81,96,132;
// QuadcopterLogic.cpp:428-437).  glibc raises EDOM exactly when |x| > 1 (NaN raises nothing).
inline bool acos_domain_error(float x) { return x > 1.0f || x < -1.0f; }

// ---------------------------------------------------------------------------------------------
// Vec3 (Common/Common/Math/Vec3.hpp)
// ---------------------------------------------------------------------------------------------
template<typename R>
struct V3 {
  R x, y, z;
  V3() : x(std::numeric_limits<R>::quiet_NaN()), y(x), z(x) {}  // :35 NaN default
  V3(R a, R b, R c) : x(a), y(b), z(c) {}
  template<typename S>
  explicit V3(const V3<S>& o) : x(R(o.x)), y(R(o.y)), z(R(o.z)) {}  // :54-64
  R dot(const V3& r) const { return x * r.x + y * r.y + z * r.z; }  // :101
  V3 cross(const V3& r) const {                                      // :106
    return V3(y * r.z - z * r.y, z * r.x - x * r.z, x * r.y - y * r.x);
  }
  R norm2sq() const { return dot(*this); }
  R norm() const { return m_sqrt(norm2sq()); }  // :117
  V3 unit() const {                             // :126-129: the norm is truncated to float
    float const n = float(norm());
    return (*this) / R(n);
  }
  V3 operator+(const V3& r) const { return V3(x + r.x, y + r.y, z + r.z); }
  V3 operator-(const V3& r) const { return V3(x - r.x, y - r.y, z - r.z); }
  V3 operator/(R s) const { return V3(x / s, y / s, z / s); }
  V3 operator-() const { return V3(R(-1) * x, R(-1) * y, R(-1) * z); }  // :143-145 (*this)*Real(-1)
  R get(int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
// :172-179 both scalar orders compute scalar*component
template<typename R> V3<R> operator*(const V3<R>& v, R s) { return V3<R>(s * v.x, s * v.y, s * v.z); }
template<typename R> V3<R> operator*(R s, const V3<R>& v) { return V3<R>(s * v.x, s * v.y, s * v.z); }
// :183-190 integer scalar
template<typename R> V3<R> muli(const V3<R>& v, int s) { return V3<R>(s * v.x, s * v.y, s * v.z); }

typedef V3<float> V3f;
typedef V3<double> V3d;

template<typename R>
struct M33 {
  R m[3][3];
};
// Vec3.hpp:202-210: accumulate from 0 in j order
template<typename R>
V3<R> mul(const M33<R>& a, const V3<R>& v) {
  R o[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      o[i] += a.m[i][j] * v.get(j);
  return V3<R>(o[0], o[1], o[2]);
}

// ---------------------------------------------------------------------------------------------
// Rotation (Common/Common/Math/Rotation.hpp)
// ---------------------------------------------------------------------------------------------
template<typename R>
struct Rot {
  R v[4];
  Rot() {}
  Rot(R a, R b, R c, R d) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
  static R min_angle() { return R(4.84813681e-6); }  // :39
  static Rot identity() { return Rot(1, 0, 0, 0); }
  Rot inverse() const { return Rot(v[0], -v[1], -v[2], -v[3]); }  // :67
  static Rot from_axis_angle(const V3<R>& u, R angle) {             // :92-97 sin evaluated 3x
    const R half = R(0.5);
    return Rot(m_cos(angle * half), m_sin(angle * half) * u.x, m_sin(angle * half) * u.y,
               m_sin(angle * half) * u.z);
  }
  static Rot from_rotation_vector(const V3<R>& rv) {  // :84-89
    const R theta = rv.norm();
    if (theta < min_angle()) return identity();
    return from_axis_angle(rv / theta, theta);
  }
  static Rot from_euler_ypr(R y, R p, R r) {  // :99-110
    const R half = R(0.5);
    Rot o;
    o.v[0] = m_cos(half * y) * m_cos(half * p) * m_cos(half * r) + m_sin(half * y) * m_sin(half * p) * m_sin(half * r);
    o.v[1] = m_cos(half * y) * m_cos(half * p) * m_sin(half * r) - m_sin(half * y) * m_sin(half * p) * m_cos(half * r);
    o.v[2] = m_cos(half * y) * m_sin(half * p) * m_cos(half * r) + m_sin(half * y) * m_cos(half * p) * m_sin(half * r);
    o.v[3] = m_sin(half * y) * m_cos(half * p) * m_cos(half * r) - m_cos(half * y) * m_sin(half * p) * m_sin(half * r);
    return o;
  }
  // :124-131  (*this) * r1
  Rot operator*(const Rot& r1) const {
    R c0 = r1.v[0] * v[0] - r1.v[1] * v[1] - r1.v[2] * v[2] - r1.v[3] * v[3];
    R c1 = r1.v[1] * v[0] + r1.v[0] * v[1] + r1.v[3] * v[2] - r1.v[2] * v[3];
    R c2 = r1.v[2] * v[0] - r1.v[3] * v[1] + r1.v[0] * v[2] + r1.v[1] * v[3];
    R c3 = r1.v[3] * v[0] + r1.v[2] * v[1] - r1.v[1] * v[2] + r1.v[0] * v[3];
    return Rot(c0, c1, c2, c3);
  }
  V3<R> vector_part() const {  // :155-161
    if (v[0] > 0) return V3<R>(v[1], v[2], v[3]);
    return V3<R>(-v[1], -v[2], -v[3]);
  }
  V3<R> to_rotation_vector() const {  // :144-153
    const V3<R> n = vector_part();
    const R norm = n.norm();
    const R angle = m_asin(norm) * 2;
    if (angle < min_angle()) return V3<R>(0, 0, 0);
    return n * (angle / norm);
  }
  void to_euler_ypr(R& y, R& p, R& r) const {  // :163-169
    y = m_atan2(R(2.0) * v[1] * v[2] + R(2.0) * v[0] * v[3],
                v[1] * v[1] + v[0] * v[0] - v[3] * v[3] - v[2] * v[2]);
    p = -m_asin(R(2.0) * v[1] * v[3] - R(2.0) * v[0] * v[2]);
    r = m_atan2(R(2.0) * v[2] * v[3] + R(2.0) * v[0] * v[1],
                v[3] * v[3] - v[2] * v[2] - v[1] * v[1] + v[0] * v[0]);
  }
  void matrix(R Rm[9]) const {  // :196-217
    const R r0 = v[0] * v[0], r1 = v[1] * v[1], r2 = v[2] * v[2], r3 = v[3] * v[3];
    Rm[0] = r0 + r1 - r2 - r3;
    Rm[1] = 2 * v[1] * v[2] - 2 * v[0] * v[3];
    Rm[2] = 2 * v[1] * v[3] + 2 * v[0] * v[2];
    Rm[3] = 2 * v[1] * v[2] + 2 * v[0] * v[3];
    Rm[4] = r0 - r1 + r2 - r3;
    Rm[5] = 2 * v[2] * v[3] - 2 * v[0] * v[1];
    Rm[6] = 2 * v[1] * v[3] - 2 * v[0] * v[2];
    Rm[7] = 2 * v[2] * v[3] + 2 * v[0] * v[1];
    Rm[8] = r0 - r1 - r2 + r3;
  }
  M33<R> matrix33() const {
    R Rm[9];
    matrix(Rm);
    M33<R> o;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        o.m[i][j] = Rm[3 * i + j];
    return o;
  }
  V3<R> rotate(const V3<R>& in) const {  // :236-245
    R Rm[9];
    matrix(Rm);
    return V3<R>(Rm[0] * in.x + Rm[1] * in.y + Rm[2] * in.z, Rm[3] * in.x + Rm[4] * in.y + Rm[5] * in.z,
                 Rm[6] * in.x + Rm[7] * in.y + Rm[8] * in.z);
  }
};
typedef Rot<float> Rotf;
typedef Rot<double> Rotd;

// ---------------------------------------------------------------------------------------------
// Timer (Common/Common/Time/Timer.hpp)
// ---------------------------------------------------------------------------------------------
struct Clock {
  uint64_t now_us;
};
struct Timer {
  const Clock* master;
  uint64_t last;
  explicit Timer(const Clock* c) : master(c) { reset(); }
  void reset() { last = master->now_us; }
  uint64_t micros() const { return master->now_us - last; }
  double seconds_d() const { return double(micros() * double(1e-6)); }  // :36-39
  float seconds_f() const { return float(micros() * float(1e-6)); }
  template<typename R>
  void adjust_by_seconds(R s) {  // :27-34
    if (s > 0) {
      last -= uint64_t(s * R(1e6));
    } else {
      last += uint64_t(s * R(-1e6));
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Low-pass filters (Common/Common/Math/LowPassFilter{First,Second}Order.hpp)
// ---------------------------------------------------------------------------------------------
inline V3f scale(float s, const V3f& v) { return s * v; }
inline float scale(float s, float v) { return s * v; }

template<typename S>
struct LPF2 {
  float a1, a2, b0, b1, b2;
  S xm0, xm1, ym0, ym1;
  LPF2() : a1(0), a2(0), b0(1), b1(0), b2(0) {}
  void init(float dt, float wc, S v0) {  // LowPassFilterSecondOrder.hpp:23-47
    float const sqrt2 = float(sqrt(2.0));
    a1 = (dt * dt * wc * wc - 2 * sqrt2 * dt * wc + 4) / (dt * dt * wc * wc + 2 * sqrt2 * dt * wc + 4);
    a2 = 2 * (dt * dt * wc * wc - 4) / (dt * dt * wc * wc + 2 * sqrt2 * dt * wc + 4);
    b0 = dt * dt * wc * wc / (dt * dt * wc * wc + 2 * sqrt2 * dt * wc + 4);
    b1 = dt * dt * wc * wc / (dt * dt * wc * wc + 2 * sqrt2 * dt * wc + 4);
    b2 = 2 * dt * dt * wc * wc / (dt * dt * wc * wc + 2 * sqrt2 * dt * wc + 4);
    xm0 = xm1 = ym0 = ym1 = v0;
  }
  S apply(S in) {  // :51-63
    S out = scale(b2, in);
    out = out + (scale(+b0, xm0) + scale(b1, xm1));
    out = out + (scale(-a1, ym0) - scale(a2, ym1));
    xm0 = xm1;
    xm1 = in;
    ym0 = ym1;
    ym1 = out;
    return out;
  }
  S value() const { return ym1; }
};

struct LPF1 {  // LowPassFilterFirstOrder.hpp, TYPE_RATE = TYPE_SAMPLE = float
  float coeff, prev;
  LPF1() : coeff(0.0f), prev(0) {}
  void init(float period, float cutoff, float v0) {  // :16-32; exp(float) resolves to the float overload
    prev = v0;
    coeff = expf(-period * cutoff);
  }
  float apply(float in) {  // :34-49
    if (coeff <= 0.0f) {
      prev = in;
      return in;
    }
    float out = coeff * prev + (1 - coeff) * in;
    prev = out;
    return out;
  }
};

// ---------------------------------------------------------------------------------------------
// Radio uplink codec (Common/Common/DataTypes/RadioTypes.hpp:103-116,189-240)
// ---------------------------------------------------------------------------------------------
struct RadioMsg {
  uint8_t type, flags;
  float f[10];
  RadioMsg() : type(0), flags(0) {
    for (int i = 0; i < 10; i++) f[i] = 0;  // uninitialised in the reference; never read before set
  }
};

inline float radio_field(const uint8_t* raw, unsigned idx, float limit) {
  int out = 0;
  for (int i = 0; i < 2; i++) {
    if (idx + i >= 23) break;
    out += raw[idx + i] << ((2 - 1 - i) * 8);
  }
  return limit * (out - 32768) / float(32768);
}

inline RadioMsg radio_decode(const uint8_t raw[23]) {
  RadioMsg m;
  m.type = raw[0];
  m.flags = raw[2];
  switch (m.type) {
    case 3:  // positionCommand
      for (int i = 0; i < 3; i++) m.f[i] = radio_field(raw, 3 + i * 2, 20);
      for (int i = 3; i < 6; i++) m.f[i] = radio_field(raw, 3 + i * 2, 10);
      for (int i = 6; i < 9; i++) m.f[i] = radio_field(raw, 3 + i * 2, 30);
      break;
    case 5:  // externalRatesCmd
      m.f[0] = radio_field(raw, 3, 35);
      for (int i = 1; i < 10; i++) m.f[i] = radio_field(raw, 3 + i * 2, 35);
      break;
    case 4:  // externalAccelerationCmd
      for (int i = 0; i < 3; i++) m.f[i] = radio_field(raw, 3 + i * 2, 30);
      m.f[3] = radio_field(raw, 3 + 3 * 2, 35);
      break;
    default:
      for (int i = 0; i < 10; i++) m.f[i] = radio_field(raw, 3 + i * 2, 1);
      break;
  }
  return m;
}

// ---------------------------------------------------------------------------------------------
// Telemetry encode (Common/Common/DataTypes/TelemetryPacket.hpp:39-63,122-166)
// ---------------------------------------------------------------------------------------------
inline float map_to_ones(float x, float a, float b) { return ((x - a) / (b - a)) * 2 - 1; }
inline uint16_t encode_ones(float t) {
  if (t < -1 || t > 1) return 0;
  float e = 32768 + 32767 * t;
  if (!(e == e)) return 0;  // NaN: x86-64 float->int conversion yields 0x80000000, low 16 bits 0
  return uint16_t(int(e));
}

// ---------------------------------------------------------------------------------------------
// Controllers and mixer (Components/Components/Logic/Quadcopter*Controller.hpp, QuadcopterMixer.hpp)
// ---------------------------------------------------------------------------------------------
struct PositionController {
  float natFreq, damping;
  V3f des_acceleration(V3f estPos, V3f estVel, V3f desPos, V3f desVel = V3f(0, 0, 0), V3f desAcc = V3f(0, 0, 0)) const {  // QuadcopterPositionController.hpp:22-27
    return (desPos - estPos) * natFreq * natFreq + muli(desVel - estVel, 2) * natFreq * damping + desAcc;
  }
};

struct AttitudeController {
  float tc_xy, tc_z;
  V3f desired_angular_velocity(const Rotf& desAtt, const Rotf& estAtt) const {  // QuadcopterAttitudeController.hpp:35-68
    Rotf errAtt = desAtt.inverse() * estAtt;
    const V3f desRotVec = errAtt.to_rotation_vector();
    V3f redAx = errAtt.inverse().rotate(V3f(0, 0, 1)).cross(V3f(0, 0, 1));
    float c = errAtt.inverse().rotate(V3f(0, 0, 1)).dot(V3f(0, 0, 1));
    float redAn;
    if (c >= 1.0f) {
      redAn = 0;
    } else if (c <= -1.0f) {
      redAn = float(M_PI);
    } else {
      redAn = m_acos(c);
    }
    float n = redAx.norm();
    if (n < 1e-12f) {
      redAx = V3f(0, 0, 0);
    } else {
      redAx = redAx / n;
    }
    float k3 = (1.0f / tc_z);
    float k12 = (1.0f / tc_xy);
    return -k3 * desRotVec - (k12 - k3) * redAn * redAx;
  }
};

struct AngVelController {
  float tc_xy, tc_z;
  M33<float> I;
  V3f desired_torques(const V3f& des, const V3f& est) const {  // QuadcopterAngularVelocityController.hpp:25-37
    V3f err = des - est;
    V3f acc(err.x / tc_xy, err.y / tc_xy, err.z / tc_z);
    V3f nonlin = est.cross(mul(I, est));
    return mul(I, acc) + nonlin;
  }
};

struct Mixer {
  float d, kt, kf, maxCmdTotal, minPer, maxPer, corr[4];
  Mixer() : d(0), kt(0), kf(0) {
    minPer = 0.0f;
    maxPer = 1000000.0f;
    maxCmdTotal = 4 * maxPer;
    for (int i = 0; i < 4; i++) corr[i] = 1.0f;
  }
  void set(float arm, float kF, float kTau, int spin, float maxP, float minP, float maxTot) {  // QuadcopterMixer.hpp:36-51
    d = arm / sqrtf(2.0f);
    kt = spin * kTau;
    kf = kF;
    maxPer = maxP;
    minPer = minP;
    if (maxTot < 0) {
      maxCmdTotal = 4 * maxP * 0.8f;
    } else {
      maxCmdTotal = maxTot;
    }
  }
  void motor_forces(float totF, const V3f& t, float out[4]) const {  // :63-86
    float desF = totF > maxCmdTotal ? maxCmdTotal : totF;
    out[0] = (-t.x / d - t.y / d - t.z / kt + desF) / 4.0f;
    out[1] = (-t.x / d + t.y / d + t.z / kt + desF) / 4.0f;
    out[2] = (+t.x / d + t.y / d - t.z / kt + desF) / 4.0f;
    out[3] = (+t.x / d - t.y / d + t.z / kt + desF) / 4.0f;
    for (int i = 0; i < 4; i++) {
      if (out[i] < minPer) {
        out[i] = minPer;
      } else if (out[i] > maxPer) {
        out[i] = maxPer;
      }
    }
  }
  void speeds_from_thrust(const float thrusts[4], float out[4]) const {  // :88-99
    for (int i = 0; i < 4; i++) {
      if (thrusts[i] <= 0) {
        out[i] = 0;
        continue;
      }
      out[i] = sqrtf(thrusts[i] / (corr[i] * kf));
    }
  }
  float uncorrected_force(float s) const { return kf * s * s; }  // :102-104
};

// ---------------------------------------------------------------------------------------------
// KalmanFilter6DOF (Components/Components/Logic/KalmanFilter6DOF.{hpp,cpp})
// ---------------------------------------------------------------------------------------------
struct Mat9 {
  float m[9][9];
};
inline Mat9 matmul(const Mat9& a, const Mat9& b) {  // shim contract: sequential k from the k=0 product
  Mat9 r;
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 9; j++) {
      float acc = a.m[i][0] * b.m[0][j];
      for (int k = 1; k < 9; k++) acc += a.m[i][k] * b.m[k][j];
      r.m[i][j] = acc;
    }
  return r;
}
inline Mat9 transpose(const Mat9& a) {
  Mat9 r;
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 9; j++) r.m[j][i] = a.m[i][j];
  return r;
}

struct KalmanFilter {
  enum { I_POS = 0, I_VEL = 3, I_ATT = 6, N = 9 };
  Timer estimateTimer, lastGoodMeas;
  bool imuInit, uwbInit;
  V3f pos, vel, angVel;
  Rotf att;
  V3f lastCorr;
  Mat9 cov;
  float initStdPos, initStdVel, initStdAttAboutG, initStdAttPerpG;
  float sGyro, sAcc, sRange, outlierDist;
  unsigned numRejected, maxRejectedSeq, numRejectedSeq, numResets, lastCheckResets;

  explicit KalmanFilter(const Clock* c)  // KalmanFilter6DOF.cpp:9-31
      : estimateTimer(c), lastGoodMeas(c), imuInit(false), uwbInit(false) {
    initStdPos = 3.0f;
    initStdVel = 3.0f;
    initStdAttPerpG = 10.0f * float(M_PI) / 180.0f;
    initStdAttAboutG = 30.0f * float(M_PI) / 180.0f;
    sAcc = 5;
    sGyro = 0.1f;
    sRange = 0.14f;
    outlierDist = 3.0f;
    numRejected = 0;
    numRejectedSeq = 0;
    maxRejectedSeq = 5;
    numResets = 0;
    lastCheckResets = 0;
    memset(&cov, 0, sizeof(cov));  // (uninitialised until Reset() in the reference)
    pos = vel = angVel = lastCorr = V3f(0, 0, 0);
    att = Rotf::identity();
  }

  void reset() {  // :33-68
    numResets++;
    imuInit = uwbInit = false;
    pos = V3f(0, 0, 0);
    vel = V3f(0, 0, 0);
    att = Rotf::identity();
    angVel = V3f(0, 0, 0);
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) cov.m[i][j] = 0;
    for (int i = 0; i < 3; i++) {
      cov.m[I_POS + i][I_POS + i] = initStdPos * initStdPos;
      cov.m[I_VEL + i][I_VEL + i] = initStdVel * initStdVel;
    }
    cov.m[I_ATT + 0][I_ATT + 0] = initStdAttPerpG * initStdAttPerpG;
    cov.m[I_ATT + 1][I_ATT + 1] = initStdAttPerpG * initStdAttPerpG;
    cov.m[I_ATT + 2][I_ATT + 2] = initStdAttAboutG * initStdAttAboutG;
    estimateTimer.reset();
    lastGoodMeas.reset();
    lastCorr = V3f(0, 0, 0);
  }

  bool was_reset_since_last_check() {  // KalmanFilter6DOF.hpp:59-63
    bool change = lastCheckResets != numResets;
    lastCheckResets = numResets;
    return change;
  }

  static float acos_guarded(float c) {
    float angle = m_acos(c);
    if (acos_domain_error(c)) angle = c < 0 ? float(M_PI) : 0.0f;
    return angle;
  }

  void predict(const V3f measGyro, const V3f measAcc) {  // :70-241
    if (!imuInit) {
      reset();
      imuInit = true;
      estimateTimer.reset();
      V3f const expAcc = att.inverse().rotate(V3f(0, 0, 1));
      V3f const accUnit = measAcc.unit();
      float const cosErr = expAcc.dot(accUnit);
      V3f rotAx = accUnit.cross(expAcc);
      if (rotAx.norm() > 1e-6f) {
        rotAx = rotAx / rotAx.norm();
      } else {
        rotAx = V3f(1, 0, 0);
      }
      float angle = acos_guarded(cosErr);
      att = att * Rotf::from_axis_angle(rotAx, angle);
      return;
    }
    float const dt = estimateTimer.seconds_f();
    estimateTimer.reset();

    if (!uwbInit) {  // :114-147 complementary filter
      angVel = measGyro;
      Rotf newAtt = att * Rotf::from_rotation_vector(measGyro * dt);
      att = newAtt;
      V3f const expAcc = att.inverse().rotate(V3f(0, 0, 1));
      V3f const accUnit = measAcc.unit();
      V3f rotAx = accUnit.cross(expAcc);
      if (rotAx.norm() > 1e-6f) {
        rotAx = rotAx / rotAx.norm();
      } else {
        rotAx = V3f(1, 0, 0);
      }
      float const cosErr = expAcc.dot(accUnit);
      float angle = acos_guarded(cosErr);
      float const corrAngle = (dt / 4.0f) * angle;  // TIME_CONST_ATT_CORR :7
      att = att * Rotf::from_axis_angle(rotAx, corrAngle);
      return;
    }

    // :149-241 full prediction
    V3f const p0(pos), v0(vel);
    Rotf const a0(att);
    V3f const acc = att.rotate(measAcc) + V3f(0, 0, -9.81f);
    pos = p0 + v0 * dt;
    vel = v0 + acc * dt;
    att = a0 * Rotf::from_rotation_vector(measGyro * dt);
    angVel = measGyro;

    float Rm[9];
    a0.matrix(Rm);
    Mat9 f;
    memset(&f, 0, sizeof(f));
    for (int i = 0; i < 3; i++) {
      f.m[I_POS + i][I_POS + i] = 1;
      f.m[I_POS + i][I_VEL + i] = dt;
      f.m[I_VEL + i][I_VEL + i] = 1;
    }
    for (int r = 0; r < 3; r++) {  // :184-209
      f.m[I_VEL + r][I_ATT + 0] = dt * (+measAcc.y * Rm[3 * r + 2] - measAcc.z * Rm[3 * r + 1]);
      f.m[I_VEL + r][I_ATT + 1] = dt * (-measAcc.x * Rm[3 * r + 2] + measAcc.z * Rm[3 * r + 0]);
      f.m[I_VEL + r][I_ATT + 2] = dt * (+measAcc.x * Rm[3 * r + 1] - measAcc.y * Rm[3 * r + 0]);
    }
    f.m[I_ATT + 0][I_ATT + 0] = 1;  // :212-228
    f.m[I_ATT + 1][I_ATT + 0] = -(dt * measGyro.z + lastCorr.z / 2.0f);
    f.m[I_ATT + 2][I_ATT + 0] = +(dt * measGyro.y + lastCorr.y / 2.0f);
    f.m[I_ATT + 0][I_ATT + 1] = +(dt * measGyro.z + lastCorr.z / 2.0f);
    f.m[I_ATT + 1][I_ATT + 1] = 1;
    f.m[I_ATT + 2][I_ATT + 1] = -(dt * measGyro.x + lastCorr.x / 2.0f);
    f.m[I_ATT + 0][I_ATT + 2] = -(dt * measGyro.y + lastCorr.y / 2.0f);
    f.m[I_ATT + 1][I_ATT + 2] = +(dt * measGyro.x + lastCorr.x / 2.0f);
    f.m[I_ATT + 2][I_ATT + 2] = 1;
    lastCorr = V3f(0, 0, 0);

    cov = matmul(matmul(f, cov), transpose(f));  // :232
    for (int i = 0; i < 3; i++) {                // :234-239
      cov.m[I_VEL + i][I_VEL + i] += sAcc * sAcc * dt * dt;
      cov.m[I_ATT + i][I_ATT + i] += sGyro * sGyro * dt * dt;
    }
  }

  void update_with_range(const V3f target, float const range) {  // :243-301
    if (!imuInit) return;
    if (!(range == range)) return;
    uwbInit = true;
    float const expRange = (pos - target).norm();
    V3f const dir = (pos - target) / expRange;
    float H[9];
    for (int i = 0; i < 3; i++) {
      H[I_POS + i] = dir.get(i);
      H[I_VEL + i] = 0;
      H[I_ATT + i] = 0;
    }
    // P*H' (9x1), sequential k
    float PHt[9];
    for (int i = 0; i < 9; i++) {
      float acc = cov.m[i][0] * H[0];
      for (int k = 1; k < 9; k++) acc += cov.m[i][k] * H[k];
      PHt[i] = acc;
    }
    float hph = H[0] * PHt[0];
    for (int k = 1; k < 9; k++) hph += H[k] * PHt[k];
    float const innovCov = hph + sRange * sRange;
    float L[9];
    float const inv = (1 / innovCov);
    for (int i = 0; i < 9; i++) L[i] = PHt[i] * inv;

    float const d2 = (range - expRange) * (range - expRange) / innovCov;
    if (d2 > outlierDist * outlierDist) {
      numRejected++;
      numRejectedSeq++;
      if (numRejectedSeq >= maxRejectedSeq) reset();
      return;
    }
    numRejectedSeq = 0;

    float dx[9];
    for (int i = 0; i < 9; i++) dx[i] = L[i] * (range - expRange);
    pos = pos + V3f(dx[0], dx[1], dx[2]);
    vel = vel + V3f(dx[3], dx[4], dx[5]);
    lastCorr = V3f(dx[6], dx[7], dx[8]);
    att = att * Rotf::from_rotation_vector(lastCorr);

    Mat9 A;  // I - L*H
    for (int i = 0; i < 9; i++)
      for (int j = 0; j < 9; j++) A.m[i][j] = (i == j ? 1.0f : 0.0f) - L[i] * H[j];
    cov = matmul(A, cov);
    for (int i = 0; i < 9; i++)  // :303-309 lower -> upper
      for (int j = i + 1; j < 9; j++) cov.m[i][j] = cov.m[j][i];
    lastGoodMeas.reset();
  }
};

// ---------------------------------------------------------------------------------------------
// QuadcopterLogic (Components/Components/Logic/QuadcopterLogic.{hpp,cpp})
// ---------------------------------------------------------------------------------------------
struct PeriodMonitor {  // QuadcopterLogic.hpp:383-399
  float lpDt;
  LPF1 filter;
  Timer since;
  PeriodMonitor(const Clock* c, float samplePeriod, float cutoff, float expected) : since(c) {
    filter.init(samplePeriod, cutoff, expected);
    lpDt = expected;
  }
  void update() {
    float dt = since.seconds_f();
    lpDt = filter.apply(dt);
    since.adjust_by_seconds(-dt);
  }
};

struct Logic {
  enum { FS_UNINIT = 0, FS_IDLE, FS_AUTO, FS_PANIC, FS_KILLED, FS_EXT_ACC, FS_EXT_RATES };
  int state;
  Timer timer;
  unsigned cycle;
  const float onboardPeriod, radioCmdPeriod;
  PositionController posCtrl;
  AttitudeController attCtrl;
  AngVelController angVelCtrl;
  Mixer mixer;
  KalmanFilter kf;
  V3f desPos;
  float desSpeeds[4], desForces[4], mass;
  M33<float> Rimu;
  struct { bool isNew; unsigned count; float vRaw, iRaw, vFilt; } batt;
  struct { bool isNew; unsigned count; V3f raw; LPF2<V3f> lp; } gyro, acc;
  struct { bool isNew; unsigned count; float raw; LPF2<float> lp; } temp;
  struct { bool isNew; unsigned count; uint8_t targetId; float range; bool failure; } uwb;
  Timer sinceUwb;
  bool shouldStartUwb;
  struct { bool isNew; unsigned count; RadioMsg msg; } radio;
  Timer sinceRadio;
  struct Target { uint8_t id; V3f p; unsigned good, bad; float last; } targets[32];
  unsigned numTargets;
  uint8_t nextTargetIdx;
  bool gyroCalEnabled;
  unsigned gyroCalN;
  V3f gyroCalAcc, gyroCalBias;
  bool testMotors;
  float testMotorsFrac;
  float debug[6];
  int firstPanic;
  struct { bool running; float active[4], accum[4]; unsigned count; float fmin, fmax; unsigned minCount; } propCal;
  uint8_t warnings;
  PeriodMonitor monCmd, monLoop;
  Timer sinceEstReset;
  bool shouldWriteParams;
  uint8_t myId;
  uint32_t telCounter;
  float battCritical, battWarning;
  LPF2<float> battLp;

  Logic(const Clock* c, float period)  // QuadcopterLogic.cpp:6-20
      : timer(c), onboardPeriod(period), radioCmdPeriod(0.02f), kf(c), sinceUwb(c), sinceRadio(c),
        monCmd(c, 0.02f, 1, 0.02f), monLoop(c, period, 50, period), sinceEstReset(c),
        battCritical(1e3f), battWarning(1e3f) {
    uwb.failure = false;
    batt.vFilt = 0;
    reset_counters();
  }

  void reset_counters() {  // :22-95
    cycle = 0;
    for (int i = 0; i < 4; i++) desSpeeds[i] = desForces[i] = 0;
    state = FS_UNINIT;
    mass = 0;
    batt.isNew = false; batt.count = 0; batt.vRaw = 0; batt.iRaw = 0;
    gyro.isNew = false; gyro.count = 0; gyro.raw = V3f(0, 0, 0);
    acc.isNew = false; acc.count = 0; acc.raw = V3f(0, 0, 0);
    temp.isNew = false; temp.count = 0; temp.raw = 25;
    radio.isNew = false; radio.count = 0;
    uwb.count = 0; uwb.isNew = false; uwb.range = 0; uwb.targetId = 0;
    shouldStartUwb = false;
    numTargets = 0;
    nextTargetIdx = 0;
    desPos = V3f(0, 0, 0.5f);
    testMotors = false;
    testMotorsFrac = 0;
    battCritical = 0;
    shouldWriteParams = false;
    propCal.running = false;
    for (int i = 0; i < 4; i++) { propCal.active[i] = 1.0f; propCal.accum[i] = 0.0f; }
    propCal.count = 0;
    propCal.minCount = 750;
    propCal.fmin = 0.7f;
    propCal.fmax = 1.0f / propCal.fmin;
    for (int i = 0; i < 6; i++) debug[i] = 0;
    firstPanic = 0;
    myId = 0;
    telCounter = 0;
    warnings = 0;
    gyroCalEnabled = false;  // ResetGyroCalibration QuadcopterLogic.hpp:139-144
    gyroCalN = 0;
    gyroCalAcc = V3f(0, 0, 0);
    gyroCalBias = V3f(0, 0, 0);
  }

  void initialise(const agf_logic_consts& k, uint8_t vehId) {  // :97-162
    reset_counters();
    const float accCut = 100.0f, gyroCut = 200.0f;
    const float battCut = 0.5f * float(2 * M_PI), tempCut = 0.5f * float(2 * M_PI);
    myId = vehId;
    mass = k.mass;
    Rimu = Rotf::from_euler_ypr(k.imu_yaw, k.imu_pitch, k.imu_roll).matrix33();
    battCritical = k.low_battery_threshold;
    battWarning = 1.05f * battCritical;
    acc.lp.init(onboardPeriod, accCut, acc.raw);
    gyro.lp.init(onboardPeriod, gyroCut, gyro.raw);
    temp.lp.init(onboardPeriod, tempCut, temp.raw);
    battLp.init(onboardPeriod, battCut, k.low_battery_threshold * 1.2f);
    posCtrl.natFreq = k.pos_control_nat_freq;
    posCtrl.damping = k.pos_control_damping;
    attCtrl.tc_xy = k.att_control_time_const_xy;
    attCtrl.tc_z = k.att_control_time_const_z;
    if (attCtrl.tc_z < attCtrl.tc_xy) attCtrl.tc_z = attCtrl.tc_xy;  // QuadcopterAttitudeController.hpp:19-24
    angVelCtrl.tc_xy = k.ang_vel_control_time_const_xy;
    angVelCtrl.tc_z = k.ang_vel_control_time_const_z;
    memset(&angVelCtrl.I, 0, sizeof(angVelCtrl.I));
    angVelCtrl.I.m[0][0] = k.inertia_xx;
    angVelCtrl.I.m[1][1] = k.inertia_xx;
    angVelCtrl.I.m[2][2] = k.inertia_zz;
    mixer.set(k.arm_length, k.prop_thrust_from_speed_sqr, k.prop_torque_from_thrust, k.prop0_spin_dir,
              k.max_thrust_per_propeller, k.min_thrust_per_propeller, k.max_cmd_total_thrust);
    timer.reset();
    if (k.valid) {
      state = FS_IDLE;
    } else {
      state = FS_KILLED;
      firstPanic = 6;
    }
    kf.reset();
  }

  // QuadcopterLogic.hpp:32-69
  void set_battery(float v, float i) {
    batt.isNew = true; batt.count++; batt.vRaw = v; batt.iRaw = i;
    batt.vFilt = battLp.apply(v);
  }
  void set_gyro(float x, float y, float z) {
    gyro.isNew = true; gyro.count++;
    gyro.raw = mul(Rimu, V3f(x, y, z));
    gyro.lp.apply(gyro.raw - gyroCalBias);
  }
  void set_acc(float x, float y, float z) {
    acc.isNew = true; acc.count++;
    acc.raw = mul(Rimu, V3f(x, y, z));
    acc.lp.apply(acc.raw);
  }
  void set_temp(float t) {
    temp.isNew = true; temp.count++; temp.raw = t;
    temp.lp.apply(temp.raw);
  }
  void set_uwb(float range, uint8_t responder, bool failure) {
    sinceUwb.reset();
    uwb.isNew = true; uwb.targetId = responder; uwb.range = range; uwb.failure = failure;
  }
  void set_radio(const RadioMsg& m) {  // :110-116
    radio.isNew = true; radio.count++; radio.msg = m;
    sinceRadio.reset();
    monCmd.update();
  }
  bool motors_running() const {
    for (int i = 0; i < 4; i++)
      if (desSpeeds[i] > 0) return true;
    return false;
  }
  uint8_t next_ranging_target() const {  // :181-186
    if (!numTargets) return 0;
    return targets[nextTargetIdx].id;
  }
  int add_target(uint8_t id, V3f p) {  // :220-232
    if (numTargets >= 32) return -1;
    targets[numTargets].id = id; targets[numTargets].p = p; targets[numTargets].last = -1.0f;
    targets[numTargets].bad = 0; targets[numTargets].good = 0;
    numTargets++;
    return 0;
  }

  void run() {  // QuadcopterLogic.cpp:164-219
    if (state == FS_UNINIT) return;
    cycle++;
    monLoop.update();
    update_estimator();
    parse_comms();
    update_warnings();
    check_panic();
    debug[0] = temp.lp.value();
    if (testMotors) {
      V3f const t2 = angVelCtrl.desired_torques(V3f(0, 0, 0), kf.angVel);
      mixer.motor_forces(testMotorsFrac * 9.81f * mass, t2, desForces);
      mixer.speeds_from_thrust(desForces, desSpeeds);
      return;
    }
    switch (state) {
      case FS_AUTO: ctrl_autonomous(); return;
      case FS_EXT_ACC: ctrl_ext_acc(); return;
      case FS_EXT_RATES: ctrl_ext_rates(); return;
      default: break;
    }
    for (int i = 0; i < 4; i++) desSpeeds[i] = desForces[i] = 0;
  }

  void update_estimator() {  // :221-273
    if (gyro.isNew && acc.isNew) {
      kf.predict(gyro.lp.value(), acc.lp.value());
      if (gyroCalEnabled) {
        gyroCalAcc = gyroCalAcc + gyro.raw;
        gyroCalN++;
      }
      gyro.isNew = false;
      acc.isNew = false;
    }
    shouldStartUwb = false;
    if (cycle == 100) shouldStartUwb = true;
    if (uwb.isNew) {
      uwb.isNew = false;
      shouldStartUwb = true;
      for (unsigned i = 0; i < numTargets; i++) {  // UpdateRangingDiagnostics :602-618
        if (uwb.targetId == targets[i].id) {
          if (uwb.failure) { targets[i].bad++; targets[i].last = -1.0f; }
          else { targets[i].good++; targets[i].last = uwb.range; }
          break;
        }
      }
      if (uwb.failure) {
        nextTargetIdx = (nextTargetIdx + 1) % numTargets;
      } else {
        uwb.count++;
        nextTargetIdx = (nextTargetIdx + 1) % numTargets;
        for (unsigned i = 0; i < numTargets; i++) {  // GetRangingTargetPosition :590-600
          if (uwb.targetId == targets[i].id) {
            kf.update_with_range(targets[i].p, uwb.range);
            break;
          }
        }
      }
    }
  }

  void parse_comms() {  // :275-303
    if (!radio.isNew) return;
    radio.isNew = false;
    if (state == FS_PANIC || state == FS_KILLED) return;
    switch (radio.msg.type) {
      case 2: state = FS_KILLED; if (!firstPanic) firstPanic = 7; break;
      case 3: state = FS_AUTO; break;
      case 4: state = FS_EXT_ACC; break;
      case 5: state = FS_EXT_RATES; break;
      case 6: state = FS_IDLE; break;
    }
  }

  void update_warnings() {  // :305-342
    batt.isNew = false;
    if (batt.vFilt <= battWarning) warnings |= 0x01;
    if (fabsf(monCmd.lpDt - radioCmdPeriod) > (0.1f * radioCmdPeriod)) warnings |= 0x02;
    if (sinceRadio.seconds_f() > 3 * radioCmdPeriod) warnings |= 0x10;
    if (fabsf(monLoop.lpDt - onboardPeriod) > (0.05f * onboardPeriod)) warnings |= 0x08;
    if (kf.was_reset_since_last_check()) sinceEstReset.reset();
    if (sinceEstReset.seconds_f() < 0.02f) warnings |= 0x04;
  }

  bool safety_critical() const {  // QuadcopterLogic.hpp:253-264
    return !(state == FS_UNINIT || state == FS_IDLE || state == FS_PANIC || state == FS_KILLED);
  }

  void check_panic() {  // :344-391
    V3f const estPos = kf.pos;
    Rotf const estAtt = kf.att;
    int unsafe = 0;
    if (motors_running()) {
      if ((estPos.z < -2.0f) && !(radio.msg.flags & 0x02)) unsafe = 1;
      if ((sinceUwb.micros() > 1500u * 1000u) && (state == FS_AUTO)) unsafe = 2;
      if (estAtt.rotate(V3f(0, 0, 1)).z < 0 && !(radio.msg.flags & 0x02)) unsafe = 3;
      if (sinceRadio.micros() > 1500u * 1000u) unsafe = 4;
      if (batt.vFilt <= battCritical) unsafe = 5;
    }
    if (unsafe && safety_critical()) {
      if (state != FS_PANIC) {
        state = FS_PANIC;
        firstPanic = unsafe;
      }
    }
  }

  static Rotf attitude_from_thrust_dir(const V3f& dir) {  // :423-445 / :485-507
    V3f const e3(0, 0, 1);
    const float cosAngle = dir.dot(e3);
    float angle = m_acos(cosAngle);
    if (acos_domain_error(cosAngle)) angle = cosAngle < 0 ? float(M_PI) : 0.0f;
    V3f rotAx = e3.cross(dir);
    const float n = rotAx.norm();
    if (n < 1e-6f) return Rotf::identity();
    return Rotf::from_rotation_vector(rotAx * (angle / n));
  }

  void ctrl_autonomous() {  // :393-457
    V3f const estPos = kf.pos, estVel = kf.vel, estAngVel = kf.angVel;
    Rotf const estAtt = kf.att;
    desPos = V3f(radio.msg.f[0], radio.msg.f[1], radio.msg.f[2]);
    V3f const desAcc = posCtrl.des_acceleration(estPos, estVel, desPos);
    V3f const proper = desAcc + V3f(0, 0, 9.81f);
    float const nProper = proper.norm();
    V3f const dir = proper / nProper;
    float const corr = estAtt.rotate(V3f(0, 0, 1)).z;
    float const corrSat = corr < 1.00f ? 1.00f : corr;
    const float thrust = nProper / corrSat;
    Rotf desAtt = attitude_from_thrust_dir(dir);
    V3f const desW = attCtrl.desired_angular_velocity(desAtt, estAtt);
    V3f const tq = angVelCtrl.desired_torques(desW, estAngVel);
    mixer.motor_forces(thrust * mass, tq, desForces);
    mixer.speeds_from_thrust(desForces, desSpeeds);
  }

  void ctrl_ext_acc() {  // :459-526
    Rotf const estAtt = kf.att;
    V3f const estAngVel = kf.angVel;
    V3f const desAcc(radio.msg.f[0], radio.msg.f[1], radio.msg.f[2]);
    float const yawRate = radio.msg.f[3];
    if (desAcc.z < -9.81f / 2) {
      for (int i = 0; i < 4; i++) desSpeeds[i] = desForces[i] = 0;
      return;
    }
    V3f const proper = desAcc + V3f(0, 0, 9.81f);
    const float thrust = proper.norm();
    V3f const dir = proper / thrust;
    Rotf desAtt = attitude_from_thrust_dir(dir);
    float y, p, r;
    estAtt.to_euler_ypr(y, p, r);
    Rotf noYaw = Rotf::from_euler_ypr(0, p, r);
    V3f desW = attCtrl.desired_angular_velocity(desAtt, noYaw);
    desW.z = yawRate;
    V3f const tq = angVelCtrl.desired_torques(desW, estAngVel);
    mixer.motor_forces(thrust * mass, tq, desForces);
    mixer.speeds_from_thrust(desForces, desSpeeds);
  }

  void ctrl_ext_rates() {  // :528-588
    V3f const estAngVel = kf.angVel;
    float thrust = radio.msg.f[0];
    V3f desW(radio.msg.f[1], radio.msg.f[2], radio.msg.f[3]);
    V3f const tq = angVelCtrl.desired_torques(desW, estAngVel);
    mixer.motor_forces(thrust * mass, tq, desForces);
    mixer.speeds_from_thrust(desForces, desSpeeds);
    if (radio.msg.flags & 0x01) {
      if (!propCal.running) {
        propCal.running = true;
        propCal.count = 0;
        for (int i = 0; i < 4; i++) propCal.accum[i] = 0;
      }
      for (int i = 0; i < 4; i++) propCal.accum[i] += mixer.uncorrected_force(desSpeeds[i]);
      propCal.count++;
    } else if (propCal.running) {
      propCal.running = false;
      if (propCal.count >= propCal.minCount) {
        float truePer = mass * 9.81f / 4.0f;
        for (int i = 0; i < 4; i++) {
          float f = (propCal.count * truePer) / propCal.accum[i];
          if (f > propCal.fmax) f = propCal.fmax;
          if (f < propCal.fmin) f = propCal.fmin;
          propCal.active[i] = f;
        }
        for (int i = 0; i < 4; i++) mixer.corr[i] = propCal.active[i];
        shouldWriteParams = true;
      }
    }
  }

  void telemetry(uint8_t p1[30], uint8_t p2[30]) {  // :621-679
    uint16_t d[14];
    memset(p1, 0, 30);
    memset(p2, 0, 30);
    V3f a = acc.lp.value(), g = gyro.lp.value();
    p1[0] = 0;
    p1[1] = uint8_t(telCounter % 256);
    memset(d, 0, sizeof(d));
    for (int i = 0; i < 3; i++) {
      d[i + 0] = encode_ones(map_to_ones(a.get(i), -30, 30));
      d[i + 3] = encode_ones(map_to_ones(g.get(i), -35, 35));
    }
    for (int i = 0; i < 4; i++) d[i + 6] = encode_ones(map_to_ones(desForces[i], 0, 10));
    for (int i = 0; i < 3; i++) d[i + 10] = encode_ones(map_to_ones(kf.pos.get(i), -30, 30));
    d[13] = encode_ones(map_to_ones(batt.vRaw, 0, 15));
    memcpy(p1 + 2, d, 28);
    p2[0] = 1;
    p2[1] = uint8_t(telCounter % 256);
    memset(d, 0, sizeof(d));
    V3f av = kf.att.vector_part();
    for (int i = 0; i < 3; i++) {
      d[i + 0] = encode_ones(map_to_ones(kf.vel.get(i), -30, 30));
      d[i + 3] = encode_ones(map_to_ones(av.get(i), -1, 1));
    }
    for (int i = 0; i < 6; i++) d[i + 6] = encode_ones(map_to_ones(debug[i], -100, 100));
    d[12] = uint8_t(firstPanic);
    d[13] = warnings;
    memcpy(p2 + 2, d, 28);
    telCounter++;
    warnings = 0;
  }
};

// ---------------------------------------------------------------------------------------------
// Motor (Components/Components/Simulation/Motor.{hpp,cpp})
// ---------------------------------------------------------------------------------------------
struct Motor {
  Timer timer;
  double minSpeed, maxSpeed, kF, kTau, tau, J, speed;
  V3d position, rotAxis, thrustAxis, thrust, torque, angMom;
  double power, speedCmd;
  Motor(const Clock* c, V3d pos, V3d axis, bool clockwise, double mn, double mx, double kf, double kt,
        double tc, double inertia)
      : timer(c), minSpeed(mn), maxSpeed(mx), kF(kf), kTau(kt), tau(tc), J(inertia), speed(0),
        position(pos), rotAxis(axis), thrustAxis(0, 0, 0), thrust(0, 0, 0), torque(0, 0, 0),
        angMom(0, 0, 0), power(0), speedCmd(0) {
    thrustAxis = clockwise ? rotAxis : -rotAxis;  // Motor.cpp:32-36
  }
  void run() {  // Motor.cpp:39-84
    const double dt = timer.seconds_d();
    if (dt < 1e-6) return;
    timer.reset();
    double old = speed;
    if (speedCmd < 0) speedCmd = 0;
    double c;
    if (tau == 0) {
      c = 0;
    } else {
      c = exp(-dt / tau);
    }
    speed = c * speed + (1 - c) * speedCmd;
    if (speed > maxSpeed) {
      speed = maxSpeed;
    } else if (speed < minSpeed) {
      speed = minSpeed;
    }
    angMom = speed * J * rotAxis;
    thrust = kF * speed * fabs(speed) * thrustAxis;
    torque = V3d(0, 0, 0);
    torque = torque + (-kTau * speed * fabs(speed) * rotAxis);
    torque = torque + position.cross(thrust);
    double angAcc = (speed - old) / dt;
    torque = torque - (angAcc * J * rotAxis);
    power = speed * torque.norm();
  }
};

// ---------------------------------------------------------------------------------------------
// UWB radio + network (Components/Components/Simulation/UWBRadio.hpp, UWBNetwork.cpp)
// ---------------------------------------------------------------------------------------------
struct Radio {
  uint8_t id, nextTarget;
  V3d truePos;
  struct { bool haveNew; float range; uint8_t responder; bool failure; } meas;
  explicit Radio(uint8_t i) : id(i), nextTarget(0) { meas.haveNew = false; meas.range = 0; meas.responder = 0; meas.failure = false; }
};

struct Network {
  std::vector<Radio*> radios;
  double commPeriod;
  Timer sinceLast;
  uint8_t requester, responder;
  double noiseStd;
  Network(const Clock* c, double period) : commPeriod(period), sinceLast(c), requester(0), responder(0), noiseStd(0) {}
  void run() {  // UWBNetwork.cpp:22-89 (noise term: noise-free only in the port; sigma = 0 adds exactly 0)
    if (sinceLast.seconds_d() < commPeriod) return;
    if (!requester || !responder) {
      for (Radio* r : radios) {
        if (r->nextTarget) {
          requester = r->id;
          responder = r->nextTarget;
          break;
        }
      }
      sinceLast.reset();
      return;
    }
    V3d reqPos, resPos;
    bool haveReq = false, haveRes = false;
    for (Radio* r : radios) {
      if (r->id == requester) { reqPos = r->truePos; haveReq = true; }
      if (r->id == responder) { resPos = r->truePos; haveRes = true; }
    }
    if (haveReq && haveRes) {
      double measNoise = 0.0 * noiseStd;
      float range = float((reqPos - resPos).norm() + measNoise);
      for (Radio* r : radios) {
        r->meas.range = range;
        r->meas.responder = responder;
        r->meas.failure = false;
        r->meas.haveNew = true;
      }
    }
    requester = 0;
    responder = 0;
  }
};

// ---------------------------------------------------------------------------------------------
// Quadcopter_T (Components/Components/Simulation/Quadcopter_T.{hpp,cpp})
// ---------------------------------------------------------------------------------------------
struct Quadcopter {
  Timer integ;
  V3d pos, vel, angVel;
  Rotd att;
  Radio radio;
  Logic logic;
  float cmd[4];
  M33<double> I, Iinv;
  double mass;
  V3d extForce, extTorque, drag;
  std::vector<Motor> motors;
  float battV, battI;
  double sAcc, sGyro;
  Timer logicTimer;
  double logicPeriod;
  M33<float> RimuInv;

  static M33<double> inverse3(const M33<double>& a) {  // oracle/shim/Eigen/Dense inverse() contract
    auto cof = [&](int i, int j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      return a.m[i1][j1] * a.m[i2][j2] - a.m[i1][j2] * a.m[i2][j1];
    };
    const double c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
    const double det = c0 * a.m[0][0] + (c1 * a.m[1][0] + c2 * a.m[2][0]);
    const double invdet = 1.0 / det;
    M33<double> r;
    r.m[0][0] = c0 * invdet; r.m[0][1] = c1 * invdet; r.m[0][2] = c2 * invdet;
    for (int row = 1; row < 3; row++)
      for (int col = 0; col < 3; col++) r.m[row][col] = cof(col, row) * invdet;
    return r;
  }

  Quadcopter(const Clock* c, const agf_vehicle_cfg& k, double period)  // Quadcopter_T.cpp:9-83
      : integ(c), pos(0, 0, 0), vel(0, 0, 0), angVel(0, 0, 0), att(Rotd::identity()),
        radio(uint8_t(k.vehicle_id)), logic(c, float(period)), mass(k.mass), extForce(0, 0, 0),
        extTorque(0, 0, 0), drag(k.lin_drag_coeff_b[0], k.lin_drag_coeff_b[1], k.lin_drag_coeff_b[2]),
        sAcc(0.2), sGyro(0.1), logicTimer(c), logicPeriod(period) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) I.m[i][j] = k.inertia[3 * i + j];
    Iinv = inverse3(I);
    V3d const com(k.com_error[0], k.com_error[1], k.com_error[2]);
    V3d const spin(0, 0, 1);
    const double a = k.arm_length / sqrt(2);
    const double sx[4] = {+1, -1, -1, +1}, sy[4] = {-1, -1, +1, +1};
    for (int i = 0; i < 4; i++) {
      bool cw = (i % 2 == 0);
      motors.push_back(Motor(c, a * V3d(sx[i], sy[i], 0) + com, cw ? spin : -spin, cw, k.motor_min_speed,
                             k.motor_max_speed, k.prop_thrust_from_speed_sqr, k.prop_torque_from_speed_sqr,
                             k.motor_time_const, k.motor_inertia));
      cmd[i] = 0;
    }
    battV = 1.2 * k.logic.low_battery_threshold;
    battI = -1.0;
    RimuInv = Rotf::from_euler_ypr(k.logic.imu_yaw, k.logic.imu_pitch, k.logic.imu_roll).inverse().matrix33();
    logic.initialise(k.logic, uint8_t(k.vehicle_id));
  }

  void run() {  // Quadcopter_T.cpp:86-203
    const double dt = integ.seconds_d();
    if (dt < 1e-6) return;
    integ.reset();
    V3d F(0, 0, 0), T(0, 0, 0);
    for (int i = 0; i < 4; i++) {
      motors[i].speedCmd = double(cmd[i]);
      motors[i].run();
      F = F + motors[i].thrust;
      T = T + motors[i].torque;
    }
    T = T + att.inverse().rotate(extTorque);
    V3d L = mul(I, angVel);
    for (int i = 0; i < 4; i++) L = L + motors[i].angMom;
    V3d angAcc = mul(Iinv, T - angVel.cross(L));
    V3d vel_b = att.inverse().rotate(vel);
    V3d dragF(drag.x * (-vel_b.x), drag.y * (-vel_b.y), drag.z * (-vel_b.z));
    F = F + dragF;
    V3d acc(0, 0, -9.81);
    acc = acc + (att.rotate(F) + extForce) / mass;
    V3d newpos = pos + vel * dt + 0.5 * acc * dt * dt;
    V3d newvel = vel + acc * dt;
    Rotd newatt = att * Rotd::from_rotation_vector(angVel * dt);
    V3d neww = angVel + angAcc * dt;
    if ((newpos.z <= 0) && (newvel.z < 0)) {
      newpos.z = 0;
      newvel.z = 0;
      acc.z = 0;
      neww = V3d(0, 0, 0);
    }
    pos = newpos;
    vel = newvel;
    att = newatt;
    angVel = neww;

    if (logicTimer.seconds_d() > logicPeriod) {
      logicTimer.adjust_by_seconds(-logicPeriod);
      logic.set_battery(battV, battI);
      V3f g = mul(RimuInv, V3f(angVel));
      g = g + V3f(0.0f, 0.0f, 0.0f) * float(sGyro);  // noise-free port: N(0,1) draws replaced by 0
      logic.set_gyro(g.x, g.y, g.z);
      V3f a = V3f(att.inverse().rotate(acc + V3d(0, 0, 9.81)));
      a = mul(RimuInv, a);
      a = a + V3f(0.0f, 0.0f, 0.0f) * float(sAcc);
      logic.set_acc(a.x, a.y, a.z);
      logic.set_temp(25);
      logic.run();
      for (int i = 0; i < 4; i++) cmd[i] = logic.desSpeeds[i];
      radio.truePos = pos;
      radio.nextTarget = logic.next_ranging_target();
      if (radio.meas.haveNew) {
        radio.meas.haveNew = false;
        logic.set_uwb(radio.meas.range, radio.meas.responder, radio.meas.failure);
      }
    }
  }
};

}  // namespace port

// =============================================================================================
// C interface (oracle/oracle_api.h)
// =============================================================================================
namespace port {
// RadioMessageDecoded::encodeToRadioByte (RadioTypes.hpp:73-100): 16-bit field as the two bytes it is sent in
inline uint16_t radio_encode_field(float valIn, float limit) {
  int out;
  if ((valIn > -limit) && (valIn < limit)) {
    out = int(valIn * 32768 / limit + 0.5f) + 32768;
  } else if (valIn > -limit) {
    out = 65536 - 1;
  } else {
    out = 0;  // min value, and NaN
  }
  return uint16_t(((out >> 8) % 256) << 8 | (out % 256));
}
// Offboard::QuadcopterController::Run (QuadcopterController.cpp:11-74) followed by CreateRatesCommand +
// RadioMessageDecoded(raw) (RadioTypes.hpp:158-171,210-219): returns the four floats the vehicle will decode
struct OffboardCmd {
  float f[4];
  double thrust;  // before quantisation: what SetPredictedValues is told (main.cpp:652-654)
  V3d angVel;
};
inline OffboardCmd quantise_rates(double thrust, const V3d& w) {
  const float tx[4] = {float(thrust), float(w.x), float(w.y), float(w.z)};
  OffboardCmd o;
  o.thrust = thrust;
  o.angVel = w;
  for (int i = 0; i < 4; i++) {
    const float limit = 35;  // MAX_VAL_CMD_THRUST == MAX_VAL_CMD_ANG_RATES == 35
    const int q = radio_encode_field(tx[i], limit);
    o.f[i] = limit * (q - 32768) / float(32768);
  }
  return o;
}
// shortest rotation taking e3 to `dir`, then the yaw rotation (QuadcopterController.cpp:46-71, 107-127)
inline Rotf att_from_thrust_dir_yawed(const V3f& dir, double yaw) {
  const V3f e3(0, 0, 1);
  Rotf att;
  const float cosAngle = dir.dot(e3);
  float angle;
  if (cosAngle >= (1 - 1e-12f)) {
    angle = 0;
  } else if (cosAngle <= -(1 - 1e-12f)) {
    angle = float(M_PI);
  } else {
    angle = m_acos(cosAngle);
  }
  V3f rotAx = e3.cross(dir);
  const float n = rotAx.norm();
  if (n < 1e-6f) {
    att = Rotf::identity();
  } else {
    att = Rotf::from_rotation_vector(rotAx * (angle / n));
  }
  return att * Rotf::from_rotation_vector(V3f(0, 0, float(yaw)));
}
// QuadcopterController::RunTracking (QuadcopterController.cpp:76-131) + CreateRatesCommand
inline OffboardCmd offboard_tracking_command(const agf_offboard_cfg& c, const V3d& curPos, const V3d& curVel, const Rotd& curAtt,
                                             const V3d& refPos, const V3d& refVel, const V3d& refAcc, double yaw, double refThrust,
                                             const V3d& refAngVel) {
  PositionController posCtrl;
  posCtrl.natFreq = c.pos_control_nat_freq;
  posCtrl.damping = c.pos_control_damping;
  AttitudeController attCtr;
  attCtr.tc_xy = c.att_control_time_const_xy;
  attCtr.tc_z = c.att_control_time_const_z;
  const Rotf attf(float(curAtt.v[0]), float(curAtt.v[1]), float(curAtt.v[2]), float(curAtt.v[3]));
  const V3f accErr = posCtrl.des_acceleration(V3f(curPos), V3f(curVel), V3f(refPos), V3f(refVel), V3f(0.0f, 0.0f, 0.0f));
  const double outCmdThrust = refThrust + accErr.dot(attf.rotate(V3f(0, 0, 1)));
  const V3d sum = refAcc + V3d(accErr) + V3d(V3f(0, 0, 9.81f));  // mixed Vec3d/Vec3f sum is carried in double
  const float normRefProperAcc = float(sum.norm());
  const V3f refThrustDir(sum / double(normRefProperAcc));
  const Rotf refAttYawed = att_from_thrust_dir_yawed(refThrustDir, yaw);
  const V3f angVelErr = attCtr.desired_angular_velocity(refAttYawed, attf);
  return quantise_rates(outCmdThrust, refAngVel + V3d(angVelErr));
}
inline OffboardCmd offboard_rates_command(const agf_offboard_cfg& c, const V3d& curPos, const V3d& curVel, const Rotd& curAtt,
                                          const V3d& desPos, const V3d& desVel = V3d(0, 0, 0), const V3d& desAcc = V3d(0, 0, 0),
                                          double yawOverride = std::numeric_limits<double>::quiet_NaN()) {
  const double yawAngle = (yawOverride == yawOverride) ? yawOverride : c.yaw_angle;
  PositionController posCtrl;
  posCtrl.natFreq = c.pos_control_nat_freq;
  posCtrl.damping = c.pos_control_damping;
  AttitudeController attCtr;
  attCtr.tc_xy = c.att_control_time_const_xy;
  attCtr.tc_z = c.att_control_time_const_z;
  const V3f e3(0, 0, 1);
  const V3f cmdAcc = posCtrl.des_acceleration(V3f(curPos), V3f(curVel), V3f(desPos), V3f(desVel), V3f(desAcc));
  V3f cmdProperAcc = cmdAcc + V3f(0, 0, 9.81f);
  if (cmdProperAcc.norm() > c.max_proper_acc) {  // float norm compared (and divided) in double, product back in float
    cmdProperAcc = cmdProperAcc * float(c.max_proper_acc / cmdProperAcc.norm());
  }
  if (cmdProperAcc.z < c.min_vertical_proper_acc) cmdProperAcc.z = float(c.min_vertical_proper_acc);
  const float normCmdProperAcc = cmdProperAcc.norm();
  const V3f cmdThrustDir = cmdProperAcc / normCmdProperAcc;
  const Rotf attf(float(curAtt.v[0]), float(curAtt.v[1]), float(curAtt.v[2]), float(curAtt.v[3]));
  double outCmdThrust = normCmdProperAcc * attf.rotate(V3f(0, 0, 1)).dot(cmdThrustDir);
  if (outCmdThrust < c.min_proper_acc) outCmdThrust = c.min_proper_acc;
  Rotf cmdAtt;
  const float cosAngle = cmdThrustDir.dot(e3);
  float angle;
  if (cosAngle >= (1 - 1e-12f)) {
    angle = 0;
  } else if (cosAngle <= -(1 - 1e-12f)) {
    angle = float(M_PI);
  } else {
    angle = m_acos(cosAngle);
  }
  V3f rotAx = e3.cross(cmdThrustDir);
  const float n = rotAx.norm();
  if (n < 1e-6f) {
    cmdAtt = Rotf::identity();
  } else {
    cmdAtt = Rotf::from_rotation_vector(rotAx * (angle / n));
  }
  Rotf cmdAttYawed = cmdAtt * Rotf::from_rotation_vector(V3f(0, 0, float(yawAngle)));
  const V3d outCmdAngVel(attCtr.desired_angular_velocity(cmdAttYawed, attf));
  return quantise_rates(outCmdThrust, outCmdAngVel);
}
}  // namespace port

namespace port {
// ---------------------------------------------------------------------------------------------
// Offboard::MocapStateEstimator (Components/Offboard/MocapStateEstimator.cpp) and its PredictionPipe
// (Components/Offboard/PredictionPipe.hpp).  2x2 products follow the oracle's Eigen contract: coefficient-wise,
// sequential in k from the k = 0 product, left to right.
// ---------------------------------------------------------------------------------------------
struct M22 {
  double d[2][2];
};
inline M22 mm2(const M22& a, const M22& b) {
  M22 c;
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++) {
      double acc = a.d[i][0] * b.d[0][j];
      acc += a.d[i][1] * b.d[1][j];
      c.d[i][j] = acc;
    }
  return c;
}
inline M22 add2(const M22& a, const M22& b) {
  M22 c;
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++) c.d[i][j] = a.d[i][j] + b.d[i][j];
  return c;
}
inline M22 tr2(const M22& a) {
  M22 c;
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++) c.d[j][i] = a.d[i][j];
  return c;
}
struct MocapEstimator {
  static constexpr double SMALL_TIME = 1e-6;
  static constexpr unsigned MAX_NUM_CONSECUTIVE_REJECTION = 10;
  struct Msg {
    double timeActive;
    V3d acc, angVel;
    bool ballistic;
  };
  Timer timer, lastGood, pipeTimer;
  uint64_t estTime_us;  // ManualTimer _estimateTimer
  bool initialized;
  V3d pos, vel, angVel;
  Rotd att;
  M22 varPos, varAtt;
  unsigned numRej, numRejCons;
  double tcAngVel, rejectDist, nMeasPos, nMeasAtt, nProcPos, nProcAtt, pipeDelay;
  std::deque<Msg> pipe;

  MocapEstimator(const Clock* c, const agf_offboard_estimator& e)
      : timer(c), lastGood(c), pipeTimer(c), estTime_us(0), initialized(false), numRej(0), numRejCons(0) {
    rejectDist = e.meas_reject_dist;
    tcAngVel = e.angvel_time_const;
    nMeasPos = e.meas_noise_pos;
    nMeasAtt = e.meas_noise_att;
    nProcPos = e.proc_noise_pos;
    nProcAtt = e.proc_noise_att;
    pipeDelay = e.prediction_delay;
    reset();
  }
  double est_seconds() const { return double(estTime_us * double(1e-6)); }
  void reset_variance() {  // :52-60
    varPos.d[0][0] = 25.0; varPos.d[1][1] = 25.0; varPos.d[0][1] = varPos.d[1][0] = 0.0;
    varAtt.d[0][0] = 1.0; varAtt.d[1][1] = 400; varAtt.d[0][1] = varAtt.d[1][0] = 0.0;
  }
  void reset() {  // :37-50
    initialized = false;
    pos = V3d(0, 0, 0);
    vel = V3d(0, 0, 0);
    att = Rotd::identity();
    angVel = V3d(0, 0, 0);
    reset_variance();
    estTime_us = timer.micros();
    lastGood.reset();
  }
  void set_predicted(const V3d& w, const V3d& a, bool ballistic = false) {  // hpp:74-80, PredictionPipe.hpp:25-30
    Msg m;
    m.timeActive = pipeTimer.seconds_d() + pipeDelay;
    m.acc = a;
    m.angVel = w;
    m.ballistic = ballistic;
    pipe.push_back(m);
  }
  bool active_message(double t, Msg& out, double& timeRemaining) const {  // PredictionPipe.hpp:32-53
    if (pipe.empty()) return false;
    double tLastMsg = 1e10;
    for (size_t k = pipe.size(); k-- > 0;) {
      if ((t + SMALL_TIME) >= pipe[k].timeActive) {
        out = pipe[k];
        timeRemaining = tLastMsg - pipe[k].timeActive;
        return true;
      }
      tLastMsg = pipe[k].timeActive;
    }
    return false;
  }
  void clear_expired(double currentTime) {  // PredictionPipe.hpp:55-68
    const int N = int(pipe.size());
    for (int i = 0; i < N; i++) {
      if (pipe.size() < 2) return;
      if (pipe[1].timeActive <= currentTime) pipe.pop_front();
    }
  }
  static void fetch(const MocapEstimator& e, double t, Msg& cmd, double& predictionTime) {
    predictionTime = 0;
    if (!e.active_message(t, cmd, predictionTime)) {
      cmd.acc = V3d(0, 0, 0);
      cmd.angVel = V3d(0, 0, 0);
      cmd.ballistic = true;
      predictionTime = 1e10;
    }
  }
  void prediction(double dt, V3d& oPos, V3d& oVel, Rotd& oAtt, V3d& oAngVel) const {  // GetPrediction :61-118
    const double tEnd = dt + timer.seconds_d();
    const double tStart = est_seconds();
    oPos = pos; oVel = vel; oAtt = att; oAngVel = angVel;
    double t = tStart;
    while ((t + SMALL_TIME) < tEnd) {
      Msg cmd;
      double predictionTime;
      fetch(*this, t, cmd, predictionTime);
      double dtInt = tEnd - t;
      if (dtInt > (predictionTime + SMALL_TIME)) dtInt = predictionTime;
      const V3d newPos = oPos + vel * dtInt + cmd.acc * dtInt * dtInt / 2.0;  // sic: _vel, not est.vel (:90)
      const V3d newVel = oVel + cmd.acc * dtInt;
      const Rotd newAtt = oAtt * Rotd::from_rotation_vector(angVel * dtInt);  // sic: _angVel (:92)
      double discrete = m_exp(-dtInt / tcAngVel);
      if (cmd.ballistic) discrete = 1;
      const V3d newAngVel = discrete * oAngVel + (1 - discrete) * cmd.angVel;
      oPos = newPos; oVel = newVel; oAtt = newAtt; oAngVel = newAngVel;
      t += dtInt;
    }
  }
  void update(const V3d& measPos, const Rotd& measAtt) {  // UpdateWithMeasurement :120-265
    if (!initialized) {
      initialized = true;
      pos = measPos;
      vel = V3d(0, 0, 0);
      att = measAtt;
      angVel = V3d(0, 0, 0);
      lastGood.reset();
      reset_variance();
      return;
    }
    const double t0 = est_seconds();
    const double tEnd = timer.seconds_d();
    if (tEnd > t0) {
      for (;;) {
        const double tNow = est_seconds();
        if ((tNow + SMALL_TIME) >= tEnd) break;
        Msg p;
        double predictionTime;
        fetch(*this, est_seconds(), p, predictionTime);
        double dtInt = tEnd - est_seconds();
        if (dtInt > (predictionTime + SMALL_TIME)) dtInt = predictionTime;
        const V3d pos0(pos), vel0(vel), angVel0(angVel);
        const Rotd att0(att);
        pos = pos0 + vel0 * dtInt;
        vel = vel0 + p.acc * dtInt;
        att = att0 * Rotd::from_rotation_vector(angVel0 * dtInt);
        double discrete = m_exp(-dtInt / tcAngVel);
        if (p.ballistic) discrete = 1;
        angVel = discrete * angVel0 + (1 - discrete) * p.angVel;
        estTime_us += uint64_t(0.5 + dtInt * 1e6);
        M22 A, Qp, Qa;
        A.d[0][0] = 1; A.d[0][1] = dtInt; A.d[1][0] = 0; A.d[1][1] = 1;
        Qp.d[0][0] = dtInt * dtInt * dtInt * dtInt * nProcPos / 4; Qp.d[0][1] = 0; Qp.d[1][0] = 0; Qp.d[1][1] = dtInt * dtInt * nProcPos;
        Qa.d[0][0] = dtInt * dtInt * dtInt * dtInt * nProcAtt / 4; Qa.d[0][1] = 0; Qa.d[1][0] = 0; Qa.d[1][1] = dtInt * dtInt * nProcAtt;
        const M22 newPosVar = add2(mm2(mm2(A, varPos), tr2(A)), Qp);
        const M22 newAttVar = add2(mm2(mm2(A, varAtt), tr2(A)), Qa);
        varPos = newPosVar;
        varAtt = newAttVar;
      }
    }
    double innovationCovPos = varPos.d[0][0] + nMeasPos * nMeasPos;
    double innovationCovAtt = varAtt.d[0][0] + nMeasAtt * nMeasAtt;
    const double distMeasPos = (measPos - pos).norm() / m_sqrt(3 * innovationCovPos);
    const Rotd dq = measAtt.inverse() * att;
    const double distMeasAtt = (m_acos(fabs(dq.v[0])) * 2.0) / m_sqrt(innovationCovAtt);  // Rotation::GetAngle :138-142
    bool shouldRejectMeas = false;
    if ((distMeasPos > rejectDist) || (distMeasAtt > rejectDist)) shouldRejectMeas = true;
    if (shouldRejectMeas && numRejCons < MAX_NUM_CONSECUTIVE_REJECTION) {
      numRej++;
      numRejCons++;
    } else {
      if (numRejCons >= MAX_NUM_CONSECUTIVE_REJECTION) {
        reset();
        innovationCovPos = varPos.d[0][0] + nMeasPos * nMeasPos;
        innovationCovAtt = varAtt.d[0][0] + nMeasAtt * nMeasAtt;
      }
      numRejCons = 0;
      lastGood.reset();
      // gain = V * H^T * (1 / S), H = [1 0]
      const double sP = 1 / innovationCovPos, sA = 1 / innovationCovAtt;
      double gP[2], gA[2];
      for (int i = 0; i < 2; i++) {
        double a = varPos.d[i][0] * 1.0;
        a += varPos.d[i][1] * 0.0;
        gP[i] = a * sP;
        double b = varAtt.d[i][0] * 1.0;
        b += varAtt.d[i][1] * 0.0;
        gA[i] = b * sA;
      }
      const V3d measErrPos = measPos - pos;
      pos = pos + gP[0] * measErrPos;
      vel = vel + gP[1] * measErrPos;
      const V3d measErrAtt = (att.inverse() * measAtt).to_rotation_vector();
      att = att * Rotd::from_rotation_vector(gA[0] * measErrAtt);
      angVel = angVel + gA[1] * measErrAtt;
      M22 Ip, Ia;
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) {
          const double h = (j == 0) ? 1.0 : 0.0, id = (i == j) ? 1.0 : 0.0;
          Ip.d[i][j] = id - gP[i] * h;
          Ia.d[i][j] = id - gA[i] * h;
        }
      const M22 newPosVar = mm2(Ip, varPos), newAttVar = mm2(Ia, varAtt);
      varPos = newPosVar;
      varAtt = newAttVar;
    }
    M22 sp, sa;
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++) {
        sp.d[i][j] = (varPos.d[i][j] + varPos.d[j][i]) * 0.5;
        sa.d[i][j] = (varAtt.d[i][j] + varAtt.d[j][i]) * 0.5;
      }
    varPos = sp;
    varAtt = sa;
    clear_expired(est_seconds());
  }
};
}  // namespace port

struct orc_vehicle {
  port::Clock clock;
  port::Quadcopter* quad;
  port::Network* net;
  std::vector<port::Radio*> anchors;
  uint64_t tick;
  // offboard loop (orc_run_offboard)
  port::Timer* offTimer = nullptr;
  struct Queued { uint64_t due; port::OffboardCmd cmd; int type; };
  std::deque<Queued> offQueue;
  // offboard state estimator (orc_set_offboard_estimator)
  port::MocapEstimator* est = nullptr;
  port::Timer* timerMocap = nullptr;
  double periodMocap = 0, delayEst = 0;
  void mocap_step() {  // main.cpp:451-457
    if (!est) return;
    if (timerMocap->seconds_d() > periodMocap) {
      timerMocap->adjust_by_seconds(-periodMocap);
      est->update(quad->pos, quad->att);
    }
  }
  void estimate(port::V3d& p, port::V3d& vl, port::Rotd& a) {
    port::V3d w;
    if (est) {
      est->prediction(delayEst, p, vl, a, w);
    } else {
      p = quad->pos; vl = quad->vel; a = quad->att;
    }
  }
  void set_predicted(const port::V3d& w, double thrust, const port::Rotd& a) {  // main.cpp:652-654
    if (est) est->set_predicted(w, a.rotate(port::V3d(0, 0, 1)) * thrust - port::V3d(0, 0, 9.81));
  }
  // reference generators (orc_run_offboard_ref): ExampleVehicleStateMachine members
  int stage = AGF_STAGE_WAIT_FOR_START, lastStage = AGF_STAGE_COMPLETE;
  uint64_t stageStart = 0;
  port::V3d initPosition = port::V3d(0, 0, 0), lastPos = port::V3d(0, 0, 0), lastVel = port::V3d(0, 0, 0), lastAcc = port::V3d(0, 0, 0);  // ExampleVehicleStateMachine.cpp:16-25
  double cmdYawAngle = 0;
};

static void record(orc_vehicle* v, double* r) {
  port::Quadcopter& q = *v->quad;
  r[0] = q.pos.x; r[1] = q.pos.y; r[2] = q.pos.z;
  r[3] = q.vel.x; r[4] = q.vel.y; r[5] = q.vel.z;
  for (int i = 0; i < 4; i++) r[6 + i] = q.att.v[i];
  r[10] = q.angVel.x; r[11] = q.angVel.y; r[12] = q.angVel.z;
  for (int i = 0; i < 4; i++) r[13 + i] = q.motors[i].speed;
  for (int i = 0; i < 4; i++) r[17 + i] = q.cmd[i];
  const port::KalmanFilter& kf = q.logic.kf;
  r[21] = kf.pos.x; r[22] = kf.pos.y; r[23] = kf.pos.z;
  r[24] = kf.vel.x; r[25] = kf.vel.y; r[26] = kf.vel.z;
  for (int i = 0; i < 4; i++) r[27 + i] = kf.att.v[i];
  r[31] = kf.angVel.x; r[32] = kf.angVel.y; r[33] = kf.angVel.z;
  r[34] = q.logic.state;
  r[35] = q.logic.firstPanic;
  r[36] = q.logic.cycle;
  r[37] = kf.numResets;
  r[38] = kf.numRejected;
  r[39] = q.logic.uwb.count;
}

extern "C" {

const char* orc_flavour(void) { return ORC_FLAVOUR; }

orc_vehicle* orc_create(const agf_vehicle_cfg* cfg, const orc_opts* opts) {
  orc_vehicle* v = new orc_vehicle();
  v->clock.now_us = 0;
  v->tick = 0;
  v->quad = new port::Quadcopter(&v->clock, *cfg, opts->onboard_logic_period);
  v->quad->sAcc = opts->sigma_acc;
  v->quad->sGyro = opts->sigma_gyro;
  v->net = nullptr;
  if (opts->uwb_comm_period > 0) {
    v->net = new port::Network(&v->clock, opts->uwb_comm_period);
    v->net->noiseStd = opts->uwb_noise_std_dev;
    v->net->radios.push_back(&v->quad->radio);
  }
  return v;
}

void orc_destroy(orc_vehicle* v) {
  for (auto* a : v->anchors) delete a;
  delete v->net;
  delete v->quad;
  delete v->offTimer;
  delete v;
}

void orc_set_state(orc_vehicle* v, const double p[3], const double vel[3], const double a[4],
                   const double w[3]) {
  v->quad->pos = port::V3d(p[0], p[1], p[2]);
  v->quad->vel = port::V3d(vel[0], vel[1], vel[2]);
  v->quad->att = port::Rotd(a[0], a[1], a[2], a[3]);
  v->quad->angVel = port::V3d(w[0], w[1], w[2]);
}

void orc_set_external(orc_vehicle* v, const double f[3], const double t[3]) {
  if (f) v->quad->extForce = port::V3d(f[0], f[1], f[2]);
  if (t) v->quad->extTorque = port::V3d(t[0], t[1], t[2]);
}

int orc_add_anchor(orc_vehicle* v, uint8_t id, float x, float y, float z) {
  if (v->quad->logic.add_target(id, port::V3f(x, y, z))) return -1;
  if (v->net) {
    port::Radio* r = new port::Radio(id);
    r->truePos = port::V3d(port::V3f(x, y, z));
    v->anchors.push_back(r);
    v->net->radios.push_back(r);
  }
  return 0;
}

void orc_set_radio(orc_vehicle* v, const uint8_t raw[23]) {
  v->quad->logic.set_radio(port::radio_decode(raw));
}

void orc_run(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_cmd_entry* sched,
             uint32_t nsched, const uint8_t* slot_raw, double* traj) {
  uint32_t si = 0;
  while (si < nsched && sched[si].tick < v->tick) si++;
  for (uint32_t k = 0; k < nticks; k++) {
    if (si < nsched && sched[si].tick == v->tick) {
      const uint8_t* raw = sched[si].raw;
      if (sched[si].slot >= 0 && slot_raw) raw = slot_raw + AGF_RADIO_PACKET_SIZE * sched[si].slot;
      orc_set_radio(v, raw);
      si++;
    }
    v->quad->run();
    if (v->net) v->net->run();
    if (traj) record(v, traj + size_t(k) * ORC_NTRAJ);
    v->clock.now_us += dt_us;
    v->tick++;
  }
}

void orc_run_offboard(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_offboard_cfg* cfg,
                      const agf_offboard_target* targets, uint32_t n_targets, const double* offset, double* traj) {
  if (!v->offTimer) v->offTimer = new port::Timer(&v->clock);
  const double period = double(cfg->period_us) * 1e-6;
  for (uint32_t k = 0; k < nticks; k++) {
    if (!v->offQueue.empty() && v->clock.now_us >= v->offQueue.front().due) {  // CommunicationsDelay.hpp:27-35, main.cpp:737-739
      port::RadioMsg m;
      m.type = uint8_t(v->offQueue.front().type);
      m.flags = uint8_t(cfg->radio_flags);
      for (int i = 0; i < 4; i++) m.f[i] = v->offQueue.front().cmd.f[i];
      for (int i = 4; i < 10; i++) m.f[i] = 35.0f * (0 - 32768) / float(32768);  // zero-filled packet bytes decode to -limit; never read
      v->quad->logic.set_radio(m);
      v->offQueue.pop_front();
    }
    v->quad->run();
    if (v->net) v->net->run();
    if (traj) record(v, traj + size_t(k) * ORC_NTRAJ);
    v->clock.now_us += dt_us;
    v->tick++;
    v->mocap_step();
    if (v->offTimer->seconds_d() > period) {  // main.cpp:471
      v->offTimer->adjust_by_seconds(-period);  // main.cpp:476
      int ti = -1;
      for (uint32_t j = 0; j < n_targets; j++)
        if (targets[j].time_us <= v->clock.now_us) ti = int(j);
      if (ti < 0) continue;
      port::V3d des(targets[ti].pos[0], targets[ti].pos[1], targets[ti].pos[2]);
      if (offset) des = des + port::V3d(offset[0], offset[1], offset[2]);
      orc_vehicle::Queued q;
      q.due = v->clock.now_us + cfg->delay_us;
      q.type = AGF_RADIO_EXTERNAL_RATES_CMD;
      port::V3d ePos, eVel;
      port::Rotd eAtt;
      v->estimate(ePos, eVel, eAtt);
      q.cmd = port::offboard_rates_command(*cfg, ePos, eVel, eAtt, des);
      v->set_predicted(q.cmd.angVel, q.cmd.thrust, eAtt);
      v->offQueue.push_back(q);
    }
  }
}

namespace port {
// SingleAxisTrajectory.hpp:57-63; q: p0 v0 a0 alpha beta gamma
inline double sat_pos(const double* q, double t) {
  return q[0] + q[1] * t + (1 / 2.0) * q[2] * t * t + (1 / 6.0) * q[5] * t * t * t + (1 / 24.0) * q[4] * t * t * t * t +
         (1 / 120.0) * q[3] * t * t * t * t * t;
}
inline double sat_vel(const double* q, double t) {
  return q[1] + q[2] * t + (1 / 2.0) * q[5] * t * t + (1 / 6.0) * q[4] * t * t * t + (1 / 24.0) * q[3] * t * t * t * t;
}
inline double sat_acc(const double* q, double t) { return q[2] + q[5] * t + (1 / 2.0) * q[4] * t * t + (1 / 6.0) * q[3] * t * t * t; }
inline V3d prim_at(const double* tr, double t, double (*f)(const double*, double)) { return V3d(f(tr, t), f(tr + 6, t), f(tr + 12, t)); }
inline V3d prim_thrust_vec(const double* tr, double t) { return prim_at(tr, t, sat_acc) - V3d(tr[18], tr[19], tr[20]); }
// RapidTrajectoryGenerator::GetOmega (RapidTrajectoryGenerator.cpp:264-286)
inline V3d prim_omega(const double* tr, double t, double timeStep) {
  const V3d n0 = prim_thrust_vec(tr, t).unit();
  const V3d n1 = prim_thrust_vec(tr, t + timeStep).unit();
  const V3d crossProd = n0.cross(n1);
  if (crossProd.norm() <= 1e-6) return V3d(0, 0, 0);
  const V3d n = crossProd.unit();
  const double d = n0.dot(n1);
  if (d > 1.0 || d < -1.0 || d != d) return V3d(0, 0, 0);  // errno after acos
  const double angle = m_acos(d) / timeStep;
  return angle * n;
}
}  // namespace port

void orc_run_offboard_ref(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_offboard_cfg* cfg,
                          const agf_offboard_ref* ref, const double* offset, const double* tr, double* traj) {
  using port::V3d;
  if (!v->offTimer) v->offTimer = new port::Timer(&v->clock);
  const double period = double(cfg->period_us) * 1e-6;
  V3d desired(ref->desired_pos[0], ref->desired_pos[1], ref->desired_pos[2]);
  if (offset) desired = desired + V3d(offset[0], offset[1], offset[2]);
  const V3d zero(0, 0, 0);
  for (uint32_t k = 0; k < nticks; k++) {
    if (!v->offQueue.empty() && v->clock.now_us >= v->offQueue.front().due) {
      port::RadioMsg m;
      m.type = uint8_t(v->offQueue.front().type);
      m.flags = uint8_t(cfg->radio_flags);
      for (int i = 0; i < 4; i++) m.f[i] = v->offQueue.front().cmd.f[i];
      for (int i = 4; i < 10; i++) m.f[i] = 35.0f * (0 - 32768) / float(32768);
      v->quad->logic.set_radio(m);
      v->offQueue.pop_front();
    }
    v->quad->run();
    if (v->net) v->net->run();
    if (traj) record(v, traj + size_t(k) * ORC_NTRAJ);
    v->clock.now_us += dt_us;
    v->tick++;
    v->mocap_step();
    if (!(v->offTimer->seconds_d() > period)) continue;
    v->offTimer->adjust_by_seconds(-period);
    const uint64_t now = v->clock.now_us;
    V3d estPos, estVel;
    port::Rotd estAtt;
    v->estimate(estPos, estVel, estAtt);
    int predicted = 2;  // 0: nothing, 1: SetPredictedValues(0, 0), 2: from the command
    orc_vehicle::Queued q;
    q.due = now + cfg->delay_us;
    q.type = AGF_RADIO_EXTERNAL_RATES_CMD;
    bool send = true;
    if (ref->kind == AGF_OFFREF_TRAJECTORY) {
      if (!(now > ref->start_us)) {
        q.cmd = port::offboard_rates_command(*cfg, estPos, estVel, estAtt, desired, zero, zero, ref->desired_yaw);
      } else {
        double traj_t = double(now - ref->start_us) * 1e-6;
        const double tEnd = tr[21];
        V3d tp, tv, ta;
        if (traj_t < tEnd) {
          traj_t += 0.04;
          tp = port::prim_at(tr, traj_t, port::sat_pos);
          tv = port::prim_at(tr, traj_t, port::sat_vel);
          ta = port::prim_at(tr, traj_t, port::sat_acc);
        } else {
          tp = port::prim_at(tr, tEnd, port::sat_pos);
          tv = zero;
          ta = zero;
        }
        if (tp.z < 0) {
          tp.z = 0;
          if (tv.z < 0) tv.z = 0;
          if (ta.z < 0) ta.z = 0;
        }
        const port::Rotd trajAtt(tr[22], tr[23], tr[24], tr[25]);
        const V3d refPos = trajAtt.rotate(tp) + V3d(tr[26], tr[27], tr[28]);
        const V3d refVel = trajAtt.rotate(tv), refAcc = trajAtt.rotate(ta);
        const double refThrust = port::prim_thrust_vec(tr, traj_t).norm();
        const V3d refAngVel = (estAtt.inverse() * trajAtt).rotate(port::prim_omega(tr, traj_t, 0.02));
        q.cmd = port::offboard_tracking_command(*cfg, estPos, estVel, estAtt, refPos, refVel, refAcc, ref->desired_yaw, refThrust,
                                                refAngVel);
      }
    } else {
      const bool shouldStart = now >= ref->start_us, shouldStop = now >= ref->stop_us;
      bool safe = true;  // SafetyNet::UpdateWithEstimator + IsSafe (Components/Offboard/SafetyNet.hpp:70-106)
      if (ref->safety_net) {
        const double since = v->est ? v->est->lastGood.seconds_d() : 0.0;
        bool notSeen = since > ref->not_seen_timeout, unsafePos = false, upsideDownAndLow = false;
        for (int a = 0; a < 3; a++) {
          if (estPos.get(a) < ref->safe_min[a]) unsafePos = true;
          if (estPos.get(a) > ref->safe_max[a]) unsafePos = true;
        }
        if (estPos.z < ref->min_normal_height) {
          if (estAtt.rotate(V3d(0, 0, 1)).z < 0) upsideDownAndLow = true;
        }
        safe = !(notSeen || unsafePos || upsideDownAndLow);
      }
      const bool stageChange = v->stage != v->lastStage;
      v->lastStage = v->stage;
      if (stageChange) v->stageStart = now;
      const double ts = double(now - v->stageStart) * 1e-6;
      switch (v->stage) {
        case AGF_STAGE_WAIT_FOR_START:
          if (shouldStart) v->stage = AGF_STAGE_SPOOL_UP;
          send = false;
          predicted = 0;
          break;
        case AGF_STAGE_SPOOL_UP:
          if (!safe) v->stage = AGF_STAGE_EMERGENCY;
          predicted = 1;
          q.cmd = port::quantise_rates(9.81 * 0.25, zero);
          if (ts > 0.5) v->stage = AGF_STAGE_TAKEOFF;
          break;
        case AGF_STAGE_TAKEOFF: {
          if (stageChange) v->initPosition = estPos;
          if (!safe) v->stage = AGF_STAGE_EMERGENCY;
          double frac = ts / 2.0;
          if (frac >= 1.0) {
            v->stage = AGF_STAGE_FLIGHT;
            frac = 1.0;
          }
          const V3d cmdPos = (1 - frac) * v->initPosition + frac * desired;
          q.cmd = port::offboard_rates_command(*cfg, estPos, estVel, estAtt, cmdPos, zero, zero, v->cmdYawAngle);
        } break;
        case AGF_STAGE_FLIGHT: {
          if (!safe) v->stage = AGF_STAGE_EMERGENCY;
          V3d cmdPos(0, 0, 0), cmdVel(0, 0, 0), cmdAcc(0, 0, 0);
          const double t = ts;
          const double frac = std::min(t / 2.0, 1.0);
          const double yaw0 = ref->desired_yaw;
          switch (ref->traj_id) {
            case 0:
              cmdPos = desired;
              v->cmdYawAngle = 0;
              break;
            case 1: {
              const double radius = 1.0, angSpeed = 0.5;
              cmdPos = V3d(0.0, -2.0, desired.z) + radius * V3d(port::m_cos(angSpeed * t), port::m_sin(angSpeed * t), 0);
              cmdVel = (radius * angSpeed) * V3d(-port::m_sin(angSpeed * t), port::m_cos(angSpeed * t), 0);
              cmdAcc = (radius * (angSpeed * angSpeed)) * V3d(-port::m_cos(angSpeed * t), -port::m_sin(angSpeed * t), 0);
              v->cmdYawAngle = yaw0 + angSpeed * t;
            } break;
            case 2: {
              const double amplitude = 1.0, angFreq = 2.0;
              cmdPos = desired + amplitude * V3d(0, port::m_sin(angFreq * t), 0);
              cmdVel = (amplitude * angFreq) * V3d(0, port::m_cos(angFreq * t), 0);
              cmdAcc = (amplitude * (angFreq * angFreq)) * V3d(0, -port::m_sin(angFreq * t), 0);
              v->cmdYawAngle = yaw0;
            } break;
            case 3: {
              const double radius = 0.5, angSpeed = 1;
              cmdPos = V3d(0.0, 0.0, desired.z) + radius * V3d(port::m_cos(angSpeed * t), port::m_sin(angSpeed * t), 0);
              cmdVel = (radius * angSpeed) * V3d(-port::m_sin(angSpeed * t), port::m_cos(angSpeed * t), 0);
              cmdAcc = (radius * (angSpeed * angSpeed)) * V3d(-port::m_cos(angSpeed * t), -port::m_sin(angSpeed * t), 0);
              v->cmdYawAngle = 0;
            } break;
            case 4: {
              const double radius = 0.5, angSpeed = 0.5;
              cmdPos = V3d(0.0, 0.0, desired.z) +
                       radius * V3d(port::m_cos(angSpeed * t), port::m_sin(angSpeed * t), port::m_cos(angSpeed * t * 4));
              cmdVel = (radius * angSpeed) * V3d(-port::m_sin(angSpeed * t), port::m_cos(angSpeed * t), -port::m_sin(angSpeed * t * 4));
              cmdAcc = (radius * (angSpeed * angSpeed)) *
                       V3d(-port::m_cos(angSpeed * t), -port::m_sin(angSpeed * t), -port::m_cos(angSpeed * t * 4));
              v->cmdYawAngle = angSpeed * t;
            } break;
            default:
              cmdPos = desired;
              v->cmdYawAngle = 0.2 * t;
              break;
          }
          v->lastPos = (1 - frac) * desired + frac * cmdPos;
          v->lastVel = frac * cmdVel;
          v->lastAcc = frac * cmdAcc;
          q.cmd = port::offboard_rates_command(*cfg, estPos, estVel, estAtt, v->lastPos, v->lastVel, v->lastAcc, v->cmdYawAngle);
          if (shouldStop) v->stage = AGF_STAGE_LANDING;
        } break;
        case AGF_STAGE_LANDING: {
          if (!safe) v->stage = AGF_STAGE_EMERGENCY;
          const double frac = std::min(ts / 2.0, 1.0);
          const V3d land(0, 0, -0.5);
          const V3d cmdPos = v->lastPos + ts * land;
          if (cmdPos.z < 0) v->stage = AGF_STAGE_COMPLETE;
          q.cmd = port::offboard_rates_command(*cfg, estPos, estVel, estAtt, (1 - frac) * v->lastPos + frac * cmdPos,
                                               (1 - frac) * v->lastVel + frac * land, (1 - frac) * v->lastAcc + frac * zero,
                                               v->cmdYawAngle);
        } break;
        case AGF_STAGE_COMPLETE:
          predicted = 1;
          q.type = AGF_RADIO_IDLE_CMD;
          for (int i = 0; i < 4; i++) q.cmd.f[i] = 35.0f * (0 - 32768) / float(32768);  // zero bytes decode to -limit; never read
          break;
        default:  // AGF_STAGE_EMERGENCY
          predicted = 0;
          q.type = AGF_RADIO_EMERGENCY_KILL;
          for (int i = 0; i < 4; i++) q.cmd.f[i] = 35.0f * (0 - 32768) / float(32768);
          break;
      }
    }
    if (v->est && predicted == 1) v->est->set_predicted(zero, zero);
    if (predicted == 2) v->set_predicted(q.cmd.angVel, q.cmd.thrust, estAtt);
    if (send) v->offQueue.push_back(q);
  }
}

void orc_set_offboard_estimator(orc_vehicle* v, const agf_offboard_estimator* e) {
  delete v->est;
  delete v->timerMocap;
  v->est = nullptr;
  v->timerMocap = nullptr;
  if (!e || e->kind != AGF_OFFEST_MOCAP) return;
  v->est = new port::MocapEstimator(&v->clock, *e);
  v->timerMocap = new port::Timer(&v->clock);
  v->periodMocap = double(e->mocap_period_us) * 1e-6;
  v->delayEst = e->prediction_delay;
}

void orc_get_offboard_estimate(orc_vehicle* v, double horizon, double* o13, double* c4) {
  port::V3d p, vl, w;
  port::Rotd a;
  if (v->est) {
    v->est->prediction(horizon, p, vl, a, w);
  } else {
    p = v->quad->pos; vl = v->quad->vel; a = v->quad->att; w = v->quad->angVel;
  }
  o13[0] = p.x; o13[1] = p.y; o13[2] = p.z;
  o13[3] = vl.x; o13[4] = vl.y; o13[5] = vl.z;
  for (int i = 0; i < 4; i++) o13[6 + i] = a.v[i];
  o13[10] = w.x; o13[11] = w.y; o13[12] = w.z;
  if (c4) {
    c4[0] = v->est ? double(v->est->initialized) : 0.0;
    c4[1] = v->est ? double(v->est->numRej) : 0.0;
    c4[2] = v->est ? double(v->est->numRejCons) : 0.0;
    c4[3] = v->est ? double(v->est->pipe.size()) : 0.0;
  }
}

void orc_get_offboard_state(orc_vehicle* v, double* o) {
  o[0] = v->stage;
  o[1] = v->lastStage;
  o[2] = double(v->stageStart);
  const port::V3d* q[4] = {&v->initPosition, &v->lastPos, &v->lastVel, &v->lastAcc};
  for (int i = 0; i < 4; i++) {
    o[3 + 3 * i] = q[i]->x; o[4 + 3 * i] = q[i]->y; o[5 + 3 * i] = q[i]->z;
  }
  o[15] = v->cmdYawAngle;
}

static void dump3(const port::LPF2<port::V3f>& f, float out[4][3]) {
  const port::V3f* s[4] = {&f.xm0, &f.xm1, &f.ym0, &f.ym1};
  for (int i = 0; i < 4; i++) { out[i][0] = s[i]->x; out[i][1] = s[i]->y; out[i][2] = s[i]->z; }
}

void orc_get_full(orc_vehicle* v, orc_full_state* o) {
  memset(o, 0, sizeof(*o));
  port::Quadcopter& q = *v->quad;
  port::Logic& L = q.logic;
  o->pos[0] = q.pos.x; o->pos[1] = q.pos.y; o->pos[2] = q.pos.z;
  o->vel[0] = q.vel.x; o->vel[1] = q.vel.y; o->vel[2] = q.vel.z;
  for (int i = 0; i < 4; i++) o->att[i] = q.att.v[i];
  o->ang_vel[0] = q.angVel.x; o->ang_vel[1] = q.angVel.y; o->ang_vel[2] = q.angVel.z;
  for (int i = 0; i < 4; i++) {
    o->motor_speed[i] = q.motors[i].speed;
    o->motor_force_z[i] = q.motors[i].thrust.z;
    o->motor_speed_cmd[i] = q.cmd[i];
    o->des_motor_speeds[i] = L.desSpeeds[i];
    o->des_motor_forces[i] = L.desForces[i];
  }
  o->flight_state = L.state;
  o->first_panic_reason = L.firstPanic;
  o->cycle_counter = int(L.cycle);
  o->tel_warnings = L.warnings;
  dump3(L.gyro.lp, o->gyro_lpf);
  dump3(L.acc.lp, o->acc_lpf);
  o->temp_lpf[0] = L.temp.lp.xm0; o->temp_lpf[1] = L.temp.lp.xm1; o->temp_lpf[2] = L.temp.lp.ym0; o->temp_lpf[3] = L.temp.lp.ym1;
  o->batt_lpf[0] = L.battLp.xm0; o->batt_lpf[1] = L.battLp.xm1; o->batt_lpf[2] = L.battLp.ym0; o->batt_lpf[3] = L.battLp.ym1;
  o->batt_voltage_filtered = L.batt.vFilt;
  o->monitor_cmd_rate_lpdt = L.monCmd.lpDt;
  o->monitor_main_loop_lpdt = L.monLoop.lpDt;
  o->des_pos[0] = L.desPos.x; o->des_pos[1] = L.desPos.y; o->des_pos[2] = L.desPos.z;
  o->radio_type = L.radio.msg.type;
  o->radio_flags = L.radio.msg.flags;
  o->radio_count = L.radio.count;
  if (L.radio.count)
    for (int i = 0; i < 10; i++) o->radio_floats[i] = L.radio.msg.f[i];
  o->uwb_meas_count = L.uwb.count;
  o->next_ranging_target_idx = L.nextTargetIdx;
  const port::KalmanFilter& kf = L.kf;
  o->kf_pos[0] = kf.pos.x; o->kf_pos[1] = kf.pos.y; o->kf_pos[2] = kf.pos.z;
  o->kf_vel[0] = kf.vel.x; o->kf_vel[1] = kf.vel.y; o->kf_vel[2] = kf.vel.z;
  for (int i = 0; i < 4; i++) o->kf_att[i] = kf.att.v[i];
  o->kf_ang_vel[0] = kf.angVel.x; o->kf_ang_vel[1] = kf.angVel.y; o->kf_ang_vel[2] = kf.angVel.z;
  o->kf_last_corr[0] = kf.lastCorr.x; o->kf_last_corr[1] = kf.lastCorr.y; o->kf_last_corr[2] = kf.lastCorr.z;
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 9; j++) o->kf_cov[9 * i + j] = kf.cov.m[i][j];
  o->kf_imu_init = kf.imuInit;
  o->kf_uwb_init = kf.uwbInit;
  o->kf_num_resets = kf.numResets;
  o->kf_num_rejected = kf.numRejected;
  o->kf_num_rejected_seq = kf.numRejectedSeq;
  for (int i = 0; i < 6; i++) o->debug[i] = L.debug[i];
}

void orc_get_telemetry(orc_vehicle* v, uint8_t p1[30], uint8_t p2[30]) { v->quad->logic.telemetry(p1, p2); }

void orc_get_imu(orc_vehicle* v, double acc[3], double gyro[3]) {
  port::V3f a = v->quad->logic.acc.lp.value(), g = v->quad->logic.gyro.lp.value();
  acc[0] = a.x; acc[1] = a.y; acc[2] = a.z;
  gyro[0] = g.x; gyro[1] = g.y; gyro[2] = g.z;
}

uint64_t orc_time_us(orc_vehicle* v) { return v->clock.now_us; }

double orc_run_population(const agf_vehicle_cfg* cfgs, uint32_t n_cfgs, uint32_t n,
                          const orc_opts* opts, const double* init13, const float* anchors,
                          uint32_t n_anchors, uint32_t dt_us, uint32_t nticks,
                          const agf_cmd_entry* sched, uint32_t nsched, const uint8_t* slot_raw,
                          uint32_t threads, double* final_out) {
  std::vector<orc_vehicle*> vs(n);
  for (uint32_t i = 0; i < n; i++) {
    vs[i] = orc_create(&cfgs[n_cfgs == 1 ? 0 : i], opts);
    if (init13) {
      const double* s = init13 + 13 * size_t(i);
      orc_set_state(vs[i], s, s + 3, s + 6, s + 10);
    }
    for (uint32_t a = 0; a < n_anchors; a++)
      orc_add_anchor(vs[i], uint8_t(anchors[4 * a]), anchors[4 * a + 1], anchors[4 * a + 2], anchors[4 * a + 3]);
  }
  if (threads < 1) threads = 1;
  auto work = [&](uint32_t t) {
    uint32_t lo = uint32_t(uint64_t(n) * t / threads), hi = uint32_t(uint64_t(n) * (t + 1) / threads);
    uint8_t mine[AGF_MAX_CMD_SLOTS * AGF_RADIO_PACKET_SIZE];
    for (uint32_t i = lo; i < hi; i++) {
      const uint8_t* sr = nullptr;
      if (slot_raw) {
        for (int s = 0; s < AGF_MAX_CMD_SLOTS; s++)
          memcpy(mine + s * AGF_RADIO_PACKET_SIZE, slot_raw + (size_t(s) * n + i) * AGF_RADIO_PACKET_SIZE,
                 AGF_RADIO_PACKET_SIZE);
        sr = mine;
      }
      orc_run(vs[i], dt_us, nticks, sched, nsched, sr, nullptr);
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  if (threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < threads; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (uint32_t i = 0; i < n; i++) {
    if (final_out) record(vs[i], final_out + size_t(i) * ORC_NTRAJ);
    orc_destroy(vs[i]);
  }
  return secs;
}

#include "../orc_population_traj.inc"

// codec entry points of the port are only the decode side (the product owns the encoders and is
// checked against the reference's encoders through oracle/_ref)
void orc_radio_decode(const uint8_t raw[23], uint8_t* type, uint8_t* flags, float floats[10]) {
  port::RadioMsg m = port::radio_decode(raw);
  *type = m.type;
  *flags = m.flags;
  for (int i = 0; i < 10; i++) floats[i] = m.f[i];
}

}  // extern "C"
