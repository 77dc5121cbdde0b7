// agf_quadcopter.hpp -- host-side C++ mirror of the reference's Components/Simulation object API on top of
// the C ABI (include/agrifly_b200.h).  Header only; link with libagrifly_b200.so.
//
// What it mirrors (reference file:line):
//   Simulation::SimulationObject6DOF          Components/Components/Simulation/SimulationObject6DOF.hpp:12-86
//   Simulation::Quadcopter_T<QuadcopterLogic> Components/Components/Simulation/Quadcopter_T.hpp:21-134
//   BaseTimer / ManualTimer                   Common/Common/Time/BaseTimer.hpp, ManualTimer.hpp:20-45
//   Vec3 / Rotation value types               Common/Common/Math/Vec3.hpp:28, Rotation.hpp:28
//   RadioMessageDecoded::RawMessage           Common/Common/DataTypes/RadioTypes.hpp:68-70
//   TelemetryPacket::data_packet_t            Common/Common/DataTypes/TelemetryPacket.hpp:32-36
//
// Two ways to use it:
//   * stand-alone (default): the small value types below live in namespace agf;
//   * inside the reference tree: compile with -DAGF_WITH_REFERENCE_HEADERS and the reference's include paths
//     (-I<agri-fly>/Common -I<agri-fly>/Components); the facade then takes and returns the reference's own
//     Vec3d / Rotationd / BaseTimer / RawMessage / data_packet_t, so a Rappids_Simulator-style loop
//     (Simulator/Rappids_Simulator/main.cpp:211-218,279-280,391-395,738) compiles unchanged after replacing
//     `Simulation::Quadcopter` by `agf::Simulation::Quadcopter`.
//
// A QuadcopterBatch owns one agf_batch handle (N vehicles on one GPU); a Quadcopter is either a view of one
// vehicle of a batch or, built with the reference's 15-argument constructor, a batch of one.  Semantics of
// Run(): the reference's Run() integrates over the time since the previous Run() as read from the shared
// BaseTimer (Quadcopter_T.cpp:87-91).  The facade reads the same timer, ages the batch's stopwatches by the
// elapsed microseconds (agf_batch_advance_clock) and issues exactly one Run() (agf_batch_run(b, 0, 1)).  Every
// vehicle of a batch steps on the first Run() after the timer moved; further Run() calls at the same timer
// reading (the reference loop `for (auto v : vehicles) v->Run();`, Simulator/main.cpp:323-325) return at once,
// like the reference's own early return for dt < 1 us.
//
// Errors: the reference's step has no error channel; a failing CUDA call is exceptional and throws
// std::runtime_error with agf_last_error_string().  There is no CPU fallback.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "agrifly_b200.h"

#if defined(AGF_WITH_REFERENCE_HEADERS)
#include "Common/DataTypes/RadioTypes.hpp"
#include "Common/DataTypes/TelemetryPacket.hpp"
#include "Common/Math/Rotation.hpp"
#include "Common/Math/Vec3.hpp"
#include "Common/Time/BaseTimer.hpp"
#include "Common/Time/ManualTimer.hpp"
#include "Components/Logic/QuadcopterConstants.hpp"
#endif

namespace agf {

#if defined(AGF_WITH_REFERENCE_HEADERS)
using ::BaseTimer;
using ::ManualTimer;
using ::Rotationd;
using ::Rotationf;
using ::Vec3d;
using ::Vec3f;
typedef RadioTypes::RadioMessageDecoded::RawMessage RawMessage;
typedef TelemetryPacket::data_packet_t data_packet_t;
typedef Onboard::QuadcopterConstants::QuadcopterType QuadcopterType;
inline int quad_type_to_abi(QuadcopterType t) {  // QuadcopterConstants.hpp:16-24, same order as AGF_QC_TYPE_*
  return int(t);
}
template<class M> inline double mat3_at(const M& m, int r, int c) { return m(r, c); }
#else
// ---- minimal stand-ins with the reference's member names ----------------------------------------
template<typename Real>
struct Vec3 {  // Vec3.hpp:28 (default = NaN, :35)
  Real x, y, z;
  Vec3() : x(std::numeric_limits<Real>::quiet_NaN()), y(x), z(x) {}
  Vec3(Real a, Real b, Real c) : x(a), y(b), z(c) {}
  template<typename R2> explicit Vec3(const Vec3<R2>& o) : x(Real(o.x)), y(Real(o.y)), z(Real(o.z)) {}
  Vec3 operator+(const Vec3& r) const { return Vec3(x + r.x, y + r.y, z + r.z); }
  Vec3 operator-(const Vec3& r) const { return Vec3(x - r.x, y - r.y, z - r.z); }
  Vec3 operator*(Real s) const { return Vec3(x * s, y * s, z * s); }
  Vec3 operator/(Real s) const { return Vec3(x / s, y / s, z / s); }
  Real Dot(const Vec3& r) const { return x * r.x + y * r.y + z * r.z; }
  Vec3 Cross(const Vec3& r) const { return Vec3(y * r.z - z * r.y, z * r.x - x * r.z, x * r.y - y * r.x); }
  Real GetNorm2Squared() const { return Dot(*this); }
  Real GetNorm2() const { return std::sqrt(Dot(*this)); }
  Real operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
typedef Vec3<double> Vec3d;
typedef Vec3<float> Vec3f;

template<typename Real>
class Rotation {  // Rotation.hpp:28: unit quaternion [w,x,y,z], <world> = att * <body>
 public:
  Rotation() : _v{1, 0, 0, 0} {}
  Rotation(Real a, Real b, Real c, Real d) : _v{a, b, c, d} {}
  static Rotation Identity() { return Rotation(1, 0, 0, 0); }
  Rotation Inverse() const { return Rotation(_v[0], -_v[1], -_v[2], -_v[3]); }
  Real operator[](unsigned i) const { return _v[i]; }
  Real& operator[](unsigned i) { return _v[i]; }
  static Rotation FromEulerYPR(Real y, Real p, Real r) {  // Rotation.hpp:99-110
    const Real h = Real(0.5);
    Rotation o;
    o._v[0] = std::cos(h * y) * std::cos(h * p) * std::cos(h * r) + std::sin(h * y) * std::sin(h * p) * std::sin(h * r);
    o._v[1] = std::cos(h * y) * std::cos(h * p) * std::sin(h * r) - std::sin(h * y) * std::sin(h * p) * std::cos(h * r);
    o._v[2] = std::cos(h * y) * std::sin(h * p) * std::cos(h * r) + std::sin(h * y) * std::cos(h * p) * std::sin(h * r);
    o._v[3] = std::sin(h * y) * std::cos(h * p) * std::cos(h * r) - std::cos(h * y) * std::sin(h * p) * std::sin(h * r);
    return o;
  }
  void ToEulerYPR(Real& y, Real& p, Real& r) const {  // Rotation.hpp:163-169
    const Real *q = _v;
    y = std::atan2(Real(2) * q[1] * q[2] + Real(2) * q[0] * q[3], q[1] * q[1] + q[0] * q[0] - q[3] * q[3] - q[2] * q[2]);
    p = -std::asin(Real(2) * q[1] * q[3] - Real(2) * q[0] * q[2]);
    r = std::atan2(Real(2) * q[2] * q[3] + Real(2) * q[0] * q[1], q[3] * q[3] - q[2] * q[2] - q[1] * q[1] + q[0] * q[0]);
  }
  Rotation operator*(const Rotation& r1) const {  // Rotation.hpp:124-131
    const Real* a = _v;
    return Rotation(r1[0] * a[0] - r1[1] * a[1] - r1[2] * a[2] - r1[3] * a[3],
                    r1[1] * a[0] + r1[0] * a[1] + r1[3] * a[2] - r1[2] * a[3],
                    r1[2] * a[0] - r1[3] * a[1] + r1[0] * a[2] + r1[1] * a[3],
                    r1[3] * a[0] + r1[2] * a[1] - r1[1] * a[2] + r1[0] * a[3]);
  }
  Vec3<Real> operator*(const Vec3<Real>& v) const {  // Rotate, Rotation.hpp:196-245
    const Real w = _v[0], x = _v[1], y = _v[2], z = _v[3];
    const Real r0 = w * w, r1 = x * x, r2 = y * y, r3 = z * z;
    return Vec3<Real>((r0 + r1 - r2 - r3) * v.x + (2 * x * y - 2 * w * z) * v.y + (2 * x * z + 2 * w * y) * v.z,
                      (2 * x * y + 2 * w * z) * v.x + (r0 - r1 + r2 - r3) * v.y + (2 * y * z - 2 * w * x) * v.z,
                      (2 * x * z - 2 * w * y) * v.x + (2 * y * z + 2 * w * x) * v.y + (r0 - r1 - r2 + r3) * v.z);
  }

 private:
  Real _v[4];
};
typedef Rotation<double> Rotationd;
typedef Rotation<float> Rotationf;

struct Mat3d {  // stands in for Eigen::Matrix<double,3,3> in the constructor (Quadcopter_T.hpp:25)
  double d[3][3];
  Mat3d() { std::memset(d, 0, sizeof(d)); }
  double& operator()(int r, int c) { return d[r][c]; }
  double operator()(int r, int c) const { return d[r][c]; }
};
template<class M> inline double mat3_at(const M& m, int r, int c) { return m(r, c); }

class BaseTimer {  // BaseTimer.hpp
 public:
  virtual ~BaseTimer() {}
  virtual uint64_t GetMicroSeconds(void) const = 0;
};
class ManualTimer : public BaseTimer {  // ManualTimer.hpp:20-45
 public:
  ManualTimer() : _currentTime(0) {}
  void ResetMicroseconds(uint64_t t_us) { _currentTime = t_us; }
  void AdvanceMicroSeconds(uint64_t dt_us) { _currentTime += dt_us; }
  template<typename Real> Real GetSeconds(void) const { return (Real)(GetMicroSeconds() * Real(1e-6)); }
  virtual uint64_t GetMicroSeconds(void) const { return _currentTime; }

 private:
  uint64_t _currentTime;
};

struct RawMessage {  // RadioTypes.hpp:68-70
  uint8_t raw[AGF_RADIO_PACKET_SIZE];
};
struct data_packet_t {  // TelemetryPacket.hpp:32-36
  uint8_t type;
  uint8_t packetNumber;
  uint16_t data[14];
} __attribute__((packed));
static_assert(sizeof(data_packet_t) == AGF_TELEMETRY_PACKET_SIZE, "telemetry packet layout");

enum QuadcopterType {  // QuadcopterConstants.hpp:16-24
  QC_TYPE_INVALID = AGF_QC_TYPE_INVALID,
  QC_TYPE_CF_STANDARD = AGF_QC_TYPE_CF_STANDARD,
  QC_TYPE_CF_BIGMOTORSPROPS = AGF_QC_TYPE_CF_BIGMOTORSPROPS,
  QC_TYPE_CF_FEEDTHROUGH = AGF_QC_TYPE_CF_FEEDTHROUGH,
  QC_TYPE_CF_LARGEQUAD = AGF_QC_TYPE_CF_LARGEQUAD,
  QC_TYPE_CF_MINIQUAD = AGF_QC_TYPE_CF_MINIQUAD
};
inline int quad_type_to_abi(QuadcopterType t) { return int(t); }

// RadioTypes.hpp:123-187 command builders (the lossy 16-bit quantisation is part of the loop)
namespace RadioTypes {
struct RadioMessageDecoded {
  enum { RAW_PACKET_SIZE = AGF_RADIO_PACKET_SIZE };
  typedef agf::RawMessage RawMessage;
  static void CreateKillCommand(uint8_t flags, uint8_t raw[RAW_PACKET_SIZE]) { agf_radio_encode_kill(flags, raw); }
  static void CreateIdleCommand(uint8_t flags, uint8_t raw[RAW_PACKET_SIZE]) { agf_radio_encode_idle(flags, raw); }
  static void CreatePositionCommand(uint8_t flags, Vec3f pos, Vec3f vel, Vec3f acc, uint8_t raw[RAW_PACKET_SIZE]) {
    const float p[3] = {pos.x, pos.y, pos.z}, v[3] = {vel.x, vel.y, vel.z}, a[3] = {acc.x, acc.y, acc.z};
    agf_radio_encode_position(flags, p, v, a, raw);
  }
  static void CreateRatesCommand(uint8_t flags, float totalThrust, Vec3f angVel, uint8_t raw[RAW_PACKET_SIZE]) {
    const float w[3] = {angVel.x, angVel.y, angVel.z};
    agf_radio_encode_rates(flags, totalThrust, w, raw);
  }
  static void CreateAccelerationCommand(uint8_t flags, Vec3f acc, float yawRate, uint8_t raw[RAW_PACKET_SIZE]) {
    const float a[3] = {acc.x, acc.y, acc.z};
    agf_radio_encode_acceleration(flags, a, yawRate, raw);
  }
};
}  // namespace RadioTypes
#endif  // AGF_WITH_REFERENCE_HEADERS

inline void check(int rc, const char* what) {
  if (rc != AGF_OK) throw std::runtime_error(std::string(what) + ": " + agf_last_error_string());
}

namespace Simulation {

// The 15 constructor arguments of Quadcopter_T (Quadcopter_T.hpp:24-32) as an agf_vehicle_cfg; the firmware
// constants come from the airframe table of `quadcopterType` exactly as Quadcopter_T.cpp:71-82 looks them up.
template<class Mat3>
inline agf_vehicle_cfg MakeVehicleConfig(double mass, const Mat3& inertiaMatrix, double armLength, Vec3d centreOfMassError,
                                         double motorMinSpeed, double motorMaxSpeed, double propThrustFromSpeedSqr,
                                         double propTorqueFromSpeedSqr, double motorTimeConst, double motorInertia,
                                         Vec3d linDragCoeffB, uint8_t id, QuadcopterType quadcopterType) {
  agf_vehicle_cfg c;
  check(agf_vehicle_cfg_from_type(quad_type_to_abi(quadcopterType), id, &c), "agf_vehicle_cfg_from_type");
  c.mass = mass;
  for (int r = 0; r < 3; r++)
    for (int q = 0; q < 3; q++) c.inertia[3 * r + q] = mat3_at(inertiaMatrix, r, q);
  c.arm_length = armLength;
  c.com_error[0] = centreOfMassError.x; c.com_error[1] = centreOfMassError.y; c.com_error[2] = centreOfMassError.z;
  c.motor_min_speed = motorMinSpeed;
  c.motor_max_speed = motorMaxSpeed;
  c.prop_thrust_from_speed_sqr = propThrustFromSpeedSqr;
  c.prop_torque_from_speed_sqr = propTorqueFromSpeedSqr;
  c.motor_time_const = motorTimeConst;
  c.motor_inertia = motorInertia;
  c.lin_drag_coeff_b[0] = linDragCoeffB.x; c.lin_drag_coeff_b[1] = linDragCoeffB.y; c.lin_drag_coeff_b[2] = linDragCoeffB.z;
  c.vehicle_id = id;
  return c;
}

// N vehicles behind one handle.  Fast path for populations: RunTicks(); object-API path: Run().
class QuadcopterBatch {
 public:
  // cfgs.size() == 1: all vehicles share it; == n: one per vehicle (parameter sweeps)
  QuadcopterBatch(BaseTimer* masterTimer, const std::vector<agf_vehicle_cfg>& cfgs, size_t n, const agf_batch_opts& opts)
      : _timer(masterTimer), _h(nullptr), _n(n) {
    check(agf_batch_create(cfgs.data(), cfgs.size(), n, &opts, &_h), "agf_batch_create");
    _lastTimerUs = masterTimer ? masterTimer->GetMicroSeconds() : 0;
  }
  ~QuadcopterBatch() {
    if (_h) agf_batch_destroy(_h);
  }
  QuadcopterBatch(const QuadcopterBatch&) = delete;
  QuadcopterBatch& operator=(const QuadcopterBatch&) = delete;

  static agf_batch_opts DefaultOptions(double onboardLogicPeriod = 1.0 / 500) {
    agf_batch_opts o;
    agf_batch_opts_default(&o);
    o.onboard_logic_period = onboardLogicPeriod;
    return o;
  }

  size_t size() const { return _n; }
  agf_batch* handle() const { return _h; }

  // Quadcopter_T::Run for every vehicle, at the master timer's current reading (see the header comment)
  void Run() {
    const uint64_t now = _timer->GetMicroSeconds();
    if (_ranOnce && now == _lastTimerUs) return;  // same reading: dt < 1 us, Quadcopter_T.cpp:88-90
    const uint64_t adv = now - _lastTimerUs;
    if (adv > 0xFFFFFFFFull) throw std::runtime_error("QuadcopterBatch::Run: timer advanced by more than 2^32 us");
    if (adv) check(agf_batch_advance_clock(_h, uint32_t(adv)), "agf_batch_advance_clock");
    check(agf_batch_run(_h, 0, 1), "agf_batch_run");
    _lastTimerUs = now;
    _ranOnce = true;
  }
  // nticks x { Run(); timer += dt_us } in one kernel launch, state in registers in between.  The caller owns
  // the timer: advance it by nticks*dt_us afterwards (a ManualTimer can be passed to have that done here).
  void RunTicks(uint32_t nticks, uint32_t dt_us, ManualTimer* advance = nullptr) {
    const uint64_t now = _timer->GetMicroSeconds();
    const uint64_t adv = now - _lastTimerUs;
    if (adv) check(agf_batch_advance_clock(_h, uint32_t(adv)), "agf_batch_advance_clock");
    check(agf_batch_run(_h, dt_us, nticks), "agf_batch_run");
    _lastTimerUs = now + uint64_t(nticks) * dt_us;
    _ranOnce = true;
    if (advance) advance->AdvanceMicroSeconds(uint64_t(nticks) * dt_us);
  }
  void Sync() { check(agf_batch_sync(_h), "agf_batch_sync"); }

  void AddUWBRadioTarget(uint8_t id, Vec3f pos) {
    const float p[3] = {pos.x, pos.y, pos.z};
    check(agf_batch_add_uwb_anchor(_h, id, p), "agf_batch_add_uwb_anchor");
  }
  void SetCommandRadioMsgAll(const RawMessage& raw) { check(agf_batch_set_radio_cmd(_h, raw.raw, 0, _n, 1), "agf_batch_set_radio_cmd"); }

  // whole-population field access, [n][ncomp] (AGF_F_*)
  void GetField(int field, void* dst, size_t first, size_t count) { check(agf_batch_get_field(_h, field, dst, first, count), "agf_batch_get_field"); }
  void SetField(int field, const void* src, size_t first, size_t count) { check(agf_batch_set_field(_h, field, src, first, count), "agf_batch_set_field"); }

 private:
  BaseTimer* _timer;
  agf_batch* _h;
  size_t _n;
  uint64_t _lastTimerUs = 0;
  bool _ranOnce = false;
};

// One vehicle with the reference's method names.
class Quadcopter {
 public:
  // Quadcopter_T::Quadcopter_T (Quadcopter_T.hpp:24-32): a batch of one
  template<class Mat3>
  Quadcopter(BaseTimer* const masterTimer, double mass, const Mat3& inertiaMatrix, double armLength, Vec3d centreOfMassError,
             double motorMinSpeed, double motorMaxSpeed, double propThrustFromSpeedSqr, double propTorqueFromSpeedSqr,
             double motorTimeConst, double motorInertia, Vec3d linDragCoeffB, uint8_t id, QuadcopterType quadcopterType,
             double onboardLogicPeriod, const agf_batch_opts* opts = nullptr)
      : _i(0) {
    std::vector<agf_vehicle_cfg> c(1, MakeVehicleConfig(mass, inertiaMatrix, armLength, centreOfMassError, motorMinSpeed, motorMaxSpeed,
                                                        propThrustFromSpeedSqr, propTorqueFromSpeedSqr, motorTimeConst, motorInertia,
                                                        linDragCoeffB, id, quadcopterType));
    agf_batch_opts o = opts ? *opts : QuadcopterBatch::DefaultOptions(onboardLogicPeriod);
    o.onboard_logic_period = onboardLogicPeriod;
    _b = std::make_shared<QuadcopterBatch>(masterTimer, c, 1, o);
  }
  // view of vehicle i of an existing batch
  Quadcopter(std::shared_ptr<QuadcopterBatch> batch, size_t i) : _b(batch), _i(i) {
    if (i >= batch->size()) throw std::out_of_range("Quadcopter: index outside the batch");
  }

  void Run() { _b->Run(); }

  // SimulationObject6DOF.hpp:26-56
  Vec3d GetPosition() const { double v[3]; _b->GetField(AGF_F_POSITION, v, _i, 1); return Vec3d(v[0], v[1], v[2]); }
  Vec3d GetVelocity() const { double v[3]; _b->GetField(AGF_F_VELOCITY, v, _i, 1); return Vec3d(v[0], v[1], v[2]); }
  Rotationd GetAttitude() const { double v[4]; _b->GetField(AGF_F_ATTITUDE, v, _i, 1); return Rotationd(v[0], v[1], v[2], v[3]); }
  Vec3d GetAngularVelocity() const { double v[3]; _b->GetField(AGF_F_ANGULAR_VELOCITY, v, _i, 1); return Vec3d(v[0], v[1], v[2]); }
  void SetPosition(Vec3d in) { const double v[3] = {in.x, in.y, in.z}; _b->SetField(AGF_F_POSITION, v, _i, 1); }
  void SetVelocity(Vec3d in) { const double v[3] = {in.x, in.y, in.z}; _b->SetField(AGF_F_VELOCITY, v, _i, 1); }
  void SetAttitude(Rotationd in) { const double v[4] = {in[0], in[1], in[2], in[3]}; _b->SetField(AGF_F_ATTITUDE, v, _i, 1); }
  void SetAngularVelocity(Vec3d in) { const double v[3] = {in.x, in.y, in.z}; _b->SetField(AGF_F_ANGULAR_VELOCITY, v, _i, 1); }

  // Quadcopter_T.hpp:39-83
  double GetMotorForce(unsigned i) const { double v[4]; _b->GetField(AGF_F_MOTOR_FORCE, v, _i, 1); return v[i]; }
  void SetExternalForce(Vec3d in) { const double v[3] = {in.x, in.y, in.z}; check(agf_batch_set_external_wrench(_b->handle(), v, nullptr, _i, 1), "agf_batch_set_external_wrench"); }
  void SetExternalTorque(Vec3d in) { const double v[3] = {in.x, in.y, in.z}; check(agf_batch_set_external_wrench(_b->handle(), nullptr, v, _i, 1), "agf_batch_set_external_wrench"); }
  void GetEstimate(Vec3f& pos, Vec3f& vel, Rotationf& att, Vec3f& angVel) const {
    float p[3], v[3], q[4], w[3];
    _b->GetField(AGF_F_EST_POSITION, p, _i, 1);
    _b->GetField(AGF_F_EST_VELOCITY, v, _i, 1);
    _b->GetField(AGF_F_EST_ATTITUDE, q, _i, 1);
    _b->GetField(AGF_F_EST_ANGULAR_VELOCITY, w, _i, 1);
    pos = Vec3f(p[0], p[1], p[2]); vel = Vec3f(v[0], v[1], v[2]); att = Rotationf(q[0], q[1], q[2], q[3]); angVel = Vec3f(w[0], w[1], w[2]);
  }
  // anchors are shared by the vehicles of a batch (each vehicle ranges in a private network)
  void AddUWBRadioTarget(uint8_t id, Vec3f pos) { _b->AddUWBRadioTarget(id, pos); }
  void SetCommandRadioMsg(RawMessage const raw) { check(agf_batch_set_radio_cmd(_b->handle(), raw.raw, _i, 1, 1), "agf_batch_set_radio_cmd"); }
  void GetTelemetryDataPackets(data_packet_t& dataPacket1, data_packet_t& dataPacket2) {
    check(agf_batch_get_telemetry(_b->handle(), reinterpret_cast<uint8_t*>(&dataPacket1), reinterpret_cast<uint8_t*>(&dataPacket2), _i, 1), "agf_batch_get_telemetry");
  }
  void GetAccelerometer(Vec3d& acc) { float v[3]; _b->GetField(AGF_F_ACCELEROMETER, v, _i, 1); acc = Vec3d(v[0], v[1], v[2]); }
  void GetRateGyro(Vec3d& rateGyro) { float v[3]; _b->GetField(AGF_F_RATE_GYRO, v, _i, 1); rateGyro = Vec3d(v[0], v[1], v[2]); }

  // beyond the reference's public surface (its members are private there)
  int GetFlightState() const { int32_t v; _b->GetField(AGF_F_FLIGHT_STATE, &v, _i, 1); return v; }
  int GetFirstPanicReason() const { int32_t v; _b->GetField(AGF_F_PANIC_REASON, &v, _i, 1); return v; }
  double GetMotorSpeed(unsigned i) const { double v[4]; _b->GetField(AGF_F_MOTOR_SPEED, v, _i, 1); return v[i]; }
  std::shared_ptr<QuadcopterBatch> GetBatch() const { return _b; }
  size_t GetIndex() const { return _i; }

 private:
  std::shared_ptr<QuadcopterBatch> _b;
  size_t _i;
};

}  // namespace Simulation
}  // namespace agf
