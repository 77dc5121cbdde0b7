// agf_host_params.h -- host-side evaluation of everything the reference computes once per vehicle
// at construction time (Quadcopter_T::Quadcopter_T, Quadcopter_T.cpp:9-83; QuadcopterLogic::Initialise,
// QuadcopterLogic.cpp:97-162; QuadcopterMixer::SetParameters; the LowPassFilter Initialise methods),
// turned into the launch parameters of the step kernels.  Float arithmetic stays float, in the
// reference's order; must be compiled with -ffp-contract=off.
#pragma once
#include <math.h>
#include <string.h>

#include <vector>

#include "agf_math.h"
#include "agf_types.h"

namespace agf {

// scalar k of vehicle i inside a quad-grouped SoA with `lanes` scalars per quad
AGF_HDI size_t sidx(int k, size_t n, size_t i, int lanes) {
  return (size_t(k / lanes) * n + i) * lanes + (k % lanes);
}

inline Lpf2Coef lpf2_coef(float dt, float wc) {  // LowPassFilterSecondOrder.hpp:23-45 (float arithmetic)
  Lpf2Coef c;
  float const sqrt2 = float(sqrt(2.0));
  c.a1 = (dt * dt * wc * wc - 2 * sqrt2 * dt * wc + 4) / (dt * dt * wc * wc + 2 * sqrt2 * dt * wc + 4);
  c.a2 = 2 * (dt * dt * wc * wc - 4) / (dt * dt * wc * wc + 2 * sqrt2 * dt * wc + 4);
  c.b0 = dt * dt * wc * wc / (dt * dt * wc * wc + 2 * sqrt2 * dt * wc + 4);
  c.b1 = dt * dt * wc * wc / (dt * dt * wc * wc + 2 * sqrt2 * dt * wc + 4);
  c.b2 = 2 * dt * dt * wc * wc / (dt * dt * wc * wc + 2 * sqrt2 * dt * wc + 4);
  return c;
}

inline void quat_from_ypr(float y, float p, float r, float q[4]) {  // Rotation.hpp:99-110 with the shared libm
  const float h = 0.5f;
  q[0] = agf_cosf(h * y) * agf_cosf(h * p) * agf_cosf(h * r) + agf_sinf(h * y) * agf_sinf(h * p) * agf_sinf(h * r);
  q[1] = agf_cosf(h * y) * agf_cosf(h * p) * agf_sinf(h * r) - agf_sinf(h * y) * agf_sinf(h * p) * agf_cosf(h * r);
  q[2] = agf_cosf(h * y) * agf_sinf(h * p) * agf_cosf(h * r) + agf_sinf(h * y) * agf_cosf(h * p) * agf_sinf(h * r);
  q[3] = agf_sinf(h * y) * agf_cosf(h * p) * agf_cosf(h * r) - agf_cosf(h * y) * agf_sinf(h * p) * agf_sinf(h * r);
}
inline void quat_matrix(const float v[4], float R[9]) {  // Rotation.hpp:196-217
  const float r0 = v[0] * v[0], r1 = v[1] * v[1], r2 = v[2] * v[2], r3 = v[3] * v[3];
  R[0] = r0 + r1 - r2 - r3;
  R[1] = 2 * v[1] * v[2] - 2 * v[0] * v[3];
  R[2] = 2 * v[1] * v[3] + 2 * v[0] * v[2];
  R[3] = 2 * v[1] * v[2] + 2 * v[0] * v[3];
  R[4] = r0 - r1 + r2 - r3;
  R[5] = 2 * v[2] * v[3] - 2 * v[0] * v[1];
  R[6] = 2 * v[1] * v[3] - 2 * v[0] * v[2];
  R[7] = 2 * v[2] * v[3] + 2 * v[0] * v[1];
  R[8] = r0 - r1 - r2 + r3;
}

inline void derive_logic(const agf_logic_consts& k, float onboard_period, LogicParams& o) {  // QuadcopterLogic.cpp:97-162
  memset(&o, 0, sizeof(o));
  o.lp_acc = lpf2_coef(onboard_period, 100.0f);
  o.lp_gyro = lpf2_coef(onboard_period, 200.0f);
  o.lp_batt = lpf2_coef(onboard_period, 0.5f * float(2 * M_PI));
  o.lp_temp = lpf2_coef(onboard_period, 0.5f * float(2 * M_PI));
  float q[4], qi[4];
  quat_from_ypr(k.imu_yaw, k.imu_pitch, k.imu_roll, q);
  quat_matrix(q, o.R_imu);
  qi[0] = q[0]; qi[1] = -q[1]; qi[2] = -q[2]; qi[3] = -q[3];
  quat_matrix(qi, o.R_imu_inv);
  o.imu_identity = (k.imu_yaw == 0 && k.imu_pitch == 0 && k.imu_roll == 0) ? 1 : 0;
  o.mass = k.mass;
  o.ixx = k.inertia_xx;
  o.izz = k.inertia_zz;
  o.nat_freq = k.pos_control_nat_freq;
  o.damping = k.pos_control_damping;
  o.tc_att_xy = k.att_control_time_const_xy;
  o.tc_att_z = k.att_control_time_const_z;
  if (o.tc_att_z < o.tc_att_xy) o.tc_att_z = o.tc_att_xy;  // QuadcopterAttitudeController.hpp:19-24
  o.tc_w_xy = k.ang_vel_control_time_const_xy;
  o.tc_w_z = k.ang_vel_control_time_const_z;
  // QuadcopterMixer::SetParameters (QuadcopterMixer.hpp:36-51)
  o.mix_d = k.arm_length / sqrtf(2.0f);
  o.mix_kt = k.prop0_spin_dir * k.prop_torque_from_thrust;
  o.mix_kf = k.prop_thrust_from_speed_sqr;
  o.max_thrust = k.max_thrust_per_propeller;
  o.min_thrust = k.min_thrust_per_propeller;
  o.max_cmd_total = k.max_cmd_total_thrust < 0 ? 4 * k.max_thrust_per_propeller * 0.8f : k.max_cmd_total_thrust;
  o.batt_voltage = 1.2 * k.low_battery_threshold;  // Quadcopter_T.cpp:72 (double product, stored as float)
  o.batt_critical = k.low_battery_threshold;
  o.batt_warning = 1.05f * o.batt_critical;
  o.onboard_period = onboard_period;
  o.mon_cmd_coef = expf(-0.02f * 1.0f);             // QuadcopterLogic.cpp:14, LowPassFilterFirstOrder.hpp:31
  o.mon_loop_coef = expf(-onboard_period * 50.0f);  // QuadcopterLogic.cpp:15
  o.inv_mix_d = 1.0f / o.mix_d;
  o.inv_mix_kt = 1.0f / o.mix_kt;
  o.inv_mix_kf = 1.0f / o.mix_kf;
  o.inv_tc_w_xy = 1.0f / o.tc_w_xy;
  o.inv_tc_w_z = 1.0f / o.tc_w_z;
  o.k3_att = 1.0f / o.tc_att_z;
  o.k12_att = 1.0f / o.tc_att_xy;
  o.valid = k.valid;
}

// 3x3 inverse with the algorithm Eigen uses for fixed 3x3 (cofactors / determinant), which is what
// Quadcopter_T.cpp:20 `inertiaMatrix.inverse()` evaluates
inline void inverse3(const double a[9], double r[9]) {
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return a[3 * i1 + j1] * a[3 * i2 + j2] - a[3 * i1 + j2] * a[3 * i2 + j1];
  };
  const double c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
  const double det = c0 * a[0] + (c1 * a[3] + c2 * a[6]);
  const double invdet = 1.0 / det;
  r[0] = c0 * invdet; r[1] = c1 * invdet; r[2] = c2 * invdet;
  for (int row = 1; row < 3; row++)
    for (int col = 0; col < 3; col++) r[3 * row + col] = cof(col, row) * invdet;
}


// exp(-dt/tau) of Motor.cpp:53-57, evaluated with the host libm for the dt of the coming launch
inline double motor_c_host(double tau, uint32_t dt_us) {
  if (tau == 0) return 0;
  const double dt = double(uint64_t(dt_us) * double(1e-6));
  return exp(-dt / tau);
}

template<typename P>
inline void fill_plant(const agf_vehicle_cfg& c, PlantPV<P>& pv) {
  double inv[9];
  inverse3(c.inertia, inv);
  pv.mass = P(c.mass);
  for (int k = 0; k < 9; k++) { pv.I[k] = P(c.inertia[k]); pv.Iinv[k] = P(inv[k]); }
  pv.kF = P(c.prop_thrust_from_speed_sqr);
  pv.kTau = P(c.prop_torque_from_speed_sqr);
  pv.motor_c = P(0);
  pv.inv_mass = P(1.0 / c.mass);
}

// offboard loop parameters (agf_offboard_cfg -> OffboardParams + the clock-only thresholds); returns 0 or a reason
inline const char* fill_offboard(const agf_offboard_cfg& c, OffboardParams& off, TimingConsts& tc) {
  if (c.period_us == 0) return "offboard loop period must be > 0";
  if ((uint64_t(c.delay_us) + c.period_us - 1) / c.period_us + 1 > uint64_t(AGF_OFFQ))
    return "offboard loop: ceil(delay_us / period_us) + 1 commands in flight exceed AGF_OFFBOARD_QUEUE";
  if (!(c.att_control_time_const_xy > 0) || !(c.att_control_time_const_z > 0)) return "offboard loop: attitude time constants must be > 0";
  off.nat_freq = c.pos_control_nat_freq;
  off.damping = c.pos_control_damping;
  off.tc_att_xy = c.att_control_time_const_xy;
  off.tc_att_z = c.att_control_time_const_z;
  off.k3_att = 1.0f / c.att_control_time_const_z;
  off.k12_att = 1.0f / c.att_control_time_const_xy;
  off.max_proper = c.max_proper_acc;
  off.min_vert = c.min_vertical_proper_acc;
  off.min_proper = c.min_proper_acc;
  off.yaw = float(c.yaw_angle);
  off.flags = c.radio_flags & 0xFFu;
  tc.off_enabled = 1;
  timing_thresholds_offboard(tc, double(c.period_us) * 1e-6);
  tc.off_delay_us = c.delay_us;
  return nullptr;
}

// everything of StepShared that does not depend on device pointers, anchors or noise settings
template<typename P>
inline void build_shared(const agf_vehicle_cfg& c0, double onboard_logic_period, double uwb_comm_period, StepShared<P>& sh) {
  memset(&sh, 0, sizeof(sh));
  derive_logic(c0.logic, float(onboard_logic_period), sh.logic);
  sh.tc.logic_period = onboard_logic_period;
  sh.tc.logic_adj_us = uint32_t(uint64_t((-onboard_logic_period) * double(-1e6)));  // Timer.hpp:31-33
  sh.tc.comm_period = uwb_comm_period;
  sh.tc.net_enabled = uwb_comm_period > 0 ? 1 : 0;
  sh.tc.n_anchors = 0;
  timing_thresholds(sh.tc);
  sh.motor_min = P(c0.motor_min_speed);
  sh.motor_max = P(c0.motor_max_speed);
  sh.motor_J = P(c0.motor_inertia);
  const double a = c0.arm_length / sqrt(2);  // Quadcopter_T.cpp:45-65
  const double sx[4] = {+1, -1, -1, +1}, sy[4] = {-1, -1, +1, +1};
  for (int m = 0; m < 4; m++) {
    sh.motor_pos[m][0] = P(a * sx[m] + c0.com_error[0]);
    sh.motor_pos[m][1] = P(a * sy[m] + c0.com_error[1]);
    sh.motor_pos[m][2] = P(a * 0.0 + c0.com_error[2]);
  }
  sh.has_drag = 0;
  for (int k = 0; k < 3; k++) {
    sh.drag[k] = P(c0.lin_drag_coeff_b[k]);
    if (c0.lin_drag_coeff_b[k] != 0.0) sh.has_drag = 1;
  }
}

// constructor-time state (SimulationObject6DOF.hpp:14-19, QuadcopterLogic::ResetCounters/Initialise,
// KalmanFilter6DOF::Reset) replicated for every vehicle, in the HBM layout of agf_types.h
template<typename P>
inline void initial_state(size_t n, const LogicParams& lg, float low_battery_threshold, bool uwb, std::vector<P>& hp,
                          std::vector<float>& hf, std::vector<uint32_t>& hu, std::vector<float>& hc) {
  const int VP = int(16 / sizeof(P));
  hp.assign(size_t(NP_PAD) * n, P(0));
  hf.assign(size_t(NF_PAD) * n, 0.0f);
  hu.assign(size_t(NU_PAD) * n, 0u);
  const float battInit = low_battery_threshold * 1.2f;  // QuadcopterLogic.cpp:138-139
  for (size_t i = 0; i < n; i++) {
    hp[sidx(SP_ATT, n, i, VP)] = P(1);
    hf[sidx(SF_KATT, n, i, 4)] = 1.0f;
    for (int k = 0; k < 4; k++) {
      hf[sidx(SF_TEMP_LP + k, n, i, 4)] = 25.0f;
      hf[sidx(SF_BATT_LP + k, n, i, 4)] = battInit;
      hf[sidx(SF_PC_CORR + k, n, i, 4)] = 1.0f;
    }
    hf[sidx(SF_MON_CMD, n, i, 4)] = 0.02f;
    hf[sidx(SF_MON_LOOP, n, i, 4)] = lg.onboard_period;
    uint32_t bits = 0;
    if (lg.valid) {
      bits |= AGF_FS_IDLE;
    } else {
      bits |= AGF_FS_KILLED;
      bits |= uint32_t(AGF_PANIC_KILLED_INTERNALLY) << 3;
    }
    bits |= (1u << 21);  // numResets (1 after Initialise) != lastCheckNumResets (0)
    hu[sidx(SU_BITS, n, i, 4)] = bits;
    hu[sidx(SU_KFCNT, n, i, 4)] = 1u;
  }
  hc.clear();
  if (uwb) {
    hc.assign(size_t(NC_PAD) * n, 0.0f);
    const float sp = 3.0f, sv = 3.0f;
    const float sperp = 10.0f * float(M_PI) / 180.0f, sabout = 30.0f * float(M_PI) / 180.0f;
    for (size_t i = 0; i < n; i++) {
      for (int k = 0; k < 3; k++) {
        hc[sidx(9 * k + k, n, i, 4)] = sp * sp;
        hc[sidx(9 * (3 + k) + 3 + k, n, i, 4)] = sv * sv;
      }
      hc[sidx(9 * 6 + 6, n, i, 4)] = sperp * sperp;
      hc[sidx(9 * 7 + 7, n, i, 4)] = sperp * sperp;
      hc[sidx(9 * 8 + 8, n, i, 4)] = sabout * sabout;
    }
  }
}

}  // namespace agf
