// agf_types.h -- data layout and launch-parameter types shared by the host handle and the kernels.
//
// HBM layout of a batch of N vehicles (DESIGN.md "Data layout"): four structure-of-arrays
// groups, each an array of 16-byte quads indexed [quad][vehicle]:
//   sp  plant state            NP_PAD scalars of the plant precision P  (double2 or float4 quads)
//   sf  onboard-logic floats   NF_PAD floats                             (float4 quads)
//   su  flags/counters/ages    NU_PAD uint32                             (uint4 quads)
//   sc  EKF covariance         NC_PAD floats, only for UWB batches       (float4 quads)
// A warp reading quad q of 32 consecutive vehicles touches 512 (or 1024) contiguous bytes.
#pragma once

#include <stdint.h>
#include <stddef.h>

#include "agrifly_b200.h"

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define AGF_HDI __host__ __device__ __forceinline__
#else
#define AGF_HDI inline
#endif

namespace agf {

// ---- flat state tables ------------------------------------------------------------------------
enum {  // plant scalars (precision P)
  SP_POS = 0, SP_VEL = 3, SP_ATT = 6, SP_W = 10, SP_MS = 13,
  SP_CPOS = 17,  // FP32 fast variants: running compensation of the position sum (see tick(): compensated integration)
  SP_RPOS = 20,  // UWB radio's latched true position (UWBRadio::_uwbTruePosition); 23 pad
  SP_CVEL = 24,  // FP32 fast variants: running compensation of the velocity sum; 27 pad
  NP_CORE = 20, NP_REF = 24 /* what the FP64-plant variants load and store */, NP_PAD = 28
};
enum {  // logic floats; the HK block follows the core block
  SF_CMD = 0, SF_DFORCE = 4, SF_RADIO = 8, SF_KATT = 12, SF_GYRO_LP = 16, SF_ACC_LP = 28,
  SF_KPOS = 40, SF_KVEL = 43, SF_KW = 46, SF_KCORR = 49, SF_UWB_RANGE = 52, SF_LOGIC_RANGE = 53,
  NF_CORE = 56,
  SF_TEMP_LP = 56, SF_BATT_LP = 60, SF_PC_ACCUM = 64, SF_PC_CORR = 68, SF_BATT_VFILT = 72,
  SF_MON_CMD = 73, SF_MON_LOOP = 74,
  NF_PAD = 76
};
enum {  // uint32 words
  SU_BITS = 0, SU_CNT = 1, SU_CYCLE = 2, SU_KFCNT = 3, SU_UWB_COUNT = 4, SU_AGE_RADIO = 5,
  SU_AGE_UWB = 6, SU_UWBW = 7,
  NU_CORE = 8,
  SU_AGE_EST_RESET = 8, SU_PC_COUNT = 9, SU_AGE_MON_CMD = 10, SU_AGE_MON_LOOP = 11,
  NU_PAD = 12
};
enum { NC_PAD = 84 };
enum { AGF_OFFQ = AGF_OFFBOARD_QUEUE };  // offboard-loop commands in flight per vehicle

template<typename P> struct VecOf;
#if defined(__CUDACC__)
template<> struct VecOf<double> {
  typedef double2 type;
  static constexpr int lanes = 2;
  static AGF_HDI void unpack(const double2& v, double* o) { o[0] = v.x; o[1] = v.y; }
  static AGF_HDI double2 pack(const double* o) { return make_double2(o[0], o[1]); }
};
template<> struct VecOf<float> {
  typedef float4 type;
  static constexpr int lanes = 4;
  static AGF_HDI void unpack(const float4& v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
  static AGF_HDI float4 pack(const float* o) { return make_float4(o[0], o[1], o[2], o[3]); }
};

template<typename P>
struct StateArrays {
  typename VecOf<P>::type* sp;
  float4* sf;
  uint4* su;
  float4* sc;  // null unless the batch ranges against UWB anchors
  float4* sq;  // offboard-loop command queue [AGF_OFFQ][N], null unless the loop is on
};
#endif

// ---- shared (per batch) parameters -----------------------------------------------------------
struct Lpf2Coef {  // LowPassFilterSecondOrder coefficients, computed on the host in float
  float a1, a2, b0, b1, b2;
};

struct LogicParams {  // what QuadcopterLogic::Initialise derives from QuadcopterConstants
  Lpf2Coef lp_gyro, lp_acc, lp_temp, lp_batt;
  float R_imu[9], R_imu_inv[9];
  int imu_identity;
  float mass, ixx, izz;
  float nat_freq, damping;
  float tc_att_xy, tc_att_z, tc_w_xy, tc_w_z;
  float mix_d, mix_kt, mix_kf, max_cmd_total, min_thrust, max_thrust;
  float batt_voltage, batt_warning, batt_critical;
  float onboard_period;
  float mon_cmd_coef, mon_loop_coef;  // LowPassFilterFirstOrder coefficients (expf on the host)
  // reciprocals used by the fast-arithmetic variants only
  float inv_mix_d, inv_mix_kt, inv_mix_kf, inv_tc_w_xy, inv_tc_w_z, k3_att, k12_att;
  int valid;
};

struct TimingConsts {
  double logic_period;    // Quadcopter_T::_onboardLogicPeriod
  uint32_t logic_adj_us;  // uint64_t(period * 1e6) of Timer::AdjustTimeBySeconds
  double comm_period;     // UWBNetwork::_commPeriod
  int net_enabled;
  int n_anchors;
  // The reference compares stopwatch readings as doubles, `us * 1e-6 > period` etc.  Those
  // predicates are monotone in the integer microsecond reading, so the host evaluates them once
  // (timing_thresholds) and the kernels compare integers: same decisions, no FP64 per tick.
  uint32_t logic_min_age_us;  // smallest reading with  double(us)*1e-6 >  logic_period  (Quadcopter_T.cpp:159)
  uint32_t net_min_age_us;    // smallest reading with !(double(us)*1e-6 <  comm_period)  (UWBNetwork.cpp:28)
  uint32_t plant_min_age_us;  // smallest reading with !(double(us)*1e-6 <  1e-6)         (Quadcopter_T.cpp:88)
  // offboard loop (Simulator/Rappids_Simulator/main.cpp:471-476, CommunicationsDelay.hpp:18-39): clock-only as well
  int off_enabled;
  uint32_t off_min_age_us;    // smallest reading with  double(us)*1e-6 >  period               (main.cpp:471)
  uint32_t off_adj_us;        // uint64_t(period * 1e6) of Timer::AdjustTimeBySeconds            (main.cpp:476)
  uint32_t off_delay_us;      // CommunicationsDelay::_delayTime_us
  uint64_t off_first_target_us;  // no command is generated before the first target applies
  // simulated mocap packets for the offboard estimator (main.cpp:451-457): clock-only
  int mocap_enabled;
  uint32_t mocap_min_age_us, mocap_adj_us;
};

inline uint32_t first_true_us(double guess_us, bool (*pred)(double, double), double arg) {
  int64_t a = int64_t(guess_us) - 4;
  if (a < 0) a = 0;
  while (!pred(double(uint64_t(a)) * double(1e-6), arg)) a++;
  return uint32_t(a);
}
inline void timing_thresholds(TimingConsts& tc) {
  tc.logic_min_age_us = first_true_us(tc.logic_period * 1e6, [](double t, double p) { return t > p; }, tc.logic_period);
  tc.net_min_age_us = tc.net_enabled ? first_true_us(tc.comm_period * 1e6, [](double t, double p) { return !(t < p); }, tc.comm_period) : 0xFFFFFFFFu;
  tc.plant_min_age_us = first_true_us(1.0, [](double t, double p) { return !(t < p); }, 1e-6);
}
inline void timing_thresholds_mocap(TimingConsts& tc, double period) {
  tc.mocap_min_age_us = first_true_us(period * 1e6, [](double t, double p) { return t > p; }, period);
  tc.mocap_adj_us = uint32_t(uint64_t((-period) * double(-1e6)));
}
inline void timing_thresholds_offboard(TimingConsts& tc, double period) {
  tc.off_min_age_us = first_true_us(period * 1e6, [](double t, double p) { return t > p; }, period);
  tc.off_adj_us = uint32_t(uint64_t((-period) * double(-1e6)));  // Timer.hpp:31-33
}

// Stopwatch readings that are identical for every vehicle of a batch (they depend on the clock
// only): Quadcopter_T::_integrationTimer, ::_timerOnboardLogic, KalmanFilter6DOF::_estimateTimer,
// UWBNetwork::_timeSinceLastRange and the network's transaction flag.  Evolved by the same
// function on the host (to carry them across launches) and on the device (per tick).
struct Timing {
  uint32_t integ_age, logic_age, kf_age, net_age;
  uint32_t net_active, has_target;
  // offboard loop: main-loop stopwatch and the delay queue's control state (payloads are per vehicle)
  uint32_t off_age, off_head, off_count;
  uint32_t off_wait[AGF_OFFQ];  // microseconds until each queued command is due (0: deliverable)
  uint32_t mocap_age;           // timerMocap
  uint64_t now_us;              // simulation clock (ManualTimer::GetMicroSeconds)
};

// off_wait[] by a run-time slot without dynamic register-array indexing (which would put the whole struct in local memory)
AGF_HDI uint32_t off_wait_get(const Timing& ts, uint32_t slot) {
  uint32_t v = ts.off_wait[0];
#pragma unroll
  for (uint32_t q = 1; q < AGF_OFFQ; q++) v = (slot == q) ? ts.off_wait[q] : v;
  return v;
}
AGF_HDI void off_wait_set(Timing& ts, uint32_t slot, uint32_t value) {
#pragma unroll
  for (uint32_t q = 0; q < AGF_OFFQ; q++) ts.off_wait[q] = (slot == q) ? value : ts.off_wait[q];
}

struct TickPlan {
  uint32_t plant_dt_us, kf_dt_us;
  float plant_dt_f32, kf_dt_f32;  // the same intervals in seconds as the FP32 code needs them, converted once per launch on the host
  bool run_plant, run_logic, run_net, net_start, net_complete, net_reset, has_target;
  bool off_deliver, off_generate, mocap_update;
  uint32_t off_deliver_slot, off_gen_slot;
};

// dt_us: the clock advance that follows this tick's Run() (the offboard loop acts after the advance)
AGF_HDI TickPlan timing_plan(const Timing& ts, const TimingConsts& tc, uint32_t dt_us) {
  TickPlan p;
  p.plant_dt_us = ts.integ_age;
  p.kf_dt_us = ts.kf_age;
  p.kf_dt_f32 = float(ts.kf_age) * 1e-6f;
  p.plant_dt_f32 = float(double(ts.integ_age) * 1e-6);
  // Quadcopter_T.cpp:87-90: dt = GetSeconds<double>(); if (dt < 1e-6) return;
  p.run_plant = ts.integ_age >= tc.plant_min_age_us;
  // Quadcopter_T.cpp:159: if (_timerOnboardLogic.GetSeconds<double>() > _onboardLogicPeriod)
  p.run_logic = p.run_plant && ts.logic_age >= tc.logic_min_age_us;
  p.has_target = ts.has_target || (p.run_logic && tc.n_anchors > 0);
  // UWBNetwork.cpp:28: if (_timeSinceLastRange.GetSeconds<double>() < _commPeriod) return;
  p.run_net = tc.net_enabled && ts.net_age >= tc.net_min_age_us;
  p.net_start = p.net_complete = p.net_reset = false;
  if (p.run_net) {
    if (!ts.net_active) {
      p.net_start = p.has_target;  // :32-41 a requester with a target exists
      p.net_reset = true;          // :43 the stopwatch restarts whether or not one was found
    } else {
      p.net_complete = true;
    }
  }
  // offboard loop.  Delivery (main.cpp:737-739 of the previous iteration): the oldest queued command, once due.
  p.off_deliver = tc.off_enabled && ts.off_count > 0 && off_wait_get(ts, ts.off_head % AGF_OFFQ) == 0;
  p.off_deliver_slot = ts.off_head % AGF_OFFQ;
  // Generation (main.cpp:471): after Run() and the clock advance, when the stopwatch exceeds the period
  p.off_generate = tc.off_enabled && (ts.off_age + dt_us >= tc.off_min_age_us) && (ts.now_us + dt_us >= tc.off_first_target_us);
  p.off_gen_slot = (ts.off_head + ts.off_count) % AGF_OFFQ;  // the delivered one (if any) frees the head, not the tail
  // mocap packet (main.cpp:451): after Run() and the clock advance, before the offboard main loop
  p.mocap_update = tc.mocap_enabled && (ts.mocap_age + dt_us >= tc.mocap_min_age_us);
  return p;
}

AGF_HDI void timing_advance(Timing& ts, const TimingConsts& tc, const TickPlan& p, uint32_t dt_us) {
  if (p.run_plant) ts.integ_age = 0;
  if (p.run_logic) {
    ts.logic_age -= tc.logic_adj_us;
    ts.kf_age = 0;
  }
  ts.has_target = p.has_target;
  if (p.net_start) ts.net_active = 1;
  if (p.net_complete) ts.net_active = 0;
  if (p.net_reset) ts.net_age = 0;
  ts.integ_age += dt_us;
  ts.logic_age += dt_us;
  ts.kf_age += dt_us;
  ts.net_age += dt_us;
  ts.now_us += dt_us;
  if (tc.mocap_enabled) {
    ts.mocap_age += dt_us;
    if (ts.mocap_age >= tc.mocap_min_age_us) ts.mocap_age -= tc.mocap_adj_us;
  }
  if (tc.off_enabled) {
    if (p.off_deliver) {
      ts.off_head = (ts.off_head + 1) % AGF_OFFQ;
      ts.off_count--;
    }
    for (uint32_t q = 0; q < AGF_OFFQ; q++) ts.off_wait[q] = ts.off_wait[q] > dt_us ? ts.off_wait[q] - dt_us : 0;
    ts.off_age += dt_us;
    if (ts.off_age >= tc.off_min_age_us) {  // stopwatch restarts whether or not a target applies yet (main.cpp:476)
      ts.off_age -= tc.off_adj_us;
      if (p.off_generate) {
        off_wait_set(ts, p.off_gen_slot, tc.off_delay_us);
        ts.off_count++;
      }
    }
  }
}

// A launch's tick plans, evaluated ONCE on the host (they depend on the clock only) and read by the kernel as one
// 16-byte word per tick instead of evolving the stopwatch recurrence in every thread: that state cost the hot loop
// ~10 registers and, in the kernels with the offboard loop, was spilled and re-read every tick (ncu, round 1: 30 % of
// the stall samples of the offboard kernels were local-memory loads of it).
enum { PP_RUN_PLANT = 1u, PP_RUN_LOGIC = 2u, PP_NET_START = 4u, PP_NET_COMPLETE = 8u, PP_OFF_DELIVER = 16u, PP_OFF_GENERATE = 32u,
       PP_MOCAP = 64u, PP_DSLOT_SHIFT = 8, PP_GSLOT_SHIFT = 12 };
struct PackedPlan {
  uint32_t flags, plant_dt_us;
  float kf_dt_f32;     // float(kf_dt_us) * 1e-6f          (KalmanFilter6DOF's dt, QuadcopterLogic.cpp:229)
  float plant_dt_f32;  // float(double(plant_dt_us) * 1e-6) (Quadcopter_T.cpp:87 in the FP32 plant)
};
inline PackedPlan pack_plan(const TickPlan& p) {
  PackedPlan q;
  q.flags = (p.run_plant ? PP_RUN_PLANT : 0u) | (p.run_logic ? PP_RUN_LOGIC : 0u) | (p.net_start ? PP_NET_START : 0u) |
            (p.net_complete ? PP_NET_COMPLETE : 0u) | (p.off_deliver ? PP_OFF_DELIVER : 0u) | (p.off_generate ? PP_OFF_GENERATE : 0u) |
            (p.mocap_update ? PP_MOCAP : 0u) | (p.off_deliver_slot << PP_DSLOT_SHIFT) | (p.off_gen_slot << PP_GSLOT_SHIFT);
  q.plant_dt_us = p.plant_dt_us;
  q.kf_dt_f32 = float(p.kf_dt_us) * 1e-6f;
  q.plant_dt_f32 = float(double(p.plant_dt_us) * 1e-6);
  return q;
}
AGF_HDI TickPlan unpack_plan(uint32_t flags, uint32_t plant_dt_us, float kf_dt_f32, float plant_dt_f32) {
  TickPlan p;
  p.plant_dt_us = plant_dt_us;
  p.kf_dt_us = 0;  // host-side bookkeeping only
  p.kf_dt_f32 = kf_dt_f32;
  p.plant_dt_f32 = plant_dt_f32;
  p.run_plant = (flags & PP_RUN_PLANT) != 0;
  p.run_logic = (flags & PP_RUN_LOGIC) != 0;
  p.net_start = (flags & PP_NET_START) != 0;
  p.net_complete = (flags & PP_NET_COMPLETE) != 0;
  p.off_deliver = (flags & PP_OFF_DELIVER) != 0;
  p.off_generate = (flags & PP_OFF_GENERATE) != 0;
  p.mocap_update = (flags & PP_MOCAP) != 0;
  p.off_deliver_slot = (flags >> PP_DSLOT_SHIFT) & 0xFu;
  p.off_gen_slot = (flags >> PP_GSLOT_SHIFT) & 0xFu;
  p.run_net = p.net_reset = p.has_target = false;  // host-side bookkeeping only
  return p;
}

struct AnchorDev {
  float x, y, z;
  uint32_t id;
};

// Offboard::MocapStateEstimator of the in-kernel offboard loop (agrifly_b200.h "offboard loop: state estimator")
// state fields: position, velocity, angular velocity, attitude, 2x2 variances (row-major), estimate time [us], time of
// the last accepted measurement, initialised flag, rejection counters, prediction pipe (message count, then
// AGF_OFFEST_PIPE message slots of time-active, acceleration, angular velocity, ballistic flag).
// Storage is blocked by warp, [vehicle / 32][field][vehicle % 32]: field k of vehicle i sits at
// state[est_index(i) + k * E_LANES], so every access of the device code is base register + immediate (no index
// arithmetic) and a warp's access is 256 contiguous bytes.
// The pipe is an unordered set of slots: a free slot holds time-active = E_SLOT_FREE.  Activation times are strictly
// increasing in the order messages are added (PredictionPipe.hpp:25-30 stamps them with the clock), so "the newest
// message that is already active" and "the oldest one that is not yet" (GetActiveMessage, :32-53) are a maximum and a
// minimum over the slots, and ClearExpiredMessages (:55-68) frees every active message but the newest: no message ever moves.
enum { E_POS = 0, E_VEL = 3, E_W = 6, E_ATT = 9, E_VP = 13, E_VA = 17, E_TEST = 21, E_LASTGOOD = 22, E_INIT = 23, E_NREJ = 24,
       E_NREJC = 25, E_NPIPE = 26, E_PIPE = 27, E_MSG = 8, E_FIELDS = E_PIPE + E_MSG * AGF_OFFEST_PIPE, E_LANES = 32 };
#define E_SLOT_FREE 1e300
AGF_HDI size_t est_index(size_t i) { return (i / E_LANES) * (size_t(E_FIELDS) * E_LANES) + (i % E_LANES); }
AGF_HDI size_t est_doubles(size_t n) { return ((n + E_LANES - 1) / E_LANES) * (size_t(E_FIELDS) * E_LANES); }
struct EstParams {
  int kind;
  uint64_t t0_us;  // clock reading at construction: origin of the estimator's and its pipe's Timer
  double delay, reject, tc_angvel, inv_tc_angvel, meas_pos, meas_att, proc_pos, proc_att;
  double* state;   // device, est_doubles(N) values, see est_index
};
// Offboard::QuadcopterController + radio link of the in-kernel offboard loop (agrifly_b200.h "offboard rates loop")
struct OffboardParams {
  float nat_freq, damping, tc_att_xy, tc_att_z;
  float k3_att, k12_att;   // reciprocals, fast variants
  double max_proper, min_vert, min_proper;
  float yaw;
  uint32_t flags;
  uint32_t n_targets;
  const agf_offboard_target* targets;  // device, sorted by time
  const double* offsets;               // device [3][N] or null
  // reference generator (agf_offboard_ref)
  int ref_kind, traj_id;
  uint64_t start_us, stop_us;
  double desired[3], desired_yaw;
  int safety_net;  // Offboard::SafetyNet on the estimate (stages)
  double safe_min[3], safe_max[3], min_normal_height, not_seen_timeout;
  double* state;       // device [AGF_OFFSTATE_DOUBLES][N]: stage machine state per vehicle
  const double* traj;  // device [AGF_OFFTRAJ_DOUBLES][N]: motion primitive per vehicle
  EstParams est;
};

template<typename P>
struct StepShared {
  LogicParams logic;
  OffboardParams off;
  TimingConsts tc;
  P motor_min, motor_max, motor_J;
  P motor_pos[4][3];
  P drag[3];
  int has_drag;
  const P* ext_force;   // [3][N] world frame, or null
  const P* ext_torque;  // [3][N]
  AnchorDev anchors[AGF_MAX_UWB_ANCHORS];
  uint32_t n_anchors;
  uint64_t seed;
  uint32_t philox_rk[20];  // Philox4x32-10 round keys of `seed` (agf_step.cuh philox_round_keys)
  int noise_on, bias_on, uwb_noise_on;
  float sigma_gyro, sigma_acc, bias_sigma_gyro, bias_sigma_acc, uwb_sigma;
  float uwb_outlier_prob, uwb_outlier_sigma;  // UWBNetwork::SetNoiseProperties (UWBNetwork.hpp:28-33)
};

// plant parameters that may vary per vehicle (parameter sweeps)
template<typename P>
struct PlantPV {
  P mass;
  P I[9], Iinv[9];
  P kF, kTau;
  P motor_c;  // exp(-dt/tau) of Motor.cpp:53-57 for the dt of this launch, evaluated on the host
  P inv_mass; // 1/mass, used by the fast-arithmetic variants only
};

// the same for a per-vehicle parameter sweep: diagonal inertia, everything in registers
template<typename P>
struct PlantPVDiag {
  P mass, inv_mass;
  P Id[3], Iinvd[3];
  P kF, kTau;
  P motor_c;
};

#if defined(__CUDACC__)
// per-vehicle plant parameters in HBM: 3 quads of P-lanes... stored as plain component arrays
// grouped in 16-byte quads like the state: {mass, ixx, iyy, izz}, {iinv_xx, iinv_yy, iinv_zz, kF},
// {kTau, motor_c, 1/mass, 0}  -> NPV_PAD scalars
enum { PV_MASS = 0, PV_IXX = 1, PV_IYY = 2, PV_IZZ = 3, PV_IIXX = 4, PV_IIYY = 5, PV_IIZZ = 6, PV_KF = 7,
       PV_KTAU = 8, PV_MOTOR_C = 9, PV_INV_MASS = 10, NPV_PAD = 12 };

struct SchedEntryDev {
  uint64_t tick;
  int32_t slot;
  uint32_t type, flags;
  float f[4];
  uint32_t pad_;
};

struct SlotArrays {
  const float4* f;      // [N] floats[0..3] of each vehicle's packet
  const uint32_t* tf;   // [N] type | flags << 8
};

template<typename P>
struct StepLaunch {
  StepShared<P> sh;
  StateArrays<P> st;
  const typename VecOf<P>::type* pv;  // per-vehicle plant parameters [NPV_PAD/lanes][N], or null
  PlantPV<P> pv_shared;
  size_t n;
  uint64_t tick0;
  uint32_t nticks, dt_us;
  const uint4* plans;  // [nticks] PackedPlan of every tick of this launch
  uint64_t now0_us;    // simulation clock at the first tick
  const SchedEntryDev* sched;
  uint32_t sched_begin, sched_end;
  SlotArrays slots[AGF_MAX_CMD_SLOTS];
  // trajectory log ring or null: [capacity] records; a record is (AGF_LOG_FIELDS - 1) / lanes vectors [quad][log_n] holding the
  // first 16 values of every vehicle, then the 17th value as [log_n] scalars; log_n = N rounded up to whole 128-byte lines
  P* log;
  size_t log_n;
  uint32_t log_stride, log_capacity;
  uint32_t log_first_off, log_slot0;  // tick offset (within the launch) and ring slot of the launch's first record
  uint64_t first_global_index;
  // Work distribution (agf_step.cuh "balanced schedule"): the population is cut into nblocks vehicle blocks of
  // blockDim.x vehicles.  balanced == 0: CTA b steps block b for all nticks.  balanced == 1: the grid is one
  // wave of co-resident CTAs and the nblocks x nticks block-ticks are dealt out evenly, a block's ticks being
  // split between two neighbouring CTAs where a boundary falls inside it; flags[b] == epoch publishes that the
  // first part of block b has been stored.
  uint32_t nblocks, balanced;
  uint32_t* flags;
  uint32_t epoch;
};
#endif

}  // namespace agf
