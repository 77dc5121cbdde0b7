// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE (builds oracle/_ref/libagf_ref_*.so).
//
// Drives the UNMODIFIED reference implementation of the hot path
//   Components/Components/Simulation/{Quadcopter_T,Motor,UWBNetwork}.cpp
//   Components/Components/Logic/{QuadcopterLogic,KalmanFilter6DOF}.cpp
// (compiled where they lie under /root/reference by oracle/Makefile, against the header shims in
// oracle/shim) through the C interface of oracle/oracle_api.h.  No reference source is copied
// into this repository; this file only *calls* the reference's public classes and, for parity
// dumps and noise control, reads/writes private members via the access-specifier macro below
// (class layout is unaffected by access specifiers with GCC).
//
// The reference has no noise-free switch: IMU noise sigmas are constructor-initialised private
// doubles (Quadcopter_T.cpp:5-6,25-26); the harness overwrites them.
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <fstream>
#include <iostream>
#include <limits>
#include <memory>
#include <queue>
#include <random>
#include <thread>
#include <sstream>
#include <string>
#include <vector>

#include <Eigen/Dense>

#define private public
#define protected public
#include "Common/Time/ManualTimer.hpp"
#include "Components/Simulation/Quadcopter_T.hpp"
#include "Components/Simulation/UWBNetwork.hpp"
#include "Components/Simulation/CommunicationsDelay.hpp"
#include "Components/Offboard/QuadcopterController.hpp"
#include "Components/Offboard/MocapStateEstimator.hpp"
#include "Components/Offboard/SafetyNet.hpp"
#include "Components/TrajectoryGenerator/RapidTrajectoryGenerator.hpp"
#include "ExampleVehicleStateMachine.hpp"  // AIFS_ROS/hiperlab_rostools/src/QuadMocapRatesControl, against oracle/shim/ros
#undef private
#undef protected

#include "hiperlab_rostools/telemetry.h"
#include "oracle_api.h"

#ifndef ORC_FLAVOUR
#define ORC_FLAVOUR "ref-glibc"
#endif

struct orc_vehicle {
  ManualTimer timer;
  std::shared_ptr<Simulation::Quadcopter> quad;
  std::unique_ptr<Simulation::UWBNetwork> net;
  std::vector<std::shared_ptr<Simulation::UWBRadio>> anchors;
  uint64_t tick;
  // offboard loop (orc_run_offboard): created on first use
  std::unique_ptr<Timer> offTimer;
  std::unique_ptr<Simulation::CommunicationsDelay<RadioTypes::RadioMessageDecoded::RawMessage>> offChannel;
  // offboard state estimator (orc_set_offboard_estimator): main.cpp:221-225,286
  std::unique_ptr<Offboard::MocapStateEstimator> est;
  std::unique_ptr<Timer> timerMocap;
  double periodMocap = 0, delayEst = 0;
  // after Run() + clock advance: simulated mocap packet (main.cpp:451-457)
  void mocapStep() {
    if (!est) return;
    if (timerMocap->GetSeconds<double>() > periodMocap) {
      timerMocap->AdjustTimeBySeconds(-periodMocap);
      est->UpdateWithMeasurement(quad->GetPosition(), quad->GetAttitude());
    }
  }
  Offboard::EstimatedState estimate() {  // main.cpp:468-469
    if (est) return est->GetPrediction(delayEst);
    Offboard::EstimatedState e;
    e.pos = quad->GetPosition();
    e.vel = quad->GetVelocity();
    e.att = quad->GetAttitude();
    e.angVel = quad->GetAngularVelocity();
    return e;
  }
  // the ROS rates-control node's own state machine object (orc_run_stages_node)
  std::unique_ptr<Offboard::ExampleVehicleStateMachine> node;
  std::unique_ptr<Timer> nodeTimerMocap;
  // reference generators of the offboard loop (orc_run_offboard_ref)
  int stage = AGF_STAGE_WAIT_FOR_START, lastStage = AGF_STAGE_COMPLETE;  // ExampleVehicleStateMachine.cpp:10-11
  std::unique_ptr<Timer> stageTimer;
  Vec3d initPosition = Vec3d(0, 0, 0), lastPos = Vec3d(0, 0, 0), lastVel = Vec3d(0, 0, 0), lastAcc = Vec3d(0, 0, 0);  // ExampleVehicleStateMachine.cpp:16-25
  double cmdYawAngle = 0;                         // ExampleVehicleStateMachine.cpp:19
};

template<typename LPF>
static void dump_lpf3(const LPF& f, float out[4][3]) {
  const Vec3f* s[4] = {&f._xm0, &f._xm1, &f._ym0, &f._ym1};
  for (int i = 0; i < 4; i++) {
    out[i][0] = s[i]->x; out[i][1] = s[i]->y; out[i][2] = s[i]->z;
  }
}

template<typename Real>
static std::string toCSV(const Vec3<Real> v) {
  std::stringstream ss;
  ss << v.x << "," << v.y << "," << v.z << ",";
  return ss.str();
}

extern "C" {

const char* orc_flavour(void) { return ORC_FLAVOUR; }

orc_vehicle* orc_create(const agf_vehicle_cfg* cfg, const orc_opts* opts) {
  orc_vehicle* v = new orc_vehicle();
  v->tick = 0;
  Eigen::Matrix<double, 3, 3> inertia;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      inertia(i, j) = cfg->inertia[3 * i + j];
  v->quad.reset(new Simulation::Quadcopter(
      &v->timer, cfg->mass, inertia, cfg->arm_length,
      Vec3d(cfg->com_error[0], cfg->com_error[1], cfg->com_error[2]), cfg->motor_min_speed,
      cfg->motor_max_speed, cfg->prop_thrust_from_speed_sqr, cfg->prop_torque_from_speed_sqr,
      cfg->motor_time_const, cfg->motor_inertia,
      Vec3d(cfg->lin_drag_coeff_b[0], cfg->lin_drag_coeff_b[1], cfg->lin_drag_coeff_b[2]),
      uint8_t(cfg->vehicle_id),
      Onboard::QuadcopterConstants::QuadcopterType(cfg->quad_type), opts->onboard_logic_period));
  v->quad->_stdDevAccNoise = opts->sigma_acc;
  v->quad->_stdDevRateGyroNoise = opts->sigma_gyro;
  if (opts->uwb_comm_period > 0) {
    v->net.reset(new Simulation::UWBNetwork(&v->timer, opts->uwb_comm_period));
    v->net->SetNoiseProperties(opts->uwb_noise_std_dev, opts->uwb_outlier_probability, opts->uwb_outlier_std_dev);
    v->net->AddRadio(v->quad->GetRadio());
  }
  return v;
}

void orc_destroy(orc_vehicle* v) { delete v; }

void orc_set_state(orc_vehicle* v, const double pos[3], const double vel[3], const double att[4],
                   const double ang_vel[3]) {
  v->quad->SetPosition(Vec3d(pos[0], pos[1], pos[2]));
  v->quad->SetVelocity(Vec3d(vel[0], vel[1], vel[2]));
  v->quad->SetAttitude(Rotationd(att[0], att[1], att[2], att[3]));
  v->quad->SetAngularVelocity(Vec3d(ang_vel[0], ang_vel[1], ang_vel[2]));
}

void orc_set_external(orc_vehicle* v, const double force[3], const double torque[3]) {
  if (force) v->quad->SetExternalForce(Vec3d(force[0], force[1], force[2]));
  if (torque) v->quad->SetExternalTorque(Vec3d(torque[0], torque[1], torque[2]));
}

int orc_add_anchor(orc_vehicle* v, uint8_t id, float x, float y, float z) {
  if (v->quad->_logic._numRangingTargets >= 32) return -1;
  v->quad->AddUWBRadioTarget(id, Vec3f(x, y, z));
  if (v->net) {
    std::shared_ptr<Simulation::UWBRadio> r(new Simulation::UWBRadio(&v->timer, id));
    r->SetPosition(Vec3d(Vec3f(x, y, z)));
    v->anchors.push_back(r);
    v->net->AddRadio(r);
  }
  return 0;
}

void orc_set_radio(orc_vehicle* v, const uint8_t raw[AGF_RADIO_PACKET_SIZE]) {
  RadioTypes::RadioMessageDecoded::RawMessage m;
  memcpy(m.raw, raw, AGF_RADIO_PACKET_SIZE);
  v->quad->SetCommandRadioMsg(m);
}

static void record(orc_vehicle* v, double* r) {
  Simulation::Quadcopter& q = *v->quad;
  Vec3d p = q.GetPosition(), vel = q.GetVelocity(), w = q.GetAngularVelocity();
  Rotationd a = q.GetAttitude();
  r[0] = p.x; r[1] = p.y; r[2] = p.z;
  r[3] = vel.x; r[4] = vel.y; r[5] = vel.z;
  for (int i = 0; i < 4; i++) r[6 + i] = a[i];
  r[10] = w.x; r[11] = w.y; r[12] = w.z;
  for (int i = 0; i < 4; i++) r[13 + i] = q._motors[i]._speed;
  for (int i = 0; i < 4; i++) r[17 + i] = q._motorSpeedCommands[i];
  Vec3f ep, ev, ew;
  Rotationf ea;
  q.GetEstimate(ep, ev, ea, ew);
  r[21] = ep.x; r[22] = ep.y; r[23] = ep.z;
  r[24] = ev.x; r[25] = ev.y; r[26] = ev.z;
  for (int i = 0; i < 4; i++) r[27 + i] = ea[i];
  r[31] = ew.x; r[32] = ew.y; r[33] = ew.z;
  r[34] = int(q._logic._state);
  r[35] = q._logic._firstPanicReason;
  r[36] = q._logic._cycleCounter;
  r[37] = q._logic._kf._numResets;
  r[38] = q._logic._kf._numMeasRejected;
  r[39] = q._logic._uwbRangeMeas.count;
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
    v->quad->Run();
    if (v->net) v->net->Run();
    if (traj) record(v, traj + size_t(k) * ORC_NTRAJ);
    v->timer.AdvanceMicroSeconds(dt_us);
    v->tick++;
  }
}

// Loop body of Simulator/Rappids_Simulator/main.cpp:391-392,471-476,625-627,666-673,737-739 with the reference's own
// Timer, QuadcopterController, RadioTypes and CommunicationsDelay; estimate = truth.
void orc_run_offboard(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_offboard_cfg* cfg,
                      const agf_offboard_target* targets, uint32_t n_targets, const double* offset, double* traj) {
  if (!v->offTimer) {
    v->offTimer.reset(new Timer(&v->timer));
    v->offChannel.reset(new Simulation::CommunicationsDelay<RadioTypes::RadioMessageDecoded::RawMessage>(
        &v->timer, double(cfg->delay_us) * 1e-6));
    v->offChannel->_delayTime_us = cfg->delay_us;  // exact integer microseconds
  }
  Offboard::QuadcopterController ctrl;
  ctrl.SetParameters(cfg->pos_control_nat_freq, cfg->pos_control_damping, cfg->att_control_time_const_xy,
                     cfg->att_control_time_const_z);
  ctrl._minVerticalProperAcceleration = cfg->min_vertical_proper_acc;
  ctrl._maxProperAcc = cfg->max_proper_acc;
  ctrl._minProperAcc = cfg->min_proper_acc;
  const double period = double(cfg->period_us) * 1e-6;
  for (uint32_t k = 0; k < nticks; k++) {
    // main.cpp:737-739 of the previous iteration == before this Run()
    if (v->offChannel->HaveNewMessage()) v->quad->SetCommandRadioMsg(v->offChannel->GetMessage());
    v->quad->Run();
    if (v->net) v->net->Run();
    if (traj) record(v, traj + size_t(k) * ORC_NTRAJ);
    v->timer.AdvanceMicroSeconds(dt_us);
    v->tick++;
    v->mocapStep();
    if (v->offTimer->GetSeconds<double>() > period) {  // main.cpp:471
      v->offTimer->AdjustTimeBySeconds(-period);       // main.cpp:476
      const uint64_t now = v->timer.GetMicroSeconds();
      int ti = -1;
      for (uint32_t j = 0; j < n_targets; j++)
        if (targets[j].time_us <= now) ti = int(j);
      if (ti < 0) continue;
      Vec3d des(targets[ti].pos[0], targets[ti].pos[1], targets[ti].pos[2]);
      if (offset) des = des + Vec3d(offset[0], offset[1], offset[2]);
      Vec3d cmdAngVel;
      double cmdThrust;
      const Offboard::EstimatedState estState = v->estimate();
      ctrl.Run(estState.pos, estState.vel, estState.att, des, Vec3d(0, 0, 0),
               Vec3d(0, 0, 0), cfg->yaw_angle, cmdAngVel, cmdThrust);  // main.cpp:625-627
      RadioTypes::RadioMessageDecoded::RawMessage rawMsg;
      memset(rawMsg.raw, 0, sizeof(rawMsg.raw));
      RadioTypes::RadioMessageDecoded::CreateRatesCommand(uint8_t(cfg->radio_flags), float(cmdThrust), Vec3f(cmdAngVel),
                                                          rawMsg.raw);  // main.cpp:666-669
      if (v->est)  // main.cpp:652-654
        v->est->SetPredictedValues(cmdAngVel, (estState.att * Vec3d(0, 0, 1) * cmdThrust - Vec3d(0, 0, 9.81)));
      v->offChannel->AddMessage(rawMsg);                                 // main.cpp:673
    }
  }
}

// The same loop with the desired state coming from a reference generator (include/agrifly_b200.h "offboard loop:
// reference generators"): the stage logic of ExampleVehicleStateMachine::Run restated line by line around the
// reference's own Timer / QuadcopterController / RadioTypes (the node itself needs ROS), and Rappids_Simulator's
// trajectory tracking (main.cpp:560-634) around the reference's RapidTrajectoryGenerator and RunTracking.
void orc_run_offboard_ref(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_offboard_cfg* cfg,
                          const agf_offboard_ref* ref, const double* offset, const double* tr /* [AGF_OFFTRAJ_DOUBLES] */,
                          double* traj) {
  if (!v->offTimer) {
    v->offTimer.reset(new Timer(&v->timer));
    v->offChannel.reset(new Simulation::CommunicationsDelay<RadioTypes::RadioMessageDecoded::RawMessage>(
        &v->timer, double(cfg->delay_us) * 1e-6));
    v->offChannel->_delayTime_us = cfg->delay_us;
  }
  if (!v->stageTimer) v->stageTimer.reset(new Timer(&v->timer));
  Offboard::QuadcopterController ctrl;
  ctrl.SetParameters(cfg->pos_control_nat_freq, cfg->pos_control_damping, cfg->att_control_time_const_xy,
                     cfg->att_control_time_const_z);
  ctrl._minVerticalProperAcceleration = cfg->min_vertical_proper_acc;
  ctrl._maxProperAcc = cfg->max_proper_acc;
  ctrl._minProperAcc = cfg->min_proper_acc;
  const double period = double(cfg->period_us) * 1e-6;
  Vec3d _desiredPosition(ref->desired_pos[0], ref->desired_pos[1], ref->desired_pos[2]);
  if (offset) _desiredPosition = _desiredPosition + Vec3d(offset[0], offset[1], offset[2]);
  const double _desiredYawAngle = ref->desired_yaw;
  std::unique_ptr<RapidQuadrocopterTrajectoryGenerator::RapidTrajectoryGenerator> _traj;
  Rotationd trajAtt = Rotationd::Identity();
  Vec3d trajOffset(0, 0, 0);
  double trajEndTime = 0;
  if (ref->kind == AGF_OFFREF_TRAJECTORY) {
    _traj.reset(new RapidQuadrocopterTrajectoryGenerator::RapidTrajectoryGenerator(
        Vec3d(tr[0], tr[6], tr[12]), Vec3d(tr[1], tr[7], tr[13]), Vec3d(tr[2], tr[8], tr[14]), Vec3d(tr[18], tr[19], tr[20])));
    for (int a = 0; a < 3; a++) {
      _traj->_axis[a]._a = tr[6 * a + 3];
      _traj->_axis[a]._b = tr[6 * a + 4];
      _traj->_axis[a]._g = tr[6 * a + 5];
    }
    trajEndTime = tr[21];
    trajAtt = Rotationd(tr[22], tr[23], tr[24], tr[25]);
    trajOffset = Vec3d(tr[26], tr[27], tr[28]);
  }
  for (uint32_t k = 0; k < nticks; k++) {
    if (v->offChannel->HaveNewMessage()) v->quad->SetCommandRadioMsg(v->offChannel->GetMessage());
    v->quad->Run();
    if (v->net) v->net->Run();
    if (traj) record(v, traj + size_t(k) * ORC_NTRAJ);
    v->timer.AdvanceMicroSeconds(dt_us);
    v->tick++;
    v->mocapStep();
    if (!(v->offTimer->GetSeconds<double>() > period)) continue;  // main.cpp:471
    v->offTimer->AdjustTimeBySeconds(-period);                     // main.cpp:476
    const uint64_t now = v->timer.GetMicroSeconds();
    const Offboard::EstimatedState estState = v->estimate();
    const Vec3d estPos = estState.pos, estVel = estState.vel;
    const Rotationd estAtt = estState.att;
    // 0: nothing, 1: SetPredictedValues(0, 0), 2: SetPredictedValues(cmdAngVel, att * e3 * cmdThrust - g)
    int predicted = 2;
    RadioTypes::RadioMessageDecoded::RawMessage rawMsg;
    memset(rawMsg.raw, 0, sizeof(rawMsg.raw));
    bool send = true;
    Vec3d cmdAngVel;
    double cmdThrust;
    if (ref->kind == AGF_OFFREF_TRAJECTORY) {
      if (!(now > ref->start_us)) {  // main.cpp:623-627 (t < startFlightTime; the clocks are integer microseconds)
        ctrl.Run(estPos, estVel, estAtt, _desiredPosition, Vec3d(0, 0, 0), Vec3d(0, 0, 0), _desiredYawAngle, cmdAngVel,
                 cmdThrust);
      } else {
        double traj_t = double(now - ref->start_us) * 1e-6;  // trackTrajTime.GetSeconds<double>() (main.cpp:560)
        Vec3d trajPos, trajVel, trajAcc;
        if (traj_t < trajEndTime) {  // main.cpp:562-571
          traj_t += 0.04;
          trajPos = _traj->GetPosition(traj_t);
          trajVel = _traj->GetVelocity(traj_t);
          trajAcc = _traj->GetAcceleration(traj_t);
        } else {
          trajPos = _traj->GetPosition(trajEndTime);
          trajVel = Vec3d(0, 0, 0);
          trajAcc = Vec3d(0, 0, 0);
        }
        if (trajPos.z < 0) {  // main.cpp:580-591
          trajPos.z = 0;
          if (trajVel.z < 0) trajVel.z = 0;
          if (trajAcc.z < 0) trajAcc.z = 0;
        }
        Vec3d refPos = trajAtt * trajPos + trajOffset;  // main.cpp:593-600
        Vec3d refVel = trajAtt * trajVel;
        Vec3d refAcc = trajAtt * trajAcc;
        double refThrust = _traj->GetThrust(traj_t);
        Vec3d refAngVel = estAtt.Inverse() * trajAtt * _traj->GetOmega(traj_t, 0.02);
        Rotationf cmdAtt;
        ctrl.RunTracking(estPos, estVel, estAtt, refPos, refVel, refAcc, _desiredYawAngle, refThrust, refAngVel,
                         cmdAngVel, cmdThrust, cmdAtt);  // main.cpp:629-634
      }
      RadioTypes::RadioMessageDecoded::CreateRatesCommand(uint8_t(cfg->radio_flags), float(cmdThrust), Vec3f(cmdAngVel),
                                                          rawMsg.raw);
    } else {  // AGF_OFFREF_STAGES: ExampleVehicleStateMachine::Run(shouldStart, shouldStop)
      const bool shouldStart = now >= ref->start_us, shouldStop = now >= ref->stop_us;
      bool safe = true;  // _safetyNet->UpdateWithEstimator(...) :104-105, the reference's own SafetyNet with the configured box
      if (ref->safety_net) {
        Offboard::SafetyNet net;
        net.SetSafeCorners(Vec3d(ref->safe_min[0], ref->safe_min[1], ref->safe_min[2]),
                           Vec3d(ref->safe_max[0], ref->safe_max[1], ref->safe_max[2]), ref->min_normal_height);
        net._vehicleNotSeenTimeout = ref->not_seen_timeout;
        net.UpdateWithEstimator(estState, v->est ? v->est->GetTimeSinceLastGoodMeasurement() : 0.0);
        safe = net.GetIsSafe();
      }
      bool stageChange = v->stage != v->lastStage;  // :96-100
      v->lastStage = v->stage;
      if (stageChange) v->stageTimer->Reset();
      auto runController = [&](Vec3d desPos, Vec3d desVel, Vec3d desAcc) {  // RunControllerAndUpdateEstimator :407-432
        ctrl.Run(estPos, estVel, estAtt, desPos, desVel, desAcc, v->cmdYawAngle, cmdAngVel, cmdThrust);
        RadioTypes::RadioMessageDecoded::CreateRatesCommand(uint8_t(cfg->radio_flags), float(cmdThrust), Vec3f(cmdAngVel),
                                                            rawMsg.raw);
      };
      switch (v->stage) {
        case AGF_STAGE_WAIT_FOR_START:  // :113-120
          if (shouldStart) v->stage = AGF_STAGE_SPOOL_UP;
          send = false;
          predicted = 0;
          break;
        case AGF_STAGE_SPOOL_UP: {  // :122-160
          if (!safe) v->stage = AGF_STAGE_EMERGENCY;
          double const motorSpoolUpTime = 0.5;
          double const spoolUpThrustByWeight = 0.25;
          predicted = 1;  // :134
          cmdThrust = 9.81 * spoolUpThrustByWeight;
          cmdAngVel = Vec3d(0, 0, 0);
          RadioTypes::RadioMessageDecoded::CreateRatesCommand(uint8_t(cfg->radio_flags), float(cmdThrust), Vec3f(cmdAngVel),
                                                              rawMsg.raw);
          if (v->stageTimer->GetSeconds<double>() > motorSpoolUpTime) v->stage = AGF_STAGE_TAKEOFF;
        } break;
        case AGF_STAGE_TAKEOFF: {  // :162-189
          if (stageChange) v->initPosition = estPos;
          if (!safe) v->stage = AGF_STAGE_EMERGENCY;
          double const takeOffTime = 2.0;
          double frac = v->stageTimer->GetSeconds<double>() / takeOffTime;
          if (frac >= 1.0) {
            v->stage = AGF_STAGE_FLIGHT;
            frac = 1.0;
          }
          Vec3d cmdPos = (1 - frac) * v->initPosition + frac * _desiredPosition;
          runController(cmdPos, Vec3d(0, 0, 0), Vec3d(0, 0, 0));
        } break;
        case AGF_STAGE_FLIGHT: {  // :191-298
          if (!safe) v->stage = AGF_STAGE_EMERGENCY;
          Vec3d cmdPos(0, 0, 0), cmdVel(0, 0, 0), cmdAcc(0, 0, 0);
          double t = v->stageTimer->GetSeconds<double>();
          double const getIntoActionTime = 2.0;
          double frac = std::min(t / getIntoActionTime, 1.0);
          switch (ref->traj_id) {
            case 0:
              cmdPos = _desiredPosition;
              cmdVel = Vec3d(0, 0, 0);
              cmdAcc = Vec3d(0, 0, 0);
              v->cmdYawAngle = 0;
              break;
            case 1: {
              Vec3d circleCenter(0.0, -2.0, _desiredPosition.z);
              double radius = 1.0;
              double angSpeed = 0.5;
              cmdPos = circleCenter + radius * Vec3d(cos(angSpeed * t), sin(angSpeed * t), 0);
              cmdVel = radius * angSpeed * Vec3d(-sin(angSpeed * t), cos(angSpeed * t), 0);
              cmdAcc = radius * pow(angSpeed, 2) * Vec3d(-cos(angSpeed * t), -sin(angSpeed * t), 0);
              v->cmdYawAngle = _desiredYawAngle + angSpeed * t;
            } break;
            case 2: {
              double amplitude = 1.0;
              double angFreq = 2.0;
              cmdPos = _desiredPosition + amplitude * Vec3d(0, sin(angFreq * t), 0);
              cmdVel = amplitude * angFreq * Vec3d(0, cos(angFreq * t), 0);
              cmdAcc = amplitude * pow(angFreq, 2) * Vec3d(0, -sin(angFreq * t), 0);
              v->cmdYawAngle = _desiredYawAngle;
            } break;
            case 3: {
              Vec3d circleCenter(0.0, 0.0, _desiredPosition.z);
              double radius = 0.5;
              double angSpeed = 1;
              cmdPos = circleCenter + radius * Vec3d(cos(angSpeed * t), sin(angSpeed * t), 0);
              cmdVel = radius * angSpeed * Vec3d(-sin(angSpeed * t), cos(angSpeed * t), 0);
              cmdAcc = radius * pow(angSpeed, 2) * Vec3d(-cos(angSpeed * t), -sin(angSpeed * t), 0);
              v->cmdYawAngle = 0;
            } break;
            case 4: {
              Vec3d circleCenter(0.0, 0.0, _desiredPosition.z);
              double radius = 0.5;
              double angSpeed = 0.5;
              cmdPos = circleCenter + radius * Vec3d(cos(angSpeed * t), sin(angSpeed * t), cos(angSpeed * t * 4));
              cmdVel = radius * angSpeed * Vec3d(-sin(angSpeed * t), cos(angSpeed * t), -sin(angSpeed * t * 4));
              cmdAcc = radius * pow(angSpeed, 2) * Vec3d(-cos(angSpeed * t), -sin(angSpeed * t), -cos(angSpeed * t * 4));
              v->cmdYawAngle = angSpeed * t;
            } break;
            case 5:
              cmdPos = _desiredPosition;
              v->cmdYawAngle = 0.2 * t;
              break;
          }
          v->lastPos = (1 - frac) * _desiredPosition + frac * cmdPos;
          v->lastVel = frac * cmdVel;
          v->lastAcc = frac * cmdAcc;
          runController(v->lastPos, v->lastVel, v->lastAcc);
          if (shouldStop) v->stage = AGF_STAGE_LANDING;
        } break;
        case AGF_STAGE_LANDING: {  // :300-324
          if (!safe) v->stage = AGF_STAGE_EMERGENCY;
          double const LANDING_SPEED = 0.5;
          double const getIntoActionTime = 2.0;
          double frac = std::min(v->stageTimer->GetSeconds<double>() / getIntoActionTime, 1.0);
          Vec3d cmdPos = v->lastPos + v->stageTimer->GetSeconds<double>() * Vec3d(0, 0, -LANDING_SPEED);
          if (cmdPos.z < 0) v->stage = AGF_STAGE_COMPLETE;
          runController((1 - frac) * v->lastPos + frac * cmdPos, (1 - frac) * v->lastVel + frac * Vec3d(0, 0, -LANDING_SPEED),
                        (1 - frac) * v->lastAcc + frac * Vec3d(0, 0, 0));
        } break;
        case AGF_STAGE_COMPLETE:  // :326-343
          predicted = 1;  // :332
          RadioTypes::RadioMessageDecoded::CreateIdleCommand(uint8_t(cfg->radio_flags), rawMsg.raw);
          break;
        default:  // AGF_STAGE_EMERGENCY :350-363
          predicted = 0;
          RadioTypes::RadioMessageDecoded::CreateKillCommand(uint8_t(cfg->radio_flags), rawMsg.raw);
          break;
      }
    }
    if (v->est && predicted == 1) v->est->SetPredictedValues(Vec3d(0, 0, 0), Vec3d(0, 0, 0));
    if (v->est && predicted == 2)
      v->est->SetPredictedValues(cmdAngVel, (estState.att * Vec3d(0, 0, 1) * cmdThrust - Vec3d(0, 0, 9.81)));
    if (send) v->offChannel->AddMessage(rawMsg);
  }
}

// The UNMODIFIED flight-stage state machine of the ROS rates-control node (ExampleVehicleStateMachine.cpp) in the loop: its
// own MocapStateEstimator fed through CallbackEstimator at the mocap rate, Run(shouldStart, shouldStop) at the offboard
// period, the radio_command it publishes taken from the roscpp shim and sent through the delay queue.  This pins the
// restatement of the stage logic in orc_run_offboard_ref / the port / the device code.  A message whose type byte is 0 (the
// default-constructed message the node publishes while it waits for the start signal) is not transmitted.
void orc_run_stages_node(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_offboard_cfg* cfg, const agf_offboard_ref* ref,
                         const agf_offboard_estimator* e, int traj_id_check, double* traj) {
  using hiperlab_rostools::radio_command;
  (void)traj_id_check;
  if (!v->offTimer) {
    v->offTimer.reset(new Timer(&v->timer));
    v->offChannel.reset(new Simulation::CommunicationsDelay<RadioTypes::RadioMessageDecoded::RawMessage>(
        &v->timer, double(cfg->delay_us) * 1e-6));
    v->offChannel->_delayTime_us = cfg->delay_us;
  }
  if (!v->node) {
    std::cout.setstate(std::ios_base::failbit);  // the node narrates its stages on stdout
    v->node.reset(new Offboard::ExampleVehicleStateMachine());
    ros::NodeHandle n;
    v->node->Initialize(1, "oracle", n, &v->timer, e->prediction_delay);  // id 1: the airframe of the tests (GetVehicleTypeFromID)
    std::cout.clear();
    v->node->_ctrl->SetParameters(cfg->pos_control_nat_freq, cfg->pos_control_damping, cfg->att_control_time_const_xy,
                                  cfg->att_control_time_const_z);
    v->node->_ctrl->_minVerticalProperAcceleration = cfg->min_vertical_proper_acc;
    v->node->_ctrl->_maxProperAcc = cfg->max_proper_acc;
    v->node->_ctrl->_minProperAcc = cfg->min_proper_acc;
    v->node->_est->SetStatistics(e->meas_noise_pos, e->meas_noise_att, e->proc_noise_pos, e->proc_noise_att);
    v->node->_est->SetAngularVelocityTimeConstant(e->angvel_time_const);
    v->node->_est->_measRejectDist = e->meas_reject_dist;
    v->node->SetDesiredPosition(Vec3d(ref->desired_pos[0], ref->desired_pos[1], ref->desired_pos[2]));
    v->node->SetDesiredYaw(ref->desired_yaw);
    if (ref->safety_net) {
      v->node->_safetyNet->SetSafeCorners(Vec3d(ref->safe_min[0], ref->safe_min[1], ref->safe_min[2]),
                                          Vec3d(ref->safe_max[0], ref->safe_max[1], ref->safe_max[2]), ref->min_normal_height);
      v->node->_safetyNet->_vehicleNotSeenTimeout = ref->not_seen_timeout;
    } else {  // switched off: a box nothing leaves, no time-out
      v->node->_safetyNet->SetSafeCorners(Vec3d(-1e30, -1e30, -1e30), Vec3d(1e30, 1e30, 1e30), -1e30);
      v->node->_safetyNet->_vehicleNotSeenTimeout = 1e30;
    }
    v->nodeTimerMocap.reset(new Timer(&v->timer));
  }
  const double period = double(cfg->period_us) * 1e-6, periodMocap = double(e->mocap_period_us) * 1e-6;
  std::cout.setstate(std::ios_base::failbit);
  for (uint32_t k = 0; k < nticks; k++) {
    if (v->offChannel->HaveNewMessage()) v->quad->SetCommandRadioMsg(v->offChannel->GetMessage());
    v->quad->Run();
    if (v->net) v->net->Run();
    if (traj) record(v, traj + size_t(k) * ORC_NTRAJ);
    v->timer.AdvanceMicroSeconds(dt_us);
    v->tick++;
    if (v->nodeTimerMocap->GetSeconds<double>() > periodMocap) {  // the mocap node's packet
      v->nodeTimerMocap->AdjustTimeBySeconds(-periodMocap);
      hiperlab_rostools::mocap_output m;
      const Vec3d p = v->quad->GetPosition();
      const Rotationd a = v->quad->GetAttitude();
      m.vehicleID = 1;
      m.posx = p.x; m.posy = p.y; m.posz = p.z;
      m.attq0 = a[0]; m.attq1 = a[1]; m.attq2 = a[2]; m.attq3 = a[3];
      v->node->CallbackEstimator(m);
    }
    if (!(v->offTimer->GetSeconds<double>() > period)) continue;
    v->offTimer->AdjustTimeBySeconds(-period);
    const uint64_t now = v->timer.GetMicroSeconds();
    ros::LastPublished<radio_command>::fresh() = false;
    v->node->Run(now >= ref->start_us, now >= ref->stop_us);
    if (!ros::LastPublished<radio_command>::fresh()) continue;
    const radio_command& c = ros::LastPublished<radio_command>::get();
    if (c.raw[0] == 0) continue;  // RadioTypes::invalid: the wait stage's default-constructed message
    RadioTypes::RadioMessageDecoded::RawMessage rawMsg;
    memcpy(rawMsg.raw, c.raw, sizeof(rawMsg.raw));
    v->offChannel->AddMessage(rawMsg);
  }
  std::cout.clear();
}

void orc_get_stages_node_state(orc_vehicle* v, double* o) {
  for (int i = 0; i < AGF_OFFSTATE_DOUBLES; i++) o[i] = 0.0;
  if (!v->node) return;
  Offboard::ExampleVehicleStateMachine& m = *v->node;
  o[0] = int(m._flightStage);
  o[1] = int(m._lastFlightStage);
  o[2] = double(m._stageTimer->_lastResetTime_usec);
  const Vec3d* q[4] = {&m._initPosition, &m._lastPos, &m._lastVel, &m._lastAcc};
  for (int i = 0; i < 4; i++) {
    o[3 + 3 * i] = q[i]->x; o[4 + 3 * i] = q[i]->y; o[5 + 3 * i] = q[i]->z;
  }
  o[15] = m._cmdYawAngle;
}

void orc_set_offboard_estimator(orc_vehicle* v, const agf_offboard_estimator* e) {
  if (!e || e->kind != AGF_OFFEST_MOCAP) {
    v->est.reset();
    v->timerMocap.reset();
    return;
  }
  v->est.reset(new Offboard::MocapStateEstimator(&v->timer, 1, e->prediction_delay));  // main.cpp:221-224
  v->est->SetStatistics(e->meas_noise_pos, e->meas_noise_att, e->proc_noise_pos, e->proc_noise_att);
  v->est->SetAngularVelocityTimeConstant(e->angvel_time_const);
  v->est->_measRejectDist = e->meas_reject_dist;
  v->timerMocap.reset(new Timer(&v->timer));  // main.cpp:286
  v->periodMocap = double(e->mocap_period_us) * 1e-6;
  v->delayEst = e->prediction_delay;
}

void orc_get_offboard_estimate(orc_vehicle* v, double horizon, double* o13, double* c4) {
  Offboard::EstimatedState e = v->est ? v->est->GetPrediction(horizon) : v->estimate();
  o13[0] = e.pos.x; o13[1] = e.pos.y; o13[2] = e.pos.z;
  o13[3] = e.vel.x; o13[4] = e.vel.y; o13[5] = e.vel.z;
  for (int i = 0; i < 4; i++) o13[6 + i] = e.att[i];
  o13[10] = e.angVel.x; o13[11] = e.angVel.y; o13[12] = e.angVel.z;
  if (c4) {
    c4[0] = v->est ? double(v->est->_initialized) : 0.0;
    c4[1] = v->est ? double(v->est->_numMeasRejected) : 0.0;
    c4[2] = v->est ? double(v->est->_numMeasRejectedConsecutively) : 0.0;
    c4[3] = v->est ? double(v->est->_predictionPipe._messages.size()) : 0.0;
  }
}

void orc_get_offboard_state(orc_vehicle* v, double* o) {
  o[0] = v->stage;
  o[1] = v->lastStage;
  o[2] = v->stageTimer ? double(v->stageTimer->_lastResetTime_usec) : 0.0;
  const Vec3d* q[4] = {&v->initPosition, &v->lastPos, &v->lastVel, &v->lastAcc};
  for (int i = 0; i < 4; i++) {
    o[3 + 3 * i] = q[i]->x; o[4 + 3 * i] = q[i]->y; o[5 + 3 * i] = q[i]->z;
  }
  o[15] = v->cmdYawAngle;
}

// One row of Rappids_Simulator's simulation.csv written with the reference's own types and stream operators
// (main.cpp:55-59 toCSV, :676-733), from the same record the product formats.
size_t orc_csv_row(const agf_csv_record* r, char* buf, size_t cap) {
  std::ostringstream logfile;
  logfile << r->t << ",";
  logfile << toCSV(Vec3d(r->pos[0], r->pos[1], r->pos[2]));
  logfile << toCSV(Vec3d(r->vel[0], r->vel[1], r->vel[2]));
  logfile << toCSV(Rotationd(r->att[0], r->att[1], r->att[2], r->att[3]).ToEulerYPR());
  logfile << toCSV(Vec3d(r->ang_vel[0], r->ang_vel[1], r->ang_vel[2]));
  for (int i = 0; i < 4; i++) logfile << r->motor_forces[i] << ",";
  Vec3f pos(r->est_pos[0], r->est_pos[1], r->est_pos[2]), vel(r->est_vel[0], r->est_vel[1], r->est_vel[2]);
  Rotationf att(r->est_att[0], r->est_att[1], r->est_att[2], r->est_att[3]);
  Vec3f angVel(r->est_ang_vel[0], r->est_ang_vel[1], r->est_ang_vel[2]);
  logfile << toCSV(pos);
  logfile << toCSV(vel);
  logfile << toCSV(att.ToEulerYPR());
  logfile << toCSV(angVel);
  logfile << toCSV(Vec3d(r->des_pos[0], r->des_pos[1], r->des_pos[2]));
  logfile << toCSV(Vec3d(r->des_vel[0], r->des_vel[1], r->des_vel[2]));
  logfile << int(r->panic_reason) << ",";
  for (int i = 0; i < 4; i++) logfile << double(r->last_radio_cmd[i]) << ",";
  logfile << "\n";
  const std::string s = logfile.str();
  if (buf && cap) {
    strncpy(buf, s.c_str(), cap - 1);
    buf[cap - 1] = 0;
  }
  return s.size();
}

// The telemetry and simulator_truth messages as the ROS Simulator node fills them (AIFS_ROS/hiperlab_rostools/src/Simulator/
// main.cpp:455-475, 501-546; that file itself needs AirSim, Boost and cv_bridge): the same statements on the reference's own
// TelemetryPacket / Rotation types and the shim's message structs.
void orc_msg_telemetry(const uint8_t* raw1, const uint8_t* raw2, agf_msg_telemetry* o) {
  TelemetryPacket::data_packet_t dataPacketRaw1, dataPacketRaw2;
  memcpy(&dataPacketRaw1, raw1, AGF_TELEMETRY_PACKET_SIZE);
  memcpy(&dataPacketRaw2, raw2, AGF_TELEMETRY_PACKET_SIZE);
  TelemetryPacket::TelemetryPacket dataPacket1, dataPacket2;
  TelemetryPacket::DecodeTelemetryPacket(dataPacketRaw1, dataPacket1);
  TelemetryPacket::DecodeTelemetryPacket(dataPacketRaw2, dataPacket2);
  hiperlab_rostools::telemetry telMsgOut;
  telMsgOut.packetNumber = dataPacket1.packetNumber;
  for (int i = 0; i < 3; i++) {
    telMsgOut.accelerometer[i] = dataPacket1.accel[i];
    telMsgOut.rateGyro[i] = dataPacket1.gyro[i];
    telMsgOut.position[i] = dataPacket1.position[i];
  }
  for (int i = 0; i < 4; i++) telMsgOut.motorForces[i] = dataPacket1.motorForces[i];
  telMsgOut.batteryVoltage = dataPacket1.battVoltage;
  for (int i = 0; i < TelemetryPacket::TelemetryPacket::NUM_DEBUG_FLOATS; i++) telMsgOut.debugVals[i] = dataPacket2.debugVals[i];
  Vec3f attYPR = Rotationf::FromVectorPartOfQuaternion(
      Vec3f(dataPacket2.attitude[0], dataPacket2.attitude[1], dataPacket2.attitude[2])).ToEulerYPR();
  for (int i = 0; i < 3; i++) {
    telMsgOut.velocity[i] = dataPacket2.velocity[i];
    telMsgOut.attitude[i] = dataPacket2.attitude[i];
    telMsgOut.attitudeYPR[i] = attYPR[i];
  }
  telMsgOut.panicReason = dataPacket2.panicReason;
  telMsgOut.warnings = dataPacket2.warnings;
  memset(o, 0, sizeof(*o));
  o->packetNumber = telMsgOut.packetNumber;
  for (int i = 0; i < 3; i++) {
    o->accelerometer[i] = telMsgOut.accelerometer[i]; o->rateGyro[i] = telMsgOut.rateGyro[i]; o->position[i] = telMsgOut.position[i];
    o->attitude[i] = telMsgOut.attitude[i]; o->velocity[i] = telMsgOut.velocity[i]; o->attitudeYPR[i] = telMsgOut.attitudeYPR[i];
  }
  for (int i = 0; i < 4; i++) o->motorForces[i] = telMsgOut.motorForces[i];
  for (int i = 0; i < 6; i++) o->debugVals[i] = telMsgOut.debugVals[i];
  o->batteryVoltage = telMsgOut.batteryVoltage;
  o->panicReason = telMsgOut.panicReason;
  o->warnings = telMsgOut.warnings;
}
void orc_msg_simulator_truth(orc_vehicle* v, agf_msg_simulator_truth* o) {
  memset(o, 0, sizeof(*o));
  o->posx = v->quad->GetPosition().x; o->posy = v->quad->GetPosition().y; o->posz = v->quad->GetPosition().z;
  o->velx = v->quad->GetVelocity().x; o->vely = v->quad->GetVelocity().y; o->velz = v->quad->GetVelocity().z;
  o->attq0 = v->quad->GetAttitude()[0]; o->attq1 = v->quad->GetAttitude()[1];
  o->attq2 = v->quad->GetAttitude()[2]; o->attq3 = v->quad->GetAttitude()[3];
  v->quad->GetAttitude().ToEulerYPR(o->attyaw, o->attpitch, o->attroll);
  o->angvelx = v->quad->GetAngularVelocity().x; o->angvely = v->quad->GetAngularVelocity().y; o->angvelz = v->quad->GetAngularVelocity().z;
}

void orc_get_full(orc_vehicle* v, orc_full_state* o) {
  memset(o, 0, sizeof(*o));
  Simulation::Quadcopter& q = *v->quad;
  Onboard::QuadcopterLogic& L = q._logic;
  Vec3d p = q.GetPosition(), vel = q.GetVelocity(), w = q.GetAngularVelocity();
  Rotationd a = q.GetAttitude();
  o->pos[0] = p.x; o->pos[1] = p.y; o->pos[2] = p.z;
  o->vel[0] = vel.x; o->vel[1] = vel.y; o->vel[2] = vel.z;
  for (int i = 0; i < 4; i++) o->att[i] = a[i];
  o->ang_vel[0] = w.x; o->ang_vel[1] = w.y; o->ang_vel[2] = w.z;
  for (int i = 0; i < 4; i++) {
    o->motor_speed[i] = q._motors[i]._speed;
    o->motor_force_z[i] = q.GetMotorForce(i);
    o->motor_speed_cmd[i] = q._motorSpeedCommands[i];
    o->des_motor_speeds[i] = L._desMotorSpeeds[i];
    o->des_motor_forces[i] = L._desMotorForcesForTelemetry[i];
  }
  o->flight_state = int(L._state);
  o->first_panic_reason = L._firstPanicReason;
  o->cycle_counter = int(L._cycleCounter);
  o->tel_warnings = L._telWarnings;
  dump_lpf3(L._imuRateGyro.lowPass, o->gyro_lpf);
  dump_lpf3(L._imuAccelerometer.lowPass, o->acc_lpf);
  o->temp_lpf[0] = L._imuTemperature.lowPass._xm0; o->temp_lpf[1] = L._imuTemperature.lowPass._xm1;
  o->temp_lpf[2] = L._imuTemperature.lowPass._ym0; o->temp_lpf[3] = L._imuTemperature.lowPass._ym1;
  o->batt_lpf[0] = L._battVoltageLowPass._xm0; o->batt_lpf[1] = L._battVoltageLowPass._xm1;
  o->batt_lpf[2] = L._battVoltageLowPass._ym0; o->batt_lpf[3] = L._battVoltageLowPass._ym1;
  o->batt_voltage_filtered = L._battMeas.voltageFiltered;
  o->monitor_cmd_rate_lpdt = L._monitorCmdRate.lpDt;
  o->monitor_main_loop_lpdt = L._monitorMainLoopPeriod.lpDt;
  o->des_pos[0] = L._desPos.x; o->des_pos[1] = L._desPos.y; o->des_pos[2] = L._desPos.z;
  o->radio_type = L._radioMessage.msg.type;
  o->radio_flags = L._radioMessage.msg.flags;
  o->radio_count = L._radioMessage.count;
  if (L._radioMessage.count)  // floats are uninitialised before the first message (RadioTypes.hpp:118-121)
    for (int i = 0; i < 10; i++) o->radio_floats[i] = L._radioMessage.msg.floats[i];
  o->uwb_meas_count = L._uwbRangeMeas.count;
  o->next_ranging_target_idx = L._nextRangingTargetIdx;
  Onboard::KalmanFilter6DOF& kf = L._kf;
  o->kf_pos[0] = kf._pos.x; o->kf_pos[1] = kf._pos.y; o->kf_pos[2] = kf._pos.z;
  o->kf_vel[0] = kf._vel.x; o->kf_vel[1] = kf._vel.y; o->kf_vel[2] = kf._vel.z;
  for (int i = 0; i < 4; i++) o->kf_att[i] = kf._att[i];
  o->kf_ang_vel[0] = kf._angVel.x; o->kf_ang_vel[1] = kf._angVel.y; o->kf_ang_vel[2] = kf._angVel.z;
  o->kf_last_corr[0] = kf._lastMeasUpdateAttCorrection.x;
  o->kf_last_corr[1] = kf._lastMeasUpdateAttCorrection.y;
  o->kf_last_corr[2] = kf._lastMeasUpdateAttCorrection.z;
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 9; j++)
      o->kf_cov[9 * i + j] = kf._cov(i, j);
  o->kf_imu_init = kf._IMUInitialized;
  o->kf_uwb_init = kf._UWBInitialized;
  o->kf_num_resets = kf._numResets;
  o->kf_num_rejected = kf._numMeasRejected;
  o->kf_num_rejected_seq = kf._numMeasRejectedSequentially;
  for (int i = 0; i < 6; i++) o->debug[i] = L._debug[i];
}

void orc_get_telemetry(orc_vehicle* v, uint8_t p1[30], uint8_t p2[30]) {
  TelemetryPacket::data_packet_t a, b;
  memset(&a, 0, sizeof(a));
  memset(&b, 0, sizeof(b));
  v->quad->GetTelemetryDataPackets(a, b);
  static_assert(sizeof(a) == 30, "packet size");
  memcpy(p1, &a, 30);
  memcpy(p2, &b, 30);
}

void orc_get_imu(orc_vehicle* v, double acc[3], double gyro[3]) {
  Vec3d a, g;
  v->quad->GetAccelerometer(a);
  v->quad->GetRateGyro(g);
  acc[0] = a.x; acc[1] = a.y; acc[2] = a.z;
  gyro[0] = g.x; gyro[1] = g.y; gyro[2] = g.z;
}

uint64_t orc_time_us(orc_vehicle* v) { return v->timer.GetMicroSeconds(); }

// The reference's UWB range noise comes from ONE file-scope generator (UWBNetwork.cpp:4, thread_local here: ref_tls_rng.h)
// that every UWBNetwork constructor re-seeds with 0 (:19).  Left alone, the vehicles of a population stepped on T threads
// would replay the same T-fold copies of one stream; a Monte-Carlo population needs independent realisations, so the
// generator is seeded per vehicle right before that vehicle runs (no reference source is touched; vehicle 0 keeps seed 0).
// The distribution objects next to it (:5-6) stay shared file-scope state: populations WITH range noise are stepped on one
// thread by the tests (threads = 1), which makes them reproducible.
}  // extern "C"
extern std::mt19937 rng;  // thread_local through the forced include (the -include of ref_tls_rng.h covers every file of the command)
extern std::normal_distribution<double> distNormal;  // UWBNetwork.cpp:6: keeps the second value of every pair it draws
extern "C" {
static void seed_range_noise(uint32_t vehicle) {
  rng.seed(vehicle);
  distNormal.reset();
}

double orc_run_population(const agf_vehicle_cfg* cfgs, uint32_t n_cfgs, uint32_t n,
                          const orc_opts* opts, const double* init13, const float* anchors,
                          uint32_t n_anchors, uint32_t dt_us, uint32_t nticks,
                          const agf_cmd_entry* sched, uint32_t nsched, const uint8_t* slot_raw,
                          uint32_t threads, double* final_out) {
  std::vector<orc_vehicle*> vs(n);
  for (uint32_t i = 0; i < n; i++) {
    vs[i] = orc_create(&cfgs[n_cfgs == 1 ? 0 : i], opts);
    // The reference default-seeds every vehicle's IMU noise engine identically (Quadcopter_T.hpp:122, seed 1); a
    // Monte-Carlo population needs independent streams: vehicle i is seeded i + 1 (vehicle 0 = the reference default).
    vs[i]->quad->_generator.seed(i + 1);
    if (init13) {
      const double* s = init13 + 13 * size_t(i);
      orc_set_state(vs[i], s, s + 3, s + 6, s + 10);
    }
    for (uint32_t a = 0; a < n_anchors; a++)
      orc_add_anchor(vs[i], uint8_t(anchors[4 * a]), anchors[4 * a + 1], anchors[4 * a + 2],
                     anchors[4 * a + 3]);
  }
  if (threads < 1) threads = 1;
  auto work = [&](uint32_t t) {
    uint32_t lo = uint32_t(uint64_t(n) * t / threads), hi = uint32_t(uint64_t(n) * (t + 1) / threads);
    uint8_t mine[AGF_MAX_CMD_SLOTS * AGF_RADIO_PACKET_SIZE];
    for (uint32_t i = lo; i < hi; i++) {
      const uint8_t* sr = nullptr;
      if (slot_raw) {
        for (int s = 0; s < AGF_MAX_CMD_SLOTS; s++)
          memcpy(mine + s * AGF_RADIO_PACKET_SIZE,
                 slot_raw + (size_t(s) * n + i) * AGF_RADIO_PACKET_SIZE, AGF_RADIO_PACKET_SIZE);
        sr = mine;
      }
      seed_range_noise(i);
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
    if (final_out) {
      // state after the last tick (time already advanced; record reads the stored state)
      record(vs[i], final_out + size_t(i) * ORC_NTRAJ);
    }
    orc_destroy(vs[i]);
  }
  return secs;
}

#define ORC_POP_SEED(v, i) ((v)->quad->_generator.seed((i) + 1), seed_range_noise(i))
#include "orc_population_traj.inc"

void orc_radio_encode_rates(uint8_t flags, float thrust, const float w[3], uint8_t raw[23]) {
  memset(raw, 0, 23);
  RadioTypes::RadioMessageDecoded::CreateRatesCommand(flags, thrust, Vec3f(w[0], w[1], w[2]), raw);
}
void orc_radio_encode_position(uint8_t flags, const float p[3], const float vv[3], const float a[3],
                               uint8_t raw[23]) {
  memset(raw, 0, 23);
  RadioTypes::RadioMessageDecoded::CreatePositionCommand(
      flags, Vec3f(p[0], p[1], p[2]), Vec3f(vv[0], vv[1], vv[2]), Vec3f(a[0], a[1], a[2]), raw);
}
void orc_radio_encode_acceleration(uint8_t flags, const float a[3], float yaw_rate,
                                   uint8_t raw[23]) {
  memset(raw, 0, 23);
  RadioTypes::RadioMessageDecoded::CreateAccelerationCommand(flags, Vec3f(a[0], a[1], a[2]),
                                                             yaw_rate, raw);
}
void orc_radio_decode(const uint8_t raw[23], uint8_t* type, uint8_t* flags, float floats[10]) {
  RadioTypes::RadioMessageDecoded m(raw);
  *type = m.type;
  *flags = m.flags;
  for (int i = 0; i < 10; i++) floats[i] = m.floats[i];
}
void orc_telemetry_decode(const uint8_t packet[30], agf_telemetry* out) {
  TelemetryPacket::data_packet_t p;
  memcpy(&p, packet, 30);
  TelemetryPacket::TelemetryPacket t;
  memset(&t, 0, sizeof(t));
  TelemetryPacket::DecodeTelemetryPacket(p, t);
  memset(out, 0, sizeof(*out));
  out->type = t.type;
  out->packet_number = t.packetNumber;
  for (int i = 0; i < 3; i++) {
    out->accel[i] = t.accel[i]; out->gyro[i] = t.gyro[i]; out->position[i] = t.position[i];
    out->velocity[i] = t.velocity[i]; out->attitude[i] = t.attitude[i];
  }
  for (int i = 0; i < 4; i++) out->motor_forces[i] = t.motorForces[i];
  for (int i = 0; i < 6; i++) out->debug_vals[i] = t.debugVals[i];
  out->batt_voltage = t.battVoltage;
  out->panic_reason = t.panicReason;
  out->warnings = t.warnings;
}

void orc_logic_consts(int quad_type, agf_logic_consts* o) {
  Onboard::QuadcopterConstants c((Onboard::QuadcopterConstants::QuadcopterType)quad_type);
  memset(o, 0, sizeof(*o));
  o->mass = c.mass; o->inertia_xx = c.inertia_xx; o->inertia_zz = c.inertia_zz;
  o->arm_length = c.armLength;
  o->prop_thrust_from_speed_sqr = c.propellerThrustFromSpeedSqr;
  o->prop_torque_from_thrust = c.propellerTorqueFromThrust;
  o->max_thrust_per_propeller = c.maxThrustPerPropeller;
  o->min_thrust_per_propeller = c.minThrustPerPropeller;
  o->max_cmd_total_thrust = c.maxCmdTotalThrust;
  o->prop0_spin_dir = c.prop0SpinDir;
  o->pos_control_nat_freq = c.posControl_natFreq; o->pos_control_damping = c.posControl_damping;
  o->ang_vel_control_time_const_xy = c.angVelControl_timeConst_xy;
  o->att_control_time_const_xy = c.attControl_timeConst_xy;
  o->ang_vel_control_time_const_z = c.angVelControl_timeConst_z;
  o->att_control_time_const_z = c.attControl_timeConst_z;
  o->imu_yaw = c.IMU_yaw; o->imu_pitch = c.IMU_pitch; o->imu_roll = c.IMU_roll;
  o->low_battery_threshold = c.lowBatteryThreshold;
  o->lin_drag_coeff_b[0] = c.linDragCoeffBx; o->lin_drag_coeff_b[1] = c.linDragCoeffBy;
  o->lin_drag_coeff_b[2] = c.linDragCoeffBz;
  o->motor_time_const = c.motorTimeConst; o->motor_inertia = c.motorInertia;
  o->motor_min_speed = c.motorMinSpeed; o->motor_max_speed = c.motorMaxSpeed;
  o->valid = c.valid;
}

// The vehicle configuration as the reference's apps build it (Simulator/Rappids_Simulator/main.cpp:147-165): vehicle ID ->
// QuadcopterConstants::GetVehicleTypeFromID -> the float table entries widened to double, inertia_yy = inertia_xx,
// propTorqueFromSpeedSqr = kTau * kF.  Lets bench.py's reference arm run without loading the product library.
int orc_vehicle_cfg_from_id(int vehicle_id, agf_vehicle_cfg* c) {
  const Onboard::QuadcopterConstants::QuadcopterType t = Onboard::QuadcopterConstants::GetVehicleTypeFromID(vehicle_id);
  memset(c, 0, sizeof(*c));
  orc_logic_consts(int(t), &c->logic);
  Onboard::QuadcopterConstants k(t);
  c->mass = k.mass;
  c->inertia[0] = k.inertia_xx;
  c->inertia[4] = k.inertia_xx;
  c->inertia[8] = k.inertia_zz;
  c->arm_length = k.armLength;
  c->prop_thrust_from_speed_sqr = k.propellerThrustFromSpeedSqr;
  c->prop_torque_from_speed_sqr = k.propellerTorqueFromThrust * k.propellerThrustFromSpeedSqr;
  c->motor_time_const = k.motorTimeConst;
  c->motor_inertia = k.motorInertia;
  c->motor_min_speed = k.motorMinSpeed;
  c->motor_max_speed = k.motorMaxSpeed;
  c->lin_drag_coeff_b[0] = k.linDragCoeffBx;
  c->lin_drag_coeff_b[1] = k.linDragCoeffBy;
  c->lin_drag_coeff_b[2] = k.linDragCoeffBz;
  c->vehicle_id = vehicle_id;
  c->quad_type = int(t);
  return k.valid ? 0 : -1;
}
void orc_radio_encode_idle(uint8_t flags, uint8_t raw[23]) {
  memset(raw, 0, 23);
  RadioTypes::RadioMessageDecoded::CreateIdleCommand(flags, raw);
}

}  // extern "C"
