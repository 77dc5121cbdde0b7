// oracle/hostsim/agf_hostsim.cu -- TEST INFRASTRUCTURE: the product's device step header
// (agri-fly_b200/csrc/agf_step.cuh) compiled for the HOST and driven through oracle/oracle_api.h.
//
// Purpose: separates "is the re-formulated step (sparse EKF algebra, packed flags, uniform timing,
// quad-packed state serialisation) logically identical to the reference?" -- answerable in the
// build container, which has no GPU, by comparing this library with oracle/_ref bit for bit --
// from "does the GPU execute the same arithmetic?" (the -m gpu parity tests).
// It is NOT a product path: nothing in agri-fly_b200/ loads it, and agf_batch_create refuses to
// run without a CUDA device.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <limits>
#include <vector>

#include "../oracle_api.h"
#include "../../agri-fly_b200/csrc/agf_host_params.h"
#include "../../agri-fly_b200/csrc/agf_step.cuh"

using namespace agf;

#ifndef ORC_FLAVOUR
#define ORC_FLAVOUR "hostsim-shared"
#endif

struct orc_vehicle {
  agf_vehicle_cfg cfg;
  orc_opts opts;
  StepShared<double> sh;
  PlantPV<double> pv;
  Timing ts;
  bool uwb;
  std::vector<double> hp, ext_f, ext_t;
  std::vector<float> hf, hc, hq;
  std::vector<uint32_t> hu;
  uint64_t tick, now_us;
  std::vector<double> offstate, offtraj;  // reference generators of the offboard loop (n = 1)
  std::vector<double> offest;             // offboard estimator state
  StateArrays<double> arrays() {
    StateArrays<double> a;
    a.sp = (double2*)hp.data();
    a.sf = (float4*)hf.data();
    a.su = (uint4*)hu.data();
    a.sc = uwb ? (float4*)hc.data() : nullptr;
    a.sq = hq.empty() ? nullptr : (float4*)hq.data();
    return a;
  }
};

template<bool UWB>
static void run_impl(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_cmd_entry* sched, uint32_t nsched,
                     const uint8_t* slot_raw, double* traj) {
  VState<double, true, UWB, true> s;
  StateArrays<double> a = v->arrays();
  Scratch sc;
  sc.q = nullptr;

  state_load(s, a, 1, 0, sc, v->sh.logic.mix_kf);
  v->sh.ext_force = v->ext_f.empty() ? nullptr : v->ext_f.data();
  v->sh.ext_torque = v->ext_t.empty() ? nullptr : v->ext_t.data();
  uint32_t si = 0;
  while (si < nsched && sched[si].tick < v->tick) si++;
  for (uint32_t k = 0; k < nticks; k++) {
    // the launch-constant motor coefficient for the dt the plant will integrate over
    const uint32_t plant_dt = v->ts.integ_age ? v->ts.integ_age : dt_us;
    v->pv.motor_c = motor_c_host(v->cfg.motor_time_const, plant_dt);
    if (si < nsched && sched[si].tick == v->tick) {
      const uint8_t* raw = sched[si].raw;
      if (sched[si].slot >= 0 && slot_raw) raw = slot_raw + AGF_RADIO_PACKET_SIZE * sched[si].slot;
      uint8_t type, flags;
      float f[10];
      agf_radio_decode(raw, &type, &flags, f);
      radio_deliver(s, v->sh.logic, type, flags, f);
      si++;
    }
    {
      const TickPlan plan = timing_plan(v->ts, v->sh.tc, dt_us);
      tick<double, true, UWB, true, true>(s, sc, v->sh, v->pv, plan, v->ts.now_us, dt_us, v->tick, 0, 0, 1);
      timing_advance(v->ts, v->sh.tc, plan, dt_us);
    }
    if (traj) {
      double* r = traj + size_t(k) * ORC_NTRAJ;
      for (int c = 0; c < 3; c++) { r[c] = s.pos[c]; r[3 + c] = s.vel[c]; r[10 + c] = s.w[c]; r[21 + c] = s.kpos[c]; r[24 + c] = s.kvel[c]; r[31 + c] = s.kw[c]; }
      for (int c = 0; c < 4; c++) { r[6 + c] = s.att[c]; r[13 + c] = s.ms[c]; r[17 + c] = s.cmd[c]; r[27 + c] = s.katt[c]; }
      r[34] = s.bits & 7u;
      r[35] = (s.bits >> 3) & 7u;
      r[36] = s.cycle;
      r[37] = s.kfcnt & 0xFFFFu;
      r[38] = s.kfcnt >> 16;
      r[39] = s.uwb_count;
    }
    v->tick++;
    v->now_us += dt_us;
  }
  state_store(s, a, 1, 0, sc, v->sh.logic.mix_kf);
}

extern "C" {

const char* orc_flavour(void) { return ORC_FLAVOUR; }

orc_vehicle* orc_create(const agf_vehicle_cfg* cfg, const orc_opts* opts) {
  orc_vehicle* v = new orc_vehicle();
  v->cfg = *cfg;
  v->opts = *opts;
  v->uwb = opts->uwb_comm_period > 0;
  build_shared(*cfg, opts->onboard_logic_period, opts->uwb_comm_period, v->sh);
  fill_plant(*cfg, v->pv);
  memset(&v->ts, 0, sizeof(v->ts));
  initial_state<double>(1, v->sh.logic, cfg->logic.low_battery_threshold, v->uwb, v->hp, v->hf, v->hu, v->hc);
  v->tick = 0;
  v->now_us = 0;
  return v;
}
void orc_destroy(orc_vehicle* v) { delete v; }

void orc_set_state(orc_vehicle* v, const double p[3], const double vel[3], const double a[4], const double w[3]) {
  for (int k = 0; k < 3; k++) {
    v->hp[sidx(SP_POS + k, 1, 0, 2)] = p[k];
    v->hp[sidx(SP_VEL + k, 1, 0, 2)] = vel[k];
    v->hp[sidx(SP_W + k, 1, 0, 2)] = w[k];
  }
  for (int k = 0; k < 4; k++) v->hp[sidx(SP_ATT + k, 1, 0, 2)] = a[k];
}
void orc_set_external(orc_vehicle* v, const double f[3], const double t[3]) {
  if (v->ext_f.empty()) { v->ext_f.assign(3, 0.0); v->ext_t.assign(3, 0.0); }
  if (f) for (int k = 0; k < 3; k++) v->ext_f[k] = f[k];
  if (t) for (int k = 0; k < 3; k++) v->ext_t[k] = t[k];
}
int orc_add_anchor(orc_vehicle* v, uint8_t id, float x, float y, float z) {
  if (v->sh.n_anchors >= AGF_MAX_UWB_ANCHORS) return -1;
  AnchorDev& a = v->sh.anchors[v->sh.n_anchors++];
  a.id = id; a.x = x; a.y = y; a.z = z;
  v->sh.tc.n_anchors = int(v->sh.n_anchors);
  return 0;
}
void orc_set_radio(orc_vehicle* v, const uint8_t raw[23]) {
  agf_cmd_entry e;
  memset(&e, 0, sizeof(e));
  e.tick = uint32_t(v->tick);
  e.slot = -1;
  memcpy(e.raw, raw, 23);
  // deliver through a zero-tick run: load, deliver, store
  if (v->uwb) {
    VState<double, true, true, true> s; StateArrays<double> a = v->arrays(); Scratch sc{nullptr}; state_load(s, a, 1, 0, sc, v->sh.logic.mix_kf);
    uint8_t ty, fl; float f[10]; agf_radio_decode(raw, &ty, &fl, f); radio_deliver(s, v->sh.logic, ty, fl, f); state_store(s, a, 1, 0, sc, v->sh.logic.mix_kf);
  } else {
    VState<double, true, false, true> s; StateArrays<double> a = v->arrays(); Scratch sc{nullptr}; state_load(s, a, 1, 0, sc, v->sh.logic.mix_kf);
    uint8_t ty, fl; float f[10]; agf_radio_decode(raw, &ty, &fl, f); radio_deliver(s, v->sh.logic, ty, fl, f); state_store(s, a, 1, 0, sc, v->sh.logic.mix_kf);
  }
}
void orc_run(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_cmd_entry* sched, uint32_t nsched,
             const uint8_t* slot_raw, double* traj) {
  if (v->uwb) run_impl<true>(v, dt_us, nticks, sched, nsched, slot_raw, traj);
  else run_impl<false>(v, dt_us, nticks, sched, nsched, slot_raw, traj);
}

// the device step's own offboard loop (tick(): plan.off_deliver / plan.off_generate), compiled for the host
void orc_run_offboard(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_offboard_cfg* cfg,
                      const agf_offboard_target* targets, uint32_t n_targets, const double* offset, double* traj) {
  if (v->hq.empty()) {
    v->hq.assign(4 * AGF_OFFQ, 0.0f);
    const char* why = fill_offboard(*cfg, v->sh.off, v->sh.tc);
    if (why) { fprintf(stderr, "hostsim: %s\n", why); abort(); }
  }
  v->sh.off.targets = targets;
  v->sh.off.n_targets = n_targets;
  v->sh.off.offsets = offset;  // [3][1]
  v->sh.tc.off_first_target_us = n_targets ? targets[0].time_us : ~0ull;
  v->ts.now_us = v->now_us;
  orc_run(v, dt_us, nticks, nullptr, 0, nullptr, traj);
}

void orc_run_offboard_ref(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_offboard_cfg* cfg,
                          const agf_offboard_ref* ref, const double* offset, const double* tr, double* traj) {
  if (v->hq.empty()) {
    v->hq.assign(4 * AGF_OFFQ, 0.0f);
    const char* why = fill_offboard(*cfg, v->sh.off, v->sh.tc);
    if (why) { fprintf(stderr, "hostsim: %s\n", why); abort(); }
  }
  if (v->offstate.empty()) {  // what agf_batch_set_offboard_reference writes
    v->offstate.assign(AGF_OFFSTATE_DOUBLES, 0.0);
    v->offstate[0] = AGF_STAGE_WAIT_FOR_START;
    v->offstate[1] = AGF_STAGE_COMPLETE;
    v->offstate[2] = double(v->now_us);
    v->offstate[15] = 0.0;
  }
  if (tr) v->offtraj.assign(tr, tr + AGF_OFFTRAJ_DOUBLES);
  OffboardParams& o = v->sh.off;
  o.targets = nullptr;
  o.n_targets = 0;
  o.offsets = offset;
  o.ref_kind = ref->kind;
  o.traj_id = ref->traj_id;
  o.start_us = ref->start_us;
  o.stop_us = ref->stop_us;
  for (int k = 0; k < 3; k++) o.desired[k] = ref->desired_pos[k];
  o.desired_yaw = ref->desired_yaw;
  o.safety_net = ref->kind == AGF_OFFREF_STAGES ? ref->safety_net : 0;
  for (int k = 0; k < 3; k++) { o.safe_min[k] = ref->safe_min[k]; o.safe_max[k] = ref->safe_max[k]; }
  o.min_normal_height = ref->min_normal_height;
  o.not_seen_timeout = ref->not_seen_timeout;
  o.state = v->offstate.data();
  o.traj = v->offtraj.empty() ? nullptr : v->offtraj.data();
  v->sh.tc.off_first_target_us = 0;
  v->ts.now_us = v->now_us;
  orc_run(v, dt_us, nticks, nullptr, 0, nullptr, traj);
}

// what agf_batch_set_offboard_estimator does (the offboard loop's parameters are filled at the first run call)
void orc_set_offboard_estimator(orc_vehicle* v, const agf_offboard_estimator* e) {
  EstParams& p = v->sh.off.est;
  if (!e || e->kind != AGF_OFFEST_MOCAP) {
    p.kind = AGF_OFFEST_TRUTH;
    v->sh.tc.mocap_enabled = 0;
    return;
  }
  v->offest.assign(est_doubles(1), 0.0);  // vehicle 0 of a warp-blocked array: field k at [k * E_LANES]
  v->offest[E_ATT * E_LANES] = 1.0;
  v->offest[E_VP * E_LANES] = 25.0; v->offest[(E_VP + 3) * E_LANES] = 25.0;
  v->offest[E_VA * E_LANES] = 1.0; v->offest[(E_VA + 3) * E_LANES] = 400.0;
  v->offest[E_LASTGOOD * E_LANES] = double(v->now_us);
  for (int k = 0; k < AGF_OFFEST_PIPE; k++) v->offest[(E_PIPE + E_MSG * k) * E_LANES] = E_SLOT_FREE;
  p.kind = AGF_OFFEST_MOCAP;
  p.t0_us = v->now_us;
  p.delay = e->prediction_delay;
  p.reject = e->meas_reject_dist;
  p.tc_angvel = e->angvel_time_const;
  p.inv_tc_angvel = 1.0 / e->angvel_time_const;
  p.meas_pos = e->meas_noise_pos; p.meas_att = e->meas_noise_att;
  p.proc_pos = e->proc_noise_pos; p.proc_att = e->proc_noise_att;
  p.state = v->offest.data();
  v->sh.tc.mocap_enabled = 1;
  timing_thresholds_mocap(v->sh.tc, double(e->mocap_period_us) * 1e-6);
  v->ts.mocap_age = 0;
}

void orc_get_offboard_estimate(orc_vehicle* v, double horizon, double* o, double* c4) {
  EstCore e;
  EstPipe pipe;
  mocap_predict<true, double>(v->sh.off.est, 0, v->now_us, horizon, e, pipe);
  const double x[13] = {e.pos.x, e.pos.y, e.pos.z, e.vel.x, e.vel.y, e.vel.z, e.att.w, e.att.x, e.att.y, e.att.z, e.w.x, e.w.y, e.w.z};
  for (int k = 0; k < 13; k++) o[k] = x[k];
  if (c4) {
    c4[0] = v->offest[E_INIT * E_LANES]; c4[1] = v->offest[E_NREJ * E_LANES]; c4[2] = v->offest[E_NREJC * E_LANES]; c4[3] = v->offest[E_NPIPE * E_LANES];
  }
}

void orc_get_offboard_state(orc_vehicle* v, double* out) {
  for (int k = 0; k < AGF_OFFSTATE_DOUBLES; k++) out[k] = v->offstate.empty() ? 0.0 : v->offstate[k];
}

void orc_get_full(orc_vehicle* v, orc_full_state* o) {
  memset(o, 0, sizeof(*o));
  auto P_ = [&](int k) { return v->hp[sidx(k, 1, 0, 2)]; };
  auto F_ = [&](int k) { return v->hf[sidx(k, 1, 0, 4)]; };
  auto U_ = [&](int k) { return v->hu[sidx(k, 1, 0, 4)]; };
  for (int k = 0; k < 3; k++) { o->pos[k] = P_(SP_POS + k); o->vel[k] = P_(SP_VEL + k); o->ang_vel[k] = P_(SP_W + k); }
  for (int k = 0; k < 4; k++) {
    o->att[k] = P_(SP_ATT + k);
    o->motor_speed[k] = P_(SP_MS + k);
    o->motor_force_z[k] = (v->pv.kF * o->motor_speed[k]) * fabs(o->motor_speed[k]);
    o->motor_speed_cmd[k] = F_(SF_CMD + k);
    o->des_motor_speeds[k] = F_(SF_CMD + k);
    o->des_motor_forces[k] = F_(SF_DFORCE + k);
    o->kf_att[k] = F_(SF_KATT + k);
    o->temp_lpf[k] = F_(SF_TEMP_LP + k);
    o->batt_lpf[k] = F_(SF_BATT_LP + k);
  }
  const uint32_t bits = U_(SU_BITS), cnt = U_(SU_CNT), kc = U_(SU_KFCNT), w = U_(SU_UWBW);
  o->flight_state = bits & 7u;
  o->first_panic_reason = (bits >> 3) & 7u;
  o->cycle_counter = int(U_(SU_CYCLE));
  o->tel_warnings = (cnt >> 24) & 0xFFu;
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < 4; k++) { o->gyro_lpf[k][c] = F_(SF_GYRO_LP + 4 * c + k); o->acc_lpf[k][c] = F_(SF_ACC_LP + 4 * c + k); }
  o->batt_voltage_filtered = F_(SF_BATT_VFILT);
  o->monitor_cmd_rate_lpdt = F_(SF_MON_CMD);
  o->monitor_main_loop_lpdt = F_(SF_MON_LOOP);
  o->radio_type = (bits >> 6) & 7u;
  o->radio_flags = (bits >> 9) & 0xFFu;
  for (int k = 0; k < 4; k++) o->radio_floats[k] = F_(SF_RADIO + k);
  o->uwb_meas_count = int(U_(SU_UWB_COUNT));
  o->next_ranging_target_idx = (w >> 24) & 0xFFu;
  for (int k = 0; k < 3; k++) { o->kf_pos[k] = F_(SF_KPOS + k); o->kf_vel[k] = F_(SF_KVEL + k); o->kf_ang_vel[k] = F_(SF_KW + k); o->kf_last_corr[k] = F_(SF_KCORR + k); }
  if (v->uwb) for (int k = 0; k < 81; k++) o->kf_cov[k] = v->hc[sidx(k, 1, 0, 4)];
  o->kf_imu_init = (bits >> 18) & 1u;
  o->kf_uwb_init = (bits >> 19) & 1u;
  o->kf_num_resets = kc & 0xFFFFu;
  o->kf_num_rejected = kc >> 16;
  o->kf_num_rejected_seq = cnt & 0xFFu;
  o->debug[0] = o->cycle_counter ? F_(SF_TEMP_LP + 3) : 0.0f;
}

uint64_t orc_time_us(orc_vehicle* v) { return v->now_us; }

}  // extern "C"
