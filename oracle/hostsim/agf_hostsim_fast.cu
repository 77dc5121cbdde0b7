// oracle/hostsim/agf_hostsim_fast.cu -- TEST INFRASTRUCTURE: the FAST instantiations of the product's device step
// header (agri-fly_b200/csrc/agf_step.cuh: FP32 or FP64 plant, fast arithmetic, packed symmetric EKF in the scratch)
// compiled for the HOST behind oracle/oracle_api.h.
//
// Purpose: a development aid for the build container, which has no GPU.  The fast kernels are held to a tolerance, not
// to bit equality; with this library their formulation (block-wise EKF propagation, closed-form rotations, reciprocal
// parameters, FP32 plant integration) can be compared with the reference trajectories on the CPU before spending GPU
// time.  It is NOT the kernel: the host compiler contracts FMAs differently and its libm is not CUDA's, so numbers from
// here are indicative only -- the -m gpu tests (tests/test_fast_population_gpu.py) are the check.  Nothing in
// agri-fly_b200/ loads it.  Flavours: hostsim-fast32 (-DHOSTSIM_F64=0), hostsim-fast64 (-DHOSTSIM_F64=1).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>

#include "../oracle_api.h"
#include "../../agri-fly_b200/csrc/agf_host_params.h"
#include "../../agri-fly_b200/csrc/agf_step.cuh"

using namespace agf;

#ifndef HOSTSIM_F64
#define HOSTSIM_F64 0
#endif
#if HOSTSIM_F64
typedef double HP;
#define ORC_FLAVOUR "hostsim-fast64"
#else
typedef float HP;
#define ORC_FLAVOUR "hostsim-fast32"
#endif
static constexpr int kVP = int(16 / sizeof(HP));

struct orc_vehicle {
  agf_vehicle_cfg cfg;
  orc_opts opts;
  StepShared<HP> sh;
  PlantPV<HP> pv;
  Timing ts;
  bool uwb;
  std::vector<HP> hp;
  std::vector<float> hf, hc;
  std::vector<uint32_t> hu;
  std::vector<float4> scratch;  // the thread's shared-memory scratch: [quad][SQ_STRIDE], this "thread" is lane 0
  uint64_t tick, now_us;
  StateArrays<HP> arrays() {
    StateArrays<HP> a;
    a.sp = (typename VecOf<HP>::type*)hp.data();
    a.sf = (float4*)hf.data();
    a.su = (uint4*)hu.data();
    a.sc = uwb ? (float4*)hc.data() : nullptr;
    a.sq = nullptr;
    return a;
  }
};

template<bool UWB>
static void run_impl(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_cmd_entry* sched, uint32_t nsched,
                     const uint8_t* slot_raw, double* traj) {
  VState<HP, false, UWB, true> s;
  StateArrays<HP> a = v->arrays();
  Scratch sc;
  sc.q = v->scratch.data();
  state_load(s, a, 1, 0, sc, v->sh.logic.mix_kf);
  uint32_t si = 0;
  while (si < nsched && sched[si].tick < v->tick) si++;
  for (uint32_t k = 0; k < nticks; k++) {
    const uint32_t plant_dt = v->ts.integ_age ? v->ts.integ_age : dt_us;
    v->pv.motor_c = HP(motor_c_host(v->cfg.motor_time_const, plant_dt));
    if (si < nsched && sched[si].tick == v->tick) {
      const uint8_t* raw = sched[si].raw;
      if (sched[si].slot >= 0 && slot_raw) raw = slot_raw + AGF_RADIO_PACKET_SIZE * sched[si].slot;
      uint8_t type, flags;
      float f[10];
      agf_radio_decode(raw, &type, &flags, f);
      radio_deliver(s, v->sh.logic, type, flags, f);
      si++;
    }
    const TickPlan plan = timing_plan(v->ts, v->sh.tc, dt_us);
    tick<HP, false, UWB, true, false>(s, sc, v->sh, v->pv, plan, v->ts.now_us, dt_us, v->tick, 0, 0, 1);
    timing_advance(v->ts, v->sh.tc, plan, dt_us);
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

static void record(orc_vehicle* v, double* r) {  // the stored state, as agf_batch_get_* would return it
  auto P_ = [&](int k) { return double(v->hp[sidx(k, 1, 0, kVP)]); };
  auto F_ = [&](int k) { return double(v->hf[sidx(k, 1, 0, 4)]); };
  auto U_ = [&](int k) { return v->hu[sidx(k, 1, 0, 4)]; };
  for (int c = 0; c < 3; c++) { r[c] = P_(SP_POS + c); r[3 + c] = P_(SP_VEL + c); r[10 + c] = P_(SP_W + c); r[21 + c] = F_(SF_KPOS + c); r[24 + c] = F_(SF_KVEL + c); r[31 + c] = F_(SF_KW + c); }
  for (int c = 0; c < 4; c++) { r[6 + c] = P_(SP_ATT + c); r[13 + c] = P_(SP_MS + c); r[17 + c] = F_(SF_CMD + c); r[27 + c] = F_(SF_KATT + c); }
  const uint32_t bits = U_(SU_BITS), kc = U_(SU_KFCNT);
  r[34] = bits & 7u;
  r[35] = (bits >> 3) & 7u;
  r[36] = U_(SU_CYCLE);
  r[37] = kc & 0xFFFFu;
  r[38] = kc >> 16;
  r[39] = U_(SU_UWB_COUNT);
}

static void unsupported(const char* what) {
  fprintf(stderr, "%s: %s is not part of this development aid\n", ORC_FLAVOUR, what);
  abort();
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
  initial_state<HP>(1, v->sh.logic, cfg->logic.low_battery_threshold, v->uwb, v->hp, v->hf, v->hu, v->hc);
  v->scratch.assign(size_t(SQ_QUADS_UWB) * SQ_STRIDE, make_float4(0, 0, 0, 0));
  v->tick = 0;
  v->now_us = 0;
  return v;
}
void orc_destroy(orc_vehicle* v) { delete v; }

void orc_set_state(orc_vehicle* v, const double p[3], const double vel[3], const double a[4], const double w[3]) {
  for (int k = 0; k < 3; k++) {
    v->hp[sidx(SP_POS + k, 1, 0, kVP)] = HP(p[k]);
    v->hp[sidx(SP_VEL + k, 1, 0, kVP)] = HP(vel[k]);
    v->hp[sidx(SP_W + k, 1, 0, kVP)] = HP(w[k]);
  }
  for (int k = 0; k < 4; k++) v->hp[sidx(SP_ATT + k, 1, 0, kVP)] = HP(a[k]);
}
void orc_set_external(orc_vehicle*, const double*, const double*) { unsupported("orc_set_external"); }
int orc_add_anchor(orc_vehicle* v, uint8_t id, float x, float y, float z) {
  if (v->sh.n_anchors >= AGF_MAX_UWB_ANCHORS) return -1;
  AnchorDev& a = v->sh.anchors[v->sh.n_anchors++];
  a.id = id; a.x = x; a.y = y; a.z = z;
  v->sh.tc.n_anchors = int(v->sh.n_anchors);
  return 0;
}
void orc_set_radio(orc_vehicle*, const uint8_t*) { unsupported("orc_set_radio"); }
void orc_run(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_cmd_entry* sched, uint32_t nsched,
             const uint8_t* slot_raw, double* traj) {
  if (v->uwb) run_impl<true>(v, dt_us, nticks, sched, nsched, slot_raw, traj);
  else run_impl<false>(v, dt_us, nticks, sched, nsched, slot_raw, traj);
}
void orc_run_offboard(orc_vehicle*, uint32_t, uint32_t, const agf_offboard_cfg*, const agf_offboard_target*, uint32_t, const double*,
                      double*) {
  unsupported("orc_run_offboard");
}
void orc_get_full(orc_vehicle*, orc_full_state*) { unsupported("orc_get_full"); }
uint64_t orc_time_us(orc_vehicle* v) { return v->now_us; }

#include "../orc_population_traj.inc"

}  // extern "C"
