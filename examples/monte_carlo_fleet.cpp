// monte_carlo_fleet.cpp -- BASELINE config 3 from a C++ host with no Python in the process: a Monte-Carlo population of
// vehicles tracking a 4-waypoint square with the full onboard loop (IMU noise, 9-state EKF on 8 UWB anchors, position /
// attitude / rate control), sharded by index range over the GPUs of one node, ONE host thread and one agf_batch per GPU,
// nothing exchanged while stepping, and the tracking-error statistics reduced over NCCL at each read-out
// (agf_batch_reduce_stats_nccl: statistics kernel + one all-gather + combine on each batch's stream).
//
//   g++ -std=c++14 -pthread -Iinclude examples/monte_carlo_fleet.cpp -Lagri-fly_b200 -lagrifly_b200 -o monte_carlo_fleet
//   ./monte_carlo_fleet [gpus] [vehicles per gpu] [seconds] [fp32|fp64]
//
// Prints one line per read-out (every simulated 2.5 s) and a final line
//   STATS <vehicles> <rms error> <max error> <panics> <non-finite> <sum e^2>
// that tests/test_host_facade.py compares with the same population stepped in one batch.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "agrifly_b200.h"

static void check(int rc, const char* what) {
  if (rc != AGF_OK) {
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, agf_last_error_string());
    exit(1);
  }
}

// Philox-free initial states: a small hash of the GLOBAL vehicle index, so that a shard's vehicles start exactly where they
// would in an unsharded run
static double unit(uint64_t i, uint32_t k) {
  uint64_t z = (i + 1) * 0x9E3779B97F4A7C15ull + k * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return double(z >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char** argv) {
  const int gpus = argc > 1 ? atoi(argv[1]) : 1;
  const size_t per_gpu = argc > 2 ? size_t(atol(argv[2])) : 131072;
  const double seconds = argc > 3 ? atof(argv[3]) : 10.0;
  const bool fp64 = argc > 4 && !strcmp(argv[4], "fp64");
  const uint32_t ticks_total = uint32_t(seconds * 500.0 + 0.5), ticks_per_launch = 1250;  // 2.5 s = one leg of the square

  // the command schedule shared by the population: idle for 0.5 s, then the waypoints (+-1, +-1, 1.5), 2.5 s each, as 50 Hz
  // position commands with the 20 ms uplink delay of the ROS simulator (Simulator/main.cpp:234)
  std::vector<agf_cmd_entry> sched;
  const float wp[4][3] = {{1, 1, 1.5f}, {-1, 1, 1.5f}, {-1, -1, 1.5f}, {1, -1, 1.5f}}, zero[3] = {0, 0, 0};
  for (uint32_t g = 10; g + 11 < ticks_total; g += 10) {
    agf_cmd_entry e;
    memset(&e, 0, sizeof e);
    e.tick = g + 11;
    e.slot = -1;
    const double t = (g + 1) * 0.002;
    if (t < 0.5) agf_radio_encode_idle(0, e.raw);
    else agf_radio_encode_position(0, wp[int((t - 0.5) / 2.5) % 4], zero, zero, e.raw);
    sched.push_back(e);
  }

  std::vector<void*> comms(gpus, nullptr);
  if (gpus > 1) {
    std::vector<int> devs(gpus);
    for (int d = 0; d < gpus; d++) devs[d] = d;
    check(agf_nccl_comm_init_all(gpus, devs.data(), comms.data()), "agf_nccl_comm_init_all");
  }
  std::vector<std::vector<double>> results(gpus, std::vector<double>(AGF_STATS_LEN, 0.0));

  auto worker = [&](int rank) {
    agf_vehicle_cfg cfg;
    check(agf_vehicle_cfg_from_type(agf_quad_type_from_id(1), 1, &cfg), "agf_vehicle_cfg_from_type");
    cfg.motor_time_const = 0.015;
    agf_batch_opts o;
    agf_batch_opts_default(&o);  // the reference's IMU noise is on by default
    o.device = rank;
    o.precision = fp64 ? AGF_PREC_FP64 : AGF_PREC_FP32;
    o.math = AGF_MATH_FAST;
    o.uwb_comm_period = 0.004;
    o.seed = 7;
    o.first_global_index = uint64_t(rank) * per_gpu;  // noise streams are keyed by the population-wide vehicle index
    agf_batch* b = nullptr;
    check(agf_batch_create(&cfg, 1, per_gpu, &o, &b), "agf_batch_create");
    for (int a = 0; a < 8; a++) {
      const float p[3] = {(a & 2) ? -3.0f : 3.0f, (a & 1) ? -3.0f : 3.0f, (a & 4) ? 3.0f : 0.1f};
      check(agf_batch_add_uwb_anchor(b, uint8_t(101 + a), p), "agf_batch_add_uwb_anchor");
    }
    std::vector<double> s13(per_gpu * 13, 0.0);
    for (size_t i = 0; i < per_gpu; i++) {
      const uint64_t gi = o.first_global_index + i;
      double* s = &s13[13 * i];
      s[0] = 2.0 * unit(gi, 0) - 1.0;
      s[1] = 2.0 * unit(gi, 1) - 1.0;
      const double yaw = (2.0 * unit(gi, 2) - 1.0) * 1.0;
      s[6] = cos(0.5 * yaw);
      s[9] = sin(0.5 * yaw);
    }
    check(agf_batch_set_state(b, s13.data(), 0, per_gpu), "agf_batch_set_state");
    check(agf_batch_set_cmd_schedule(b, sched.data(), sched.size()), "agf_batch_set_cmd_schedule");
    double st[AGF_STATS_LEN];
    for (uint32_t done = 0; done < ticks_total;) {
      const uint32_t c = ticks_total - done < ticks_per_launch ? ticks_total - done : ticks_per_launch;
      check(agf_batch_run(b, 2000, c), "agf_batch_run");
      done += c;
      check(agf_batch_reduce_stats_nccl(b, comms[rank], nullptr, st), "agf_batch_reduce_stats_nccl");  // collective
      if (rank == 0)
        printf("t = %5.2f s: %.0f vehicles on %d GPU(s), rms tracking error %.4f m, max %.3f m, panics %.0f\n", done * 0.002,
               st[AGF_ST_COUNT], gpus, sqrt(st[AGF_ST_SUM_E2] / st[AGF_ST_COUNT]), st[AGF_ST_MAX_ENORM], st[AGF_ST_N_PANIC]);
    }
    memcpy(results[rank].data(), st, sizeof st);
    check(agf_batch_destroy(b), "agf_batch_destroy");
  };

  std::vector<std::thread> th;
  for (int r = 0; r < gpus; r++) th.emplace_back(worker, r);
  for (auto& t : th) t.join();
  for (int r = 1; r < gpus; r++)  // every rank holds the same combined vector, bit for bit
    if (memcmp(results[r].data(), results[0].data(), sizeof(double) * AGF_STATS_LEN)) {
      fprintf(stderr, "rank %d disagrees with rank 0\n", r);
      return 2;
    }
  for (int r = 0; r < gpus; r++) check(agf_nccl_comm_destroy(comms[r]), "agf_nccl_comm_destroy");
  const double* s = results[0].data();
  printf("STATS %.0f %.9g %.9g %.0f %.0f %.12g\n", s[AGF_ST_COUNT], sqrt(s[AGF_ST_SUM_E2] / s[AGF_ST_COUNT]), s[AGF_ST_MAX_ENORM],
         s[AGF_ST_N_PANIC], s[AGF_ST_N_NONFINITE], s[AGF_ST_SUM_E2]);
  return 0;
}
