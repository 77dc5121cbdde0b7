// rappids_fleet.cpp -- what Rappids_Simulator does for one vehicle (Simulator/Rappids_Simulator/main.cpp:330-745), for a
// fleet and with the whole loop on the device: Run() at 500 Hz, mocap packets at 200 Hz into the offboard estimator, the
// offboard controller at 100 Hz, 16-bit rates commands through a 30 ms delay queue -- all inside the step kernel
// (include/agrifly_b200.h "offboard rates loop", "reference generators", "state estimator").  The desired state comes from the
// flight stages of the ROS rates-control node (spool-up, take-off, circle, landing).  The host only reads vehicle 0 back
// every 10 ms and writes the reference's simulation.csv (main.cpp:266-270,676-733) to stdout.
//
//   g++ -std=c++14 -Iinclude examples/rappids_fleet.cpp -Lagri-fly_b200 -lagrifly_b200 -o rappids_fleet
//   ./rappids_fleet [vehicles] [seconds] [imu noise 1|0] > simulation.csv
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "agrifly_b200.h"

static void check(int rc, const char* what) {
  if (rc != AGF_OK) {
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, agf_last_error_string());
    exit(1);
  }
}

int main(int argc, char** argv) {
  const size_t n = argc > 1 ? size_t(atol(argv[1])) : 4096;
  const double seconds = argc > 2 ? atof(argv[2]) : 6.0;
  agf_vehicle_cfg cfg;
  check(agf_vehicle_cfg_from_type(agf_quad_type_from_id(1), 1, &cfg), "agf_vehicle_cfg_from_type");  // main.cpp:147-150
  cfg.motor_time_const = 0.015;
  agf_batch_opts o;
  agf_batch_opts_default(&o);  // the reference's IMU noise (Quadcopter_T.cpp:5-6) is on by default
  if (argc > 3 && atoi(argv[3]) == 0) o.sigma_acc = o.sigma_gyro = 0.0;
  agf_batch* b = nullptr;
  check(agf_batch_create(&cfg, 1, n, &o, &b), "agf_batch_create");

  agf_offboard_cfg oc;
  check(agf_offboard_cfg_default(cfg.quad_type, &oc), "agf_offboard_cfg_default");  // 100 Hz, 30 ms (main.cpp:175-178)
  const agf_offboard_target unused = {0, {0.0, 0.0, 0.0}};
  check(agf_batch_set_offboard_loop(b, &oc, &unused, 1, nullptr), "agf_batch_set_offboard_loop");
  agf_offboard_estimator est;
  check(agf_offboard_estimator_default(&est), "agf_offboard_estimator_default");  // MocapStateEstimator, 200 Hz (main.cpp:174)
  check(agf_batch_set_offboard_estimator(b, &est), "agf_batch_set_offboard_estimator");
  agf_offboard_ref ref;
  memset(&ref, 0, sizeof ref);
  ref.kind = AGF_OFFREF_STAGES;
  ref.traj_id = 3;                       // circle at fixed height (ExampleVehicleStateMachine.cpp:252-264)
  ref.start_us = 500000;
  ref.stop_us = uint64_t((seconds - 3.0) * 1e6);
  ref.desired_pos[2] = 1.0;              // QuadMocapRatesControl/main.cpp:82
  check(agf_batch_set_offboard_reference(b, &ref), "agf_batch_set_offboard_reference");

  char line[2048];
  agf_csv_header(line, sizeof line);
  fputs(line, stdout);
  const unsigned rounds = unsigned(seconds * 100.0 + 0.5);
  for (unsigned k = 1; k <= rounds; k++) {
    check(agf_batch_run(b, 2000, 5), "agf_batch_run");  // 10 ms: five ticks, one launch for the whole fleet
    agf_csv_record r;
    memset(&r, 0, sizeof r);
    r.t = k * 0.01;
    check(agf_batch_get_field(b, AGF_F_POSITION, r.pos, 0, 1), "get position");
    check(agf_batch_get_field(b, AGF_F_VELOCITY, r.vel, 0, 1), "get velocity");
    check(agf_batch_get_field(b, AGF_F_ATTITUDE, r.att, 0, 1), "get attitude");
    check(agf_batch_get_field(b, AGF_F_ANGULAR_VELOCITY, r.ang_vel, 0, 1), "get angular velocity");
    uint8_t p1[AGF_TELEMETRY_PACKET_SIZE], p2[AGF_TELEMETRY_PACKET_SIZE];
    check(agf_batch_get_telemetry(b, p1, p2, 0, 1), "agf_batch_get_telemetry");  // main.cpp:662-670
    agf_telemetry t1, t2;
    agf_telemetry_decode(p1, &t1);
    agf_telemetry_decode(p2, &t2);
    for (int m = 0; m < 4; m++) r.motor_forces[m] = t1.motor_forces[m];
    r.panic_reason = t2.panic_reason;
    double e[13], st[AGF_OFFSTATE_DOUBLES];
    check(agf_batch_get_offboard_estimate(b, 0.0, e, nullptr, 0, 1), "agf_batch_get_offboard_estimate");  // est->GetPrediction(0), :698
    for (int a = 0; a < 3; a++) {
      r.est_pos[a] = float(e[a]);
      r.est_vel[a] = float(e[3 + a]);
      r.est_ang_vel[a] = float(e[10 + a]);
    }
    for (int a = 0; a < 4; a++) r.est_att[a] = float(e[6 + a]);
    check(agf_batch_get_offboard_state(b, st, 0, 1), "agf_batch_get_offboard_state");
    for (int a = 0; a < 3; a++) {  // desired position / velocity as last commanded by the stage machine
      r.des_pos[a] = st[6 + a];
      r.des_vel[a] = st[9 + a];
    }
    agf_csv_format_row(&r, line, sizeof line);
    fputs(line, stdout);
  }
  double stats[16];
  check(agf_batch_reduce_stats(b, nullptr, stats), "agf_batch_reduce_stats");
  fprintf(stderr, "%zu vehicles, %.1f s: %llu kernel launches, panics %.0f, non-finite %.0f\n", n, seconds,
          (unsigned long long)agf_batch_launch_count(b), stats[6], stats[9]);
  check(agf_batch_destroy(b), "agf_batch_destroy");
  return 0;
}
