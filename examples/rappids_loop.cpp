// rappids_loop.cpp -- the stepping loop of Simulator/Rappids_Simulator/main.cpp (:140-150,211-218,330-392,
// 471-739) with the Unity/AirSim calls and the offboard estimator/controller removed: one vehicle behind the
// reference's object API (agf::Simulation::Quadcopter), a 500 Hz simulation clock, a 100 Hz command source and
// a 30 ms uplink delay queue.  The command sequence is the "rates" scenario of agri-fly_b200/scenarios.py, so
// tests/test_host_facade.py can compare the printed trajectory with the CPU oracle bit for bit.
//
//   g++ -std=c++14 -Iinclude -Iagri-fly_b200/host examples/rappids_loop.cpp -Lagri-fly_b200 -lagrifly_b200 -o rappids_loop
//   (inside the reference tree add -DAGF_WITH_REFERENCE_HEADERS -I<agri-fly>/Common -I<agri-fly>/Components)
#include <cstdio>
#include <cstdlib>
#include <deque>

#include "agf_quadcopter.hpp"

using namespace agf;

int main(int argc, char** argv) {
  const unsigned nticks = argc > 1 ? unsigned(atoi(argv[1])) : 5000;
  const bool fused_tail = argc > 2 && atoi(argv[2]) != 0;  // run the last 1000 ticks with QuadcopterBatch::RunTicks

  ManualTimer simTimer;                       // main.cpp:140
  const uint32_t dt_us = 2000;                // 500 Hz, main.cpp:143
  const double cmdPeriod = 0.01;              // main.cpp:177 (offboard loop, 100 Hz)
  const uint64_t uplinkDelay_us = 30000;      // main.cpp:178

  // vehicle constants exactly as main.cpp:147-218 widens them from the float airframe table
  agf_vehicle_cfg c;
  check(agf_vehicle_cfg_from_type(agf_quad_type_from_id(1), 1, &c), "agf_vehicle_cfg_from_type");
#if defined(AGF_WITH_REFERENCE_HEADERS)
  Eigen::Matrix<double, 3, 3> inertia;
#else
  Mat3d inertia;
#endif
  for (int r = 0; r < 3; r++)
    for (int q = 0; q < 3; q++) inertia(r, q) = c.inertia[3 * r + q];
  Simulation::Quadcopter quad(&simTimer, c.mass, inertia, c.arm_length, Vec3d(0, 0, 0), c.motor_min_speed, c.motor_max_speed,
                              c.prop_thrust_from_speed_sqr, c.prop_torque_from_speed_sqr, /*motorTimeConst*/ 0.0, /*motorInertia*/ 0.0,
                              Vec3d(0, 0, 0), 1, QuadcopterType(agf_quad_type_from_id(1)), 1.0 / 500);
  quad.SetPosition(Vec3d(0, 0, 0));           // main.cpp:279-280
  quad.SetAttitude(Rotationd::Identity());
  // parity run: zero the IMU noise like the oracle harness does
  check(agf_batch_set_noise(quad.GetBatch()->handle(), 0, 0.0, 0.0, 0.0, 0.0), "agf_batch_set_noise");

  struct Queued { uint64_t due_us; RawMessage msg; };
  std::deque<Queued> cmdRadioChannel;         // CommunicationsDelay.hpp:18-39
  uint64_t sinceCmd_us = 0;                   // Timer slaved to simTimer, main.cpp:440-459

  for (unsigned k = 0; k < nticks; k++) {
    if (fused_tail && k + 1000 == nticks) {
      // the same 1000 iterations in ONE kernel launch: commands that would be generated meanwhile are scheduled
      // up front; this path is what populations use.  (No commands are generated here, so it is only valid for
      // demonstration when the remaining commands are already queued; the test uses fused_tail = 0.)
      quad.GetBatch()->RunTicks(1000, dt_us, &simTimer);
      break;
    }
    quad.Run();                               // main.cpp:391
    simTimer.AdvanceMicroSeconds(dt_us);      // main.cpp:392
    sinceCmd_us += dt_us;
    if (double(sinceCmd_us) * 1e-6 > cmdPeriod) {  // strict '>', main.cpp:471
      sinceCmd_us -= uint64_t(cmdPeriod * 1e6);
      const double t = simTimer.GetSeconds<double>();
      Vec3f w(0, 0, 0);
      if (t > 2.0 && t < 2.1) w = Vec3f(1.0f, 0.5f, 0.2f);
      else if (t >= 2.1 && t < 2.2) w = Vec3f(-1.0f, -0.5f, -0.2f);
      Queued q;
      q.due_us = simTimer.GetMicroSeconds() + uplinkDelay_us;
      RadioTypes::RadioMessageDecoded::CreateRatesCommand(0, float(1.05 * 9.81), w, q.msg.raw);  // main.cpp:666-673
      cmdRadioChannel.push_back(q);
    }
    if (!cmdRadioChannel.empty() && simTimer.GetMicroSeconds() >= cmdRadioChannel.front().due_us) {  // main.cpp:737-739
      quad.SetCommandRadioMsg(cmdRadioChannel.front().msg);
      cmdRadioChannel.pop_front();
    }
    if ((k + 1) % 1000 == 0 || k + 1 == nticks) {
      const Vec3d p = quad.GetPosition(), v = quad.GetVelocity(), w = quad.GetAngularVelocity();
      const Rotationd a = quad.GetAttitude();
      printf("%u %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g", k, p.x, p.y, p.z, v.x, v.y, v.z,
             a[0], a[1], a[2], a[3], w.x, w.y, w.z);
      for (unsigned m = 0; m < 4; m++) printf(" %.17g", quad.GetMotorSpeed(m));
      printf("\n");
    }
  }
  return 0;
}
