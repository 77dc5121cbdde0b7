"""Workload driver for the step kernels with the in-kernel offboard loop (OFFB instantiations).
usage: python profiles/prof_offboard.py <fp32|fp64> <truth|mocap> <targets|stages> <vehicles> <ticks> <launches>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import agrifly_b200 as agf  # noqa: E402

prec, est, refk = sys.argv[1], sys.argv[2], sys.argv[3]
n, ticks, launches = int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
b = agf.Batch(agf.vehicle_cfg(vehicle_id=1, motor_time_const=0.015), n,
              precision=agf.abi.PREC_FP32 if prec == "fp32" else agf.abi.PREC_FP64, math=agf.abi.MATH_FAST, telemetry_warnings=False)
b.set_offboard_loop(agf.offboard_cfg(5), [(0, (0.0, 0.0, 2.0)), (3000000, (1.0, -0.5, 2.5))])
if est == "mocap":
    b.set_offboard_estimator(agf.offboard_estimator())
if refk == "stages":
    b.set_offboard_reference(agf.abi.OFFREF_STAGES, start_us=500000, desired_pos=(0.0, 0.0, 1.0), traj_id=3)
b.run(ticks)
b.sync()
b.step_kernel_time()
for _ in range(launches):
    b.run(ticks)
b.sync()
ms, nl = b.step_kernel_time()
print("%s %s %s: %d launches, %.3f ms total -> %.3e vehicle-steps/s" % (prec, est, refk, nl, ms, n * ticks * nl / (ms * 1e-3)))
b.close()
