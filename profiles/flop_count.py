"""Workload for counting executed floating-point instructions per vehicle-step with ncu hardware counters
(SURVEY.md 8d asks for an instrumented re-derivation of the hand count).

    ncu --metrics <op counters> -k regex:step_kernel -s 1 -c 1 python profiles/flop_count.py parity|fast fp64|fp32 uwb|rates [n] [ticks] [hk|nohk]

The first launch (600 ticks: take-off, EKF initialised, ranging active) is skipped by `-s 1`; the second launch
of `ticks` ticks is the one counted.  flop/step = (fadd + fmul + 2 ffma [+ dadd + dmul + 2 dfma]) / (n * ticks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import agrifly_b200 as agf  # noqa: E402

math = sys.argv[1] if len(sys.argv) > 1 else "parity"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp64"
uwb = (sys.argv[3] if len(sys.argv) > 3 else "uwb") == "uwb"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
ticks = int(sys.argv[5]) if len(sys.argv) > 5 else 300
hk = (sys.argv[6] == "hk") if len(sys.argv) > 6 else (math == "parity")  # housekeeping: always on in the parity kernels
s = agf.scenarios
cfg = agf.vehicle_cfg(vehicle_id=1, motor_time_const=0.015)
b = agf.Batch(cfg, n, precision=agf.abi.PREC_FP32 if prec == "fp32" else agf.abi.PREC_FP64,
              math=agf.abi.MATH_PARITY if math == "parity" else agf.abi.MATH_FAST,
              uwb_comm_period=0.004 if uwb else 0.0, sigma_gyro=0.1, sigma_acc=0.2, seed=7,
              telemetry_warnings=hk)
if uwb:
    for i, p in s.ANCHORS_8:
        b.add_anchor(i, p)
    b.set_schedule(s.waypoint_square_schedule(agf.codec, nticks=600 + ticks + 8))
else:
    raw = agf.codec.encode_rates(0, 9.81 * 1.02, (0.0, 0.0, 0.0))
    b.set_schedule([(k, raw, -1) for k in range(0, 600 + ticks + 8, 10)])
b.set_state13(s.monte_carlo_initial_states(n, seed=1234, yaw_max=np.pi / 3))
b.run(600)
b.run(ticks)
b.sync()
st = b.stats()
print("flop_count workload: %s %s %s hk=%d n=%d ticks=%d panic=%d nonfinite=%d" % (math, prec, "uwb" if uwb else "rates", hk, n, ticks, st[6], st[9]))
