"""Workload driver for the ncu captures in profiles/ (see profiles/README.md for the commands).
Usage: python profiles/prof_step.py [fp32|fp64] [uwb|rates] [vehicles] [ticks] [launches]
Environment: AGF_PROF_HK=0|1 (housekeeping, default 1), AGF_PROF_MATH=fast|parity, AGF_PROF_C4=1 (per-vehicle parameter sweep +
trajectory log every tick, ring of 32 records), AGF_NO_WARM=1 (no clock warm-up launches)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import agrifly_b200 as agf  # noqa: E402
import bench  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
uwb = (sys.argv[2] if len(sys.argv) > 2 else "uwb") == "uwb"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 131072
ticks = int(sys.argv[4]) if len(sys.argv) > 4 else 100
launches = int(sys.argv[5]) if len(sys.argv) > 5 else 4
import time  # noqa: E402

warm = 0 if os.environ.get("AGF_NO_WARM") else 12  # extra warm-up launches: clocks ramp up over tens of ms
hk = os.environ.get("AGF_PROF_HK", "1") != "0"
c4 = os.environ.get("AGF_PROF_C4", "0") == "1"
b, _ = bench.workload(agf, n, 0, prec, ticks * (launches + warm + 1) + 600, uwb=uwb, hk=hk, math=os.environ.get("AGF_PROF_MATH", "fast"),
                      sweep=c4)
if c4 and os.environ.get("AGF_PROF_LOG", "1") != "0":  # AGF_PROF_LOG=0: the sweep without the log
    b.enable_log(1, 32)
b.run(500)  # first launch: take-off, EKF initialised, UWB ranging active
t0 = time.time()
for _ in range(warm):
    b.run(ticks)
    b.sync()
    if time.time() - t0 > 0.4:
        break
b.step_kernel_time()
for _ in range(launches):
    b.run(ticks)
b.sync()
ms, nl = b.step_kernel_time()
print("step kernel: %d launches, %.3f ms total; last %d: %.3e vehicle-steps/s" %
      (nl, ms, launches, 0 if nl == 0 else n * ticks * nl / (ms * 1e-3)))
st = b.stats()
print("panic", st[6], "nonfinite", st[9], "rms err", np.sqrt(st[4] / st[0]))
