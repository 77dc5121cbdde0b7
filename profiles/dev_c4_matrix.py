"""C4 decomposition: what do per-vehicle plant parameters and the per-tick trajectory log each cost the step kernel?
Usage: python profiles/dev_c4_matrix.py [rates|uwb] [vehicles] [ticks] [stride ...]
Kernel-only rates (CUDA events inside the library), one warm launch + two timed, for
  shared parameters / sweep  x  no log / log every `stride` ticks."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import agrifly_b200 as agf  # noqa: E402
import bench  # noqa: E402

uwb = (sys.argv[1] if len(sys.argv) > 1 else "rates") == "uwb"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 21
ticks = int(sys.argv[3]) if len(sys.argv) > 3 else 500
strides = [int(a) for a in sys.argv[4:]] or [1]
for sweep in (False, True):
    for stride in [0] + strides:
        b, _ = bench.workload(agf, n, 0, "fp32", ticks * 4 + 600, uwb=uwb, hk=True, sweep=sweep)
        if stride:
            b.enable_log(stride, 32)
        b.run(500)
        sps, ms, nl = bench.kernel_rate(b, n, ticks, launches=2)
        print("%s sweep=%d log_stride=%d: %.3e vehicle-steps/s (%.3f ms per %d-tick launch)%s" %
              ("uwb" if uwb else "rates", sweep, stride, sps, ms / nl, ticks,
               "" if not stride else "  log %.0f GB/s" % (sps / stride * 68 / 1e9)), flush=True)
        b.close()
