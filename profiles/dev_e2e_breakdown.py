"""Where does the e2e leg of bench.py spend its time?  Same calls as bench.py's e2e_step, each followed by a sync and timed."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import agrifly_b200 as agf  # noqa: E402
import bench  # noqa: E402

n, S = 131072, int(sys.argv[1]) if len(sys.argv) > 1 else 15000
b, init = bench.workload(agf, n, 0, "fp32", S * 12 + 16, hk=True, uwb=True)
pin13 = torch.from_numpy(np.ascontiguousarray(init[:, 0:13])).pin_memory()
pos_out = torch.empty((n, 3), dtype=torch.float64).pin_memory()
for _ in range(2):
    b.run(S)
b.sync()
b.step_kernel_time()
for it in range(4):
    t = [time.time()]
    agf._check(b.L.agf_batch_set_state(b.h, pin13.data_ptr(), 0, n)); b.sync(); t.append(time.time())
    b.run(S); b.sync(); t.append(time.time())
    agf._check(b.L.agf_batch_get_field(b.h, 0, pos_out.data_ptr(), 0, n)); t.append(time.time())
    st = b.stats(); t.append(time.time())
    kms, kl = b.step_kernel_time()
    d = np.diff(t) * 1e3
    print("set_state %.2f ms, run %.2f ms (kernel %.2f ms), get_field %.2f ms, stats %.2f ms; panic %d killed %d nonfinite %d" %
          (d[0], d[1], kms, d[2], d[3], st[6], st[7], st[9]), flush=True)
for it in range(3):
    t0 = time.time(); b.run(S); b.sync(); t1 = time.time()
    kms, kl = b.step_kernel_time()
    print("run only: %.2f ms (kernel %.2f ms)" % ((t1 - t0) * 1e3, kms), flush=True)
