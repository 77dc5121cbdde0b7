"""Workload driver for the batched RAPPIDS planner (bench C5 workload at a chosen size).
usage: python profiles/prof_rappids.py <fast|parity> <n_vehicles> <k_candidates> <reps> [hard]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import agrifly_b200 as agf


def main():
    math = agf.abi.MATH_FAST if sys.argv[1] == "fast" else agf.abi.MATH_PARITY
    n, k, reps = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    hard = len(sys.argv) > 5 and sys.argv[5] == "hard"
    kw = dict(speed_max=4.5, acc_max=3.0, box_depth=(1.0, 3.0), n_boxes=(2, 4)) if hard else {}
    t = time.time()
    pop = agf.scenarios.rappids_population(n, seed=2024, **kw)
    print("population %.1fs" % (time.time() - t))
    cfg = agf.rappids_cfg(math=math)
    with agf.Rappids(cfg, n, k) as pl:
        pl.render_scenes(pop["row_bg"], pop["boxes"])
        pl.set_states(pop["vel0"], pop["acc0"], pop["grav"])
        pl.sample_candidates(k, seed=7)
        if os.environ.get("AGF_PROF_DISPATCH") == "index":
            pl.set_dispatch(False)
        pl.plan()  # warm-up: the first launch carries the lazy module load between its timing events
        pl.sync()
        pl.plan_kernel_time()
        for _ in range(reps):
            pl.plan()
        pl.sync()
        ms, cnt = pl.plan_kernel_time()
        st = pl.stats()
        print("n=%d k=%d %s: %.3f ms per plan launch (%d launches) -> %.3e plans/s, %.3e candidates/s" %
              (n, k, sys.argv[1], ms, cnt, n / ms * 1e3, n * k / ms * 1e3))
        print({a: b / n for a, b in st.items()})


if __name__ == "__main__":
    main()
