"""Per-vehicle work of a plan launch next to its counters (which plans are the long ones?).
usage: python profiles/dev_rappids_work.py <n> <k> <out.npz> [hard]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import agrifly_b200 as agf

n, k, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
hard = len(sys.argv) > 4 and sys.argv[4] == "hard"
kw = dict(speed_max=4.5, acc_max=3.0, box_depth=(1.0, 3.0), n_boxes=(2, 4)) if hard else {}
pop = agf.scenarios.rappids_population(n, seed=2024, **kw)
with agf.Rappids(agf.rappids_cfg(math=agf.abi.MATH_FAST), n, k) as pl:
    pl.render_scenes(pop["row_bg"], pop["boxes"])
    pl.set_states(pop["vel0"], pop["acc0"], pop["grav"])
    pl.sample_candidates(k, seed=7)
    for _ in range(3):
        pl.plan()
    pl.sync()
    res = pl.results()
    work = pl.plan_work()
    flags = pl.candidate_flags() if hasattr(pl, "candidate_flags") else None
    np.savez_compressed(out, work=work, res=res)
    o = np.argsort(-work.astype(np.int64))
    print("work cycles: mean %.3g  median %.3g  p99 %.3g  max %.3g  sum %.4g" % (work.mean(), np.median(work), np.percentile(work, 99), work.max(), work.sum()))
    for v in o[:12]:
        r = res[v]
        print(v, work[v], {f: int(r[f]) for f in ("found", "n_cost_checks", "n_collision_checks", "n_velocity_checks", "n_collision_free", "n_pyramids")})
