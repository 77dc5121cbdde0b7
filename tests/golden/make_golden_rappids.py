"""Generates tests/golden/rappids_vectors.npz from the UNMODIFIED reference planner (oracle/_ref/
libagf_rappids_ref_*.so = DepthImagePlanner.cpp + RapidTrajectoryGenerator.cpp + SingleAxisTrajectory.cpp
compiled where they lie under /root/reference by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden_rappids.py

The reference holds no fixtures for the planner; these are outputs of the reference itself,
recorded so the port -- and through it the CUDA kernel -- stays pinned where /root/reference does not exist.
Per libm flavour and scene family ("easy" = SURVEY 8d C5 scenes, "hard" = faster vehicles, closer/more boxes):
scene descriptions (integer, rasterised identically by numpy and the device), initial states, the candidates the
reference's own RandomTrajectoryGenerator(std::mt19937(seed = vehicle index)) drew, and everything
FindLowestCostTrajectory returned: result struct, per-candidate TrajectoryTestResult flags, pyramids.
Also known-answer vectors of the pieces: RootFinder cubic/quartic, RapidTrajectoryGenerator coefficients and
feasibility verdicts.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import agrifly_b200 as agf  # noqa: E402
import orc_rappids as R  # noqa: E402

N_VEH, K_CAND, MAX_PYR = 16, 256, 32
FAMILIES = {"easy": {}, "hard": dict(speed_max=4.5, acc_max=3.0, box_depth=(1.0, 3.0), n_boxes=(2, 4))}
OUT_INTS = ("found", "best_index", "n_generated", "n_cost_checks", "n_collision_checks", "n_velocity_checks",
            "n_collision_free", "n_pyramids")


def plan_population(P, pop, imgs, k=K_CAND, cfg=None, candidates=None):
    """Runs one reference planner call per vehicle -> dict of stacked arrays."""
    n = imgs.shape[0]
    cfg = cfg or R.default_cfg(max_pyramids=MAX_PYR)
    ints = np.zeros((n, len(OUT_INTS)), dtype=np.int32)
    cost = np.zeros(n)
    tf = np.zeros(n)
    coeffs = np.zeros((n, 6, 3))
    flags = np.zeros((n, k), dtype=np.uint8)
    cands = np.zeros((n, k, 4))
    pyr = np.full((n, MAX_PYR, R.PYR), np.nan)
    for i in range(n):
        r = P.plan(cfg, imgs[i], pop["vel0"][i], pop["acc0"][i], pop["grav"][i], n=k, seed=i,
                   candidates=None if candidates is None else candidates[i], max_pyr=MAX_PYR)
        ints[i] = [r[f] for f in OUT_INTS]
        cost[i], tf[i], coeffs[i] = r["best_cost"], r["best_tf"], r["best_coeffs"]
        flags[i], cands[i] = r["results"], r["candidates"]
        pyr[i, :len(r["pyramids"])] = r["pyramids"]
    return dict(ints=ints, cost=cost, tf=tf, coeffs=coeffs, flags=flags, cands=cands, pyr=pyr)


def piece_inputs():
    rng = np.random.default_rng(99)
    cubics = np.concatenate([rng.uniform(-10, 10, (60, 3)), [[0, 0, 0], [-6, 11, -6], [3, 3, 1], [0, -1, 0], [1e-9, 0, -1]]])
    quartics = np.concatenate([rng.uniform(-10, 10, (60, 4)), [[-10, 35, -50, 24], [0, 0, 0, 0], [0, -5, 0, 4],
                                                              [0, 2, 0, 1], [-4, 6, -4, 1], [0, 0, 0, -1]]])
    prim = np.zeros((64, 13))  # vel0 acc0 grav goal T
    prim[:, 0:3] = rng.uniform(-3, 3, (64, 3))
    prim[:, 3:6] = rng.uniform(-4, 4, (64, 3))
    prim[:, 6:9] = [0.0, 9.81, 0.0]
    prim[:, 9:12] = rng.uniform(-2, 2, (64, 3)) + [0, 0, 2.5]
    prim[:, 12] = rng.uniform(0.3, 3.0, 64)
    return cubics, quartics, prim


def pieces(P, cubics, quartics, prim):
    cr = np.full((len(cubics), 4), np.nan)
    for i, c in enumerate(cubics):
        n, r = P.solve_cubic(*c)
        cr[i, 0] = n
        cr[i, 1:1 + n] = r[:n]
    qr = np.full((len(quartics), 5), np.nan)
    for i, c in enumerate(quartics):
        n, r = P.solve_quartic(*c)
        qr[i, 0] = n
        qr[i, 1:1 + n] = r[:n]
    pr = np.zeros((len(prim), 11))
    for i, p in enumerate(prim):
        abg, ir, vr = P.primitive(p[0:3], p[3:6], p[6:9], p[9:12], p[12])
        pr[i, :9] = abg.ravel()
        pr[i, 9:] = (ir, vr)
    return cr, qr, pr


def main():
    out = {}
    scn = agf.scenarios
    cubics, quartics, prim = piece_inputs()
    out["pieces/cubics"], out["pieces/quartics"], out["pieces/prim"] = cubics, quartics, prim
    for flavour in ("ref-glibc", "ref-shared"):
        P = R.Planner(flavour)
        for fam, kw in FAMILIES.items():
            pop = scn.rappids_population(N_VEH, seed=11, **kw)
            imgs = scn.rappids_render(pop["row_bg"], pop["boxes"], pop["width"])
            g = plan_population(P, pop, imgs)
            key = "%s/%s/" % (flavour, fam)
            for k in ("row_bg", "boxes", "vel0", "acc0", "grav"):
                out[key + k] = pop[k]
            for k, v in g.items():
                out[key + k] = v
        # goal-directed cost (Rappids_Simulator main.cpp:95-109) and a tight pyramid budget, on the hard scenes
        pop = scn.rappids_population(N_VEH, seed=11, **FAMILIES["hard"])
        imgs = scn.rappids_render(pop["row_bg"], pop["boxes"], pop["width"])
        g = plan_population(P, pop, imgs, cfg=R.default_cfg(max_pyramids=3, cost_kind=1, cost_vec=(0.5, -0.2, 6.0)))
        for k, v in g.items():
            out["%s/goal3/%s" % (flavour, k)] = v
        cr, qr, pr = pieces(P, cubics, quartics, prim)
        out[flavour + "/pieces/cubic_roots"], out[flavour + "/pieces/quartic_roots"], out[flavour + "/pieces/prim_out"] = cr, qr, pr
    path = os.path.join(HERE, "rappids_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
