"""GPU parity tests of the batched RAPPIDS planner (K6, include/agrifly_b200_rappids.h) through the C ABI:
the parity variant of the kernel against the oracle (bit-exact: flags, counters, pyramids, coefficients), against
the golden vectors recorded from the unmodified reference, the throughput variant within a stated tolerance, the
device rasteriser and sampler, and size-independent properties at BASELINE config 5's full size (64K vehicles)."""
import os

import numpy as np
import pytest

from common import bit_equal
from conftest import ROOT

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "rappids_vectors.npz"))
OUT_INTS = ("found", "best_index", "n_generated", "n_cost_checks", "n_collision_checks", "n_velocity_checks",
            "n_collision_free", "n_pyramids")
HARD = dict(speed_max=4.5, acc_max=3.0, box_depth=(1.0, 3.0), n_boxes=(2, 4))


def oracle(flavour="port-shared"):
    import orc_rappids as R
    if not R.available(flavour):
        pytest.skip("oracle %s not built" % flavour)
    return R, R.Planner(flavour)


def run_gpu(agf, pop, cands, math, **cfg_edits):
    cfg = agf.rappids_cfg(math=math, **cfg_edits)
    n, k = cands.shape[0], cands.shape[1]
    with agf.Rappids(cfg, n, max(k, 1)) as pl:
        pl.render_scenes(pop["row_bg"], pop["boxes"])
        pl.set_states(pop["vel0"], pop["acc0"], pop["grav"])
        pl.set_candidates(cands)
        pl.plan()
        pl.sync()
        return dict(res=pl.results(), flags=pl.candidate_flags(), pyr=pl.pyramids(), stats=pl.stats(),
                    images=pl.get_images(), launches=pl.launch_count)


def assert_vehicle_equal(i, res, flags, pyr, e):
    for f in OUT_INTS:
        assert res[i][f] == e[f], (i, f, res[i][f], e[f])
    assert np.array_equal(flags[i], e["results"]), (i, "flags")
    npyr = e["n_pyramids"]
    assert bit_equal(pyr[i, :npyr], e["pyramids"][:npyr]), (i, "pyramids")
    assert np.all(np.isnan(pyr[i, npyr:])), (i, "unused pyramid records must be NaN")
    if e["found"]:
        assert bit_equal(res[i]["best_coeffs"], e["best_coeffs"]), (i, "coeffs")
        assert res[i]["best_cost"] == e["best_cost"] and res[i]["best_tf"] == e["best_tf"], (i, "cost")


@pytest.mark.parametrize("family", ["easy", "hard"])
def test_parity_variant_is_bit_identical_to_oracle(agf, family):
    """Every output of FindLowestCostTrajectory for every vehicle of a population, candidates drawn by the oracle's
    restatement of the reference's std::mt19937 sampler."""
    R, P = oracle("port-shared")
    n, k = 192, 384
    pop = agf.scenarios.rappids_population(n, seed=31, **(HARD if family == "hard" else {}))
    imgs = agf.scenarios.rappids_render(pop["row_bg"], pop["boxes"], 320)
    ocfg = R.default_cfg(max_pyramids=32)
    exp, cands = [], np.zeros((n, k, 4))
    for i in range(n):
        e = P.plan(ocfg, imgs[i], pop["vel0"][i], pop["acc0"][i], pop["grav"][i], n=k, seed=1000 + i)
        exp.append(e)
        cands[i] = e["candidates"]
    g = run_gpu(agf, pop, cands, agf.abi.MATH_PARITY)
    assert np.array_equal(g["images"], imgs), "device rasteriser differs from the host rasteriser"
    for i in range(n):
        assert_vehicle_equal(i, g["res"], g["flags"], g["pyr"], exp[i])
    st = g["stats"]
    assert st["found"] == sum(e["found"] for e in exp) and st["generated"] == n * k
    assert st["pyramids"] == sum(e["n_pyramids"] for e in exp)
    assert st["collision_free"] == sum(e["n_collision_free"] for e in exp)
    assert g["launches"] >= 1


@pytest.mark.parametrize("fam", ["easy", "hard", "goal3"])
def test_parity_variant_matches_reference_golden(agf, fam):
    """Against outputs of the unmodified reference planner (shared-libm build) committed under tests/golden."""
    key = "ref-shared/%s/" % fam
    sk = "ref-shared/%s/" % ("hard" if fam == "goal3" else fam)
    pop = {k: GOLD[sk + k] for k in ("row_bg", "boxes", "vel0", "acc0", "grav")}
    edits = dict(max_pyramids=3, cost_kind=1, cost_vec=(0.5, -0.2, 6.0)) if fam == "goal3" else {}
    g = run_gpu(agf, pop, GOLD[key + "cands"], agf.abi.MATH_PARITY, **edits)
    n = pop["vel0"].shape[0]
    for i in range(n):
        ints = GOLD[key + "ints"][i]
        e = dict(zip(OUT_INTS, ints.tolist()))
        e.update(results=GOLD[key + "flags"][i], pyramids=GOLD[key + "pyr"][i][:ints[7]], best_coeffs=GOLD[key + "coeffs"][i],
                 best_cost=GOLD[key + "cost"][i], best_tf=GOLD[key + "tf"][i])
        assert_vehicle_equal(i, g["res"], g["flags"], g["pyr"], e)


def test_images_from_host_equal_rendered_scenes_and_per_vehicle_goals(agf):
    """agf_rappids_set_images (the cv::Mat path) and agf_rappids_set_goals (per-vehicle cost vectors)."""
    R, P = oracle("port-shared")
    n, k = 48, 200
    pop = agf.scenarios.rappids_population(n, seed=8, **HARD)
    imgs = agf.scenarios.rappids_render(pop["row_bg"], pop["boxes"], 320)
    cands = agf.scenarios.rappids_candidates(n, k, seed=5)
    rng = np.random.default_rng(2)
    goals = np.column_stack([rng.uniform(-2, 2, n), rng.uniform(-1, 1, n), rng.uniform(4, 9, n)])
    cfg = agf.rappids_cfg(math=agf.abi.MATH_PARITY, cost_kind=1, max_pyramids=8)
    with agf.Rappids(cfg, n, k) as pl:
        pl.set_images(imgs[:20])
        pl.set_images(imgs[20:], first=20)
        pl.set_states(pop["vel0"], pop["acc0"], pop["grav"])
        pl.set_goals(goals)
        pl.set_candidates(cands)
        pl.plan()
        pl.sync()
        res, flags, pyr = pl.results(), pl.candidate_flags(), pl.pyramids()
        assert np.array_equal(pl.get_images(first=7, count=9), imgs[7:16])
        assert bit_equal(pl.get_candidates(), cands)
    for i in range(n):
        e = P.plan(R.default_cfg(max_pyramids=8, cost_kind=1, cost_vec=tuple(goals[i])), imgs[i], pop["vel0"][i],
                   pop["acc0"][i], pop["grav"][i], candidates=cands[i])
        assert_vehicle_equal(i, res, flags, pyr, e)


def test_edge_cases(agf):
    """Far scene (saturated pixels), wall inside the minimum checking distance, vehicle-specific mixes; the planner
    can be called again on the same handle after the inputs changed."""
    R, P = oracle("port-shared")
    n, k = 4, 64
    far = np.full((240, 320), 65535, dtype=np.uint16)
    wall = np.full((240, 320), int(0.3 / (10.0 / 256.0)), dtype=np.uint16)
    half = far.copy()
    half[:, :160] = int(1.2 / (10.0 / 256.0))
    zero = np.zeros((240, 320), dtype=np.uint16)
    imgs = np.stack([far, wall, half, zero])
    v = np.tile([0.0, 0.0, 1.0], (n, 1))
    a = np.zeros((n, 3))
    g = np.tile([0.0, 9.81, 0.0], (n, 1))
    cands = agf.scenarios.rappids_candidates(n, k, seed=9)
    cfg = agf.rappids_cfg(math=agf.abi.MATH_PARITY)
    with agf.Rappids(cfg, n, k) as pl:
        pl.set_images(imgs)
        pl.set_states(v, a, g)
        pl.set_candidates(cands)
        for rep in range(2):
            pl.plan()
            pl.sync()
            res, flags, pyr = pl.results(), pl.candidate_flags(), pl.pyramids()
            for i in range(n):
                e = P.plan(R.default_cfg(max_pyramids=32), imgs[i], v[i], a[i], g[i], candidates=cands[i])
                assert_vehicle_equal(i, res, flags, pyr, e)
            assert res[0]["found"] == 1 and res[1]["found"] == 0 and res[1]["best_index"] == -1
            # second round: new candidates, same images
            cands = agf.scenarios.rappids_candidates(n, k, seed=10)
            pl.set_candidates(cands)


@pytest.mark.parametrize("size", [(640, 480), (1088, 400), (96, 64)])
def test_other_image_sizes(agf, size):
    """Images that are not 320 x 240: twice the size (20 / 15 groups of 32 pixels per line), wider than 1 024 pixels (more
    than 32 groups per row: the one-line-per-step form of the shrink scans), and smaller than three groups.  Synthetic scenes:
    a background whose depth varies by row and by column, boxes in front of it; parity variant against the oracle, bit for bit."""
    R, P = oracle("port-shared")
    W, H = size
    n, k = 6, 160
    rng = np.random.default_rng(W + H)
    scale = 10.0 / 256.0
    imgs = np.zeros((n, H, W), dtype=np.uint16)
    for i in range(n):
        bg = (rng.uniform(3.0, 6.0) - 1.5 * np.linspace(0, 1, H)[:, None] + 0.5 * np.linspace(0, 1, W)[None, :]) / scale
        im = bg.copy()
        for _ in range(rng.integers(1, 4)):
            x0, y0 = rng.integers(0, W - 8), rng.integers(0, H - 8)
            w_, h_ = rng.integers(4, max(5, W // 4)), rng.integers(4, max(5, H // 3))
            im[y0:y0 + h_, x0:x0 + w_] = rng.uniform(0.8, 2.5) / scale
        imgs[i] = np.clip(im, 0, 65535).astype(np.uint16)
    v = np.column_stack([rng.uniform(-0.5, 0.5, n), rng.uniform(-0.5, 0.5, n), rng.uniform(0.0, 1.5, n)])
    a = rng.uniform(-0.5, 0.5, (n, 3))
    g = np.tile([0.0, 9.81, 0.0], (n, 1))
    cands = agf.scenarios.rappids_candidates(n, k, seed=11, width=W, height=H)
    cfg = agf.rappids_cfg(width=W, height=H, math=agf.abi.MATH_PARITY)
    with agf.Rappids(cfg, n, k) as pl:
        pl.set_images(imgs)
        pl.set_states(v, a, g)
        pl.set_candidates(cands)
        pl.plan()
        pl.sync()
        res, flags, pyr = pl.results(), pl.candidate_flags(), pl.pyramids()
    ocfg = R.default_cfg(width=W, height=H, max_pyramids=32)
    npyr = 0
    for i in range(n):
        e = P.plan(ocfg, imgs[i], v[i], a[i], g[i], candidates=cands[i])
        assert_vehicle_equal(i, res, flags, pyr, e)
        npyr += e["n_pyramids"]
    assert npyr > 0 or min(size) < 100


def test_fast_variant_agrees_with_glibc_oracle(agf):
    """Throughput variant (FMA contraction, CUDA libm) against the pure-reference arithmetic.  Stated tolerance:
    at least 97 % of the vehicles get exactly the same flags for all candidates and the same returned candidate;
    for those the cost and coefficients agree to 1e-9 relative.  (A root that moves by an ulp can flip a verdict of
    a candidate that grazes a pyramid face, which then changes the pruning of later candidates.)"""
    R, P = oracle("port-glibc")
    n, k = 256, 256
    pop = agf.scenarios.rappids_population(n, seed=77, **HARD)
    imgs = agf.scenarios.rappids_render(pop["row_bg"], pop["boxes"], 320)
    cands = agf.scenarios.rappids_candidates(n, k, seed=78)
    g = run_gpu(agf, pop, cands, agf.abi.MATH_FAST)
    same = 0
    for i in range(n):
        e = P.plan(R.default_cfg(max_pyramids=32), imgs[i], pop["vel0"][i], pop["acc0"][i], pop["grav"][i], candidates=cands[i])
        if np.array_equal(g["flags"][i], e["results"]) and g["res"][i]["best_index"] == e["best_index"]:
            same += 1
            if e["found"]:
                assert abs(g["res"][i]["best_cost"] - e["best_cost"]) <= 1e-9 * max(1.0, abs(e["best_cost"]))
                assert np.allclose(g["res"][i]["best_coeffs"], e["best_coeffs"], rtol=1e-9, atol=1e-12)
    assert same >= 0.97 * n, same


@pytest.mark.parametrize("family", ["easy", "hard"])
def test_collision_verdicts_against_the_reference_ground_truth(agf, family):
    """A second, algorithm-independent check of K6 on the GPU scenes: the reference's own ray-traced collision test
    (DepthImagePlanner::IsCollisionFreeGroundTruth, DepthImagePlanner.cpp:1031-1097, the yardstick of its
    MeasureConservativeness, :972-1003) run on every candidate that reached the collision test of the THROUGHPUT variant.
    The pyramid method is conservative: what it calls free must be free for the ray tracer (no unsafe accept, ever), and
    what it wrongly calls colliding stays at the reference planner's own rate on the same scenes."""
    import orc_rappids as R
    fl = "ref-glibc" if R.available("ref-glibc") else "port-glibc"
    if not R.available(fl):
        pytest.skip("no planner oracle built")
    P = R.Planner(fl)
    n, k = 32, 256
    pop = agf.scenarios.rappids_population(n, seed=91, **(HARD if family == "hard" else {}))
    imgs = agf.scenarios.rappids_render(pop["row_bg"], pop["boxes"], 320)
    cands = agf.scenarios.rappids_candidates(n, k, seed=92)
    g = run_gpu(agf, pop, cands, agf.abi.MATH_FAST)
    ocfg = R.default_cfg(max_pyramids=32)
    tab, tab_ref = np.zeros((2, 2), int), np.zeros((2, 2), int)  # [planner says free][ground truth says free]
    for i in range(n):
        e = P.plan(ocfg, imgs[i], pop["vel0"][i], pop["acc0"][i], pop["grav"][i], candidates=cands[i])
        chk = np.nonzero((g["flags"][i] & 4) | (e["results"] & 4))[0]
        if len(chk) == 0:
            continue
        gt = P.ground_truth(ocfg, imgs[i], pop["vel0"][i], pop["acc0"][i], pop["grav"][i], cands[i][chk])
        for j, t in zip(chk, gt):
            if g["flags"][i][j] & 4:
                tab[int(bool(g["flags"][i][j] & 8)), int(t)] += 1
            if e["results"][j] & 4:
                tab_ref[int(bool(e["results"][j] & 8)), int(t)] += 1
    print("%s (%s): GPU fast [planner free][truth free] = %s, reference planner = %s" % (family, fl, tab.tolist(), tab_ref.tolist()))
    assert tab.sum() > 100 and tab[1].sum() > 20
    assert tab[1, 0] == 0 and tab_ref[1, 0] == 0              # never free for the planner and colliding for the ray tracer
    assert abs(int(tab[0, 1]) - int(tab_ref[0, 1])) <= 2 + 0.05 * tab_ref[0, 1]   # same conservativeness as the reference


def test_device_sampler(agf):
    """agf_rappids_sample_candidates: draws lie in the sampling box (RandomTrajectoryGenerator, DepthImagePlanner.hpp:
    334-393: uniform pixel, depth, duration; end point = depth * back-projected pixel), are uniform, and do not
    depend on how the population is sharded."""
    n, k = 512, 256
    cfg = agf.rappids_cfg(math=agf.abi.MATH_FAST)
    with agf.Rappids(cfg, n, k) as pl:
        pl.sample_candidates(k, seed=5)
        c = pl.get_candidates()
    with agf.Rappids(cfg, 100, k) as pl:
        pl.sample_candidates(k, seed=5, first_global_index=300)
        c2 = pl.get_candidates()
    assert bit_equal(c2, c[300:400])
    with agf.Rappids(cfg, n, k) as pl:
        pl.sample_candidates(k, seed=6)
        assert not np.array_equal(pl.get_candidates(), c)
    depth, T = c[..., 2], c[..., 3]
    px = c[..., 0] / depth * cfg.focal_length + cfg.cx
    py = c[..., 1] / depth * cfg.focal_length + cfg.cy
    for x, lo, hi in ((px, cfg.sample_min_x, cfg.sample_max_x), (py, cfg.sample_min_y, cfg.sample_max_y),
                      (depth, cfg.sample_min_depth, cfg.sample_max_depth), (T, cfg.sample_min_time, cfg.sample_max_time)):
        assert x.min() >= lo - 1e-9 and x.max() <= hi + 1e-9
        u = (x.ravel() - lo) / (hi - lo)
        m = u.size
        assert abs(u.mean() - 0.5) < 5 * np.sqrt(1 / 12.0 / m) and abs(u.var() - 1 / 12.0) < 5 * np.sqrt(1 / 180.0 / m)
        hist = np.histogram(u, bins=16, range=(0, 1))[0]
        assert np.all(np.abs(hist - m / 16) < 6 * np.sqrt(m / 16))
    # columns are independent
    U = np.stack([px.ravel(), py.ravel(), depth.ravel(), T.ravel()])
    cc = np.corrcoef(U)
    assert np.max(np.abs(cc - np.eye(4))) < 5 / np.sqrt(U.shape[1])


def test_full_size_population_properties(agf):
    """BASELINE config 5 at full size: 64K vehicles, one synthetic 320x240 depth image each (rendered on the
    device), 256 device-sampled candidates.  Size-independent properties + an oracle spot check of 48 random vehicles."""
    R, P = oracle("port-glibc")
    n, k = 65536, 256
    pop = agf.scenarios.rappids_population(n, seed=2024)
    cfg = agf.rappids_cfg(math=agf.abi.MATH_FAST)
    with agf.Rappids(cfg, n, k) as pl:
        pl.render_scenes(pop["row_bg"], pop["boxes"])
        pl.set_states(pop["vel0"], pop["acc0"], pop["grav"])
        pl.sample_candidates(k, seed=7)
        pl.plan()
        pl.sync()
        res, flags, st = pl.results(), pl.candidate_flags(), pl.stats()
        ms, cnt = pl.plan_kernel_time()
        idx = np.random.default_rng(0).choice(n, 48, replace=False)
        spot = [(int(i), pl.get_images(first=int(i), count=1)[0], pl.get_candidates(first=int(i), count=1)[0]) for i in idx]
    print("C5 full size: %.2f ms per plan launch -> %.3e planner calls/s, %.3e candidates/s" % (ms, n / ms * 1e3, n * k / ms * 1e3))
    # counters are the flag counts; flags are nested (a later test is only made if the earlier passed)
    assert np.all(res["n_generated"] == k)
    for bit, name in ((1, "n_cost_checks"), (2, "n_collision_checks"), (4, "n_velocity_checks"), (8, "n_collision_free")):
        assert np.array_equal((flags & bit != 0).sum(axis=1), res[name]), name
    assert np.all(np.isin(flags, [0, 1, 3, 7, 15]))
    assert np.all((res["best_index"] >= 0) == (res["found"] == 1))
    f = res["found"] == 1
    assert np.all(flags[f, res["best_index"][f]] == 15)
    last_free = k - 1 - np.argmax((flags == 15)[:, ::-1], axis=1)
    assert np.array_equal(last_free[f], res["best_index"][f])
    assert np.all(res["n_pyramids"] <= 32) and np.all(res["n_pyramids"][f] >= 1)
    assert np.all(np.isfinite(res["best_cost"][f])) and np.all(res["best_cost"][f] < 0)
    assert st["found"] == f.sum() and st["generated"] == n * k and st["collision_free"] == res["n_collision_free"].sum()
    assert abs(st["sum_best_cost"] - res["best_cost"][f].sum()) < 1e-6 * f.sum()
    assert f.mean() > 0.9
    same = 0
    for i, img, cand in spot:
        e = P.plan(R.default_cfg(max_pyramids=32), img, pop["vel0"][i], pop["acc0"][i], pop["grav"][i], candidates=cand)
        same += int(np.array_equal(flags[i], e["results"]) and res[i]["best_index"] == e["best_index"])
    assert same >= 46, same


def test_planned_primitives_fly_in_the_tracking_loop(agf):
    """N3 -> N1: depth image -> planner -> the returned primitive in SingleAxisTrajectory's own variables
    (agf_rappids_get_tracking_primitives) -> RunTracking inside the step kernel.  The records reproduce the planner's
    polynomial exactly (alpha / 120 == t^5 coefficient, ...), and a population flying its planned primitives matches the
    oracle flying the same records bit for bit."""
    import orc
    from common import cfg_for, make_batch_offboard, make_batch_offboard_ref
    if not orc.available("port-shared"):
        pytest.skip("oracle port not built")
    n, k = 24, 256
    pop = agf.scenarios.rappids_population(n, seed=77)
    with agf.Rappids(agf.rappids_cfg(math=agf.abi.MATH_PARITY), n, k) as pl:
        pl.render_scenes(pop["row_bg"], pop["boxes"])
        pl.set_states(pop["vel0"], pop["acc0"], pop["grav"])
        pl.sample_candidates(k, seed=3)
        pl.plan()
        pl.sync()
        res = pl.results()
        rec = pl.tracking_primitives()
        # the same records handed over on the device: planner -> the batch's trajectory table, no host hop
        sc_dev = agf.scenarios.tracking_scenario(nticks=2600)
        sc_dev["ref"] = dict(sc_dev["ref"], desired_pos=(0.0, 0.0, 2.0), start_us=3000000)
        sc_dev["pos"] = (0.0, 0.0, 0.0)
        sc_dev["primitive"] = None
        bd = make_batch_offboard(agf, dict(sc_dev, targets=[(0, (0.0, 0.0, 0.0))]), n=n)
        pl.export_tracking_primitives(bd, att=np.tile([0.5, -0.5, 0.5, -0.5], (n, 1)), offset=np.tile([0.0, 0.0, 2.0], (n, 1)))
        bd.set_offboard_reference(**sc_dev["ref"])
        bd.run(2600)
        got_dev = bd.record()
        bd.close()
    assert rec.shape == (n, agf.abi.OFFTRAJ_DOUBLES) and np.sum(res["found"]) >= n // 2
    for i in range(n):
        if not res[i]["found"]:
            assert rec[i, 21] == 0.0
            continue
        c = res[i]["best_coeffs"].reshape(6, 3)
        for a in range(3):
            p0, v0, a0, al, be, ga = rec[i, 6 * a:6 * a + 6]
            assert (al / 120, be / 24, ga / 6, a0 / 2, v0, p0) == tuple(c[:, a]), (i, a)
        assert rec[i, 21] == res[i]["best_tf"] and bit_equal(rec[i, 18:21], pop["grav"][i])
    # fly them: camera frame (x right, y down, z forward) -> world (x forward, y left, z up), hover point 2 m up
    s = np.sqrt(0.5)
    cam_to_world = np.array([0.5, -0.5, 0.5, -0.5])  # body x = camera z, body y = -camera x, body z = -camera y
    fly = [i for i in range(n) if res[i]["found"]][:6]
    recs = rec[fly].copy()
    recs[:, 22:26] = cam_to_world
    recs[:, 26:29] = (0.0, 0.0, 2.0)
    sc = agf.scenarios.tracking_scenario(nticks=2600)
    sc["ref"] = dict(sc["ref"], desired_pos=(0.0, 0.0, 2.0), start_us=3000000)
    sc["pos"] = (0.0, 0.0, 0.0)
    b = make_batch_offboard_ref(agf, sc, n=len(fly), primitives=recs)
    b.run(1300)
    b.run(1300)
    got = b.record()
    b.close()
    O = orc.Oracle("port-shared")
    for j in range(len(fly)):
        v = O.vehicle(cfg_for(agf, sc), uwb_comm_period=0.0)
        v.set_state(pos=sc["pos"], att=sc["att"])
        ref = v.run_offboard_ref(2600, agf.offboard_cfg(sc["quad_type"]), agf.offboard_ref(**sc["ref"]), trajectory=recs[j])
        assert bit_equal(got[j], ref[-1]), (j, got[j][0:3], ref[-1][0:3])
        assert bit_equal(got_dev[fly[j]], ref[-1]), ("device hand-over", j)
        assert ref[-1, 35] == 0 and np.linalg.norm(ref[-1, 0:3] - np.array([0.0, 0.0, 2.0])) > 0.3  # it went somewhere


@pytest.mark.parametrize("family", ["easy", "hard"])
def test_frame_jumps_and_dispatch_order_do_not_change_results(agf, family, monkeypatch):
    """The planning pass takes K iterations of InflatePyramid's spiral expansion in one step where their frame holds no
    blocker, folds the unblocked shrink updates of a 32-pixel span into one reduction, and hands vehicles out by their
    previous plan's work: none of it may change a bit of any output.  8 192 vehicles
    (about 30 000 pyramids), both scene families: line-by-line expansion in index order against jumps of 8 and of 3
    iterations and against the second plan of a handle (dispatch by work)."""
    n, k = 8192, 256
    pop = agf.scenarios.rappids_population(n, seed=909, **(HARD if family == "hard" else {}))
    cands = agf.scenarios.rappids_candidates(n, k, seed=910)
    out = {}
    for jump in (0, 8, 3):
        monkeypatch.setenv("AGF_RAPPIDS_FRAME_JUMP", str(jump))
        monkeypatch.setenv("AGF_RAPPIDS_SHRINK_FOLD", "0" if jump == 0 else "1")  # the reference's one update per pixel / folded spans
        with agf.Rappids(agf.rappids_cfg(math=agf.abi.MATH_PARITY), n, k) as pl:
            pl.render_scenes(pop["row_bg"], pop["boxes"])
            pl.set_states(pop["vel0"], pop["acc0"], pop["grav"])
            pl.set_candidates(cands)
            pl.plan()
            pl.sync()
            out[jump] = (pl.results(), pl.candidate_flags(), pl.pyramids())
            if jump == 8:
                work = pl.plan_work()
                pl.plan()  # handed out by the work of the first plan
                pl.sync()
                out["second"] = (pl.results(), pl.candidate_flags(), pl.pyramids())
                assert work.min() > 0
    # the vehicles whose previous plan was long are planned by a whole CTA each (speculative collision tests committed in
    # order): with the threshold at 0.0002 x the launch's ideal length that is most of the population
    monkeypatch.setenv("AGF_RAPPIDS_FRAME_JUMP", "8")
    monkeypatch.setenv("AGF_RAPPIDS_SHRINK_FOLD", "1")
    monkeypatch.setenv("AGF_RAPPIDS_COOP_FACTOR", "0.0002")
    monkeypatch.setenv("AGF_RAPPIDS_COOP_CAP", "1.0")
    with agf.Rappids(agf.rappids_cfg(math=agf.abi.MATH_PARITY), n, k) as pl:
        pl.render_scenes(pop["row_bg"], pop["boxes"])
        pl.set_states(pop["vel0"], pop["acc0"], pop["grav"])
        pl.set_candidates(cands)
        for _ in range(2):
            pl.plan()
        pl.sync()
        out["coop"] = (pl.results(), pl.candidate_flags(), pl.pyramids())
    ref = out[0]
    assert ref[0]["n_pyramids"].sum() > n
    for key in (8, 3, "second", "coop"):
        res, flags, pyr = out[key]
        assert res.tobytes() == ref[0].tobytes(), key
        assert np.array_equal(flags, ref[1]), key
        assert pyr.tobytes() == ref[2].tobytes(), key
