"""CPU tests of the RAPPIDS planner oracle: the restatement oracle/port/
agf_rappids_port.cpp against golden vectors recorded from the UNMODIFIED reference planner
(tests/golden/rappids_vectors.npz, made by tests/golden/make_golden_rappids.py), against the live reference
build when oracle/_ref is present, and against independent mathematics (numpy root finding, boundary conditions
of the motion primitive, geometric meaning of the flags)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from common import bit_equal
from conftest import ROOT

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "rappids_vectors.npz"))
OUT_INTS = ("found", "best_index", "n_generated", "n_cost_checks", "n_collision_checks", "n_velocity_checks",
            "n_collision_free", "n_pyramids")
MAX_PYR = 32


def planner_or_skip(flavour):
    import orc_rappids as R
    if not R.available(flavour):
        pytest.skip("oracle flavour %s not built here" % flavour)
    return R.Planner(flavour)


def family_cfg(R, fam):
    if fam == "goal3":
        return R.default_cfg(max_pyramids=3, cost_kind=1, cost_vec=(0.5, -0.2, 6.0))
    return R.default_cfg(max_pyramids=MAX_PYR)


def scene_key(fam):
    return "hard" if fam == "goal3" else fam


def check_against_golden(P, R, agf, math, fam, use_sampler):
    key = "ref-%s/%s/" % (math, fam)
    sk = "ref-%s/%s/" % (math, scene_key(fam))
    imgs = agf.scenarios.rappids_render(GOLD[sk + "row_bg"], GOLD[sk + "boxes"], 320)
    cfg = family_cfg(R, fam)
    n = imgs.shape[0]
    k = GOLD[key + "cands"].shape[1]
    for i in range(n):
        r = P.plan(cfg, imgs[i], GOLD[sk + "vel0"][i], GOLD[sk + "acc0"][i], GOLD[sk + "grav"][i], n=k, seed=i,
                   candidates=None if use_sampler else GOLD[key + "cands"][i], max_pyr=MAX_PYR)
        assert [r[f] for f in OUT_INTS] == GOLD[key + "ints"][i].tolist(), (fam, i)
        assert np.array_equal(r["results"], GOLD[key + "flags"][i]), (fam, i)
        assert bit_equal(r["candidates"], GOLD[key + "cands"][i]), (fam, i)   # std::mt19937 + uniform_real draws
        npyr = r["n_pyramids"]
        assert bit_equal(r["pyramids"][:npyr], GOLD[key + "pyr"][i][:npyr]), (fam, i)
        if r["found"]:
            assert bit_equal(r["best_coeffs"], GOLD[key + "coeffs"][i]), (fam, i)
            assert r["best_cost"] == GOLD[key + "cost"][i] and r["best_tf"] == GOLD[key + "tf"][i]


@pytest.mark.parametrize("math", ["glibc", "shared"])
@pytest.mark.parametrize("fam", ["easy", "hard", "goal3"])
def test_port_matches_reference_golden(agf, math, fam):
    """The restatement reproduces every output of the reference's FindLowestCostTrajectory bit for bit, drawing the
    candidates itself (its own mt19937 / uniform_real_distribution restatement)."""
    import orc_rappids as R
    check_against_golden(planner_or_skip("port-" + math), R, agf, math, fam, use_sampler=True)


@pytest.mark.parametrize("math", ["glibc", "shared"])
def test_reference_golden_still_reproducible(agf, math):
    import orc_rappids as R
    check_against_golden(planner_or_skip("ref-" + math), R, agf, math, "hard", use_sampler=True)
    check_against_golden(planner_or_skip("ref-" + math), R, agf, math, "goal3", use_sampler=False)


@pytest.mark.parametrize("math", ["glibc", "shared"])
def test_port_matches_live_reference_on_fresh_population(agf, math):
    """A population that is NOT in the golden file: port and unmodified reference side by side."""
    import orc_rappids as R
    ref, port = planner_or_skip("ref-" + math), planner_or_skip("port-" + math)
    pop = agf.scenarios.rappids_population(12, seed=4242, speed_max=3.5, acc_max=2.0, box_depth=(1.0, 4.0), n_boxes=(1, 4))
    imgs = agf.scenarios.rappids_render(pop["row_bg"], pop["boxes"], 320)
    samp = R.Sampler(20.0, 300.0, 30.0, 200.0, 1.0, 4.0, 1.0, 2.5)
    for i in range(12):
        for cfg in (R.default_cfg(max_pyramids=0), R.default_cfg(max_pyramids=2, cost_kind=1, cost_vec=(0.0, 0.0, 8.0))):
            a = ref.plan(cfg, imgs[i], pop["vel0"][i], pop["acc0"][i], pop["grav"][i], n=384, seed=100 + i, sampler=samp, max_pyr=64)
            b = port.plan(cfg, imgs[i], pop["vel0"][i], pop["acc0"][i], pop["grav"][i], n=384, seed=100 + i, sampler=samp, max_pyr=64)
            for f in OUT_INTS:
                assert a[f] == b[f], (i, f)
            assert np.array_equal(a["results"], b["results"]) and bit_equal(a["candidates"], b["candidates"])
            assert bit_equal(a["pyramids"], b["pyramids"])
            if a["found"]:
                assert bit_equal(a["best_coeffs"], b["best_coeffs"]) and a["best_cost"] == b["best_cost"]


@pytest.mark.parametrize("math", ["glibc", "shared"])
def test_pieces_match_reference_golden(math):
    """RootFinder (Common/Math/RootFinder.hpp) and RapidTrajectoryGenerator known answers recorded from the reference."""
    P = planner_or_skip("port-" + math)
    for i, c in enumerate(GOLD["pieces/cubics"]):
        n, r = P.solve_cubic(*c)
        g = GOLD["ref-%s/pieces/cubic_roots" % math][i]
        assert n == int(g[0]) and bit_equal(r[:n], g[1:1 + n]), ("cubic", i)
    for i, c in enumerate(GOLD["pieces/quartics"]):
        n, r = P.solve_quartic(*c)
        g = GOLD["ref-%s/pieces/quartic_roots" % math][i]
        assert n == int(g[0]) and bit_equal(r[:n], g[1:1 + n]), ("quartic", i)
    for i, p in enumerate(GOLD["pieces/prim"]):
        abg, ir, vr = P.primitive(p[0:3], p[3:6], p[6:9], p[9:12], p[12])
        g = GOLD["ref-%s/pieces/prim_out" % math][i]
        assert bit_equal(abg.ravel(), g[:9]) and (ir, vr) == (int(g[9]), int(g[10])), ("primitive", i)


def test_root_finder_against_numpy():
    """The reference's closed-form solvers find the real roots numpy finds (to the accuracy the float 2*pi and float
    eps inside them allow: SURVEY appendix A item 16)."""
    P = planner_or_skip("port-glibc")
    rng = np.random.default_rng(3)
    for _ in range(200):
        roots = np.sort(rng.uniform(-5, 5, 4))
        if np.min(np.diff(roots)) < 0.2:
            continue
        co = np.poly(roots)       # monic quartic with four well-separated real roots
        n, r = P.solve_quartic(*co[1:])
        assert n == 4 and np.allclose(np.sort(r), roots, atol=1e-5), (roots, r)
        co3 = np.poly(roots[:3])
        n, r = P.solve_cubic(*co3[1:])
        assert n == 3 and np.allclose(np.sort(r[:3]), roots[:3], atol=1e-5), (roots, r)
    # one real root + complex pair
    n, r = P.solve_cubic(*np.poly([2.0, 1 + 1j, 1 - 1j])[1:].real)
    assert n == 1 and abs(r[0] - 2.0) < 1e-6


def test_primitive_meets_boundary_conditions():
    """RapidTrajectoryGenerator::Generate (RapidTrajectoryGenerator.cpp): the 5th-order primitive starts at
    (0, v0, a0) and ends at rest at the goal."""
    P = planner_or_skip("port-glibc")
    rng = np.random.default_rng(5)
    for _ in range(50):
        v0, a0 = rng.uniform(-2, 2, 3), rng.uniform(-2, 2, 3)
        goal, T = rng.uniform(-1, 1, 3) + [0, 0, 2], rng.uniform(0.5, 3.0)
        abg, _, _ = P.primitive(v0, a0, [0, 9.81, 0], goal, T)
        al, be, ga = abg[:, 0], abg[:, 1], abg[:, 2]
        pos = al / 120 * T**5 + be / 24 * T**4 + ga / 6 * T**3 + a0 / 2 * T**2 + v0 * T
        vel = al / 24 * T**4 + be / 6 * T**3 + ga / 2 * T**2 + a0 * T + v0
        acc = al / 6 * T**3 + be / 2 * T**2 + ga * T + a0
        assert np.allclose(pos, goal, atol=1e-9) and np.allclose(vel, 0, atol=1e-9) and np.allclose(acc, 0, atol=1e-9)


def test_flags_mean_what_they_say(agf):
    """Geometric check of the collision verdicts, independent of the planner's pyramid machinery: a candidate flagged
    COLLISION_FREE never comes closer to an occupied pixel's back-projection than the TRUE vehicle radius (sampled
    along the trajectory), and the counters are consistent with the flags."""
    import orc_rappids as R
    P = planner_or_skip("port-glibc")
    pop = agf.scenarios.rappids_population(6, seed=77, speed_max=3.0, box_depth=(1.2, 3.0), n_boxes=(2, 4))
    imgs = agf.scenarios.rappids_render(pop["row_bg"], pop["boxes"], 320)
    cfg = R.default_cfg(max_pyramids=MAX_PYR)
    f, cx, cy, ds = cfg.focal_length, cfg.cx, cfg.cy, cfg.depth_scale
    ys, xs = np.mgrid[0:240, 0:320]
    n_free = 0
    for i in range(6):
        r = P.plan(cfg, imgs[i], pop["vel0"][i], pop["acc0"][i], pop["grav"][i], n=400, seed=i)
        fl = r["results"]
        assert r["n_generated"] == 400
        assert r["n_cost_checks"] == int(np.sum(fl & R.LOW_COST != 0))
        assert r["n_collision_checks"] == int(np.sum(fl & R.DYN_FEASIBLE != 0))
        assert r["n_velocity_checks"] == int(np.sum(fl & R.VEL_ADMISSIBLE != 0))
        assert r["n_collision_free"] == int(np.sum(fl & R.COLLISION_FREE != 0))
        assert (r["best_index"] >= 0) == bool(r["found"])
        if r["found"]:
            assert fl[r["best_index"]] == 15
            # the returned trajectory is the last collision-free one (each accepted candidate lowers the bar)
            assert r["best_index"] == int(np.nonzero(fl == 15)[0][-1])
        z = imgs[i].astype(np.float64) * ds
        pts = np.stack([(xs - cx) / f * z, (ys - cy) / f * z, z], axis=-1).reshape(-1, 3)[::7]  # occupied surface samples
        for k in np.nonzero(fl == 15)[0][:3]:
            goal, T = r["candidates"][k, :3], r["candidates"][k, 3]
            abg, _, _ = P.primitive(pop["vel0"][i], pop["acc0"][i], pop["grav"][i], goal, T)
            t = np.linspace(0, T, 60)[:, None]
            p = (abg[:, 0] / 120 * t**5 + abg[:, 1] / 24 * t**4 + abg[:, 2] / 6 * t**3 + pop["acc0"][i] / 2 * t**2 + pop["vel0"][i] * t)
            d = np.min(np.linalg.norm(p[:, None, :] - pts[None, :, :], axis=2))
            assert d > cfg.true_radius, (i, k, d)
            n_free += 1
    assert n_free >= 6


def test_ground_truth_collision_check_port_matches_reference_and_planner_is_conservative(agf):
    """DepthImagePlanner::IsCollisionFreeGroundTruth (DepthImagePlanner.cpp:1031-1097): the restatement equals the
    unmodified reference on every candidate that reached the collision test, and the pyramid method never calls a
    trajectory free that the ray tracer finds colliding (Section IV.A of the RAPPIDS paper, MeasureConservativeness)."""
    import orc_rappids as R
    if not R.available("ref-glibc"):
        pytest.skip("oracle/_ref not built")
    ref, port = R.Planner("ref-glibc"), R.Planner("port-glibc")
    n, k = 6, 192
    for fam, kw in (("easy", {}), ("hard", dict(speed_max=4.5, acc_max=3.0, box_depth=(1.0, 3.0), n_boxes=(2, 4)))):
        pop = agf.scenarios.rappids_population(n, seed=55, **kw)
        imgs = agf.scenarios.rappids_render(pop["row_bg"], pop["boxes"], 320)
        cfg = R.default_cfg(max_pyramids=32)
        checked = unsafe = 0
        for i in range(n):
            e = ref.plan(cfg, imgs[i], pop["vel0"][i], pop["acc0"][i], pop["grav"][i], n=k, seed=900 + i)
            chk = np.nonzero(e["results"] & 4)[0]
            if len(chk) == 0:
                continue
            args = (cfg, imgs[i], pop["vel0"][i], pop["acc0"][i], pop["grav"][i], e["candidates"][chk])
            g_ref, g_port = ref.ground_truth(*args), port.ground_truth(*args)
            assert np.array_equal(g_ref, g_port), (fam, i)
            checked += len(chk)
            unsafe += int(np.sum(((e["results"][chk] & 8) != 0) & ~g_ref))
        assert checked > 20 and unsafe == 0, (fam, checked, unsafe)


def test_edge_cases(agf):
    """Empty candidate list, a wall closer than the minimum checking distance, an empty (far) scene, saturated pixels."""
    import orc_rappids as R
    P = planner_or_skip("port-glibc")
    cfg = R.default_cfg(max_pyramids=MAX_PYR)
    v, a, g = [0.0, 0.0, 1.0], [0.0, 0.0, 0.0], [0.0, 9.81, 0.0]
    far = np.full((240, 320), 65535, dtype=np.uint16)
    r = P.plan(cfg, far, v, a, g, n=64, seed=1)
    assert r["found"] == 1 and r["n_pyramids"] >= 1 and r["n_collision_free"] >= 1
    wall = np.full((240, 320), int(0.3 / cfg.depth_scale), dtype=np.uint16)   # 0.3 m < min_checking_dist 0.5 m
    r = P.plan(cfg, wall, v, a, g, n=64, seed=1)
    assert r["found"] == 0 and r["best_index"] == -1 and r["n_collision_free"] == 0
    cands = np.zeros((0, 4))
    r = P.plan(cfg, far, v, a, g, candidates=cands)
    assert r["found"] == 0 and r["n_generated"] == 0
    if R.available("ref-glibc"):
        ref = R.Planner("ref-glibc")
        for img in (far, wall):
            x, y = ref.plan(cfg, img, v, a, g, n=64, seed=1), P.plan(cfg, img, v, a, g, n=64, seed=1)
            assert all(x[f] == y[f] for f in OUT_INTS) and np.array_equal(x["results"], y["results"])


def test_plan_many_threads_equal_single_calls(agf):
    import orc_rappids as R
    P = planner_or_skip("port-glibc")
    pop = agf.scenarios.rappids_population(10, seed=5)
    imgs = agf.scenarios.rappids_render(pop["row_bg"], pop["boxes"], 320)
    cands = agf.scenarios.rappids_candidates(10, 128, seed=3)
    cfg = R.default_cfg(max_pyramids=MAX_PYR)
    outs, res = P.plan_many(cfg, imgs, pop["vel0"], pop["acc0"], pop["grav"], cands, threads=3)
    for i in range(10):
        r = P.plan(cfg, imgs[i], pop["vel0"][i], pop["acc0"][i], pop["grav"][i], candidates=cands[i])
        assert outs[i].found == r["found"] and outs[i].best_index == r["best_index"] and np.array_equal(res[i], r["results"])


def test_rappids_struct_sizes_match_header(agf, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "agrifly_b200_rappids.h"\nint main(void){printf("%zu %zu\\n",'
                   'sizeof(agf_rappids_cfg),sizeof(agf_rappids_result));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(agf.abi.RappidsCfg), C.sizeof(agf.abi.RappidsResult)]
    c = agf.rappids_cfg()
    assert (c.width, c.height, c.focal_length, c.cx, c.cy) == (320, 240, 160.0, 160.0, 120.0)
    assert (c.true_radius, c.planning_radius, c.min_checking_dist) == (0.116, 0.174, 0.5)
    assert c.depth_scale == 10.0 / 256.0


def test_rappids_has_no_cpu_fallback(agf):
    from conftest import has_cuda
    if has_cuda():
        pytest.skip("GPU present")
    h = C.c_void_p()
    cfg = agf.rappids_cfg()
    rc = agf.lib().agf_rappids_create(C.byref(cfg), 4, 16, C.byref(h))
    assert rc == agf.abi.ENODEVICE and not h
