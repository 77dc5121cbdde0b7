"""Population- and trajectory-level parity of the FAST kernels (the ones bench.py times) against the reference.

The bit-exact evidence (tests/test_parity_gpu.py) is for the parity instantiations.  The throughput kernels use
different arithmetic (FMA contraction, packed symmetric EKF, closed forms, CUDA libm), so they can only be held to a
tolerance -- and the closed loop through the float EKF is chaotic enough that the REFERENCE ITSELF does not reproduce
its own trajectories across benign rebuilds.  This file quantifies both, on the same population:

  envelope   the unmodified reference sources (oracle/_ref) rebuilt (a) with FMA contraction (-ffp-contract=fast -mfma,
             what -march=native does: ref-fma) and (b) with a different, equally accurate float libm (ref-shared), each
             against the plain build (ref-glibc), per vehicle, every 50th tick, all 17 plant columns and the estimator
             p^ v^ q^;
  fast       FP32-fast and FP64-fast kernels against ref-glibc, same statistic.

Asserted: at the median, 90th and 99th percentile over vehicles the fast kernels' deviation is at most 2x the larger of
the two reference-vs-reference deviations, over the whole 10 s horizon and over its first 2 s; where the loop is well
conditioned (rates mode: no estimator in the control path) the north star's FP32 tolerance 1e-4 holds outright; and
the set of vehicles that panic is the reference's up to the few whose margin is inside the envelope.  On the GPU box
/root/reference does not exist: oracle/_ref/*.so (built in the build container) travel with the snapshot; without
them the envelope falls back to port-shared vs port-glibc (the restatement, bit-identical to ref-* where both exist).
"""
import os

import numpy as np
import pytest

from common import cfg_for

pytestmark = pytest.mark.gpu

STRIDE = 50
GROUPS = (("position", slice(0, 3)), ("plant17", slice(0, 17)), ("estimator", slice(21, 31)))
WINDOWS = (("0-2s", slice(0, 20)), ("0-10s", slice(0, 100)))
PCTS = (50, 90, 99)


def per_vehicle_err(a, b, cols, window):
    """max over the window's records and the group's columns of |a-b| / max(|b|, 1)  -> [vehicle]"""
    e = np.abs(a[window][..., cols] - b[window][..., cols]) / np.maximum(np.abs(b[window][..., cols]), 1.0)
    return np.nanmax(e, axis=(0, 2))


def table(a, ref, ok):
    return {(g, w): np.percentile(per_vehicle_err(a, ref, cols, win)[ok], PCTS + (100,))
            for g, cols in GROUPS for w, win in WINDOWS}


def gpu_population_traj(agf, cfg, sc, init, slot, sched, nt, **kw):
    n = len(init)
    b = agf.Batch(cfg, n, uwb_comm_period=sc["uwb_comm_period"], **kw)
    for i, p in sc["anchors"]:
        b.add_anchor(i, p)
    b.set_state13(init)
    if slot is not None:
        b.set_slot(0, slot)
    b.set_schedule(sched)
    recs = []
    for _ in range(nt // STRIDE):
        b.run(STRIDE)
        recs.append(b.record())
    b.close()
    return np.stack(recs)


def oracles(orc_mod):
    have_ref = all(orc_mod.available(f) for f in ("ref-glibc", "ref-shared", "ref-fma"))
    if have_ref:
        return "ref-glibc", ("ref-shared", "ref-fma")
    return "port-glibc", ("port-shared",)


def test_fast_kernels_population_trajectory_parity_c2(agf, orc_mod):
    """BASELINE config 2's population at FULL size (4 096 randomized vehicles, full onboard loop, 10 s), noise-free."""
    n, nt = 4096, 5000
    s = agf.scenarios
    sc = s.full_scenario(agf.codec, nticks=nt)
    init = s.monte_carlo_initial_states(n, seed=1234)
    idle = agf.codec.encode_idle(0)
    slot = np.array([np.frombuffer(agf.codec.encode_position(0, (p[0], p[1], 1.5)), np.uint8) for p in init])
    sched = [(d, idle, -1) if sl == -2 else (d, None, 0) for d, _, sl in s.hover_slot_schedule(nt)]
    cfg = cfg_for(agf, sc)
    anchors = np.array([[i, *p] for i, p in sc["anchors"]], np.float32)
    slots = np.zeros((4, n, 23), np.uint8)
    slots[0] = slot
    base, others = oracles(orc_mod)
    cpu = {}
    for fl in (base,) + others:
        cpu[fl], _ = orc_mod.Oracle(fl).run_population_traj(cfg, n, STRIDE, init13=init, anchors=anchors, nticks=nt, sched=sched,
                                                            slot_raw=slots, threads=os.cpu_count() or 1,
                                                            uwb_comm_period=sc["uwb_comm_period"])
    ref = cpu[base]
    # vehicles the reference flies normally for the whole horizon (the rest panic -- |yaw| near 180 deg -- and then slide or
    # tumble: their state is not a tracking result; they are compared as a SET below)
    ok = (ref[-1, :, 35] == 0) & np.isfinite(ref[..., 0:34]).all(axis=(0, 2))
    assert ok.mean() > 0.75
    env = {}
    for fl in others:
        t = table(cpu[fl], ref, ok)
        for k, v in t.items():
            env[k] = np.maximum(env.get(k, 0.0), v)
            print("envelope %-10s %-9s %-5s median %.2e p90 %.2e p99 %.2e max %.2e" % ((fl,) + k + tuple(v)))
    for name, kw in (("fp32-fast", dict(precision=agf.abi.PREC_FP32, math=agf.abi.MATH_FAST)),
                     ("fp64-fast", dict(precision=agf.abi.PREC_FP64, math=agf.abi.MATH_FAST)),
                     ("fp32-fast+hk", dict(precision=agf.abi.PREC_FP32, math=agf.abi.MATH_FAST, telemetry_warnings=True))):
        got = gpu_population_traj(agf, cfg, sc, init, slot, sched, nt, **kw)
        t = table(got, ref, ok)
        for k, v in t.items():
            print("%-19s %-9s %-5s median %.2e p90 %.2e p99 %.2e max %.2e   (x envelope: %.2f %.2f %.2f)" %
                  ((name,) + k + tuple(v) + tuple(v[:3] / np.maximum(env[k][:3], 1e-300))))
        for k, v in t.items():
            # percentile by percentile within 2x of what the reference does to itself; an absolute floor of 2e-6 covers the
            # windows in which the reference builds still agree to the last bit (all vehicles idle on the ground)
            assert np.all(v[:3] <= 2.0 * env[k][:3] + 2e-6), (name, k, v, env[k])
        # the same vehicles panic (allowing the few whose decision sits inside the envelope)
        p_ref, p_got = ref[-1, :, 35] != 0, got[-1, :, 35] != 0
        differ = np.sum(p_ref != p_got)
        differ_env = max(np.sum((cpu[fl][-1, :, 35] != 0) != p_ref) for fl in others)
        print("%s: panicked %d (reference %d), %d vehicles differ (reference rebuilds: %d)" %
              (name, p_got.sum(), p_ref.sum(), differ, differ_env))
        assert differ <= 2 * differ_env + 8
        # and the population still hovers where it should
        fin = got[-1, ok]
        assert np.nanmedian(np.abs(fin[:, 2] - 1.5)) < 0.02


def test_fast_kernels_population_trajectory_parity_rates(agf, orc_mod):
    """Rates mode (no estimator in the control path: the well-conditioned case): 1 024 vehicles with randomized initial
    attitude, thrust 1.05 g and a body-rate doublet, 10 s.  North star: 1e-4 relative in FP32 mode."""
    n, nt = 1024, 5000
    s = agf.scenarios
    sc = s.rates_scenario(agf.codec, nticks=nt)
    init = s.monte_carlo_initial_states(n, seed=77)
    cfg = cfg_for(agf, sc)
    base, others = oracles(orc_mod)
    cpu = {}
    for fl in (base,) + others:
        cpu[fl], _ = orc_mod.Oracle(fl).run_population_traj(cfg, n, STRIDE, init13=init, nticks=nt, sched=sc["sched"],
                                                            threads=os.cpu_count() or 1)
    ref = cpu[base]
    ok = np.isfinite(ref[..., 0:34]).all(axis=(0, 2)) & (ref[-1, :, 35] == 0)
    assert ok.mean() > 0.95
    env = {}
    for fl in others:
        t = table(cpu[fl], ref, ok)
        for k, v in t.items():
            env[k] = np.maximum(env.get(k, 0.0), v)
            print("envelope %-10s %-9s %-5s median %.2e p90 %.2e p99 %.2e max %.2e" % ((fl,) + k + tuple(v)))
    for name, kw, tol in (("fp32-fast", dict(precision=agf.abi.PREC_FP32, math=agf.abi.MATH_FAST), 1e-4),
                          ("fp64-fast", dict(precision=agf.abi.PREC_FP64, math=agf.abi.MATH_FAST), 1e-4)):
        got = gpu_population_traj(agf, cfg, sc, init, None, sc["sched"], nt, **kw)
        t = table(got, ref, ok)
        for k, v in t.items():
            print("%-10s %-9s %-5s median %.2e p90 %.2e p99 %.2e max %.2e" % ((name,) + k + tuple(v)))
        # north star, FP32 mode: position, velocity, attitude, motor speed within 1e-4 relative over 10 s -- every vehicle;
        # the typical vehicle an order of magnitude better (compensated FP32 integration, agf_step.cuh tick())
        assert t[("plant17", "0-10s")][3] <= tol, (name, t[("plant17", "0-10s")])
        assert t[("plant17", "0-10s")][0] <= 2e-5, (name, t[("plant17", "0-10s")])
