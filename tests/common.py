"""Shared helpers for the parity tests: run a scenario on an oracle or on the CUDA batch."""
import numpy as np


def cfg_for(agf, sc, **extra):
    kw = dict(motor_time_const=sc["motor_time_const"], motor_inertia=sc["motor_inertia"])
    kw.update(extra)
    return agf.vehicle_cfg(sc["quad_type"], sc["vehicle_id"], **kw)


def run_oracle(O, agf, sc, nticks=None, record=True, cfg=None, **kw):
    v = O.vehicle(cfg if cfg is not None else cfg_for(agf, sc), uwb_comm_period=sc["uwb_comm_period"], **kw)
    v.set_state(pos=sc["pos"], att=sc["att"])
    for i, p in sc["anchors"]:
        v.add_anchor(i, p)
    tr = v.run(nticks or sc["nticks"], sched=sc["sched"], record=record)
    return tr, v


def make_batch(agf, sc, n=1, cfg=None, **kw):
    b = agf.Batch(cfg if cfg is not None else cfg_for(agf, sc), n, uwb_comm_period=sc["uwb_comm_period"], **kw)
    for i, p in sc["anchors"]:
        b.add_anchor(i, p)
    s13 = np.zeros((n, 13))
    s13[:, 0:3] = sc["pos"]
    s13[:, 6:10] = sc["att"]
    b.set_state13(s13)
    b.set_schedule(sc["sched"])
    return b


def run_batch_traj(b, nticks, every=1):
    """Step the batch `every` ticks at a time and record vehicle 0 after each chunk -> [k][40]."""
    out = []
    done = 0
    while done < nticks:
        c = min(every, nticks - done)
        b.run(c)
        done += c
        out.append(b.record()[0])
    return np.array(out)


def rel_err(a, b, floor=1.0):
    """max |a-b| / max(|b|, floor) -- relative where the quantity is large, absolute near zero."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def bit_equal(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def run_oracle_offboard(O, agf, sc, nticks=None, chunks=None, offset=None, cfg=None, ocfg=None, **kw):
    """Offboard-loop scenario on an oracle; `chunks` splits the run into several calls."""
    v = O.vehicle(cfg if cfg is not None else cfg_for(agf, sc), uwb_comm_period=sc["uwb_comm_period"], **kw)
    v.set_state(pos=sc["pos"], att=sc["att"])
    oc = ocfg if ocfg is not None else agf.offboard_cfg(sc["quad_type"])
    n = nticks or sc["nticks"]
    parts = []
    for c in (chunks or [n]):
        parts.append(v.run_offboard(c, oc, sc["targets"], offset=offset))
    return np.vstack(parts), v


def make_batch_offboard(agf, sc, n=1, offsets=None, cfg=None, ocfg=None, **kw):
    b = agf.Batch(cfg if cfg is not None else cfg_for(agf, sc), n, uwb_comm_period=sc["uwb_comm_period"], **kw)
    s13 = np.zeros((n, 13))
    s13[:, 0:3] = sc["pos"]
    s13[:, 6:10] = sc["att"]
    b.set_state13(s13)
    b.set_offboard_loop(ocfg if ocfg is not None else agf.offboard_cfg(sc["quad_type"]), sc["targets"], offsets)
    return b


def run_oracle_offboard_ref(O, agf, sc, chunks=None, offset=None, nticks=None):
    """Offboard loop with a reference generator (flight stages / primitive tracking) on an oracle."""
    v = O.vehicle(cfg_for(agf, sc), uwb_comm_period=sc["uwb_comm_period"])
    v.set_state(pos=sc["pos"], att=sc["att"])
    oc = agf.offboard_cfg(sc["quad_type"])
    ref = agf.offboard_ref(**sc["ref"])
    rec = None if sc["primitive"] is None else agf.primitive_record(**sc["primitive"])
    n = nticks or sc["nticks"]
    parts = [v.run_offboard_ref(c, oc, ref, offset=offset, trajectory=rec) for c in (chunks or [n])]
    return np.vstack(parts), v


def make_batch_offboard_ref(agf, sc, n=1, offsets=None, primitives=None, **kw):
    b = make_batch_offboard(agf, dict(sc, targets=[(0, (0.0, 0.0, 0.0))]), n=n, offsets=offsets, **kw)
    if sc["primitive"] is not None or primitives is not None:
        recs = primitives if primitives is not None else np.tile(agf.primitive_record(**sc["primitive"]), (n, 1))
        b.set_offboard_trajectories(recs)
    b.set_offboard_reference(**sc["ref"])
    return b


def run_oracle_estimator(O, agf, sc, chunks=None, jump_at=None, est=None):
    """Offboard loop fed by the MocapStateEstimator on an oracle.  sc: offboard / stages / tracking scenario.
    jump_at: tick at which the vehicle is teleported by (3, 0.5, 0) m, so that measurements are rejected until the forced
    reset.  Returns (trajectory, [estimate(0), estimate(0.03), counters])."""
    v = O.vehicle(cfg_for(agf, sc), uwb_comm_period=sc["uwb_comm_period"])
    v.set_state(pos=sc["pos"], att=sc["att"])
    v.set_offboard_estimator(est if est is not None else agf.offboard_estimator())
    oc = agf.offboard_cfg(sc["quad_type"])
    if "ref" in sc:
        ref = agf.offboard_ref(**sc["ref"])
        rec = None if sc["primitive"] is None else agf.primitive_record(**sc["primitive"])
        step = lambda c: v.run_offboard_ref(c, oc, ref, trajectory=rec)
    else:
        step = lambda c: v.run_offboard(c, oc, sc["targets"])
    parts = []
    n = sc["nticks"]
    if jump_at is not None:
        parts.append(step(jump_at))
        a = parts[0][-1]
        v.set_state(pos=a[0:3] + np.array([3.0, 0.5, 0.0]), vel=a[3:6], att=a[6:10], ang_vel=a[10:13])
        n -= jump_at
    for c in (chunks or [n]):
        parts.append(step(c))
    e0, c4 = v.offboard_estimate(0.0)
    e1, _ = v.offboard_estimate(0.03)
    return np.vstack(parts), np.concatenate([e0, e1, c4])


def make_batch_estimator(agf, sc, n=1, est=None, **kw):
    """Batch flying `sc` (offboard / stages / tracking scenario) with the MocapStateEstimator in the loop."""
    if "ref" in sc:
        b = make_batch_offboard_ref(agf, sc, n=n, **kw)
    else:
        b = make_batch_offboard(agf, sc, n=n, **kw)
    b.set_offboard_estimator(est if est is not None else agf.offboard_estimator())
    return b
