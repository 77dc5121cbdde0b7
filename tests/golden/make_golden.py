"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference
by oracle/Makefile).  Run in the build container:  python tests/golden/make_golden.py

The reference has no golden vectors of its own for this path; these fixtures
are outputs of the reference itself, recorded so that the oracle port -- and through it the CUDA
path -- stays pinned on machines where /root/reference does not exist.
Contents per scenario and libm flavour: the 40-column trajectory record (oracle/oracle_api.h) at
every 25th tick plus the last 20 ticks, and the full internal state at the end.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import agrifly_b200 as agf  # noqa: E402
import orc  # noqa: E402
from agrifly_b200 import scenarios as scen  # noqa: E402
from common import run_oracle, run_oracle_estimator, run_oracle_offboard, run_oracle_offboard_ref  # noqa: E402


def sample_ticks(n):
    return np.unique(np.concatenate([np.arange(24, n, 25), np.arange(max(0, n - 20), n)]))


def main():
    codec = agf.codec
    out = {}
    for flavour in ("ref-glibc", "ref-shared"):
        O = orc.Oracle(flavour)
        for sc in (scen.rates_scenario(codec), scen.full_scenario(codec), scen.accel_scenario(codec)):
            tr, v = run_oracle(O, agf, sc)
            idx = sample_ticks(len(tr))
            key = "%s/%s" % (flavour, sc["name"])
            out[key + "/ticks"] = idx
            out[key + "/traj"] = tr[idx]
            full = v.full()
            for k, val in full.items():
                out[key + "/full/" + k] = np.asarray(val)
            p1, p2 = v.telemetry()
            out[key + "/tel1"] = p1
            out[key + "/tel2"] = p2
        # offboard loop (Rappids_Simulator's closed loop, truth-fed): reference QuadcopterController + CommunicationsDelay
        sc = scen.offboard_scenario()
        tr, v = run_oracle_offboard(O, agf, sc)
        idx = sample_ticks(len(tr))
        key = "%s/%s" % (flavour, sc["name"])
        out[key + "/ticks"] = idx
        out[key + "/traj"] = tr[idx]
        for k, val in v.full().items():
            out[key + "/full/" + k] = np.asarray(val)
        # reference generators of the offboard loop: flight stages (ROS rates-control node) and primitive tracking
        for sc in [scen.stages_scenario(t) for t in range(6)] + [scen.tracking_scenario()]:
            tr, v = run_oracle_offboard_ref(O, agf, sc)
            idx = sample_ticks(len(tr))
            key = "%s/%s" % (flavour, sc["name"])
            out[key + "/ticks"] = idx
            out[key + "/traj"] = tr[idx]
            out[key + "/offstate"] = v.offboard_state()
        # the UNMODIFIED flight-stage state machine of the ROS rates-control node in the loop (roscpp shim): pins the stage logic
        for sc in (scen.stages_scenario(3), scen.stages_emergency_scenario()):
            v = O.vehicle(agf.vehicle_cfg(sc["quad_type"], sc["vehicle_id"], motor_time_const=sc["motor_time_const"],
                                          motor_inertia=sc["motor_inertia"]), uwb_comm_period=0.0)
            v.set_state(pos=sc["pos"], att=sc["att"])
            tr = v.run_stages_node(sc["nticks"], agf.offboard_cfg(sc["quad_type"]), agf.offboard_ref(**sc["ref"]), agf.offboard_estimator())
            idx = sample_ticks(len(tr))
            key = "%s/node/%s" % (flavour, sc["name"])
            out[key + "/ticks"] = idx
            out[key + "/traj"] = tr[idx]
            out[key + "/offstate"] = v.stages_node_state()
        # the offboard loop fed by the reference's MocapStateEstimator (+ measurement rejection and forced reset)
        for sc, jump in ((scen.offboard_scenario(), None), (scen.stages_scenario(1), None), (scen.tracking_scenario(), None),
                         (scen.offboard_scenario(2000), 1000)):
            tr, est = run_oracle_estimator(O, agf, sc, jump_at=jump)
            idx = sample_ticks(len(tr))
            key = "%s/est/%s%s" % (flavour, sc["name"], "" if jump is None else "-jump")
            out[key + "/ticks"] = idx
            out[key + "/traj"] = tr[idx]
            out[key + "/estimate"] = est
    # codec known-answer vectors from the reference's own RadioTypes / TelemetryPacket code
    O = orc.Oracle("ref-glibc")
    rng = np.random.default_rng(7)
    import ctypes as C
    vals = np.concatenate([rng.uniform(-40, 40, (200, 10)), np.array([[0] * 10, [35] * 10, [-35] * 10, [20] * 10,
                          [19.9997] * 10, [np.nan] * 10, [1e-4] * 10, [-1e-4] * 10])]).astype(np.float32)
    raws = {"rates": [], "position": [], "acceleration": []}
    decs = {"rates": [], "position": [], "acceleration": []}
    for row in vals:
        f3 = lambda a: (C.c_float * 3)(*[float(x) for x in a])
        for kind in raws:
            raw = (C.c_uint8 * 23)()
            if kind == "rates":
                O.L.orc_radio_encode_rates(3, float(row[0]), f3(row[1:4]), raw)
            elif kind == "position":
                O.L.orc_radio_encode_position(1, f3(row[0:3]), f3(row[3:6]), f3(row[6:9]), raw)
            else:
                O.L.orc_radio_encode_acceleration(2, f3(row[0:3]), float(row[3]), raw)
            t, fl, fo = C.c_uint8(), C.c_uint8(), (C.c_float * 10)()
            O.L.orc_radio_decode(raw, C.byref(t), C.byref(fl), fo)
            raws[kind].append(np.frombuffer(bytes(raw), np.uint8))
            d = np.array(fo, np.float32)
            if kind == "position":
                d[9] = 0
            if kind == "acceleration":
                d[4:] = 0
            decs[kind].append(d)
    out["codec/values"] = vals
    for kind in raws:
        out["codec/%s/raw" % kind] = np.array(raws[kind])
        out["codec/%s/decoded" % kind] = np.array(decs[kind])
    # airframe tables
    for t in (1, 2, 4, 5):
        lc = agf.abi.LogicConsts()
        O.L.orc_logic_consts(t, C.byref(lc))
        out["consts/%d" % t] = np.frombuffer(bytes(lc), np.uint8)
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
