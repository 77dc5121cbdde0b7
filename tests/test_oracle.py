"""CPU tests of the oracle: the literal port (oracle/port) against golden vectors recorded from the
unmodified reference (tests/golden/reference_vectors.npz, made by tests/golden/make_golden.py), against
the live reference build when oracle/_ref is present, and against the host-compiled product step."""
import os
import subprocess

import numpy as np
import pytest

from common import bit_equal, run_oracle, run_oracle_estimator, run_oracle_offboard, run_oracle_offboard_ref
from conftest import ROOT, oracle_or_skip

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
SCENARIOS = ["rates", "full", "accel"]


def scenario(agf, name):
    s = agf.scenarios
    return {"rates": s.rates_scenario, "full": s.full_scenario, "accel": s.accel_scenario}[name](agf.codec)


FULL_SKIP = {"radio_floats"}  # floats beyond those a message type decodes are uninitialised in the reference


@pytest.mark.parametrize("math", ["glibc", "shared"])
@pytest.mark.parametrize("name", SCENARIOS)
def test_port_matches_reference_golden(agf, orc_mod, math, name):
    O = oracle_or_skip(orc_mod, "port-" + math)
    tr, v = run_oracle(O, agf, scenario(agf, name))
    key = "ref-%s/%s" % (math, name)
    idx = GOLD[key + "/ticks"]
    assert bit_equal(tr[idx], GOLD[key + "/traj"]), "port trajectory differs from the reference's"
    full = v.full()
    for k, val in full.items():
        if k in FULL_SKIP:
            continue
        assert bit_equal(np.asarray(val), GOLD[key + "/full/" + k]), k
    p1, p2 = v.telemetry()
    assert np.array_equal(p1, GOLD[key + "/tel1"]) and np.array_equal(p2, GOLD[key + "/tel2"])


@pytest.mark.parametrize("math", ["glibc", "shared"])
@pytest.mark.parametrize("name", SCENARIOS)
def test_port_matches_live_reference(agf, orc_mod, math, name):
    R = oracle_or_skip(orc_mod, "ref-" + math)
    P = oracle_or_skip(orc_mod, "port-" + math)
    sc = scenario(agf, name)
    a, _ = run_oracle(R, agf, sc)
    b, _ = run_oracle(P, agf, sc)
    assert bit_equal(a, b)


@pytest.mark.parametrize("math", ["glibc", "shared"])
def test_offboard_loop_port_matches_reference(agf, orc_mod, math):
    """the offboard rates loop (reference QuadcopterController::Run + CreateRatesCommand +
    CommunicationsDelay around Run(), truth-fed) restated in the port: golden vectors, and the live reference."""
    sc = agf.scenarios.offboard_scenario()
    P = oracle_or_skip(orc_mod, "port-" + math)
    tr, v = run_oracle_offboard(P, agf, sc, chunks=[1777, 2223])  # the loop's stopwatch and queue persist across calls
    key = "ref-%s/offboard" % math
    assert bit_equal(tr[GOLD[key + "/ticks"]], GOLD[key + "/traj"])
    full = v.full()
    for k, val in full.items():
        if k not in FULL_SKIP:
            assert bit_equal(np.asarray(val), GOLD[key + "/full/" + k]), k
    # the loop really flies the vehicle: set-points reached, no panic
    assert np.linalg.norm(tr[2999, 0:3] - np.array([1.0, -0.5, 2.5])) < 0.05 and tr[-1, 35] == 0
    if orc_mod.available("ref-" + math):
        R = orc_mod.Oracle("ref-" + math)
        a, _ = run_oracle_offboard(R, agf, sc)
        assert bit_equal(a, tr)


def test_offboard_loop_device_step_on_host_matches_port(agf, orc_mod, port_shared):
    """The product's own offboard loop (agf_step.cuh tick(): clock-driven delivery/generation, per-vehicle command
    queue) compiled for the host equals the port bit for bit, also when the run is cut into launches."""
    if not orc_mod.available("hostsim-shared"):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "hostsim"], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("hostsim not buildable here: " + r.stderr[-300:])
    H = orc_mod.Oracle("hostsim-shared")
    sc = agf.scenarios.offboard_scenario(nticks=2500)
    off = (0.3, -0.2, 0.1)
    a, _ = run_oracle_offboard(port_shared, agf, sc, offset=off)
    b, _ = run_oracle_offboard(H, agf, sc, chunks=[1, 1, 13, 985, 1500], offset=off)
    assert bit_equal(a, b)


REF_SCENARIOS = ["stages0", "stages1", "stages2", "stages3", "stages4", "stages5", "tracking"]


def ref_scenario(agf, name):
    return agf.scenarios.tracking_scenario() if name == "tracking" else agf.scenarios.stages_scenario(int(name[-1]))


@pytest.mark.parametrize("math", ["glibc", "shared"])
@pytest.mark.parametrize("name", REF_SCENARIOS)
def test_offboard_reference_generators_port_matches_reference(agf, orc_mod, math, name):
    """flight stages of the ROS rates-control node (spool-up, take-off ramp, the six flight
    trajectories, landing, idle) and Rappids_Simulator's primitive tracking (RunTracking + GetThrust/GetOmega
    feed-forward) restated in the port: golden vectors from the reference, and the live reference where present."""
    sc = ref_scenario(agf, name)
    P = oracle_or_skip(orc_mod, "port-" + math)
    tr, v = run_oracle_offboard_ref(P, agf, sc, chunks=[1234, sc["nticks"] - 1234])
    key = "ref-%s/%s" % (math, name)
    assert bit_equal(tr[GOLD[key + "/ticks"]], GOLD[key + "/traj"])
    assert bit_equal(v.offboard_state(), GOLD[key + "/offstate"])
    assert tr[-1, 35] == 0  # no panic
    if name == "tracking":  # the primitive ends 1.5 m ahead in a frame yawed by 0.4 rad, 0.3 m up
        end = np.array([0.2, -0.1, 2.0]) + np.array([1.5 * np.cos(0.4) - 0.5 * np.sin(0.4), 1.5 * np.sin(0.4) + 0.5 * np.cos(0.4), 0.3])
        assert np.linalg.norm(tr[-1, 0:3] - end) < 0.1
    else:  # took off, flew at 1 m, landed, idles
        assert abs(tr[2500, 2] - 1.0) < (0.55 if name == "stages4" else 0.05)  # trajectory 4 oscillates in height
        assert tr[-1, 2] < 0.02 and v.offboard_state()[0] == agf.abi.STAGE_COMPLETE
    if orc_mod.available("ref-" + math):
        a, _ = run_oracle_offboard_ref(orc_mod.Oracle("ref-" + math), agf, sc)
        assert bit_equal(a, tr)


def test_offboard_reference_generators_device_code_on_host(agf, orc_mod, port_shared):
    """The product's device code for the generators and the tracking controller (agf_step.cuh offboard_generate),
    compiled for the host, equals the port bit for bit, also across launch boundaries and with a set-point offset."""
    if not orc_mod.available("hostsim-shared"):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "hostsim"], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("hostsim not buildable here: " + r.stderr[-300:])
    H = orc_mod.Oracle("hostsim-shared")
    if not hasattr(H.L, "orc_run_offboard_ref"):
        pytest.skip("stale hostsim build")
    for name in ("stages1", "stages4", "tracking", "emergency"):
        sc = agf.scenarios.stages_emergency_scenario() if name == "emergency" else ref_scenario(agf, name)
        off = (0.3, -0.2, 0.1)
        a, va = run_oracle_offboard_ref(port_shared, agf, sc, offset=off)
        b, vb = run_oracle_offboard_ref(H, agf, sc, chunks=[1, 1, 13, 985, sc["nticks"] - 1000], offset=off)
        assert bit_equal(a, b), name
        if name != "tracking":
            assert bit_equal(va.offboard_state(), vb.offboard_state())


EST_CASES = [("offboard", None), ("stages1", None), ("tracking", None), ("offboard", 1000)]


def est_scenario(agf, name, jump):
    if name == "offboard":
        return agf.scenarios.offboard_scenario(2000 if jump else 4000)
    return ref_scenario(agf, name)


@pytest.mark.parametrize("math", ["glibc", "shared"])
@pytest.mark.parametrize("name,jump", EST_CASES)
def test_offboard_estimator_port_matches_reference(agf, orc_mod, math, name, jump):
    """Offboard::MocapStateEstimator in the loop (200 Hz pose measurements, prediction through the queued
    commands, GetPrediction(30 ms) feeding the controller, SetPredictedValues after each command), restated in the
    port: golden vectors from the reference incl. the estimates and counters, and the live reference where present.
    The jump case teleports the vehicle: ten measurements are rejected, then the estimator resets and re-initialises."""
    sc = est_scenario(agf, name, jump)
    P = oracle_or_skip(orc_mod, "port-" + math)
    tr, est = run_oracle_estimator(P, agf, sc, jump_at=jump, chunks=None if jump else [777, sc["nticks"] - 777])
    key = "ref-%s/est/%s%s" % (math, sc["name"], "" if jump is None else "-jump")
    assert bit_equal(tr[GOLD[key + "/ticks"]], GOLD[key + "/traj"])
    assert bit_equal(est, GOLD[key + "/estimate"])
    assert tr[-1, 35] == 0 and est[26] == 1 and est[27] == (10 if jump else 0)
    # the estimate is the state: prediction over the loop latency within centimetres of the truth
    assert np.linalg.norm(est[0:3] - tr[-1, 0:3]) < 0.02
    if orc_mod.available("ref-" + math):
        a, ea = run_oracle_estimator(orc_mod.Oracle("ref-" + math), agf, sc, jump_at=jump)
        assert bit_equal(a, tr) and bit_equal(ea, est)


@pytest.mark.parametrize("case", ["nominal", "emergency"])
@pytest.mark.parametrize("math", ["glibc", "shared"])
def test_stage_logic_is_pinned_by_the_unmodified_ros_state_machine(agf, orc_mod, math, case):
    """AIFS_ROS/hiperlab_rostools/src/QuadMocapRatesControl/ExampleVehicleStateMachine.cpp compiles UNMODIFIED
    into oracle/_ref against a roscpp / message shim; driven in the same loop (its own MocapStateEstimator, Run() at 100 Hz,
    the radio_command it publishes through the delay queue) it flies wait -> spool-up -> take-off -> circle -> landing ->
    idle.  The restated stage logic (port; and through it the device code) reproduces that trajectory and the machine's
    final state bit for bit: golden vectors from the node, and the live node where the reference build exists.  The
    emergency case switches the node's SafetyNet to a set-point outside its lab-space box: the stage machine latches
    StageEmergency during take-off and kills the vehicle."""
    sc = agf.scenarios.stages_scenario(3) if case == "nominal" else agf.scenarios.stages_emergency_scenario()
    P = oracle_or_skip(orc_mod, "port-" + math)
    v = P.vehicle(agf.vehicle_cfg(sc["quad_type"], sc["vehicle_id"], motor_time_const=sc["motor_time_const"],
                                  motor_inertia=sc["motor_inertia"]), uwb_comm_period=0.0)
    v.set_state(pos=sc["pos"], att=sc["att"])
    v.set_offboard_estimator(agf.offboard_estimator())
    tr = v.run_offboard_ref(sc["nticks"], agf.offboard_cfg(sc["quad_type"]), agf.offboard_ref(**sc["ref"]))
    key = "ref-%s/node/%s" % (math, sc["name"])
    assert bit_equal(tr[GOLD[key + "/ticks"]], GOLD[key + "/traj"])
    assert bit_equal(v.offboard_state(), GOLD[key + "/offstate"])
    if case == "nominal":
        assert v.offboard_state()[0] == agf.abi.STAGE_COMPLETE and tr[-1, 35] == 0 and abs(tr[2500, 2] - 1.0) < 0.05
    else:
        assert v.offboard_state()[0] == agf.abi.STAGE_EMERGENCY and tr[-1, 34] == agf.abi.FS_KILLED and tr[-1, 2] < 0.01
    if orc_mod.available("ref-" + math):
        R = orc_mod.Oracle("ref-" + math)
        if hasattr(R.L, "orc_run_stages_node"):
            n = R.vehicle(agf.vehicle_cfg(sc["quad_type"], sc["vehicle_id"], motor_time_const=sc["motor_time_const"],
                                          motor_inertia=sc["motor_inertia"]), uwb_comm_period=0.0)
            n.set_state(pos=sc["pos"], att=sc["att"])
            tn = n.run_stages_node(sc["nticks"], agf.offboard_cfg(sc["quad_type"]), agf.offboard_ref(**sc["ref"]), agf.offboard_estimator())
            assert bit_equal(tn, tr) and bit_equal(n.stages_node_state(), v.offboard_state())


def test_offboard_estimator_device_code_on_host(agf, orc_mod, port_shared):
    """The product's device code of the estimator (agf_step.cuh mocap_update / mocap_predict / mocap_set_predicted),
    compiled for the host, equals the port bit for bit, across launch boundaries and through the reset path."""
    if not orc_mod.available("hostsim-shared"):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "hostsim"], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("hostsim not buildable here: " + r.stderr[-300:])
    H = orc_mod.Oracle("hostsim-shared")
    if not hasattr(H.L, "orc_set_offboard_estimator"):
        pytest.skip("stale hostsim build")
    for name, jump in EST_CASES:
        sc = est_scenario(agf, name, jump)
        a, ea = run_oracle_estimator(port_shared, agf, sc, jump_at=jump)
        b, eb = run_oracle_estimator(H, agf, sc, jump_at=jump, chunks=None if jump else [1, 13, 985, sc["nticks"] - 999])
        assert bit_equal(a, b) and bit_equal(ea, eb), (name, jump)


def test_simulation_csv_rows_match_reference_stream_output(agf, orc_mod):
    """rows of Rappids_Simulator's simulation.csv.  The product's formatter against the reference's own
    stream operators, Rotation::ToEulerYPR and Vec3 types (oracle/_ref harness), character for character, on records
    taken from a flight with the estimator in the loop and on awkward values; the header against main.cpp:266-270."""
    import ctypes as C
    hdr = agf.csv_header()
    assert hdr.startswith("t,posx,posy,posz,velx,vely,velz,attY,attP,attR,angvelx") and hdr.endswith("panic,r1,r2,r3,r4\n")
    assert hdr.count(",") == 39
    R = oracle_or_skip(orc_mod, "ref-glibc")
    sc = agf.scenarios.offboard_scenario(1500)
    v = R.vehicle(agf.vehicle_cfg(sc["quad_type"], sc["vehicle_id"], motor_time_const=sc["motor_time_const"]), uwb_comm_period=0.0)
    v.set_state(pos=sc["pos"], att=sc["att"])
    v.set_offboard_estimator(agf.offboard_estimator())
    rng = np.random.default_rng(5)
    rows = 0
    for k in range(60):
        tr = v.run_offboard(25, agf.offboard_cfg(sc["quad_type"]), sc["targets"])[-1]
        e, _ = v.offboard_estimate(0.0)
        p1, p2 = v.telemetry()
        t1, t2 = agf.abi.Telemetry(), agf.abi.Telemetry()
        agf.lib().agf_telemetry_decode(p1.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(t1))
        agf.lib().agf_telemetry_decode(p2.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(t2))
        text, rec = agf.csv_row(0.05 * (k + 1), tr[0:3], tr[3:6], tr[6:10], tr[10:13], list(t1.motor_forces), e[0:3], e[3:6], e[6:10],
                                e[10:13], (1.0, -0.5, 2.5), (0, 0, 0), t2.panic_reason, rng.normal(size=4) * 10)
        buf = C.create_string_buffer(2048)
        R.L.orc_csv_row(C.byref(rec), buf, len(buf))
        assert text == buf.value.decode(), (text, buf.value.decode())
        assert text.count(",") == 40 and text.endswith(",\n")
        rows += 1
    for vals in ([0.0] * 3, [1e-7, -1e-7, 123456789.0], [1e21, -1e-21, 0.1], [float("nan"), float("inf"), -float("inf")],
                 [100000.0, 1000000.0, 999999.5]):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        text, rec = agf.csv_row(vals[0], vals, vals[::-1], q, vals, vals + [0.5], vals, vals, q, vals, vals, vals, 3, vals + [1.5])
        buf = C.create_string_buffer(2048)
        R.L.orc_csv_row(C.byref(rec), buf, len(buf))
        assert text == buf.value.decode(), (text, buf.value.decode())
    assert rows == 60


def test_ros_message_fields_match_the_simulator_node_statements(agf, orc_mod):
    """the fields of the hiperlab_rostools telemetry and simulator_truth messages.  agf_msg_*_fill against the
    Simulator node's own statements (Simulator/main.cpp:455-475, 501-546) executed on the reference's TelemetryPacket /
    Rotation types in the harness, byte for byte, on packets and states taken from a flight."""
    import ctypes as C
    R = oracle_or_skip(orc_mod, "ref-glibc")
    sc = agf.scenarios.full_scenario(agf.codec, nticks=1500)
    v = R.vehicle(agf.vehicle_cfg(sc["quad_type"], sc["vehicle_id"], motor_time_const=sc["motor_time_const"]),
                  uwb_comm_period=sc["uwb_comm_period"])
    v.set_state(pos=sc["pos"], att=sc["att"])
    for i, p in sc["anchors"]:
        v.add_anchor(i, p)
    d3 = lambda a: (C.c_double * len(a))(*[float(x) for x in a])
    for k in range(30):
        tr = v.run(50, sched=[e for e in sc["sched"]])[-1]
        p1, p2 = v.telemetry()
        mine, ref = agf.abi.MsgTelemetry(), agf.abi.MsgTelemetry()
        agf.lib().agf_msg_telemetry_fill(p1.ctypes.data, p2.ctypes.data, C.byref(mine))
        R.L.orc_msg_telemetry(p1.ctypes.data, p2.ctypes.data, C.byref(ref))
        assert bytes(mine) == bytes(ref), k
        st_mine, st_ref = agf.abi.MsgSimulatorTruth(), agf.abi.MsgSimulatorTruth()
        agf.lib().agf_msg_simulator_truth_fill(0, d3(tr[0:3]), d3(tr[3:6]), d3(tr[6:10]), d3(tr[10:13]), C.byref(st_mine))
        R.L.orc_msg_simulator_truth(v.h, C.byref(st_ref))
        assert bytes(st_mine) == bytes(st_ref), k
    assert abs(mine.position[2] - tr[2]) < 0.1 and mine.motorForces[0] > 0.05  # it flies, and the fields are the flight's


def test_reference_golden_still_reproducible(agf, orc_mod):
    """The committed vectors are what the reference build in this container produces today."""
    R = oracle_or_skip(orc_mod, "ref-glibc")
    tr, _ = run_oracle(R, agf, scenario(agf, "rates"))
    assert bit_equal(tr[GOLD["ref-glibc/rates/ticks"]], GOLD["ref-glibc/rates/traj"])


def test_survey_sanity_vectors(agf, port_glibc):
    """SURVEY.md section 8c probe values for the rates-mode scenario (recorded independently)."""
    tr, _ = run_oracle(port_glibc, agf, scenario(agf, "rates"))
    assert tr[999, 0] == 0 and tr[999, 1] == 0
    assert abs(tr[999, 2] - 0.939083617525108) < 1e-14
    assert abs(tr[999, 5] - 0.960208197878) < 1e-11
    assert abs(tr[999, 13] - 2909.44384766) < 1e-7
    np.testing.assert_allclose(tr[4999, 0:3], [0.342036496412748, -0.886737158754551, 24.3015536037538], rtol=0, atol=1e-12)
    np.testing.assert_allclose(tr[4999, 6:10], [0.999999946981528, 1.30944008436e-4, -1.00535276674e-4, 2.80683562528e-4], rtol=0, atol=1e-12)


def test_full_mode_counters(agf, port_glibc):
    tr, _ = run_oracle(port_glibc, agf, scenario(agf, "full"))
    assert tr[-1, 39] == 1665 and tr[-1, 38] == 0 and tr[-1, 37] == 2 and tr[-1, 35] == 0 and tr[-1, 34] == 2
    # closed-loop tracking of the last set-point
    assert np.linalg.norm(tr[-1, 0:3] - np.array([1.5, 0.7, 2.0])) < 0.02


def test_propeller_calibration_port_and_device_code_match_reference(agf, orc_mod, port_shared):
    """QuadcopterLogic.cpp:553-587: rates commands with the CALIBRATE_MOTORS flag for > 750 logic cycles, then the flag
    cleared: the per-motor thrust corrections change the motor commands (thrust 1.08 g -> correction 1 / 1.08).  The port and
    the device step compiled for the host reproduce the unmodified reference bit for bit, every tick."""
    sc = agf.scenarios.calibration_scenario(agf.codec)
    a, _ = run_oracle(port_shared, agf, sc)
    plain, _ = run_oracle(port_shared, agf, agf.scenarios.calibration_scenario(agf.codec, flag_until=0))
    assert np.allclose(a[1800, 13:17], plain[1800, 13:17], rtol=1e-6)          # no effect while calibrating
    assert np.all(a[-1, 13:17] > 1.03 * plain[-1, 13:17])                       # corrections at work afterwards
    for fl in ("ref-shared", "hostsim-shared"):
        if not orc_mod.available(fl):
            continue
        b, _ = run_oracle(orc_mod.Oracle(fl), agf, sc)
        assert bit_equal(a, b), fl


def test_hostsim_matches_port(agf, orc_mod, port_shared):
    """The product's device step header compiled for the host == the literal port, bit for bit."""
    if not orc_mod.available("hostsim-shared"):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "hostsim"], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("hostsim does not build here: " + r.stderr[-300:])
    H = orc_mod.Oracle("hostsim-shared")
    for name in SCENARIOS:
        sc = scenario(agf, name)
        a, va = run_oracle(port_shared, agf, sc)
        b, vb = run_oracle(H, agf, sc)
        assert bit_equal(a, b), name
        fa, fb = va.full(), vb.full()
        keys = ["pos", "vel", "att", "ang_vel", "motor_speed", "motor_force_z", "motor_speed_cmd", "flight_state",
                "first_panic_reason", "cycle_counter", "tel_warnings", "des_motor_forces", "gyro_lpf", "acc_lpf",
                "temp_lpf", "batt_lpf", "batt_voltage_filtered", "monitor_cmd_rate_lpdt", "monitor_main_loop_lpdt",
                "uwb_meas_count", "next_ranging_target_idx", "kf_pos", "kf_vel", "kf_att", "kf_ang_vel", "kf_last_corr",
                "kf_imu_init", "kf_uwb_init", "kf_num_resets", "kf_num_rejected", "kf_num_rejected_seq"]
        if sc["uwb_comm_period"] > 0:
            keys.append("kf_cov")
        for k in keys:
            assert bit_equal(fa[k], fb[k]), (name, k)


def test_hostsim_chunked_runs_equal_single_run(agf, orc_mod):
    """Serialising the state to the HBM layout between launches loses nothing."""
    if not orc_mod.available("hostsim-shared"):
        pytest.skip("hostsim not built")
    H = orc_mod.Oracle("hostsim-shared")
    sc = scenario(agf, "full")
    a, _ = run_oracle(H, agf, sc, nticks=1500)
    v = H.vehicle(agf.vehicle_cfg(sc["quad_type"], sc["vehicle_id"], motor_time_const=sc["motor_time_const"]),
                  uwb_comm_period=sc["uwb_comm_period"])
    v.set_state(pos=sc["pos"], att=sc["att"])
    for i, p in sc["anchors"]:
        v.add_anchor(i, p)
    parts = [v.run(c, sched=sc["sched"]) for c in (1, 1, 7, 91, 400, 1000)]
    assert bit_equal(a, np.concatenate(parts))


# ---- physics properties of the oracle (the same ones the GPU tests use at full size) -------------
def test_hover_equilibrium_speed(agf, port_glibc):
    """omega_hover = sqrt(m g / (4 kF)) holds the vehicle still."""
    cfg = agf.vehicle_cfg(vehicle_id=1)
    v = port_glibc.vehicle(cfg)
    v.set_state(pos=(0, 0, 1.0))
    # rates command with thrust exactly g: the mixer asks each motor for m g / 4
    raw = agf.codec.encode_rates(0, 9.81, (0, 0, 0))
    sched = [(k, raw, -1) for k in range(0, 1000, 10)]
    tr = v.run(1000, sched=sched)
    w_hover = np.sqrt(cfg.mass * 9.81 / (4 * cfg.prop_thrust_from_speed_sqr))
    assert abs(tr[-1, 13] / w_hover - 1) < 2e-4  # 16-bit thrust quantisation
    assert abs(tr[-1, 5]) < 0.05


def test_quaternion_norm_drift_is_bounded(agf, port_glibc):
    tr, _ = run_oracle(port_glibc, agf, scenario(agf, "rates"))
    assert np.max(np.abs(np.linalg.norm(tr[:, 6:10], axis=1) - 1)) < 1e-12


def test_time_base_latencies(agf, port_glibc):
    """Run() before Advance(): tick 0 is a no-op, logic first runs at tick 2."""
    tr, _ = run_oracle(port_glibc, agf, scenario(agf, "rates"), nticks=10)
    assert list(tr[:5, 36]) == [0, 0, 1, 2, 3]


def test_population_trajectories_port_equals_reference_and_envelope_exists(agf, orc_mod):
    """oracle/orc_population_traj.inc: the population helper with a trajectory (every 50th tick).  The port reproduces
    the unmodified reference bit for bit at trajectory level for a randomized full-mode population, with either libm; and
    the reference's own sensitivity envelope -- the same sources rebuilt with FMA contraction (ref-fma) -- is non-zero in
    full mode, which is what tests/test_fast_population_gpu.py measures the fast kernels against."""
    n, nt, stride = 48, 2500, 50
    s = agf.scenarios
    sc = s.full_scenario(agf.codec, nticks=nt)
    init = s.monte_carlo_initial_states(n, seed=1234)
    idle = agf.codec.encode_idle(0)
    slot = np.array([np.frombuffer(agf.codec.encode_position(0, (p[0], p[1], 1.5)), np.uint8) for p in init])
    sched = [(d, idle, -1) if sl == -2 else (d, None, 0) for d, _, sl in s.hover_slot_schedule(nt)]
    cfg = agf.vehicle_cfg(sc["quad_type"], sc["vehicle_id"], motor_time_const=sc["motor_time_const"])
    anchors = np.array([[i, *p] for i, p in sc["anchors"]], np.float32)
    slots = np.zeros((4, n, 23), np.uint8)
    slots[0] = slot
    tr = {}
    for fl in ("port-glibc", "port-shared", "ref-glibc", "ref-shared", "ref-fma"):
        if not orc_mod.available(fl):
            continue
        tr[fl], _ = orc_mod.Oracle(fl).run_population_traj(cfg, n, stride, init13=init, anchors=anchors, nticks=nt, sched=sched,
                                                           slot_raw=slots, threads=4, uwb_comm_period=sc["uwb_comm_period"])
        assert tr[fl].shape == (nt // stride, n, 40)
    # record r of the trajectory helper == the final record of a plain run of (r + 1) * stride ticks
    fin, _ = orc_mod.Oracle("port-glibc").run_population(cfg, n, init13=init, anchors=anchors, nticks=nt, sched=sched, slot_raw=slots,
                                                         threads=4, uwb_comm_period=sc["uwb_comm_period"])
    assert np.array_equal(tr["port-glibc"][-1], fin, equal_nan=True)
    if "ref-glibc" not in tr:
        pytest.skip("oracle/_ref not built here")
    assert np.array_equal(tr["port-glibc"], tr["ref-glibc"], equal_nan=True)
    assert np.array_equal(tr["port-shared"], tr["ref-shared"], equal_nan=True)
    d = np.abs(tr["ref-fma"][..., 0:3] - tr["ref-glibc"][..., 0:3]).max(axis=(0, 2))
    assert np.median(d) > 1e-7 and np.isfinite(d).all()   # the reference moves under a benign rebuild ...
    assert np.median(d) < 5e-2                            # ... but remains the same flight


def test_fast_formulation_on_the_host_holds_the_fp32_tolerance_in_rates_mode(agf, orc_mod):
    """oracle/hostsim/agf_hostsim_fast.cu: the FAST instantiations of the device step header compiled for the host (a
    development aid; the host compiler's FMA contraction and libm are not the GPU's, so this is indicative -- the binding
    check is tests/test_fast_population_gpu.py on the GPU).  Rates mode, randomized initial attitudes, 10 s: the FP32 plant
    (compensated position / velocity sums, quaternion kept at unit norm) stays within the north star's 1e-4 of the
    reference for every vehicle and within 2e-5 for the typical one; the FP64 plant within the reference's own FMA
    sensitivity (1e-5)."""
    import shutil
    if not (orc_mod.available("hostsim-fast32") and orc_mod.available("hostsim-fast64")):
        if not shutil.which("nvcc") and not os.path.exists("/usr/local/cuda/bin/nvcc"):
            pytest.skip("nvcc not available to build oracle/hostsim")
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "hostsim-fast"], stdout=subprocess.DEVNULL)
    n, nt, stride = 64, 5000, 50
    s = agf.scenarios
    sc = s.rates_scenario(agf.codec, nticks=nt)
    init = s.monte_carlo_initial_states(n, seed=77)
    cfg = agf.vehicle_cfg(sc["quad_type"], sc["vehicle_id"], motor_time_const=sc["motor_time_const"], motor_inertia=sc["motor_inertia"])
    tr = {}
    for fl in ("port-glibc", "hostsim-fast32", "hostsim-fast64"):
        tr[fl], _ = orc_mod.Oracle(fl).run_population_traj(cfg, n, stride, init13=init, nticks=nt, sched=sc["sched"], threads=4)
    ref = tr["port-glibc"]
    for fl, tol_max, tol_med in (("hostsim-fast32", 1e-4, 2e-5), ("hostsim-fast64", 1e-5, 2e-6)):
        e = (np.abs(tr[fl][..., 0:17] - ref[..., 0:17]) / np.maximum(np.abs(ref[..., 0:17]), 1.0)).max(axis=(0, 2))
        assert e.max() <= tol_max and np.median(e) <= tol_med, (fl, np.median(e), e.max())
