"""Scenario definitions shared by tests, golden-vector generation and bench.py.

A scenario is (vehicle config edits, initial state, UWB anchors, command schedule); the same
schedule drives the CUDA batch (agf_batch_set_cmd_schedule) and the CPU oracles.  The timing
reproduces the offboard loop of Simulator/Rappids_Simulator/main.cpp:471-739: a command is
computed right after the Run() of tick g (the first g whose elapsed time exceeds the loop period,
strict '>'), queued with the uplink delay (CommunicationsDelay.hpp:18-39) and therefore delivered
before the Run() of tick g + delay_ticks + 1.
"""
import numpy as np

DT_US = 2000


def command_ticks(nticks, period_ticks, delay_ticks):
    """(generation tick g, delivery tick) pairs.  g = period, 2*period, ... (strict '>' on the
    stopwatch: with Run() before Advance() the first generation is after the Run of tick `period`)."""
    out = []
    g = period_ticks
    while g + delay_ticks + 1 < nticks:
        out.append((g, g + delay_ticks + 1))
        g += period_ticks
    return out


def rates_scenario(codec, nticks=5000):
    """SURVEY section 8c "Rates mode": external rates command, thrust 1.05 g, a short body-rate
    doublet at t in (2.0, 2.2) s, 100 Hz commands, 30 ms uplink delay."""
    sched = []
    for g, d in command_ticks(nticks, 5, 15):
        t = (g + 1) * DT_US * 1e-6
        if 2.0 < t < 2.1:
            w = (1.0, 0.5, 0.2)
        elif 2.1 <= t < 2.2:
            w = (-1.0, -0.5, -0.2)
        else:
            w = (0.0, 0.0, 0.0)
        sched.append((d, codec.encode_rates(0, np.float32(1.05 * 9.81), w), -1))
    return dict(name="rates", quad_type=5, vehicle_id=1, motor_time_const=0.0, motor_inertia=0.0,
                pos=(0.0, 0.0, 0.0), att=(1.0, 0.0, 0.0, 0.0), anchors=[], uwb_comm_period=0.0,
                sched=sched, nticks=nticks)


def calibration_scenario(codec, nticks=2600, flag_until=1900):
    """Propeller calibration (QuadcopterLogic.cpp:553-587): rates commands with the CALIBRATE_MOTORS flag set for more
    than 750 logic cycles (thrust 1.08 g, a gentle constant body rate so that the four motors differ), then the flag
    cleared -- the logic turns the accumulated thrusts into per-motor correction factors and the mixer uses them."""
    sched = []
    for g, d in command_ticks(nticks, 5, 15):
        flags = 0x01 if d < flag_until else 0x00
        sched.append((d, codec.encode_rates(flags, np.float32(1.08 * 9.81), (0.05, -0.03, 0.1)), -1))
    return dict(name="calibration", quad_type=5, vehicle_id=1, motor_time_const=0.015, motor_inertia=0.0,
                pos=(0.0, 0.0, 0.0), att=(1.0, 0.0, 0.0, 0.0), anchors=[], uwb_comm_period=0.0,
                sched=sched, nticks=nticks)


ANCHORS_8 = [(101 + i, (x, y, z)) for i, (x, y, z) in enumerate(
    [(sx * 3.0, sy * 3.0, z) for z in (0.1, 3.0) for sx in (1, -1) for sy in (1, -1)])]


def full_scenario(codec, nticks=5000, start=(0.5, -0.3, 0.0), wp1=None, wp2=None, motor_time_const=0.015):
    """SURVEY section 8c "Full mode": onboard position control with the 9-state EKF and an 8-anchor
    UWB network; idle commands for the first second, then a hover set-point, then a step."""
    wp1 = (start[0], start[1], 1.5) if wp1 is None else wp1
    wp2 = (start[0] + 1.0, start[1] + 1.0, 2.0) if wp2 is None else wp2
    sched = []
    zero = (0.0, 0.0, 0.0)
    for g, d in command_ticks(nticks, 10, 10):
        t = (g + 1) * DT_US * 1e-6
        if t < 1.0:
            raw = codec.encode_idle(0)
        elif t < 5.0:
            raw = codec.encode_position(0, wp1, zero, zero)
        else:
            raw = codec.encode_position(0, wp2, zero, zero)
        sched.append((d, raw, -1))
    return dict(name="full", quad_type=5, vehicle_id=1, motor_time_const=motor_time_const,
                motor_inertia=0.0, pos=tuple(start), att=(1.0, 0.0, 0.0, 0.0), anchors=ANCHORS_8,
                uwb_comm_period=0.004, sched=sched, nticks=nticks)


def accel_scenario(codec, nticks=2500):
    """External acceleration command mode (QuadcopterLogic.cpp:459-526): climb, then a lateral
    acceleration pulse with yaw rate; exercises ToEulerYPR / FromEulerYPR (atan2f, asinf)."""
    sched = []
    for g, d in command_ticks(nticks, 10, 10):
        t = (g + 1) * DT_US * 1e-6
        if t < 0.5:
            raw = codec.encode_idle(0)
        elif t < 2.0:
            raw = codec.encode_acceleration(0, (0.0, 0.0, 1.0), 0.0)
        elif t < 3.0:
            raw = codec.encode_acceleration(0, (1.5, -0.8, 0.0), 0.6)
        else:
            raw = codec.encode_acceleration(0, (-1.0, 0.5, -0.5), -0.3)
        sched.append((d, raw, -1))
    return dict(name="accel", quad_type=4, vehicle_id=13, motor_time_const=0.02, motor_inertia=1e-6,
                pos=(0.0, 0.0, 0.0), att=(1.0, 0.0, 0.0, 0.0), anchors=[], uwb_comm_period=0.0,
                sched=sched, nticks=nticks)


def waypoint_square_schedule(codec, nticks=5000, z=1.5, leg_s=2.5, idle_s=0.5):
    """4-waypoint square (+-1, +-1, z), leg_s seconds each, 50 Hz position commands with
    20 ms uplink delay, shared by the whole population."""
    wps = [(1.0, 1.0, z), (-1.0, 1.0, z), (-1.0, -1.0, z), (1.0, -1.0, z)]
    zero = (0.0, 0.0, 0.0)
    sched = []
    for g, d in command_ticks(nticks, 10, 10):
        t = (g + 1) * DT_US * 1e-6
        if t < idle_s:
            raw = codec.encode_idle(0)
        else:
            raw = codec.encode_position(0, wps[int((t - idle_s) / leg_s) % 4], zero, zero)
        sched.append((d, raw, -1))
    return sched


def hover_slot_schedule(nticks=5000, idle_s=0.5):
    """every vehicle hovers at its own set-point (per-vehicle packets in command slot 0,
    idle packets in slot 1 are not needed: idle is a broadcast)."""
    sched = []
    for g, d in command_ticks(nticks, 10, 10):
        t = (g + 1) * DT_US * 1e-6
        sched.append((d, None, -2 if t < idle_s else 0))  # -2: caller substitutes an idle broadcast
    return sched


def monte_carlo_initial_states(n, seed=1234, yaw_max=np.pi):
    """SURVEY C2 initial conditions: p_xy ~ U(-1,1) m, p_z = 0, yaw ~ U(-yaw_max,yaw_max) (C2: pi),
    roll/pitch ~ U(-5,5) deg, v = w = 0.  Returns [n][13] (pos3 vel3 att4 angvel3), float64.
    numpy's Philox bit generator keyed by `seed` (counter-based, reproducible anywhere).

    Note: the reference's onboard EKF initialises its yaw estimate to 0 (sigma 30 deg about gravity,
    KalmanFilter6DOF.cpp:16-19,77-108); with |yaw| > ~120 deg its position loop is unstable and the
    vehicle panics -- on the reference CPU code exactly as on the GPU (about 21 % of a U(-pi,pi)
    population in the waypoint scenario).  The throughput workload (C3) therefore draws yaw from
    +-60 deg so that every vehicle flies; C2 keeps the full range."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    u = rng.random((n, 5))
    px = -1.0 + 2.0 * u[:, 0]
    py = -1.0 + 2.0 * u[:, 1]
    yaw = -yaw_max + 2.0 * yaw_max * u[:, 2]
    roll = np.deg2rad(-5.0 + 10.0 * u[:, 3])
    pitch = np.deg2rad(-5.0 + 10.0 * u[:, 4])
    cy, sy = np.cos(0.5 * yaw), np.sin(0.5 * yaw)
    cp, sp = np.cos(0.5 * pitch), np.sin(0.5 * pitch)
    cr, sr = np.cos(0.5 * roll), np.sin(0.5 * roll)
    # Rotation::FromEulerYPR (Rotation.hpp:99-110)
    q = np.stack([cy * cp * cr + sy * sp * sr, cy * cp * sr - sy * sp * cr,
                  cy * sp * cr + sy * cp * sr, sy * cp * cr - cy * sp * sr], axis=1)
    out = np.zeros((n, 13))
    out[:, 0] = px
    out[:, 1] = py
    out[:, 6:10] = q
    return out


def offboard_scenario(nticks=4000):
    """the closed loop Rappids_Simulator flies (main.cpp:471-739, CTRL_OFFBOARD_RATES) with a
    truth-fed estimate: offboard position controller at 100 Hz -> 16-bit rates commands -> 30 ms uplink delay ->
    onboard rate controller.  Take-off to a hover set-point, then a step to a second set-point.  No UWB, the
    onboard estimator runs its complementary attitude filter only."""
    targets = [(0, (0.0, 0.0, 2.0)), (3000000, (1.0, -0.5, 2.5)), (6000000, (1.0, -0.5, 1.0))]
    return dict(name="offboard", quad_type=5, vehicle_id=1, motor_time_const=0.015, motor_inertia=0.0,
                pos=(0.0, 0.0, 0.0), att=(1.0, 0.0, 0.0, 0.0), anchors=[], uwb_comm_period=0.0, sched=[],
                targets=targets, nticks=nticks)


def stages_scenario(traj_id=3, nticks=7000):
    """the flight stages of the ROS rates-control node (ExampleVehicleStateMachine.cpp:93-370) as the
    reference generator of the offboard loop: start signal at 0.5 s -> spool-up 0.5 s -> 2 s take-off ramp -> flight
    trajectory `traj_id` -> stop signal at 9 s -> landing -> idle."""
    sc = offboard_scenario(nticks)
    sc.update(name="stages%d" % traj_id, pos=(0.2, -0.1, 0.0),
              ref=dict(kind=1, start_us=500000, stop_us=9000000, desired_pos=(0.0, 0.0, 1.0), desired_yaw=0.3, traj_id=traj_id),
              primitive=None)
    return sc


def stages_emergency_scenario(nticks=4000):
    """The flight stages with Offboard::SafetyNet on and a set-point outside the lab-space box (x = 2.5 m > 1.8 m): the
    take-off ramp carries the estimate across the box, the stage machine latches its emergency stage and sends kill commands."""
    sc = stages_scenario(3, nticks)
    sc.update(name="stages3-emergency", ref=dict(sc["ref"], desired_pos=(2.5, 0.0, 1.0), safety_net=True))
    return sc


def tracking_scenario(nticks=3400):
    """Rappids_Simulator's tracking of a planned motion primitive (main.cpp:560-634): hover at 2 m with
    QuadcopterController::Run until 4 s, then RunTracking along a 2.5 s rest-to-rest quintic given in a frame yawed by
    0.4 rad (trajAtt) and anchored at the hover point (trajOffset)."""
    import numpy as np
    sc = offboard_scenario(nticks)
    sc.update(name="tracking", pos=(0.2, -0.1, 0.0),
              ref=dict(kind=2, start_us=4000000, stop_us=0, desired_pos=(0.2, -0.1, 2.0), desired_yaw=0.1, traj_id=0),
              primitive=dict(pf=(1.5, 0.5, 0.3), T=2.5, offset=(0.2, -0.1, 2.0), att=(float(np.cos(0.2)), 0.0, 0.0, float(np.sin(0.2)))))
    return sc


# ---------------------------------------------------------------------------------------------
# RAPPIDS planner workloads: synthetic depth scenes + planner initial states
# ---------------------------------------------------------------------------------------------
RAPPIDS_DEPTH_SCALE = 10.0 / 256.0   # Simulator/Rappids_Simulator/main.cpp:121-122
RAPPIDS_MAX_BOXES = 4


def rappids_population(n, seed=2024, width=320, height=240, camera_height=1.5, first=0, speed_max=2.0,
                       acc_max=0.5, box_depth=(1.5, 4.0), n_boxes=(1, 3)):
    """Scene descriptions and initial states of vehicles first .. first+n-1 of a population (independent of the
    sharding: every vehicle draws from its own Philox stream).

    Scene: background at 8 m, a floor band below the horizon (camera `camera_height` above a flat floor),
    1-3 axis-aligned boxes at 1.5-4 m.  Returned as integer data so that host (numpy) and device rasterisers
    produce identical images:
      row_bg  [n, height] uint16   background/floor pixel value of every image row
      boxes   [n, RAPPIDS_MAX_BOXES, 5] int32  (x0, x1, y0, y1, value), x1/y1 exclusive, value 0 = unused
    State in the camera frame (x right, y down, z forward; Rappids_Simulator main.cpp:491-497):
      vel0 forward 0-2 m/s with a small lateral part, acc0 small, grav = 9.81 m/s^2 along +y tilted by up to 5 deg.
    """
    f = width / 2.0
    row_bg = np.zeros((n, height), dtype=np.uint16)
    boxes = np.zeros((n, RAPPIDS_MAX_BOXES, 5), dtype=np.int32)
    vel0 = np.zeros((n, 3))
    acc0 = np.zeros((n, 3))
    grav = np.zeros((n, 3))
    ys = np.arange(height)
    for i in range(n):
        rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 0, first + i]))
        h = camera_height * rng.uniform(0.8, 1.2)
        with np.errstate(divide="ignore"):
            floor = np.where(ys > height / 2.0, h * f / np.maximum(ys - height / 2.0, 1e-9), np.inf)
        depth = np.minimum(8.0, floor)
        row_bg[i] = np.floor(depth / RAPPIDS_DEPTH_SCALE).astype(np.uint16)
        nb = int(rng.integers(n_boxes[0], n_boxes[1] + 1))
        for b in range(nb):
            d = rng.uniform(box_depth[0], box_depth[1])
            w = int(rng.integers(width // 16, width // 4))
            hh = int(rng.integers(height // 8, height // 2))
            x0 = int(rng.integers(0, width - w))
            y0 = int(rng.integers(0, height - hh))
            boxes[i, b] = (x0, x0 + w, y0, y0 + hh, int(d / RAPPIDS_DEPTH_SCALE))
        vel0[i] = (rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2), rng.uniform(0.0, speed_max))
        acc0[i] = rng.uniform(-acc_max, acc_max, size=3)
        tilt = np.deg2rad(rng.uniform(-5.0, 5.0, size=2))
        grav[i] = 9.81 * np.array([np.sin(tilt[0]), np.cos(tilt[0]) * np.cos(tilt[1]), np.sin(tilt[1])])
    return dict(row_bg=row_bg, boxes=boxes, vel0=vel0, acc0=acc0, grav=grav, width=width, height=height)


def rappids_render(row_bg, boxes, width):
    """Host rasteriser of a scene description: pixel = min(row background, values of the boxes covering it)."""
    n, height = row_bg.shape
    img = np.repeat(row_bg[:, :, None], width, axis=2).astype(np.uint16)
    for i in range(n):
        for x0, x1, y0, y1, v in boxes[i]:
            if v > 0:
                img[i, y0:y1, x0:x1] = np.minimum(img[i, y0:y1, x0:x1], np.uint16(v))
    return img


def rappids_candidates(n, k, seed=77, width=320, height=240, first=0):
    """[n, k, 4] candidate end points (camera frame) + durations drawn like the reference's
    RandomTrajectoryGenerator (DepthImagePlanner.hpp:334-352: pixel in the central 80 % of the image, depth
    1.5-3 m, duration 2-3 s), from numpy's Philox (the population generator; the parity tests also use the
    reference's own std::mt19937 draws)."""
    out = np.zeros((n, k, 4))
    f = width / 2.0
    for i in range(n):
        rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 1, first + i]))
        px = rng.uniform(0.1 * width, 0.9 * width, size=k)
        py = rng.uniform(0.1 * height, 0.9 * height, size=k)
        d = rng.uniform(1.5, 3.0, size=k)
        out[i, :, 0] = d * ((px - width / 2.0) / f)
        out[i, :, 1] = d * ((py - height / 2.0) / f)
        out[i, :, 2] = d * 1
        out[i, :, 3] = rng.uniform(2.0, 3.0, size=k)
    return out
