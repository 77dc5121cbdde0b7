"""agri-fly_b200 -- Python driver (ctypes) for libagrifly_b200.so, the B200-native batched
implementation of agri-fly's per-vehicle simulation step.

The package directory name carries the reference's hyphen; import it as `agrifly_b200` (alias
package at the repo root) or with importlib.import_module("agri-fly_b200").

The product is the CUDA library behind include/agrifly_b200.h.  This module only marshals numpy
arrays across that C ABI; it has no CPU implementation of the step and raises if the library or a
CUDA device is missing.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as abi
from . import scenarios  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AGF_LIB_PATH") or os.path.join(HERE, "libagrifly_b200.so")  # override: tuning variants only
_lib = None


def offboard_cfg(quad_type=5, **edits):
    """agf_offboard_cfg_default(quad_type) with field edits (period_us=..., delay_us=..., yaw_angle=...)."""
    c = abi.OffboardCfg()
    _check(lib().agf_offboard_cfg_default(int(quad_type), C.byref(c)))
    for k, v in edits.items():
        setattr(c, k, v)
    return c


def offboard_ref(kind, start_us=0, stop_us=2**64 - 1, desired_pos=(0.0, 0.0, 1.0), desired_yaw=0.0, traj_id=0, safety_net=False):
    r = abi.OffboardRef(kind=int(kind), traj_id=int(traj_id), start_us=int(start_us), stop_us=int(stop_us),
                        desired_yaw=float(desired_yaw))
    r.desired_pos[:] = [float(x) for x in desired_pos]
    if safety_net:
        lib().agf_offboard_ref_safety_default(C.byref(r))  # Offboard::SafetyNet's lab-space box, 1 m, 0.5 s
    return r


def csv_header():
    """Header line of Rappids_Simulator's simulation.csv (agf_csv_header)."""
    buf = C.create_string_buffer(1024)
    lib().agf_csv_header(buf, len(buf))
    return buf.value.decode()


def csv_row(t, pos, vel, att, ang_vel, motor_forces, est_pos, est_vel, est_att, est_ang_vel, des_pos, des_vel, panic_reason,
            last_radio_cmd):
    """One row of simulation.csv (agf_csv_format_row) from the values Rappids_Simulator logs."""
    r = abi.CsvRecord(t=float(t), panic_reason=int(panic_reason))
    for name, val in (("pos", pos), ("vel", vel), ("att", att), ("ang_vel", ang_vel), ("motor_forces", motor_forces),
                      ("est_pos", est_pos), ("est_vel", est_vel), ("est_att", est_att), ("est_ang_vel", est_ang_vel),
                      ("des_pos", des_pos), ("des_vel", des_vel), ("last_radio_cmd", last_radio_cmd)):
        getattr(r, name)[:] = [float(x) for x in val]
    buf = C.create_string_buffer(2048)
    lib().agf_csv_format_row(C.byref(r), buf, len(buf))
    return buf.value.decode(), r


def sweep_cfgs(base, n, seed=5):
    """BASELINE config 4's parameter sweep as an [n][sizeof(agf_vehicle_cfg)] byte array for Batch(): mass x U(0.8, 1.2),
    I_xx = I_yy and I_zz x U(0.7, 1.3), kF x U(0.9, 1.1), k_tau x U(0.8, 1.2), motor time constant U(0, 0.05) s
    (SURVEY.md 8d).  cfg_at(arr, i) gives vehicle i's configuration back as a struct."""
    V = abi.VehicleCfg
    raw = np.tile(np.frombuffer(bytes(base), np.uint8), (n, 1))
    rng = np.random.default_rng(seed)

    def col(field, k=0):
        off = getattr(V, field).offset + 8 * k
        return raw[:, off:off + 8].view(np.float64)[:, 0]

    def put(field, values, k=0):
        off = getattr(V, field).offset + 8 * k
        raw[:, off:off + 8] = np.ascontiguousarray(values, dtype=np.float64).view(np.uint8).reshape(n, 8)

    put("mass", col("mass") * rng.uniform(0.8, 1.2, n))
    ixx = col("inertia", 0) * rng.uniform(0.7, 1.3, n)
    put("inertia", ixx, 0)
    put("inertia", ixx, 4)
    put("inertia", col("inertia", 8) * rng.uniform(0.7, 1.3, n), 8)
    put("prop_thrust_from_speed_sqr", col("prop_thrust_from_speed_sqr") * rng.uniform(0.9, 1.1, n))
    put("prop_torque_from_speed_sqr", col("prop_torque_from_speed_sqr") * rng.uniform(0.8, 1.2, n))
    put("motor_time_const", rng.uniform(0.0, 0.05, n))
    return raw


def cfg_at(arr, i):
    return abi.VehicleCfg.from_buffer_copy(arr[i].tobytes())


def offboard_estimator(**edits):
    """agf_offboard_estimator_default() (MocapStateEstimator as Rappids_Simulator sets it up) with field edits."""
    e = abi.OffboardEstimator()
    _check(lib().agf_offboard_estimator_default(C.byref(e)))
    for k, v in edits.items():
        setattr(e, k, v)
    return e


def primitive_record(pf, T, p0=(0, 0, 0), v0=(0, 0, 0), a0=(0, 0, 0), grav=(0.0, 0.0, -9.81), att=(1.0, 0.0, 0.0, 0.0),
                     offset=(0.0, 0.0, 0.0)):
    """AGF_OFFTRAJ_DOUBLES record of the minimum-jerk primitive from (p0, v0, a0) to rest at pf in T seconds
    (fully defined end state: SingleAxisTrajectory.cpp:59-103 with position, velocity and acceleration goals)."""
    rec = np.zeros(abi.OFFTRAJ_DOUBLES)
    for a in range(3):
        dp = pf[a] - p0[a] - v0[a] * T - 0.5 * a0[a] * T * T
        dv = 0.0 - v0[a] - a0[a] * T
        da = 0.0 - a0[a]
        T2, T3, T4, T5 = T * T, T ** 3, T ** 4, T ** 5
        alpha = (60 * T2 * da - 360 * T * dv + 720 * dp) / T5
        beta = (-24 * T3 * da + 168 * T2 * dv - 360 * T * dp) / T5
        gamma = (3 * T4 * da - 24 * T3 * dv + 60 * T2 * dp) / T5
        rec[6 * a:6 * a + 6] = [p0[a], v0[a], a0[a], alpha, beta, gamma]
    rec[18:21] = grav
    rec[21] = T
    rec[22:26] = att
    rec[26:29] = offset
    return rec


class AgfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("agrifly_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load libagrifly_b200.so (built in-tree by agri-fly_b200/build.py). No fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libagrifly_b200.so is not built: run `python agri-fly_b200/build.py` "
                              "(there is no CPU fallback for the simulation step)")
        _lib = abi.bind(C.CDLL(LIB_PATH))
    return _lib


def _check(rc):
    if rc != 0:
        raise AgfError(rc, lib().agf_last_error_string().decode(errors="replace"))


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class Codec:
    """RadioTypes / TelemetryPacket codecs of the C ABI."""

    @staticmethod
    def _raw():
        return (C.c_uint8 * abi.RADIO_PACKET_SIZE)()

    def encode_rates(self, flags, thrust, w):
        raw = self._raw()
        lib().agf_radio_encode_rates(flags, float(thrust), _f3(w), raw)
        return bytes(raw)

    def encode_position(self, flags, p, v=(0, 0, 0), a=(0, 0, 0)):
        raw = self._raw()
        lib().agf_radio_encode_position(flags, _f3(p), _f3(v), _f3(a), raw)
        return bytes(raw)

    def encode_acceleration(self, flags, a, yaw_rate):
        raw = self._raw()
        lib().agf_radio_encode_acceleration(flags, _f3(a), float(yaw_rate), raw)
        return bytes(raw)

    def encode_idle(self, flags=0):
        raw = self._raw()
        lib().agf_radio_encode_idle(flags, raw)
        return bytes(raw)

    def encode_kill(self, flags=0):
        raw = self._raw()
        lib().agf_radio_encode_kill(flags, raw)
        return bytes(raw)

    def decode(self, raw):
        buf = (C.c_uint8 * abi.RADIO_PACKET_SIZE)(*bytes(raw))
        t, f = C.c_uint8(), C.c_uint8()
        fl = (C.c_float * 10)()
        lib().agf_radio_decode(buf, C.byref(t), C.byref(f), fl)
        return t.value, f.value, np.array(fl, dtype=np.float32)

    def decode_telemetry(self, packet):
        buf = (C.c_uint8 * abi.TELEMETRY_PACKET_SIZE)(*bytes(packet))
        t = abi.Telemetry()
        lib().agf_telemetry_decode(buf, C.byref(t))
        return t


codec = Codec()


def quad_type_from_id(vehicle_id):
    return lib().agf_quad_type_from_id(int(vehicle_id))


def vehicle_cfg(quad_type=None, vehicle_id=1, **overrides):
    """agf_vehicle_cfg for an airframe type (default: the type of `vehicle_id`)."""
    if quad_type is None:
        quad_type = quad_type_from_id(vehicle_id)
    c = abi.VehicleCfg()
    _check(lib().agf_vehicle_cfg_from_type(int(quad_type), int(vehicle_id), C.byref(c)))
    for k, v in overrides.items():
        setattr(c, k, v)
    return c


def make_schedule(entries):
    """[(tick, raw|None, slot)] -> ctypes array of agf_cmd_entry."""
    arr = (abi.CmdEntry * max(1, len(entries)))()
    for i, e in enumerate(entries):
        arr[i].tick = int(e[0])
        arr[i].slot = int(e[2]) if len(e) > 2 else -1
        if e[1] is not None:
            C.memmove(arr[i].raw, bytes(e[1]), abi.RADIO_PACKET_SIZE)
    return arr


class Batch:
    """A batch of vehicles on one GPU (opaque agf_batch handle)."""

    def __init__(self, cfgs, n, precision=abi.PREC_FP64, math=abi.MATH_PARITY, device=0,
                 onboard_logic_period=1.0 / 500.0, uwb_comm_period=0.0, sigma_acc=0.0, sigma_gyro=0.0,
                 bias_sigma_acc=0.0, bias_sigma_gyro=0.0, uwb_noise_std_dev=0.0, seed=1,
                 first_global_index=0, stream=None, telemetry_warnings=True, block_threads=0):
        L = lib()
        o = abi.BatchOpts()
        L.agf_batch_opts_default(C.byref(o))
        o.device, o.precision, o.math, o.block_threads = device, precision, math, block_threads
        o.onboard_logic_period, o.uwb_comm_period = onboard_logic_period, uwb_comm_period
        o.sigma_acc, o.sigma_gyro = sigma_acc, sigma_gyro
        o.bias_sigma_acc, o.bias_sigma_gyro = bias_sigma_acc, bias_sigma_gyro
        o.uwb_noise_std_dev, o.seed, o.first_global_index = uwb_noise_std_dev, seed, first_global_index
        o.stream = stream
        o.telemetry_warnings = 1 if telemetry_warnings else 0
        if isinstance(cfgs, abi.VehicleCfg):
            arr, ncfg = (abi.VehicleCfg * 1)(cfgs), 1
        elif isinstance(cfgs, np.ndarray):  # [n][sizeof(agf_vehicle_cfg)] bytes: big sweeps built with numpy (sweep_cfgs())
            assert cfgs.dtype == np.uint8 and cfgs.ndim == 2 and cfgs.shape[1] == C.sizeof(abi.VehicleCfg) and cfgs.flags.c_contiguous
            arr, ncfg = cfgs.ctypes.data_as(C.POINTER(abi.VehicleCfg)), cfgs.shape[0]
        else:
            arr, ncfg = (abi.VehicleCfg * len(cfgs))(*cfgs), len(cfgs)
        h = C.c_void_p()
        _check(L.agf_batch_create(arr, ncfg, n, C.byref(o), C.byref(h)))
        self.h, self.n, self.L = h, n, L

    def close(self):
        if getattr(self, "h", None):
            self.L.agf_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # stepping
    def run(self, nticks, dt_us=2000):
        _check(self.L.agf_batch_run(self.h, dt_us, nticks))

    def advance_clock(self, dt_us):
        """ManualTimer::AdvanceMicroSeconds for the batch's clock; pair with run(1, dt_us=0) == Run()."""
        _check(self.L.agf_batch_advance_clock(self.h, dt_us))

    def sync(self):
        _check(self.L.agf_batch_sync(self.h))

    @property
    def time_us(self):
        return self.L.agf_batch_time_us(self.h)

    @property
    def ticks(self):
        return self.L.agf_batch_ticks(self.h)

    @property
    def stream(self):
        return self.L.agf_batch_stream(self.h)

    # state
    def get(self, name, first=0, count=None):
        fid, dt, nc = abi.FIELDS[name]
        count = self.n - first if count is None else count
        out = np.empty((count, nc), dtype=dt)
        _check(self.L.agf_batch_get_field(self.h, fid, out.ctypes.data, first, count))
        return out

    def set(self, name, values, first=0):
        fid, dt, nc = abi.FIELDS[name]
        v = np.ascontiguousarray(values, dtype=dt).reshape(-1, nc)
        _check(self.L.agf_batch_set_field(self.h, fid, v.ctypes.data, first, len(v)))

    def set_state13(self, s13, first=0):
        """position 3, velocity 3, attitude 4, angular velocity 3 per vehicle: one call, one copy (agf_batch_set_state)"""
        s13 = np.ascontiguousarray(s13, dtype=np.float64).reshape(-1, 13)
        _check(self.L.agf_batch_set_state(self.h, s13.ctypes.data, first, len(s13)))

    def get_state13(self, first=0, count=None):
        """the way back (agf_batch_get_state): [count][13] doubles"""
        count = self.n - first if count is None else count
        out = np.empty((count, 13), np.float64)
        _check(self.L.agf_batch_get_state(self.h, out.ctypes.data, first, count))
        return out

    def record(self):
        """[n][40] in the oracle's trajectory-record column order (oracle/oracle_api.h)."""
        r = np.zeros((self.n, 40))
        r[:, 0:3] = self.get("position")
        r[:, 3:6] = self.get("velocity")
        r[:, 6:10] = self.get("attitude")
        r[:, 10:13] = self.get("angular_velocity")
        r[:, 13:17] = self.get("motor_speed")
        r[:, 17:21] = self.get("motor_speed_cmd")
        r[:, 21:24] = self.get("est_position")
        r[:, 24:27] = self.get("est_velocity")
        r[:, 27:31] = self.get("est_attitude")
        r[:, 31:34] = self.get("est_angular_velocity")
        r[:, 34] = self.get("flight_state")[:, 0]
        r[:, 35] = self.get("panic_reason")[:, 0]
        r[:, 36] = self.get("cycle_counter")[:, 0]
        kc = self.get("kf_counters")
        r[:, 37], r[:, 38], r[:, 39] = kc[:, 0], kc[:, 1], kc[:, 2]
        return r

    # commands
    def set_radio(self, raw, first=0, count=None, broadcast=True):
        count = self.n - first if count is None else count
        if broadcast:
            buf = np.frombuffer(bytes(raw), dtype=np.uint8).copy()
        else:
            buf = np.ascontiguousarray(raw, dtype=np.uint8).reshape(count, abi.RADIO_PACKET_SIZE)
        _check(self.L.agf_batch_set_radio_cmd(self.h, buf.ctypes.data, first, count, 1 if broadcast else 0))

    def set_schedule(self, entries):
        entries = list(entries)
        _check(self.L.agf_batch_set_cmd_schedule(self.h, make_schedule(entries), len(entries)))

    def set_slot(self, slot, raw):
        buf = np.ascontiguousarray(raw, dtype=np.uint8).reshape(self.n, abi.RADIO_PACKET_SIZE)
        _check(self.L.agf_batch_set_cmd_slot(self.h, slot, buf.ctypes.data))

    def telemetry(self, first=0, count=None):
        count = self.n - first if count is None else count
        p1 = np.zeros((count, abi.TELEMETRY_PACKET_SIZE), np.uint8)
        p2 = np.zeros((count, abi.TELEMETRY_PACKET_SIZE), np.uint8)
        _check(self.L.agf_batch_get_telemetry(self.h, p1.ctypes.data, p2.ctypes.data, first, count))
        return p1, p2

    def set_external_wrench(self, force=None, torque=None, first=0, count=None):
        count = self.n - first if count is None else count
        f = None if force is None else np.ascontiguousarray(np.broadcast_to(force, (count, 3)), dtype=np.float64)
        t = None if torque is None else np.ascontiguousarray(np.broadcast_to(torque, (count, 3)), dtype=np.float64)
        _check(self.L.agf_batch_set_external_wrench(self.h, None if f is None else f.ctypes.data,
                                                    None if t is None else t.ctypes.data, first, count))

    def set_offboard_loop(self, cfg, targets, offsets=None):
        """In-kernel offboard rates loop (agrifly_b200.h).  cfg: abi.OffboardCfg (offboard_cfg()) or None to switch it
        off; targets: [(time_us, (x, y, z)), ...]; offsets: [n][3] per-vehicle shift of the targets or None."""
        if cfg is None:
            _check(self.L.agf_batch_set_offboard_loop(self.h, None, None, 0, None))
            return
        tarr = (abi.OffboardTarget * max(1, len(targets)))()
        for i, (t, p) in enumerate(targets):
            tarr[i].time_us = int(t)
            tarr[i].pos[:] = [float(x) for x in p]
        off = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.float64).reshape(self.n, 3)
        _check(self.L.agf_batch_set_offboard_loop(self.h, C.byref(cfg), tarr, len(targets),
                                                  None if off is None else off.ctypes.data))

    def set_offboard_reference(self, kind, start_us=0, stop_us=2**64 - 1, desired_pos=(0.0, 0.0, 1.0), desired_yaw=0.0, traj_id=0,
                               safety_net=False):
        """Reference generator of the offboard loop (agrifly_b200.h): abi.OFFREF_STAGES (flight stages of the ROS
        rates-control node) or abi.OFFREF_TRAJECTORY (tracking of per-vehicle motion primitives); None: targets."""
        if kind is None:
            _check(self.L.agf_batch_set_offboard_reference(self.h, None))
            return
        _check(self.L.agf_batch_set_offboard_reference(self.h, C.byref(offboard_ref(kind, start_us, stop_us, desired_pos,
                                                                                   desired_yaw, traj_id, safety_net))))

    def set_offboard_estimator(self, est):
        """est: abi.OffboardEstimator (offboard_estimator()) or None for the true state"""
        _check(self.L.agf_batch_set_offboard_estimator(self.h, None if est is None else C.byref(est)))

    def offboard_estimate(self, horizon=0.0, first=0, count=None):
        count = self.n - first if count is None else count
        e, c = np.empty((count, 13)), np.empty((count, 4))
        _check(self.L.agf_batch_get_offboard_estimate(self.h, float(horizon), e.ctypes.data, c.ctypes.data, first, count))
        return e, c

    def set_offboard_trajectories(self, traj, first=0):
        t = np.ascontiguousarray(traj, dtype=np.float64).reshape(-1, abi.OFFTRAJ_DOUBLES)
        _check(self.L.agf_batch_set_offboard_trajectories(self.h, t.ctypes.data, first, len(t)))

    def offboard_state(self, first=0, count=None):
        count = self.n - first if count is None else count
        out = np.empty((count, abi.OFFSTATE_DOUBLES))
        _check(self.L.agf_batch_get_offboard_state(self.h, out.ctypes.data, first, count))
        return out

    def add_anchor(self, id_, pos):
        _check(self.L.agf_batch_add_uwb_anchor(self.h, id_, _f3(pos)))

    def set_noise(self, seed, sigma_gyro, sigma_acc, bias_sigma_gyro=0.0, bias_sigma_acc=0.0):
        _check(self.L.agf_batch_set_noise(self.h, seed, sigma_gyro, sigma_acc, bias_sigma_gyro, bias_sigma_acc))

    # logging / statistics
    def set_uwb_noise(self, noise_std_dev, outlier_probability=0.0, outlier_std_dev=0.0):
        """UWBNetwork::SetNoiseProperties (UWBNetwork.hpp:28-33)"""
        _check(self.L.agf_batch_set_uwb_noise(self.h, float(noise_std_dev), float(outlier_probability), float(outlier_std_dev)))

    def enable_log(self, stride, capacity):
        _check(self.L.agf_batch_enable_log(self.h, stride, capacity))

    @property
    def log_count(self):
        return self.L.agf_batch_log_count(self.h)

    def read_log(self, rec, first=0, count=None):
        count = self.n - first if count is None else count
        out = np.empty((count, abi.LOG_FIELDS))
        _check(self.L.agf_batch_read_log(self.h, rec, out.ctypes.data, first, count))
        return out

    def stats(self, target=None):
        out = np.zeros(abi.STATS_LEN)
        t = None if target is None else np.ascontiguousarray(target, dtype=np.float64).reshape(self.n, 3)
        _check(self.L.agf_batch_reduce_stats(self.h, None if t is None else t.ctypes.data, out.ctypes.data))
        return out

    def stats_nccl(self, comm, target=None):
        """The population statistics over every rank of `comm` (an ncclComm_t as an int / c_void_p; None: this GPU only):
        stats kernel + one all-gather + combine on the batch's stream, result on the host."""
        out = np.zeros(abi.STATS_LEN)
        t = None if target is None else np.ascontiguousarray(target, dtype=np.float64).reshape(self.n, 3)
        _check(self.L.agf_batch_reduce_stats_nccl(self.h, comm, None if t is None else t.ctypes.data, out.ctypes.data))
        return out

    def stats_nccl_device(self, comm, dev_ptr, target=None):
        t = None if target is None else np.ascontiguousarray(target, dtype=np.float64).reshape(self.n, 3)
        _check(self.L.agf_batch_reduce_stats_nccl_device(self.h, comm, None if t is None else t.ctypes.data, dev_ptr))

    def stats_device(self, dev_ptr, target=None):
        t = None if target is None else np.ascontiguousarray(target, dtype=np.float64).reshape(self.n, 3)
        _check(self.L.agf_batch_reduce_stats_device(self.h, None if t is None else t.ctypes.data, dev_ptr))

    @property
    def launch_count(self):
        return self.L.agf_batch_launch_count(self.h)

    def step_kernel_time(self):
        ms, nl = C.c_double(), C.c_uint64()
        _check(self.L.agf_batch_step_kernel_time(self.h, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value


RESULT_DTYPE = np.dtype([("found", "<i4"), ("best_index", "<i4"), ("n_generated", "<i4"), ("n_cost_checks", "<i4"),
                         ("n_collision_checks", "<i4"), ("n_velocity_checks", "<i4"), ("n_collision_free", "<i4"),
                         ("n_pyramids", "<i4"), ("best_cost", "<f8"), ("best_coeffs", "<f8", (6, 3)), ("best_tf", "<f8"),
                         ("pyramid_cap_hit", "<i4"), ("reserved_", "<i4")])
assert RESULT_DTYPE.itemsize == C.sizeof(abi.RappidsResult)


def rappids_cfg(width=320, height=240, **edits):
    """agf_rappids_cfg_default(width, height) with field edits (cost_kind=..., cost_vec=(..), math=..., ...)."""
    c = abi.RappidsCfg()
    _check(lib().agf_rappids_cfg_default(int(width), int(height), C.byref(c)))
    for k, v in edits.items():
        if k == "cost_vec":
            c.cost_vec[:] = [float(x) for x in v]
        else:
            setattr(c, k, v)
    return c


class Rappids:
    """Batched RAPPIDS planner handle (include/agrifly_b200_rappids.h): one planner call per vehicle."""

    def __init__(self, cfg, n, max_candidates):
        self.n, self.kcap, self.cfg = int(n), int(max_candidates), cfg
        self.W, self.H = cfg.width, cfg.height
        self._h = C.c_void_p()
        _check(lib().agf_rappids_create(C.byref(cfg), self.n, self.kcap, C.byref(self._h)))
        self.k = 0

    def close(self):
        if self._h:
            lib().agf_rappids_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _cnt(self, first, count):
        return self.n - first if count is None else count

    def set_images(self, images, first=0):
        images = np.ascontiguousarray(images, dtype=np.uint16)
        assert images.shape[1:] == (self.H, self.W), images.shape
        _check(lib().agf_rappids_set_images(self._h, images.ctypes.data, first, images.shape[0]))

    def render_scenes(self, row_bg, boxes, first=0):
        row_bg = np.ascontiguousarray(row_bg, dtype=np.uint16)
        boxes = np.ascontiguousarray(boxes, dtype=np.int32)
        assert row_bg.shape[1] == self.H and boxes.shape[1:] == (abi.RAPPIDS_MAX_BOXES, 5)
        _check(lib().agf_rappids_render_scenes(self._h, row_bg.ctypes.data, boxes.ctypes.data, first, row_bg.shape[0]))

    def get_images(self, first=0, count=None):
        count = self._cnt(first, count)
        out = np.zeros((count, self.H, self.W), dtype=np.uint16)
        _check(lib().agf_rappids_get_images(self._h, out.ctypes.data, first, count))
        return out

    def set_states(self, vel0, acc0, grav, first=0):
        v, a, g = (np.ascontiguousarray(x, dtype=np.float64) for x in (vel0, acc0, grav))
        assert v.shape == a.shape == g.shape and v.shape[1] == 3
        _check(lib().agf_rappids_set_states(self._h, v.ctypes.data, a.ctypes.data, g.ctypes.data, first, v.shape[0]))

    def set_goals(self, goals, first=0):
        g = np.ascontiguousarray(goals, dtype=np.float64)
        _check(lib().agf_rappids_set_goals(self._h, g.ctypes.data, first, g.shape[0]))

    def set_candidates(self, cands, first=0):
        c = np.ascontiguousarray(cands, dtype=np.float64)
        assert c.ndim == 3 and c.shape[2] == 4
        _check(lib().agf_rappids_set_candidates(self._h, c.ctypes.data, c.shape[1], first, c.shape[0]))
        self.k = c.shape[1]

    def sample_candidates(self, k, seed, first_global_index=0):
        _check(lib().agf_rappids_sample_candidates(self._h, int(k), int(seed), int(first_global_index)))
        self.k = int(k)

    def get_candidates(self, first=0, count=None):
        count = self._cnt(first, count)
        out = np.zeros((count, self.k, 4))
        _check(lib().agf_rappids_get_candidates(self._h, out.ctypes.data, first, count))
        return out

    def plan(self):
        _check(lib().agf_rappids_plan(self._h))

    def sync(self):
        _check(lib().agf_rappids_sync(self._h))

    def set_dispatch(self, by_last_work=True):
        """Dispatch order of the planning pass: by the previous plan's work per vehicle (default) or index order."""
        _check(lib().agf_rappids_set_dispatch(self._h, 1 if by_last_work else 0))

    def plan_work(self, first=0, count=None):
        """Device clock cycles the last plan spent on each vehicle."""
        count = self._cnt(first, count)
        out = np.zeros(count, dtype=np.uint32)
        _check(lib().agf_rappids_get_plan_work(self._h, out.ctypes.data, first, count))
        return out

    def results(self, first=0, count=None):
        count = self._cnt(first, count)
        out = np.zeros(count, dtype=RESULT_DTYPE)
        _check(lib().agf_rappids_get_results(self._h, out.ctypes.data, first, count))
        return out

    def tracking_primitives(self, first=0, count=None):
        """[count][29] records for Batch.set_offboard_trajectories (trajAtt identity, trajOffset zero: fill in columns
        22:26 and 26:29)."""
        count = self._cnt(first, count)
        out = np.zeros((count, abi.OFFTRAJ_DOUBLES))
        _check(lib().agf_rappids_get_tracking_primitives(self._h, out.ctypes.data, first, count))
        return out

    def export_tracking_primitives(self, batch, att=None, offset=None, dst_first=0):
        """Writes the planned primitives into `batch`'s trajectory table on the device (no host hop)."""
        ptr, nv = C.c_void_p(), C.c_size_t()
        _check(batch.L.agf_batch_offboard_trajectories_device_ptr(batch.h, C.byref(ptr), C.byref(nv)))
        a = None if att is None else np.ascontiguousarray(att, dtype=np.float64).reshape(self.n, 4)
        o = None if offset is None else np.ascontiguousarray(offset, dtype=np.float64).reshape(self.n, 3)
        _check(lib().agf_rappids_export_tracking_primitives(self._h, ptr, nv.value, dst_first, None if a is None else a.ctypes.data,
                                                            None if o is None else o.ctypes.data))

    def candidate_flags(self, first=0, count=None):
        count = self._cnt(first, count)
        out = np.zeros((count, self.k), dtype=np.uint8)
        _check(lib().agf_rappids_get_candidate_flags(self._h, out.ctypes.data, first, count))
        return out

    def pyramids(self, first=0, count=None):
        count = self._cnt(first, count)
        out = np.zeros((count, abi.RAPPIDS_MAX_PYRAMIDS, abi.RAPPIDS_PYRAMID_DOUBLES))
        _check(lib().agf_rappids_get_pyramids(self._h, out.ctypes.data, first, count))
        return out

    def stats(self):
        out = np.zeros(8)
        _check(lib().agf_rappids_reduce_stats(self._h, out.ctypes.data))
        return dict(zip(("found", "generated", "cost_checks", "input_feasible", "velocity_admissible",
                         "collision_free", "pyramids", "sum_best_cost"), out.tolist()))

    def stats_device(self, dev_ptr):
        _check(lib().agf_rappids_reduce_stats_device(self._h, C.c_void_p(dev_ptr)))

    @property
    def stream(self):
        return lib().agf_rappids_stream(self._h)

    @property
    def launch_count(self):
        return int(lib().agf_rappids_launch_count(self._h))

    def plan_kernel_time(self):
        ms, n = C.c_double(), C.c_uint64()
        _check(lib().agf_rappids_plan_kernel_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


def build_info():
    return lib().agf_build_info().decode()
