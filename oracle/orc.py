"""oracle/orc.py -- TEST INFRASTRUCTURE: ctypes driver for the CPU oracles (oracle/oracle_api.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this.  Flavours: ref-glibc / ref-shared (unmodified reference sources, oracle/_ref/*.so) and
port-glibc / port-shared (independent restatement, oracle/libagf_port_*.so).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _load_abi():
    # struct layouts come from the product's ABI mirror (a header-like module; loading it does not
    # load the CUDA library)
    import sys
    root = os.path.dirname(HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    import agrifly_b200
    return agrifly_b200.abi


abi = _load_abi()
NTRAJ = 40

# trajectory record columns
COL = dict(pos=slice(0, 3), vel=slice(3, 6), att=slice(6, 10), angvel=slice(10, 13),
           motor=slice(13, 17), motor_cmd=slice(17, 21), est_pos=slice(21, 24),
           est_vel=slice(24, 27), est_att=slice(27, 31), est_angvel=slice(31, 34),
           flight_state=34, panic=35, cycle=36, kf_resets=37, kf_rejected=38, uwb_count=39)


class OrcOpts(C.Structure):
    _fields_ = [("onboard_logic_period", C.c_double), ("uwb_comm_period", C.c_double),
                ("sigma_acc", C.c_double), ("sigma_gyro", C.c_double),
                ("uwb_noise_std_dev", C.c_double), ("uwb_outlier_probability", C.c_double), ("uwb_outlier_std_dev", C.c_double)]


class FullState(C.Structure):
    _fields_ = [
        ("pos", C.c_double * 3), ("vel", C.c_double * 3), ("att", C.c_double * 4),
        ("ang_vel", C.c_double * 3), ("motor_speed", C.c_double * 4),
        ("motor_force_z", C.c_double * 4), ("motor_speed_cmd", C.c_float * 4),
        ("flight_state", C.c_int32), ("first_panic_reason", C.c_int32),
        ("cycle_counter", C.c_int32), ("tel_warnings", C.c_int32),
        ("des_motor_speeds", C.c_float * 4), ("des_motor_forces", C.c_float * 4),
        ("gyro_lpf", (C.c_float * 3) * 4), ("acc_lpf", (C.c_float * 3) * 4),
        ("temp_lpf", C.c_float * 4), ("batt_lpf", C.c_float * 4),
        ("batt_voltage_filtered", C.c_float), ("monitor_cmd_rate_lpdt", C.c_float),
        ("monitor_main_loop_lpdt", C.c_float), ("des_pos", C.c_float * 3),
        ("radio_floats", C.c_float * 10), ("radio_type", C.c_int32), ("radio_flags", C.c_int32),
        ("radio_count", C.c_int32), ("uwb_meas_count", C.c_int32),
        ("next_ranging_target_idx", C.c_int32),
        ("kf_pos", C.c_float * 3), ("kf_vel", C.c_float * 3), ("kf_att", C.c_float * 4),
        ("kf_ang_vel", C.c_float * 3), ("kf_last_corr", C.c_float * 3), ("kf_cov", C.c_float * 81),
        ("kf_imu_init", C.c_int32), ("kf_uwb_init", C.c_int32), ("kf_num_resets", C.c_int32),
        ("kf_num_rejected", C.c_int32), ("kf_num_rejected_seq", C.c_int32), ("debug", C.c_float * 6),
    ]

    def as_dict(self):
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = np.array(v) if hasattr(v, "__len__") else v
        return out


PATHS = {
    "ref-glibc": os.path.join(HERE, "_ref", "libagf_ref_glibc.so"),
    "ref-shared": os.path.join(HERE, "_ref", "libagf_ref_shared.so"),
    "ref-fma": os.path.join(HERE, "_ref", "libagf_ref_fma.so"),
    "port-glibc": os.path.join(HERE, "libagf_port_glibc.so"),
    "port-shared": os.path.join(HERE, "libagf_port_shared.so"),
    "hostsim-shared": os.path.join(HERE, "libagf_hostsim_shared.so"),
    "hostsim-fast32": os.path.join(HERE, "libagf_hostsim_fast32.so"),
    "hostsim-fast64": os.path.join(HERE, "libagf_hostsim_fast64.so"),
}


def available(flavour):
    return os.path.exists(PATHS[flavour])


def make_schedule(entries):
    """entries: list of (tick, raw bytes[23] or None, slot) -> ctypes array of agf_cmd_entry."""
    arr = (abi.CmdEntry * max(1, len(entries)))()
    for i, e in enumerate(entries):
        tick, raw = e[0], e[1]
        slot = e[2] if len(e) > 2 else -1
        arr[i].tick = int(tick)
        arr[i].slot = int(slot)
        if raw is not None:
            C.memmove(arr[i].raw, bytes(raw), abi.RADIO_PACKET_SIZE)
    return arr


class Oracle:
    def __init__(self, flavour):
        self.flavour = flavour
        L = C.CDLL(PATHS[flavour])
        vp, P = C.c_void_p, C.POINTER
        L.orc_flavour.restype = C.c_char_p
        L.orc_create.restype = vp
        L.orc_create.argtypes = [P(abi.VehicleCfg), P(OrcOpts)]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_state.argtypes = [vp] + [P(C.c_double)] * 4
        L.orc_set_external.argtypes = [vp, P(C.c_double), P(C.c_double)]
        L.orc_add_anchor.argtypes = [vp, C.c_uint8, C.c_float, C.c_float, C.c_float]
        L.orc_add_anchor.restype = C.c_int
        L.orc_set_radio.argtypes = [vp, C.c_char_p]
        L.orc_run.argtypes = [vp, C.c_uint32, C.c_uint32, P(abi.CmdEntry), C.c_uint32, C.c_void_p, C.c_void_p]
        if True:
            L.orc_run_offboard.argtypes = [vp, C.c_uint32, C.c_uint32, P(abi.OffboardCfg), P(abi.OffboardTarget), C.c_uint32,
                                           C.c_void_p, C.c_void_p]
        if hasattr(L, "orc_run_offboard_ref"):
            L.orc_run_offboard_ref.argtypes = [vp, C.c_uint32, C.c_uint32, P(abi.OffboardCfg), P(abi.OffboardRef), C.c_void_p,
                                               C.c_void_p, C.c_void_p]
            L.orc_get_offboard_state.argtypes = [vp, C.c_void_p]
        if hasattr(L, "orc_run_stages_node"):
            L.orc_run_stages_node.argtypes = [vp, C.c_uint32, C.c_uint32, P(abi.OffboardCfg), P(abi.OffboardRef),
                                              P(abi.OffboardEstimator), C.c_int, C.c_void_p]
            L.orc_get_stages_node_state.argtypes = [vp, C.c_void_p]
        if hasattr(L, "orc_msg_telemetry"):
            L.orc_msg_telemetry.argtypes = [C.c_void_p, C.c_void_p, P(abi.MsgTelemetry)]
            L.orc_msg_simulator_truth.argtypes = [vp, P(abi.MsgSimulatorTruth)]
        if hasattr(L, "orc_csv_row"):
            L.orc_csv_row.restype = C.c_size_t
            L.orc_csv_row.argtypes = [P(abi.CsvRecord), C.c_char_p, C.c_size_t]
        if hasattr(L, "orc_set_offboard_estimator"):
            L.orc_set_offboard_estimator.argtypes = [vp, P(abi.OffboardEstimator)]
            L.orc_get_offboard_estimate.argtypes = [vp, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_get_full.argtypes = [vp, P(FullState)]
        L.orc_time_us.restype = C.c_uint64
        L.orc_time_us.argtypes = [vp]
        if not flavour.startswith("hostsim"):
            L.orc_get_telemetry.argtypes = [vp, C.c_void_p, C.c_void_p]
            L.orc_get_imu.argtypes = [vp, P(C.c_double), P(C.c_double)]
            L.orc_run_population.restype = C.c_double
            L.orc_run_population.argtypes = [P(abi.VehicleCfg), C.c_uint32, C.c_uint32, P(OrcOpts), C.c_void_p,
                                             C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, P(abi.CmdEntry),
                                             C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        if hasattr(L, "orc_run_population_traj"):
            if True:
                L.orc_run_population_traj.restype = C.c_double
                L.orc_run_population_traj.argtypes = [P(abi.VehicleCfg), C.c_uint32, C.c_uint32, P(OrcOpts), C.c_void_p,
                                                      C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, P(abi.CmdEntry),
                                                      C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        if flavour.startswith("ref-"):  # codec cross-checks exist only against the real reference
            L.orc_radio_encode_rates.argtypes = [C.c_uint8, C.c_float, P(C.c_float), C.c_void_p]
            L.orc_radio_encode_position.argtypes = [C.c_uint8, P(C.c_float), P(C.c_float), P(C.c_float), C.c_void_p]
            L.orc_radio_encode_acceleration.argtypes = [C.c_uint8, P(C.c_float), C.c_float, C.c_void_p]
            L.orc_telemetry_decode.argtypes = [C.c_void_p, P(abi.Telemetry)]
            L.orc_logic_consts.argtypes = [C.c_int, P(abi.LogicConsts)]
            if hasattr(L, "orc_vehicle_cfg_from_id"):
                L.orc_vehicle_cfg_from_id.argtypes = [C.c_int, P(abi.VehicleCfg)]
                L.orc_radio_encode_idle.argtypes = [C.c_uint8, C.c_void_p]
        if not flavour.startswith("hostsim"):
            L.orc_radio_decode.argtypes = [C.c_void_p, P(C.c_uint8), P(C.c_uint8), P(C.c_float)]
        self.L = L
        assert L.orc_flavour().decode() == flavour, (L.orc_flavour(), flavour)

    def vehicle(self, cfg, **kw):
        return OracleVehicle(self, cfg, **kw)

    def run_population(self, cfgs, n, init13=None, anchors=None, dt_us=2000, nticks=1, sched=(),
                       slot_raw=None, threads=1, onboard_logic_period=1.0 / 500.0, uwb_comm_period=0.0,
                       sigma_acc=0.0, sigma_gyro=0.0, uwb_noise_std_dev=0.0, uwb_outlier_probability=0.0, uwb_outlier_std_dev=0.0):
        opts = OrcOpts(onboard_logic_period, uwb_comm_period, sigma_acc, sigma_gyro, uwb_noise_std_dev, uwb_outlier_probability,
                       uwb_outlier_std_dev)
        if isinstance(cfgs, abi.VehicleCfg):
            carr = (abi.VehicleCfg * 1)(cfgs)
            ncfg = 1
        else:
            carr = (abi.VehicleCfg * len(cfgs))(*cfgs)
            ncfg = len(cfgs)
        out = np.zeros((n, NTRAJ))
        i13 = None if init13 is None else np.ascontiguousarray(init13, dtype=np.float64)
        anc = None if anchors is None else np.ascontiguousarray(anchors, dtype=np.float32)
        sr = None if slot_raw is None else np.ascontiguousarray(slot_raw, dtype=np.uint8)
        sch = make_schedule(list(sched))
        secs = self.L.orc_run_population(
            carr, ncfg, n, C.byref(opts), None if i13 is None else i13.ctypes.data,
            None if anc is None else anc.ctypes.data, 0 if anc is None else len(anc), dt_us, nticks,
            sch, len(sched), None if sr is None else sr.ctypes.data, threads, out.ctypes.data)
        return out, secs

    def run_population_traj(self, cfgs, n, stride, init13=None, anchors=None, dt_us=2000, nticks=1, sched=(),
                            slot_raw=None, threads=1, onboard_logic_period=1.0 / 500.0, uwb_comm_period=0.0,
                            sigma_acc=0.0, sigma_gyro=0.0):
        """run_population with a trajectory: -> ([nticks // stride][n][NTRAJ], seconds)"""
        opts = OrcOpts(onboard_logic_period, uwb_comm_period, sigma_acc, sigma_gyro, 0.0)
        if isinstance(cfgs, abi.VehicleCfg):
            carr = (abi.VehicleCfg * 1)(cfgs)
            ncfg = 1
        else:
            carr = (abi.VehicleCfg * len(cfgs))(*cfgs)
            ncfg = len(cfgs)
        out = np.zeros((nticks // stride, n, NTRAJ))
        i13 = None if init13 is None else np.ascontiguousarray(init13, dtype=np.float64)
        anc = None if anchors is None else np.ascontiguousarray(anchors, dtype=np.float32)
        sr = None if slot_raw is None else np.ascontiguousarray(slot_raw, dtype=np.uint8)
        sch = make_schedule(list(sched))
        secs = self.L.orc_run_population_traj(
            carr, ncfg, n, C.byref(opts), None if i13 is None else i13.ctypes.data,
            None if anc is None else anc.ctypes.data, 0 if anc is None else len(anc), dt_us, nticks,
            sch, len(sched), None if sr is None else sr.ctypes.data, threads, stride, out.ctypes.data)
        return out, secs


def reference_vehicle_cfg(O, vehicle_id=1, **overrides):
    """agf_vehicle_cfg for `vehicle_id` built by the reference's own QuadcopterConstants inside the harness (ref flavours),
    so that bench.py's reference arm loads no product library; the port has no table of its own and borrows the product's."""
    c = abi.VehicleCfg()
    if hasattr(O.L, "orc_vehicle_cfg_from_id"):
        if O.L.orc_vehicle_cfg_from_id(int(vehicle_id), C.byref(c)) != 0:
            raise ValueError("unknown vehicle id %r" % (vehicle_id,))
    else:
        import agrifly_b200
        c = agrifly_b200.vehicle_cfg(vehicle_id=vehicle_id)
    for k, v in overrides.items():
        setattr(c, k, v)
    return c


class RefCodec:
    """The radio command encoders of the reference itself (RadioTypes.hpp through the harness), with the call signatures of
    the product's codec object, for agrifly_b200.scenarios' schedule builders.  Port flavours fall back to the product codec."""

    def __init__(self, O):
        self.L = O.L
        self.ref = hasattr(O.L, "orc_radio_encode_idle")
        if not self.ref:
            import agrifly_b200
            self.fallback = agrifly_b200.codec

    @staticmethod
    def _f3(v):
        return (C.c_float * 3)(*[float(x) for x in v])

    def encode_position(self, flags, p, v=(0, 0, 0), a=(0, 0, 0)):
        if not self.ref:
            return self.fallback.encode_position(flags, p, v, a)
        raw = (C.c_uint8 * abi.RADIO_PACKET_SIZE)()
        self.L.orc_radio_encode_position(flags, self._f3(p), self._f3(v), self._f3(a), raw)
        return bytes(raw)

    def encode_rates(self, flags, thrust, w):
        if not self.ref:
            return self.fallback.encode_rates(flags, thrust, w)
        raw = (C.c_uint8 * abi.RADIO_PACKET_SIZE)()
        self.L.orc_radio_encode_rates(flags, float(thrust), self._f3(w), raw)
        return bytes(raw)

    def encode_idle(self, flags=0):
        if not self.ref:
            return self.fallback.encode_idle(flags)
        raw = (C.c_uint8 * abi.RADIO_PACKET_SIZE)()
        self.L.orc_radio_encode_idle(flags, raw)
        return bytes(raw)


class OracleVehicle:
    def __init__(self, orc, cfg, onboard_logic_period=1.0 / 500.0, uwb_comm_period=0.0,
                 sigma_acc=0.0, sigma_gyro=0.0, uwb_noise_std_dev=0.0, uwb_outlier_probability=0.0, uwb_outlier_std_dev=0.0):
        self.orc = orc
        self.L = orc.L
        opts = OrcOpts(onboard_logic_period, uwb_comm_period, sigma_acc, sigma_gyro, uwb_noise_std_dev, uwb_outlier_probability,
                       uwb_outlier_std_dev)
        self.h = self.L.orc_create(C.byref(cfg), C.byref(opts))

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, pos=(0, 0, 0), vel=(0, 0, 0), att=(1, 0, 0, 0), ang_vel=(0, 0, 0)):
        a = lambda x, n: (C.c_double * n)(*[float(v) for v in x])
        self.L.orc_set_state(self.h, a(pos, 3), a(vel, 3), a(att, 4), a(ang_vel, 3))

    def set_external(self, force=None, torque=None):
        a = lambda x: None if x is None else (C.c_double * 3)(*[float(v) for v in x])
        self.L.orc_set_external(self.h, a(force), a(torque))

    def add_anchor(self, id_, pos):
        return self.L.orc_add_anchor(self.h, id_, float(pos[0]), float(pos[1]), float(pos[2]))

    def set_radio(self, raw):
        self.L.orc_set_radio(self.h, bytes(raw))

    def run(self, nticks, dt_us=2000, sched=(), slot_raw=None, record=True):
        sch = make_schedule(list(sched))
        traj = np.zeros((nticks, NTRAJ)) if record else None
        sr = None
        if slot_raw is not None:
            sr = np.ascontiguousarray(slot_raw, dtype=np.uint8)
        self.L.orc_run(self.h, dt_us, nticks, sch, len(sched), None if sr is None else sr.ctypes.data,
                       None if traj is None else traj.ctypes.data)
        return traj

    def run_offboard(self, nticks, cfg, targets, offset=None, dt_us=2000, record=True):
        """targets: list of (time_us, (x, y, z)); cfg: abi.OffboardCfg"""
        tarr = (abi.OffboardTarget * max(1, len(targets)))()
        for i, (t, p) in enumerate(targets):
            tarr[i].time_us = int(t)
            tarr[i].pos[:] = [float(x) for x in p]
        traj = np.zeros((nticks, NTRAJ)) if record else None
        off = None if offset is None else np.ascontiguousarray(offset, dtype=np.float64)
        self.L.orc_run_offboard(self.h, dt_us, nticks, C.byref(cfg), tarr, len(targets),
                                None if off is None else off.ctypes.data, None if traj is None else traj.ctypes.data)
        return traj

    def run_offboard_ref(self, nticks, cfg, ref, offset=None, trajectory=None, dt_us=2000, record=True):
        """cfg: abi.OffboardCfg, ref: abi.OffboardRef, trajectory: [29] doubles (AGF_OFFREF_TRAJECTORY)"""
        traj = np.zeros((nticks, NTRAJ)) if record else None
        off = None if offset is None else np.ascontiguousarray(offset, dtype=np.float64)
        tr = None if trajectory is None else np.ascontiguousarray(trajectory, dtype=np.float64)
        self.L.orc_run_offboard_ref(self.h, dt_us, nticks, C.byref(cfg), C.byref(ref), None if off is None else off.ctypes.data,
                                    None if tr is None else tr.ctypes.data, None if traj is None else traj.ctypes.data)
        return traj

    def run_stages_node(self, nticks, cfg, ref, est, dt_us=2000):
        """the unmodified ExampleVehicleStateMachine in the loop (ref flavours only)"""
        traj = np.zeros((nticks, NTRAJ))
        self.L.orc_run_stages_node(self.h, dt_us, nticks, C.byref(cfg), C.byref(ref), C.byref(est), 3, traj.ctypes.data)
        return traj

    def stages_node_state(self):
        o = np.zeros(abi.OFFSTATE_DOUBLES)
        self.L.orc_get_stages_node_state(self.h, o.ctypes.data)
        return o

    def set_offboard_estimator(self, est):
        """est: abi.OffboardEstimator or None (true state)"""
        self.L.orc_set_offboard_estimator(self.h, None if est is None else C.byref(est))

    def offboard_estimate(self, horizon=0.0):
        e, c = np.zeros(13), np.zeros(4)
        self.L.orc_get_offboard_estimate(self.h, horizon, e.ctypes.data, c.ctypes.data)
        return e, c

    def offboard_state(self):
        o = np.zeros(abi.OFFSTATE_DOUBLES)
        self.L.orc_get_offboard_state(self.h, o.ctypes.data)
        return o

    def full(self):
        fs = FullState()
        self.L.orc_get_full(self.h, C.byref(fs))
        return fs.as_dict()

    def telemetry(self):
        p1 = np.zeros(30, np.uint8)
        p2 = np.zeros(30, np.uint8)
        self.L.orc_get_telemetry(self.h, p1.ctypes.data, p2.ctypes.data)
        return p1, p2

    def imu(self):
        a = (C.c_double * 3)()
        g = (C.c_double * 3)()
        self.L.orc_get_imu(self.h, a, g)
        return np.array(a), np.array(g)

    def time_us(self):
        return self.L.orc_time_us(self.h)
