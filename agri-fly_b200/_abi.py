"""ctypes mirror of include/agrifly_b200.h (struct layouts and prototypes)."""
import ctypes as C

RADIO_PACKET_SIZE = 23
TELEMETRY_PACKET_SIZE = 30
MAX_CMD_SLOTS = 4
STATS_LEN = 16
LOG_FIELDS = 17

OK, EINVAL, ENOMEM, ECUDA, EUNSUPPORTED, ERANGE, ENODEVICE, EFULL, ENCCL = 0, -1, -2, -3, -4, -5, -6, -7, -8
NCCL_UNIQUE_ID_BYTES = 128

QC_TYPE_INVALID, QC_TYPE_CF_STANDARD, QC_TYPE_CF_BIGMOTORSPROPS, QC_TYPE_CF_FEEDTHROUGH, \
    QC_TYPE_CF_LARGEQUAD, QC_TYPE_CF_MINIQUAD = range(6)
FS_UNINITIALIZED, FS_IDLE, FS_FULLY_AUTONOMOUS, FS_PANIC, FS_KILLED, \
    FS_EXTERNAL_ACCELERATION_CONTROL, FS_EXTERNAL_RATES_CONTROL = range(7)
PREC_FP64, PREC_FP32 = 0, 1
MATH_PARITY, MATH_FAST = 0, 1

# field ids -> (numpy dtype, ncomp)
FIELDS = {
    "position": (0, "f8", 3), "velocity": (1, "f8", 3), "attitude": (2, "f8", 4),
    "angular_velocity": (3, "f8", 3), "motor_speed": (4, "f8", 4), "motor_speed_cmd": (5, "f4", 4),
    "est_position": (6, "f4", 3), "est_velocity": (7, "f4", 3), "est_attitude": (8, "f4", 4),
    "est_angular_velocity": (9, "f4", 3), "accelerometer": (10, "f4", 3), "rate_gyro": (11, "f4", 3),
    "flight_state": (12, "i4", 1), "panic_reason": (13, "i4", 1), "motor_force": (14, "f8", 4),
    "est_covariance": (15, "f4", 81), "cycle_counter": (16, "i4", 1), "kf_counters": (17, "i4", 4),
    "des_motor_force": (18, "f4", 4), "uwb_measurement": (19, "f4", 2),
}


class LogicConsts(C.Structure):
    _fields_ = [
        ("mass", C.c_float), ("inertia_xx", C.c_float), ("inertia_zz", C.c_float),
        ("arm_length", C.c_float), ("prop_thrust_from_speed_sqr", C.c_float),
        ("prop_torque_from_thrust", C.c_float), ("max_thrust_per_propeller", C.c_float),
        ("min_thrust_per_propeller", C.c_float), ("max_cmd_total_thrust", C.c_float),
        ("prop0_spin_dir", C.c_int32), ("pos_control_nat_freq", C.c_float),
        ("pos_control_damping", C.c_float), ("ang_vel_control_time_const_xy", C.c_float),
        ("att_control_time_const_xy", C.c_float), ("ang_vel_control_time_const_z", C.c_float),
        ("att_control_time_const_z", C.c_float), ("imu_yaw", C.c_float), ("imu_pitch", C.c_float),
        ("imu_roll", C.c_float), ("low_battery_threshold", C.c_float),
        ("lin_drag_coeff_b", C.c_float * 3), ("motor_time_const", C.c_float),
        ("motor_inertia", C.c_float), ("motor_min_speed", C.c_float), ("motor_max_speed", C.c_float),
        ("valid", C.c_int32),
    ]


class VehicleCfg(C.Structure):
    _fields_ = [
        ("mass", C.c_double), ("inertia", C.c_double * 9), ("arm_length", C.c_double),
        ("com_error", C.c_double * 3), ("motor_min_speed", C.c_double),
        ("motor_max_speed", C.c_double), ("prop_thrust_from_speed_sqr", C.c_double),
        ("prop_torque_from_speed_sqr", C.c_double), ("motor_time_const", C.c_double),
        ("motor_inertia", C.c_double), ("lin_drag_coeff_b", C.c_double * 3),
        ("vehicle_id", C.c_int32), ("quad_type", C.c_int32), ("logic", LogicConsts),
    ]


class Telemetry(C.Structure):
    _fields_ = [
        ("type", C.c_uint8), ("packet_number", C.c_uint8), ("accel", C.c_float * 3),
        ("gyro", C.c_float * 3), ("motor_forces", C.c_float * 4), ("position", C.c_float * 3),
        ("batt_voltage", C.c_float), ("velocity", C.c_float * 3), ("attitude", C.c_float * 3),
        ("debug_vals", C.c_float * 6), ("panic_reason", C.c_uint8), ("warnings", C.c_uint8),
    ]


class BatchOpts(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("precision", C.c_int32), ("math", C.c_int32),
        ("block_threads", C.c_int32), ("onboard_logic_period", C.c_double),
        ("uwb_comm_period", C.c_double), ("sigma_acc", C.c_double), ("sigma_gyro", C.c_double),
        ("bias_sigma_acc", C.c_double), ("bias_sigma_gyro", C.c_double),
        ("uwb_noise_std_dev", C.c_double), ("seed", C.c_uint64), ("first_global_index", C.c_uint64),
        ("stream", C.c_void_p), ("telemetry_warnings", C.c_int32), ("reserved", C.c_int32),
    ]


class CmdEntry(C.Structure):
    _fields_ = [("tick", C.c_uint32), ("slot", C.c_int32), ("raw", C.c_uint8 * RADIO_PACKET_SIZE),
                ("pad_", C.c_uint8)]


OFFBOARD_QUEUE = 4


class OffboardCfg(C.Structure):
    _fields_ = [("period_us", C.c_uint32), ("delay_us", C.c_uint32),
                ("pos_control_nat_freq", C.c_float), ("pos_control_damping", C.c_float),
                ("att_control_time_const_xy", C.c_float), ("att_control_time_const_z", C.c_float),
                ("min_vertical_proper_acc", C.c_double), ("max_proper_acc", C.c_double), ("min_proper_acc", C.c_double),
                ("yaw_angle", C.c_double), ("radio_flags", C.c_uint32), ("reserved", C.c_uint32)]


class OffboardTarget(C.Structure):
    _fields_ = [("time_us", C.c_uint64), ("pos", C.c_double * 3)]


OFFREF_TARGETS, OFFREF_STAGES, OFFREF_TRAJECTORY = 0, 1, 2
STAGE_WAIT_FOR_START, STAGE_SPOOL_UP, STAGE_TAKEOFF, STAGE_FLIGHT, STAGE_LANDING, STAGE_COMPLETE, STAGE_EMERGENCY = range(7)
OFFTRAJ_DOUBLES, OFFSTATE_DOUBLES = 29, 16


class MsgSimulatorTruth(C.Structure):
    _fields_ = [("vehicleID", C.c_int64)] + [(k, C.c_double) for k in (
        "posx", "posy", "posz", "velx", "vely", "velz", "attyaw", "attpitch", "attroll", "attq0", "attq1", "attq2", "attq3",
        "angvelx", "angvely", "angvelz")]


class MsgTelemetry(C.Structure):
    _fields_ = [("vehicleID", C.c_uint8), ("type", C.c_uint8), ("packetNumber", C.c_uint8), ("seqNum", C.c_uint8),
                ("accelerometer", C.c_double * 3), ("rateGyro", C.c_double * 3), ("position", C.c_double * 3),
                ("attitude", C.c_double * 3), ("velocity", C.c_double * 3), ("attitudeYPR", C.c_double * 3),
                ("motorForces", C.c_double * 4), ("debugVals", C.c_double * 6), ("batteryVoltage", C.c_double),
                ("panicReason", C.c_uint8), ("warnings", C.c_uint8)]


class CsvRecord(C.Structure):
    _fields_ = [("t", C.c_double), ("pos", C.c_double * 3), ("vel", C.c_double * 3), ("att", C.c_double * 4),
                ("ang_vel", C.c_double * 3), ("motor_forces", C.c_float * 4), ("est_pos", C.c_float * 3),
                ("est_vel", C.c_float * 3), ("est_att", C.c_float * 4), ("est_ang_vel", C.c_float * 3),
                ("des_pos", C.c_double * 3), ("des_vel", C.c_double * 3), ("panic_reason", C.c_int32),
                ("last_radio_cmd", C.c_float * 4)]


OFFEST_TRUTH, OFFEST_MOCAP = 0, 1


class OffboardEstimator(C.Structure):
    _fields_ = [("kind", C.c_int32), ("mocap_period_us", C.c_uint32), ("prediction_delay", C.c_double),
                ("meas_reject_dist", C.c_double), ("angvel_time_const", C.c_double),
                ("meas_noise_pos", C.c_double), ("meas_noise_att", C.c_double),
                ("proc_noise_pos", C.c_double), ("proc_noise_att", C.c_double)]


class OffboardRef(C.Structure):
    _fields_ = [("kind", C.c_int32), ("traj_id", C.c_int32), ("start_us", C.c_uint64), ("stop_us", C.c_uint64),
                ("desired_pos", C.c_double * 3), ("desired_yaw", C.c_double), ("safety_net", C.c_int32), ("reserved", C.c_int32),
                ("safe_min", C.c_double * 3), ("safe_max", C.c_double * 3), ("min_normal_height", C.c_double),
                ("not_seen_timeout", C.c_double)]


# ---- include/agrifly_b200_rappids.h ------------------------------------------------------------
RAPPIDS_MAX_PYRAMIDS = 32
RAPPIDS_PYRAMID_DOUBLES = 17
RAPPIDS_MAX_BOXES = 4
RAPPIDS_COST_DIRECTION, RAPPIDS_COST_GOAL = 0, 1
RAPPIDS_LOW_COST, RAPPIDS_DYNAMICS_FEASIBLE, RAPPIDS_VELOCITY_ADMISSIBLE, RAPPIDS_COLLISION_FREE = 1, 2, 4, 8


class RappidsCfg(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("depth_scale", C.c_double),
                ("focal_length", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("true_radius", C.c_double), ("planning_radius", C.c_double), ("min_checking_dist", C.c_double),
                ("min_thrust", C.c_double), ("max_thrust", C.c_double), ("max_angvel", C.c_double),
                ("min_section_time", C.c_double), ("max_velocity", C.c_double),
                ("max_pyramids", C.c_int32), ("cost_kind", C.c_int32), ("cost_vec", C.c_double * 3),
                ("sample_min_x", C.c_double), ("sample_max_x", C.c_double), ("sample_min_y", C.c_double),
                ("sample_max_y", C.c_double), ("sample_min_depth", C.c_double), ("sample_max_depth", C.c_double),
                ("sample_min_time", C.c_double), ("sample_max_time", C.c_double),
                ("math", C.c_int32), ("device", C.c_int32)]


class RappidsResult(C.Structure):
    _fields_ = [("found", C.c_int32), ("best_index", C.c_int32), ("n_generated", C.c_int32),
                ("n_cost_checks", C.c_int32), ("n_collision_checks", C.c_int32), ("n_velocity_checks", C.c_int32),
                ("n_collision_free", C.c_int32), ("n_pyramids", C.c_int32), ("best_cost", C.c_double),
                ("best_coeffs", C.c_double * 18), ("best_tf", C.c_double),
                ("pyramid_cap_hit", C.c_int32), ("reserved_", C.c_int32)]


# every symbol include/agrifly_b200.h and include/agrifly_b200_rappids.h declare: name -> (restype, argtypes)
_P = C.POINTER
PROTOTYPES = {
    "agf_rappids_cfg_default": (C.c_int, [C.c_int32, C.c_int32, _P(RappidsCfg)]),
    "agf_rappids_create": (C.c_int, [_P(RappidsCfg), C.c_size_t, C.c_int32, _P(C.c_void_p)]),
    "agf_rappids_destroy": (C.c_int, [C.c_void_p]),
    "agf_rappids_size": (C.c_size_t, [C.c_void_p]),
    "agf_rappids_stream": (C.c_void_p, [C.c_void_p]),
    "agf_rappids_set_images": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_rappids_render_scenes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_rappids_get_images": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_rappids_set_states": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_rappids_set_goals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_rappids_set_candidates": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_size_t, C.c_size_t]),
    "agf_rappids_sample_candidates": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.c_uint64]),
    "agf_rappids_get_candidates": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_rappids_plan": (C.c_int, [C.c_void_p]),
    "agf_rappids_sync": (C.c_int, [C.c_void_p]),
    "agf_rappids_get_results": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_rappids_get_candidate_flags": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_rappids_get_tracking_primitives": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_rappids_export_tracking_primitives": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]),
    "agf_rappids_get_pyramids": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_rappids_reduce_stats": (C.c_int, [C.c_void_p, C.c_void_p]),
    "agf_rappids_reduce_stats_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "agf_rappids_plan_kernel_time": (C.c_int, [C.c_void_p, _P(C.c_double), _P(C.c_uint64)]),
    "agf_rappids_launch_count": (C.c_uint64, [C.c_void_p]),
    "agf_rappids_set_dispatch": (C.c_int, [C.c_void_p, C.c_int32]),
    "agf_rappids_get_plan_work": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_quad_type_from_id": (C.c_int, [C.c_uint]),
    "agf_logic_consts_from_type": (C.c_int, [C.c_int, _P(LogicConsts)]),
    "agf_vehicle_cfg_from_type": (C.c_int, [C.c_int, C.c_int, _P(VehicleCfg)]),
    "agf_radio_encode_rates": (None, [C.c_uint8, C.c_float, _P(C.c_float), _P(C.c_uint8)]),
    "agf_radio_encode_position": (None, [C.c_uint8, _P(C.c_float), _P(C.c_float), _P(C.c_float), _P(C.c_uint8)]),
    "agf_radio_encode_acceleration": (None, [C.c_uint8, _P(C.c_float), C.c_float, _P(C.c_uint8)]),
    "agf_radio_encode_idle": (None, [C.c_uint8, _P(C.c_uint8)]),
    "agf_radio_encode_kill": (None, [C.c_uint8, _P(C.c_uint8)]),
    "agf_radio_decode": (None, [_P(C.c_uint8), _P(C.c_uint8), _P(C.c_uint8), _P(C.c_float)]),
    "agf_telemetry_decode": (None, [_P(C.c_uint8), _P(Telemetry)]),
    "agf_batch_opts_default": (None, [_P(BatchOpts)]),
    "agf_batch_create": (C.c_int, [_P(VehicleCfg), C.c_size_t, C.c_size_t, _P(BatchOpts), _P(C.c_void_p)]),
    "agf_batch_destroy": (C.c_int, [C.c_void_p]),
    "agf_batch_size": (C.c_size_t, [C.c_void_p]),
    "agf_batch_stream": (C.c_void_p, [C.c_void_p]),
    "agf_batch_run": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "agf_batch_advance_clock": (C.c_int, [C.c_void_p, C.c_uint32]),
    "agf_batch_sync": (C.c_int, [C.c_void_p]),
    "agf_batch_time_us": (C.c_uint64, [C.c_void_p]),
    "agf_batch_ticks": (C.c_uint64, [C.c_void_p]),
    "agf_batch_get_field": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_batch_set_field": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_field_size": (C.c_size_t, [C.c_int]),
    "agf_batch_set_radio_cmd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]),
    "agf_batch_set_cmd_schedule": (C.c_int, [C.c_void_p, _P(CmdEntry), C.c_size_t]),
    "agf_batch_set_cmd_slot": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "agf_offboard_cfg_default": (C.c_int, [C.c_int, _P(OffboardCfg)]),
    "agf_batch_set_offboard_loop": (C.c_int, [C.c_void_p, _P(OffboardCfg), _P(OffboardTarget), C.c_size_t, C.c_void_p]),
    "agf_offboard_ref_safety_default": (None, [_P(OffboardRef)]),
    "agf_batch_set_offboard_reference": (C.c_int, [C.c_void_p, _P(OffboardRef)]),
    "agf_batch_set_offboard_trajectories": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_batch_offboard_trajectories_device_ptr": (C.c_int, [C.c_void_p, _P(C.c_void_p), _P(C.c_size_t)]),
    "agf_batch_get_offboard_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_msg_simulator_truth_fill": (None, [C.c_int64, _P(C.c_double), _P(C.c_double), _P(C.c_double), _P(C.c_double),
                                            _P(MsgSimulatorTruth)]),
    "agf_msg_telemetry_fill": (None, [C.c_void_p, C.c_void_p, _P(MsgTelemetry)]),
    "agf_csv_header": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "agf_csv_format_row": (C.c_size_t, [_P(CsvRecord), C.c_char_p, C.c_size_t]),
    "agf_batch_set_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_batch_get_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_offboard_estimator_default": (C.c_int, [_P(OffboardEstimator)]),
    "agf_batch_set_offboard_estimator": (C.c_int, [C.c_void_p, _P(OffboardEstimator)]),
    "agf_batch_get_offboard_estimate": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_batch_get_telemetry": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_batch_set_external_wrench": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_batch_add_uwb_anchor": (C.c_int, [C.c_void_p, C.c_uint8, _P(C.c_float)]),
    "agf_batch_set_noise": (C.c_int, [C.c_void_p, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_double]),
    "agf_batch_enable_log": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "agf_batch_log_count": (C.c_uint64, [C.c_void_p]),
    "agf_batch_read_log": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.c_size_t]),
    "agf_batch_log_device_ptr": (C.c_int, [C.c_void_p, _P(C.c_void_p), _P(C.c_size_t), _P(C.c_size_t)]),
    "agf_batch_reduce_stats_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "agf_batch_reduce_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "agf_batch_reduce_stats_nccl_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agf_batch_reduce_stats_nccl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agf_nccl_version": (C.c_int, [_P(C.c_int)]),
    "agf_nccl_get_unique_id": (C.c_int, [C.c_void_p]),
    "agf_nccl_comm_init_rank": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _P(C.c_void_p)]),
    "agf_nccl_comm_init_all": (C.c_int, [C.c_int, _P(C.c_int), _P(C.c_void_p)]),
    "agf_nccl_comm_destroy": (C.c_int, [C.c_void_p]),
    "agf_batch_set_uwb_noise": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double]),
    "agf_batch_launch_count": (C.c_uint64, [C.c_void_p]),
    "agf_batch_step_kernel_time": (C.c_int, [C.c_void_p, _P(C.c_double), _P(C.c_uint64)]),
    "agf_last_error_string": (C.c_char_p, []),
    "agf_build_info": (C.c_char_p, []),
}


def bind(lib):
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
