/* agrifly_b200.h -- C ABI of the B200-native batched quadrotor simulation step.
 *
 * Drop-in boundary for ONE path of muellerlab/agri-fly: the per-vehicle
 * simulation step `Simulation::Quadcopter_T<Onboard::QuadcopterLogic>::Run()`
 * (Components/Components/Simulation/Quadcopter_T.cpp:86-203) together with the
 * object API it sits behind, `Simulation::SimulationObject6DOF`
 * (Components/Components/Simulation/SimulationObject6DOF.hpp:12-86).
 * The reference has no FFI layer for this path -- the C++ class API *is* the
 * operator API -- so every entry point below cites the reference member it
 * replaces.  `agri-fly_b200/host/agf_quadcopter.hpp` layers the reference's
 * method names (Run, Get/SetPosition, SetCommandRadioMsg, ...) on top of this
 * header so that a `Rappids_Simulator`-style loop compiles unchanged.
 *
 * Conventions: plain C, no exceptions cross the ABI, every function returns an
 * int (AGF_OK == 0, negative == error), opaque handle, caller-owned host
 * buffers, one handle per GPU, thread-compatible (external synchronisation per
 * handle).  There is NO CPU fallback: without a CUDA device agf_batch_create
 * fails with AGF_ENODEVICE.
 *
 * Units: SI.  Attitude is the reference's Rotation quaternion [w,x,y,z] with
 * <world vector> = att * <body vector> (Common/Common/Math/Rotation.hpp:28).
 */
#ifndef AGRIFLY_B200_H_
#define AGRIFLY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGF_VERSION_MAJOR 0
#define AGF_VERSION_MINOR 1

/* ---- error codes ------------------------------------------------------- */
#define AGF_OK 0
#define AGF_EINVAL (-1)       /* bad argument                                   */
#define AGF_ENOMEM (-2)       /* host or device allocation failed               */
#define AGF_ECUDA (-3)        /* CUDA runtime error (see agf_last_error_string) */
#define AGF_EUNSUPPORTED (-4) /* valid request this build cannot serve          */
#define AGF_ERANGE (-5)       /* index / count outside the batch                */
#define AGF_ENODEVICE (-6)    /* no CUDA device: there is no CPU fallback       */
#define AGF_ENCCL (-8)        /* NCCL unavailable or an NCCL call failed        */
#define AGF_EFULL (-7)        /* table full (reference returns -1 likewise:
                                 QuadcopterLogic.hpp:224-227)                   */

/* ---- enums mirrored from the reference --------------------------------- */
/* Onboard::QuadcopterConstants::QuadcopterType (QuadcopterConstants.hpp:16-24) */
enum {
  AGF_QC_TYPE_INVALID = 0,
  AGF_QC_TYPE_CF_STANDARD = 1,
  AGF_QC_TYPE_CF_BIGMOTORSPROPS = 2,
  AGF_QC_TYPE_CF_FEEDTHROUGH = 3,
  AGF_QC_TYPE_CF_LARGEQUAD = 4,
  AGF_QC_TYPE_CF_MINIQUAD = 5
};
/* Onboard::QuadcopterLogic::FlightState (QuadcopterLogic.hpp:145-154) */
enum {
  AGF_FS_UNINITIALIZED = 0,
  AGF_FS_IDLE = 1,
  AGF_FS_FULLY_AUTONOMOUS = 2,
  AGF_FS_PANIC = 3,
  AGF_FS_KILLED = 4,
  AGF_FS_EXTERNAL_ACCELERATION_CONTROL = 5,
  AGF_FS_EXTERNAL_RATES_CONTROL = 6
};
/* Onboard::PanicReason (PanicReason.hpp:5-14) */
enum {
  AGF_PANIC_NO_PANIC = 0,
  AGF_PANIC_ONBOARD_ESTIMATE_CRAZY = 1,
  AGF_PANIC_UWB_TIMEOUT = 2,
  AGF_PANIC_UPSIDE_DOWN = 3,
  AGF_PANIC_RADIO_CMD_TIMEOUT = 4,
  AGF_PANIC_LOW_BATTERY = 5,
  AGF_PANIC_KILLED_INTERNALLY = 6,
  AGF_PANIC_KILLED_EXTERNALLY = 7
};
/* RadioTypes::Type / ReservedFlags (RadioTypes.hpp:17-37) */
enum {
  AGF_RADIO_INVALID = 0,
  AGF_RADIO_RESERVED_FUTURE = 1,
  AGF_RADIO_EMERGENCY_KILL = 2,
  AGF_RADIO_POSITION_CMD = 3,
  AGF_RADIO_EXTERNAL_ACCELERATION_CMD = 4,
  AGF_RADIO_EXTERNAL_RATES_CMD = 5,
  AGF_RADIO_IDLE_CMD = 6
};
#define AGF_RADIO_FLAG_CALIBRATE_MOTORS 0x01
#define AGF_RADIO_FLAG_DISABLE_ONBOARD_SAFETY 0x02
#define AGF_RADIO_PACKET_SIZE 23 /* RadioMessageDecoded::RAW_PACKET_SIZE (RadioTypes.hpp:50) */
#define AGF_RADIO_NUM_FLOATS 10
#define AGF_TELEMETRY_PACKET_SIZE 30 /* sizeof(TelemetryPacket::data_packet_t) (TelemetryPacket.hpp:32-36) */
/* TelemetryPacket::TelemetryWarnings (TelemetryPacket.hpp:21-30) */
#define AGF_WARN_LOW_BATT 0x01
#define AGF_WARN_CMD_RATE 0x02
#define AGF_WARN_UWB_RESET 0x04
#define AGF_WARN_ONBOARD_FREQ 0x08
#define AGF_WARN_CMD_BATCH_DROP 0x10

#define AGF_MAX_UWB_ANCHORS 32 /* QuadcopterLogic::MAX_NUM_RANGING_TARGETS (QuadcopterLogic.hpp:346) */

/* ---- vehicle configuration --------------------------------------------- */
/* The onboard "firmware" constants, i.e. the members of
 * Onboard::QuadcopterConstants (QuadcopterConstants.hpp:334-367) that
 * QuadcopterLogic::Initialise reads (QuadcopterLogic.cpp:97-162).  All float,
 * as in the reference. */
typedef struct agf_logic_consts {
  float mass;
  float inertia_xx, inertia_zz; /* inertiaMatrix = diag(xx, xx, zz) (QuadcopterConstants.hpp:269-271) */
  float arm_length;
  float prop_thrust_from_speed_sqr;
  float prop_torque_from_thrust;
  float max_thrust_per_propeller, min_thrust_per_propeller;
  float max_cmd_total_thrust; /* < 0: mixer default 0.8*4*max (QuadcopterMixer.hpp:45-50) */
  int32_t prop0_spin_dir;
  float pos_control_nat_freq, pos_control_damping;
  float ang_vel_control_time_const_xy, att_control_time_const_xy;
  float ang_vel_control_time_const_z, att_control_time_const_z;
  float imu_yaw, imu_pitch, imu_roll;
  float low_battery_threshold;
  float lin_drag_coeff_b[3];
  float motor_time_const, motor_inertia, motor_min_speed, motor_max_speed;
  int32_t valid; /* 0: Initialise() latches FS_KILLED / PANIC_KILLED_INTERNALLY */
} agf_logic_consts;

/* One vehicle = the 15 constructor arguments of Simulation::Quadcopter_T
 * (Quadcopter_T.hpp:24-32) minus the timer, plus the firmware constants the
 * reference looks up from `quadcopterType` inside the constructor
 * (Quadcopter_T.cpp:71-82). */
typedef struct agf_vehicle_cfg {
  double mass;
  double inertia[9]; /* row major 3x3 */
  double arm_length;
  double com_error[3];
  double motor_min_speed, motor_max_speed;
  double prop_thrust_from_speed_sqr;
  double prop_torque_from_speed_sqr;
  double motor_time_const;
  double motor_inertia;
  double lin_drag_coeff_b[3];
  int32_t vehicle_id; /* uint8_t id in the reference                     */
  int32_t quad_type;  /* AGF_QC_TYPE_*                                   */
  agf_logic_consts logic; /* filled by agf_vehicle_cfg_from_type; may be edited (extension) */
} agf_vehicle_cfg;

/* QuadcopterConstants::GetVehicleTypeFromID (QuadcopterConstants.hpp:297-332) */
int agf_quad_type_from_id(unsigned id);
/* QuadcopterConstants::QuadcopterConstants(t) (QuadcopterConstants.hpp:31-274) */
int agf_logic_consts_from_type(int quad_type, agf_logic_consts* out);
/* The widening the reference's apps perform before calling the constructor
 * (Simulator/Rappids_Simulator/main.cpp:147-218): double(float table value),
 * inertia_yy = inertia_xx, propTorqueFromSpeedSqr = kTau*kF, comError = 0. */
int agf_vehicle_cfg_from_type(int quad_type, int vehicle_id, agf_vehicle_cfg* out);

/* ---- radio uplink codec (Common/Common/DataTypes/RadioTypes.hpp) -------- */
/* CreateRatesCommand :158 */
void agf_radio_encode_rates(uint8_t flags, float total_thrust, const float ang_vel[3],
                            uint8_t raw[AGF_RADIO_PACKET_SIZE]);
/* CreatePositionCommand :137 */
void agf_radio_encode_position(uint8_t flags, const float pos[3], const float vel[3],
                               const float acc[3], uint8_t raw[AGF_RADIO_PACKET_SIZE]);
/* CreateAccelerationCommand :173 */
void agf_radio_encode_acceleration(uint8_t flags, const float acc[3], float yaw_rate,
                                   uint8_t raw[AGF_RADIO_PACKET_SIZE]);
/* CreateIdleCommand :130 / CreateKillCommand :123 (float fields are zero-filled) */
void agf_radio_encode_idle(uint8_t flags, uint8_t raw[AGF_RADIO_PACKET_SIZE]);
void agf_radio_encode_kill(uint8_t flags, uint8_t raw[AGF_RADIO_PACKET_SIZE]);
/* RadioMessageDecoded(raw) :189-240 */
void agf_radio_decode(const uint8_t raw[AGF_RADIO_PACKET_SIZE], uint8_t* type, uint8_t* flags,
                      float floats[AGF_RADIO_NUM_FLOATS]);

/* ---- telemetry downlink codec (Common/Common/DataTypes/TelemetryPacket.hpp) */
typedef struct agf_telemetry { /* TelemetryPacket::TelemetryPacket :100-119 */
  uint8_t type, packet_number;
  float accel[3], gyro[3], motor_forces[4], position[3], batt_voltage;
  float velocity[3], attitude[3], debug_vals[6];
  uint8_t panic_reason, warnings;
} agf_telemetry;
/* DecodeTelemetryPacket :169-207; fills only the fields the packet type carries */
void agf_telemetry_decode(const uint8_t packet[AGF_TELEMETRY_PACKET_SIZE], agf_telemetry* out);

/* ---- log format: Rappids_Simulator's Logs/rappids_simulator/simulation.csv -------------
 * Header line (Simulator/Rappids_Simulator/main.cpp:266-270) and one row per offboard main-loop round (:676-733):
 * time, true position / velocity / attitude as Euler yaw-pitch-roll (Rotation.hpp:163-169) / angular velocity,
 * the four motor forces of the decoded telemetry packet, estimated position / velocity / Euler angles / angular
 * velocity (floats, :696-716), desired position and velocity, panic reason, last radio command.  Every value is
 * followed by a comma, numbers are written as the reference's ofstream writes them (6 significant digits). */
typedef struct agf_csv_record {
  double t;
  double pos[3], vel[3], att[4], ang_vel[3];
  float motor_forces[4];
  float est_pos[3], est_vel[3], est_att[4], est_ang_vel[3];
  double des_pos[3], des_vel[3];
  int32_t panic_reason;
  float last_radio_cmd[4];
} agf_csv_record;
/* Both return the length of the text (without the terminating NUL) and write at most cap bytes incl. the NUL. */
size_t agf_csv_header(char* buf, size_t cap);
size_t agf_csv_format_row(const agf_csv_record* r, char* buf, size_t cap);

/* ---- ROS message fields: what the hiperlab_rostools Simulator node publishes per vehicle ----------
 * simulator_truth (AIFS_ROS/hiperlab_rostools/msg/simulator_truth.msg, filled at Simulator/main.cpp:455-475) and telemetry
 * (msg/telemetry.msg, filled at Simulator/main.cpp:501-546 from the two decoded downlink packets; attitudeYPR through
 * Rotationf::FromVectorPartOfQuaternion + ToEulerYPR).  Plain structs with the messages' fields in order, headers left to the
 * caller (there is no ROS here). */
typedef struct agf_msg_simulator_truth {
  int64_t vehicleID;
  double posx, posy, posz, velx, vely, velz, attyaw, attpitch, attroll, attq0, attq1, attq2, attq3, angvelx, angvely, angvelz;
} agf_msg_simulator_truth;
typedef struct agf_msg_telemetry {
  uint8_t vehicleID, type, packetNumber, seqNum;
  double accelerometer[3], rateGyro[3], position[3], attitude[3], velocity[3], attitudeYPR[3], motorForces[4], debugVals[6],
      batteryVoltage;
  uint8_t panicReason, warnings;
} agf_msg_telemetry;
void agf_msg_simulator_truth_fill(int64_t vehicle_id, const double pos[3], const double vel[3], const double att[4],
                                  const double ang_vel[3], agf_msg_simulator_truth* out);
void agf_msg_telemetry_fill(const uint8_t packet1[AGF_TELEMETRY_PACKET_SIZE], const uint8_t packet2[AGF_TELEMETRY_PACKET_SIZE],
                            agf_msg_telemetry* out);

/* ---- the batched handle -------------------------------------------------- */
typedef struct agf_batch agf_batch;

/* arithmetic variants */
#define AGF_PREC_FP64 0 /* the reference's own mixed precision: double plant, float onboard logic */
#define AGF_PREC_FP32 1 /* plant in float as well (new; tolerance 1e-4)                         */
#define AGF_MATH_PARITY 0 /* no FMA contraction, agf_math.h shared libm: bit-comparable with the oracle */
#define AGF_MATH_FAST 1   /* FMA, CUDA fast paths                                                     */

typedef struct agf_batch_opts {
  int32_t device;    /* CUDA ordinal                                                     */
  int32_t precision; /* AGF_PREC_*                                                       */
  int32_t math;      /* AGF_MATH_*                                                       */
  int32_t block_threads; /* 0 = default                                                  */
  double onboard_logic_period; /* [s] Quadcopter_T ctor arg (Quadcopter_T.hpp:32); default 1/500 */
  double uwb_comm_period;      /* [s] UWBNetwork ctor arg (UWBNetwork.hpp:19); <= 0: no network    */
  /* IMU noise (Quadcopter_T.cpp:5-6: sigma_acc 0.2, sigma_gyro 0.1). Defaults are the
   * reference's values; a noise-free run sets both to 0.  Bias is an extension (0 = reference). */
  double sigma_acc, sigma_gyro;
  double bias_sigma_acc, bias_sigma_gyro;
  double uwb_noise_std_dev; /* UWBNetwork::SetNoiseProperties' noiseStdDev (UWBNetwork.hpp:28); outliers: agf_batch_set_uwb_noise */
  uint64_t seed;            /* Philox key                                                       */
  uint64_t first_global_index; /* index of vehicle 0 of this shard in the whole population (RNG counter) */
  void* stream;             /* cudaStream_t to launch on; NULL = a stream owned by the handle   */
  int32_t telemetry_warnings; /* 1: maintain TelemetryPacket warnings bitmask (default 1)         */
  int32_t reserved;
} agf_batch_opts;

void agf_batch_opts_default(agf_batch_opts* opts);

/* Quadcopter_T::Quadcopter_T (Quadcopter_T.cpp:9-83) for n_vehicles vehicles.
 * n_cfgs == 1: all vehicles share cfgs[0]; n_cfgs == n_vehicles: one each. */
int agf_batch_create(const agf_vehicle_cfg* cfgs, size_t n_cfgs, size_t n_vehicles,
                     const agf_batch_opts* opts, agf_batch** out);
int agf_batch_destroy(agf_batch* b);
size_t agf_batch_size(const agf_batch* b);
void* agf_batch_stream(const agf_batch* b);

/* `nticks` repetitions of [deliver scheduled radio commands] -> Run() -> [UWB network Run()]
 * -> clock += dt_us, for every vehicle: the loop body of
 * Simulator/Rappids_Simulator/main.cpp:391-392,737-739.  Asynchronous on the handle's stream. */
int agf_batch_run(agf_batch* b, uint32_t dt_us, uint32_t nticks);
/* The two halves of a tick for callers that own the clock, as the reference's apps do:
 * agf_batch_run(b, 0, 1) is exactly `Run()` at the current clock reading (a second call without an advance
 * in between returns early like Quadcopter_T.cpp:88-90), and agf_batch_advance_clock is
 * ManualTimer::AdvanceMicroSeconds (ManualTimer.hpp:29) for every stopwatch slaved to the batch's clock. */
int agf_batch_advance_clock(agf_batch* b, uint32_t dt_us);
int agf_batch_sync(agf_batch* b);
/* simulation clock (ManualTimer::GetMicroSeconds, ManualTimer.hpp:38) and ticks run so far */
uint64_t agf_batch_time_us(const agf_batch* b);
uint64_t agf_batch_ticks(const agf_batch* b);

/* state fields: SimulationObject6DOF getters/setters (SimulationObject6DOF.hpp:26-56) and the
 * read-outs of Quadcopter_T.hpp:39-83.  Host buffers are [count][ncomp], element type as listed. */
enum {
  AGF_F_POSITION = 0,      /* double[3]  Get/SetPosition                                  */
  AGF_F_VELOCITY = 1,      /* double[3]  Get/SetVelocity                                  */
  AGF_F_ATTITUDE = 2,      /* double[4]  Get/SetAttitude  [w,x,y,z]                       */
  AGF_F_ANGULAR_VELOCITY = 3, /* double[3]  Get/SetAngularVelocity                        */
  AGF_F_MOTOR_SPEED = 4,   /* double[4]  Motor::_speed (Motor.hpp:53)                     */
  AGF_F_MOTOR_SPEED_CMD = 5, /* float[4]   _motorSpeedCommands (Quadcopter_T.hpp:99)      */
  AGF_F_EST_POSITION = 6,  /* float[3]   GetEstimate (Quadcopter_T.hpp:53)                */
  AGF_F_EST_VELOCITY = 7,  /* float[3]                                                    */
  AGF_F_EST_ATTITUDE = 8,  /* float[4]                                                    */
  AGF_F_EST_ANGULAR_VELOCITY = 9, /* float[3]                                             */
  AGF_F_ACCELEROMETER = 10, /* float[3]   GetAccelerometer (LPF output; Quadcopter_T.hpp:74) */
  AGF_F_RATE_GYRO = 11,    /* float[3]   GetRateGyro                                      */
  AGF_F_FLIGHT_STATE = 12, /* int32[1]   QuadcopterLogic::GetFlightState                  */
  AGF_F_PANIC_REASON = 13, /* int32[1]   QuadcopterLogic::GetFirstPanicReason             */
  AGF_F_MOTOR_FORCE = 14,  /* double[4]  GetMotorForce(i) (Quadcopter_T.hpp:39; z thrust of the last Run) */
  AGF_F_EST_COVARIANCE = 15, /* float[81] KalmanFilter6DOF::_cov row major (read only)   */
  AGF_F_CYCLE_COUNTER = 16, /* int32[1]  QuadcopterLogic::GetCycleCounter                 */
  AGF_F_KF_COUNTERS = 17,  /* int32[4]  {numResets, numMeasRejected, uwbMeasCount, imuInit|uwbInit<<1} (read only) */
  AGF_F_DES_MOTOR_FORCE = 18, /* float[4] _desMotorForcesForTelemetry (read only)          */
  AGF_F_UWB_MEASUREMENT = 19, /* float[2] UWBRadio::_meas of the vehicle's radio: {range [m], index of the responding anchor in
                                 AddUWBRadioTarget order} (UWBRadio.hpp:17-94; read only)                       */
  AGF_F_COUNT_
};
int agf_batch_get_field(agf_batch* b, int field, void* host_dst, size_t first, size_t count);
int agf_batch_set_field(agf_batch* b, int field, const void* host_src, size_t first, size_t count);
/* SetPosition + SetVelocity + SetAttitude + SetAngularVelocity (SimulationObject6DOF.hpp:42-56) of a vehicle range in one
 * call and one host-to-device copy: state13 = [count][13] doubles, position 3, velocity 3, attitude w x y z, angular velocity 3. */
int agf_batch_set_state(agf_batch* b, const double* state13, size_t first, size_t count);
/* GetPosition + GetVelocity + GetAttitude + GetAngularVelocity (SimulationObject6DOF.hpp:26-40) of a vehicle range in one
 * call and one device-to-host copy, same [count][13] layout: a host that owns the population's state between steps reads
 * it with this and hands it back with agf_batch_set_state. */
int agf_batch_get_state(agf_batch* b, double* state13, size_t first, size_t count);
/* bytes per vehicle of a field's host representation */
size_t agf_field_size(int field);

/* UWBNetwork::SetNoiseProperties(noiseStdDev, outlierProbability, outlierStdDev) (UWBNetwork.hpp:28-33) for every vehicle's
 * private ranging network, with the reference's measurement model (UWBNetwork.cpp:62-75): with probability
 * outlier_probability a completed range is an outlier, range = N(0,1) * outlier_std_dev (NOT centred on the true distance),
 * otherwise range = true distance + N(0,1) * noise_std_dev.  Draws are counter-based Philox (vehicle, tick, stream 2), so they
 * do not depend on sharding or launch chunking; the reference's single global std::mt19937 is not reproduced. */
int agf_batch_set_uwb_noise(agf_batch* b, double noise_std_dev, double outlier_probability, double outlier_std_dev);

/* SimulationObject6DOF::SetCommandRadioMsg (SimulationObject6DOF.hpp:64): deliver now, i.e. before
 * the next Run().  raw is [count][23], or one packet for all vehicles in [first, first+count) when
 * broadcast != 0. */
int agf_batch_set_radio_cmd(agf_batch* b, const uint8_t* raw, size_t first, size_t count,
                            int broadcast);

/* In-kernel command delivery, replacing the host loop + CommunicationsDelay queue
 * (CommunicationsDelay.hpp:18-39; main.cpp:737-739).  An entry is delivered before the Run() of
 * absolute tick `tick` (tick 0 = first Run after create).  slot < 0: `raw` goes to every vehicle;
 * slot >= 0: vehicle i receives packet i of per-vehicle slot `slot` (agf_batch_set_cmd_slot).
 * Entries must be sorted by tick; at most one entry per tick (as the reference delivers at most
 * one message per tick). */
typedef struct agf_cmd_entry {
  uint32_t tick;
  int32_t slot;
  uint8_t raw[AGF_RADIO_PACKET_SIZE];
  uint8_t pad_;
} agf_cmd_entry;
int agf_batch_set_cmd_schedule(agf_batch* b, const agf_cmd_entry* entries, size_t n);
#define AGF_MAX_CMD_SLOTS 4
int agf_batch_set_cmd_slot(agf_batch* b, int slot, const uint8_t* raw /* [n_vehicles][23] */);

/* ---- offboard rates loop inside the kernel ---------------------------------
 * The control loop Rappids_Simulator wraps around Run() (Simulator/Rappids_Simulator/main.cpp:471-739,
 * controllerType == CTRL_OFFBOARD_RATES), executed per vehicle on the device so that populations fly the
 * reference's own demo without a host round trip per command:
 *   every `period_us` of simulation time (stopwatch with strict '>', main.cpp:471-476), right after Run() and
 *   the clock advance: state estimate -> Offboard::QuadcopterController::Run (QuadcopterController.cpp:11-74)
 *   -> RadioMessageDecoded::CreateRatesCommand (16-bit quantisation, RadioTypes.hpp:73-100,158-171) ->
 *   CommunicationsDelay queue (`delay_us`, CommunicationsDelay.hpp:18-39) -> SetCommandRadioMsg before the
 *   first Run() at or after the due time (main.cpp:737-739).
 * The state estimate is the true state at command generation (a MocapStateEstimator-equivalent predictor is a
 * later row); the desired position is piecewise constant in time (`targets`, sorted by time_us; the entry
 * with the largest time_us <= the simulation clock at command generation applies; before the first entry no
 * command is generated), optionally shifted per vehicle by `per_vehicle_offset`. */
#define AGF_OFFBOARD_QUEUE 4 /* commands in flight: ceil(delay_us / period_us) + 1 must not exceed this */
typedef struct agf_offboard_cfg {
  uint32_t period_us; /* periodOffboardMainLoop, main.cpp:175: 10000 */
  uint32_t delay_us;  /* timeDelayOffboardControlLoopTrue, main.cpp:178,282-283: 30000 */
  /* QuadcopterController::SetParameters (main.cpp:227-229) from the airframe table */
  float pos_control_nat_freq, pos_control_damping, att_control_time_const_xy, att_control_time_const_z;
  /* QuadcopterController::QuadcopterController (QuadcopterController.cpp:5-9): 0.5*9.81, 20, -1 */
  double min_vertical_proper_acc, max_proper_acc, min_proper_acc;
  double yaw_angle; /* desiredYawAngle [rad] (main.cpp:244,627) */
  uint32_t radio_flags;
  uint32_t reserved;
} agf_offboard_cfg;
typedef struct agf_offboard_target {
  uint64_t time_us;
  double pos[3];
} agf_offboard_target;
/* period/delay of Rappids_Simulator, controller parameters of the airframe `quad_type` */
int agf_offboard_cfg_default(int quad_type, agf_offboard_cfg* out);
/* cfg == NULL switches the loop off.  per_vehicle_offset: [n_vehicles][3] doubles or NULL. */
int agf_batch_set_offboard_loop(agf_batch* b, const agf_offboard_cfg* cfg, const agf_offboard_target* targets,
                                size_t n_targets, const double* per_vehicle_offset);

/* ---- offboard loop: reference generators, tracking controller ---------
 * What the desired position / velocity / acceleration handed to the offboard controller is, per vehicle and on the
 * device.  The default (AGF_OFFREF_TARGETS) is the piecewise-constant table of agf_batch_set_offboard_loop.
 *
 * AGF_OFFREF_STAGES -- the flight stages of the ROS rates-control node, ExampleVehicleStateMachine::Run
 *   (AIFS_ROS/hiperlab_rostools/src/QuadMocapRatesControl/ExampleVehicleStateMachine.cpp:93-370), evaluated once per
 *   offboard period: wait for start -> spool-up (rates command 0.25 g, 0.5 s, :122-153) -> take-off (2 s ramp from
 *   the position at stage entry to desired_pos, :162-189) -> flight (trajectory `traj_id` 0..5 of :211-287, blended
 *   in over 2 s, :289-294) -> landing after the stop signal (0.5 m/s, :300-323) -> complete (idle command,
 *   :326-343).  The start / stop signals (joystick buttons in the node) are the clock reaching start_us / stop_us.
 *   Safety net, battery check and ROS publishing are not part of the path; in the wait stage no command is sent.
 * AGF_OFFREF_TRAJECTORY -- Rappids_Simulator's tracking of a planned motion primitive (Simulator/Rappids_Simulator/
 *   main.cpp:560-618): until start_us the vehicle holds desired_pos with QuadcopterController::Run (:623-627); from
 *   then on it follows its own quintic (agf_batch_set_offboard_trajectories) with QuadcopterController::RunTracking
 *   (QuadcopterController.cpp:76-131; main.cpp:629-634): reference position / velocity / acceleration rotated by
 *   trajAtt and shifted by trajOffset, thrust and angular-velocity feed-forward from the primitive
 *   (RapidTrajectoryGenerator::GetThrust / GetOmega(t, 0.02), RapidTrajectoryGenerator.cpp:264-286), with the loop's
 *   own quirks kept: the sample time runs 0.04 s ahead of the tracking stopwatch while t < end (:562-563), and
 *   the z clamps behind the camera (:580-591). */
enum { AGF_OFFREF_TARGETS = 0, AGF_OFFREF_STAGES = 1, AGF_OFFREF_TRAJECTORY = 2 };
enum { AGF_STAGE_WAIT_FOR_START = 0, AGF_STAGE_SPOOL_UP = 1, AGF_STAGE_TAKEOFF = 2, AGF_STAGE_FLIGHT = 3,
       AGF_STAGE_LANDING = 4, AGF_STAGE_COMPLETE = 5, AGF_STAGE_EMERGENCY = 6 }; /* ExampleVehicleStateMachine.hpp FlightStage */
typedef struct agf_offboard_ref {
  int32_t kind;          /* AGF_OFFREF_* */
  int32_t traj_id;       /* STAGES: trajID of the flight stage (ExampleVehicleStateMachine.cpp:213), 0..5 */
  uint64_t start_us;     /* STAGES: start signal; TRAJECTORY: start of tracking (startFlightTime, main.cpp:141) */
  uint64_t stop_us;      /* STAGES: stop signal (UINT64_MAX: never) */
  double desired_pos[3]; /* _desiredPosition (QuadMocapRatesControl/main.cpp:82: (0,0,1)) / hover point (main.cpp:505) */
  double desired_yaw;    /* _desiredYawAngle [rad]; TRAJECTORY: desYawAngleDeg * pi / 180 (main.cpp:244) */
  /* STAGES: Offboard::SafetyNet (Components/Offboard/SafetyNet.hpp:52-106), checked on the state estimate in the spool-up,
   * take-off, flight and landing stages; a violation latches AGF_STAGE_EMERGENCY, which sends kill commands
   * (ExampleVehicleStateMachine.cpp:126-129,167-170,195-198,304-307,350-363).  safety_net == 0: off. */
  int32_t safety_net;
  int32_t reserved;
  double safe_min[3], safe_max[3]; /* lab-space corners: (-2.4,-3.1,-0.5) .. (1.8,3.1,4.5) */
  double min_normal_height;        /* below it the vehicle must point upwards: 1.0 */
  double not_seen_timeout;         /* [s] since the estimator's last accepted measurement: 0.5 */
} agf_offboard_ref;
/* SafetyNet's defaults into the four fields above, safety_net = 1 */
void agf_offboard_ref_safety_default(agf_offboard_ref* ref);
/* Needs agf_batch_set_offboard_loop first (period, delay, controller gains; its targets are ignored for the other
 * kinds, its per-vehicle offsets shift desired_pos).  ref == NULL: back to AGF_OFFREF_TARGETS. */
int agf_batch_set_offboard_reference(agf_batch* b, const agf_offboard_ref* ref);
/* One motion primitive per vehicle, [count][AGF_OFFTRAJ_DOUBLES]:
 *   [0..17]  per axis a = 0..2 at [6a..6a+5]: p0, v0, a0, alpha, beta, gamma of SingleAxisTrajectory
 *            (TrajectoryGenerator/SingleAxisTrajectory.hpp; position = p0 + v0 t + a0 t^2/2 + gamma t^3/6 + beta t^4/24
 *            + alpha t^5/120)
 *   [18..20] gravity in the trajectory frame (RapidTrajectoryGenerator constructor, main.cpp:495-497)
 *   [21]     end time tf
 *   [22..25] trajAtt quaternion w, x, y, z: world = trajAtt * trajectory frame (main.cpp:519)
 *   [26..28] trajOffset, world frame (main.cpp:522) */
#define AGF_OFFTRAJ_DOUBLES 29
int agf_batch_set_offboard_trajectories(agf_batch* b, const double* traj, size_t first, size_t count);
/* Device pointer of the trajectory records, [AGF_OFFTRAJ_DOUBLES][n_vehicles] doubles (field-major), allocated on first use:
 * for producers that already live on the GPU (agf_rappids_export_tracking_primitives, agrifly_b200_rappids.h). */
int agf_batch_offboard_trajectories_device_ptr(agf_batch* b, double** dev_ptr, size_t* n_vehicles);
/* Per-vehicle state of the reference generator, [count][AGF_OFFSTATE_DOUBLES]: stage, last stage, stage start
 * [us], position at take-off [3], last commanded position / velocity / acceleration [3 each], commanded yaw. */
#define AGF_OFFSTATE_DOUBLES 16
int agf_batch_get_offboard_state(agf_batch* b, double* out, size_t first, size_t count);

/* ---- offboard loop: state estimator ------------------------------------------
 * Where the offboard controller's state estimate comes from.  AGF_OFFEST_TRUTH: the true state at command
 * generation.  AGF_OFFEST_MOCAP: Offboard::MocapStateEstimator (Components/Offboard/MocapStateEstimator.cpp) as
 * Rappids_Simulator runs it (Simulator/Rappids_Simulator/main.cpp:221-225,451-457,468-469,652-654), per vehicle on
 * the device: every `mocap_period_us` (stopwatch, strict '>') UpdateWithMeasurement(true position, true attitude)
 * -- prediction through the queued commands (PredictionPipe.hpp), per-axis position/velocity and attitude/rate
 * Kalman filters with 2x2 covariances, 6-sigma measurement rejection, forced reset after 10 rejections -- and, at
 * command generation, GetPrediction(prediction_delay) (:61-118, incl. its use of the un-predicted velocity and
 * angular velocity for the position and attitude increments); after each command SetPredictedValues(cmdAngVel,
 * att * e3 * cmdThrust - g) queues it with the pipe's delay. */
enum { AGF_OFFEST_TRUTH = 0, AGF_OFFEST_MOCAP = 1 };
#define AGF_OFFEST_PIPE 8 /* prediction messages kept per vehicle; agf_batch_set_offboard_estimator returns AGF_EUNSUPPORTED when
                            ceil(prediction_delay / loop period) + ceil(mocap period / loop period) + 2 exceeds it (the reference's pipe is unbounded) */
typedef struct agf_offboard_estimator {
  int32_t kind;             /* AGF_OFFEST_* */
  uint32_t mocap_period_us; /* periodMocapSystem, main.cpp:174: 5000 */
  double prediction_delay;  /* timeDelayOffboardControlLoopEstimate, main.cpp:179,223-224,469: 0.03 s */
  /* MocapStateEstimator::MocapStateEstimator (MocapStateEstimator.cpp:20-32) */
  double meas_reject_dist, angvel_time_const;
  double meas_noise_pos, meas_noise_att, proc_noise_pos, proc_noise_att;
} agf_offboard_estimator;
int agf_offboard_estimator_default(agf_offboard_estimator* out);
/* Needs agf_batch_set_offboard_loop first; est == NULL or kind TRUTH: back to the true state.  The estimator objects
 * are constructed (Reset()) at the current clock reading. */
int agf_batch_set_offboard_estimator(agf_batch* b, const agf_offboard_estimator* est);
/* MocapStateEstimator::GetPrediction(horizon) now, [count][13]: position, velocity, attitude (w,x,y,z), angular
 * velocity; plus the counters [count][4] (may be NULL): initialised, measurements rejected, rejected consecutively,
 * messages in the prediction pipe. */
int agf_batch_get_offboard_estimate(agf_batch* b, double horizon, double* est13, double* counters4, size_t first, size_t count);

/* SimulationObject6DOF::GetTelemetryDataPackets (SimulationObject6DOF.hpp:67;
 * QuadcopterLogic.cpp:621-679), including its side effects (packet counter++, warnings cleared).
 * p1, p2: [count][30]. */
int agf_batch_get_telemetry(agf_batch* b, uint8_t* p1, uint8_t* p2, size_t first, size_t count);

/* Quadcopter_T::SetExternalForce / SetExternalTorque (Quadcopter_T.hpp:45-51), world frame.
 * force/torque: [count][3] or NULL to leave unchanged. */
int agf_batch_set_external_wrench(agf_batch* b, const double* force, const double* torque,
                                  size_t first, size_t count);

/* Quadcopter_T::AddUWBRadioTarget (Quadcopter_T.hpp:58) + an anchor radio at the same place in the
 * vehicle's private UWBNetwork.  Anchors are shared by all vehicles of the batch. */
int agf_batch_add_uwb_anchor(agf_batch* b, uint8_t id, const float pos[3]);

/* change noise parameters after creation (see agf_batch_opts) */
int agf_batch_set_noise(agf_batch* b, uint64_t seed, double sigma_gyro, double sigma_acc,
                        double bias_sigma_gyro, double bias_sigma_acc);

/* Trajectory logging to HBM (new capability; replaces the CSV logger of main.cpp:676-733).
 * Every `stride` ticks the step kernel appends one record per vehicle: 17 values {pos3 vel3 att4 angvel3 motorspeed4}
 * in the plant precision.  Device layout of a record (agf_batch_log_device_ptr): the first 16 values of a vehicle as
 * 16-byte vectors [quad][stride_vehicles] (float4: 4 quads; double2: 8), so that a warp's store instruction writes 512
 * contiguous bytes, followed by the 17th value as [stride_vehicles] scalars; value f of vehicle i of ring slot r is
 *   f < 16:  base[(r*17 + (f/L)*L) * stride_vehicles + i*L + f%L]   (L = 16 / elem_size lanes per vector)
 *   f == 16: base[(r*17 + 16) * stride_vehicles + i]
 * with stride_vehicles = the vehicle count rounded up to a multiple of 32.
 * capacity_records bounds the ring; agf_batch_read_log gathers a record on the device and copies it out. */
#define AGF_LOG_FIELDS 17
int agf_batch_enable_log(agf_batch* b, uint32_t stride_ticks, uint32_t capacity_records);
/* number of records written since enable (may exceed capacity: ring) */
uint64_t agf_batch_log_count(const agf_batch* b);
/* copies record `rec` (absolute index) for vehicles [first, first+count) as double[count][17] */
int agf_batch_read_log(agf_batch* b, uint64_t rec, double* host_dst, size_t first, size_t count);
/* raw device pointer, element size and stride_vehicles of the log ring, for consumers that stay on the GPU (any may be NULL) */
int agf_batch_log_device_ptr(agf_batch* b, void** dev_ptr, size_t* elem_size, size_t* stride_vehicles);

/* Monte-Carlo statistics (new capability): one launch reduces, over the vehicles of this handle,
 * the tracking error e = position - target, with warp shuffles -> one atomic per block.
 * target: [n][3] doubles on the host, or NULL for "the position command last delivered".
 * The device-side result is a vector of AGF_STATS_LEN doubles that sums across shards (all entries
 * except the last two, which combine with max), so the multi-GPU reduction is one
 * all-reduce(SUM) + one all-reduce(MAX) on that vector. */
#define AGF_STATS_LEN 16
enum {
  AGF_ST_COUNT = 0,      /* vehicles                                      */
  AGF_ST_SUM_EX = 1, AGF_ST_SUM_EY = 2, AGF_ST_SUM_EZ = 3,
  AGF_ST_SUM_E2 = 4,     /* sum |e|^2                                     */
  AGF_ST_SUM_ENORM = 5,  /* sum |e|                                       */
  AGF_ST_N_PANIC = 6,    /* vehicles in FS_PANIC                          */
  AGF_ST_N_KILLED = 7,
  AGF_ST_N_AUTONOMOUS = 8,
  AGF_ST_N_NONFINITE = 9, /* vehicles with a non-finite position          */
  AGF_ST_SUM_SPEED = 10,  /* sum |v|                                      */
  AGF_ST_SUM_EST_ERR = 11, /* sum |est_pos - pos|                         */
  AGF_ST_RESERVED12 = 12, AGF_ST_RESERVED13 = 13,
  AGF_ST_MAX_ENORM = 14,  /* max |e|   (combine with MAX)                 */
  AGF_ST_MAX_EST_ERR = 15 /* max |est_pos - pos| (combine with MAX)       */
};
/* dev_out: device pointer to AGF_STATS_LEN doubles (e.g. a torch tensor later handed to
 * torch.distributed / ncclAllReduce); asynchronous on the handle's stream. */
int agf_batch_reduce_stats_device(agf_batch* b, const double* host_target, double* dev_out);
int agf_batch_reduce_stats(agf_batch* b, const double* host_target, double host_out[AGF_STATS_LEN]);

/* ---- multi-GPU: the statistics read-out over NCCL --------------------------------
 * Vehicles shard by index range over one handle per GPU and nothing is exchanged inside the step; the one
 * collective of the product is this read-out.  `nccl_comm` is an ncclComm_t (passed as void* so that this header
 * needs no nccl.h) whose rank owns this batch's GPU, or NULL for a single GPU.  On the batch's stream: the
 * statistics kernel, ONE ncclAllGather of the AGF_STATS_LEN doubles of every rank, and a combine kernel (entries
 * below AGF_ST_MAX_ENORM are summed in rank order, the rest maximised), so every rank ends with the same bits.  From
 * the second call on with the same (communicator, output pointer, no host target) the three are replayed as one
 * captured CUDA graph.  Collective: every rank of the communicator must call it, one host thread per GPU.  NCCL keeps a
 * communicator alive while a graph that captured one of its collectives exists: destroy the batch (or use
 * agf_nccl_comm_destroy, which drops such graphs) BEFORE calling ncclCommDestroy on a communicator of your own.
 * AGF_STATS_GRAPH=0 in the environment keeps the read-out as three eager launches. */
int agf_batch_reduce_stats_nccl_device(agf_batch* b, void* nccl_comm, const double* host_target, double* dev_out);
int agf_batch_reduce_stats_nccl(agf_batch* b, void* nccl_comm, const double* host_target, double host_out[AGF_STATS_LEN]);
/* Communicator helpers for hosts that do not link NCCL themselves (libnccl.so.2 is resolved at run time; in a process
 * that already carries an NCCL, e.g. torch's, that one is used).  agf_nccl_comm_init_rank = ncclCommInitRank on `device`
 * with an id from agf_nccl_get_unique_id shared by the caller's own means (MPI, a file, torch.distributed);
 * agf_nccl_comm_init_all = ncclCommInitAll for one process driving several GPUs with one thread each. */
#define AGF_NCCL_UNIQUE_ID_BYTES 128
int agf_nccl_version(int* version);
int agf_nccl_get_unique_id(uint8_t id[AGF_NCCL_UNIQUE_ID_BYTES]);
int agf_nccl_comm_init_rank(const uint8_t id[AGF_NCCL_UNIQUE_ID_BYTES], int nranks, int rank, int device, void** comm_out);
int agf_nccl_comm_init_all(int ndev, const int* devices, void** comms_out /* [ndev] */);
/* ncclCommDestroy; first drops the read-out graphs that batches captured with this communicator (no read-out of those
 * batches may be in flight).  Destroy communicators created here before or after their batches, in any order. */
int agf_nccl_comm_destroy(void* comm);

/* number of kernel launches issued by this handle so far (bench.py's gpu_launches claim) */
uint64_t agf_batch_launch_count(const agf_batch* b);
/* device time of the step kernels launched since the last call, measured with CUDA events on the
 * handle's stream: *ms = sum of kernel durations, *launches = how many.  Synchronises. */
int agf_batch_step_kernel_time(agf_batch* b, double* ms, uint64_t* launches);

const char* agf_last_error_string(void);
const char* agf_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* AGRIFLY_B200_H_ */
