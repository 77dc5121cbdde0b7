/* oracle/oracle_api.h -- TEST INFRASTRUCTURE, never linked into the product.
 *
 * One C interface, two implementations:
 *   oracle/_ref/libagf_ref_{glibc,shared}.so  the UNMODIFIED reference sources compiled from
 *        /root/reference with the header shims in oracle/shim (oracle/ref_harness.cpp), and
 *   oracle/libagf_port_{glibc,shared}.so      the independent literal CPU restatement
 *        (oracle/port/agf_port.cpp).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load these.  `glibc` = the platform libm, `shared` = libm calls redirected to
 * agri-fly_b200/csrc/agf_math.h (see DESIGN.md "Parity definition").
 *
 * Tick semantics (identical to agf_batch_run):  for each tick
 *     [deliver the schedule entry for this tick]  ->  vehicle.Run()  ->  uwbNetwork.Run() (if any)
 *     -> record  ->  clock += dt_us
 * which is the loop body of Simulator/Rappids_Simulator/main.cpp:391-392,737-739.
 */
#ifndef AGF_ORACLE_API_H_
#define AGF_ORACLE_API_H_

#include <stdint.h>
#include "agrifly_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_vehicle orc_vehicle;

typedef struct orc_opts {
  double onboard_logic_period; /* Quadcopter_T ctor arg */
  double uwb_comm_period;      /* <= 0: no UWB network  */
  double sigma_acc, sigma_gyro; /* IMU noise std dev (reference: 0.2 / 0.1, Quadcopter_T.cpp:5-6) */
  double uwb_noise_std_dev;
  double uwb_outlier_probability, uwb_outlier_std_dev; /* UWBNetwork::SetNoiseProperties (ref flavours only) */
} orc_opts;

/* everything observable, for deep parity checks */
typedef struct orc_full_state {
  double pos[3], vel[3], att[4], ang_vel[3];
  double motor_speed[4];
  double motor_force_z[4];
  float motor_speed_cmd[4];
  /* logic */
  int32_t flight_state, first_panic_reason, cycle_counter, tel_warnings;
  float des_motor_speeds[4], des_motor_forces[4];
  float gyro_lpf[4][3]; /* xm0, xm1, ym0, ym1 */
  float acc_lpf[4][3];
  float temp_lpf[4], batt_lpf[4];
  float batt_voltage_filtered;
  float monitor_cmd_rate_lpdt, monitor_main_loop_lpdt;
  float des_pos[3];
  float radio_floats[10];
  int32_t radio_type, radio_flags, radio_count;
  int32_t uwb_meas_count, next_ranging_target_idx;
  /* estimator */
  float kf_pos[3], kf_vel[3], kf_att[4], kf_ang_vel[3], kf_last_corr[3];
  float kf_cov[81];
  int32_t kf_imu_init, kf_uwb_init, kf_num_resets, kf_num_rejected, kf_num_rejected_seq;
  float debug[6];
} orc_full_state;

/* per-tick trajectory record: 40 doubles */
#define ORC_NTRAJ 40
/* 0-2 pos, 3-5 vel, 6-9 att, 10-12 angvel, 13-16 motor speed, 17-20 motor speed cmd,
 * 21-23 est pos, 24-26 est vel, 27-30 est att, 31-33 est angvel, 34 flight state, 35 panic,
 * 36 cycle counter, 37 kf resets, 38 kf rejected, 39 uwb count */

const char* orc_flavour(void); /* "ref-glibc", "ref-shared", "port-glibc", "port-shared" */

orc_vehicle* orc_create(const agf_vehicle_cfg* cfg, const orc_opts* opts);
void orc_destroy(orc_vehicle* v);
void orc_set_state(orc_vehicle* v, const double pos[3], const double vel[3], const double att[4],
                   const double ang_vel[3]);
void orc_set_external(orc_vehicle* v, const double force[3], const double torque[3]);
int orc_add_anchor(orc_vehicle* v, uint8_t id, float x, float y, float z);
void orc_set_radio(orc_vehicle* v, const uint8_t raw[AGF_RADIO_PACKET_SIZE]);
/* slot >= 0 entries use `slot_raw` (one packet: this vehicle's entry of that slot) */
void orc_run(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_cmd_entry* sched,
             uint32_t nsched, const uint8_t* slot_raw /* [AGF_MAX_CMD_SLOTS][23] or NULL */,
             double* traj /* [nticks][ORC_NTRAJ] or NULL */);
/* The same ticks with the offboard rates loop of Simulator/Rappids_Simulator/main.cpp:471-739 around Run()
 * (semantics: include/agrifly_b200.h "offboard rates loop"): truth-fed QuadcopterController::Run ->
 * CreateRatesCommand -> CommunicationsDelay -> SetCommandRadioMsg.  The loop's stopwatch and queue live in the
 * vehicle object, so consecutive calls continue seamlessly.  offset: this vehicle's [3] shift or NULL. */
void orc_run_offboard(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_offboard_cfg* cfg,
                      const agf_offboard_target* targets, uint32_t n_targets, const double* offset,
                      double* traj /* [nticks][ORC_NTRAJ] or NULL */);
/* The same loop with a reference generator (agf_offboard_ref: flight stages of the ROS rates-control node, or
 * Rappids_Simulator's tracking of a motion primitive `tr` [AGF_OFFTRAJ_DOUBLES] with RunTracking). */
void orc_run_offboard_ref(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_offboard_cfg* cfg,
                          const agf_offboard_ref* ref, const double* offset, const double* tr,
                          double* traj /* [nticks][ORC_NTRAJ] or NULL */);
void orc_get_offboard_state(orc_vehicle* v, double* out /* [AGF_OFFSTATE_DOUBLES] */);
/* ref flavours only: the unmodified ExampleVehicleStateMachine of the ROS rates-control node in the same loop (pins the
 * restated stage logic); traj_id must be the value compiled into the node (ExampleVehicleStateMachine.cpp:213: 3) */
void orc_run_stages_node(orc_vehicle* v, uint32_t dt_us, uint32_t nticks, const agf_offboard_cfg* cfg, const agf_offboard_ref* ref,
                         const agf_offboard_estimator* est, int traj_id_check, double* traj);
void orc_get_stages_node_state(orc_vehicle* v, double* out /* [AGF_OFFSTATE_DOUBLES] */);
/* Offboard::MocapStateEstimator in the loop (agf_offboard_estimator); NULL: back to the true state */
void orc_set_offboard_estimator(orc_vehicle* v, const agf_offboard_estimator* est);
void orc_get_offboard_estimate(orc_vehicle* v, double horizon, double* est13, double* counters4 /* or NULL */);
/* ROS telemetry / simulator_truth message fields through the reference's own types (ref flavours only) */
void orc_msg_telemetry(const uint8_t* raw1, const uint8_t* raw2, agf_msg_telemetry* out);
void orc_msg_simulator_truth(orc_vehicle* v, agf_msg_simulator_truth* out);
/* simulation.csv row through the reference's own stream operators and Euler conversion (ref flavours only) */
size_t orc_csv_row(const agf_csv_record* r, char* buf, size_t cap);
void orc_get_full(orc_vehicle* v, orc_full_state* out);
void orc_get_telemetry(orc_vehicle* v, uint8_t p1[AGF_TELEMETRY_PACKET_SIZE],
                       uint8_t p2[AGF_TELEMETRY_PACKET_SIZE]);
void orc_get_imu(orc_vehicle* v, double acc[3], double gyro[3]);
uint64_t orc_time_us(orc_vehicle* v);

/* population helper for the CPU baseline: creates n vehicles (cfgs: 1 or n), applies init
 * states ([n][13]: pos3 vel3 att4 angvel3, or NULL), runs nticks with the schedule on `threads`
 * std::threads (contiguous chunks), writes final [n][ORC_NTRAJ] records, returns seconds spent in
 * the stepping loop only (construction excluded). */
double orc_run_population(const agf_vehicle_cfg* cfgs, uint32_t n_cfgs, uint32_t n,
                          const orc_opts* opts, const double* init13, const float* anchors /*[na][4]: id,x,y,z*/,
                          uint32_t n_anchors, uint32_t dt_us, uint32_t nticks,
                          const agf_cmd_entry* sched, uint32_t nsched, const uint8_t* slot_raw /*[slots][n][23]*/,
                          uint32_t threads, double* final_out);

/* the same with a trajectory: the record of every vehicle after every `stride` ticks ->
 * traj_out [nticks / stride][n][ORC_NTRAJ] (oracle/orc_population_traj.inc) */
double orc_run_population_traj(const agf_vehicle_cfg* cfgs, uint32_t n_cfgs, uint32_t n, const orc_opts* opts,
                               const double* init13, const float* anchors, uint32_t n_anchors, uint32_t dt_us,
                               uint32_t nticks, const agf_cmd_entry* sched, uint32_t nsched, const uint8_t* slot_raw,
                               uint32_t threads, uint32_t stride, double* traj_out);

/* codec cross-checks against the reference's own RadioTypes / TelemetryPacket code */
void orc_radio_encode_rates(uint8_t flags, float thrust, const float w[3], uint8_t raw[23]);
void orc_radio_encode_position(uint8_t flags, const float p[3], const float v[3], const float a[3],
                               uint8_t raw[23]);
void orc_radio_encode_acceleration(uint8_t flags, const float a[3], float yaw_rate, uint8_t raw[23]);
void orc_radio_decode(const uint8_t raw[23], uint8_t* type, uint8_t* flags, float floats[10]);
void orc_telemetry_decode(const uint8_t packet[30], agf_telemetry* out);
void orc_logic_consts(int quad_type, agf_logic_consts* out);
/* ref flavours only: vehicle ID -> the constructor arguments as the reference's apps derive them (main.cpp:147-165) */
int orc_vehicle_cfg_from_id(int vehicle_id, agf_vehicle_cfg* out);
void orc_radio_encode_idle(uint8_t flags, uint8_t raw[23]);

#ifdef __cplusplus
}
#endif
#endif
