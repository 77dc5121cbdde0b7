// agf_config.cpp -- vehicle configuration tables and the radio / telemetry codecs (host side).
//
// Restates, for the batched handle's construction path:
//   Components/Components/Logic/QuadcopterConstants.hpp:31-274   airframe tables
//   Components/Components/Logic/QuadcopterConstants.hpp:297-332  vehicle-ID -> type map
//   Components/Components/Logic/QuadcopterConstants.hpp:370-406  max motor speed from PWM constants
//   Simulator/Rappids_Simulator/main.cpp:147-218                 widening of the table into ctor args
//   Common/Common/DataTypes/RadioTypes.hpp:73-240                16-bit uplink codec
//   Common/Common/DataTypes/TelemetryPacket.hpp:39-207           16-bit downlink codec
// All table arithmetic is float, in the reference's order, with the platform powf/sqrtf
// (run once on the host, never on the device).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "agrifly_b200.h"

namespace {

// QuadcopterConstants.hpp:370-392
float max_cf_speed_from_pwm_consts(const float c[3][2]) {
  int MAX_PWM = 255;
  float MAX_BATT = 4.1;
  float k_1 = c[0][0] + c[0][1] * MAX_BATT;
  float k_2 = c[1][0] + c[1][1] * MAX_BATT;
  float k_3 = c[2][0] + c[2][1] * MAX_BATT;
  return (-k_2 + sqrtf(powf(k_2, 2) - 4 * k_3 * (k_1 - MAX_PWM))) / (2 * k_3);
}

// QuadcopterConstants.hpp:394-406
float max_esc_speed_from_pwm_consts(const float c[2]) {
  int ESC_PERIOD_MAX = 2000;
  return (ESC_PERIOD_MAX - c[0]) / c[1];
}

}  // namespace

extern "C" {

int agf_quad_type_from_id(unsigned id) {
  switch (id) {
    case 3: case 4: case 10:
      return AGF_QC_TYPE_CF_STANDARD;
    case 2: case 5: case 6: case 7: case 9: case 12: case 15: case 17:
      return AGF_QC_TYPE_CF_BIGMOTORSPROPS;
    case 13: case 14: case 18: case 19:
      return AGF_QC_TYPE_CF_LARGEQUAD;
    case 1: case 16: case 20: case 21: case 22: case 24: case 26:
      return AGF_QC_TYPE_CF_MINIQUAD;
    default:
      return AGF_QC_TYPE_INVALID;
  }
}

int agf_logic_consts_from_type(int quad_type, agf_logic_consts* o) {
  if (!o) return AGF_EINVAL;
  memset(o, 0, sizeof(*o));
  // defaults, QuadcopterConstants.hpp:33-51
  o->pos_control_nat_freq = 2.0f;
  o->pos_control_damping = 0.7f;
  o->ang_vel_control_time_const_xy = 0.03f;
  o->att_control_time_const_xy = 0.20f;
  o->ang_vel_control_time_const_z = 0.5f;
  o->att_control_time_const_z = 1.0f;
  o->motor_time_const = 0;
  o->motor_inertia = 0;
  o->motor_min_speed = 0;
  o->motor_max_speed = 10000;
  o->min_thrust_per_propeller = 0.0f;
  o->max_cmd_total_thrust = -1;
  const float perCellLowVoltage = 3.0f;
  o->low_battery_threshold = 0.0f;

  switch (quad_type) {
    case AGF_QC_TYPE_CF_STANDARD: {  // :54-90
      o->mass = 38e-3;
      o->inertia_xx = 16e-6f;
      o->inertia_zz = 29e-6f;
      o->arm_length = 46e-3f;
      o->prop_thrust_from_speed_sqr = float(3.58e-8f);
      o->prop_torque_from_thrust = 0.0006;
      o->prop0_spin_dir = 1;
      const float pwm[3][2] = {{-86.19993685f, 22.87189816f},
                               {0.30208677f, -0.07345602f},
                               {-1.59346434e-05f, 1.53209239e-05f}};
      o->motor_max_speed = max_cf_speed_from_pwm_consts(pwm);
      o->max_thrust_per_propeller = o->prop_thrust_from_speed_sqr * powf(o->motor_max_speed, 2);
      o->max_cmd_total_thrust = 0.9f * o->max_thrust_per_propeller * 4;
      o->ang_vel_control_time_const_xy = 0.04f;
      o->att_control_time_const_xy = 0.40f;
      o->low_battery_threshold = 1 * perCellLowVoltage;
      o->valid = 1;
      break;
    }
    case AGF_QC_TYPE_CF_BIGMOTORSPROPS: {  // :91-124
      o->mass = 39e-3;
      o->inertia_xx = 30e-6f;
      o->inertia_zz = 60e-6f;
      o->arm_length = 48e-3f;
      o->prop_thrust_from_speed_sqr = float(4.14e-8f);
      o->prop_torque_from_thrust = 0.001;
      o->prop0_spin_dir = 1;
      const float pwm[3][2] = {{-379.31113434f, 84.84738207f},
                               {0.65309704f, -0.13852527f},
                               {-1.34462353e-04f, 3.57662798e-05f}};
      o->motor_max_speed = max_cf_speed_from_pwm_consts(pwm);
      o->max_thrust_per_propeller = o->prop_thrust_from_speed_sqr * powf(o->motor_max_speed, 2);
      o->max_cmd_total_thrust = 0.8f * o->max_thrust_per_propeller * 4;
      o->low_battery_threshold = 1 * perCellLowVoltage;
      o->lin_drag_coeff_b[0] = 0.0206185f;
      o->lin_drag_coeff_b[1] = 0.0216621f;
      o->lin_drag_coeff_b[2] = 0.0f;
      o->valid = 1;
      break;
    }
    case AGF_QC_TYPE_CF_LARGEQUAD: {  // :157-195
      o->mass = 0.760;
      o->inertia_xx = 0.004406f;
      o->inertia_zz = 0.008611f;
      o->arm_length = 0.166f;
      o->prop_thrust_from_speed_sqr = 7.64e-6f;
      o->prop_torque_from_thrust = 0.0140f;
      o->prop0_spin_dir = 1;
      const float esc[2] = {972.0f, 0.742f};
      o->motor_max_speed = max_esc_speed_from_pwm_consts(esc);
      o->max_thrust_per_propeller = o->prop_thrust_from_speed_sqr * powf(o->motor_max_speed, 2);
      o->low_battery_threshold = 3 * perCellLowVoltage;
      o->ang_vel_control_time_const_xy = 0.0457f;
      o->att_control_time_const_xy = 0.0914f;
      o->ang_vel_control_time_const_z = 0.2545f;
      o->att_control_time_const_z = 0.5089f;
      o->lin_drag_coeff_b[0] = 0.1286181f;
      o->lin_drag_coeff_b[1] = 0.1286181f;
      o->lin_drag_coeff_b[2] = 0.1286181f;
      o->valid = 1;
      break;
    }
    case AGF_QC_TYPE_CF_MINIQUAD: {  // :196-235
      o->mass = 0.142;
      o->inertia_xx = 92.7e-6f;
      o->inertia_zz = 158.57e-6f;
      o->arm_length = 58e-3f;
      o->prop_thrust_from_speed_sqr = 4.32e-8f;
      o->prop_torque_from_thrust = 0.00808f;
      o->prop0_spin_dir = 1;
      const float esc[2] = {999.0f, 0.14f};
      o->motor_max_speed = max_esc_speed_from_pwm_consts(esc);
      o->max_thrust_per_propeller = o->prop_thrust_from_speed_sqr * powf(o->motor_max_speed, 2);
      o->min_thrust_per_propeller = 0.03f;
      o->max_cmd_total_thrust = 0.7f * (o->max_thrust_per_propeller * 4);
      o->low_battery_threshold = 2 * perCellLowVoltage;
      o->pos_control_nat_freq = 2.0f;
      o->pos_control_damping = 0.7f;
      o->ang_vel_control_time_const_xy = 0.04f;
      o->att_control_time_const_xy = o->ang_vel_control_time_const_xy * 2;
      o->ang_vel_control_time_const_z = o->ang_vel_control_time_const_xy * 5;
      o->att_control_time_const_z = o->ang_vel_control_time_const_z * 2;
      o->valid = 1;
      break;
    }
    case AGF_QC_TYPE_CF_FEEDTHROUGH:  // :125-156 (never allowed to fly)
    default:                          // :237-266
      o->valid = 0;
      o->mass = 1;
      o->inertia_xx = 1;
      o->inertia_zz = 1;
      o->arm_length = 1;
      o->prop_thrust_from_speed_sqr = 0;
      o->prop_torque_from_thrust = 0;
      o->max_thrust_per_propeller = 0;
      o->prop0_spin_dir = 0;
      if (quad_type == AGF_QC_TYPE_CF_FEEDTHROUGH) {
        o->low_battery_threshold = 1 * perCellLowVoltage;
      } else {
        o->motor_max_speed = 0;
      }
      break;
  }
  return AGF_OK;
}

int agf_vehicle_cfg_from_type(int quad_type, int vehicle_id, agf_vehicle_cfg* c) {
  if (!c) return AGF_EINVAL;
  memset(c, 0, sizeof(*c));
  int rc = agf_logic_consts_from_type(quad_type, &c->logic);
  if (rc) return rc;
  const agf_logic_consts& k = c->logic;
  // main.cpp:151-165: every value is the float table entry widened to double
  c->mass = k.mass;
  const double ixx = k.inertia_xx, izz = k.inertia_zz;
  c->inertia[0] = ixx;
  c->inertia[4] = ixx;  // inertia_yy = inertia_xx (main.cpp:154)
  c->inertia[8] = izz;
  c->arm_length = k.arm_length;
  c->prop_thrust_from_speed_sqr = k.prop_thrust_from_speed_sqr;
  // main.cpp:158-159: float * float product, then widened
  c->prop_torque_from_speed_sqr = k.prop_torque_from_thrust * k.prop_thrust_from_speed_sqr;
  c->motor_time_const = k.motor_time_const;
  c->motor_inertia = k.motor_inertia;
  c->motor_min_speed = k.motor_min_speed;
  c->motor_max_speed = k.motor_max_speed;
  for (int i = 0; i < 3; i++) c->lin_drag_coeff_b[i] = k.lin_drag_coeff_b[i];
  c->vehicle_id = vehicle_id;
  c->quad_type = quad_type;
  return AGF_OK;
}

// ---------------------------------------------------------------------------------------------
// radio uplink codec, RadioTypes.hpp
// ---------------------------------------------------------------------------------------------
enum {
  kIdxType = 0, kIdxReserved = 1, kIdxFlags = 2, kIdxFloats = 3,
  kEncSize = 2, kEncMax = 1 << 16, kEncHalf = kEncMax / 2,
  kMaxThrust = 35, kMaxAngRates = 35, kMaxPos = 20, kMaxVel = 10, kMaxAcc = 30, kMaxDefault = 1
};

// encodeToRadioByte :73-101
static void encode_field(float v, float limit, unsigned indx, uint8_t* bytes) {
  int out;
  if ((v > -limit) && (v < limit)) {
    out = int(v * kEncHalf / limit + 0.5f) + kEncHalf;
  } else if (v > -limit) {
    out = kEncMax - 1;
  } else if (v < limit) {
    out = 0;
  } else {
    out = 0;  // NaN
  }
  for (int i = 0; i < kEncSize; i++) {
    if (indx + i >= AGF_RADIO_PACKET_SIZE) break;
    bytes[indx + i] = uint8_t((out >> ((kEncSize - i - 1) * 8)) % 256);
  }
}

// decodeFromRadioBytes :103-116
static float decode_field(const uint8_t* bytes, unsigned indx, float limit) {
  int out = 0;
  for (int i = 0; i < kEncSize; i++) {
    if (indx + i >= AGF_RADIO_PACKET_SIZE) break;
    out += bytes[indx + i] << ((kEncSize - 1 - i) * 8);
  }
  return limit * (out - kEncHalf) / float(kEncHalf);
}

void agf_radio_encode_rates(uint8_t flags, float total_thrust, const float w[3], uint8_t* raw) {
  memset(raw, 0, AGF_RADIO_PACKET_SIZE);
  raw[kIdxType] = AGF_RADIO_EXTERNAL_RATES_CMD;
  raw[kIdxFlags] = flags;
  encode_field(total_thrust, kMaxThrust, kIdxFloats, raw);
  for (int i = 0; i < 3; i++) encode_field(w[i], kMaxAngRates, kIdxFloats + (i + 1) * kEncSize, raw);
}

void agf_radio_encode_position(uint8_t flags, const float p[3], const float v[3], const float a[3],
                               uint8_t* raw) {
  memset(raw, 0, AGF_RADIO_PACKET_SIZE);
  raw[kIdxType] = AGF_RADIO_POSITION_CMD;
  raw[kIdxFlags] = flags;
  for (int i = 0; i < 3; i++) {
    encode_field(p[i], kMaxPos, kIdxFloats + (0 + i) * kEncSize, raw);
    encode_field(v[i], kMaxVel, kIdxFloats + (3 + i) * kEncSize, raw);
    encode_field(a[i], kMaxAcc, kIdxFloats + (6 + i) * kEncSize, raw);
  }
}

void agf_radio_encode_acceleration(uint8_t flags, const float a[3], float yaw_rate, uint8_t* raw) {
  memset(raw, 0, AGF_RADIO_PACKET_SIZE);
  raw[kIdxType] = AGF_RADIO_EXTERNAL_ACCELERATION_CMD;
  raw[kIdxFlags] = flags;
  for (int i = 0; i < 3; i++) encode_field(a[i], kMaxAcc, kIdxFloats + i * kEncSize, raw);
  encode_field(yaw_rate, kMaxAngRates, kIdxFloats + 3 * kEncSize, raw);
}

void agf_radio_encode_idle(uint8_t flags, uint8_t* raw) {
  memset(raw, 0, AGF_RADIO_PACKET_SIZE);
  raw[kIdxType] = AGF_RADIO_IDLE_CMD;
  raw[kIdxFlags] = flags;
}

void agf_radio_encode_kill(uint8_t flags, uint8_t* raw) {
  memset(raw, 0, AGF_RADIO_PACKET_SIZE);
  raw[kIdxType] = AGF_RADIO_EMERGENCY_KILL;
  raw[kIdxFlags] = flags;
}

void agf_radio_decode(const uint8_t* raw, uint8_t* type, uint8_t* flags, float* f) {
  *type = raw[kIdxType];
  *flags = raw[kIdxFlags];
  switch (*type) {
    case AGF_RADIO_POSITION_CMD:
      for (int i = 0; i < 3; i++) f[i] = decode_field(raw, kIdxFloats + i * kEncSize, kMaxPos);
      for (int i = 3; i < 6; i++) f[i] = decode_field(raw, kIdxFloats + i * kEncSize, kMaxVel);
      for (int i = 6; i < 9; i++) f[i] = decode_field(raw, kIdxFloats + i * kEncSize, kMaxAcc);
      f[9] = 0;  // left uninitialised by the reference (RadioTypes.hpp:195-209); never read
      break;
    case AGF_RADIO_EXTERNAL_RATES_CMD:
      f[0] = decode_field(raw, kIdxFloats, kMaxThrust);
      for (int i = 1; i < AGF_RADIO_NUM_FLOATS; i++)
        f[i] = decode_field(raw, kIdxFloats + i * kEncSize, kMaxAngRates);
      break;
    case AGF_RADIO_EXTERNAL_ACCELERATION_CMD:
      for (int i = 0; i < 3; i++) f[i] = decode_field(raw, kIdxFloats + i * kEncSize, kMaxAcc);
      f[3] = decode_field(raw, kIdxFloats + 3 * kEncSize, kMaxAngRates);
      for (int i = 4; i < AGF_RADIO_NUM_FLOATS; i++) f[i] = 0;  // uninitialised in the reference
      break;
    default:
      for (int i = 0; i < AGF_RADIO_NUM_FLOATS; i++)
        f[i] = decode_field(raw, kIdxFloats + i * kEncSize, kMaxDefault);
      break;
  }
}

// ---------------------------------------------------------------------------------------------
// telemetry downlink codec, TelemetryPacket.hpp
// ---------------------------------------------------------------------------------------------
static float map_to_ab(float x, float a, float b) { return ((x + 1) / 2) * (b - a) + a; }  // :44-46
static float decode_ones_range(uint16_t t) {                                            // :66-71
  if (t == 0) return NAN;
  return (t - 32768) / float(32768);
}

void agf_telemetry_decode(const uint8_t* packet, agf_telemetry* o) {
  memset(o, 0, sizeof(*o));
  o->type = packet[0];
  o->packet_number = packet[1];
  uint16_t d[14];
  memcpy(d, packet + 2, 28);
  if (o->type == 0) {  // PACKET_TYPE_QUAD_TELEMETRY_PT1 :173-190
    for (int i = 0; i < 3; i++) {
      o->accel[i] = map_to_ab(decode_ones_range(d[i + 0]), -30, 30);
      o->gyro[i] = map_to_ab(decode_ones_range(d[i + 3]), -35, 35);
    }
    for (int i = 0; i < 4; i++) o->motor_forces[i] = map_to_ab(decode_ones_range(d[i + 6]), 0, 10);
    for (int i = 0; i < 3; i++) o->position[i] = map_to_ab(decode_ones_range(d[i + 10]), -30, 30);
    o->batt_voltage = map_to_ab(decode_ones_range(d[13]), 0, 15);
  } else if (o->type == 1) {  // PACKET_TYPE_QUAD_TELEMETRY_PT2 :192-205
    for (int i = 0; i < 3; i++) {
      o->velocity[i] = map_to_ab(decode_ones_range(d[i + 0]), -30, 30);
      o->attitude[i] = map_to_ab(decode_ones_range(d[i + 3]), -1, 1);
    }
    for (int i = 0; i < 6; i++) o->debug_vals[i] = map_to_ab(decode_ones_range(d[i + 6]), -100, 100);
    memcpy(&o->panic_reason, &d[12], 1);
    memcpy(&o->warnings, &d[13], 1);
  }
}

// ---- simulation.csv (Simulator/Rappids_Simulator/main.cpp:266-270,676-733) -------------------------------
size_t agf_csv_header(char* buf, size_t cap) {
  static const char h[] =
      "t,posx,posy,posz,velx,vely,velz,attY,attP,attR,angvelx,angvely,angvelz,m1,m2,m3,m4,"
      "estposx,estposy,estposz,estvelx,estvely,estvelz,esty,estp,estr,estangx,estangy,estangz,"
      "desposx,desposy,desposz,desvelx,desvely,desvelz,panic,r1,r2,r3,r4\n";
  if (buf && cap) {
    strncpy(buf, h, cap - 1);
    buf[cap - 1] = 0;
  }
  return sizeof(h) - 1;
}

namespace {
struct CsvOut {
  char* buf;
  size_t cap, len;
  void num(double v) {  // operator<<(double) of a default-constructed stream == "%g" (precision 6)
    char tmp[40];
    const int k = snprintf(tmp, sizeof(tmp), "%g,", v);
    for (int i = 0; i < k; i++, len++)
      if (buf && len + 1 < cap) buf[len] = tmp[i];
  }
  void integer(int v) {
    char tmp[24];
    const int k = snprintf(tmp, sizeof(tmp), "%d,", v);
    for (int i = 0; i < k; i++, len++)
      if (buf && len + 1 < cap) buf[len] = tmp[i];
  }
};
}  // namespace

size_t agf_csv_format_row(const agf_csv_record* r, char* buf, size_t cap) {
  CsvOut o{buf, cap, 0};
  o.num(r->t);
  for (int i = 0; i < 3; i++) o.num(r->pos[i]);
  for (int i = 0; i < 3; i++) o.num(r->vel[i]);
  {  // Rotationd::ToEulerYPR (Rotation.hpp:163-169)
    const double* v = r->att;
    const double y = atan2(2.0 * v[1] * v[2] + 2.0 * v[0] * v[3], v[1] * v[1] + v[0] * v[0] - v[3] * v[3] - v[2] * v[2]);
    const double p = -asin(2.0 * v[1] * v[3] - 2.0 * v[0] * v[2]);
    const double rr = atan2(2.0 * v[2] * v[3] + 2.0 * v[0] * v[1], v[3] * v[3] - v[2] * v[2] - v[1] * v[1] + v[0] * v[0]);
    o.num(y);
    o.num(p);
    o.num(rr);
  }
  for (int i = 0; i < 3; i++) o.num(r->ang_vel[i]);
  for (int i = 0; i < 4; i++) o.num(double(r->motor_forces[i]));
  for (int i = 0; i < 3; i++) o.num(double(r->est_pos[i]));
  for (int i = 0; i < 3; i++) o.num(double(r->est_vel[i]));
  {  // Rotationf::ToEulerYPR
    const float* v = r->est_att;
    const float y = atan2f(2.0f * v[1] * v[2] + 2.0f * v[0] * v[3], v[1] * v[1] + v[0] * v[0] - v[3] * v[3] - v[2] * v[2]);
    const float p = -asinf(2.0f * v[1] * v[3] - 2.0f * v[0] * v[2]);
    const float rr = atan2f(2.0f * v[2] * v[3] + 2.0f * v[0] * v[1], v[3] * v[3] - v[2] * v[2] - v[1] * v[1] + v[0] * v[0]);
    o.num(double(y));
    o.num(double(p));
    o.num(double(rr));
  }
  for (int i = 0; i < 3; i++) o.num(double(r->est_ang_vel[i]));
  for (int i = 0; i < 3; i++) o.num(r->des_pos[i]);
  for (int i = 0; i < 3; i++) o.num(r->des_vel[i]);
  o.integer(r->panic_reason);
  for (int i = 0; i < 4; i++) o.num(double(r->last_radio_cmd[i]));
  if (buf && o.len + 1 < cap) buf[o.len] = '\n';
  o.len++;
  if (buf && cap) buf[o.len < cap ? o.len : cap - 1] = 0;
  return o.len;
}

// ---- ROS message fields (AIFS_ROS/hiperlab_rostools/src/Simulator/main.cpp:455-475, 501-546) -----------------------
void agf_msg_simulator_truth_fill(int64_t vehicle_id, const double pos[3], const double vel[3], const double att[4],
                                  const double ang_vel[3], agf_msg_simulator_truth* o) {
  memset(o, 0, sizeof(*o));
  o->vehicleID = vehicle_id;
  o->posx = pos[0]; o->posy = pos[1]; o->posz = pos[2];
  o->velx = vel[0]; o->vely = vel[1]; o->velz = vel[2];
  o->attq0 = att[0]; o->attq1 = att[1]; o->attq2 = att[2]; o->attq3 = att[3];
  const double* v = att;  // Rotationd::ToEulerYPR (Rotation.hpp:163-169)
  o->attyaw = atan2(2.0 * v[1] * v[2] + 2.0 * v[0] * v[3], v[1] * v[1] + v[0] * v[0] - v[3] * v[3] - v[2] * v[2]);
  o->attpitch = -asin(2.0 * v[1] * v[3] - 2.0 * v[0] * v[2]);
  o->attroll = atan2(2.0 * v[2] * v[3] + 2.0 * v[0] * v[1], v[3] * v[3] - v[2] * v[2] - v[1] * v[1] + v[0] * v[0]);
  o->angvelx = ang_vel[0]; o->angvely = ang_vel[1]; o->angvelz = ang_vel[2];
}

void agf_msg_telemetry_fill(const uint8_t* packet1, const uint8_t* packet2, agf_msg_telemetry* o) {
  memset(o, 0, sizeof(*o));
  agf_telemetry p1, p2;
  agf_telemetry_decode(packet1, &p1);
  agf_telemetry_decode(packet2, &p2);
  o->packetNumber = p1.packet_number;
  for (int i = 0; i < 3; i++) {
    o->accelerometer[i] = p1.accel[i];
    o->rateGyro[i] = p1.gyro[i];
    o->position[i] = p1.position[i];
  }
  for (int i = 0; i < 4; i++) o->motorForces[i] = p1.motor_forces[i];
  o->batteryVoltage = p1.batt_voltage;
  for (int i = 0; i < 6; i++) o->debugVals[i] = p2.debug_vals[i];
  // Rotationf::FromVectorPartOfQuaternion (Rotation.hpp:112-121; the unit vector's norm is a float, Vec3.hpp:126-129)
  float vx = p2.attitude[0], vy = p2.attitude[1], vz = p2.attitude[2];
  float tmp = vx * vx + vy * vy + vz * vz;
  if (tmp > 1.0f) {
    const float nrm = sqrtf(tmp);
    vx = vx / nrm; vy = vy / nrm; vz = vz / nrm;
    tmp = 1.0f;
  }
  const float q[4] = {float(sqrt(1 - tmp)), vx, vy, vz};
  const float y = atan2f(2.0f * q[1] * q[2] + 2.0f * q[0] * q[3], q[1] * q[1] + q[0] * q[0] - q[3] * q[3] - q[2] * q[2]);
  const float pch = -asinf(2.0f * q[1] * q[3] - 2.0f * q[0] * q[2]);
  const float r = atan2f(2.0f * q[2] * q[3] + 2.0f * q[0] * q[1], q[3] * q[3] - q[2] * q[2] - q[1] * q[1] + q[0] * q[0]);
  const float ypr[3] = {y, pch, r};
  for (int i = 0; i < 3; i++) {
    o->velocity[i] = p2.velocity[i];
    o->attitude[i] = p2.attitude[i];
    o->attitudeYPR[i] = ypr[i];
  }
  o->panicReason = p2.panic_reason;
  o->warnings = p2.warnings;
}

}  // extern "C"
