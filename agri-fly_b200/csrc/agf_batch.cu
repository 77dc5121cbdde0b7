// agf_batch.cu -- the batched handle behind include/agrifly_b200.h.
//
// Host side of the drop-in boundary: construction (Quadcopter_T::Quadcopter_T,
// Quadcopter_T.cpp:9-83, and QuadcopterLogic::Initialise, QuadcopterLogic.cpp:97-162, evaluated
// once on the host for the whole batch), state get/set (SimulationObject6DOF.hpp:26-56), radio
// command delivery, telemetry read-out, and the launch of the step kernels.  Everything that
// touches vehicle state runs on the GPU; there is no CPU execution path for the step.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "agf_host_params.h"
#include "agf_launch.h"
#include "agf_math.h"
#include "agf_nccl.h"
#include "agf_types.h"
#include "agrifly_b200.h"

namespace agf {

static thread_local std::string g_last_error;
static int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
  char buf[512];
  if (e != cudaSuccess) {
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  } else {
    snprintf(buf, sizeof(buf), "%s", what);
  }
  g_last_error = buf;
  return code;
}
// error reporting for the other translation units of the library (agf_rappids.cu)
int fail_from(int code, const char* what, int cuda_error) { return fail(code, what, (cudaError_t)cuda_error); }
#define AGF_CUDA(call)                                      \
  do {                                                      \
    cudaError_t e_ = (call);                                \
    if (e_ != cudaSuccess) return fail(AGF_ECUDA, #call, e_); \
  } while (0)

// Batches that hold a captured read-out graph (stats kernel -> ncclAllGather -> combine).  NCCL keeps a communicator alive
// while a graph that captured one of its collectives exists -- ncclCommDestroy would wait for ever -- so
// agf_nccl_comm_destroy first drops the graphs captured with that communicator (release_stats_graphs_for).
struct Batch;
static std::mutex g_graph_mutex;
static std::vector<Batch*> g_graph_owners;
void release_stats_graphs_for(void* comm);

// ---------------------------------------------------------------------------------------------
// small kernels around the step: field gather/scatter, immediate radio delivery, telemetry, stats
// ---------------------------------------------------------------------------------------------
template<typename P>
struct FieldCtx {
  P* sp; float* sf; uint32_t* su; float* sc;
  size_t n;
  P kF_shared;
  const P* pv;  // per-vehicle plant scalars or null
};

template<typename P>
__device__ inline double f_get(const FieldCtx<P>& c, int field, size_t i, int comp, int* as_int, float* as_float, int* kind) {
  constexpr int VP = VecOf<P>::lanes;
  *kind = 0;  // 0 double, 1 float, 2 int
  switch (field) {
    case AGF_F_POSITION: return double(c.sp[sidx(SP_POS + comp, c.n, i, VP)]);
    case AGF_F_VELOCITY: return double(c.sp[sidx(SP_VEL + comp, c.n, i, VP)]);
    case AGF_F_ATTITUDE: return double(c.sp[sidx(SP_ATT + comp, c.n, i, VP)]);
    case AGF_F_ANGULAR_VELOCITY: return double(c.sp[sidx(SP_W + comp, c.n, i, VP)]);
    case AGF_F_MOTOR_SPEED: return double(c.sp[sidx(SP_MS + comp, c.n, i, VP)]);
    case AGF_F_MOTOR_FORCE: {
      const P sp = c.sp[sidx(SP_MS + comp, c.n, i, VP)];
      const P kF = c.pv ? c.pv[sidx(PV_KF, c.n, i, VP)] : c.kF_shared;
      return double((kF * sp) * (sp < 0 ? -sp : sp));
    }
    case AGF_F_MOTOR_SPEED_CMD: *kind = 1; *as_float = c.sf[sidx(SF_CMD + comp, c.n, i, 4)]; return 0;
    case AGF_F_DES_MOTOR_FORCE: *kind = 1; *as_float = c.sf[sidx(SF_DFORCE + comp, c.n, i, 4)]; return 0;
    case AGF_F_EST_POSITION: *kind = 1; *as_float = c.sf[sidx(SF_KPOS + comp, c.n, i, 4)]; return 0;
    case AGF_F_EST_VELOCITY: *kind = 1; *as_float = c.sf[sidx(SF_KVEL + comp, c.n, i, 4)]; return 0;
    case AGF_F_EST_ATTITUDE: *kind = 1; *as_float = c.sf[sidx(SF_KATT + comp, c.n, i, 4)]; return 0;
    case AGF_F_EST_ANGULAR_VELOCITY: *kind = 1; *as_float = c.sf[sidx(SF_KW + comp, c.n, i, 4)]; return 0;
    case AGF_F_ACCELEROMETER: *kind = 1; *as_float = c.sf[sidx(SF_ACC_LP + 4 * comp + 3, c.n, i, 4)]; return 0;
    case AGF_F_RATE_GYRO: *kind = 1; *as_float = c.sf[sidx(SF_GYRO_LP + 4 * comp + 3, c.n, i, 4)]; return 0;
    case AGF_F_EST_COVARIANCE: *kind = 1; *as_float = c.sc ? c.sc[sidx(comp, c.n, i, 4)] : 0.0f; return 0;
    case AGF_F_UWB_MEASUREMENT:
      *kind = 1;
      *as_float = comp == 0 ? c.sf[sidx(SF_UWB_RANGE, c.n, i, 4)] : float((c.su[sidx(SU_UWBW, c.n, i, 4)] >> 8) & 0xFFu);
      return 0;
    case AGF_F_FLIGHT_STATE: *kind = 2; *as_int = int(c.su[sidx(SU_BITS, c.n, i, 4)] & 0x7u); return 0;
    case AGF_F_PANIC_REASON: *kind = 2; *as_int = int((c.su[sidx(SU_BITS, c.n, i, 4)] >> 3) & 0x7u); return 0;
    case AGF_F_CYCLE_COUNTER: *kind = 2; *as_int = int(c.su[sidx(SU_CYCLE, c.n, i, 4)]); return 0;
    case AGF_F_KF_COUNTERS: {
      *kind = 2;
      const uint32_t kc = c.su[sidx(SU_KFCNT, c.n, i, 4)], bits = c.su[sidx(SU_BITS, c.n, i, 4)];
      if (comp == 0) *as_int = int(kc & 0xFFFFu);
      else if (comp == 1) *as_int = int(kc >> 16);
      else if (comp == 2) *as_int = int(c.su[sidx(SU_UWB_COUNT, c.n, i, 4)]);
      else *as_int = int(((bits >> 18) & 1u) | (((bits >> 19) & 1u) << 1));
      return 0;
    }
  }
  return 0;
}

template<typename P>
__global__ void field_get_kernel(FieldCtx<P> c, int field, int ncomp, size_t first, size_t count, void* out) {
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= count * ncomp) return;
  const size_t v = t / ncomp;
  const int comp = int(t % ncomp);
  int ai = 0, kind = 0;
  float af = 0;
  const double d = f_get(c, field, first + v, comp, &ai, &af, &kind);
  if (kind == 0) ((double*)out)[t] = d;
  else if (kind == 1) ((float*)out)[t] = af;
  else ((int32_t*)out)[t] = ai;
}

template<typename P>
__global__ void field_set_kernel(FieldCtx<P> c, int field, int ncomp, size_t first, size_t count, const void* in) {
  constexpr int VP = VecOf<P>::lanes;
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= count * ncomp) return;
  const size_t i = first + t / ncomp;
  const int comp = int(t % ncomp);
  // the staged buffer holds doubles for the plant fields and floats for the logic fields (field_elem): read it with
  // the element type of the field only -- a double read of a float-sized stage would run past its end
  const double* ind = (const double*)in;
  const float* inf = (const float*)in;
  switch (field) {
    // a position / velocity written from outside starts its compensated sum afresh (FP32 fast variants, agf_step.cuh tick())
    case AGF_F_POSITION: c.sp[sidx(SP_POS + comp, c.n, i, VP)] = P(ind[t]); c.sp[sidx(SP_CPOS + comp, c.n, i, VP)] = P(0); break;
    case AGF_F_VELOCITY: c.sp[sidx(SP_VEL + comp, c.n, i, VP)] = P(ind[t]); c.sp[sidx(SP_CVEL + comp, c.n, i, VP)] = P(0); break;
    case AGF_F_ATTITUDE: c.sp[sidx(SP_ATT + comp, c.n, i, VP)] = P(ind[t]); break;
    case AGF_F_ANGULAR_VELOCITY: c.sp[sidx(SP_W + comp, c.n, i, VP)] = P(ind[t]); break;
    case AGF_F_MOTOR_SPEED: c.sp[sidx(SP_MS + comp, c.n, i, VP)] = P(ind[t]); break;
    case AGF_F_MOTOR_SPEED_CMD: c.sf[sidx(SF_CMD + comp, c.n, i, 4)] = inf[t]; break;
    case AGF_F_EST_POSITION: c.sf[sidx(SF_KPOS + comp, c.n, i, 4)] = inf[t]; break;
    case AGF_F_EST_VELOCITY: c.sf[sidx(SF_KVEL + comp, c.n, i, 4)] = inf[t]; break;
    case AGF_F_EST_ATTITUDE: c.sf[sidx(SF_KATT + comp, c.n, i, 4)] = inf[t]; break;
    case AGF_F_EST_ANGULAR_VELOCITY: c.sf[sidx(SF_KW + comp, c.n, i, 4)] = inf[t]; break;
    default: break;
  }
}

// the whole 6-DOF state of a vehicle range in one pass: in = [count][13] doubles, position 3, velocity 3, attitude 4, angular velocity 3
template<typename P>
__global__ void state13_set_kernel(FieldCtx<P> c, size_t first, size_t count, const double* __restrict__ in) {
  constexpr int VP = VecOf<P>::lanes;
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= count * 13) return;
  const size_t i = first + t / 13;
  const int comp = int(t % 13);
  const int slot = comp < 3 ? SP_POS + comp : (comp < 6 ? SP_VEL + comp - 3 : (comp < 10 ? SP_ATT + comp - 6 : SP_W + comp - 10));
  c.sp[sidx(slot, c.n, i, VP)] = P(in[t]);
  if (comp < 6) c.sp[sidx(comp < 3 ? SP_CPOS + comp : SP_CVEL + comp - 3, c.n, i, VP)] = P(0);
}

// ... and the way back: out = [count][13] doubles
template<typename P>
__global__ void state13_get_kernel(FieldCtx<P> c, size_t first, size_t count, double* __restrict__ out) {
  constexpr int VP = VecOf<P>::lanes;
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= count * 13) return;
  const size_t i = first + t / 13;
  const int comp = int(t % 13);
  const int slot = comp < 3 ? SP_POS + comp : (comp < 6 ? SP_VEL + comp - 3 : (comp < 10 ? SP_ATT + comp - 6 : SP_W + comp - 10));
  out[t] = double(c.sp[sidx(slot, c.n, i, VP)]);
}

// SetExternalForce / SetExternalTorque of a vehicle range: in = [count][3] doubles (either may be null), dst = [3][n]
template<typename P>
__global__ void wrench_set_kernel(P* __restrict__ dst_f, P* __restrict__ dst_t, size_t n, size_t first, size_t count,
                                  const double* __restrict__ in_f, const double* __restrict__ in_t) {
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= count * 3) return;
  const size_t i = first + t / 3;
  const int k = int(t % 3);
  if (in_f) dst_f[size_t(k) * n + i] = P(in_f[t]);
  if (in_t) dst_t[size_t(k) * n + i] = P(in_t[t]);
}

// SetCommandRadioMsg now (QuadcopterLogic.hpp:110-116), outside the step kernel
struct RadioNow {
  uint32_t type, flags;
  float f[4];
};
__global__ void radio_now_kernel(float* sf, uint32_t* su, size_t n, size_t first, size_t count, int broadcast,
                                 RadioNow one, const RadioNow* per, float mon_cmd_coef, int hk) {
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const size_t i = first + t;
  const RadioNow m = broadcast ? one : per[t];
  uint32_t bits = su[sidx(SU_BITS, n, i, 4)];
  bits |= (1u << 17);
  bits = (bits & ~(0x7u << 6)) | ((m.type & 0x7u) << 6);
  bits = (bits & ~(0xFFu << 9)) | ((m.flags & 0xFFu) << 9);
  su[sidx(SU_BITS, n, i, 4)] = bits;
  for (int k = 0; k < 4; k++) sf[sidx(SF_RADIO + k, n, i, 4)] = m.f[k];
  su[sidx(SU_AGE_RADIO, n, i, 4)] = 0;
  if (hk) {
    const uint32_t age = su[sidx(SU_AGE_MON_CMD, n, i, 4)];
    const float mdt = float(age) * 1e-6f;
    const float prev = sf[sidx(SF_MON_CMD, n, i, 4)];
    sf[sidx(SF_MON_CMD, n, i, 4)] = mon_cmd_coef <= 0.0f ? mdt : __fadd_rn(__fmul_rn(mon_cmd_coef, prev), __fmul_rn(__fsub_rn(1.0f, mon_cmd_coef), mdt));
    su[sidx(SU_AGE_MON_CMD, n, i, 4)] = age - uint32_t(__fmul_rn(mdt, 1e6f));
  }
}

// Clock advance without a Run(): the per-vehicle stopwatches (radio / UWB timeouts, monitors) age by dt
__global__ void advance_clock_kernel(uint4* su, size_t n, uint32_t dt_us) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  auto sat = [dt_us](uint32_t a) { return a > 0xF0000000u ? a : a + dt_us; };
  uint4 a = su[size_t(SU_AGE_RADIO / 4) * n + i];  // {uwb_count, age_radio, age_uwb, uwbw}
  a.y = sat(a.y);
  a.z = sat(a.z);
  su[size_t(SU_AGE_RADIO / 4) * n + i] = a;
  uint4 h = su[size_t(SU_AGE_EST_RESET / 4) * n + i];  // {age_est_reset, pc_count, age_mon_cmd, age_mon_loop}
  h.x = sat(h.x);
  h.z = sat(h.z);
  h.w = sat(h.w);
  su[size_t(SU_AGE_EST_RESET / 4) * n + i] = h;
}

// TelemetryPacket.hpp:39-63: MapToOnesRange + EncodeOnesRange (round-to-nearest ops, never contracted)
__device__ inline uint16_t tel_encode(float x, float a, float b) {
  const float t = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(x, a), __fsub_rn(b, a)), 2.0f), 1.0f);
  if (t < -1 || t > 1) return 0;
  const float e = __fadd_rn(32768.0f, __fmul_rn(32767.0f, t));
  if (!(e == e)) return 0;
  return uint16_t(int(e));
}

// GetTelemetryDataPackets (QuadcopterLogic.cpp:621-679)
__global__ void telemetry_kernel(const float* sf, uint32_t* su, uint32_t* tel_counter, size_t n, size_t first,
                                 size_t count, float batt_voltage, int hk, uint8_t* p1, uint8_t* p2) {
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const size_t i = first + t;
  uint16_t d[14];
  uint8_t* o1 = p1 + t * AGF_TELEMETRY_PACKET_SIZE;
  uint8_t* o2 = p2 + t * AGF_TELEMETRY_PACKET_SIZE;
  const uint32_t ctr = tel_counter[i];
  for (int k = 0; k < 3; k++) {
    d[k] = tel_encode(sf[sidx(SF_ACC_LP + 4 * k + 3, n, i, 4)], -30, 30);
    d[3 + k] = tel_encode(sf[sidx(SF_GYRO_LP + 4 * k + 3, n, i, 4)], -35, 35);
    d[10 + k] = tel_encode(sf[sidx(SF_KPOS + k, n, i, 4)], -30, 30);
  }
  for (int k = 0; k < 4; k++) d[6 + k] = tel_encode(sf[sidx(SF_DFORCE + k, n, i, 4)], 0, 10);
  d[13] = tel_encode(batt_voltage, 0, 15);
  o1[0] = 0;
  o1[1] = uint8_t(ctr % 256);
  for (int k = 0; k < 14; k++) { o1[2 + 2 * k] = uint8_t(d[k] & 0xFF); o1[3 + 2 * k] = uint8_t(d[k] >> 8); }
  const float qw = sf[sidx(SF_KATT + 0, n, i, 4)];
  for (int k = 0; k < 3; k++) {
    d[k] = tel_encode(sf[sidx(SF_KVEL + k, n, i, 4)], -30, 30);
    const float a = sf[sidx(SF_KATT + 1 + k, n, i, 4)];
    d[3 + k] = tel_encode(qw > 0 ? a : -a, -1, 1);  // ToVectorPartOfQuaternion (Rotation.hpp:155-161)
  }
  // debug[0] = temperature low-pass output (QuadcopterLogic.cpp:178); others stay 0
  const float dbg0 = hk ? sf[sidx(SF_TEMP_LP + 3, n, i, 4)] : 25.0f;
  const uint32_t cyc = su[sidx(SU_CYCLE, n, i, 4)];
  d[6] = tel_encode(cyc ? dbg0 : 0.0f, -100, 100);
  for (int k = 1; k < 6; k++) d[6 + k] = tel_encode(0.0f, -100, 100);
  const uint32_t bits = su[sidx(SU_BITS, n, i, 4)];
  const uint32_t cnt = su[sidx(SU_CNT, n, i, 4)];
  d[12] = uint16_t((bits >> 3) & 0x7u);
  d[13] = uint16_t((cnt >> 24) & 0xFFu);
  o2[0] = 1;
  o2[1] = uint8_t(ctr % 256);
  for (int k = 0; k < 14; k++) { o2[2 + 2 * k] = uint8_t(d[k] & 0xFF); o2[3 + 2 * k] = uint8_t(d[k] >> 8); }
  tel_counter[i] = ctr + 1;
  su[sidx(SU_CNT, n, i, 4)] = cnt & 0x00FFFFFFu;  // warnings cleared after sending
}

// K4: Monte-Carlo statistics. warp shuffle tree -> one atomic per warp leader per quantity.
__device__ inline double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ inline double warp_max(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ inline void atomic_max_double(double* addr, double v) {  // v >= 0
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double((long long)assumed) >= v) break;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
  } while (assumed != old);
}

// the all-gathered statistics vectors of every rank -> one vector, in rank order (deterministic, the same bits on every rank)
// one record of the trajectory log ([quad][log_n] vectors + [log_n] scalars, agf_types.h StepLaunch::log) -> double[count][17]
template<typename P>
__global__ void log_gather_kernel(const P* __restrict__ rec, size_t log_n, size_t first, size_t count, double* __restrict__ out) {
  constexpr int VP = VecOf<P>::lanes;
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= count * AGF_LOG_FIELDS) return;
  const size_t i = first + t / AGF_LOG_FIELDS;
  const int f = int(t % AGF_LOG_FIELDS);
  out[t] = double(f < AGF_LOG_FIELDS - 1 ? rec[(size_t(f / VP) * log_n + i) * VP + f % VP] : rec[size_t(AGF_LOG_FIELDS - 1) * log_n + i]);
}

__global__ void stats_combine_kernel(const double* gathered, int nranks, double* out) {
  const int k = threadIdx.x;
  if (k >= AGF_STATS_LEN) return;
  double r = gathered[k];
  for (int q = 1; q < nranks; q++) {
    const double v = gathered[q * AGF_STATS_LEN + k];
    r = k >= AGF_ST_MAX_ENORM ? fmax(r, v) : r + v;
  }
  out[k] = r;
}

template<typename P>
__global__ void stats_kernel(const P* sp, const float* sf, const uint32_t* su, size_t n, const double* target, double* out) {
  constexpr int VP = VecOf<P>::lanes;
  __shared__ double sh[AGF_STATS_LEN][8];
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  double v[AGF_STATS_LEN];
  for (int k = 0; k < AGF_STATS_LEN; k++) v[k] = 0;
  if (i < n) {
    double p[3], e[3], ee[3];
    for (int k = 0; k < 3; k++) {
      p[k] = double(sp[sidx(SP_POS + k, n, i, VP)]);
      const double tgt = target ? target[3 * i + k] : double(sf[sidx(SF_RADIO + k, n, i, 4)]);
      e[k] = p[k] - tgt;
      ee[k] = double(sf[sidx(SF_KPOS + k, n, i, 4)]) - p[k];
    }
    const double vx = double(sp[sidx(SP_VEL + 0, n, i, VP)]), vy = double(sp[sidx(SP_VEL + 1, n, i, VP)]), vz = double(sp[sidx(SP_VEL + 2, n, i, VP)]);
    const bool finite = isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]);
    const uint32_t fs = su[sidx(SU_BITS, n, i, 4)] & 0x7u;
    v[AGF_ST_COUNT] = 1;
    v[AGF_ST_N_PANIC] = fs == AGF_FS_PANIC;
    v[AGF_ST_N_KILLED] = fs == AGF_FS_KILLED;
    v[AGF_ST_N_AUTONOMOUS] = fs == AGF_FS_FULLY_AUTONOMOUS;
    v[AGF_ST_N_NONFINITE] = !finite;
    if (finite) {
      const double e2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
      const double en = sqrt(e2), est = sqrt(ee[0] * ee[0] + ee[1] * ee[1] + ee[2] * ee[2]);
      v[AGF_ST_SUM_EX] = e[0]; v[AGF_ST_SUM_EY] = e[1]; v[AGF_ST_SUM_EZ] = e[2];
      v[AGF_ST_SUM_E2] = e2; v[AGF_ST_SUM_ENORM] = en;
      v[AGF_ST_SUM_SPEED] = sqrt(vx * vx + vy * vy + vz * vz);
      v[AGF_ST_SUM_EST_ERR] = isfinite(est) ? est : 0.0;
      v[AGF_ST_MAX_ENORM] = en;
      v[AGF_ST_MAX_EST_ERR] = isfinite(est) ? est : 0.0;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  for (int k = 0; k < AGF_STATS_LEN; k++) {
    const double r = k >= AGF_ST_MAX_ENORM ? warp_max(v[k]) : warp_sum(v[k]);
    if (lane == 0) sh[k][warp] = r;
  }
  __syncthreads();
  if (threadIdx.x < AGF_STATS_LEN) {
    const int k = threadIdx.x;
    double r = sh[k][0];
    for (int w = 1; w < nwarp; w++) r = k >= AGF_ST_MAX_ENORM ? fmax(r, sh[k][w]) : r + sh[k][w];
    if (k >= AGF_ST_MAX_ENORM) atomic_max_double(&out[k], r);
    else atomicAdd(&out[k], r);
  }
}

// ---------------------------------------------------------------------------------------------
// the handle
// ---------------------------------------------------------------------------------------------
struct Batch {
  virtual ~Batch() {}
  virtual int run(uint32_t dt_us, uint32_t nticks) = 0;
  virtual int advance_clock(uint32_t dt_us) = 0;
  virtual int get_field(int field, void* dst, size_t first, size_t count) = 0;
  virtual int set_field(int field, const void* src, size_t first, size_t count) = 0;
  virtual int set_state13(const double* src, size_t first, size_t count) = 0;
  virtual int get_state13(double* dst, size_t first, size_t count) = 0;
  virtual int set_radio(const uint8_t* raw, size_t first, size_t count, int broadcast) = 0;
  virtual int set_schedule(const agf_cmd_entry* e, size_t n) = 0;
  virtual int set_slot(int slot, const uint8_t* raw) = 0;
  virtual int telemetry(uint8_t* p1, uint8_t* p2, size_t first, size_t count) = 0;
  virtual int set_wrench(const double* f, const double* t, size_t first, size_t count) = 0;
  virtual int add_anchor(uint8_t id, const float pos[3]) = 0;
  virtual int enable_log(uint32_t stride, uint32_t cap) = 0;
  virtual int read_log(uint64_t rec, double* dst, size_t first, size_t count) = 0;
  virtual int log_ptr(void** p, size_t* es, size_t* ln) = 0;
  virtual int stats(const double* target, double* dev_out) = 0;
  virtual int set_offboard(const agf_offboard_cfg* cfg, const agf_offboard_target* targets, size_t n_targets, const double* offsets) = 0;
  virtual int set_offboard_ref(const agf_offboard_ref* ref) = 0;
  virtual int set_offboard_traj(const double* traj, size_t first, size_t count) = 0;
  virtual int get_offboard_state(double* out, size_t first, size_t count) = 0;
  virtual int offboard_traj_ptr(double** p, size_t* nv) = 0;
  virtual int set_offboard_estimator(const agf_offboard_estimator* e) = 0;
  virtual int get_offboard_estimate(double horizon, double* est13, double* counters4, size_t first, size_t count) = 0;

  size_t n = 0;
  agf_batch_opts opts;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  uint64_t now_us = 0, ticks = 0, launches = 0;
  uint64_t log_records = 0;
  // Step-kernel timing (agf_batch_step_kernel_time): a fixed ring of event pairs.  When the ring is full the oldest
  // pair -- EVENT_RING launches old, long complete -- is folded into the running totals, so an object-API caller that
  // steps one tick per launch for hours never grows the handle (round-1 advice: two events leaked per launch).
  enum { EVENT_RING = 64 };
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
  size_t ev_head = 0, ev_count = 0;  // ring of pending pairs: [ev_head, ev_head + ev_count)
  double timed_ms = 0;
  uint64_t timed_launches = 0;

  int fold_oldest_event() {
    auto& e = events[ev_head];
    float t = 0;
    AGF_CUDA(cudaEventSynchronize(e.second));
    AGF_CUDA(cudaEventElapsedTime(&t, e.first, e.second));
    timed_ms += t;
    timed_launches++;
    ev_head = (ev_head + 1) % EVENT_RING;
    ev_count--;
    return AGF_OK;
  }
  // the pair that brackets the next launch
  int next_event_pair(std::pair<cudaEvent_t, cudaEvent_t>** out) {
    if (events.empty()) {  // the whole ring at first use: the ring indices below run modulo EVENT_RING
      events.reserve(EVENT_RING);
      for (int k = 0; k < EVENT_RING; k++) {
        cudaEvent_t a, b;
        AGF_CUDA(cudaEventCreate(&a));
        AGF_CUDA(cudaEventCreate(&b));
        events.push_back({a, b});
      }
    }
    if (ev_count == size_t(EVENT_RING)) {
      int rc = fold_oldest_event();
      if (rc) return rc;
    }
    *out = &events[(ev_head + ev_count) % EVENT_RING];
    ev_count++;
    return AGF_OK;
  }
  int drain_events(double* ms, uint64_t* n_launches) {
    while (ev_count) {
      int rc = fold_oldest_event();
      if (rc) return rc;
    }
    if (ms) *ms = timed_ms;
    if (n_launches) *n_launches = timed_launches;
    timed_ms = 0;
    timed_launches = 0;
    return AGF_OK;
  }

  int sync() {
    AGF_CUDA(cudaStreamSynchronize(stream));
    return AGF_OK;
  }

  // ---- statistics read-out buffers, allocated once per handle (no cudaMalloc on the read-out path) ----
  double* d_stats_local = nullptr;   // this GPU's AGF_STATS_LEN doubles
  double* d_stats_gather = nullptr;  // [nranks][AGF_STATS_LEN]
  double* d_stats_out = nullptr;     // combined vector when the caller wants it on the host
  double* h_stats = nullptr;         // pinned
  int gather_ranks = 0;
  // the captured read-out (stats kernel -> all-gather -> combine) and what it was captured for
  cudaGraphExec_t stats_graph = nullptr;
  void* graph_comm = nullptr;
  double* graph_out = nullptr;
  uint64_t nccl_readouts = 0;
  bool graph_disabled = false;

  int ensure_stats_buffers(int nranks) {
    if (!d_stats_local) AGF_CUDA(cudaMalloc(&d_stats_local, sizeof(double) * AGF_STATS_LEN));
    if (!d_stats_out) AGF_CUDA(cudaMalloc(&d_stats_out, sizeof(double) * AGF_STATS_LEN));
    if (!h_stats) AGF_CUDA(cudaMallocHost(&h_stats, sizeof(double) * AGF_STATS_LEN));
    if (nranks > gather_ranks) {
      cudaFree(d_stats_gather);
      d_stats_gather = nullptr;
      gather_ranks = 0;
      AGF_CUDA(cudaMalloc(&d_stats_gather, sizeof(double) * AGF_STATS_LEN * size_t(nranks)));
      gather_ranks = nranks;
    }
    return AGF_OK;
  }
  void drop_stats_graph() {  // under g_graph_mutex
    if (stats_graph) cudaGraphExecDestroy(stats_graph);
    stats_graph = nullptr;
    graph_comm = nullptr;
    graph_out = nullptr;
  }
  void free_stats_buffers() {
    {
      std::lock_guard<std::mutex> lock(g_graph_mutex);
      drop_stats_graph();
      g_graph_owners.erase(std::remove(g_graph_owners.begin(), g_graph_owners.end(), this), g_graph_owners.end());
    }
    cudaFree(d_stats_local);
    cudaFree(d_stats_gather);
    cudaFree(d_stats_out);
    if (h_stats) cudaFreeHost(h_stats);
  }

  // stats kernel + ONE all-gather + combine on this handle's stream (include/agrifly_b200.h "multi-GPU")
  int stats_nccl(void* comm, const double* target, double* dev_out) {
    if (!comm) return stats(target, dev_out);
    const char* why = "";
    const NcclApi* api = nccl_api(&why);
    if (!api) return fail(AGF_ENCCL, why);
    AGF_CUDA(cudaSetDevice(opts.device));
    int nranks = 0;
    int rc = api->CommCount(reinterpret_cast<NcclComm>(comm), &nranks);
    if (rc != kNcclSuccess || nranks < 1) return fail(AGF_ENCCL, "ncclCommCount failed");
    int e = ensure_stats_buffers(nranks);
    if (e) return e;
    auto body = [&](const double* tgt) -> int {
      int r = stats(tgt, d_stats_local);
      if (r) return r;
      const int nr = api->AllGather(d_stats_local, d_stats_gather, AGF_STATS_LEN, kNcclFloat64, reinterpret_cast<NcclComm>(comm), stream);
      if (nr != kNcclSuccess) {
        char buf[200];
        snprintf(buf, sizeof buf, "ncclAllGather: %s", api->GetErrorString(nr));
        return fail(AGF_ENCCL, buf);
      }
      stats_combine_kernel<<<1, 32, 0, stream>>>(d_stats_gather, nranks, dev_out);
      AGF_CUDA(cudaGetLastError());
      launches++;
      return AGF_OK;
    };
    nccl_readouts++;
    static const bool no_graph = [] { const char* e = getenv("AGF_STATS_GRAPH"); return e && e[0] == '0'; }();  // AGF_STATS_GRAPH=0: eager launches
    if (target || graph_disabled || no_graph) return body(target);  // a host target is staged per call: not captured
    if (stats_graph && graph_comm == comm && graph_out == dev_out) {
      AGF_CUDA(cudaGraphLaunch(stats_graph, stream));
      launches += 2;
      return AGF_OK;
    }
    if (nccl_readouts < 2) return body(nullptr);  // first read-out eagerly: NCCL sets its connections up outside a capture
    // capture the three operations once; any refusal falls back to eager launches for good
    if (stats_graph) {
      cudaGraphExecDestroy(stats_graph);
      stats_graph = nullptr;
    }
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      graph_disabled = true;
      return body(nullptr);
    }
    const uint64_t l0 = launches;
    const int r = body(nullptr);
    const cudaError_t ce = cudaStreamEndCapture(stream, &g);
    launches = l0;
    if (r || ce != cudaSuccess || !g || cudaGraphInstantiate(&stats_graph, g, 0) != cudaSuccess) {
      cudaGetLastError();
      if (g) cudaGraphDestroy(g);
      stats_graph = nullptr;
      graph_disabled = true;
      return body(nullptr);
    }
    cudaGraphDestroy(g);
    {
      std::lock_guard<std::mutex> lock(g_graph_mutex);
      graph_comm = comm;
      graph_out = dev_out;
      if (std::find(g_graph_owners.begin(), g_graph_owners.end(), this) == g_graph_owners.end()) g_graph_owners.push_back(this);
    }
    AGF_CUDA(cudaGraphLaunch(stats_graph, stream));
    launches += 2;
    return AGF_OK;
  }
};

void release_stats_graphs_for(void* comm) {
  std::lock_guard<std::mutex> lock(g_graph_mutex);
  for (Batch* b : g_graph_owners)
    if (b->graph_comm == comm) {
      cudaSetDevice(b->opts.device);
      cudaStreamSynchronize(b->stream);
      b->drop_stats_graph();
    }
}

static size_t field_ncomp(int field) {
  switch (field) {
    case AGF_F_POSITION: case AGF_F_VELOCITY: case AGF_F_ANGULAR_VELOCITY: case AGF_F_EST_POSITION:
    case AGF_F_EST_VELOCITY: case AGF_F_EST_ANGULAR_VELOCITY: case AGF_F_ACCELEROMETER: case AGF_F_RATE_GYRO:
      return 3;
    case AGF_F_ATTITUDE: case AGF_F_MOTOR_SPEED: case AGF_F_MOTOR_SPEED_CMD: case AGF_F_EST_ATTITUDE:
    case AGF_F_MOTOR_FORCE: case AGF_F_KF_COUNTERS: case AGF_F_DES_MOTOR_FORCE:
      return 4;
    case AGF_F_FLIGHT_STATE: case AGF_F_PANIC_REASON: case AGF_F_CYCLE_COUNTER:
      return 1;
    case AGF_F_UWB_MEASUREMENT:
      return 2;
    case AGF_F_EST_COVARIANCE:
      return 81;
  }
  return 0;
}
static size_t field_elem(int field) {
  switch (field) {
    case AGF_F_POSITION: case AGF_F_VELOCITY: case AGF_F_ATTITUDE: case AGF_F_ANGULAR_VELOCITY:
    case AGF_F_MOTOR_SPEED: case AGF_F_MOTOR_FORCE:
      return 8;
    default:
      return 4;
  }
}

template<typename P>
struct BatchImpl : Batch {
  typedef typename VecOf<P>::type PV;
  static constexpr int VP = VecOf<P>::lanes;

  StateArrays<P> st{nullptr, nullptr, nullptr, nullptr, nullptr};
  PV* d_pv = nullptr;
  P* d_ext_force = nullptr;
  P* d_ext_torque = nullptr;
  uint32_t* d_tel_counter = nullptr;
  uint32_t* d_flags = nullptr;  // balanced-schedule hand-over flags, one per 32 vehicles (smallest block)
  uint32_t epoch = 0;
  SchedEntryDev* d_sched = nullptr;
  std::vector<SchedEntryDev> sched;
  float4* d_slot_f[AGF_MAX_CMD_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  uint32_t* d_slot_tf[AGF_MAX_CMD_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  P* d_log = nullptr;
  uint32_t log_stride = 0, log_cap = 0;
  size_t log_n = 0;  // vehicles per log array: n rounded up to whole 128-byte lines
  uint64_t log_base_tick = 0;
  void* d_stage = nullptr;
  size_t stage_bytes = 0;
  double* d_target = nullptr;
  agf_offboard_target* d_off_targets = nullptr;
  double* d_off_offsets = nullptr;
  uint64_t first_target_us = 0;
  std::vector<PackedPlan> h_plans;
  PackedPlan* d_plans = nullptr;
  uint32_t plans_cap = 0;
  uint32_t off_period_us = 0;  // periodOffboardMainLoop of the loop that was set (sizes the estimator's prediction pipe)
  double* d_off_est = nullptr;  // estimator state, est_doubles(n) values blocked by warp (agf_types.h est_index)
  double* d_off_state = nullptr;  // [AGF_OFFSTATE_DOUBLES][n]
  double* d_off_traj = nullptr;   // [AGF_OFFTRAJ_DOUBLES][n]

  StepShared<P> sh;
  PlantPV<P> pv_shared;
  Timing ts;
  bool per_vehicle = false;
  std::vector<agf_vehicle_cfg> cfgs;  // 1 or n
  std::vector<double> tau;            // motor time constants (1 or n)
  uint32_t c_for_dt_us = 0xFFFFFFFFu;
  bool uwb = false, hk = true, parity = true;

  ~BatchImpl() override {
    cudaSetDevice(opts.device);
    cudaFree(st.sp); cudaFree(st.sf); cudaFree(st.su); cudaFree(st.sc);
    cudaFree(d_pv); cudaFree(d_ext_force); cudaFree(d_ext_torque); cudaFree(d_tel_counter); cudaFree(d_flags);
    cudaFree(d_sched); cudaFree(d_log); cudaFree(d_stage); cudaFree(d_target); cudaFree(d_off_targets); cudaFree(d_off_offsets); cudaFree(d_off_state); cudaFree(d_off_traj); cudaFree(d_off_est); cudaFree(d_plans); cudaFree(st.sq);
    for (int s = 0; s < AGF_MAX_CMD_SLOTS; s++) { cudaFree(d_slot_f[s]); cudaFree(d_slot_tf[s]); }
    for (auto& e : events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    free_stats_buffers();
    if (own_stream && stream) cudaStreamDestroy(stream);
  }

  int ensure_stage(size_t bytes) {
    if (bytes <= stage_bytes) return AGF_OK;
    cudaFree(d_stage);
    d_stage = nullptr;
    stage_bytes = 0;
    AGF_CUDA(cudaMalloc(&d_stage, bytes));
    stage_bytes = bytes;
    return AGF_OK;
  }

  FieldCtx<P> ctx() const {
    FieldCtx<P> c;
    c.sp = (P*)st.sp; c.sf = (float*)st.sf; c.su = (uint32_t*)st.su; c.sc = (float*)st.sc;
    c.n = n;
    c.kF_shared = pv_shared.kF;
    c.pv = per_vehicle ? (const P*)d_pv : nullptr;
    return c;
  }

  // ---- construction --------------------------------------------------------------------------
  int init(const agf_vehicle_cfg* in, size_t n_cfgs, size_t n_vehicles, const agf_batch_opts& o) {
    n = n_vehicles;
    opts = o;
    parity = (o.math == AGF_MATH_PARITY);
    hk = parity || o.telemetry_warnings != 0;
    uwb = o.uwb_comm_period > 0;
    cfgs.assign(in, in + n_cfgs);
    per_vehicle = n_cfgs > 1;
    const agf_vehicle_cfg& c0 = cfgs[0];
    if (per_vehicle) {
      for (size_t i = 0; i < n_cfgs; i++) {
        const agf_vehicle_cfg& c = cfgs[i];
        bool ok = memcmp(&c.logic, &c0.logic, sizeof(c.logic)) == 0 && c.arm_length == c0.arm_length &&
                  memcmp(c.com_error, c0.com_error, sizeof(c.com_error)) == 0 &&
                  c.motor_min_speed == c0.motor_min_speed && c.motor_max_speed == c0.motor_max_speed &&
                  c.motor_inertia == c0.motor_inertia &&
                  memcmp(c.lin_drag_coeff_b, c0.lin_drag_coeff_b, sizeof(c.lin_drag_coeff_b)) == 0;
        for (int r = 0; r < 3 && ok; r++)
          for (int q = 0; q < 3; q++)
            if (r != q && c.inertia[3 * r + q] != 0.0) ok = false;
        if (!ok)
          return fail(AGF_EUNSUPPORTED,
                      "per-vehicle configs may differ only in mass, diagonal inertia, prop thrust/torque constants and "
                      "motor time constant (one airframe type per batch)");
      }
    }
    for (const auto& c : cfgs) {
      if (!(c.prop_thrust_from_speed_sqr >= 0) || !(c.prop_torque_from_speed_sqr >= 0) ||
          !(c.motor_max_speed > c.motor_min_speed))  // the asserts of Motor.cpp:27-29
        return fail(AGF_EINVAL, "motor constants violate the reference's constructor assertions (Motor.cpp:27-29)");
    }
    // shared parameters
    build_shared(c0, o.onboard_logic_period, o.uwb_comm_period, sh);
    set_noise_params(o.seed, o.sigma_gyro, o.sigma_acc, o.bias_sigma_gyro, o.bias_sigma_acc, o.uwb_noise_std_dev);
    // shared plant
    fill_plant(c0, pv_shared);
    tau.resize(cfgs.size());
    for (size_t i = 0; i < cfgs.size(); i++) tau[i] = cfgs[i].motor_time_const;
    memset(&ts, 0, sizeof(ts));

    // device allocations
    AGF_CUDA(cudaMalloc(&st.sp, sizeof(PV) * (NP_PAD / VP) * n));
    AGF_CUDA(cudaMalloc(&st.sf, sizeof(float4) * (NF_PAD / 4) * n));
    AGF_CUDA(cudaMalloc(&st.su, sizeof(uint4) * (NU_PAD / 4) * n));
    if (uwb) AGF_CUDA(cudaMalloc(&st.sc, sizeof(float4) * (NC_PAD / 4) * n));
    AGF_CUDA(cudaMalloc(&d_tel_counter, sizeof(uint32_t) * n));
    AGF_CUDA(cudaMemsetAsync(d_tel_counter, 0, sizeof(uint32_t) * n, stream));
    AGF_CUDA(cudaMalloc(&d_flags, sizeof(uint32_t) * ((n + 31) / 32)));
    AGF_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(uint32_t) * ((n + 31) / 32), stream));
    if (per_vehicle) AGF_CUDA(cudaMalloc(&d_pv, sizeof(PV) * (NPV_PAD / VP) * n));
    return init_state();
  }

  void set_noise_params(uint64_t seed, double sg, double sa, double bg, double ba, double su_) {
    sh.seed = seed;
    {  // Philox4x32-10 round keys, read by the kernels from the constant bank (agf_step.cuh philox_round_keys)
      uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
      for (int r = 0; r < 10; r++) {
        sh.philox_rk[2 * r] = k0;
        sh.philox_rk[2 * r + 1] = k1;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
      }
    }
    sh.sigma_gyro = float(sg);
    sh.sigma_acc = float(sa);
    sh.bias_sigma_gyro = float(bg);
    sh.bias_sigma_acc = float(ba);
    sh.uwb_sigma = float(su_);
    sh.bias_on = (bg != 0.0 || ba != 0.0) ? 1 : 0;
    sh.noise_on = (sg != 0.0 || sa != 0.0 || sh.bias_on) ? 1 : 0;
    sh.uwb_noise_on = (su_ != 0.0 || sh.uwb_outlier_prob > 0.0f) ? 1 : 0;
  }
  void set_uwb_noise(double noise, double p_out, double s_out) {
    opts.uwb_noise_std_dev = noise;
    sh.uwb_sigma = float(noise);
    sh.uwb_outlier_prob = float(p_out);
    sh.uwb_outlier_sigma = float(s_out);
    sh.uwb_noise_on = (noise != 0.0 || p_out > 0.0) ? 1 : 0;
  }

  // constructor-time state (SimulationObject6DOF.hpp:14-19, QuadcopterLogic::ResetCounters/Initialise,
  // KalmanFilter6DOF::Reset) replicated for every vehicle
  int init_state() {
    std::vector<P> hp;
    std::vector<float> hf, hc;
    std::vector<uint32_t> hu;
    initial_state<P>(n, sh.logic, cfgs[0].logic.low_battery_threshold, uwb, hp, hf, hu, hc);
    AGF_CUDA(cudaMemcpyAsync(st.sp, hp.data(), hp.size() * sizeof(P), cudaMemcpyHostToDevice, stream));
    AGF_CUDA(cudaMemcpyAsync(st.sf, hf.data(), hf.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
    AGF_CUDA(cudaMemcpyAsync(st.su, hu.data(), hu.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    if (uwb) AGF_CUDA(cudaMemcpyAsync(st.sc, hc.data(), hc.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
    AGF_CUDA(cudaStreamSynchronize(stream));
    return AGF_OK;
  }

  int refresh_motor_c(uint32_t dt_us) {
    if (dt_us == c_for_dt_us) return AGF_OK;
    pv_shared.motor_c = P(motor_c_host(tau[0], dt_us));
    if (per_vehicle) {
      std::vector<P> h(size_t(NPV_PAD) * n, P(0));
      for (size_t i = 0; i < n; i++) {
        const agf_vehicle_cfg& c = cfgs[i];
        double I[9] = {c.inertia[0], 0, 0, 0, c.inertia[4], 0, 0, 0, c.inertia[8]}, inv[9];
        inverse3(I, inv);
        h[sidx(PV_MASS, n, i, VP)] = P(c.mass);
        h[sidx(PV_IXX, n, i, VP)] = P(c.inertia[0]);
        h[sidx(PV_IYY, n, i, VP)] = P(c.inertia[4]);
        h[sidx(PV_IZZ, n, i, VP)] = P(c.inertia[8]);
        h[sidx(PV_IIXX, n, i, VP)] = P(inv[0]);
        h[sidx(PV_IIYY, n, i, VP)] = P(inv[4]);
        h[sidx(PV_IIZZ, n, i, VP)] = P(inv[8]);
        h[sidx(PV_KF, n, i, VP)] = P(c.prop_thrust_from_speed_sqr);
        h[sidx(PV_KTAU, n, i, VP)] = P(c.prop_torque_from_speed_sqr);
        h[sidx(PV_MOTOR_C, n, i, VP)] = P(motor_c_host(tau[i], dt_us));
        h[sidx(PV_INV_MASS, n, i, VP)] = P(1.0 / c.mass);
      }
      AGF_CUDA(cudaMemcpyAsync(d_pv, h.data(), h.size() * sizeof(P), cudaMemcpyHostToDevice, stream));
      AGF_CUDA(cudaStreamSynchronize(stream));
    }
    c_for_dt_us = dt_us;
    return AGF_OK;
  }

  // ---- stepping ------------------------------------------------------------------------------
  int launch(uint32_t dt_us, uint32_t nticks) {
    StepLaunch<P> L;
    memset(&L, 0, sizeof(L));
    sh.ext_force = d_ext_force;
    sh.ext_torque = d_ext_torque;
    L.sh = sh;
    L.st = st;
    L.pv = per_vehicle ? d_pv : nullptr;
    L.pv_shared = pv_shared;
    L.n = n;
    L.tick0 = ticks;
    L.nticks = nticks;
    L.dt_us = dt_us;
    ts.now_us = now_us;
    // the launch's tick plans: the clock-only stopwatches evolved here, once, with the function the whole design shares
    {
      h_plans.resize(nticks);
      Timing tt = ts;
      for (uint32_t t = 0; t < nticks; t++) {
        const TickPlan p = timing_plan(tt, sh.tc, dt_us);
        h_plans[t] = pack_plan(p);
        timing_advance(tt, sh.tc, p, dt_us);
      }
      if (nticks > plans_cap) {
        cudaFree(d_plans);
        d_plans = nullptr;
        plans_cap = 0;
        AGF_CUDA(cudaMalloc(&d_plans, size_t(nticks) * sizeof(PackedPlan)));
        plans_cap = nticks;
      }
      // pageable source: the runtime stages it before returning, so h_plans may be reused by the next launch
      AGF_CUDA(cudaMemcpyAsync(d_plans, h_plans.data(), size_t(nticks) * sizeof(PackedPlan), cudaMemcpyHostToDevice, stream));
      L.plans = reinterpret_cast<const uint4*>(d_plans);
      L.now0_us = now_us;
    }
    L.sched = d_sched;
    // entries of this launch: [begin, end) with tick in [ticks, ticks + nticks)
    auto lo = std::lower_bound(sched.begin(), sched.end(), ticks, [](const SchedEntryDev& e, uint64_t t) { return e.tick < t; });
    auto hi = std::lower_bound(sched.begin(), sched.end(), ticks + nticks, [](const SchedEntryDev& e, uint64_t t) { return e.tick < t; });
    L.sched_begin = uint32_t(lo - sched.begin());
    L.sched_end = uint32_t(hi - sched.begin());
    for (auto it = lo; it != hi; ++it)
      if (it->slot >= 0 && !d_slot_f[it->slot]) return fail(AGF_EINVAL, "schedule references a command slot that was never set");
    for (int s = 0; s < AGF_MAX_CMD_SLOTS; s++) { L.slots[s].f = d_slot_f[s]; L.slots[s].tf = d_slot_tf[s]; }
    L.log = d_log;
    L.log_n = log_n;
    L.log_stride = log_stride ? log_stride : 1;
    L.log_capacity = log_cap ? log_cap : 1;
    L.log_first_off = L.log_stride - 1 - uint32_t(ticks % L.log_stride);
    L.log_slot0 = uint32_t(((ticks + L.log_first_off + 1) / L.log_stride - 1) % L.log_capacity);
    L.first_global_index = opts.first_global_index;
    L.flags = d_flags;
    L.epoch = ++epoch;

    std::pair<cudaEvent_t, cudaEvent_t>* ev = nullptr;
    {
      int rc = next_event_pair(&ev);
      if (rc) return rc;
    }
    const int block = opts.block_threads > 0 ? std::min(opts.block_threads, AGF_BLOCK_THREADS) : AGF_BLOCK_THREADS;
    AGF_CUDA(cudaEventRecord(ev->first, stream));
    cudaError_t e = do_launch(L, block);
    if (e != cudaSuccess) return fail(AGF_ECUDA, "step kernel launch", e);
    AGF_CUDA(cudaEventRecord(ev->second, stream));
    launches++;
    // carry the clock-only stopwatches across the launch with the same function the device uses
    for (uint32_t t = 0; t < nticks; t++) {
      const TickPlan p = timing_plan(ts, sh.tc, dt_us);
      timing_advance(ts, sh.tc, p, dt_us);
    }
    if (d_log) log_records = (ticks + nticks) / log_stride - log_base_tick / log_stride;
    ticks += nticks;
    now_us += uint64_t(dt_us) * nticks;
    return AGF_OK;
  }

  cudaError_t do_launch(const StepLaunch<P>& L, int block);

  // BaseTimer advancing between two Run() calls (ManualTimer::AdvanceMicroSeconds, ManualTimer.hpp:29)
  int advance_clock(uint32_t dt_us) override {
    if (!dt_us) return AGF_OK;
    AGF_CUDA(cudaSetDevice(opts.device));
    advance_clock_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(st.su, n, dt_us);
    AGF_CUDA(cudaGetLastError());
    launches++;
    ts.integ_age += dt_us;
    ts.logic_age += dt_us;
    ts.kf_age += dt_us;
    ts.net_age += dt_us;
    now_us += dt_us;
    ts.now_us = now_us;
    if (sh.tc.off_enabled && sh.tc.mocap_enabled) {  // simulated mocap packet (main.cpp:451-457), before the main loop
      ts.mocap_age += dt_us;
      if (ts.mocap_age >= sh.tc.mocap_min_age_us) {
        ts.mocap_age -= sh.tc.mocap_adj_us;
        cudaError_t e = launch_offboard_mocap(st, n, sh.off.est, now_us, stream);
        if (e != cudaSuccess) return fail(AGF_ECUDA, "mocap update kernel launch", e);
        launches++;
      }
    }
    if (sh.tc.off_enabled) {  // the offboard main loop acts after the clock advance (main.cpp:392,471-673)
      for (uint32_t q = 0; q < AGF_OFFQ; q++) ts.off_wait[q] = ts.off_wait[q] > dt_us ? ts.off_wait[q] - dt_us : 0;
      ts.off_age += dt_us;
      if (ts.off_age >= sh.tc.off_min_age_us) {
        ts.off_age -= sh.tc.off_adj_us;
        if (now_us >= sh.tc.off_first_target_us) {
          if (ts.off_count >= AGF_OFFQ) return fail(AGF_EUNSUPPORTED, "offboard loop: command queue full (Run() not called between clock advances)");
          const uint32_t slot = (ts.off_head + ts.off_count) % AGF_OFFQ;
          cudaError_t e = launch_offboard_generate(st, n, sh.off, now_us, slot, stream);
          if (e != cudaSuccess) return fail(AGF_ECUDA, "offboard command kernel launch", e);
          launches++;
          ts.off_wait[slot] = sh.tc.off_delay_us;
          ts.off_count++;
        }
      }
    }
    return AGF_OK;
  }

  int run(uint32_t dt_us, uint32_t nticks) override {
    if (dt_us == 0 && nticks > 1) return fail(AGF_EINVAL, "dt_us == 0 (Run() without a clock advance) takes nticks == 1");
    AGF_CUDA(cudaSetDevice(opts.device));
    while (nticks) {
      // the plant step of the first tick integrates over the time since the previous Run(); the motor lag
      // coefficient exp(-dt/tau) is a launch constant, so a launch never mixes two different dt
      const uint32_t first_dt = ts.integ_age;
      uint32_t chunk = nticks, c_dt = dt_us;
      if (first_dt != 0 && first_dt != dt_us) {
        chunk = 1;
        c_dt = first_dt;
      }
      int rc = refresh_motor_c(c_dt);
      if (rc) return rc;
      rc = launch(dt_us, chunk);
      if (rc) return rc;
      nticks -= chunk;
    }
    return AGF_OK;
  }

  // ---- offboard loop -------------------------------------------------------------------------
  int set_offboard(const agf_offboard_cfg* cfg, const agf_offboard_target* targets, size_t n_targets, const double* offsets) override {
    AGF_CUDA(cudaSetDevice(opts.device));
    AGF_CUDA(cudaStreamSynchronize(stream));
    if (!cfg) {
      sh.tc.off_enabled = 0;
      sh.tc.mocap_enabled = 0;
      sh.off.est.kind = AGF_OFFEST_TRUTH;
      ts.off_age = ts.off_head = ts.off_count = 0;
      memset(ts.off_wait, 0, sizeof(ts.off_wait));
      return AGF_OK;
    }
    if (!targets || !n_targets) return fail(AGF_EINVAL, "offboard loop needs at least one target");
    for (size_t k = 1; k < n_targets; k++)
      if (targets[k].time_us <= targets[k - 1].time_us) return fail(AGF_EINVAL, "offboard targets must be strictly increasing in time");
    OffboardParams off;
    memset(&off, 0, sizeof(off));
    TimingConsts tc = sh.tc;
    const char* why = fill_offboard(*cfg, off, tc);
    if (why) return fail(AGF_EINVAL, why);
    off_period_us = cfg->period_us;
    cudaFree(d_off_targets);
    d_off_targets = nullptr;
    AGF_CUDA(cudaMalloc(&d_off_targets, n_targets * sizeof(agf_offboard_target)));
    AGF_CUDA(cudaMemcpyAsync(d_off_targets, targets, n_targets * sizeof(agf_offboard_target), cudaMemcpyHostToDevice, stream));
    cudaFree(d_off_offsets);
    d_off_offsets = nullptr;
    if (offsets) {
      std::vector<double> soa(3 * n);
      for (size_t i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) soa[size_t(k) * n + i] = offsets[3 * i + k];
      AGF_CUDA(cudaMalloc(&d_off_offsets, soa.size() * sizeof(double)));
      AGF_CUDA(cudaMemcpyAsync(d_off_offsets, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
      AGF_CUDA(cudaStreamSynchronize(stream));
    }
    if (!st.sq) {
      AGF_CUDA(cudaMalloc(&st.sq, sizeof(float4) * AGF_OFFQ * n));
      AGF_CUDA(cudaMemsetAsync(st.sq, 0, sizeof(float4) * AGF_OFFQ * n, stream));
    }
    AGF_CUDA(cudaStreamSynchronize(stream));
    const bool was_on = sh.tc.off_enabled != 0;
    off.targets = d_off_targets;
    off.n_targets = uint32_t(n_targets);
    off.offsets = d_off_offsets;
    // the reference generator survives a change of the loop parameters
    off.ref_kind = sh.off.ref_kind; off.traj_id = sh.off.traj_id;
    off.start_us = sh.off.start_us; off.stop_us = sh.off.stop_us;
    for (int k = 0; k < 3; k++) off.desired[k] = sh.off.desired[k];
    off.desired_yaw = sh.off.desired_yaw;
    off.safety_net = sh.off.safety_net;
    for (int k = 0; k < 3; k++) { off.safe_min[k] = sh.off.safe_min[k]; off.safe_max[k] = sh.off.safe_max[k]; }
    off.min_normal_height = sh.off.min_normal_height;
    off.not_seen_timeout = sh.off.not_seen_timeout;
    off.state = d_off_state;
    off.traj = d_off_traj;
    off.est = sh.off.est;
    tc.mocap_enabled = sh.tc.mocap_enabled;
    tc.mocap_min_age_us = sh.tc.mocap_min_age_us;
    tc.mocap_adj_us = sh.tc.mocap_adj_us;
    sh.off = off;
    sh.tc = tc;
    sh.tc.off_first_target_us = off.ref_kind == AGF_OFFREF_TARGETS ? targets[0].time_us : 0;
    first_target_us = targets[0].time_us;
    if (!was_on) {  // the loop's Timer and queue are created now (Timer::Timer resets to the current clock reading)
      ts.off_age = ts.off_head = ts.off_count = 0;
      memset(ts.off_wait, 0, sizeof(ts.off_wait));
    }
    return AGF_OK;
  }

  int set_offboard_ref(const agf_offboard_ref* ref) override {
    AGF_CUDA(cudaSetDevice(opts.device));
    AGF_CUDA(cudaStreamSynchronize(stream));
    if (!sh.tc.off_enabled) return fail(AGF_EINVAL, "set the offboard loop (agf_batch_set_offboard_loop) before its reference generator");
    if (!ref || ref->kind == AGF_OFFREF_TARGETS) {
      sh.off.ref_kind = AGF_OFFREF_TARGETS;
      sh.tc.off_first_target_us = first_target_us;
      return AGF_OK;
    }
    if (ref->kind != AGF_OFFREF_STAGES && ref->kind != AGF_OFFREF_TRAJECTORY) return fail(AGF_EINVAL, "unknown reference generator");
    if (ref->kind == AGF_OFFREF_STAGES && (ref->traj_id < 0 || ref->traj_id > 5)) return fail(AGF_EINVAL, "traj_id must be 0..5");
    if (ref->kind == AGF_OFFREF_TRAJECTORY && !d_off_traj) return fail(AGF_EINVAL, "set the trajectories (agf_batch_set_offboard_trajectories) first");
    if (ref->kind == AGF_OFFREF_STAGES) {  // a fresh state machine (ExampleVehicleStateMachine::ExampleVehicleStateMachine, .cpp:5-26)
      std::vector<double> h(size_t(AGF_OFFSTATE_DOUBLES) * n, 0.0);
      for (size_t i = 0; i < n; i++) {
        h[0 * n + i] = AGF_STAGE_WAIT_FOR_START;
        h[1 * n + i] = AGF_STAGE_COMPLETE;
        h[2 * n + i] = double(now_us);
        h[15 * n + i] = 0.0;
      }
      if (!d_off_state) AGF_CUDA(cudaMalloc(&d_off_state, h.size() * sizeof(double)));
      AGF_CUDA(cudaMemcpy(d_off_state, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    sh.off.ref_kind = ref->kind;
    sh.off.traj_id = ref->traj_id;
    sh.off.start_us = ref->start_us;
    sh.off.stop_us = ref->stop_us;
    for (int k = 0; k < 3; k++) sh.off.desired[k] = ref->desired_pos[k];
    sh.off.desired_yaw = ref->desired_yaw;
    sh.off.safety_net = ref->kind == AGF_OFFREF_STAGES ? ref->safety_net : 0;
    for (int k = 0; k < 3; k++) { sh.off.safe_min[k] = ref->safe_min[k]; sh.off.safe_max[k] = ref->safe_max[k]; }
    sh.off.min_normal_height = ref->min_normal_height;
    sh.off.not_seen_timeout = ref->not_seen_timeout;
    sh.off.state = d_off_state;
    sh.off.traj = d_off_traj;
    sh.tc.off_first_target_us = 0;
    return AGF_OK;
  }
  int set_offboard_estimator(const agf_offboard_estimator* e) override {
    AGF_CUDA(cudaSetDevice(opts.device));
    AGF_CUDA(cudaStreamSynchronize(stream));
    if (!sh.tc.off_enabled) return fail(AGF_EINVAL, "set the offboard loop (agf_batch_set_offboard_loop) before its estimator");
    if (!e || e->kind == AGF_OFFEST_TRUTH) {
      sh.off.est.kind = AGF_OFFEST_TRUTH;
      sh.tc.mocap_enabled = 0;
      return AGF_OK;
    }
    if (e->kind != AGF_OFFEST_MOCAP) return fail(AGF_EINVAL, "unknown estimator kind");
    if (!(e->mocap_period_us > 0) || !(e->angvel_time_const > 0) || !(e->prediction_delay >= 0))
      return fail(AGF_EINVAL, "estimator: mocap period and angular-velocity time constant must be positive");
    {
      // The reference's PredictionPipe is unbounded; here it has AGF_OFFEST_PIPE slots.  In flight at most: the messages that
      // become active within the prediction delay, those added between two mocap packets, the active one and the newest.
      const uint64_t per = off_period_us ? off_period_us : 1;
      const uint64_t delay_us = uint64_t(e->prediction_delay * 1e6 + 0.5);
      const uint64_t need = (delay_us + per - 1) / per + (uint64_t(e->mocap_period_us) + per - 1) / per + 2;
      if (need > uint64_t(AGF_OFFEST_PIPE))
        return fail(AGF_EUNSUPPORTED, "estimator: ceil(prediction_delay / loop period) + ceil(mocap period / loop period) + 2 prediction "
                                      "messages in flight exceed AGF_OFFEST_PIPE");
    }
    // MocapStateEstimator::MocapStateEstimator -> Reset() (MocapStateEstimator.cpp:9-50) for every vehicle
    std::vector<double> h(est_doubles(n), 0.0);
    for (size_t i = 0; i < n; i++) {
      double* e0 = h.data() + est_index(i);  // field k at e0[k * E_LANES]
      e0[(E_ATT + 0) * E_LANES] = 1.0;
      e0[(E_VP + 0) * E_LANES] = 25.0; e0[(E_VP + 3) * E_LANES] = 25.0;
      e0[(E_VA + 0) * E_LANES] = 1.0; e0[(E_VA + 3) * E_LANES] = 400.0;
      e0[E_LASTGOOD * E_LANES] = double(now_us);
      for (int k = 0; k < AGF_OFFEST_PIPE; k++) e0[(E_PIPE + E_MSG * k) * E_LANES] = E_SLOT_FREE;  // empty prediction pipe
    }
    if (!d_off_est) AGF_CUDA(cudaMalloc(&d_off_est, h.size() * sizeof(double)));
    AGF_CUDA(cudaMemcpy(d_off_est, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    EstParams& p = sh.off.est;
    p.kind = AGF_OFFEST_MOCAP;
    p.t0_us = now_us;
    p.delay = e->prediction_delay;
    p.reject = e->meas_reject_dist;
    p.tc_angvel = e->angvel_time_const;
    p.inv_tc_angvel = 1.0 / e->angvel_time_const;
    p.meas_pos = e->meas_noise_pos; p.meas_att = e->meas_noise_att;
    p.proc_pos = e->proc_noise_pos; p.proc_att = e->proc_noise_att;
    p.state = d_off_est;
    sh.tc.mocap_enabled = 1;
    timing_thresholds_mocap(sh.tc, double(e->mocap_period_us) * 1e-6);
    ts.mocap_age = 0;  // timerMocap is created now (main.cpp:286)
    return AGF_OK;
  }
  int get_offboard_estimate(double horizon, double* est13, double* counters4, size_t first, size_t count) override {
    if (!est13) return fail(AGF_EINVAL, "null output");
    if (first + count > n) return fail(AGF_ERANGE, "vehicle range outside the batch");
    if (sh.off.est.kind != AGF_OFFEST_MOCAP || !d_off_est) return fail(AGF_EINVAL, "no offboard estimator is set");
    if (!count) return AGF_OK;
    AGF_CUDA(cudaSetDevice(opts.device));
    double* d_out = nullptr;
    AGF_CUDA(cudaMalloc(&d_out, 13 * count * sizeof(double)));
    cudaError_t e = launch_offboard_estimate(sh.off.est, n, first, count, now_us, horizon, d_out, stream);
    launches++;
    std::vector<double> h(13 * count);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), d_out, h.size() * sizeof(double), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(AGF_ECUDA, "estimator read-out", e);
    for (size_t k = 0; k < count; k++)
      for (int f = 0; f < 13; f++) est13[k * 13 + f] = h[size_t(f) * count + k];
    if (counters4) {
      if (int rc = ensure_stage(4 * count * sizeof(double))) return rc;
      AGF_CUDA(launch_offboard_counters(d_off_est, first, count, (double*)d_stage, stream));
      launches++;
      AGF_CUDA(cudaMemcpyAsync(counters4, d_stage, 4 * count * sizeof(double), cudaMemcpyDeviceToHost, stream));
      AGF_CUDA(cudaStreamSynchronize(stream));
    }
    return AGF_OK;
  }
  int set_offboard_traj(const double* traj, size_t first, size_t count) override {
    if (!traj) return fail(AGF_EINVAL, "null trajectory array");
    if (first + count > n) return fail(AGF_ERANGE, "vehicle range outside the batch");
    AGF_CUDA(cudaSetDevice(opts.device));
    AGF_CUDA(cudaStreamSynchronize(stream));
    if (!d_off_traj) {
      AGF_CUDA(cudaMalloc(&d_off_traj, size_t(AGF_OFFTRAJ_DOUBLES) * n * sizeof(double)));
      AGF_CUDA(cudaMemset(d_off_traj, 0, size_t(AGF_OFFTRAJ_DOUBLES) * n * sizeof(double)));
      sh.off.traj = d_off_traj;
    }
    std::vector<double> col(count);
    for (int k = 0; k < AGF_OFFTRAJ_DOUBLES; k++) {  // AoS in, [field][vehicle] on the device
      for (size_t i = 0; i < count; i++) col[i] = traj[i * AGF_OFFTRAJ_DOUBLES + k];
      AGF_CUDA(cudaMemcpy(d_off_traj + size_t(k) * n + first, col.data(), count * sizeof(double), cudaMemcpyHostToDevice));
    }
    return AGF_OK;
  }
  int offboard_traj_ptr(double** p, size_t* nv) override {
    if (!p) return fail(AGF_EINVAL, "null output");
    AGF_CUDA(cudaSetDevice(opts.device));
    if (!d_off_traj) {
      AGF_CUDA(cudaMalloc(&d_off_traj, size_t(AGF_OFFTRAJ_DOUBLES) * n * sizeof(double)));
      AGF_CUDA(cudaMemset(d_off_traj, 0, size_t(AGF_OFFTRAJ_DOUBLES) * n * sizeof(double)));
      sh.off.traj = d_off_traj;
    }
    *p = d_off_traj;
    if (nv) *nv = n;
    return AGF_OK;
  }
  int get_offboard_state(double* out, size_t first, size_t count) override {
    if (!out) return fail(AGF_EINVAL, "null output");
    if (first + count > n) return fail(AGF_ERANGE, "vehicle range outside the batch");
    if (!d_off_state) return fail(AGF_EINVAL, "no stage reference generator is set");
    AGF_CUDA(cudaSetDevice(opts.device));
    AGF_CUDA(cudaStreamSynchronize(stream));
    std::vector<double> col(count);
    for (int k = 0; k < AGF_OFFSTATE_DOUBLES; k++) {
      AGF_CUDA(cudaMemcpy(col.data(), d_off_state + size_t(k) * n + first, count * sizeof(double), cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < count; i++) out[i * AGF_OFFSTATE_DOUBLES + k] = col[i];
    }
    return AGF_OK;
  }

  // ---- field access --------------------------------------------------------------------------
  int get_field(int field, void* dst, size_t first, size_t count) override {
    const size_t nc = field_ncomp(field);
    if (!nc) return fail(AGF_EINVAL, "unknown field");
    if (first + count > n) return fail(AGF_ERANGE, "vehicle range outside the batch");
    if (!count) return AGF_OK;
    AGF_CUDA(cudaSetDevice(opts.device));
    const size_t bytes = count * nc * field_elem(field);
    int rc = ensure_stage(bytes);
    if (rc) return rc;
    if (field == AGF_F_EST_COVARIANCE && !uwb) {
      // no ranging: the covariance stays at its Reset() value (KalmanFilter6DOF.cpp:42-61)
      float* o = (float*)dst;
      const float sp = 3.0f, sperp = 10.0f * float(M_PI) / 180.0f, sabout = 30.0f * float(M_PI) / 180.0f;
      for (size_t v = 0; v < count; v++) {
        float* m = o + 81 * v;
        memset(m, 0, 81 * sizeof(float));
        for (int k = 0; k < 6; k++) m[9 * k + k] = sp * sp;
        m[9 * 6 + 6] = m[9 * 7 + 7] = sperp * sperp;
        m[9 * 8 + 8] = sabout * sabout;
      }
      return AGF_OK;
    }
    const size_t total = count * nc;
    field_get_kernel<P><<<unsigned((total + 255) / 256), 256, 0, stream>>>(ctx(), field, int(nc), first, count, d_stage);
    AGF_CUDA(cudaGetLastError());
    launches++;
    AGF_CUDA(cudaMemcpyAsync(dst, d_stage, bytes, cudaMemcpyDeviceToHost, stream));
    AGF_CUDA(cudaStreamSynchronize(stream));
    return AGF_OK;
  }

  int set_field(int field, const void* src, size_t first, size_t count) override {
    switch (field) {
      case AGF_F_POSITION: case AGF_F_VELOCITY: case AGF_F_ATTITUDE: case AGF_F_ANGULAR_VELOCITY:
      case AGF_F_MOTOR_SPEED: case AGF_F_MOTOR_SPEED_CMD: case AGF_F_EST_POSITION: case AGF_F_EST_VELOCITY:
      case AGF_F_EST_ATTITUDE: case AGF_F_EST_ANGULAR_VELOCITY:
        break;
      default:
        return fail(AGF_EINVAL, "field is read-only or unknown");
    }
    const size_t nc = field_ncomp(field);
    if (first + count > n) return fail(AGF_ERANGE, "vehicle range outside the batch");
    if (!count) return AGF_OK;
    AGF_CUDA(cudaSetDevice(opts.device));
    const size_t bytes = count * nc * field_elem(field);
    int rc = ensure_stage(bytes);
    if (rc) return rc;
    AGF_CUDA(cudaMemcpyAsync(d_stage, src, bytes, cudaMemcpyHostToDevice, stream));
    const size_t total = count * nc;
    field_set_kernel<P><<<unsigned((total + 255) / 256), 256, 0, stream>>>(ctx(), field, int(nc), first, count, d_stage);
    AGF_CUDA(cudaGetLastError());
    launches++;
    AGF_CUDA(cudaStreamSynchronize(stream));
    return AGF_OK;
  }

  int set_state13(const double* src, size_t first, size_t count) override {
    if (!src) return fail(AGF_EINVAL, "null state array");
    if (first + count > n) return fail(AGF_ERANGE, "vehicle range outside the batch");
    if (!count) return AGF_OK;
    AGF_CUDA(cudaSetDevice(opts.device));
    const size_t bytes = count * 13 * sizeof(double);
    int rc = ensure_stage(bytes);
    if (rc) return rc;
    AGF_CUDA(cudaMemcpyAsync(d_stage, src, bytes, cudaMemcpyHostToDevice, stream));
    const size_t total = count * 13;
    state13_set_kernel<P><<<unsigned((total + 255) / 256), 256, 0, stream>>>(ctx(), first, count, reinterpret_cast<const double*>(d_stage));
    AGF_CUDA(cudaGetLastError());
    launches++;
    AGF_CUDA(cudaStreamSynchronize(stream));  // the caller may reuse its buffer
    return AGF_OK;
  }

  int get_state13(double* dst, size_t first, size_t count) override {
    if (!dst) return fail(AGF_EINVAL, "null state array");
    if (first + count > n) return fail(AGF_ERANGE, "vehicle range outside the batch");
    if (!count) return AGF_OK;
    AGF_CUDA(cudaSetDevice(opts.device));
    const size_t bytes = count * 13 * sizeof(double);
    int rc = ensure_stage(bytes);
    if (rc) return rc;
    const size_t total = count * 13;
    state13_get_kernel<P><<<unsigned((total + 255) / 256), 256, 0, stream>>>(ctx(), first, count, reinterpret_cast<double*>(d_stage));
    AGF_CUDA(cudaGetLastError());
    launches++;
    AGF_CUDA(cudaMemcpyAsync(dst, d_stage, bytes, cudaMemcpyDeviceToHost, stream));
    AGF_CUDA(cudaStreamSynchronize(stream));
    return AGF_OK;
  }

  // ---- radio -----------------------------------------------------------------------------------
  static RadioNow decode_now(const uint8_t* raw) {
    RadioNow m;
    uint8_t type, flags;
    float f[AGF_RADIO_NUM_FLOATS];
    agf_radio_decode(raw, &type, &flags, f);
    m.type = type;
    m.flags = flags;
    for (int k = 0; k < 4; k++) m.f[k] = f[k];
    return m;
  }

  int set_radio(const uint8_t* raw, size_t first, size_t count, int broadcast) override {
    if (!raw) return fail(AGF_EINVAL, "raw is null");
    if (first + count > n) return fail(AGF_ERANGE, "vehicle range outside the batch");
    if (!count) return AGF_OK;
    AGF_CUDA(cudaSetDevice(opts.device));
    RadioNow one;
    memset(&one, 0, sizeof(one));
    const RadioNow* per = nullptr;
    if (broadcast) {
      one = decode_now(raw);
    } else {
      std::vector<RadioNow> h(count);
      for (size_t i = 0; i < count; i++) h[i] = decode_now(raw + i * AGF_RADIO_PACKET_SIZE);
      int rc = ensure_stage(count * sizeof(RadioNow));
      if (rc) return rc;
      AGF_CUDA(cudaMemcpyAsync(d_stage, h.data(), count * sizeof(RadioNow), cudaMemcpyHostToDevice, stream));
      AGF_CUDA(cudaStreamSynchronize(stream));
      per = (const RadioNow*)d_stage;
    }
    radio_now_kernel<<<unsigned((count + 255) / 256), 256, 0, stream>>>((float*)st.sf, (uint32_t*)st.su, n, first, count, broadcast,
                                                                      one, per, sh.logic.mon_cmd_coef, hk ? 1 : 0);
    AGF_CUDA(cudaGetLastError());
    launches++;
    if (!broadcast) AGF_CUDA(cudaStreamSynchronize(stream));
    return AGF_OK;
  }

  int set_schedule(const agf_cmd_entry* e, size_t ne) override {
    std::vector<SchedEntryDev> v(ne);
    for (size_t i = 0; i < ne; i++) {
      if (i && e[i].tick <= e[i - 1].tick) return fail(AGF_EINVAL, "schedule must be strictly increasing in tick");
      if (e[i].slot >= AGF_MAX_CMD_SLOTS) return fail(AGF_EINVAL, "schedule slot out of range");
      memset(&v[i], 0, sizeof(v[i]));
      v[i].tick = e[i].tick;
      v[i].slot = e[i].slot < 0 ? -1 : e[i].slot;
      if (e[i].slot < 0) {
        const RadioNow m = decode_now(e[i].raw);
        v[i].type = m.type;
        v[i].flags = m.flags;
        for (int k = 0; k < 4; k++) v[i].f[k] = m.f[k];
      }
    }
    AGF_CUDA(cudaSetDevice(opts.device));
    AGF_CUDA(cudaStreamSynchronize(stream));
    cudaFree(d_sched);
    d_sched = nullptr;
    sched.swap(v);
    if (ne) {
      AGF_CUDA(cudaMalloc(&d_sched, ne * sizeof(SchedEntryDev)));
      AGF_CUDA(cudaMemcpyAsync(d_sched, sched.data(), ne * sizeof(SchedEntryDev), cudaMemcpyHostToDevice, stream));
      AGF_CUDA(cudaStreamSynchronize(stream));
    }
    return AGF_OK;
  }

  int set_slot(int slot, const uint8_t* raw) override {
    if (slot < 0 || slot >= AGF_MAX_CMD_SLOTS) return fail(AGF_EINVAL, "slot out of range");
    if (!raw) return fail(AGF_EINVAL, "raw is null");
    AGF_CUDA(cudaSetDevice(opts.device));
    std::vector<float4> hf(n);
    std::vector<uint32_t> ht(n);
    for (size_t i = 0; i < n; i++) {
      const RadioNow m = decode_now(raw + i * AGF_RADIO_PACKET_SIZE);
      hf[i] = make_float4(m.f[0], m.f[1], m.f[2], m.f[3]);
      ht[i] = (m.type & 0xFFu) | ((m.flags & 0xFFu) << 8);
    }
    if (!d_slot_f[slot]) {
      AGF_CUDA(cudaMalloc(&d_slot_f[slot], n * sizeof(float4)));
      AGF_CUDA(cudaMalloc(&d_slot_tf[slot], n * sizeof(uint32_t)));
    }
    AGF_CUDA(cudaMemcpyAsync(d_slot_f[slot], hf.data(), n * sizeof(float4), cudaMemcpyHostToDevice, stream));
    AGF_CUDA(cudaMemcpyAsync(d_slot_tf[slot], ht.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    AGF_CUDA(cudaStreamSynchronize(stream));
    return AGF_OK;
  }

  int telemetry(uint8_t* p1, uint8_t* p2, size_t first, size_t count) override {
    if (!p1 || !p2) return fail(AGF_EINVAL, "null packet buffer");
    if (first + count > n) return fail(AGF_ERANGE, "vehicle range outside the batch");
    if (!count) return AGF_OK;
    AGF_CUDA(cudaSetDevice(opts.device));
    const size_t bytes = count * AGF_TELEMETRY_PACKET_SIZE;
    int rc = ensure_stage(2 * bytes);
    if (rc) return rc;
    uint8_t* d1 = (uint8_t*)d_stage;
    uint8_t* d2 = d1 + bytes;
    telemetry_kernel<<<unsigned((count + 127) / 128), 128, 0, stream>>>((const float*)st.sf, (uint32_t*)st.su, d_tel_counter, n, first,
                                                                      count, sh.logic.batt_voltage, hk ? 1 : 0, d1, d2);
    AGF_CUDA(cudaGetLastError());
    launches++;
    AGF_CUDA(cudaMemcpyAsync(p1, d1, bytes, cudaMemcpyDeviceToHost, stream));
    AGF_CUDA(cudaMemcpyAsync(p2, d2, bytes, cudaMemcpyDeviceToHost, stream));
    AGF_CUDA(cudaStreamSynchronize(stream));
    return AGF_OK;
  }

  int set_wrench(const double* f, const double* t, size_t first, size_t count) override {
    if (first + count > n) return fail(AGF_ERANGE, "vehicle range outside the batch");
    AGF_CUDA(cudaSetDevice(opts.device));
    if (!d_ext_force) {
      AGF_CUDA(cudaMalloc(&d_ext_force, sizeof(P) * 3 * n));
      AGF_CUDA(cudaMalloc(&d_ext_torque, sizeof(P) * 3 * n));
      AGF_CUDA(cudaMemsetAsync(d_ext_force, 0, sizeof(P) * 3 * n, stream));
      AGF_CUDA(cudaMemsetAsync(d_ext_torque, 0, sizeof(P) * 3 * n, stream));
    }
    if (!count || (!f && !t)) return AGF_OK;
    // one staged copy per array and one kernel (the reference's setters are per vehicle; a range costs one round trip)
    const size_t bytes = count * 3 * sizeof(double);
    if (int rc = ensure_stage(2 * bytes)) return rc;
    double* sf_ = (double*)d_stage;
    double* st_ = sf_ + count * 3;
    if (f) AGF_CUDA(cudaMemcpyAsync(sf_, f, bytes, cudaMemcpyHostToDevice, stream));
    if (t) AGF_CUDA(cudaMemcpyAsync(st_, t, bytes, cudaMemcpyHostToDevice, stream));
    wrench_set_kernel<P><<<unsigned((count * 3 + 255) / 256), 256, 0, stream>>>(d_ext_force, d_ext_torque, n, first, count,
                                                                                f ? sf_ : nullptr, t ? st_ : nullptr);
    AGF_CUDA(cudaGetLastError());
    launches++;
    AGF_CUDA(cudaStreamSynchronize(stream));  // the caller may reuse its buffers
    return AGF_OK;
  }

  int add_anchor(uint8_t id, const float pos[3]) override {
    if (!pos) return fail(AGF_EINVAL, "pos is null");
    if (sh.n_anchors >= AGF_MAX_UWB_ANCHORS) return fail(AGF_EFULL, "anchor table full (QuadcopterLogic.hpp:224-227)");
    if (id == 0) return fail(AGF_EINVAL, "anchor id 0 means 'no target' in the reference (UWBRadio.hpp:17-24)");
    if (!uwb) return fail(AGF_EUNSUPPORTED, "batch was created without a UWB network (uwb_comm_period <= 0)");
    AnchorDev& a = sh.anchors[sh.n_anchors++];
    a.id = id;
    a.x = pos[0]; a.y = pos[1]; a.z = pos[2];
    sh.tc.n_anchors = int(sh.n_anchors);
    return AGF_OK;
  }

  // ---- logging -------------------------------------------------------------------------------
  int enable_log(uint32_t stride, uint32_t cap) override {
    AGF_CUDA(cudaSetDevice(opts.device));
    AGF_CUDA(cudaStreamSynchronize(stream));
    cudaFree(d_log);
    d_log = nullptr;
    log_stride = log_cap = 0;
    log_records = 0;
    if (!stride || !cap) return AGF_OK;  // disable
    log_n = (n + 31) / 32 * 32;
    const size_t bytes = sizeof(P) * size_t(cap) * AGF_LOG_FIELDS * log_n;
    cudaError_t e = cudaMalloc(&d_log, bytes);
    if (e != cudaSuccess) {
      d_log = nullptr;
      return fail(AGF_ENOMEM, "log ring allocation", e);
    }
    log_stride = stride;
    log_cap = cap;
    log_base_tick = ticks;
    return AGF_OK;
  }

  int read_log(uint64_t rec, double* dst, size_t first, size_t count) override {
    if (!d_log) return fail(AGF_EINVAL, "logging is not enabled");
    if (first + count > n) return fail(AGF_ERANGE, "vehicle range outside the batch");
    if (rec >= log_records || rec + log_cap < log_records) return fail(AGF_ERANGE, "record not in the ring");
    AGF_CUDA(cudaSetDevice(opts.device));
    // absolute record index in the kernel's numbering
    const uint64_t abs_rec = log_base_tick / log_stride + rec;
    const size_t total = count * AGF_LOG_FIELDS;
    if (int rc = ensure_stage(total * sizeof(double))) return rc;
    const P* rec_ptr = d_log + size_t(abs_rec % log_cap) * AGF_LOG_FIELDS * log_n;
    log_gather_kernel<P><<<unsigned((total + 255) / 256), 256, 0, stream>>>(rec_ptr, log_n, first, count, (double*)d_stage);
    AGF_CUDA(cudaGetLastError());
    launches++;
    AGF_CUDA(cudaMemcpyAsync(dst, d_stage, total * sizeof(double), cudaMemcpyDeviceToHost, stream));
    AGF_CUDA(cudaStreamSynchronize(stream));
    return AGF_OK;
  }

  int log_ptr(void** p, size_t* es, size_t* ln) override {
    if (p) *p = d_log;
    if (es) *es = sizeof(P);
    if (ln) *ln = log_n;
    return d_log ? AGF_OK : fail(AGF_EINVAL, "logging is not enabled");
  }

  int stats(const double* target, double* dev_out) override {
    AGF_CUDA(cudaSetDevice(opts.device));
    const double* dt = nullptr;
    if (target) {
      if (!d_target) AGF_CUDA(cudaMalloc(&d_target, sizeof(double) * 3 * n));
      AGF_CUDA(cudaMemcpyAsync(d_target, target, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, stream));
      dt = d_target;
    }
    AGF_CUDA(cudaMemsetAsync(dev_out, 0, sizeof(double) * AGF_STATS_LEN, stream));
    stats_kernel<P><<<unsigned((n + 255) / 256), 256, 0, stream>>>((const P*)st.sp, (const float*)st.sf, (const uint32_t*)st.su, n, dt, dev_out);
    AGF_CUDA(cudaGetLastError());
    launches++;
    return AGF_OK;
  }
};

template<>
cudaError_t BatchImpl<double>::do_launch(const StepLaunch<double>& L, int block) {
  if (parity) return launch_step_parity(L, uwb, block, stream);
  return uwb ? launch_step_fast_f64_uwb(L, hk, stream) : launch_step_fast_f64_rates(L, hk, stream);
}
template<>
cudaError_t BatchImpl<float>::do_launch(const StepLaunch<float>& L, int block) {
  return uwb ? launch_step_fast_f32_uwb(L, hk, stream) : launch_step_fast_f32_rates(L, hk, stream);
}

}  // namespace agf

// =============================================================================================
// C ABI
// =============================================================================================
using agf::Batch;
using agf::fail;

static Batch* B(agf_batch* b) { return reinterpret_cast<Batch*>(b); }
static const Batch* B(const agf_batch* b) { return reinterpret_cast<const Batch*>(b); }

extern "C" {

void agf_batch_opts_default(agf_batch_opts* o) {
  if (!o) return;
  memset(o, 0, sizeof(*o));
  o->device = 0;
  o->precision = AGF_PREC_FP64;
  o->math = AGF_MATH_PARITY;
  o->block_threads = 0;
  o->onboard_logic_period = 1.0 / 500.0;
  o->uwb_comm_period = 0.0;
  o->sigma_acc = 0.2;   // Quadcopter_T.cpp:5
  o->sigma_gyro = 0.1;  // Quadcopter_T.cpp:6
  o->seed = 1;
  o->telemetry_warnings = 1;
}

int agf_batch_create(const agf_vehicle_cfg* cfgs, size_t n_cfgs, size_t n_vehicles, const agf_batch_opts* opts,
                     agf_batch** out) {
  if (!out) return fail(AGF_EINVAL, "out is null");
  *out = nullptr;
  if (!cfgs || !opts) return fail(AGF_EINVAL, "cfgs/opts is null");
  if (n_vehicles == 0) return fail(AGF_EINVAL, "n_vehicles must be > 0");
  if (n_cfgs != 1 && n_cfgs != n_vehicles) return fail(AGF_EINVAL, "n_cfgs must be 1 or n_vehicles");
  if (opts->precision != AGF_PREC_FP64 && opts->precision != AGF_PREC_FP32) return fail(AGF_EINVAL, "bad precision");
  if (opts->precision == AGF_PREC_FP32 && opts->math == AGF_MATH_PARITY)
    return fail(AGF_EUNSUPPORTED, "parity arithmetic exists for the reference's own precision (AGF_PREC_FP64) only");
  if (!(opts->onboard_logic_period > 0)) return fail(AGF_EINVAL, "onboard_logic_period must be > 0");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(AGF_ENODEVICE, "no CUDA device available: agrifly_b200 has no CPU execution path");
  }
  if (opts->device < 0 || opts->device >= ndev) return fail(AGF_EINVAL, "device ordinal out of range");
  e = cudaSetDevice(opts->device);
  if (e != cudaSuccess) return fail(AGF_ECUDA, "cudaSetDevice", e);
  Batch* b = nullptr;
  if (opts->precision == AGF_PREC_FP64) {
    b = new (std::nothrow) agf::BatchImpl<double>();
  } else {
    b = new (std::nothrow) agf::BatchImpl<float>();
  }
  if (!b) return fail(AGF_ENOMEM, "host allocation");
  b->opts = *opts;
  if (opts->stream) {
    b->stream = (cudaStream_t)opts->stream;
  } else {
    e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete b;
      return fail(AGF_ECUDA, "cudaStreamCreate", e);
    }
    b->own_stream = true;
  }
  int rc = opts->precision == AGF_PREC_FP64
               ? static_cast<agf::BatchImpl<double>*>(b)->init(cfgs, n_cfgs, n_vehicles, *opts)
               : static_cast<agf::BatchImpl<float>*>(b)->init(cfgs, n_cfgs, n_vehicles, *opts);
  if (rc) {
    delete b;
    return rc;
  }
  *out = reinterpret_cast<agf_batch*>(b);
  return AGF_OK;
}

int agf_batch_destroy(agf_batch* b) {
  if (!b) return AGF_OK;
  cudaSetDevice(B(b)->opts.device);
  cudaStreamSynchronize(B(b)->stream);
  delete B(b);
  return AGF_OK;
}

size_t agf_batch_size(const agf_batch* b) { return b ? B(b)->n : 0; }
void* agf_batch_stream(const agf_batch* b) { return b ? (void*)B(b)->stream : nullptr; }
int agf_batch_run(agf_batch* b, uint32_t dt_us, uint32_t nticks) { return b ? B(b)->run(dt_us, nticks) : fail(AGF_EINVAL, "null handle"); }
int agf_batch_advance_clock(agf_batch* b, uint32_t dt_us) { return b ? B(b)->advance_clock(dt_us) : fail(AGF_EINVAL, "null handle"); }
int agf_batch_sync(agf_batch* b) { return b ? B(b)->sync() : fail(AGF_EINVAL, "null handle"); }
uint64_t agf_batch_time_us(const agf_batch* b) { return b ? B(b)->now_us : 0; }
uint64_t agf_batch_ticks(const agf_batch* b) { return b ? B(b)->ticks : 0; }

int agf_batch_get_field(agf_batch* b, int field, void* dst, size_t first, size_t count) {
  if (!b || !dst) return fail(AGF_EINVAL, "null argument");
  return B(b)->get_field(field, dst, first, count);
}
int agf_batch_set_field(agf_batch* b, int field, const void* src, size_t first, size_t count) {
  if (!b || !src) return fail(AGF_EINVAL, "null argument");
  return B(b)->set_field(field, src, first, count);
}
size_t agf_field_size(int field) { return agf::field_ncomp(field) * agf::field_elem(field); }

int agf_batch_set_radio_cmd(agf_batch* b, const uint8_t* raw, size_t first, size_t count, int broadcast) {
  if (!b) return fail(AGF_EINVAL, "null handle");
  return B(b)->set_radio(raw, first, count, broadcast);
}
int agf_batch_set_cmd_schedule(agf_batch* b, const agf_cmd_entry* e, size_t n) {
  if (!b || (n && !e)) return fail(AGF_EINVAL, "null argument");
  return B(b)->set_schedule(e, n);
}
int agf_batch_set_cmd_slot(agf_batch* b, int slot, const uint8_t* raw) {
  if (!b) return fail(AGF_EINVAL, "null handle");
  return B(b)->set_slot(slot, raw);
}
int agf_offboard_cfg_default(int quad_type, agf_offboard_cfg* out) {
  if (!out) return fail(AGF_EINVAL, "null output");
  agf_logic_consts lc;
  const int rc = agf_logic_consts_from_type(quad_type, &lc);
  if (rc) return rc;
  memset(out, 0, sizeof(*out));
  out->period_us = 10000;  // main.cpp:175
  out->delay_us = 30000;   // main.cpp:178
  out->pos_control_nat_freq = lc.pos_control_nat_freq;  // main.cpp:227-229
  out->pos_control_damping = lc.pos_control_damping;
  out->att_control_time_const_xy = lc.att_control_time_const_xy;
  out->att_control_time_const_z = lc.att_control_time_const_z;
  out->min_vertical_proper_acc = 0.5 * 9.81;  // QuadcopterController.cpp:6-8
  out->max_proper_acc = 20;
  out->min_proper_acc = -1;
  out->yaw_angle = 0;
  return AGF_OK;
}
int agf_offboard_estimator_default(agf_offboard_estimator* o) {
  if (!o) return fail(AGF_EINVAL, "null output");
  memset(o, 0, sizeof(*o));
  o->kind = AGF_OFFEST_MOCAP;
  o->mocap_period_us = 5000;          // 1/200 s, Simulator/Rappids_Simulator/main.cpp:174
  o->prediction_delay = 0.03;         // main.cpp:179
  o->meas_reject_dist = 6.0;          // MocapStateEstimator.cpp:20-32
  o->angvel_time_const = 0.04;
  o->meas_noise_pos = 0.02;
  o->meas_noise_att = 5 * M_PI / 180;
  o->proc_noise_pos = 1.0 * 9.81;
  o->proc_noise_att = 200;
  return AGF_OK;
}
int agf_batch_set_offboard_estimator(agf_batch* b, const agf_offboard_estimator* est) {
  return b ? B(b)->set_offboard_estimator(est) : fail(AGF_EINVAL, "null handle");
}
int agf_batch_get_offboard_estimate(agf_batch* b, double horizon, double* est13, double* counters4, size_t first, size_t count) {
  return b ? B(b)->get_offboard_estimate(horizon, est13, counters4, first, count) : fail(AGF_EINVAL, "null handle");
}
int agf_batch_get_state(agf_batch* b, double* state13, size_t first, size_t count) {
  return b ? B(b)->get_state13(state13, first, count) : fail(AGF_EINVAL, "null handle");
}
int agf_batch_set_state(agf_batch* b, const double* state13, size_t first, size_t count) {
  return b ? B(b)->set_state13(state13, first, count) : fail(AGF_EINVAL, "null handle");
}
void agf_offboard_ref_safety_default(agf_offboard_ref* r) {  // SafetyNet::SafetyNet (SafetyNet.hpp:52-58)
  if (!r) return;
  r->safety_net = 1;
  r->safe_min[0] = -2.4; r->safe_min[1] = -3.1; r->safe_min[2] = -0.5;
  r->safe_max[0] = +1.8; r->safe_max[1] = +3.1; r->safe_max[2] = 4.5;
  r->min_normal_height = 1.0;
  r->not_seen_timeout = 0.5;
}
int agf_batch_set_offboard_reference(agf_batch* b, const agf_offboard_ref* ref) {
  return b ? B(b)->set_offboard_ref(ref) : fail(AGF_EINVAL, "null handle");
}
int agf_batch_set_offboard_trajectories(agf_batch* b, const double* traj, size_t first, size_t count) {
  return b ? B(b)->set_offboard_traj(traj, first, count) : fail(AGF_EINVAL, "null handle");
}
int agf_batch_offboard_trajectories_device_ptr(agf_batch* b, double** dev_ptr, size_t* n_vehicles) {
  return b ? B(b)->offboard_traj_ptr(dev_ptr, n_vehicles) : fail(AGF_EINVAL, "null handle");
}
int agf_batch_get_offboard_state(agf_batch* b, double* out, size_t first, size_t count) {
  return b ? B(b)->get_offboard_state(out, first, count) : fail(AGF_EINVAL, "null handle");
}
int agf_batch_set_offboard_loop(agf_batch* b, const agf_offboard_cfg* cfg, const agf_offboard_target* targets, size_t n_targets,
                                const double* per_vehicle_offset) {
  return b ? B(b)->set_offboard(cfg, targets, n_targets, per_vehicle_offset) : fail(AGF_EINVAL, "null handle");
}
int agf_batch_get_telemetry(agf_batch* b, uint8_t* p1, uint8_t* p2, size_t first, size_t count) {
  if (!b) return fail(AGF_EINVAL, "null handle");
  return B(b)->telemetry(p1, p2, first, count);
}
int agf_batch_set_external_wrench(agf_batch* b, const double* f, const double* t, size_t first, size_t count) {
  if (!b) return fail(AGF_EINVAL, "null handle");
  return B(b)->set_wrench(f, t, first, count);
}
int agf_batch_add_uwb_anchor(agf_batch* b, uint8_t id, const float pos[3]) {
  if (!b) return fail(AGF_EINVAL, "null handle");
  return B(b)->add_anchor(id, pos);
}
int agf_batch_set_noise(agf_batch* b, uint64_t seed, double sg, double sa, double bg, double ba) {
  if (!b) return fail(AGF_EINVAL, "null handle");
  if (B(b)->opts.precision == AGF_PREC_FP64) {
    auto* x = static_cast<agf::BatchImpl<double>*>(B(b));
    x->set_noise_params(seed, sg, sa, bg, ba, x->opts.uwb_noise_std_dev);
  } else {
    auto* x = static_cast<agf::BatchImpl<float>*>(B(b));
    x->set_noise_params(seed, sg, sa, bg, ba, x->opts.uwb_noise_std_dev);
  }
  return AGF_OK;
}
int agf_batch_set_uwb_noise(agf_batch* b, double noise, double p_out, double s_out) {
  if (!b) return fail(AGF_EINVAL, "null handle");
  if (!(noise >= 0.0) || !(p_out >= 0.0 && p_out <= 1.0) || !(s_out >= 0.0)) return fail(AGF_EINVAL, "bad UWB noise properties");
  if (B(b)->opts.precision == AGF_PREC_FP64) {
    static_cast<agf::BatchImpl<double>*>(B(b))->set_uwb_noise(noise, p_out, s_out);
  } else {
    static_cast<agf::BatchImpl<float>*>(B(b))->set_uwb_noise(noise, p_out, s_out);
  }
  return AGF_OK;
}
int agf_batch_enable_log(agf_batch* b, uint32_t stride, uint32_t cap) {
  if (!b) return fail(AGF_EINVAL, "null handle");
  return B(b)->enable_log(stride, cap);
}
uint64_t agf_batch_log_count(const agf_batch* b) { return b ? B(b)->log_records : 0; }
int agf_batch_read_log(agf_batch* b, uint64_t rec, double* dst, size_t first, size_t count) {
  if (!b || !dst) return fail(AGF_EINVAL, "null argument");
  return B(b)->read_log(rec, dst, first, count);
}
int agf_batch_log_device_ptr(agf_batch* b, void** p, size_t* es, size_t* ln) {
  if (!b) return fail(AGF_EINVAL, "null handle");
  return B(b)->log_ptr(p, es, ln);
}
int agf_batch_reduce_stats_device(agf_batch* b, const double* target, double* dev_out) {
  if (!b || !dev_out) return fail(AGF_EINVAL, "null argument");
  return B(b)->stats(target, dev_out);
}
int agf_batch_reduce_stats_nccl_device(agf_batch* b, void* nccl_comm, const double* target, double* dev_out) {
  if (!b || !dev_out) return fail(AGF_EINVAL, "null argument");
  return B(b)->stats_nccl(nccl_comm, target, dev_out);
}
int agf_batch_reduce_stats_nccl(agf_batch* b, void* nccl_comm, const double* target, double* host_out) {
  if (!b || !host_out) return fail(AGF_EINVAL, "null argument");
  Batch* x = B(b);
  cudaSetDevice(x->opts.device);
  int rc = x->ensure_stats_buffers(1);
  if (!rc) rc = x->stats_nccl(nccl_comm, target, x->d_stats_out);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(x->h_stats, x->d_stats_out, sizeof(double) * AGF_STATS_LEN, cudaMemcpyDeviceToHost, x->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(x->stream);
    if (e != cudaSuccess) return fail(AGF_ECUDA, "stats copy", e);
    memcpy(host_out, x->h_stats, sizeof(double) * AGF_STATS_LEN);
  }
  return rc;
}
int agf_batch_reduce_stats(agf_batch* b, const double* target, double* host_out) {
  return agf_batch_reduce_stats_nccl(b, nullptr, target, host_out);
}

uint64_t agf_batch_launch_count(const agf_batch* b) { return b ? B(b)->launches : 0; }

int agf_batch_step_kernel_time(agf_batch* b, double* ms, uint64_t* launches) {
  if (!b) return fail(AGF_EINVAL, "null handle");
  Batch* x = B(b);
  cudaSetDevice(x->opts.device);
  cudaError_t e = cudaStreamSynchronize(x->stream);
  if (e != cudaSuccess) return fail(AGF_ECUDA, "cudaStreamSynchronize", e);
  return x->drain_events(ms, launches);
}

const char* agf_last_error_string(void) { return agf::g_last_error.c_str(); }

const char* agf_build_info(void) {
  static char buf[4096];
  static std::once_flag once;
  std::call_once(once, [] {
    int o = snprintf(buf, sizeof(buf), "agrifly_b200 %d.%d sm_100a block=%d | ", AGF_VERSION_MAJOR, AGF_VERSION_MINOR, AGF_BLOCK_THREADS);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) {
      agf::kernel_attrs_parity(buf + o, sizeof(buf) - o);
      o = int(strlen(buf));
      agf::kernel_attrs_fast_f64_uwb(buf + o, sizeof(buf) - o);
      o = int(strlen(buf));
      agf::kernel_attrs_fast_f64_rates(buf + o, sizeof(buf) - o);
      o = int(strlen(buf));
      agf::kernel_attrs_fast_f32_uwb(buf + o, sizeof(buf) - o);
      o = int(strlen(buf));
      agf::kernel_attrs_fast_f32_rates(buf + o, sizeof(buf) - o);
    } else {
      cudaGetLastError();
      snprintf(buf + o, sizeof(buf) - o, "no CUDA device visible");
    }
  });
  return buf;
}

}  // extern "C"
