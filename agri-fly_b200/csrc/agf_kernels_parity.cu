// agf_kernels_parity.cu -- step kernels with bit-comparable arithmetic.
// MUST be compiled with -fmad=false (no FMA contraction), default -prec-div/-prec-sqrt (IEEE).
#include <stdio.h>

#define AGF_MATH_OUTLINE 1  // agf_math.h: the shared libm as real calls (code size; same arithmetic)
#include "agf_launch.h"
#include "agf_step.cuh"

namespace agf {

// The in-kernel offboard loop is a template axis here as in the fast kernels: batches without it run kernels that do
// not contain it (half the code).
cudaError_t launch_step_parity(const StepLaunch<double>& L, bool uwb, int block, cudaStream_t stream) {
  if (L.sh.tc.off_enabled) {
    if (uwb) return launch_step_kernel<double>(step_kernel<double, true, true, true, false, true>, L, block, 0, stream);
    return launch_step_kernel<double>(step_kernel<double, true, false, true, false, true>, L, block, 0, stream);
  }
  if (uwb) return launch_step_kernel<double>(step_kernel<double, true, true, true, false, false>, L, block, 0, stream);
  return launch_step_kernel<double>(step_kernel<double, true, false, true, false, false>, L, block, 0, stream);
}

// offboard main loop outside the step kernel (split Run()/advance stepping): same arithmetic as tick()'s
template<typename P>
__global__ void offboard_generate_kernel(StateArrays<P> st, size_t n, OffboardParams off, uint64_t t_gen_us, uint32_t slot) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int VP = VecOf<P>::lanes;
  P r[12];
#pragma unroll
  for (int q = 0; q < 12 / VP; q++) VecOf<P>::unpack(st.sp[size_t(q) * n + i], &r[q * VP]);  // pos3 vel3 att4 (+2)
  const V3<P> cp(r[SP_POS], r[SP_POS + 1], r[SP_POS + 2]), cv(r[SP_VEL], r[SP_VEL + 1], r[SP_VEL + 2]);
  const Q4<P> ca(r[SP_ATT], r[SP_ATT + 1], r[SP_ATT + 2], r[SP_ATT + 3]);
  st.sq[size_t(slot) * n + i] = offboard_generate<true, P>(off, i, n, t_gen_us, cp, cv, ca);
}
// simulated mocap packet of the offboard estimator for the split Run()/advance stepping: same code as tick()'s
template<typename P>
__global__ void offboard_mocap_kernel(StateArrays<P> st, size_t n, EstParams ep, uint64_t now_us) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int VP = VecOf<P>::lanes;
  P r[12];
#pragma unroll
  for (int q = 0; q < 12 / VP; q++) VecOf<P>::unpack(st.sp[size_t(q) * n + i], &r[q * VP]);
  mocap_update<true, double>(ep, i, now_us, V3<double>(double(r[SP_POS]), double(r[SP_POS + 1]), double(r[SP_POS + 2])),
                     Q4<double>(double(r[SP_ATT]), double(r[SP_ATT + 1]), double(r[SP_ATT + 2]), double(r[SP_ATT + 3])));
}
cudaError_t launch_offboard_mocap(const StateArrays<double>& st, size_t n, const EstParams& ep, uint64_t now_us, cudaStream_t stream) {
  offboard_mocap_kernel<double><<<unsigned((n + 127) / 128), 128, 0, stream>>>(st, n, ep, now_us);
  return cudaGetLastError();
}
cudaError_t launch_offboard_mocap(const StateArrays<float>& st, size_t n, const EstParams& ep, uint64_t now_us, cudaStream_t stream) {
  offboard_mocap_kernel<float><<<unsigned((n + 127) / 128), 128, 0, stream>>>(st, n, ep, now_us);
  return cudaGetLastError();
}
// initialised flag, rejection counters and message count of vehicles first .. first+count-1 -> out [count][4]
__global__ void offboard_counters_kernel(const double* state, size_t first, size_t count, double* out) {
  const size_t k = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const double* st = state + est_index(first + k);
  out[4 * k + 0] = st[size_t(E_INIT) * E_LANES];
  out[4 * k + 1] = st[size_t(E_NREJ) * E_LANES];
  out[4 * k + 2] = st[size_t(E_NREJC) * E_LANES];
  out[4 * k + 3] = st[size_t(E_NPIPE) * E_LANES];
}
cudaError_t launch_offboard_counters(const double* state, size_t first, size_t count, double* out, cudaStream_t stream) {
  offboard_counters_kernel<<<unsigned((count + 127) / 128), 128, 0, stream>>>(state, first, count, out);
  return cudaGetLastError();
}
// MocapStateEstimator::GetPrediction(horizon) for vehicles first .. first+count-1 -> out [13][count]
__global__ void offboard_estimate_kernel(EstParams ep, size_t n, size_t first, size_t count, uint64_t now_us, double horizon, double* out) {
  const size_t k = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= count) return;
  EstCore e;
  EstPipe pipe;
  mocap_predict<true, double>(ep, first + k, now_us, horizon, e, pipe);
  const double v[13] = {e.pos.x, e.pos.y, e.pos.z, e.vel.x, e.vel.y, e.vel.z, e.att.w, e.att.x, e.att.y, e.att.z, e.w.x, e.w.y, e.w.z};
  for (int f = 0; f < 13; f++) out[size_t(f) * count + k] = v[f];
}
cudaError_t launch_offboard_estimate(const EstParams& ep, size_t n, size_t first, size_t count, uint64_t now_us, double horizon,
                                     double* out, cudaStream_t stream) {
  offboard_estimate_kernel<<<unsigned((count + 127) / 128), 128, 0, stream>>>(ep, n, first, count, now_us, horizon, out);
  return cudaGetLastError();
}
cudaError_t launch_offboard_generate(const StateArrays<double>& st, size_t n, const OffboardParams& off, uint64_t t_gen_us,
                                     uint32_t slot, cudaStream_t stream) {
  offboard_generate_kernel<double><<<unsigned((n + 127) / 128), 128, 0, stream>>>(st, n, off, t_gen_us, slot);
  return cudaGetLastError();
}
cudaError_t launch_offboard_generate(const StateArrays<float>& st, size_t n, const OffboardParams& off, uint64_t t_gen_us,
                                     uint32_t slot, cudaStream_t stream) {
  offboard_generate_kernel<float><<<unsigned((n + 127) / 128), 128, 0, stream>>>(st, n, off, t_gen_us, slot);
  return cudaGetLastError();
}

template<typename K>
static int attr_line(char* buf, size_t n, const char* name, K kernel) {
  cudaFuncAttributes a;
  if (cudaFuncGetAttributes(&a, kernel) != cudaSuccess) {
    cudaGetLastError();
    return snprintf(buf, n, "%s: n/a; ", name);
  }
  return snprintf(buf, n, "%s: %d regs, %zu B local, %zu B smem; ", name, a.numRegs, a.localSizeBytes,
                  a.sharedSizeBytes);
}

void kernel_attrs_parity(char* buf, size_t n) {
  int o = attr_line(buf, n, "step<f64,parity,uwb>", step_kernel<double, true, true, true, false, false>);
  if (o > 0 && size_t(o) < n) o += attr_line(buf + o, n - o, "step<f64,parity,nouwb>", step_kernel<double, true, false, true, false, false>);
  if (o > 0 && size_t(o) < n) o += attr_line(buf + o, n - o, "step<f64,parity,uwb>+offboard", step_kernel<double, true, true, true, false, true>);
  if (o > 0 && size_t(o) < n) attr_line(buf + o, n - o, "step<f64,parity,nouwb>+offboard", step_kernel<double, true, false, true, false, true>);
}

}  // namespace agf
