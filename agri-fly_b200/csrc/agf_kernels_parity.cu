// agf_kernels_parity.cu -- step kernels with bit-comparable arithmetic.
// MUST be compiled with -fmad=false (no FMA contraction), default -prec-div/-prec-sqrt (IEEE).
#include <stdio.h>

#include "agf_launch.h"
#include "agf_step.cuh"

namespace agf {

cudaError_t launch_step_parity(const StepLaunch<double>& L, bool uwb, int block, cudaStream_t stream) {
  if (uwb) return launch_step_kernel<double>(step_kernel<double, true, true, true, false>, L, block, 0, stream);
  return launch_step_kernel<double>(step_kernel<double, true, false, true, false>, L, block, 0, stream);
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
  int o = attr_line(buf, n, "step<f64,parity,uwb>", step_kernel<double, true, true, true, false>);
  if (o > 0 && size_t(o) < n) attr_line(buf + o, n - o, "step<f64,parity,nouwb>", step_kernel<double, true, false, true, false>);
}

}  // namespace agf
