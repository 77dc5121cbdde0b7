// agf_kernels_parity.cu -- step kernels with bit-comparable arithmetic.
// MUST be compiled with -fmad=false (no FMA contraction), default -prec-div/-prec-sqrt (IEEE).
#include <stdio.h>

#include "agf_launch.h"
#include "agf_step.cuh"

namespace agf {

cudaError_t launch_step_parity(const StepLaunch<double>& L, bool uwb, int block, cudaStream_t stream) {
  const unsigned grid = unsigned((L.n + block - 1) / block);
  if (uwb) {
    step_kernel<double, true, true, true><<<grid, block, 0, stream>>>(L);
  } else {
    step_kernel<double, true, false, true><<<grid, block, 0, stream>>>(L);
  }
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
  int o = attr_line(buf, n, "step<f64,parity,uwb>", step_kernel<double, true, true, true>);
  if (o > 0 && size_t(o) < n) attr_line(buf + o, n - o, "step<f64,parity,nouwb>", step_kernel<double, true, false, true>);
}

}  // namespace agf
