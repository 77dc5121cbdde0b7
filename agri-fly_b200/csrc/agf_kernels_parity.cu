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
