// agf_kernels_fast_f64.cu -- step kernels, double plant, FMA contraction and CUDA libm.
#include <stdio.h>

#include "agf_launch.h"
#include "agf_step.cuh"

namespace agf {

cudaError_t launch_step_fast_f64(const StepLaunch<double>& L, bool uwb, bool hk, int block, cudaStream_t stream) {
  const unsigned grid = unsigned((L.n + block - 1) / block);
  static bool carveout_set = false;
  if (!carveout_set) {  // the scratch of 4 resident blocks needs most of the SM's shared memory
    cudaFuncSetAttribute(step_kernel<double, false, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(step_kernel<double, false, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    carveout_set = true;
  }
  if (uwb) {
    if (hk) step_kernel<double, false, true, true><<<grid, block, step_smem_bytes<false, true>(block), stream>>>(L);
    else step_kernel<double, false, true, false><<<grid, block, step_smem_bytes<false, true>(block), stream>>>(L);
  } else {
    if (hk) step_kernel<double, false, false, true><<<grid, block, step_smem_bytes<false, false>(block), stream>>>(L);
    else step_kernel<double, false, false, false><<<grid, block, step_smem_bytes<false, false>(block), stream>>>(L);
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

void kernel_attrs_fast_f64(char* buf, size_t n) {
  int o = 0;
  o += attr_line(buf + o, n - o, "step<f64,fast,uwb,hk>", step_kernel<double, false, true, true>);
  if (size_t(o) < n) o += attr_line(buf + o, n - o, "step<f64,fast,uwb>", step_kernel<double, false, true, false>);
  if (size_t(o) < n) o += attr_line(buf + o, n - o, "step<f64,fast,nouwb,hk>", step_kernel<double, false, false, true>);
  if (size_t(o) < n) o += attr_line(buf + o, n - o, "step<f64,fast,nouwb>", step_kernel<double, false, false, false>);
}

}  // namespace agf
