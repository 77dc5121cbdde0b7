// agf_rappids_plan.cu -- instantiation of the batched RAPPIDS planner kernel (agf_rappids_plan.cuh).
// Compiled twice by build.py:
//   -DAGF_RAPPIDS_PARITY=1 -fmad=false   bit-comparable arithmetic (agf_math.h), the parity variant
//   -DAGF_RAPPIDS_PARITY=0               FMA contraction + CUDA libm, the throughput variant
#include "agf_rappids_plan.cuh"

#ifndef AGF_RAPPIDS_PARITY
#error "define AGF_RAPPIDS_PARITY to 0 or 1"
#endif

namespace agfr {

#if AGF_RAPPIDS_PARITY
#define AGFR_LAUNCH launch_plan_parity
#define AGFR_OCC plan_blocks_per_sm_parity
#else
#define AGFR_LAUNCH launch_plan_fast
#define AGFR_OCC plan_blocks_per_sm_fast
#endif

cudaError_t AGFR_LAUNCH(const PlanParams& P, int grid, cudaStream_t stream) {
  // candidate pass: one thread per candidate
  const size_t total = (size_t)P.n * (size_t)P.k;
  const size_t cap = (size_t)148 * 64;
  const int cgrid = (int)((total + 127) / 128 < cap ? (total + 127) / 128 : cap);
  rappids_candidates_kernel<AGF_RAPPIDS_PARITY != 0><<<cgrid, 128, 0, stream>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // planning pass: one warp per vehicle, one resident wave
  rappids_plan_kernel<AGF_RAPPIDS_PARITY != 0><<<grid, kBlock, 0, stream>>>(P);
  return cudaGetLastError();
}

cudaError_t AGFR_OCC(int* blocks, int* regs) {
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, rappids_plan_kernel<AGF_RAPPIDS_PARITY != 0>);
  if (e != cudaSuccess) return e;
  *regs = a.numRegs;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, rappids_plan_kernel<AGF_RAPPIDS_PARITY != 0>, kBlock, 0);
}

}  // namespace agfr
