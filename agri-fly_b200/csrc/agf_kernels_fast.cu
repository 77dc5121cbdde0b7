// agf_kernels_fast.cu -- step kernels with FMA contraction and CUDA fast paths.
// Compiled four times (build.py): -DAGF_FAST_F64=0|1 (plant precision) x -DAGF_FAST_UWB=0|1 (EKF + ranging),
// each object holding the {housekeeping} x {per-vehicle parameters} instantiations of one kernel family.
#include <stdio.h>

#include "agf_launch.h"
#include "agf_step.cuh"

#ifndef AGF_FAST_F64
#error "define AGF_FAST_F64 to 0 or 1"
#endif
#ifndef AGF_FAST_UWB
#error "define AGF_FAST_UWB to 0 or 1"
#endif

namespace agf {

#if AGF_FAST_F64
typedef double FastP;
#define AGF_FAST_NAME(a, b) a##f64##b
#else
typedef float FastP;
#define AGF_FAST_NAME(a, b) a##f32##b
#endif
#if AGF_FAST_UWB
#define AGF_FAST_FN(stem) AGF_FAST_NAME(stem, _uwb)
static constexpr bool kUwb = true;
#else
#define AGF_FAST_FN(stem) AGF_FAST_NAME(stem, _rates)
static constexpr bool kUwb = false;
#endif

template<bool HK, bool PV, bool OFFB>
static cudaError_t go(const StepLaunch<FastP>& L, cudaStream_t stream) {
  auto kernel = step_kernel<FastP, false, kUwb, HK, PV, OFFB>;
  const size_t smem = step_smem_bytes<false, kUwb>(AGF_BLOCK_THREADS, L.st.sq != nullptr);
  static bool carveout_set = false;
  if (!carveout_set) {
    // shared-memory carve-out = what the resident blocks' scratch needs (+1 KB per block the runtime reserves), not the
    // maximum: the rest of the 228 KB stays L1, which holds the tick plans, the schedule and the few spilled words of
    // the loop -- with the trajectory log streaming through it, a minimal L1 misses on those (profiles/r2 C4 captures)
#if AGF_CARVEOUT_MAX
    int pct = cudaSharedmemCarveoutMaxShared;
#else
    int dev = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    const size_t need = size_t(step_min_blocks<FastP, false, kUwb, OFFB>()) * (smem + 1024);
    int pct = per_sm > 0 ? int((need * 100 + size_t(per_sm) - 1) / size_t(per_sm)) : 100;
    if (pct > 100) pct = 100;
#endif
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    carveout_set = true;
  }
  return launch_step_kernel<FastP>(kernel, L, AGF_BLOCK_THREADS, smem, stream);
}

// launch_step_fast_{f32,f64}_{uwb,rates}
cudaError_t AGF_FAST_FN(launch_step_fast_)(const StepLaunch<FastP>& L, bool hk, cudaStream_t stream) {
  const bool pv = L.pv != nullptr;
  if (L.sh.tc.off_enabled) {  // kernels with the in-kernel offboard loop compiled in
    if (hk) return pv ? go<true, true, true>(L, stream) : go<true, false, true>(L, stream);
    return pv ? go<false, true, true>(L, stream) : go<false, false, true>(L, stream);
  }
  if (hk) return pv ? go<true, true, false>(L, stream) : go<true, false, false>(L, stream);
  return pv ? go<false, true, false>(L, stream) : go<false, false, false>(L, stream);
}

template<typename K>
static int attr_line(char* buf, size_t n, const char* name, K kernel) {
  cudaFuncAttributes a;
  if (cudaFuncGetAttributes(&a, kernel) != cudaSuccess) {
    cudaGetLastError();
    return snprintf(buf, n, "%s: n/a; ", name);
  }
  return snprintf(buf, n, "%s: %d regs, %zu B local; ", name, a.numRegs, a.localSizeBytes);
}

#define AGF_STR2(x) #x
#define AGF_STR(x) AGF_STR2(x)
void AGF_FAST_FN(kernel_attrs_fast_)(char* buf, size_t n) {
  const char* fam = AGF_STR(AGF_FAST_FN(step_fast_));
  char nm[96];
  int o = 0;
  snprintf(nm, sizeof nm, "%s", fam);
  o += attr_line(buf + o, n - o, nm, step_kernel<FastP, false, kUwb, false, false, false>);
  snprintf(nm, sizeof nm, "%s+hk", fam);
  if (size_t(o) < n) o += attr_line(buf + o, n - o, nm, step_kernel<FastP, false, kUwb, true, false, false>);
  snprintf(nm, sizeof nm, "%s+pv", fam);
  if (size_t(o) < n) o += attr_line(buf + o, n - o, nm, step_kernel<FastP, false, kUwb, false, true, false>);
  snprintf(nm, sizeof nm, "%s+offboard", fam);
  if (size_t(o) < n) o += attr_line(buf + o, n - o, nm, step_kernel<FastP, false, kUwb, false, false, true>);
  snprintf(nm, sizeof nm, "%s+hk+pv", fam);
  if (size_t(o) < n) o += attr_line(buf + o, n - o, nm, step_kernel<FastP, false, kUwb, true, true, false>);
}

}  // namespace agf
