// agf_launch.h -- internal: launch entry points of the step kernels.
// The parity kernels live in a translation unit compiled with -fmad=false (agf_kernels_parity.cu),
// the fast ones in agf_kernels_fast.cu, compiled once per {plant precision} x {UWB} (default contraction).
#pragma once
#include <cuda_runtime.h>

#ifndef AGF_BLOCK_THREADS
#define AGF_BLOCK_THREADS 128
#endif

#include "agf_types.h"

namespace agf {

// reference mixed precision (double plant, float logic), bit-comparable arithmetic; HK always on
cudaError_t launch_step_parity(const StepLaunch<double>& L, bool uwb, int block, cudaStream_t stream);
// FMA + CUDA fast paths; blocks of AGF_BLOCK_THREADS vehicles
cudaError_t launch_step_fast_f64_uwb(const StepLaunch<double>& L, bool hk, cudaStream_t stream);
cudaError_t launch_step_fast_f64_rates(const StepLaunch<double>& L, bool hk, cudaStream_t stream);
cudaError_t launch_step_fast_f32_uwb(const StepLaunch<float>& L, bool hk, cudaStream_t stream);
cudaError_t launch_step_fast_f32_rates(const StepLaunch<float>& L, bool hk, cudaStream_t stream);

// Offboard-loop command generation as a kernel of its own, for callers that step with the split Run()/advance form
// (agf_batch_advance_clock): one command per vehicle from the stored state into queue slot `slot`.
cudaError_t launch_offboard_generate(const StateArrays<double>& st, size_t n, const OffboardParams& off, uint64_t t_gen_us,
                                     uint32_t slot, cudaStream_t stream);
cudaError_t launch_offboard_generate(const StateArrays<float>& st, size_t n, const OffboardParams& off, uint64_t t_gen_us,
                                     uint32_t slot, cudaStream_t stream);

// the offboard estimator's mocap packet and read-out as kernels of their own (split stepping, agf_batch_get_offboard_estimate)
cudaError_t launch_offboard_mocap(const StateArrays<double>& st, size_t n, const EstParams& ep, uint64_t now_us, cudaStream_t stream);
cudaError_t launch_offboard_mocap(const StateArrays<float>& st, size_t n, const EstParams& ep, uint64_t now_us, cudaStream_t stream);
cudaError_t launch_offboard_estimate(const EstParams& ep, size_t n, size_t first, size_t count, uint64_t now_us, double horizon,
                                     double* out, cudaStream_t stream);

cudaError_t launch_offboard_counters(const double* state, size_t first, size_t count, double* out, cudaStream_t stream);

// registers per thread / local memory of each instantiation, for agf_build_info()
void kernel_attrs_parity(char* buf, size_t n);
void kernel_attrs_fast_f64_uwb(char* buf, size_t n);
void kernel_attrs_fast_f64_rates(char* buf, size_t n);
void kernel_attrs_fast_f32_uwb(char* buf, size_t n);
void kernel_attrs_fast_f32_rates(char* buf, size_t n);

}  // namespace agf
