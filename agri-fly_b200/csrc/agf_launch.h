// agf_launch.h -- internal: launch entry points of the step kernels.
// The parity kernels live in a translation unit compiled with -fmad=false (agf_kernels_parity.cu),
// the fast ones in agf_kernels_fast_{f32,f64}.cu (default contraction).
#pragma once
#include <cuda_runtime.h>

#include "agf_types.h"

#ifndef AGF_BLOCK_THREADS
#define AGF_BLOCK_THREADS 128
#endif

namespace agf {

// reference mixed precision (double plant, float logic), bit-comparable arithmetic; HK always on
cudaError_t launch_step_parity(const StepLaunch<double>& L, bool uwb, int block, cudaStream_t stream);
// double plant, FMA + CUDA libm
cudaError_t launch_step_fast_f64(const StepLaunch<double>& L, bool uwb, bool hk, int block, cudaStream_t stream);
// float plant, FMA + CUDA libm
cudaError_t launch_step_fast_f32(const StepLaunch<float>& L, bool uwb, bool hk, int block, cudaStream_t stream);

// registers per thread / static smem of each instantiation, for agf_build_info()
void kernel_attrs_parity(char* buf, size_t n);
void kernel_attrs_fast_f64(char* buf, size_t n);
void kernel_attrs_fast_f32(char* buf, size_t n);

}  // namespace agf
