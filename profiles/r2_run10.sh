#!/bin/bash
# round 2, GPU call 10: packed FP32 (FFMA2) EKF predict/update: population parity of the fast kernels, timings, whole suite
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( timeout 1200 python -m pytest tests/test_fast_population_gpu.py -m gpu -q -s ) > $O/gpu_tests_fastpop_k.log 2>&1; tail -2 $O/gpu_tests_fastpop_k.log
bash profiles/r2_variants.sh k base
( timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_fast_population_gpu.py ) > $O/gpu_tests_k.log 2>&1; tail -3 $O/gpu_tests_k.log
