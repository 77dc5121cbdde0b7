#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_k.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests_k.log
: > gpurun_out/bal.log
for m in "fp64 rates 65536" "fp64 uwb 65536" "fp32 uwb 65536" "fp32 rates 50000" "fp32 uwb 40000"; do echo "== $m" >> gpurun_out/bal.log; timeout 120 python profiles/prof_step.py $m 500 4 >> gpurun_out/bal.log 2>&1; done
echo done
