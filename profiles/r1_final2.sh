#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/gpu_tests_i.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests_i.log
echo done
