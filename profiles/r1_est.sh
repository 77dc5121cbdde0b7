#!/bin/bash
# full GPU suite incl. the offboard estimator + a default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_f.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests_f.log
( time timeout 600 python bench.py ) > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
echo done
