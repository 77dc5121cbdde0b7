#!/bin/bash
# after the tick-plan table refactor: GPU suite, kernel-only throughput of all modes, offboard-loop modes, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_j.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests_j.log
: > gpurun_out/plans_times.log
for m in "fp32 uwb" "fp32 rates" "fp64 uwb" "fp64 rates"; do echo "== $m" >> gpurun_out/plans_times.log; timeout 120 python profiles/prof_step.py $m 131072 500 4 >> gpurun_out/plans_times.log 2>&1; done
for cfg in "fp32 truth targets" "fp32 mocap targets" "fp32 mocap stages"; do
  timeout 200 python profiles/prof_offboard.py $cfg 131072 500 3 >> gpurun_out/plans_times.log 2>&1
done
( time timeout 900 python bench.py ) > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err
echo done
