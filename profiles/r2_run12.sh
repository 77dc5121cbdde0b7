#!/bin/bash
# round 2, GPU call 12: whole GPU suite (reference build as checker, planner ground-truth test)
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( time timeout 2400 python -m pytest tests -m gpu -q -s --durations=8 ) > $O/gpu_tests_m.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests_m.log
grep -v "^envelope\|^fp32\|^fp64" $O/gpu_tests_m.log | tail -25
