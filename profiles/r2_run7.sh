#!/bin/bash
# round 2, GPU call 7: whole GPU suite (no -x) after the outlier-test fix
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( time timeout 2400 python -m pytest tests -m gpu -q --durations=8 ) > $O/gpu_tests_f.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests_f.log
tail -5 $O/gpu_tests_f.log
