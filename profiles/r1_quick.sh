#!/bin/bash
# quick check: variants + the fast-variant tolerance / noise / balanced tests
mkdir -p gpurun_out
bash profiles/r1_variants.sh "$@"
for m in "fp32 rates" "fp64 uwb" "fp64 rates"; do echo "== base $m" >> gpurun_out/variants.log; timeout 120 python profiles/prof_step.py $m 131072 500 3 >> gpurun_out/variants.log 2>&1; done
timeout 900 python -m pytest tests -m gpu -x -q -s -k "fast or noise or balanced or full_size or bias" > gpurun_out/gpu_tests_quick.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests_quick.log
