#!/bin/bash
# variants + full ncu capture of the FP32 full-mode step kernel + GPU tests
mkdir -p gpurun_out
bash profiles/r1_variants.sh "$@"
for m in "fp32 rates" "fp64 uwb" "fp64 rates"; do echo "== base $m" >> gpurun_out/variants.log; timeout 120 python profiles/prof_step.py $m 131072 500 3 >> gpurun_out/variants.log 2>&1; done
AGF_NO_WARM=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -o gpurun_out/prof_f32_uwb_c python profiles/prof_step.py fp32 uwb 131072 200 2 > gpurun_out/prof_full_c.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests.log
echo done
