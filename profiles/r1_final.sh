#!/bin/bash
# end-of-round evidence: GPU suite, both bench arms as the driver runs them, smoke, launch list of the bench command,
# DRAM traffic of one bench launch
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_h.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests_h.log
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/bench_ref_h.json 2> gpurun_out/bench_ref_h.err
( time timeout 900 python bench.py ) > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err
python __graft_entry__.py smoke > gpurun_out/smoke_h.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_h.log
AGF_NO_WARM=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_h.csv python bench.py --steps 3 --warmup 1 --no-extras > gpurun_out/launches_h.log 2>&1
AGF_NO_WARM=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:step_kernel -s 1 -c 1 --csv --log-file gpurun_out/traffic_fp32_uwb_131072_500_h.csv python profiles/prof_step.py fp32 uwb 131072 500 2 > gpurun_out/traffic_h.log 2>&1
echo done
