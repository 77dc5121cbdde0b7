#!/bin/bash
mkdir -p gpurun_out/r2
( time timeout 600 python -m pytest tests/test_rappids_gpu.py -x -q -m gpu ) > gpurun_out/r2/gpu_tests_rappids_final.log 2>&1
tail -8 gpurun_out/r2/gpu_tests_rappids_final.log | head -2
export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_clk.so
timeout 300 python profiles/prof_rappids.py fast 65536 512 1 2>&1 | grep "phase cycles" | tail -5 > gpurun_out/r2/rappids_phase_clocks_final.log
cat gpurun_out/r2/rappids_phase_clocks_final.log
