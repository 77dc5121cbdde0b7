#!/bin/bash
# round 2, GPU call 9: estimator state blocked by warp + unordered pipe slots: offboard tests and timings
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_host_facade.py -m gpu -q -x -k "offboard or fleet or object_api" ) > $O/gpu_tests_offb_j.log 2>&1; tail -3 $O/gpu_tests_offb_j.log
timeout 300 python profiles/prof_offboard.py fp32 mocap targets 131072 500 3 > $O/mocap_times_j.log 2>&1
timeout 300 python profiles/prof_offboard.py fp32 truth targets 131072 500 3 >> $O/mocap_times_j.log 2>&1
timeout 300 python profiles/prof_offboard.py fp32 mocap stages 131072 500 3 >> $O/mocap_times_j.log 2>&1
timeout 300 python profiles/prof_offboard.py fp64 mocap targets 131072 500 3 >> $O/mocap_times_j.log 2>&1
cat $O/mocap_times_j.log
