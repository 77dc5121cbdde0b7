#!/bin/bash
# round 2, GPU call 13: fast-variant estimator math; outlier test against the per-vehicle-seeded reference population
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_host_facade.py -m gpu -q -x -s -k "offboard or fleet or object_api or outlier or uwb" ) > $O/gpu_tests_offb_n.log 2>&1; tail -4 $O/gpu_tests_offb_n.log; grep "rejected per" $O/gpu_tests_offb_n.log
timeout 300 python profiles/prof_offboard.py fp32 mocap targets 131072 500 3 > $O/mocap_times_n.log 2>&1
timeout 300 python profiles/prof_offboard.py fp32 truth targets 131072 500 3 >> $O/mocap_times_n.log 2>&1
timeout 300 python profiles/prof_offboard.py fp32 mocap stages 131072 500 3 >> $O/mocap_times_n.log 2>&1
timeout 300 python profiles/prof_offboard.py fp64 mocap targets 131072 500 3 >> $O/mocap_times_n.log 2>&1
cat $O/mocap_times_n.log
