#!/bin/bash
# ncu --set full of the rates kernel with the in-kernel offboard loop + MocapStateEstimator
mkdir -p gpurun_out/r2
O=gpurun_out/r2
tag=${1:-i}
timeout 300 python profiles/prof_offboard.py fp32 mocap targets 131072 500 3 > $O/mocap_times_$tag.log 2>&1
timeout 300 python profiles/prof_offboard.py fp32 truth targets 131072 500 3 >> $O/mocap_times_$tag.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -o $O/prof_mocap_$tag -f python profiles/prof_offboard.py fp32 mocap targets 131072 200 2 > $O/prof_mocap_$tag.log 2>&1
cp agri-fly_b200/build/agf_kernels_fast_f32_rates.o $O/agf_kernels_fast_f32_rates_$tag.o
cat $O/mocap_times_$tag.log
