#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/offb_times.log
for lib in base offb_d0; do
  if [ "$lib" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$lib.so; fi
  echo "== $lib" >> gpurun_out/offb_times.log
  for cfg in "fp32 truth targets" "fp32 mocap targets" "fp32 mocap stages"; do
    timeout 200 python profiles/prof_offboard.py $cfg 131072 500 3 >> gpurun_out/offb_times.log 2>&1
  done
done
unset AGF_LIB_PATH
timeout 200 python profiles/prof_step.py fp32 rates 131072 500 3 >> gpurun_out/offb_times.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -o gpurun_out/prof_offb_truth python profiles/prof_offboard.py fp32 truth targets 131072 200 2 > gpurun_out/prof_offb.log 2>&1
echo done
