#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/occ3.log
run() { echo "== $1 $2 $3" >> gpurun_out/occ3.log; if [ "$1" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$1.so; fi; timeout 120 python profiles/prof_step.py $2 $3 131072 500 4 >> gpurun_out/occ3.log 2>&1; }
run base fp64 uwb; run d4 fp64 uwb; run base fp64 rates; run dr4 fp64 rates
unset AGF_LIB_PATH
for cfg in "fp32 truth targets" "fp32 mocap targets" "fp64 mocap targets"; do timeout 200 python profiles/prof_offboard.py $cfg 131072 500 3 >> gpurun_out/occ3.log 2>&1; done
echo done
