#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/occ2.log
run() { echo "== $1 $2 $3" >> gpurun_out/occ2.log; if [ "$1" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$1.so; fi; timeout 120 python profiles/prof_step.py $2 $3 131072 500 4 >> gpurun_out/occ2.log 2>&1; }
run base fp32 uwb; run u4 fp32 uwb; run base fp32 rates; run r5 fp32 rates; run r3 fp32 rates; run base fp64 uwb; run d3 fp64 uwb
echo done
