#!/bin/bash
# split planner (candidate pass + planning pass): parity tests, then occupancy variants of the planning kernel
mkdir -p gpurun_out/r2
( time timeout 600 python -m pytest tests/test_rappids_gpu.py -x -q -m gpu ) > gpurun_out/r2/gpu_tests_rappids_split.log 2>&1
tail -4 gpurun_out/r2/gpu_tests_rappids_split.log
out=gpurun_out/r2/rappids_variants_split.log
: > $out
for v in base fused mb6 mb8 mb10 mb12; do
  if [ "$v" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$v.so; fi
  echo "== $v" >> $out
  timeout 200 python profiles/prof_rappids.py fast 65536 512 4 2>&1 | grep "plans/s" >> $out
done
for v in base mb8; do
  if [ "$v" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$v.so; fi
  echo "== $v hard" >> $out
  timeout 200 python profiles/prof_rappids.py fast 65536 512 4 hard 2>&1 | grep "plans/s" >> $out
done
cat $out
