#!/bin/bash
mkdir -p gpurun_out/r2
( time timeout 900 python -m pytest tests/test_rappids_gpu.py -x -q -m gpu ) > gpurun_out/r2/gpu_tests_rappids_divcalls.log 2>&1
tail -12 gpurun_out/r2/gpu_tests_rappids_divcalls.log | head -3
out=gpurun_out/r2/rappids_variants_divcalls.log
: > $out
for fam in "" hard; do
for v in "$@"; do
  if [ "$v" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$v.so; fi
  echo "== $v $fam" >> $out
  timeout 200 python profiles/prof_rappids.py fast 65536 512 4 $fam 2>&1 | grep "plans/s" >> $out
done
done
cat $out
