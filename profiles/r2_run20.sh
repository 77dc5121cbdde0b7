#!/bin/bash
# planning pass: end-group pixels requested with the group minima (spec), L1 prefetch of the side's next line (pf)
mkdir -p gpurun_out/r2
( time timeout 600 python -m pytest tests/test_rappids_gpu.py -x -q -m gpu ) > gpurun_out/r2/gpu_tests_rappids_spec.log 2>&1
tail -6 gpurun_out/r2/gpu_tests_rappids_spec.log | head -2
out=gpurun_out/r2/rappids_variants_spec.log
: > $out
for v in "$@"; do
  if [ "$v" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$v.so; fi
  echo "== $v" >> $out
  timeout 200 python profiles/prof_rappids.py fast 65536 512 4 2>&1 | grep "plans/s" >> $out
done
cat $out
