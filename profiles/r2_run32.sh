#!/bin/bash
mkdir -p gpurun_out/r2
( time timeout 900 python -m pytest tests/test_rappids_gpu.py -x -q -m gpu ) > gpurun_out/r2/gpu_tests_rappids_fold.log 2>&1
tail -12 gpurun_out/r2/gpu_tests_rappids_fold.log | head -3
out=gpurun_out/r2/rappids_fold.log
: > $out
for fam in "" hard; do
for f in 1 0; do
  echo "== shrink fold $f $fam" >> $out
  AGF_RAPPIDS_SHRINK_FOLD=$f timeout 200 python profiles/prof_rappids.py fast 65536 512 4 $fam 2>&1 | grep "plans/s" >> $out
done
done
cat $out
