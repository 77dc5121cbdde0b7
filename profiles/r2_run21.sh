#!/bin/bash
# planning pass: dispatch by the previous plan's work against index order
mkdir -p gpurun_out/r2
( time timeout 600 python -m pytest tests/test_rappids_gpu.py -x -q -m gpu ) > gpurun_out/r2/gpu_tests_rappids_dispatch.log 2>&1
tail -6 gpurun_out/r2/gpu_tests_rappids_dispatch.log | head -2
out=gpurun_out/r2/rappids_dispatch.log
: > $out
for fam in "" hard; do
for d in work index; do
  echo "== dispatch $d $fam" >> $out
  AGF_PROF_DISPATCH=$d timeout 200 python profiles/prof_rappids.py fast 65536 512 4 $fam 2>&1 | grep "plans/s" >> $out
done
done
echo "== dispatch work, 16384 plans" >> $out
timeout 200 python profiles/prof_rappids.py fast 16384 512 4 2>&1 | grep "plans/s" >> $out
cat $out
