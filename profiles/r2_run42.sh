#!/bin/bash
mkdir -p gpurun_out/r2
( time timeout 600 python -m pytest tests/test_rappids_gpu.py -x -q -m gpu -k "frame_jumps" -s ) > gpurun_out/r2/gpu_tests_rappids_coop.log 2>&1
tail -15 gpurun_out/r2/gpu_tests_rappids_coop.log | head -3
out=gpurun_out/r2/rappids_coop_cap.log
: > $out
for fam in "" hard; do
for c in 0.04 0.1; do
  echo "== coop factor 0.3 cap $c $fam" >> $out
  AGF_RAPPIDS_COOP_FACTOR=0.3 AGF_RAPPIDS_COOP_CAP=$c timeout 120 python profiles/prof_rappids.py fast 65536 512 4 $fam 2>&1 | grep "plans/s" >> $out
done
done
cat $out
