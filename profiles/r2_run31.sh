#!/bin/bash
mkdir -p gpurun_out/r2
( time timeout 900 python -m pytest tests/test_rappids_gpu.py -x -q -m gpu ) > gpurun_out/r2/gpu_tests_rappids_jump.log 2>&1
tail -12 gpurun_out/r2/gpu_tests_rappids_jump.log | head -3
out=gpurun_out/r2/rappids_jump_final.log
: > $out
for fam in "" hard; do
for j in 8 0; do
  echo "== frame jump $j $fam" >> $out
  AGF_RAPPIDS_FRAME_JUMP=$j timeout 200 python profiles/prof_rappids.py fast 65536 512 4 $fam 2>&1 | grep "plans/s" >> $out
done
done
echo "== parity variant, frame jump 8" >> $out
timeout 200 python profiles/prof_rappids.py parity 65536 512 3 2>&1 | grep "plans/s" >> $out
echo "== index order, frame jump 8" >> $out
AGF_PROF_DISPATCH=index timeout 200 python profiles/prof_rappids.py fast 65536 512 3 2>&1 | grep "plans/s" >> $out
cat $out
