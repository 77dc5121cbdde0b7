#!/bin/bash
# planning pass with frame jumps: parity tests, then timings at jump lengths 0 / 4 / 8 / 16
mkdir -p gpurun_out/r2
( time timeout 900 python -m pytest tests/test_rappids_gpu.py -x -q -m gpu ) > gpurun_out/r2/gpu_tests_rappids_jump.log 2>&1
tail -12 gpurun_out/r2/gpu_tests_rappids_jump.log | head -8
out=gpurun_out/r2/rappids_jump.log
: > $out
for j in 0 4 8 16; do
  echo "== frame jump $j" >> $out
  AGF_RAPPIDS_FRAME_JUMP=$j timeout 200 python profiles/prof_rappids.py fast 65536 512 4 2>&1 | grep "plans/s" >> $out
done
for j in 0 8; do
  echo "== frame jump $j hard" >> $out
  AGF_RAPPIDS_FRAME_JUMP=$j timeout 200 python profiles/prof_rappids.py fast 65536 512 4 hard 2>&1 | grep "plans/s" >> $out
done
cat $out
