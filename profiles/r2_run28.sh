#!/bin/bash
mkdir -p gpurun_out/r2
( time timeout 900 python -m pytest tests/test_rappids_gpu.py -x -q -m gpu ) > gpurun_out/r2/gpu_tests_rappids_jump.log 2>&1
tail -12 gpurun_out/r2/gpu_tests_rappids_jump.log | head -8
out=gpurun_out/r2/rappids_jump.log
: > $out
for j in 0 8; do
  echo "== frame jump $j" >> $out
  AGF_RAPPIDS_FRAME_JUMP=$j timeout 200 python profiles/prof_rappids.py fast 65536 512 4 2>&1 | grep "plans/s" >> $out
done
for j in 0 8; do
  echo "== frame jump $j hard" >> $out
  AGF_RAPPIDS_FRAME_JUMP=$j timeout 200 python profiles/prof_rappids.py fast 65536 512 4 hard 2>&1 | grep "plans/s" >> $out
done
export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_clk.so
echo "== phase clocks, frame jump 8" >> $out
timeout 300 python profiles/prof_rappids.py fast 65536 512 1 2>&1 | grep "phase cycles" | tail -4 >> $out
cat $out
