#!/bin/bash
mkdir -p gpurun_out/r2
( time timeout 600 python -m pytest tests/test_rappids_gpu.py -x -q -m gpu ) > gpurun_out/r2/gpu_tests_rappids_final.log 2>&1
tail -8 gpurun_out/r2/gpu_tests_rappids_final.log | head -2
out=gpurun_out/r2/rappids_times_final.log
: > $out
timeout 120 python profiles/prof_rappids.py fast 65536 512 4 2>&1 | grep "plans/s" >> $out
timeout 120 python profiles/prof_rappids.py fast 65536 512 4 hard 2>&1 | grep "plans/s" >> $out
timeout 120 python profiles/prof_rappids.py parity 65536 512 3 2>&1 | grep "plans/s" >> $out
cat $out
