#!/bin/bash
# round 2, GPU call: planner tests + timings after a planner change
mkdir -p gpurun_out/r2
O=gpurun_out/r2
tag=${1:-q}
( timeout 900 python -m pytest tests/test_rappids_gpu.py -m gpu -q -x ) > $O/gpu_tests_rappids_$tag.log 2>&1; tail -3 $O/gpu_tests_rappids_$tag.log
timeout 300 python profiles/prof_rappids.py fast 65536 512 4 > $O/rappids_times_$tag.log 2>&1
timeout 300 python profiles/prof_rappids.py fast 65536 512 4 hard >> $O/rappids_times_$tag.log 2>&1
timeout 300 python profiles/prof_rappids.py parity 65536 512 4 >> $O/rappids_times_$tag.log 2>&1
grep "plans/s" $O/rappids_times_$tag.log
