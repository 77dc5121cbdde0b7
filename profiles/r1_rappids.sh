#!/bin/bash
# GPU test suite + RAPPIDS planner (K6) timings at C5 size + ncu of the plan kernel + default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests.log
for m in fast parity; do
  timeout 300 python profiles/prof_rappids.py $m 65536 512 3 >> gpurun_out/rappids_times.log 2>&1
done
timeout 300 python profiles/prof_rappids.py fast 65536 512 3 hard >> gpurun_out/rappids_times.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_rappids.csv python profiles/prof_rappids.py fast 16384 512 2 > gpurun_out/launches_rappids.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:plan -s 1 -c 1 -o gpurun_out/prof_rappids_fast python profiles/prof_rappids.py fast 16384 512 2 > gpurun_out/prof_rappids.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
echo done
