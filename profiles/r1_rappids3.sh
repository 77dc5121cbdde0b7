#!/bin/bash
# RAPPIDS planner with group-minimum tables: GPU parity tests, timings, occupancy variants, ncu capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rappids_gpu.py -m gpu -x -q > gpurun_out/gpu_tests_rappids.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests_rappids.log
: > gpurun_out/rappids_times6.log
for m in fast parity; do
  timeout 300 python profiles/prof_rappids.py $m 65536 512 3 >> gpurun_out/rappids_times6.log 2>&1
done
timeout 300 python profiles/prof_rappids.py fast 65536 512 3 hard >> gpurun_out/rappids_times6.log 2>&1
for v in rp_mb4 rp_mb6; do
  echo "== variant $v" >> gpurun_out/rappids_times6.log
  AGF_LIB_PATH=agri-fly_b200/variants/libagrifly_b200_$v.so timeout 300 python profiles/prof_rappids.py fast 65536 512 3 >> gpurun_out/rappids_times6.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:plan -s 1 -c 1 -o gpurun_out/prof_rappids_fast6 python profiles/prof_rappids.py fast 16384 512 2 > gpurun_out/prof_rappids6.log 2>&1
echo done
