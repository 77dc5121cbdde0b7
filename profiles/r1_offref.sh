#!/bin/bash
# full GPU suite (incl. the offboard reference generators) + RAPPIDS occupancy variants
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_e.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests_e.log
for m in fast parity; do
  timeout 300 python profiles/prof_rappids.py $m 65536 512 3 >> gpurun_out/rappids_times4.log 2>&1
done
for v in rp_mb6 rp_mb8; do
  echo "== variant $v" >> gpurun_out/rappids_times4.log
  AGF_LIB_PATH=agri-fly_b200/variants/libagrifly_b200_$v.so timeout 300 python profiles/prof_rappids.py fast 65536 512 3 >> gpurun_out/rappids_times4.log 2>&1
done
echo done
