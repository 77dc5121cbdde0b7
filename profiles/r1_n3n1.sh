#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rappids_gpu.py -m gpu -x -q > gpurun_out/gpu_tests_n3n1.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests_n3n1.log
timeout 300 python profiles/prof_rappids.py fast 65536 512 3 >> gpurun_out/gpu_tests_n3n1.log 2>&1
echo done
