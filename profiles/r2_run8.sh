#!/bin/bash
# round 2, GPU call 8: log layout [quad][vehicle] + running ring pointer + L1 carve-out: log tests, C4 decomposition
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "log or sweep or bit_identical" ) > $O/gpu_tests_log_h.log 2>&1; tail -3 $O/gpu_tests_log_h.log
bash profiles/r2_c4var.sh h base
