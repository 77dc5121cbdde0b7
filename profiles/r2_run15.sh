#!/bin/bash
# round 2, GPU call 15: new edge-case tests; a compute-sanitizer memcheck pass over the set/get, log, wrench and estimator tests
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( timeout 900 python -m pytest tests/test_rappids_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "other_image_sizes or ragged or wrench or log_ring" ) > $O/gpu_tests_new_w.log 2>&1; tail -5 $O/gpu_tests_new_w.log
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --target-processes all python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "ragged or wrench or log_ring or estimator_parity or set_field or per_vehicle_parameter" ) > $O/sanitizer_w.log 2>&1; echo "sanitizer rc=$?" >> $O/sanitizer_w.log
grep -i "ERROR SUMMARY\|passed\|failed\|sanitizer rc" $O/sanitizer_w.log | tail -5
