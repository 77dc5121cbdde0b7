#!/bin/bash
mkdir -p gpurun_out/r2
sleep 2; ( time timeout 700 compute-sanitizer --tool memcheck --error-exitcode 7 --target-processes all python -m pytest tests/test_rappids_gpu.py -x -q -m gpu -k "bit_identical or fast_variant_agrees" ) > gpurun_out/r2/sanitizer_rappids2.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r2/sanitizer_rappids2.log
grep -E "ERROR SUMMARY|passed|failed|rc=" gpurun_out/r2/sanitizer_rappids2.log | tail -5
