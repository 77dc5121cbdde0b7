#!/bin/bash
# round 2, GPU call 16: compute-sanitizer memcheck over the GPU suite minus the full-size population tests
mkdir -p gpurun_out/r2
O=gpurun_out/r2
K="not full_size and not population and not monte_carlo and not fleet and not ground_truth"
( time timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 7 --target-processes all python -m pytest tests -m gpu -q -x -k "$K" ) > $O/sanitizer_suite.log 2>&1; echo "sanitizer rc=$?" >> $O/sanitizer_suite.log
grep -i "ERROR SUMMARY\|passed\|failed\|sanitizer rc\|real" $O/sanitizer_suite.log | tail -6
