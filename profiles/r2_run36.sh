#!/bin/bash
# compute-sanitizer memcheck over the planner's GPU tests (new kernels: candidate pass, frame_pair vector loads, dispatch order)
mkdir -p gpurun_out/r2
( time timeout 800 compute-sanitizer --tool memcheck --error-exitcode 7 --target-processes all python -m pytest tests/test_rappids_gpu.py -x -q -m gpu -k "edge_cases or other_image_sizes or golden or images_from_host or device_sampler or planned_primitives" ) > gpurun_out/r2/sanitizer_rappids.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r2/sanitizer_rappids.log
grep -E "ERROR SUMMARY|passed|failed|rc=" gpurun_out/r2/sanitizer_rappids.log | tail -5
