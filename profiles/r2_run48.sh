#!/bin/bash
mkdir -p gpurun_out/r2
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2/gpu_tests_final3.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2/gpu_tests_final3.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke_final3.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2/smoke_final3.log
tail -6 gpurun_out/r2/gpu_tests_final3.log | head -3; tail -2 gpurun_out/r2/smoke_final3.log
