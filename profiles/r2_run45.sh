#!/bin/bash
mkdir -p gpurun_out/r2
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke_final3.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2/smoke_final3.log
( time timeout 600 python bench.py ) > gpurun_out/r2/bench_final3.json 2> gpurun_out/r2/bench_final3.err
tail -2 gpurun_out/r2/smoke_final3.log; head -c 250 gpurun_out/r2/bench_final3.json
