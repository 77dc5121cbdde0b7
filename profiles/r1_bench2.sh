#!/bin/bash
# default bench line with the new extras (offboard loop in the kernel, RAPPIDS C5 + its CPU baseline) and both arms as the driver runs them
mkdir -p gpurun_out
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/bench_ref_g.json 2> gpurun_out/bench_ref_g.err
( time timeout 900 python bench.py ) > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
python __graft_entry__.py smoke > gpurun_out/smoke_g.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_g.log
echo done
