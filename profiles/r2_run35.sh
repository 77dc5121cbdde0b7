#!/bin/bash
# planner throughput against the population size (is a 65 536-plan launch bound by its longest plan?)
mkdir -p gpurun_out/r2
out=gpurun_out/r2/rappids_scaling_n.log
: > $out
for n in 16384 32768 131072 262144; do
  timeout 300 python profiles/prof_rappids.py fast $n 512 3 2>&1 | grep "plans/s" >> $out
done
timeout 300 python profiles/dev_rappids_work.py 65536 512 gpurun_out/r2/rappids_work_final.npz 2>&1 | head -3 >> $out
cat $out
