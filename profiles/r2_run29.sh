#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 500 ncu --set full --clock-control none --import-source on -k regex:rappids_plan -s 1 -c 1 -o $O/prof_rappids_fold -f python profiles/prof_rappids.py fast 16384 512 2 > $O/prof_rappids_fold.log 2>&1
cp agri-fly_b200/build/agf_rappids_plan_fast.o $O/agf_rappids_plan_fast_fold.o
tail -3 $O/prof_rappids_fold.log
