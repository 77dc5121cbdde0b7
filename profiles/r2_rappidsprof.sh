#!/bin/bash
# ncu --set full of the planner kernel (fast variant), C5 scenes, + timings of easy / hard families
mkdir -p gpurun_out/r2
O=gpurun_out/r2
tag=${1:-p}
timeout 300 python profiles/prof_rappids.py fast 65536 512 3 > $O/rappids_times_$tag.log 2>&1
timeout 300 python profiles/prof_rappids.py fast 65536 512 3 hard >> $O/rappids_times_$tag.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:plan -s 1 -c 1 -o $O/prof_rappids_$tag -f python profiles/prof_rappids.py fast 16384 512 2 > $O/prof_rappids_$tag.log 2>&1
cp agri-fly_b200/build/agf_rappids_plan_fast.o $O/agf_rappids_plan_fast_$tag.o
cat $O/rappids_times_$tag.log
