#!/bin/bash
# split planner: launch list (durations of the two passes), ncu --set full of the planning pass
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rappids --csv --log-file $O/rappids_launches_split.csv python profiles/prof_rappids.py fast 65536 512 2 > $O/rappids_launches_split.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:rappids_plan -s 1 -c 1 -o $O/prof_rappids_split -f python profiles/prof_rappids.py fast 16384 512 2 > $O/prof_rappids_split.log 2>&1
cp agri-fly_b200/build/agf_rappids_plan_fast.o $O/agf_rappids_plan_fast_split.o
grep -v "^==" $O/rappids_launches_split.csv | cut -d, -f5,12- | head
