#!/bin/bash
# round 2, closing evidence after the planner work (step kernels unchanged since r2_final.sh): whole GPU suite, both bench
# arms, smoke, planner launch list and ncu --set full of the planning pass
mkdir -p gpurun_out/r2
O=gpurun_out/r2
T=final2
( time timeout 1500 python -m pytest tests -m gpu -q -s --durations=8 ) > $O/gpu_tests_$T.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests_$T.log
( time timeout 900 python bench.py ) > $O/bench_$T.json 2> $O/bench_$T.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_ref_$T.json 2> $O/bench_ref_$T.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$T.log 2>&1; echo "smoke rc=$?" >> $O/smoke_$T.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rappids|dispatch" --csv --log-file $O/rappids_launches_$T.csv python profiles/prof_rappids.py fast 65536 512 3 > $O/rappids_launches_$T.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:rappids_plan -s 1 -c 1 -o $O/prof_rappids_$T -f python profiles/prof_rappids.py fast 16384 512 2 > $O/prof_rappids_$T.log 2>&1
cp agri-fly_b200/build/agf_rappids_plan_fast.o $O/agf_rappids_plan_fast_$T.o
tail -3 $O/gpu_tests_$T.log; tail -2 $O/smoke_$T.log; head -c 400 $O/bench_$T.json; echo; head -c 300 $O/bench_ref_$T.json
