#!/bin/bash
# round 2, GPU call 1: full GPU suite (incl. the population-level parity of the fast kernels), every bench arm,
# FLOP counters of the kernels the bench times (housekeeping on), ncu --set full of the +hk FP32 kernel, the FP64 kernel
# and the C4 logging launch, launch list of the bench command.
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_a.txt 2>&1
( time timeout 2400 python -m pytest tests -m gpu -q -s -x --deselect tests/test_rappids_gpu.py ) > $O/gpu_tests_a.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests_a.log
( time timeout 900 python -m pytest tests/test_rappids_gpu.py -m gpu -q -x ) > $O/gpu_tests_rappids_a.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests_rappids_a.log
( time timeout 900 python bench.py ) > $O/bench_a.json 2> $O/bench_a.err
( time timeout 600 python bench.py --precision fp64 --no-extras ) > $O/bench_fp64_a.json 2> $O/bench_fp64_a.err
( time timeout 600 python bench.py --precision fp64 --math parity --ticks-per-step 1500 --no-extras ) > $O/bench_parity_a.json 2> $O/bench_parity_a.err
( time timeout 600 python bench.py --hk off --no-extras ) > $O/bench_hkoff_a.json 2> $O/bench_hkoff_a.err
( time timeout 600 python bench.py --config c4 --no-extras ) > $O/bench_c4_rates_a.json 2> $O/bench_c4_rates_a.err
( time timeout 600 python bench.py --config c4 --c4-mode full --no-extras ) > $O/bench_c4_full_a.json 2> $O/bench_c4_full_a.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_ref_a.json 2> $O/bench_ref_a.err
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum
for v in "fast fp32 uwb hk" "fast fp32 rates hk" "fast fp64 uwb hk" "fast fp64 rates hk" "fast fp64 rates nohk" "fast fp64 uwb nohk"; do
  set -- $v
  tag="$1_$2_$3"; [ "$4" = hk ] && tag="${tag}_hk"
  timeout 300 ncu --metrics $M --clock-control none -k regex:step_kernel -s 1 -c 1 --csv --log-file $O/flops_${tag}_4096x300.csv python profiles/flop_count.py $1 $2 $3 4096 300 $4 > $O/flops_${tag}.log 2>&1
done
# ncu --set full: FP32 +hk, FP64 fast (+hk), C4 logging (rates, sweep), parity
AGF_PROF_HK=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -o $O/prof_f32_uwb_hk python profiles/prof_step.py fp32 uwb 131072 200 2 > $O/prof_f32_uwb_hk.log 2>&1
AGF_PROF_HK=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -o $O/prof_f64_uwb_hk python profiles/prof_step.py fp64 uwb 65536 200 2 > $O/prof_f64_uwb_hk.log 2>&1
AGF_PROF_HK=1 AGF_PROF_C4=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -o $O/prof_c4_rates python profiles/prof_step.py fp32 rates 2097152 64 2 > $O/prof_c4_rates.log 2>&1
AGF_PROF_MATH=parity timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -o $O/prof_parity_uwb python profiles/prof_step.py fp64 uwb 65536 100 2 > $O/prof_parity_uwb.log 2>&1
# dram traffic of one bench-sized launch (+hk) and launch list of the bench command
AGF_NO_WARM=1 AGF_PROF_HK=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:step_kernel -s 1 -c 1 --csv --log-file $O/traffic_c3_fp32_fast_hk_131072x500.csv python profiles/prof_step.py fp32 uwb 131072 500 2 > $O/traffic_a.log 2>&1
AGF_NO_WARM=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_a.csv python bench.py --steps 3 --warmup 3 --ticks-per-step 500 --no-extras > $O/launches_a.log 2>&1
python __graft_entry__.py smoke > $O/smoke_a.log 2>&1; echo "smoke rc=$?" >> $O/smoke_a.log
echo done
