#!/bin/bash
# round 2, final evidence: whole GPU suite, every bench arm, FLOP / instruction counters, DRAM traffic, launch list,
# ncu --set full of the FP32 +hk, FP64-plant +hk and C4 logging kernels (summarised here with profiles/ncu_by_line.py).
mkdir -p gpurun_out/r2
O=gpurun_out/r2
T=final
( time timeout 2400 python -m pytest tests -m gpu -q -s --durations=8 ) > $O/gpu_tests_$T.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests_$T.log
( time timeout 900 python bench.py ) > $O/bench_$T.json 2> $O/bench_$T.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_ref_$T.json 2> $O/bench_ref_$T.err
( time timeout 600 python bench.py --precision fp64 --no-extras ) > $O/bench_fp64_$T.json 2> $O/bench_fp64_$T.err
( time timeout 600 python bench.py --precision fp64 --math parity --ticks-per-step 3000 --no-extras ) > $O/bench_parity_$T.json 2> $O/bench_parity_$T.err
( time timeout 600 python bench.py --hk off --no-extras ) > $O/bench_hkoff_$T.json 2> $O/bench_hkoff_$T.err
( time timeout 600 python bench.py --config c4 --no-extras ) > $O/bench_c4_rates_$T.json 2> $O/bench_c4_rates_$T.err
( time timeout 600 python bench.py --config c4 --c4-mode full --no-extras ) > $O/bench_c4_full_$T.json 2> $O/bench_c4_full_$T.err
ncu --query-metrics 2>/dev/null | grep -i "sass_thread_inst_executed_op\|inst_executed_pipe_fma" > $O/ncu_metric_names.txt
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__sass_thread_inst_executed_ops_fadd_fmul_ffma_pred_on.sum
for v in "fast fp32 uwb hk" "fast fp32 uwb nohk" "fast fp32 rates hk" "fast fp64 uwb hk" "fast fp64 rates hk"; do
  set -- $v
  tag="$1_$2_$3"; [ "$4" = hk ] && tag="${tag}_hk"
  timeout 300 ncu --metrics $M --clock-control none -k regex:step_kernel -s 1 -c 1 --csv --log-file $O/flops_${tag}_4096x300.csv python profiles/flop_count.py $1 $2 $3 4096 300 $4 > $O/flops_${tag}.log 2>&1
done
D=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
AGF_NO_WARM=1 AGF_PROF_HK=1 timeout 300 ncu --metrics $D --clock-control none -k regex:step_kernel -s 1 -c 1 --csv --log-file $O/traffic_c3_fp32_fast_hk_131072x500.csv python profiles/prof_step.py fp32 uwb 131072 500 2 > $O/traffic_c3.log 2>&1
AGF_NO_WARM=1 AGF_PROF_HK=1 AGF_PROF_C4=1 timeout 300 ncu --metrics $D --clock-control none -k regex:step_kernel -s 1 -c 1 --csv --log-file $O/traffic_c4_fp32_rates_2097152x300.csv python profiles/prof_step.py fp32 rates 2097152 300 2 > $O/traffic_c4_rates.log 2>&1
AGF_NO_WARM=1 AGF_PROF_HK=1 AGF_PROF_C4=1 timeout 300 ncu --metrics $D --clock-control none -k regex:step_kernel -s 1 -c 1 --csv --log-file $O/traffic_c4_fp32_full_2097152x200.csv python profiles/prof_step.py fp32 uwb 2097152 200 2 > $O/traffic_c4_full.log 2>&1
AGF_NO_WARM=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_$T.csv python bench.py --steps 3 --warmup 3 --ticks-per-step 500 --no-extras > $O/launches_$T.log 2>&1
AGF_PROF_HK=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -f -o $O/prof_f32_uwb_hk_$T python profiles/prof_step.py fp32 uwb 131072 200 2 > $O/prof_f32_uwb_hk_$T.log 2>&1
AGF_PROF_HK=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -f -o $O/prof_f64_uwb_hk_$T python profiles/prof_step.py fp64 uwb 65536 200 2 > $O/prof_f64_uwb_hk_$T.log 2>&1
AGF_PROF_HK=1 AGF_PROF_C4=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -f -o $O/prof_c4_rates_$T python profiles/prof_step.py fp32 rates 2097152 300 2 > $O/prof_c4_rates_$T.log 2>&1
for f in agf_kernels_fast_f32_uwb agf_kernels_fast_f64_uwb agf_kernels_fast_f32_rates agf_rappids_plan_fast; do cp agri-fly_b200/build/$f.o $O/${f}_$T.o; done
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$T.log 2>&1; echo "smoke rc=$?" >> $O/smoke_$T.log
tail -3 $O/gpu_tests_$T.log; tail -2 $O/smoke_$T.log; head -c 300 $O/bench_$T.json
