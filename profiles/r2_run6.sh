#!/bin/bash
# round 2, GPU call 6: whole GPU suite after the kernel diet / parity outlining / UWB noise model / estimator preload,
# the default bench (extras: parity, offboard loops), FLOP + instruction counters and an ncu --set full of the FP32 +hk kernel.
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( time timeout 2400 python -m pytest tests -m gpu -q -x --durations=8 ) > $O/gpu_tests_e.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests_e.log
( time timeout 900 python bench.py ) > $O/bench_e.json 2> $O/bench_e.err
( time timeout 600 python bench.py --precision fp64 --math parity --ticks-per-step 3000 --no-extras ) > $O/bench_parity_e.json 2> $O/bench_parity_e.err
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum
for v in "fast fp32 uwb hk" "fast fp32 rates hk" "fast fp64 uwb hk" "parity fp64 uwb hk"; do
  set -- $v
  tag="e_$1_$2_$3"; [ "$4" = hk ] && tag="${tag}_hk"
  timeout 300 ncu --metrics $M --clock-control none -k regex:step_kernel -s 1 -c 1 --csv --log-file $O/flops_${tag}_4096x300.csv python profiles/flop_count.py $1 $2 $3 4096 300 $4 > $O/flops_${tag}.log 2>&1
done
AGF_PROF_HK=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -o $O/prof_f32_uwb_hk_e python profiles/prof_step.py fp32 uwb 131072 200 2 > $O/prof_f32_uwb_hk_e.log 2>&1
AGF_PROF_MATH=parity timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -o $O/prof_parity_uwb_e python profiles/prof_step.py fp64 uwb 65536 100 2 > $O/prof_parity_uwb_e.log 2>&1
cp agri-fly_b200/build/agf_kernels_fast_f32_uwb.o $O/agf_kernels_fast_f32_uwb_e.o
echo done
