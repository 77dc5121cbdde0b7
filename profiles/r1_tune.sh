#!/bin/bash
# round-1 tuning run (one gpurun call): GPU tests, kernel-only throughput of the library and its tuning
# variants, ncu launch list + full capture of the FP32 full-mode step kernel, executed-FLOP counters.
mkdir -p gpurun_out
OPS=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_xu.sum
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/gpu_tests.log
for v in base "$@"; do
  if [ "$v" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$v.so; fi
  echo "== $v" >> gpurun_out/variants.log
  timeout 120 python profiles/prof_step.py fp32 uwb 131072 500 4 >> gpurun_out/variants.log 2>&1
done
unset AGF_LIB_PATH
for m in "fp32 rates" "fp64 uwb" "fp64 rates"; do echo "== base $m" >> gpurun_out/variants.log; timeout 120 python profiles/prof_step.py $m 131072 500 3 >> gpurun_out/variants.log 2>&1; done
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?" >> gpurun_out/bench_b.err
AGF_NO_WARM=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_b.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/launches_b.log 2>&1
AGF_NO_WARM=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -o gpurun_out/prof_f32_uwb_b python profiles/prof_step.py fp32 uwb 131072 100 2 > gpurun_out/prof_full_b.log 2>&1
for w in "parity fp64 uwb" "parity fp64 rates" "fast fp32 uwb" "fast fp32 rates" "fast fp64 uwb"; do
  timeout 200 ncu --metrics $OPS --clock-control none -k regex:step_kernel -s 1 -c 1 --csv --log-file "gpurun_out/flops_${w// /_}.csv" python profiles/flop_count.py $w 4096 300 > "gpurun_out/flops_${w// /_}.log" 2>&1
done
echo done
