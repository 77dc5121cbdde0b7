#!/bin/bash
# FLOP / instruction counters of the final kernels incl. the packed FP32 opcodes (FADD2 / FMUL2 / FFMA2 have counters of their own)
mkdir -p gpurun_out/r2
O=gpurun_out/r2
M=""
for op in fadd fmul ffma fadd2 fmul2 ffma2 dadd dmul dfma; do M="$M,smsp__sass_thread_inst_executed_op_${op}_pred_on.sum"; done
M="${M:1},smsp__inst_executed.sum,smsp__thread_inst_executed.sum"
for v in "fast fp32 uwb hk" "fast fp32 uwb nohk" "fast fp32 rates hk" "fast fp32 rates nohk" "fast fp64 uwb hk" "fast fp64 rates hk"; do
  set -- $v
  tag="$1_$2_$3"; [ "$4" = hk ] && tag="${tag}_hk"
  timeout 300 ncu --metrics $M --clock-control none -k regex:step_kernel -s 1 -c 1 --csv --log-file $O/flops_${tag}_4096x300.csv python profiles/flop_count.py $1 $2 $3 4096 300 $4 > $O/flops_${tag}.log 2>&1
  tail -1 $O/flops_${tag}.log
done
grep -c "ffma2" $O/flops_fast_fp32_uwb_hk_4096x300.csv
