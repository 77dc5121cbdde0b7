#!/bin/bash
# kernel-only throughput of library variants: FP32 full mode and rates mode, housekeeping on, 131072 vehicles x 500 ticks
# usage: bash profiles/r2_variants.sh <tag> base v1 v2 ...
mkdir -p gpurun_out/r2
tag=$1; shift
out=gpurun_out/r2/variants_$tag.log
: > $out
for v in "$@"; do
  if [ "$v" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$v.so; fi
  echo "== $v" >> $out
  AGF_PROF_HK=1 timeout 120 python profiles/prof_step.py fp32 uwb 131072 500 4 2>&1 | grep "step kernel" >> $out
  AGF_PROF_HK=1 timeout 120 python profiles/prof_step.py fp32 rates 131072 500 4 2>&1 | grep "step kernel" >> $out
done
cat $out
