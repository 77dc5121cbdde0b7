#!/bin/bash
# offboard-loop kernels of library variants: bash profiles/r2_offvariants.sh <tag> base v1 ...
mkdir -p gpurun_out/r2
tag=$1; shift
out=gpurun_out/r2/offboard_variants_$tag.log
: > $out
for v in "$@"; do
  if [ "$v" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$v.so; fi
  echo "== $v" >> $out
  timeout 200 python profiles/prof_offboard.py fp32 mocap targets 131072 500 3 >> $out 2>&1
  timeout 200 python profiles/prof_offboard.py fp32 truth targets 131072 500 3 >> $out 2>&1
  timeout 200 python profiles/prof_offboard.py fp32 mocap stages 131072 500 3 >> $out 2>&1
done
cat $out
