#!/bin/bash
# C4 decomposition over library variants: bash profiles/r2_c4var.sh <tag> base v1 v2 ...
mkdir -p gpurun_out/r2
tag=$1; shift
out=gpurun_out/r2/c4_matrix_$tag.log
: > $out
for v in "$@"; do
  if [ "$v" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$v.so; fi
  echo "== $v" >> $out
  timeout 200 python profiles/dev_c4_matrix.py rates 2097152 500 1 >> $out 2>&1
  timeout 200 python profiles/dev_c4_matrix.py uwb 2097152 300 1 >> $out 2>&1
done
cat $out
