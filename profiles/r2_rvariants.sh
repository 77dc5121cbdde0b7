#!/bin/bash
# planner timings of library variants: bash profiles/r2_rvariants.sh <tag> base v1 ...
mkdir -p gpurun_out/r2
tag=$1; shift
out=gpurun_out/r2/rappids_variants_$tag.log
: > $out
for v in "$@"; do
  if [ "$v" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$v.so; fi
  echo "== $v" >> $out
  timeout 200 python profiles/prof_rappids.py fast 65536 512 4 2>&1 | grep "plans/s" >> $out
  timeout 200 python profiles/prof_rappids.py fast 65536 512 4 hard 2>&1 | grep "plans/s" >> $out
done
cat $out
