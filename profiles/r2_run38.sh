#!/bin/bash
mkdir -p gpurun_out/r2
out=gpurun_out/r2/rappids_variants_parity_occupancy.log
: > $out
for v in "$@"; do
  if [ "$v" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$v.so; fi
  echo "== $v" >> $out
  timeout 200 python profiles/prof_rappids.py parity 65536 512 3 2>&1 | grep "plans/s" >> $out
done
cat $out
