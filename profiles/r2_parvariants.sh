#!/bin/bash
# parity-kernel occupancy variants: kernel-only rate of the FP64 parity full-mode kernel, 65536 vehicles x 300 ticks
mkdir -p gpurun_out/r2
out=gpurun_out/r2/parity_variants.log
: > $out
for v in base "$@"; do
  if [ "$v" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$v.so; fi
  echo "== $v" >> $out
  AGF_PROF_MATH=parity AGF_PROF_HK=1 timeout 200 python profiles/prof_step.py fp64 uwb 65536 300 3 2>&1 | grep "step kernel" >> $out
  AGF_PROF_MATH=parity AGF_PROF_HK=1 timeout 200 python profiles/prof_step.py fp64 rates 65536 300 3 2>&1 | grep "step kernel" >> $out
done
cat $out
