#!/bin/bash
# kernel-only throughput of the library and of tuning variants (FP32 full mode, 131072 vehicles x 500 ticks)
mkdir -p gpurun_out
: > gpurun_out/variants.log
for v in base "$@"; do
  if [ "$v" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$v.so; fi
  echo "== $v" >> gpurun_out/variants.log
  timeout 120 python profiles/prof_step.py fp32 uwb 131072 500 4 >> gpurun_out/variants.log 2>&1
done
unset AGF_LIB_PATH
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits > gpurun_out/smi_query.txt 2>&1
nvidia-smi --help-query-gpu 2>/dev/null | grep -i -A1 "reasons\." | head -60 > gpurun_out/smi_help.txt
echo done
