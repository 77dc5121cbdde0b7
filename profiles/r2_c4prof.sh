#!/bin/bash
# steady-state ncu --set full of the C4 rates kernel (sweep + log every tick) and of the same kernel with the log off
mkdir -p gpurun_out/r2
O=gpurun_out/r2
for lg in 1 0; do
AGF_PROF_HK=1 AGF_PROF_C4=1 AGF_PROF_LOG=$lg timeout 500 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1 -c 1 -o $O/prof_c4_rates_log${lg}_g -f python profiles/prof_step.py fp32 rates 2097152 300 2 > $O/prof_c4_rates_log${lg}_g.log 2>&1
done
cp agri-fly_b200/build/agf_kernels_fast_f32_rates.o $O/agf_kernels_fast_f32_rates_g.o
tail -3 $O/prof_c4_rates_log*_g.log
