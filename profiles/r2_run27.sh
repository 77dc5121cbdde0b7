#!/bin/bash
mkdir -p gpurun_out/r2
export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_clk.so
for j in 0 8; do
echo "== frame jump $j"
AGF_RAPPIDS_FRAME_JUMP=$j timeout 300 python profiles/prof_rappids.py fast 65536 512 1 2>&1 | grep "phase cycles" | tail -7
done > gpurun_out/r2/rappids_phase_clocks_jump.log 2>&1
cat gpurun_out/r2/rappids_phase_clocks_jump.log
