#!/bin/bash
mkdir -p gpurun_out/r2
for j in 0 8 16; do
echo "== frame jump $j"
AGF_RAPPIDS_FRAME_JUMP=$j timeout 300 python profiles/dev_rappids_work.py 65536 512 gpurun_out/r2/rappids_work_easy_j$j.npz 2>&1 | head -4
done > gpurun_out/r2/rappids_work_jump.log 2>&1
cat gpurun_out/r2/rappids_work_jump.log
