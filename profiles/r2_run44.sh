#!/bin/bash
mkdir -p gpurun_out/r2
( echo "== per-vehicle cycles of the third plan of a handle, cooperative long plans on (default)"; timeout 200 python profiles/dev_rappids_work.py 65536 512 gpurun_out/r2/rappids_work_coop.npz 2>&1 | head -6; echo "== off"; AGF_RAPPIDS_COOP_FACTOR=0 timeout 200 python profiles/dev_rappids_work.py 65536 512 gpurun_out/r2/rappids_work_nocoop.npz 2>&1 | head -6 ) > gpurun_out/r2/rappids_work_coop.log 2>&1
cat gpurun_out/r2/rappids_work_coop.log
