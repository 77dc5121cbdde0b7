#!/bin/bash
mkdir -p gpurun_out/r2
timeout 300 python profiles/dev_rappids_work.py 65536 512 gpurun_out/r2/rappids_work_easy.npz > gpurun_out/r2/rappids_work.log 2>&1
timeout 300 python profiles/dev_rappids_work.py 65536 512 gpurun_out/r2/rappids_work_hard.npz hard >> gpurun_out/r2/rappids_work.log 2>&1
cat gpurun_out/r2/rappids_work.log
