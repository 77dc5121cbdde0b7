#!/bin/bash
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rappids|dispatch" --csv --log-file $O/rappids_launches_jump.csv python profiles/prof_rappids.py fast 65536 512 3 > $O/rappids_launches_jump.log 2>&1
grep -v "^==" $O/rappids_launches_jump.csv | cut -d, -f5,12- | head -20
