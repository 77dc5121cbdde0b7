#!/bin/bash
# candidate pass at other occupancies (kernel durations from an ncu launch list; not bench values)
mkdir -p gpurun_out/r2
out=gpurun_out/r2/rappids_variants_candidates_occupancy.log
: > $out
for v in "$@"; do
  if [ "$v" = base ]; then unset AGF_LIB_PATH; else export AGF_LIB_PATH=$PWD/agri-fly_b200/variants/libagrifly_b200_$v.so; fi
  echo "== $v" >> $out
  timeout 200 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:rappids_candidates -c 2 --csv python profiles/prof_rappids.py fast 65536 512 1 2>&1 | grep "rappids_candidates" | cut -d, -f13- >> $out
done
cat $out
