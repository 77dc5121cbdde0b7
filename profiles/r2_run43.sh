#!/bin/bash
mkdir -p gpurun_out/r2
out=gpurun_out/r2/rappids_coop_final.log
: > $out
for fam in "" hard; do
  echo "== cooperative planning of long vehicles: default (factor 0.3, cap 0.04) $fam" >> $out
  timeout 120 python profiles/prof_rappids.py fast 65536 512 4 $fam 2>&1 | grep "plans/s" >> $out
  echo "== off $fam" >> $out
  AGF_RAPPIDS_COOP_FACTOR=0 timeout 120 python profiles/prof_rappids.py fast 65536 512 4 $fam 2>&1 | grep "plans/s" >> $out
done
echo "== default, 131072 and 32768 plans" >> $out
timeout 200 python profiles/prof_rappids.py fast 131072 512 3 2>&1 | grep "plans/s" >> $out
timeout 200 python profiles/prof_rappids.py fast 32768 512 3 2>&1 | grep "plans/s" >> $out
echo "== parity variant, default" >> $out
timeout 200 python profiles/prof_rappids.py parity 65536 512 3 2>&1 | grep "plans/s" >> $out
cat $out
