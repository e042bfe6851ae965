#!/bin/bash
# usage (GPU box): tools/gpu_lanes.sh <tag> [variant...]   - k_miller (one thread per pair) against k_miller_lanes (two
# lanes per pair) at full size and at the small sizes where the split is meant to pay; then one ncu capture of each.
tag=$1; shift
vs=${@:-l0}
mkdir -p gpurun_out
out=gpurun_out/${tag}_lanes.jsonl
: > $out
first=1
for v in $vs; do
  b=build/kbench/$v
  timeout 120 $b 20 2 >> $out 2>&1
  [ $first = 1 ] || continue
  first=0
  timeout 60 $b 17 3 >> $out 2>&1
  timeout 60 $b 0 5 17408 64 128 >> $out 2>&1      # the remainder of 2^17 pairs after three whole waves
  timeout 60 $b 0 5 18944 64 128 >> $out 2>&1      # 148 x 128 pairs: one full wave of the two-lane kernel
  timeout 60 $b 0 5 9472 32 128 >> $out 2>&1       # 148 x 64
  timeout 60 $b 0 5 9472 32 64 >> $out 2>&1
  timeout 60 $b 0 5 4096 32 64 >> $out 2>&1
  timeout 60 $b 0 5 4096 32 32 >> $out 2>&1
  timeout 60 $b 0 5 1024 32 32 >> $out 2>&1
  timeout 60 $b 0 5 64 32 32 >> $out 2>&1
  timeout 60 $b 0 5 1 32 32 >> $out 2>&1
done
cat $out
v=${vs%% *}
for kn in k_miller_lanes k_miller; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:"^${kn}\$" -c 1 -f -o /tmp/${tag}_${kn} \
      build/kbench/$v 20 1 > gpurun_out/${tag}_ncu_${kn}.log 2>&1
  ncu -i /tmp/${tag}_${kn}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw_${kn}.csv 2>/dev/null
  ncu -i /tmp/${tag}_${kn}.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_src_${kn}.csv.gz
done
ls -la gpurun_out | tail
