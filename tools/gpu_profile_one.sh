#!/bin/bash
# usage (GPU box): tools/gpu_profile_one.sh <tag> <kernel-regex> <skip> <count>
# One `ncu --set full` session for ONE kernel of tools/prof_driver.py (hardware counters; raw page only).
tag=$1; kn=$2; skip=$3; cnt=$4
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:"$kn" -s $skip -c $cnt -f -o /tmp/${tag}_${kn} \
    python tools/prof_driver.py 20 > gpurun_out/${tag}_prof_${kn}.log 2>&1
grep -E "passes|ERROR|WARNING.*fail" gpurun_out/${tag}_prof_${kn}.log | tail -4
ncu -i /tmp/${tag}_${kn}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw_${kn}.csv 2>/dev/null
ls -la gpurun_out/${tag}_raw_${kn}.csv
