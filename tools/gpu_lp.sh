#!/bin/bash
# usage (GPU box): tools/gpu_lp.sh <tag> variant...  - kbench variants at 2^20 twice, then compute-sanitizer over the two-lane kernels
tag=$1; shift
mkdir -p gpurun_out
out=gpurun_out/${tag}_kbench.jsonl
: > $out
for rep in 1 2; do for v in "$@"; do timeout 120 build/kbench/$v 20 3 >> $out 2>&1; done; done
cut -c1-330 $out
b=build/kbench/$1
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool $b 0 1 40 32 32 > gpurun_out/${tag}_sanitizer_$tool.log 2>&1
  tail -3 gpurun_out/${tag}_sanitizer_$tool.log
done
