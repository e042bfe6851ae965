#!/bin/bash
# usage (on the GPU box): tools/gpu_kbench.sh <out.jsonl> <log2n> variant...
out=$1; shift; n=$1; shift
mkdir -p gpurun_out
: > gpurun_out/$out
for v in "$@"; do timeout 120 build/kbench/$v $n 3 >> gpurun_out/$out 2>&1; done
cat gpurun_out/$out
