#!/bin/bash
# usage (GPU box): tools/gpu_lanes2.sh <tag>  - parity test of the two-lane kernels, kbench of the two-lane final
# exponentiation, the launch-policy sweep, and an ncu capture of k_final_exp_lanes.
tag=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "lanes or chunk" > gpurun_out/${tag}_pytest.log 2>&1
tail -5 gpurun_out/${tag}_pytest.log
out=gpurun_out/${tag}_fexp_lanes.jsonl
: > $out
b=build/kbench/l1
timeout 120 $b 20 1 >> $out 2>&1
timeout 60 $b 0 3 17408 64 128 >> $out 2>&1
timeout 60 $b 0 3 9472 32 128 >> $out 2>&1
timeout 60 $b 0 3 4096 32 64 >> $out 2>&1
timeout 60 $b 0 3 1 32 32 >> $out 2>&1
cat $out
timeout 900 python tools/tail_policy_sweep.py > gpurun_out/${tag}_policy_sweep.jsonl 2> gpurun_out/${tag}_policy_sweep.err
cat gpurun_out/${tag}_policy_sweep.jsonl; tail -3 gpurun_out/${tag}_policy_sweep.err
kn=k_final_exp_lanes
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"^${kn}\$" -c 1 -f -o /tmp/${tag}_${kn} \
    build/kbench/l1 17 1 > gpurun_out/${tag}_ncu_${kn}.log 2>&1
ncu -i /tmp/${tag}_${kn}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw_${kn}.csv 2>/dev/null
ls -la gpurun_out | tail -8
