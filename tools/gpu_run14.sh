#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01f.csv \
   python bench.py --steps 2 --warmup 3 --cpu-seconds 1 > gpurun_out/bench_ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_final_exp' -s 3 -c 1 -o gpurun_out/prof_r1f_fexp -f \
   python bench.py --steps 1 --warmup 3 --verify-log2n 0 --extras 0 --cpu-seconds 1 > gpurun_out/bench_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r01f.csv
