#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_gpu.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_miller|k_final_exp' -s 2 -c 2 -o gpurun_out/prof_r1c -f \
   python bench.py --log2n 17 --steps 1 --warmup 3 --verify-log2n 0 --extras 0 --cpu-seconds 1 > gpurun_out/bench_ncu_full.log 2>&1; echo "ncu full exit $?"
