#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/bench_full2.log 2>&1; echo "bench exit $?"
tail -2 gpurun_out/bench_full2.log | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"
tail -1 gpurun_out/bench_ref.log | cut -c1-600
