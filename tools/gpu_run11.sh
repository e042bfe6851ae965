#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1200 python bench.py > gpurun_out/bench_full5.log 2>&1; echo "bench exit $?"
tail -1 gpurun_out/bench_full5.log | cut -c1-300
