#!/bin/bash
# first GPU pass: probes, parity tests, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python tools/probe.py > gpurun_out/probe.log 2>&1; echo "probe exit $?"
tail -40 gpurun_out/probe.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -30 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --log2n 16 --verify-log2n 12 --steps 3 --warmup 3 --cpu-seconds 3 > gpurun_out/bench_small.log 2>&1; echo "bench exit $?"
tail -5 gpurun_out/bench_small.log
