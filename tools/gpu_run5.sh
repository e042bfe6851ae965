#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_gpu.log
python tools/tune.py 18 2>&1 | tee gpurun_out/tune4.log | sed -E 's/"n": [0-9]+, //; s/\/root\/repo\/build\/variants\///; s/"chk.*//' | cut -c1-230
