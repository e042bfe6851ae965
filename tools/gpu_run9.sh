#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/tune.py 20 > gpurun_out/tune8.log 2>&1; cat gpurun_out/tune8.log
