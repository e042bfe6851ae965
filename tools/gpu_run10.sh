#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
echo "memcheck exit ${PIPESTATUS[0]}"; tail -5 gpurun_out/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "glued or groth or threshold or verify or precompute or sum" 2>&1 | tail -5
echo "racecheck exit ${PIPESTATUS[0]}"; tail -5 gpurun_out/racecheck.log
