#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1200 python bench.py > gpurun_out/bench_full4.log 2>&1; echo "bench exit $?"
tail -1 gpurun_out/bench_full4.log | cut -c1-150
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-200
