#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1200 python bench.py > gpurun_out/bench_full6.log 2>&1; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full6.log').read().strip().splitlines()[-1])
print(d['value'], d['e2e'], d['gpu_launches'])
PY
