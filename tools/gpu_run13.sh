#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -x -q -k "pairing" 2>&1 | tail -3
timeout 1200 python bench.py --verify-log2n 0 > gpurun_out/bench_e2e.log 2>&1; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_e2e.log').read().strip().splitlines()[-1])
print(d['value'], d['e2e'], d['gpu_launches'])
PY
