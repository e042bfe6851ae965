#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.log 2>&1; echo "n8 exit $?"
tail -1 gpurun_out/bench_n8.log | cut -c1-400
