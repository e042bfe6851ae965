#!/usr/bin/env python3
"""Run hash_to_g1 / hash_to_field / sign on 2^k device-resident messages (for `ncu --metrics gpu__time_duration.sum`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import sylow_b200

n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
eng = sylow_b200.Engine(0)
dev = torch.device("cuda", 0)
rs = np.random.RandomState(1)
d_msgs = torch.from_numpy(rs.randint(0, 256, size=n * 32, dtype=np.uint8)).to(dev)
d_offs = torch.from_numpy((np.arange(n + 1, dtype=np.int64) * 32)).to(dev)
d_out = torch.empty((n, 64), dtype=torch.uint8, device=dev)
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.hash_to_g1_batch_dev(d_msgs, d_offs, d_out)
    b.record()
    torch.cuda.synchronize()
    print("hash_to_g1 n=%d: %.3f ms" % (n, a.elapsed_time(b)))
