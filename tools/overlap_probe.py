#!/usr/bin/env python3
"""Do the IMAD.WIDE pipe and the ALU pipe overlap?  One fp_mul + 8 fp add/sub per iteration (k_overlap_probe)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sylow_b200  # noqa: E402

eng = sylow_b200.Engine(0)
sms = torch.cuda.get_device_properties(0).multi_processor_count
names = {0: "fp_mul only", 39: "fp_sub + fp_add only (per pair)", 40: "mul + 8 independent adds (same block)",
         41: "mul -> 8 adds -> mul (serial phases)", 42: "serial phases, odd warps skewed by an add phase",
         43: "serial phases, odd warps skewed by ~half a mul"}
res = []
for threads, bps in ((128, 1), (256, 1), (384, 1), (512, 1), (768, 1), (1024, 1)):
    for variant in (0, 39, 40, 41, 42, 43):
        best = 1e30
        if variant == 39 and threads > 256:  # k_tower_probe is bounded to the pairing kernels' 256 threads
            continue
        for _ in range(3):
            ms, ops = eng.imad_probe(variant, sms * bps, threads, 4000)
            best = min(best, ms)
        ns_iter = best * 1e6 / 4000
        res.append({"variant": variant, "name": names[variant], "threads": threads, "ns_per_iter": ns_iter})
        print("thr=%3d  %-50s %8.1f ns / iteration" % (threads, names[variant], ns_iter), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/overlap_probe.json", "w"), indent=1)
