#!/usr/bin/env python3
"""IMAD-pipe probes on the GPU box: raw IMAD / IMAD.WIDE / IMAD.WIDE.X rates and Montgomery-multiplication
throughput at several residencies.  Writes gpurun_out/imad_probe.json (the roofline denominator)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sylow_b200  # noqa: E402

eng = sylow_b200.Engine(0)
sms = torch.cuda.get_device_properties(0).multi_processor_count
res = {"sms": sms, "name": torch.cuda.get_device_name(0), "probes": []}
names = {0: "fp_mul x1 chain", 1: "fp_mul x2 chains", 2: "fp_mul x4 chains", 10: "mad.wide.u32 independent",
         11: "mad.lo.u32 (32-bit IMAD)", 12: "mad.lo.cc/madc.hi.cc chains (IMAD.WIDE.U32.X)",
         13: "fma.rn.f64 independent (FP64 pipe)"}
for variant in (12, 13, 10, 11, 0, 1, 2):
    for threads, bps in ((128, 1), (128, 2), (256, 2), (256, 4), (256, 8), (512, 4)):
        iters = 4000 if variant >= 10 else 3000
        best = 0.0
        for _ in range(3):
            ms, ops = eng.imad_probe(variant, sms * bps, threads, iters)
            best = max(best, ops / (ms * 1e-3))
        scale = 136 if variant < 10 else 1
        rec = {"variant": variant, "name": names[variant], "threads_per_block": threads, "blocks_per_sm": bps,
               "warps_per_sm": threads * bps // 32, "ops_per_s": best, "limb_products_per_s": best * scale}
        res["probes"].append(rec)
        print("%-48s thr=%3d bps=%d warps/SM=%2d  %8.3f T/s  (limb-products %7.3f T/s)" % (
            names[variant], threads, bps, rec["warps_per_sm"], best / 1e12, best * scale / 1e12), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/imad_probe.json", "w"), indent=1)
