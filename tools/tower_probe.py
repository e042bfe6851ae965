#!/usr/bin/env python3
"""Tower-level throughput ladder (k_tower_probe): limb-product rate of dependent fp2/fp6/fp12 operations at the
pairing kernels' launch shape, next to back-to-back fp_mul.  Writes gpurun_out/tower_probe.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sylow_b200  # noqa: E402

eng = sylow_b200.Engine(0)
sms = torch.cuda.get_device_properties(0).multi_processor_count
names = {0: "fp_mul chain", 30: "fp2_mul (3 M)", 31: "fp2_sqr (2 M)", 32: "fp6_mul (18 M)", 33: "fp12_mul (54 M)",
         34: "fp12_sqr (36 M)", 35: "fp12_sparse_mul (39 M)", 36: "cyclotomic_squared (18 M)",
         37: "g2_doubling_step (26 M) + 3 fp2_add", 38: "fp2_sub + fp2_add (ops/s, no products)",
         39: "fp_sub + fp_add (ops/s, no products)"}
res = []
for variant in (0, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39):
    iters = {0: 3000, 30: 2000, 31: 2000, 38: 4000, 39: 8000}.get(variant, 200)
    best = 0.0
    for _ in range(3):
        ms, ops = eng.imad_probe(variant, sms, 256, iters)
        best = max(best, ops / (ms * 1e-3))
    rec = {"variant": variant, "name": names[variant], "fp_mul_equiv_per_s": best, "limb_products_per_s": best * 136}
    res.append(rec)
    print("%-44s %8.3f G Fp-mul-equiv/s   %6.3f T limb-products/s" % (names[variant], best / 1e9, best * 136 / 1e12),
          flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/tower_probe.json", "w"), indent=1)
