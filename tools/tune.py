#!/usr/bin/env python3
"""Time k_miller / k_final_exp for every library variant under build/variants (launch-bounds tuning)."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch

    import sylow_b200

    log2n = int(sys.argv[2])
    n = 1 << log2n
    eng = sylow_b200.Engine(0)
    dev = torch.device("cuda", 0)
    rs = np.random.RandomState(1)
    k = rs.randint(0, 256, size=(2 * n, 32), dtype=np.uint8)
    k[:, 31] &= 0x1F
    g1 = np.zeros((n, 64), np.uint8)
    g1[:, 0], g1[:, 32] = 1, 2
    G2 = (10857046999023057135944570762232829481370756359578518086990519993285655852781,
          11559732032986387107991004021392285783925812861821192530917403151452391805634,
          8495653923123431417604973247489272438418190587263600148770280649306958101930,
          4082367875863433681332203403145435568316851327593401208105741076214120093531)
    g2 = np.tile(np.frombuffer(b"".join(c.to_bytes(32, "little") for c in G2), dtype=np.uint8), (n, 1))
    d_g1 = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    d_g2 = torch.empty((n, 128), dtype=torch.uint8, device=dev)
    eng.g1_mul_batch_dev(torch.from_numpy(g1).to(dev), torch.from_numpy(k[:n]).to(dev), d_g1)
    eng.g2_mul_batch_dev(torch.from_numpy(g2).to(dev), torch.from_numpy(k[n:]).to(dev), d_g2)
    d_k = torch.from_numpy(k[:n]).to(dev)
    d_g1o = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    d_msgs = torch.from_numpy(rs.randint(0, 256, size=n * 32, dtype=np.uint8)).to(dev)
    d_offs = torch.from_numpy((np.arange(n + 1, dtype=np.int64) * 32)).to(dev)
    d_f = torch.empty((n, 384), dtype=torch.uint8, device=dev)
    d_o = torch.empty((n, 384), dtype=torch.uint8, device=dev)

    def t(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    ms_m = t(lambda: eng.miller_loop_batch_dev(d_g1, d_g2, d_f))
    ms_f = t(lambda: eng.final_exp_batch_dev(d_f, d_o))
    # Fp-only kernels: G1 ladder and hash-to-curve
    ms_g1 = t(lambda: eng.g1_mul_batch_dev(d_g1, torch.from_numpy(k[:n]).to(dev) if False else d_k, d_g1o))
    ms_h = t(lambda: eng.hash_to_g1_batch_dev(d_msgs, d_offs, d_g1o))
    d_g2o = torch.empty((n, 128), dtype=torch.uint8, device=dev)
    ms_g2 = t(lambda: eng.g2_mul_batch_dev(d_g2, d_k, d_g2o))
    chk = int(d_o[:64].to(torch.int64).sum().item())
    print(json.dumps({"lib": os.environ.get("SYLOW_B200_LIB"), "n": n, "ms_miller": ms_m, "ms_fexp": ms_f,
                      "miller_per_s": n / ms_m * 1e3, "fexp_per_s": n / ms_f * 1e3,
                      "pairings_per_s": n / (ms_m + ms_f) * 1e3, "g1_mul_per_s": n / ms_g1 * 1e3, "g2_mul_per_s": n / ms_g2 * 1e3,
                      "hash_per_s": n / ms_h * 1e3, "chk": chk}))
else:
    log2n = sys.argv[1] if len(sys.argv) > 1 else "18"
    libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so")))
    for lib in libs:
        env = dict(os.environ, SYLOW_B200_LIB=lib)
        r = subprocess.run([sys.executable, __file__, "--one", log2n], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-500:], flush=True)
