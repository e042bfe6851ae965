#!/usr/bin/env python3
"""Time sylow_b200_pairing_batch_dev for several batch sizes and slice sizes (SYLOW_B200_PAIR_CHUNK; -1 = no slicing).
One subprocess per setting (the library reads the variable once)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import numpy as np, torch, sylow_b200
    eng = sylow_b200.Engine(0)
    dev = torch.device("cuda", 0)
    out = {"chunk": os.environ.get("SYLOW_B200_PAIR_CHUNK", "default"), "tail_split": os.environ.get("SYLOW_B200_TAIL_SPLIT", "1")}
    for n in [1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 37888 * 2 + 5000, 113664, 113664 * 2 + 30000]:
        log2n = n.bit_length() - 1
        d_g1 = torch.randint(0, 255, (n, 64), dtype=torch.uint8, device=dev)
        d_g2 = torch.randint(0, 255, (n, 128), dtype=torch.uint8, device=dev)
        d_g1[:, 31::32] &= 0x1F
        d_g2[:, 31::32] &= 0x1F
        d_o = torch.empty((n, 384), dtype=torch.uint8, device=dev)
        eng.pairing_batch_dev(d_g1, d_g2, d_o)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3 if log2n >= 20 else 6
        a.record()
        for _ in range(reps):
            eng.pairing_batch_dev(d_g1, d_g2, d_o)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        out[str(n)] = round(n / ms / 1e3, 4)  # M pairings/s
    print(json.dumps(out))
else:
    for spec in sys.argv[1:] or ["-1:1", "-1:0"]:
        chunk, tail = spec.split(":")
        env = dict(os.environ, SYLOW_B200_PAIR_CHUNK=chunk, SYLOW_B200_TAIL_SPLIT=tail)
        r = subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-400:], flush=True)
