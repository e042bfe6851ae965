#!/usr/bin/env python3
"""Bucket MSM timing per window size at 2^16 .. 2^20 points (host buffers, wall clock) next to ladders + tree sum."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import sylow_b200  # noqa: E402

eng = sylow_b200.Engine(0)
rs = np.random.RandomState(5)
G1 = np.zeros((1, 64), np.uint8)
G1[0, 0], G1[0, 32] = 1, 2
for lg in (14, 16, 18, 20):
    n = 1 << lg
    k = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
    k[:, 31] &= 0x1F
    seeds, _ = eng.g1_mul_batch(np.repeat(G1, 4096, axis=0), k[:4096])
    P = np.tile(seeds, (n // 4096 + 1, 1))[:n].copy()

    def wall(fn):
        fn()
        t0 = time.perf_counter()
        r = fn()
        return (time.perf_counter() - t0) * 1e3, r

    t_l, ref = wall(lambda: eng.g1_sum(*eng.g1_mul_batch(P, k)))
    line = "n=2^%d ladders %.2f ms |" % (lg, t_l)
    for c in (0, lg - 10, lg - 8, lg - 6, lg - 4):
        if c != 0 and not 4 <= c <= 16:
            continue
        t_b, out = wall(lambda: eng.g1_msm_bucket(P, k, window_bits=c))
        assert (out[0] == ref[0]).all()
        line += " c=%d: %.2f ms" % (c, t_b)
    print(line, flush=True)
