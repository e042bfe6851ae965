import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sylow_b200
from oracle import c_oracle as c
eng = sylow_b200.Engine(0)
for n in (1 << 12, 1 << 16, 1 << 20):
    rs = np.random.RandomState(102)
    sks = rs.randint(0, 256, size=(n, 32), dtype=np.uint8); sks[:, 31] &= 0x1F
    msgs = np.zeros((n, 32), np.uint8)
    msgs[:, :8] = np.arange(n, dtype=np.uint64).view(np.uint8).reshape(n, 8)
    msgs[:, 8:] = rs.randint(0, 256, size=(n, 24), dtype=np.uint8)
    offs = np.arange(n + 1, dtype=np.uint64) * 32
    sigs = eng.sign_batch(sks, (msgs.reshape(-1), offs))
    hm, hinf = eng.hash_to_g1_batch((msgs.reshape(-1), offs))
    idx = np.unique(np.concatenate([np.arange(64), rs.randint(0, n, size=1024), np.arange(n - 64, n)]))
    sub = (msgs[idx].reshape(-1), np.arange(len(idx) + 1, dtype=np.uint64) * 32)
    ref_s = c.sign_batch(sks[idx], sub)
    ref_h, _ = c.hash_to_g1_batch(sub)
    bs = np.where((sigs[idx] != ref_s).any(axis=1))[0]
    bh = np.where((hm[idx] != ref_h).any(axis=1))[0]
    mul, _ = eng.g1_mul_batch(hm[idx], sks[idx])
    bm = np.where((mul != ref_s).any(axis=1))[0]
    print("n", n, "bad sigs", len(bs), idx[bs][:10], "bad hash", len(bh), idx[bh][:10], "bad mul-of-gpu-hash", len(bm), "hinf", int(hinf.sum()))
    if len(bm):
        j = bm[0]
        print(" scalar", bytes(sks[idx][j]).hex(), " pt", bytes(hm[idx][j]).hex())
