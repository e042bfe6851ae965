#!/usr/bin/env python3
"""verify_batch_partial_dev at 2^20 signatures with 1 / 2 / 4 signatures per thread in the Miller stage
(SYLOW_B200_VERIFY_GLUE); one subprocess per setting.  Prints ms per batch, the verdict of a valid batch and of one with a
replaced signature."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import numpy as np, torch, sylow_b200
    n = 1 << int(sys.argv[2])
    eng = sylow_b200.Engine(0)
    dev = torch.device("cuda", 0)
    rs = np.random.RandomState(3)
    sks = rs.randint(0, 256, size=(n, 32), dtype=np.uint8); sks[:, 31] &= 0x1F
    msgs = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
    offs = np.arange(n + 1, dtype=np.uint64) * 32
    sigs = eng.sign_batch(sks, (msgs.reshape(-1), offs))
    G2 = (10857046999023057135944570762232829481370756359578518086990519993285655852781,
          11559732032986387107991004021392285783925812861821192530917403151452391805634,
          8495653923123431417604973247489272438418190587263600148770280649306958101930,
          4082367875863433681332203403145435568316851327593401208105741076214120093531)
    g2 = np.tile(np.frombuffer(b"".join(c.to_bytes(32, "little") for c in G2), dtype=np.uint8), (n, 1))
    d_pk = torch.empty((n, 128), dtype=torch.uint8, device=dev)
    eng.g2_mul_batch_dev(torch.from_numpy(g2).to(dev), torch.from_numpy(sks).to(dev), d_pk)
    d_m, d_o = torch.from_numpy(msgs.reshape(-1)).to(dev), torch.from_numpy(offs.view(np.int64)).to(dev)
    d_s = torch.from_numpy(sigs).to(dev)
    d_f = torch.empty(384, dtype=torch.uint8, device=dev)
    def run(ds, seed=None):
        eng.verify_batch_partial_dev(d_pk, d_m, d_o, ds, d_f, weight_seed=seed)
        return eng.verify_batch_finish(d_f.cpu().numpy().reshape(1, 384))
    ok = run(d_s)
    bad = sigs.copy(); bad[n - 1] = bad[0]
    ok_bad = run(torch.from_numpy(bad).to(dev))
    okw = run(d_s, bytes(range(32)))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        eng.verify_batch_partial_dev(d_pk, d_m, d_o, d_s, d_f)
    b.record(); torch.cuda.synchronize()
    print(json.dumps({"glue": os.environ.get("SYLOW_B200_VERIFY_GLUE"), "n": n, "ms": a.elapsed_time(b) / 3, "valid": ok,
                      "corrupt": ok_bad, "valid_weighted": okw}))
else:
    for g in ("1", "2", "4"):
        for log2n in ("20", "13"):
            r = subprocess.run([sys.executable, __file__, "--one", log2n], env=dict(os.environ, SYLOW_B200_VERIFY_GLUE=g),
                               capture_output=True, text=True)
            print(r.stdout.strip() or r.stderr[-400:], flush=True)
