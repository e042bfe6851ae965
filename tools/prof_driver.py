#!/usr/bin/env python3
"""One launch of every hot kernel at the bench sizes, for `ncu --set full` (tools/gpu_profile.sh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, sylow_b200

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
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
d_k1, d_k2 = torch.from_numpy(k[:n]).to(dev), torch.from_numpy(k[n:]).to(dev)
d_g1 = torch.empty((n, 64), dtype=torch.uint8, device=dev)
d_g2 = torch.empty((n, 128), dtype=torch.uint8, device=dev)
for _round in range(2):  # ncu skips the first round (-s 8): cold first launches profile badly
    # launches 1, 2 (k_g1_mul, k_g2_mul): the scalar-multiplication kernels on generator inputs
    eng.g1_mul_batch_dev(torch.from_numpy(g1).to(dev), d_k1, d_g1)
    eng.g2_mul_batch_dev(torch.from_numpy(g2).to(dev), d_k2, d_g2)
    d_f = torch.empty((n, 384), dtype=torch.uint8, device=dev)
    d_o = torch.empty((n, 384), dtype=torch.uint8, device=dev)
    eng.miller_loop_batch_dev(d_g1, d_g2, d_f)          # k_miller
    eng.final_exp_batch_dev(d_f, d_o)                   # k_final_exp
    d_msgs = torch.from_numpy(rs.randint(0, 256, size=n * 32, dtype=np.uint8)).to(dev)
    d_offs = torch.from_numpy((np.arange(n + 1, dtype=np.int64) * 32)).to(dev)
    d_h = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    eng.hash_to_g1_batch_dev(d_msgs, d_offs, d_h)       # k_hash_to_g1
    nc = n // 4
    co = eng.g2_precompute(d_g2[:3].cpu().numpy())
    d_tab = torch.empty(3 * 87 * 192, dtype=torch.uint8, device=dev)
    eng.tables_to_device(co, d_tab)
    d_ok = torch.empty(nc, dtype=torch.uint8, device=dev)
    eng.pairing_check_fixed_batch_dev(d_g1, d_g2[:nc].contiguous(), d_tab, 1, 3, d_ok)  # k_glued<1,3>, k_check_products
    # verify_batch's Miller stage: k_hash_to_g1 again, then k_glued<4,0> (four signatures per thread) on (-H(m_i), pk_i);
    # the signatures here are not valid ones - the kernels are branch-free in the data
    d_part = torch.empty(384, dtype=torch.uint8, device=dev)
    eng.verify_batch_partial_dev(d_g2, d_msgs, d_offs, d_g1, d_part)
    torch.cuda.synchronize()
print("prof_driver done", n)
