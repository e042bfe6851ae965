import sys, numpy as np
sys.path.insert(0, ".")
import sylow_b200
eng = sylow_b200.Engine(0)
rs = np.random.RandomState(5)
G1 = np.zeros((1, 64), np.uint8); G1[0, 0], G1[0, 32] = 1, 2
n = 1 << 20
k = rs.randint(0, 256, size=(n, 32), dtype=np.uint8); k[:, 31] &= 0x1F
seeds, _ = eng.g1_mul_batch(np.repeat(G1, 4096, axis=0), k[:4096])
P = np.tile(seeds, (n // 4096 + 1, 1))[:n].copy()
eng.g1_msm_bucket(P, k, window_bits=10)
