import sys, time, numpy as np
sys.path.insert(0, '.')
import sylow_b200
from tests import wire as w
from oracle import bn254_py as o
eng = sylow_b200.Engine(0)
n = 1 << 16
g = eng.pairing_batch(np.frombuffer(w.g1_b(o.G1_GEN), dtype=np.uint8).reshape(1, 64), np.frombuffer(w.g2_b(o.G2_GEN), dtype=np.uint8).reshape(1, 128))
G = np.repeat(g, n, axis=0)
rs = np.random.RandomState(3)
k = rs.randint(0, 256, size=(n, 32), dtype=np.uint8); k[:, 31] &= 0x1F
eng.gt_mul_batch(G, k)
t0 = time.perf_counter(); out = eng.gt_mul_batch(G, k); dt = time.perf_counter() - t0
print("gt_mul: %.2f M/s (host buffers, n = 2^16)" % (n / dt / 1e6))
