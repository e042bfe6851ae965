"""GPU tier at BASELINE.json's full sizes, through size-independent properties (the oracle cannot run 2^20
pairings in seconds): a prefix is compared bit-for-bit with the C oracle and the WHOLE batch is pinned by a
checksum of checksums - bilinearity turns the product of all outputs into one exponentiation of
Gt::generator() that the oracle can check:  prod_i e(a_i G1, b_i G2) = GT^(sum a_i b_i mod r)."""
import numpy as np
import pytest
import torch

from oracle import bn254_py as o
from oracle import c_oracle as c
from tests import wire as w

pytestmark = pytest.mark.gpu
R = o.R_ORDER


@pytest.fixture(scope="module")
def eng():
    import sylow_b200

    e = sylow_b200.Engine(0)
    yield e
    e.close()


def _scalars(rs, n):
    k = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
    k[:, 31] &= 0x1F  # < 2^253 < r
    return k


def _ints(k):
    """(n, 32) little-endian bytes -> python ints, vectorised through 4 u64 words."""
    w64 = k.view("<u8").reshape(-1, 4)
    return [int(a) | (int(b) << 64) | (int(c_) << 128) | (int(d) << 192) for a, b, c_, d in w64]


def _gens(n):
    g1 = np.zeros((n, 64), np.uint8)
    g1[:, 0], g1[:, 32] = 1, 2
    g2 = np.tile(np.frombuffer(w.g2_b(o.G2_GEN), dtype=np.uint8), (n, 1))
    return g1, g2


def _gt_pow_gen(eng, e):
    gt = np.frombuffer(w.fp12_b(o.pairing_affine(o.G1_GEN, o.G2_GEN)), dtype=np.uint8).reshape(1, 384)
    return eng.gt_mul_batch(gt, np.frombuffer(w.fp_b(e % R), dtype=np.uint8).reshape(1, 32))[0]


def test_pairing_batch_2pow20(eng):
    """BASELINE configs[1]: 2^20 independent pairings on random G1 x G2 points."""
    n = 1 << 20
    rs = np.random.RandomState(101)
    a, b = _scalars(rs, n), _scalars(rs, n)
    g1, g2 = _gens(n)
    P, pinf = eng.g1_mul_batch(g1, a)
    Q, qinf = eng.g2_mul_batch(g2, b)
    assert not pinf.any() and not qinf.any()
    # a few infinite inputs on both sides of the host path's chunk boundaries (whole waves of both kernels = SMs * 1536 pairs per chunk):
    # pairing(inf, .) = pairing(., inf) = identity (pairing.rs:876-886)
    chunk = torch.cuda.get_device_properties(0).multi_processor_count * 1536
    holes = [chunk - 1, chunk, 3 * chunk + 5, n - 1]
    pinf[holes[:2]] = 1
    qinf[holes[2:]] = 1
    gt = eng.pairing_batch(P, Q, g1_inf=pinf, g2_inf=qinf)
    one = np.frombuffer(w.fp12_b(o.FP12_ONE), dtype=np.uint8)
    assert all((gt[h] == one).all() for h in holes)
    a, b = a.copy(), b.copy()
    a[holes] = 0  # they drop out of the checksum below
    # (1) prefix and a scattered sample, bit-exact against the CPU oracle
    idx = np.concatenate([np.arange(512), rs.randint(0, n, size=512)])
    idx = idx[~np.isin(idx, holes)]
    assert (gt[idx] == c.pairing_batch(P[idx], Q[idx])).all()
    assert (P[idx] == c.g1_mul_batch(g1[idx], a[idx])[0]).all() and (Q[idx] == c.g2_mul_batch(g2[idx], b[idx])[0]).all()
    # (2) checksum of checksums over all 2^20 outputs
    e = sum(x * y for x, y in zip(_ints(a), _ints(b))) % R
    assert (eng.fp12_product(gt) == _gt_pow_gen(eng, e)).all()
    # the same identity through the Miller-product path (glued_miller_loop + one final exponentiation)
    assert (eng.final_exp_batch(eng.miller_product(P, Q, g1_inf=pinf, g2_inf=qinf).reshape(1, 384))[0]
            == _gt_pow_gen(eng, e)).all()
    # oracle cross-check of the right-hand side
    assert w.b_fp12(bytes(_gt_pow_gen(eng, e))) == o.gt_mul(o.pairing_affine(o.G1_GEN, o.G2_GEN), e)


def test_verify_batch_2pow20(eng):
    """BASELINE configs[2]: verify_batch over 2^20 distinct-message, distinct-signer signatures."""
    n = 1 << 20
    rs = np.random.RandomState(102)
    sks = _scalars(rs, n)
    msgs = np.zeros((n, 32), np.uint8)
    msgs[:, :8] = np.arange(n, dtype=np.uint64).view(np.uint8).reshape(n, 8)
    msgs[:, 8:] = rs.randint(0, 256, size=(n, 24), dtype=np.uint8)
    offs = np.arange(n + 1, dtype=np.uint64) * 32
    packed = (msgs.reshape(-1), offs)
    sigs = eng.sign_batch(sks, packed)
    _, g2 = _gens(n)
    pks, _ = eng.g2_mul_batch(g2, sks)
    idx = rs.randint(0, n, size=256)
    sub = (msgs[idx].reshape(-1), np.arange(257, dtype=np.uint64) * 32)
    assert (sigs[idx] == c.sign_batch(sks[idx], sub)).all()  # signatures bit-exact on a sample
    assert eng.verify_batch(pks, packed, sigs) is True
    bad = sigs.copy()
    bad[n // 3] = sigs[n // 3 + 1]
    assert eng.verify_batch(pks, packed, bad) is False  # one wrong signature among 2^20 flips the verdict
    sl = slice(n // 3 - 100, n // 3 + 100)
    each = eng.verify_each(pks[sl], (msgs[sl].reshape(-1), np.arange(201, dtype=np.uint64) * 32), bad[sl])
    assert each.tolist() == [i != 100 for i in range(200)]


def test_groth16_shape_2pow18(eng):
    """BASELINE configs[3]: 2^18 four-pair product checks e(A,B) e(-alpha,beta) e(-L,gamma) e(-C,delta) == 1."""
    nc = 1 << 18
    rs = np.random.RandomState(103)
    be, ga, de = (int(x) for x in rs.randint(2, 1 << 62, size=3))
    a, b, al, l = (_scalars(rs, nc) for _ in range(4))
    ai, bi, ali, li = (_ints(x) for x in (a, b, al, l))
    de_inv = pow(de, -1, R)
    ci = [((x * y - u * be - v * ga) * de_inv) % R for x, y, u, v in zip(ai, bi, ali, li)]
    for t in (5, nc - 7):
        ci[t] = (ci[t] + 1) % R  # two invalid proofs
    neg = lambda v: [(R - x) % R for x in v]
    sc = np.zeros((nc, 4, 32), np.uint8)
    for j, col in enumerate((ai, neg(ali), neg(li), neg(ci))):
        sc[:, j, :] = np.frombuffer(b"".join(x.to_bytes(32, "little") for x in col), dtype=np.uint8).reshape(nc, 32)
    g1, g2 = _gens(4 * nc)
    G1, _ = eng.g1_mul_batch(g1, sc.reshape(-1, 32))
    Bp, _ = eng.g2_mul_batch(g2[:nc], b)
    fixed, _ = eng.g2_mul_batch(g2[:3], np.frombuffer(b"".join(x.to_bytes(32, "little") for x in (be, ga, de)),
                                                      dtype=np.uint8).reshape(3, 32))
    ok = eng.pairing_check_fixed_batch(G1, Bp, eng.g2_precompute(fixed), 1, 3)
    expect = np.ones(nc, bool)
    expect[[5, nc - 7]] = False
    assert (ok == expect).all()
    # a slice through the general (no precomputation) path and the oracle agree
    G2all = np.concatenate([Bp[:8, None, :], np.tile(fixed[None], (8, 1, 1))], axis=1).reshape(-1, 128)
    assert eng.pairing_check_batch(G1[:32], G2all, 4).tolist() == expect[:8].tolist()
    prod = c.miller_product(G1[20:24], G2all[20:24])
    assert (c.final_exp_batch(prod.reshape(1, 384))[0] == np.frombuffer(w.fp12_b(o.FP12_ONE), dtype=np.uint8)).all() == bool(expect[5])


def test_scalar_mul_2pow22(eng):
    """BASELINE configs[4]: 2^22 variable-base scalar multiplications in G1 and G2 (254-bit scalars)."""
    n = 1 << 22
    rs = np.random.RandomState(104)
    k = _scalars(rs, n)
    base_s = _scalars(rs, n)
    g1, g2 = _gens(n)
    # variable bases: B_i = s_i * G; outputs k_i * B_i = (k_i s_i) G
    B1, _ = eng.g1_mul_batch(g1, base_s)
    out1, inf1 = eng.g1_mul_batch(B1, k)
    assert not inf1.any()
    idx = rs.randint(0, n, size=384)
    assert (out1[idx] == c.g1_mul_batch(B1[idx], k[idx])[0]).all()
    ki, si = _ints(k), _ints(base_s)
    e = sum(x * y for x, y in zip(ki, si)) % R
    # checksum: prod_i e(k_i B_i, G2) = GT^(sum k_i s_i)
    f = eng.miller_product(out1, g2)
    assert (eng.final_exp_batch(f.reshape(1, 384))[0] == _gt_pow_gen(eng, e)).all()
    del out1, B1
    m = 1 << 21  # G2 (512 MiB of points at 2^22): two halves to bound host memory
    tot = 0
    for h in range(2):
        sl = slice(h * m, (h + 1) * m)
        B2, _ = eng.g2_mul_batch(g2[sl], base_s[sl])
        out2, inf2 = eng.g2_mul_batch(B2, k[sl])
        assert not inf2.any()
        j = rs.randint(0, m, size=96)
        assert (out2[j] == c.g2_mul_batch(B2[j], k[sl][j])[0]).all()
        f = eng.miller_product(g1[sl], out2)
        eh = sum(x * y for x, y in zip(ki[sl], si[sl])) % R
        assert (eng.final_exp_batch(f.reshape(1, 384))[0] == _gt_pow_gen(eng, eh)).all()
        tot += eh
    assert tot % R == e


def test_msm_2pow18(eng):
    """One 2^18-point multi-scalar multiplication (bucket method from 2^17 points on): with P_i = a_i G the sum
    sum k_i P_i is (sum k_i a_i mod r) G, a single scalar multiple the oracle can check."""
    n = 1 << 18
    rs = np.random.RandomState(105)
    a, k = _scalars(rs, n), _scalars(rs, n)
    k[:5] = 0                      # zero scalars drop out
    k[5:9] = 255                   # 2^256 - 1: reduced mod r by the group order
    g1, _ = _gens(n)
    P, pinf = eng.g1_mul_batch(g1, a)
    assert not pinf.any()
    out, inf = eng.g1_msm(P, k)
    e = sum(x * y for x, y in zip(_ints(a), _ints(k))) % R
    want = o.proj_to_affine(o.FpOps, o.proj_mul(o.FpOps, o.affine_to_proj(o.FpOps, o.G1_GEN), e))
    assert w.b_g1(bytes(out), inf) == want
    out2, inf2 = eng.g1_msm_bucket(P, k, window_bits=13)
    assert w.b_g1(bytes(out2), inf2) == want
