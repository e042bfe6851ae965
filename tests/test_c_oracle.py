"""CPU tier: pin the C oracle (oracle/sylow_oracle.c, the literal multi-threaded restatement used as the
CPU baseline and as the fast oracle of the big GPU tests) against the reference's golden vectors and
against the Python oracle."""
import random

import numpy as np
import pytest

from oracle import bn254_py as o
from oracle import c_oracle as c
from tests import wire as w

H = lambda s: int(s, 16)


def arr(bs):
    return np.frombuffer(b"".join(bs), dtype=np.uint8).reshape(len(bs), -1).copy()


@pytest.fixture(scope="module", autouse=True)
def _built():
    c.build()


def test_constants(kats):
    buf = c.constants().tobytes()
    f = lambda off: w.b_fp(buf[off:off + 32])
    k = kats["svdw_constants"]
    assert [f(32 * i) for i in range(5)] == [H(k[n]) for n in ("z", "c1", "c2", "c3", "c4")]
    off = 160
    for name, cnt in (("frobenius_coeff_fp6_c1", 6), ("frobenius_coeff_fp6_c2", 6), ("frobenius_coeff_fp12_c1", 12)):
        for i in range(cnt):
            assert [f(off), f(off + 32)] == [H(x) for x in kats[name]["table"][i]], (name, i)
            off += 64
    assert [f(off), f(off + 32)] == [H(x) for x in kats["fp2_twist_curve_constant"]["value"]]
    assert [f(off + 64), f(off + 96)] == [H(x) for x in kats["eps_exp0"]["value"]]
    assert [f(off + 128), f(off + 160)] == [H(x) for x in kats["eps_exp1"]["value"]]


def test_gt_generator_and_test_cases(kats):
    out = c.pairing_batch(arr([w.g1_b(o.G1_GEN)]), arr([w.g2_b(o.G2_GEN)]))
    assert [w.b_fp(bytes(out[0][32 * i:32 * i + 32])) for i in range(12)] == [H(x) for x in kats["gt_generator"]["fp12"]]
    t = kats["pairing_test_cases"]
    p, _ = c.g1_mul_batch(arr([w.g1_b(o.G1_GEN)]), arr([w.fp_b(H(t["g1_scalar"]))]))
    q, _ = c.g2_mul_batch(arr([w.g2_b(o.G2_GEN)]), arr([w.fp_b(H(t["g2_scalar"]))]))
    out = c.pairing_batch(p, q)
    assert [w.b_fp(bytes(out[0][32 * i:32 * i + 32])) for i in range(12)] == [H(x) for x in t["fp12"]]


def test_identities():
    g1 = arr([w.g1_b((0, 1)), w.g1_b(o.G1_GEN)])
    g2 = arr([w.g2_b(o.G2_GEN), w.g2_b((o.FP2_ZERO, o.FP2_ONE))])
    out = c.pairing_batch(g1, g2, g1_inf=[1, 0], g2_inf=[0, 1])
    assert w.b_fp12(bytes(out[0])) == o.FP12_ONE and w.b_fp12(bytes(out[1])) == o.FP12_ONE


def test_eip196_197(kats):
    inp = bytes.fromhex(kats["eip196_mul"]["input"])
    x, y, k = (int.from_bytes(inp[32 * i: 32 * i + 32], "big") for i in range(3))
    out, inf = c.g1_mul_batch(arr([w.g1_b((x, y))]), arr([w.fp_b(k)]))
    exp = bytes.fromhex(kats["eip196_mul"]["expected"])
    assert w.b_g1(bytes(out[0]))[:2] == (int.from_bytes(exp[:32], "big"), int.from_bytes(exp[32:], "big"))
    inp = bytes.fromhex(kats["eip197_pair"]["input"])
    g1s, g2s = [], []
    for off in range(0, len(inp), 192):
        cc = [int.from_bytes(inp[off + 32 * i: off + 32 * i + 32], "big") for i in range(6)]
        g1s.append((cc[0], cc[1]))
        g2s.append(((cc[3], cc[2]), (cc[5], cc[4])))
    prod = c.miller_product(arr([w.g1_b(p) for p in g1s]), arr([w.g2_b(q) for q in g2s]), threads=2)
    assert w.b_fp12(bytes(c.final_exp_batch(prod.reshape(1, 384))[0])) == o.FP12_ONE


def test_against_python_oracle():
    rng = random.Random(21)
    n = 6
    ps = [w.rand_g1(rng) for _ in range(n)]
    qs = [w.rand_g2(rng) for _ in range(n)]
    G1, G2 = arr([w.g1_b(p) for p in ps]), arr([w.g2_b(q) for q in qs])
    fs = [o.miller_loop(o.g2_precompute(q), p) for p, q in zip(ps, qs)]
    assert [w.b_fp12(bytes(r)) for r in c.miller_loop_batch(G1, G2, threads=3)] == fs
    assert [w.b_fp12(bytes(r)) for r in c.pairing_batch(G1, G2)] == [o.final_exponentiation(f) for f in fs]
    ks = [0, 1, o.R_ORDER - 1, o.P - 1] + [rng.randrange(o.P) for _ in range(n - 4)]
    out, inf = c.g1_mul_batch(G1, arr([w.fp_b(k) for k in ks]))
    assert [w.b_g1(bytes(r), i) for r, i in zip(out, inf)] == [
        o.proj_to_affine(o.FpOps, o.proj_mul(o.FpOps, o.affine_to_proj(o.FpOps, p), k)) for p, k in zip(ps, ks)]
    out, inf = c.g2_mul_batch(G2, arr([w.fp_b(k) for k in ks]))
    assert [w.b_g2(bytes(r), i) for r, i in zip(out, inf)] == [
        o.proj_to_affine(o.Fp2Ops, o.proj_mul(o.Fp2Ops, o.affine_to_proj(o.Fp2Ops, q), k)) for q, k in zip(qs, ks)]


def test_hash_sign_verify():
    rng = random.Random(22)
    msgs = [b"", b"abc", (20).to_bytes(4, "big"), bytes(range(135)), bytes(range(136)), bytes(range(200)) * 3]
    out, inf = c.hash_to_g1_batch(msgs)
    assert not inf.any()
    assert [w.b_g1(bytes(r)) for r in out] == [o.proj_to_affine(o.FpOps, o.hash_to_curve_g1(m)) for m in msgs]
    sks = [rng.randrange(1, o.R_ORDER) for _ in msgs]
    SK = arr([w.fp_b(s) for s in sks])
    sigs = c.sign_batch(SK, msgs)
    assert [w.b_g1(bytes(r)) for r in sigs] == [o.proj_to_affine(o.FpOps, o.sign(s, m)) for s, m in zip(sks, msgs)]
    pks, _ = c.g2_mul_batch(arr([w.g2_b(o.G2_GEN)] * len(msgs)), SK)
    assert c.verify_each(pks, msgs, sigs).all()
    assert c.verify_batch(pks, msgs, sigs)
    bad = sigs.copy()
    bad[1] = sigs[0]
    assert c.verify_each(pks, msgs, bad).tolist() == [True, False] + [True] * (len(msgs) - 2)
    assert not c.verify_batch(pks, msgs, bad)
    # the same batch through glued_miller_loop (shared squarings: the form the reference's example runs, and the
    # timed CPU arm of bench.py's verify_batch leg); groups of 16 signatures, ragged tail, several threads
    for t in (1, 3):
        assert c.verify_batch_glued(pks, msgs, sigs, threads=t)
        assert not c.verify_batch_glued(pks, msgs, bad, threads=t)
    pk40, m40, s40 = c._sample_signatures(40)
    assert c.verify_batch_glued(pk40, m40, s40, threads=2) and c.verify_batch(pk40, m40, s40)
    s40[39] = s40[0]
    assert not c.verify_batch_glued(pk40, m40, s40, threads=2)
