"""GPU tier (-m gpu), round-2 additions: binary-GCD inversion / Jacobi symbol, infinity flags on the BLS entry points,
random-weight batch verification, the in-library multi-device context, the chunk-interleaved device pairing path.
Everything through the C ABI, bit-exact against the oracle."""
import random

import numpy as np
import pytest

from oracle import bn254_py as o
from tests import wire as w

pytestmark = pytest.mark.gpu

MSGS = [b"", b"a", b"sylow", bytes(range(40)), b"x" * 200, b"\x00\x00\x00\x14", b"msg-6", b"msg-7", b"msg-8"]


@pytest.fixture(scope="module")
def eng():
    import sylow_b200

    e = sylow_b200.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def multi():
    import sylow_b200

    e = sylow_b200.Engine([0, 0, 0])  # three device slots on GPU 0: the slicing logic is what is under test
    yield e
    e.close()


def arr(bs):
    return np.frombuffer(b"".join(bs), dtype=np.uint8).reshape(len(bs), -1).copy()


def ints(a):
    return [w.b_fp(bytes(r)) for r in a]


def _keys(rng, n):
    sks = [rng.randrange(1, o.R_ORDER) for _ in range(n)]
    SK = arr([w.fp_b(s) for s in sks])
    return sks, SK


# ------------------------------------------------------------------------------------------ Fp: binary GCD
def test_gcd_inversion_and_jacobi(eng):
    """fp_inv (binary GCD, fp.cuh) against the oracle and against the Fermat ladder it replaced (fp.rs:418-424,
    inv(0) = 0); fp_jacobi against x^((p-1)/2) (fp.rs:625-631) - op 7 reports (x/p) + 1, plus 4 on a disagreement."""
    rng = random.Random(201)
    P = o.P
    edge = [0, 1, 2, 3, 4, P - 1, P - 2, (P + 1) // 2, (P - 1) // 2, 1 << 253, (1 << 253) - 1, (1 << 128), (1 << 128) - 1,
            (1 << 32) - 1, 1 << 32, 1 << 64, 3 ** 100 % P, pow(2, -1, P), pow(3, -1, P)]
    a = edge + [rng.randrange(P) for _ in range(5000)] + [pow(2, k, P) for k in range(0, 254, 7)]
    A = arr([w.fp_b(x) for x in a])
    Z = np.zeros_like(A)
    inv = ints(eng.fp_op_batch(3, A, Z))
    assert inv == [pow(x, P - 2, P) for x in a]
    assert inv == ints(eng.fp_op_batch(6, A, Z))
    jac = [int.from_bytes(bytes(r), "little") for r in eng.fp_op_batch(7, A, Z)]
    exp = [{0: 1, 1: 2, P - 1: 0}[pow(x, (P - 1) // 2, P)] for x in a]
    assert jac == exp


# ------------------------------------------------------------------------------------------ BLS: infinity flags
def test_verify_infinity_flags(eng):
    """pairing() maps an infinite input to the identity (pairing.rs:876-886), so verify(pk, m, sig) =
    (e(sig, G2) == e(H(m), pk)) is true for (inf, inf), false when exactly one side is infinite - in the library,
    not in the caller."""
    rng = random.Random(202)
    n = 6
    sks, SK = _keys(rng, n)
    msgs = MSGS[:n]
    sigs = eng.sign_batch(SK, msgs)
    pks, _ = eng.g2_mul_batch(arr([w.g2_b(o.G2_GEN)] * n), SK)
    pk_inf = np.array([0, 1, 0, 1, 0, 0], np.uint8)
    sg_inf = np.array([0, 0, 1, 1, 0, 0], np.uint8)
    ok = eng.verify_each(pks, msgs, sigs, pks_inf=pk_inf, sigs_inf=sg_inf)
    assert ok.tolist() == [True, False, False, True, True, True]
    # the product form skips infinite factors the same way: items 0, 3 (both infinite), 4, 5 are consistent
    keep = [0, 3, 4, 5]
    assert eng.verify_batch(pks[keep], [msgs[i] for i in keep], sigs[keep], pks_inf=pk_inf[keep],
                            sigs_inf=sg_inf[keep]) is True
    assert eng.verify_batch(pks, msgs, sigs, pks_inf=pk_inf, sigs_inf=sg_inf) is False
    # sk = 0 mod r signs to the identity, and the flag says so
    out, inf = eng.sign_batch(arr([w.fp_b(0), w.fp_b(o.R_ORDER), w.fp_b(5)]), msgs[:3], return_inf=True)
    assert inf.tolist() == [1, 1, 0]
    # the reference-API mirror passes the flags through
    from sylow_b200 import sylow as sy

    pk = sy.G2Affine.generator() * sks[0]
    sig = sy.sign(sks[0], msgs[0])
    assert sy.verify(pk, msgs[0], sig) is True
    assert sy.verify(sy.G2Affine.zero(), msgs[0], sy.G1Affine.zero()) is True
    assert sy.verify(sy.G2Affine.zero(), msgs[0], sig) is False
    assert sy.sign(0, msgs[0]).infinity is True


# ------------------------------------------------------------------------------------------ BLS: random weights
def _weights(seed: bytes, first: int, n: int):
    return [int.from_bytes(o.keccak256(seed + (first + i).to_bytes(8, "little"))[:8], "little") | 1 for i in range(n)]


def test_weighted_batch_verification(eng):
    """the unweighted product is aggregate verification - (sig_1 + D, sig_2 - D) passes.  With a weight seed
    the check is prod e(r_i sig_i, G2) e(-r_i H_i, pk_i) == 1 and that forgery fails."""
    rng = random.Random(203)
    n = 7
    sks, SK = _keys(rng, n)
    msgs = MSGS[:n]
    seed = bytes(rng.randrange(256) for _ in range(32))
    assert eng.batch_weights(seed, 5, 9).tolist() == _weights(seed, 5, 9)
    sigs = eng.sign_batch(SK, msgs)
    pks, _ = eng.g2_mul_batch(arr([w.g2_b(o.G2_GEN)] * n), SK)
    assert eng.verify_batch(pks, msgs, sigs, weight_seed=seed) is True
    assert eng.verify_batch_same_signer(pks[0], msgs, eng.sign_batch(arr([w.fp_b(sks[0])] * n), msgs), weight_seed=seed) is True
    # cancelling forgery
    D = w.rand_g1(rng)
    F = o.FpOps
    s1 = o.proj_to_affine(F, o.proj_add(F, o.affine_to_proj(F, w.b_g1(bytes(sigs[1]))), o.affine_to_proj(F, D)))
    s2 = o.proj_to_affine(F, o.proj_add(F, o.affine_to_proj(F, w.b_g1(bytes(sigs[2]))), o.proj_neg(F, o.affine_to_proj(F, D))))
    forged = sigs.copy()
    forged[1] = np.frombuffer(w.g1_b(s1), np.uint8)
    forged[2] = np.frombuffer(w.g1_b(s2), np.uint8)
    assert eng.verify_each(pks, msgs, forged).tolist() == [True, False, False] + [True] * (n - 3)
    assert eng.verify_batch(pks, msgs, forged) is True            # the reference example's semantics
    assert eng.verify_batch(pks, msgs, forged, weight_seed=seed) is False
    # the weighted partial of a slice uses the global index: two slices combine to the whole-batch verdict, and the
    # partial itself is the oracle's Miller product of the weighted pairs
    pa = eng.verify_batch_partial(pks[:3], msgs[:3], sigs[:3], weight_seed=seed, first_index=0)
    pb = eng.verify_batch_partial(pks[3:], msgs[3:], sigs[3:], weight_seed=seed, first_index=3)
    assert eng.verify_batch_finish(np.stack([pa, pb])) is True
    pb_wrong = eng.verify_batch_partial(pks[3:], msgs[3:], forged[3:], weight_seed=seed, first_index=0)
    assert eng.verify_batch_finish(np.stack([pa, pb_wrong])) is True  # valid signatures verify under any weights
    r = _weights(seed, 0, 3)
    g1s, g2s = [], []
    wsum = None
    for i in range(3):
        hm = o.proj_mul(F, o.hash_to_curve_g1(msgs[i]), r[i])
        g1s.append(o.g1_affine_neg(o.proj_to_affine(F, hm)))
        g2s.append(w.b_g2(bytes(pks[i])))
        t = o.proj_mul(F, o.affine_to_proj(F, w.b_g1(bytes(sigs[i]))), r[i])
        wsum = t if wsum is None else o.proj_add(F, wsum, t)
    g1s.append(o.proj_to_affine(F, wsum))
    g2s.append(o.G2_GEN)
    assert w.b_fp12(bytes(pa)) == o.glued_miller_loop([o.g2_precompute(q) for q in g2s], g1s)


# ------------------------------------------------------------------------------------------ multi-device context
def test_multi_device_context_matches_single(eng, multi):
    """sylow_b200_create_multi: contiguous slices, one host thread per device slot, partial products combined on the
    first device - every result bit-identical to the single-device context, ragged sizes included."""
    assert multi.device_count == 3 and eng.device_count == 1
    rng = random.Random(204)
    for n in (0, 1, 2, 7, 50):
        ks = [rng.randrange(o.P) for _ in range(2 * n)]
        K1, K2 = arr([w.fp_b(k) for k in ks[:n]]) if n else np.zeros((0, 32), np.uint8), \
            arr([w.fp_b(k) for k in ks[n:]]) if n else np.zeros((0, 32), np.uint8)
        G1 = arr([w.g1_b(o.G1_GEN)] * n) if n else np.zeros((0, 64), np.uint8)
        G2 = arr([w.g2_b(o.G2_GEN)] * n) if n else np.zeros((0, 128), np.uint8)
        p1, i1 = eng.g1_mul_batch(G1, K1)
        p2, i2 = multi.g1_mul_batch(G1, K1)
        assert (p1 == p2).all() and (i1 == i2).all()
        q1, j1 = eng.g2_mul_batch(G2, K2)
        q2, j2 = multi.g2_mul_batch(G2, K2)
        assert (q1 == q2).all() and (j1 == j2).all()
        inf1 = np.zeros(n, np.uint8)
        if n > 3:
            inf1[3] = 1
        assert (eng.pairing_batch(p1, q1, g1_inf=inf1) == multi.pairing_batch(p1, q1, g1_inf=inf1)).all()
        assert (eng.miller_loop_batch(p1, q1) == multi.miller_loop_batch(p1, q1)).all()
        assert (eng.miller_product(p1, q1, g1_inf=inf1) == multi.miller_product(p1, q1, g1_inf=inf1)).all()
        msgs = [bytes(rng.randrange(256) for _ in range(rng.randrange(0, 70))) for _ in range(n)]
        h1, hi1 = eng.hash_to_g1_batch(msgs)
        h2, hi2 = multi.hash_to_g1_batch(msgs)
        assert (h1 == h2).all() and (hi1 == hi2).all()
        sk = arr([w.fp_b(k % o.R_ORDER or 1) for k in ks[:n]]) if n else np.zeros((0, 32), np.uint8)
        s1 = eng.sign_batch(sk, msgs)
        assert (s1 == multi.sign_batch(sk, msgs)).all()
        pk, _ = eng.g2_mul_batch(G2, sk)
        assert eng.verify_each(pk, msgs, s1).tolist() == multi.verify_each(pk, msgs, s1).tolist() == [True] * n
        seed = bytes(range(32))
        assert multi.verify_batch(pk, msgs, s1) is True
        assert multi.verify_batch(pk, msgs, s1, weight_seed=seed) is True
        if n >= 2:
            bad = s1.copy()
            bad[n - 1] = s1[0]  # the corrupted signature sits in the LAST slice
            assert multi.verify_batch(pk, msgs, bad) is False
            assert multi.verify_batch(pk, msgs, bad, weight_seed=seed) is False
            assert multi.verify_each(pk, msgs, bad).tolist() == [True] * (n - 1) + [False]
            # the partials themselves differ (miller(A + B) != miller(A) miller(B) before the final exponentiation);
            # their final exponentiations agree
            pa = eng.verify_batch_partial(pk, msgs, s1, weight_seed=seed).reshape(1, 384)
            pb = multi.verify_batch_partial(pk, msgs, s1, weight_seed=seed).reshape(1, 384)
            assert (eng.final_exp_batch(pa) == eng.final_exp_batch(pb)).all()
    # pairing checks: slices over checks, k pairs each
    n, k = 10, 2
    a = [rng.randrange(1, o.R_ORDER) for _ in range(n)]
    A = arr([w.fp_b(x) for x in a])
    P1, _ = eng.g1_mul_batch(arr([w.g1_b(o.G1_GEN)] * n), A)
    NEG = arr([w.fp_b((-x) % o.R_ORDER) for x in a])
    P2, _ = eng.g1_mul_batch(arr([w.g1_b(o.G1_GEN)] * n), NEG)
    Q, _ = eng.g2_mul_batch(arr([w.g2_b(o.G2_GEN)] * n), arr([w.fp_b(7)] * n))
    g1 = np.stack([P1, P2], axis=1).reshape(2 * n, 64)
    g2 = np.stack([Q, Q], axis=1).reshape(2 * n, 128)
    g1[2 * 4] = P1[5]  # break check 4
    exp = [True] * n
    exp[4] = False
    assert eng.pairing_check_batch(g1, g2, k).tolist() == multi.pairing_check_batch(g1, g2, k).tolist() == exp
    st1 = eng.g2_validate_batch(g2)
    assert (st1 == multi.g2_validate_batch(g2)).all() and (st1 == 0).all()


def test_dev_entry_points_reject_multi_context(multi):
    import torch

    d = torch.zeros((4, 384), dtype=torch.uint8, device="cuda:0")
    with pytest.raises(Exception):
        multi.final_exp_batch_dev(d, d)


# ------------------------------------------------------------------------------------------ device path interleave
def test_pairing_batch_dev_chunk_interleave(eng):
    """pairing_batch_dev cuts large batches into slices that alternate between two streams; the results must not depend
    on the slicing (and the host-buffer path, which slices differently, must agree)."""
    import torch

    n = 148 * 768 * 2 + 12345  # more than two slices, ragged tail
    dev = torch.device("cuda", 0)
    rs = np.random.RandomState(5)
    k = rs.randint(0, 256, size=(2 * n, 32), dtype=np.uint8)
    k[:, 31] &= 0x1F
    d_g1 = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    d_g2 = torch.empty((n, 128), dtype=torch.uint8, device=dev)
    g1 = np.tile(np.frombuffer(w.g1_b(o.G1_GEN), np.uint8), (n, 1))
    g2 = np.tile(np.frombuffer(w.g2_b(o.G2_GEN), np.uint8), (n, 1))
    eng.g1_mul_batch_dev(torch.from_numpy(g1).to(dev), torch.from_numpy(k[:n]).to(dev), d_g1)
    eng.g2_mul_batch_dev(torch.from_numpy(g2).to(dev), torch.from_numpy(k[n:]).to(dev), d_g2)
    d_gt = torch.empty((n, 384), dtype=torch.uint8, device=dev)
    eng.pairing_batch_dev(d_g1, d_g2, d_gt)
    d_f = torch.empty((n, 384), dtype=torch.uint8, device=dev)
    d_ref = torch.empty((n, 384), dtype=torch.uint8, device=dev)
    eng.miller_loop_batch_dev(d_g1, d_g2, d_f)
    eng.final_exp_batch_dev(d_f, d_ref)
    torch.cuda.synchronize()
    assert torch.equal(d_gt, d_ref)
    # sampled rows against the oracle
    rows = [0, 1, 148 * 768 - 1, 148 * 768, n - 1]
    H1, H2, HG = d_g1.cpu().numpy(), d_g2.cpu().numpy(), d_gt.cpu().numpy()
    for i in rows:
        assert w.b_fp12(bytes(HG[i])) == o.pairing_affine(w.b_g1(bytes(H1[i])), w.b_g2(bytes(H2[i])))


def test_hash_failed_dev_flag(eng):
    import torch

    dev = torch.device("cuda", 0)
    msgs = b"".join(MSGS)
    offs = np.cumsum([0] + [len(m) for m in MSGS]).astype(np.int64)
    d_m = torch.from_numpy(np.frombuffer(msgs, np.uint8).copy()).to(dev)
    d_o = torch.from_numpy(offs).to(dev)
    d_out = torch.empty((len(MSGS), 64), dtype=torch.uint8, device=dev)
    eng.hash_to_g1_batch_dev(d_m, d_o, d_out)
    assert eng.hash_failed_dev() is False
    exp = [o.proj_to_affine(o.FpOps, o.hash_to_curve_g1(m)) for m in MSGS]
    assert [w.b_g1(bytes(r)) for r in d_out.cpu().numpy()] == exp


# ------------------------------------------------------------------------------------------ Fp2 / Fp6 on their own
def test_fp2_fp6_direct_ops(eng):
    """the lower tower levels directly on the GPU (fp12_op 8-15), not only through Fp12
    operations - edge coefficients (0, 1, p - 1, the value whose Montgomery form is p - 1) and random ones, against
    the oracle (src/fields/fp2.rs:99-171,269-361; fp6.rs:189-236,267-424)."""
    rng = random.Random(205)
    P = o.P
    top = (P - 1) * pow(1 << 256, -1, P) % P  # Montgomery representative p - 1: the largest limbs the kernels see
    edge = [0, 1, 2, P - 1, P - 2, top, (P + 1) // 2, (P - 1) // 2]
    cases = [([rng.choice(edge) for _ in range(12)], [rng.choice(edge) for _ in range(12)]) for _ in range(200)]
    cases += [([rng.randrange(P) for _ in range(12)], [rng.randrange(P) for _ in range(12)]) for _ in range(300)]
    cases += [([v] * 12, [w_] * 12) for v in edge for w_ in edge]
    A = arr([w.fp12_b(o.fp12_from_list(a)) for a, _ in cases])
    B = arr([w.fp12_b(o.fp12_from_list(b)) for _, b in cases])
    got = {op: [o.fp12_to_list(w.b_fp12(bytes(r))) for r in eng.fp12_op_batch(op, A, B)] for op in range(8, 16)}
    inv2 = o.TWO_INV
    for i, (a, b) in enumerate(cases):
        a2, b2 = (a[0], a[1]), (b[0], b[1])
        a6 = ((a[0], a[1]), (a[2], a[3]), (a[4], a[5]))
        b6 = ((b[0], b[1]), (b[2], b[3]), (b[4], b[5]))
        flat2 = lambda x: [x[0], x[1]]
        flat6 = lambda x: [c for t in x for c in t]
        assert got[8][i][:2] == flat2(o.fp2_mul(a2, b2)) and not any(got[8][i][2:])
        assert got[9][i][:2] == flat2(o.fp2_sqr(a2)) and not any(got[9][i][2:])
        assert got[10][i][:6] == flat6(o.fp6_mul(a6, b6)) and not any(got[10][i][6:])
        assert got[11][i][:6] == flat6(o.fp6_sqr(a6))
        assert got[12][i][:2] == flat2(o.fp2_inv(a2))
        xi = o.fp2_residue_mul(a2)
        assert got[13][i][:6] == flat2(xi) + flat2(o.fp2_add(b2, xi)) + flat2(o.fp2_sub(b2, xi))
        assert got[14][i][:6] == flat6(o.fp6_inv(a6))
        assert got[15][i][:6] == [a[0] * b[0] % P, a[1] * b[0] % P, a[0] * inv2 % P, a[1] * inv2 % P, a[0], -a[1] % P]


# ------------------------------------------------------------------------------------------ glued loops for 2-/4-pair checks
def test_glued_pairing_checks_large(eng):
    """pairing_check_batch with 2 and 4 pairs per check runs ONE glued loop per check once there are enough checks to fill
    the GPU (one Fp12 squaring per digit for all pairs of a check, pairing.rs:970-1022); below that it multiplies separate
    Miller values.  Both forms must give the same verdicts: e(a P, Q) e(-a P, Q) = 1 and e(aP,Q) e(bP,Q) e(cP,Q) e(-(a+b+c)P, Q)
    = 1, with broken checks and infinite pairs sprinkled in."""
    import torch

    rng = random.Random(206)
    nchk = 148 * 256 + 333  # more than one wave of checks
    dev = torch.device("cuda", 0)

    def scalars(vals):
        return torch.from_numpy(arr([w.fp_b(v % o.R_ORDER) for v in vals])).to(dev)

    for k in (2, 4):
        n = nchk * k
        base = [rng.randrange(1, o.R_ORDER) for _ in range(64)]
        a = [base[(i * 7 + j) % 64] + i * 1000003 + j for i in range(nchk) for j in range(k - 1)]
        ks = []
        for i in range(nchk):
            row = a[i * (k - 1):(i + 1) * (k - 1)]
            ks += row + [-sum(row)]
        bad = sorted(rng.sample(range(nchk), 40))
        for c in bad:
            ks[c * k] += 1
        d_g1 = torch.empty((n, 64), dtype=torch.uint8, device=dev)
        d_g2 = torch.empty((n, 128), dtype=torch.uint8, device=dev)
        g1gen = torch.from_numpy(np.tile(np.frombuffer(w.g1_b(o.G1_GEN), np.uint8), (n, 1))).to(dev)
        g2gen = torch.from_numpy(np.tile(np.frombuffer(w.g2_b(o.G2_GEN), np.uint8), (n, 1))).to(dev)
        eng.g1_mul_batch_dev(g1gen, scalars(ks), d_g1)
        q = rng.randrange(1, o.R_ORDER)
        eng.g2_mul_batch_dev(g2gen, scalars([q] * n), d_g2)
        inf1 = np.zeros(n, np.uint8)
        skip = [c for c in rng.sample(range(nchk), 30) if c not in bad]
        for c in skip:  # an infinite pair contributes 1: the rest of the check no longer cancels
            inf1[c * k + 1] = 1
        expect = np.ones(nchk, bool)
        expect[bad] = False
        expect[skip] = False
        d_ok = torch.empty(nchk, dtype=torch.uint8, device=dev)
        eng.pairing_check_batch_dev(d_g1, d_g2, k, d_ok, d_g1_inf=torch.from_numpy(inf1).to(dev))
        got = d_ok.cpu().numpy().astype(bool)
        assert (got == expect).all(), (k, int((got != expect).sum()))
        # the small-batch path (separate Miller values) on a prefix agrees
        m = 64
        small = eng.pairing_check_batch(d_g1[: m * k].cpu().numpy(), d_g2[: m * k].cpu().numpy(), k, g1_inf=inf1[: m * k])
        assert (small == expect[:m]).all()


# ------------------------------------------------------------------------- two lanes per Miller loop (pairing_lanes.cuh)
def test_miller_lanes_match_one_thread_kernel(eng, monkeypatch):
    """k_miller_lanes (two cooperating lanes per pair, SYLOW_B200_LANES=2) against k_miller (=0) and the oracle:
    MillerLoopResult bit for bit, ragged sizes around the block and wave boundaries, infinity flags, and the automatic
    policy (=1: small batches and wave remainders on two lanes) through pairing_batch."""
    import torch

    rng = random.Random(41)
    dev = torch.device("cuda", 0)
    n = 148 * 256 + 148 * 128 - 77  # one whole one-thread wave plus a remainder that fits one two-lane wave
    rs = np.random.RandomState(6)
    k = rs.randint(0, 256, size=(2 * n, 32), dtype=np.uint8)
    k[:, 31] &= 0x1F
    d_g1 = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    d_g2 = torch.empty((n, 128), dtype=torch.uint8, device=dev)
    g1 = np.tile(np.frombuffer(w.g1_b(o.G1_GEN), np.uint8), (n, 1))
    g2 = np.tile(np.frombuffer(w.g2_b(o.G2_GEN), np.uint8), (n, 1))
    eng.g1_mul_batch_dev(torch.from_numpy(g1).to(dev), torch.from_numpy(k[:n]).to(dev), d_g1)
    eng.g2_mul_batch_dev(torch.from_numpy(g2).to(dev), torch.from_numpy(k[n:]).to(dev), d_g2)
    torch.cuda.synchronize()
    H1, H2 = d_g1.cpu().numpy(), d_g2.cpu().numpy()
    inf1 = np.zeros(n, dtype=np.uint8)
    inf2 = np.zeros(n, dtype=np.uint8)
    inf1[[3, 64, n - 1]] = 1
    inf2[[5, 64, 1000]] = 1
    res = {}
    for mode in ("0", "2", "1"):
        monkeypatch.setenv("SYLOW_B200_LANES", mode)
        res[mode] = {m: eng.miller_loop_batch(H1[:m], H2[:m], inf1[:m], inf2[:m])
                     for m in (1, 2, 15, 16, 17, 63, 64, 65, 150, 148 * 128 + 1, n)}
    for m, a in res["0"].items():
        assert np.array_equal(a, res["2"][m]), "two-lane kernel differs at n = %d" % m
        assert np.array_equal(a, res["1"][m]), "automatic policy differs at n = %d" % m
    full = res["2"][n]
    for i in (0, 1, 2, 16, 63, 64, 65, 149, 148 * 128, n - 2):
        if inf1[i] or inf2[i]:
            continue
        assert w.b_fp12(bytes(full[i])) == o.miller_loop(o.g2_precompute(w.b_g2(bytes(H2[i]))), w.b_g1(bytes(H1[i])))
    for i in (3, 5, 64, 1000, n - 1):
        assert w.b_fp12(bytes(full[i])) == o.FP12_ONE
    # the two-lane final exponentiation against the one-thread kernel and the oracle
    fx = {}
    for mode in ("0", "2", "1"):
        monkeypatch.setenv("SYLOW_B200_LANES", mode)
        fx[mode] = {m: eng.final_exp_batch(res["0"][n][:m]) for m in (1, 2, 17, 64, 65, 150, 148 * 64 + 3)}
    for m, a in fx["0"].items():
        assert np.array_equal(a, fx["2"][m]), "two-lane final exponentiation differs at n = %d" % m
        assert np.array_equal(a, fx["1"][m]), "automatic policy differs at n = %d" % m
    for i in (0, 1, 149):
        assert w.b_fp12(bytes(fx["2"][150][i])) == o.final_exponentiation(w.b_fp12(bytes(res["0"][n][i])))
    # pairing_batch end to end under the automatic policy, small batch
    monkeypatch.setenv("SYLOW_B200_LANES", "1")
    gt = eng.pairing_batch(H1[:5], H2[:5])
    for i in range(5):
        assert w.b_fp12(bytes(gt[i])) == o.pairing_affine(w.b_g1(bytes(H1[i])), w.b_g2(bytes(H2[i])))
