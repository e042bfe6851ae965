"""GPU tier: the reference's own pairing tests (src/pairing.rs:1038-1251) re-stated against the object-level
mirror of sylow's API (sylow_b200/sylow.py), plus the byte-codec and constructor tests of src/groups/mod.rs."""
import random

import pytest

from oracle import bn254_py as o

pytestmark = pytest.mark.gpu

DST = b"WARLOCK-CHAOS-V01-CS01-SHA-256"
MSG = (20).to_bytes(4, "big")


@pytest.fixture(scope="module")
def s():
    from sylow_b200 import sylow

    sylow.engine()
    return sylow


def test_gt_generator(s, kats):
    """src/pairing.rs:1052-1057"""
    gt = s.pairing(s.G1Projective.generator(), s.G2Projective.generator())
    assert list(gt.c) == [int(x, 16) for x in kats["gt_generator"]["fp12"]]
    assert gt == s.Gt.generator()


def test_signing(s):
    """src/pairing.rs:1059-1072: e(H(m)*k, G2) == e(H(m), G2*k)"""
    expander = s.XMDExpander(DST, 128)
    rng = random.Random(41)
    for _ in range(3):
        k = rng.randrange(1, s.R)
        hm = s.G1Projective.hash_to_curve(expander, MSG)
        sig = hm * k
        pk = s.G2Projective.generator() * k
        assert s.pairing(sig, s.G2Projective.generator()) == s.pairing(hm, pk)
        assert s.sign(k, MSG) == sig and s.verify(pk, MSG, sig) and not s.verify(pk, b"other", sig)
    kp = s.KeyPair.generate()
    assert s.verify(kp.public_key, MSG, s.sign(kp.secret_key, MSG))


def test_pairing_on_ecc_key_agreement(s):
    """src/pairing.rs:1074-1099: three-party shared secret e(G1, G2)^(abc) computed three ways"""
    rng = random.Random(42)
    a, b, c = (rng.randrange(1, s.R) for _ in range(3))
    g1, g2 = s.G1Projective.generator(), s.G2Projective.generator()
    k1 = s.pairing(g1 * b, g2 * c) * a
    k2 = s.pairing(g1 * c, g2 * a) * b
    k3 = s.pairing(g1 * a, g2 * b) * c
    assert k1 == k2 == k3


def test_identities(s):
    """src/pairing.rs:1101-1120"""
    g, h = s.G1Projective.generator(), s.G2Projective.generator()
    assert s.pairing(s.G1Projective.zero(), h) == s.Gt.identity()
    assert s.pairing(g, s.G2Projective.zero()) == s.Gt.identity()
    p = -s.pairing(g, h)
    assert p == s.pairing(g, -h) == s.pairing(-g, h)


def test_cases(s, kats):
    """src/pairing.rs:1122-1189"""
    t = kats["pairing_test_cases"]
    g1 = s.G1Projective.generator() * s.Fp(int(t["g1_scalar"], 16))
    g2 = s.G2Projective.generator() * s.Fp(int(t["g2_scalar"], 16))
    assert list(s.pairing(g1, g2).c) == [int(x, 16) for x in t["fp12"]]


def test_bilinearity(s):
    """src/pairing.rs:1192-1213"""
    for _ in range(3):
        p, q, k = s.G1Projective.rand(), s.G2Projective.rand(), s.Fr.rand()
        a = s.pairing(p, q) * k
        assert a == s.pairing(p * k, q) == s.pairing(p, q * k)
        assert a != s.Gt.identity()
        assert (a * s.Fr(-1)) + a == s.Gt.identity()


def test_batches(s):
    """src/pairing.rs:1216-1242"""
    assert s.glued_pairing([], []) == s.Gt.identity()
    n = 50
    ps, qs, sps, sqs = [], [], [], []
    for _ in range(n):
        p, q, k = s.G1Projective.rand(), s.G2Projective.rand(), s.Fr.rand()
        ps.append(p), qs.append(q), sps.append(p * k), sqs.append(q * k)
    assert s.glued_pairing(sps, qs) == s.glued_pairing(ps, sqs)


def test_precomputed_matches_pairing(s):
    """examples/verify_multiple_messages_same_signer.rs:63-71: precomputed pubkey Miller loop + final exp"""
    p, q = s.G1Projective.rand(), s.G2Projective.rand()
    assert q.precompute().miller_loop(p).final_exponentiation() == s.pairing(p, q)
    ref = o.g2_precompute((q.x, q.y, False))
    b = bytes(q.precompute().coeffs)
    assert int.from_bytes(b[:32], "little") == ref[0][0][0]


def test_constructors_and_codecs(s):
    """src/groups/mod.rs:769-873 (byte round trips, corrupted bytes rejected) and the checked constructors"""
    g = s.G1Affine.rand()
    assert s.G1Affine.from_be_bytes(g.to_be_bytes()) == g
    assert s.G1Affine.from_be_bytes(s.G1Affine.zero().to_be_bytes()).is_zero()
    bad = bytearray(g.to_be_bytes())
    bad[63] ^= 1
    with pytest.raises(s.NotOnCurve):
        s.G1Affine.from_be_bytes(bytes(bad))
    with pytest.raises(s.NotOnCurve):
        s.G1Affine.new(1, 3)
    assert s.G1Affine.new(1, 2) == s.G1Affine.generator()
    gen = s.G2Affine.generator()
    assert s.G2Affine.new(gen.x, gen.y) == gen
    with pytest.raises(s.NotOnCurve):
        s.G2Affine.new(gen.x, ((gen.y[0] + 1) % s.P, gen.y[1]))
    # a point of E'(Fp2) outside the r-torsion (src/groups/mod.rs:498-505 "invalid_subgroup_check")
    rng = random.Random(43)
    while True:
        x = (rng.randrange(s.P), rng.randrange(s.P))
        y = o.fp2_sqrt(o.fp2_add(o.fp2_mul(o.fp2_sqr(x), x), o.FP2_TWIST_CURVE_CONSTANT))
        if y is not None:
            break
    with pytest.raises(s.NotInSubgroup):
        s.G2Affine.new(x, y)


def test_expander(s, kats):
    """src/hasher.rs:430-470"""
    t = kats["xmd_sha256_short"]
    e = s.XMDExpander(t["dst"].encode(), 128, "sha256")
    for m, exp in t["vectors"]:
        assert e.expand_message(m.encode(), 32).hex() == exp
        assert len(e.hash_to_field(m.encode(), 2, 48)) == 2
    for name in ("xof_shake128_short", "xof_shake128_long_dst"):  # src/hasher.rs:393-428
        t = kats[name]
        x = s.XOFExpander(t["dst"].encode(), 128)
        for m, exp in t["vectors"]:
            assert x.expand_message(m.encode(), 32).hex() == exp
    # any Expander drives hash_to_curve (GroupTrait::hash_to_curve<E: Expander>, group.rs:137)
    p = s.G1Projective.hash_to_curve(s.XOFExpander(DST, 128), MSG)
    assert (int(p.x), int(p.y)) == o.proj_to_affine(o.FpOps, o.hash_to_curve_g1(MSG, DST, "shake128"))[:2]
