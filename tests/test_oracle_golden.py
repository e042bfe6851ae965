"""Pin the Python oracle against every golden vector the reference's own tests hold for the path
(SURVEY.md section 8c).  CPU only."""
import hashlib

from oracle import bn254_py as o

H = lambda s: int(s, 16)


def test_fp_mul(kats):
    for a, b, c in kats["fp_mul"]["cases"]:
        assert H(a) * H(b) % o.P == H(c)


def test_fp2_mul_div(kats):
    for a, b, c in kats["fp2_mul"]["cases"]:
        a, b, c = (tuple(map(H, v)) for v in (a, b, c))
        assert o.fp2_mul(a, b) == c
    for a, b, c in kats["fp2_div"]["cases"]:
        a, b, c = (tuple(map(H, v)) for v in (a, b, c))
        assert o.fp2_mul(a, o.fp2_inv(b)) == c


def _fp6(v):
    v = list(map(H, v))
    return ((v[0], v[1]), (v[2], v[3]), (v[4], v[5]))


def test_fp6_mul(kats):
    for a, b, c in kats["fp6_mul"]["cases"]:
        assert o.fp6_mul(_fp6(a), _fp6(b)) == _fp6(c)
        if a == b:
            assert o.fp6_sqr(_fp6(a)) == _fp6(c)


def test_constants(kats):
    assert o.TWO_INV == H(kats["two_inv"]["value"])
    assert o.FP2_TWIST_CURVE_CONSTANT == tuple(map(H, kats["fp2_twist_curve_constant"]["value"]))
    assert o.EPS_EXP0 == tuple(map(H, kats["eps_exp0"]["value"]))
    assert o.EPS_EXP1 == tuple(map(H, kats["eps_exp1"]["value"]))
    assert o.BLS_X == H(kats["bls_x"]["value"])
    g = kats["g2_generator"]
    assert o.G2_GEN[0] == tuple(map(H, g["x"])) and o.G2_GEN[1] == tuple(map(H, g["y"]))
    assert o.g2_is_on_curve(o.G2_GEN[0], o.G2_GEN[1])
    for name, tab in (("frobenius_coeff_fp6_c1", o.FROBENIUS_COEFF_FP6_C1),
                      ("frobenius_coeff_fp6_c2", o.FROBENIUS_COEFF_FP6_C2),
                      ("frobenius_coeff_fp12_c1", o.FROBENIUS_COEFF_FP12_C1)):
        assert [tuple(map(H, e)) for e in kats[name]["table"]] == tab, name
    # 6x+2 in NAF with implicit leading 1 (src/pairing.rs:26-30)
    v = 1
    for d in o.ATE_LOOP_COUNT_NAF:
        v = 2 * v + d
    assert v == 6 * o.BLS_X + 2


def test_gt_generator(kats):
    """e(G1gen, G2gen) == GT, src/pairing.rs:1052-1057."""
    gt = o.pairing_affine(o.G1_GEN, o.G2_GEN)
    assert o.fp12_to_list(gt) == list(map(H, kats["gt_generator"]["fp12"]))


def test_pairing_test_cases(kats):
    """src/pairing.rs:1122-1189."""
    t = kats["pairing_test_cases"]
    g1 = o.proj_mul(o.FpOps, o.affine_to_proj(o.FpOps, o.G1_GEN), H(t["g1_scalar"]))
    g2 = o.proj_mul(o.Fp2Ops, o.affine_to_proj(o.Fp2Ops, o.G2_GEN), H(t["g2_scalar"]))
    assert o.fp12_to_list(o.pairing(g1, g2)) == list(map(H, t["fp12"]))


def test_pairing_identities():
    """src/pairing.rs:1101-1120."""
    g, h = o.G1_GEN, o.G2_GEN
    assert o.pairing_affine((0, 1, True), h) == o.FP12_ONE
    assert o.pairing_affine(g, (o.FP2_ZERO, o.FP2_ONE, True)) == o.FP12_ONE
    p = o.fp12_conj(o.pairing_affine(g, h))
    assert p == o.pairing_affine(g, o.g2_affine_neg(h)) == o.pairing_affine(o.g1_affine_neg(g), h)
    assert o.glued_pairing([], []) == o.FP12_ONE


def _eip_g1(b):
    x, y = int.from_bytes(b[:32], "big"), int.from_bytes(b[32:64], "big")
    return (x, y, x == 0 and y == 0)


def test_eip196(kats):
    inp = bytes.fromhex(kats["eip196_add"]["input"])
    a, b = _eip_g1(inp[:64]), _eip_g1(inp[64:])
    s = o.proj_to_affine(o.FpOps, o.proj_add(o.FpOps, o.affine_to_proj(o.FpOps, a), o.affine_to_proj(o.FpOps, b)))
    assert s[0].to_bytes(32, "big") + s[1].to_bytes(32, "big") == bytes.fromhex(kats["eip196_add"]["expected"])
    inp = bytes.fromhex(kats["eip196_mul"]["input"])
    a, k = _eip_g1(inp[:64]), int.from_bytes(inp[64:96], "big")
    s = o.proj_to_affine(o.FpOps, o.proj_mul(o.FpOps, o.affine_to_proj(o.FpOps, a), k))
    assert s[0].to_bytes(32, "big") + s[1].to_bytes(32, "big") == bytes.fromhex(kats["eip196_mul"]["expected"])


def test_eip197_pairing_check(kats):
    """2-pair ecPairing vector -> 1 (examples/reth_bn128.rs:389-416); Fp2 imaginary part first
    (src/groups/g2.rs:325-328)."""
    inp = bytes.fromhex(kats["eip197_pair"]["input"])
    g1s, g2s = [], []
    for off in range(0, len(inp), 192):
        c = [int.from_bytes(inp[off + 32 * i: off + 32 * i + 32], "big") for i in range(6)]
        g1s.append((c[0], c[1], False))
        q = ((c[3], c[2]), (c[5], c[4]), False)
        assert o.g1_is_on_curve(c[0], c[1]) and o.g2_is_on_curve(q[0], q[1])
        g2s.append(q)
    assert o.glued_pairing(g1s, g2s) == o.FP12_ONE
    # glued == product of separate Miller loops, bit-exactly
    sep = o.FP12_ONE
    for p, q in zip(g1s, g2s):
        sep = o.fp12_mul(sep, o.miller_loop(o.g2_precompute(q), p))
    assert sep == o.glued_miller_loop([o.g2_precompute(q) for q in g2s], g1s)


def test_svdw_constants(kats):
    c = kats["svdw_constants"]
    assert (o.SVDW_Z, o.SVDW_C1, o.SVDW_C2, o.SVDW_C3, o.SVDW_C4) == tuple(H(c[n]) for n in ("z", "c1", "c2", "c3", "c4"))


def test_xmd_sha256(kats):
    for name in ("xmd_sha256_short", "xmd_sha256_long_dst"):
        t = kats[name]
        for msg, exp in t["vectors"]:
            assert o.expand_message_xmd(msg.encode(), t["dst"].encode(), t["len_in_bytes"], "sha256").hex() == exp


def test_xof_shake128(kats):
    """src/hasher.rs:393-428: XOFExpander::<Shake128>, short and oversize DST."""
    for name in ("xof_shake128_short", "xof_shake128_long_dst"):
        t = kats[name]
        for msg, exp in t["vectors"]:
            assert o.expand_message_xof(msg.encode(), t["dst"].encode(), t["len_in_bytes"]).hex() == exp


def test_keccak256_kat():
    # public Keccak-256 KATs (legacy padding, not SHA3-256)
    assert o.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert o.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    assert o.keccak256(b"a" * 200) != hashlib.sha3_256(b"a" * 200).digest()


def test_hash_to_curve_properties():
    """Keccak hash_to_curve values are unpinned by the reference: check invariants."""
    for msg in (b"", b"abc", (20).to_bytes(4, "big"), bytes(range(200))):
        for u in o.hash_to_field(msg):
            x, y = o.svdw_map_to_point(u)
            assert o.g1_is_on_curve(x, y) and o.fp_sgn0(y) == o.fp_sgn0(u)
        pt = o.proj_to_affine(o.FpOps, o.hash_to_curve_g1(msg))
        assert not pt[2] and o.g1_is_on_curve(pt[0], pt[1])


def test_sign_verify_roundtrip():
    """src/pairing.rs:1059-1072 sign -> verify self-consistency, plus the batch form."""
    sk = 0x1234567890ABCDEF1234567890ABCDEF % o.R_ORDER
    pk = o.proj_mul(o.Fp2Ops, o.affine_to_proj(o.Fp2Ops, o.G2_GEN), sk)
    msg = (20).to_bytes(4, "big")
    sig = o.sign(sk, msg)
    assert o.verify(pk, msg, sig)
    assert not o.verify(pk, b"other", sig)
    pka, siga = o.proj_to_affine(o.Fp2Ops, pk), o.proj_to_affine(o.FpOps, sig)
    assert o.verify_batch([pka], [msg], [siga])
    assert not o.verify_batch([pka], [b"other"], [siga])
