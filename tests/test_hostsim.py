"""CPU tier: the DEVICE source (sylow_b200/csrc/*.cuh) compiled for the host (tests/hostsim) must agree
with the oracle.  This checks tower/curve/pairing/hash logic before a B200 is available; the PTX
primitives themselves are covered by the -m gpu tests."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import bn254_py as o
from tests import wire as w

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "hostsim", "libsylow_hostsim.so")


@pytest.fixture(scope="module")
def hs():
    src = os.path.join(HERE, "hostsim", "hostsim.cu")
    csrc = os.path.join(os.path.dirname(HERE), "sylow_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["nvcc", "-x", "cu", "-O2", "-std=c++17", "-DSYLOW_HOSTSIM", "-shared", "-Xcompiler",
                               "-fPIC,-pthread", "-w", "-o", SO, src])
    lib = ctypes.CDLL(SO)
    lib.hs_hash_to_g1.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
    lib.hs_hash_to_field.argtypes = lib.hs_hash_to_g1.argtypes
    lib.hs_keccak256.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
    lib.hs_lagrange.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_char_p]
    lib.hs_fr_op.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
    return lib


def _fp_op(hs, op, a, b=0):
    out = ctypes.create_string_buffer(32)
    hs.hs_fp_op(op, w.fp_b(a), w.fp_b(b), out)
    return w.b_fp(out.raw)


def test_fp_ops(hs):
    rng = random.Random(1)
    edge = [0, 1, 2, o.P - 1, o.P - 2, (o.P + 1) // 2, (1 << 253), (1 << 254) - 1 - (1 << 200)]
    vals = edge + [rng.randrange(o.P) for _ in range(40)]
    for a in vals:
        for b in (vals[0], vals[3], rng.choice(vals), rng.randrange(o.P)):
            assert _fp_op(hs, 0, a, b) == a * b % o.P
            assert _fp_op(hs, 1, a, b) == (a + b) % o.P
            assert _fp_op(hs, 2, a, b) == (a - b) % o.P
        assert _fp_op(hs, 3, a) == o.fp_inv(a)
        assert _fp_op(hs, 4, a) == a * o.TWO_INV % o.P
        assert _fp_op(hs, 5, a) == 9 * a % o.P
        assert _fp_op(hs, 6, a) == -a % o.P
        # binary-GCD inversion (op 3 above) against the Fermat ladder it replaced, and the Jacobi iteration against
        # x^((p-1)/2): the kernel reports (x / p) + 1, plus 4 if the two disagree
        assert _fp_op(hs, 8, a) == o.fp_inv(a)
        leg = pow(a, (o.P - 1) // 2, o.P)
        # the kernel returns the RAW limb value; _fp_op converts from the wire (canonical) form, so compare raw bytes
        out = ctypes.create_string_buffer(32)
        hs.hs_fp_op(7, w.fp_b(a), w.fp_b(0), out)
        assert int.from_bytes(out.raw, "little") == {0: 1, 1: 2, o.P - 1: 0}[leg]


def _fp12_op(hs, op, a, b=None):
    out = ctypes.create_string_buffer(384)
    hs.hs_fp12_op(op, w.fp12_b(a), w.fp12_b(b if b is not None else a), out)
    return w.b_fp12(out.raw)


def test_lin9(hs):
    """9 x + y + t mod p with the quotient estimated from the top bits (the xi multiplications): every residue
    boundary k p - 1, k p, k p + 1 and the extreme operands, on raw limb values."""
    rng = random.Random(81)
    out = ctypes.create_string_buffer(32)
    b = lambda v: v.to_bytes(32, "little")
    P = o.P
    cases = [(0, 0, 0), (P - 1, P, P - 1), (P - 1, P - 1, P - 1), (P, P, 0), (1, P, 0), (0, P, P - 1)]
    for k in range(1, 11):  # land T = 9x + y + t on k p - 1, k p, k p + 1
        for d in (-1, 0, 1):
            T = k * P + d
            x = min(P - 1, T // 9)
            rest = T - 9 * x
            y = min(P, rest)
            t = rest - y
            if 0 <= t < P:
                cases.append((x, y, t))
    cases += [(rng.randrange(P), rng.randrange(P + 1), rng.randrange(P)) for _ in range(3000)]
    for x, y, t in cases:
        hs.hs_lin9(b(x), b(y), b(t), 1, out)
        assert int.from_bytes(out.raw, "little") == (9 * x + y + t) % P, (hex(x), hex(y), hex(t))
        hs.hs_lin9(b(x), b(y), b(0), 0, out)
        assert int.from_bytes(out.raw, "little") == (9 * x + y) % P


def test_fp12_edge_coefficients(hs):
    """Extremes of the lazy-reduction bounds (tower.cuh fp6_mul): coefficients from {0, 1, p-1, p-2, ...} in every
    position; the host build aborts if a high half is not below p before a Montgomery reduction."""
    rng = random.Random(30)
    edge = [0, 1, 2, o.P - 1, o.P - 2, (o.P - 1) // 2, (o.P + 1) // 2, (1 << 32) - 1, 1 << 224, (1 << 253) + 1,
            (1 << 256) % o.P, o.P - ((1 << 256) % o.P)]
    # the bounds are on the MONTGOMERY representatives: -1/R mod p is stored as p - 1, the largest limb pattern
    top = o.P - pow(1 << 256, -1, o.P)
    edge += [top, (o.P - 2) * pow(1 << 256, -1, o.P) % o.P]
    cases = [([top] * 12, [top] * 12), ([top, 0] * 6, [0, top] * 6), ([0, top] * 6, [0, top] * 6),
             ([top, 0] * 6, [top, 0] * 6), ([top] * 6 + [0] * 6, [0] * 6 + [top] * 6), ([o.P - 1] * 12, [o.P - 1] * 12)]
    cases += [([rng.choice(edge) for _ in range(12)], [rng.choice(edge) for _ in range(12)]) for _ in range(60)]
    cases += [([rng.choice([0, top]) for _ in range(12)], [rng.choice([0, top]) for _ in range(12)]) for _ in range(60)]
    out = ctypes.create_string_buffer(192)
    for ca, cb in cases:
        a, b = o.fp12_from_list(ca), o.fp12_from_list(cb)
        hs.hs_fp6_mul(w.fp12_b(a)[:192], w.fp12_b(b)[:192], out)
        assert w.b_fp12(out.raw + bytes(192))[0] == o.fp6_mul(a[0], b[0])
        assert _fp12_op(hs, 0, a, b) == o.fp12_mul(a, b)
        assert _fp12_op(hs, 1, a) == o.fp12_sqr(a)
        assert _fp12_op(hs, 7, a, b) == o.fp12_sparse_mul(a, b[0][0], b[0][1], b[0][2])


def test_fp12_ops(hs):
    rng = random.Random(2)
    for _ in range(4):
        a, b = w.rand_fp12(rng), w.rand_fp12(rng)
        assert _fp12_op(hs, 0, a, b) == o.fp12_mul(a, b)
        assert _fp12_op(hs, 1, a) == o.fp12_sqr(a)
        assert _fp12_op(hs, 2, a) == o.fp12_inv(a)
        for e in (1, 2, 3):
            assert _fp12_op(hs, 2 + e, a) == o.fp12_frobenius(a, e)
        assert _fp12_op(hs, 7, a, b) == o.fp12_sparse_mul(a, b[0][0], b[0][1], b[0][2])
    # cyclotomic squaring is only a squaring on the cyclotomic subgroup
    g = o.pairing_affine(o.G1_GEN, o.G2_GEN)
    assert _fp12_op(hs, 6, g) == o.cyclotomic_squared(g) == o.fp12_sqr(g)


def test_pairing_generators(hs, kats):
    out = ctypes.create_string_buffer(384)
    hs.hs_pairing(w.g1_b(o.G1_GEN), w.g2_b(o.G2_GEN), out)
    assert [w.b_fp(out.raw[32 * i: 32 * i + 32]) for i in range(12)] == [int(x, 16) for x in kats["gt_generator"]["fp12"]]


def test_miller_and_final_exp_random(hs):
    rng = random.Random(3)
    p, q = w.rand_g1(rng), w.rand_g2(rng)
    out = ctypes.create_string_buffer(384)
    hs.hs_miller_loop(w.g1_b(p), w.g2_b(q), out)
    f = o.miller_loop(o.g2_precompute(q), p)
    assert w.b_fp12(out.raw) == f  # MillerLoopResult itself is bit-identical
    out2 = ctypes.create_string_buffer(384)
    hs.hs_final_exp(out.raw, out2)
    assert w.b_fp12(out2.raw) == o.final_exponentiation(f)


def test_miller_loop_two_lanes(hs):
    """pairing_lanes.cuh: two cooperating lanes (two host threads here) produce the one-thread MillerLoopResult."""
    rng = random.Random(33)
    for p, q in [(o.G1_GEN, o.G2_GEN)] + [(w.rand_g1(rng), w.rand_g2(rng)) for _ in range(2)]:
        out = ctypes.create_string_buffer(384)
        hs.hs_miller_loop_lanes(w.g1_b(p), w.g2_b(q), out)
        assert w.b_fp12(out.raw) == o.miller_loop(o.g2_precompute(q), p)


def test_final_exponentiation_two_lanes(hs):
    """pairing_lanes.cuh: the two-lane final exponentiation against the oracle (and so against the one-thread chain)."""
    rng = random.Random(34)
    p, q = w.rand_g1(rng), w.rand_g2(rng)
    for f in (o.miller_loop(o.g2_precompute(q), p), o.miller_loop(o.g2_precompute(o.G2_GEN), o.G1_GEN), w.rand_fp12(rng)):
        out = ctypes.create_string_buffer(384)
        hs.hs_final_exp_lanes(w.fp12_b(f), out)
        assert w.b_fp12(out.raw) == o.final_exponentiation(f)


def test_scalar_mul(hs, kats):
    rng = random.Random(4)
    out = ctypes.create_string_buffer(64)
    inp = bytes.fromhex(kats["eip196_mul"]["input"])
    x, y, k = (int.from_bytes(inp[32 * i: 32 * i + 32], "big") for i in range(3))
    assert hs.hs_g1_mul(w.g1_b((x, y)), 0, w.fp_b(k), out) == 0
    exp = bytes.fromhex(kats["eip196_mul"]["expected"])
    assert w.b_g1(out.raw)[:2] == (int.from_bytes(exp[:32], "big"), int.from_bytes(exp[32:], "big"))
    p = w.rand_g1(rng)
    for k in (0, 1, 2, o.R_ORDER, o.R_ORDER - 1, o.P - 1, rng.randrange(o.P)):
        inf = hs.hs_g1_mul(w.g1_b(p), 0, w.fp_b(k), out)
        ref = o.proj_to_affine(o.FpOps, o.proj_mul(o.FpOps, o.affine_to_proj(o.FpOps, p), k))
        assert (w.b_g1(out.raw, inf)) == ref
    q = w.rand_g2(rng)
    out = ctypes.create_string_buffer(128)
    for k in (0, 1, o.R_ORDER - 1, rng.randrange(o.P)):
        inf = hs.hs_g2_mul(w.g2_b(q), 0, w.fp_b(k), out)
        ref = o.proj_to_affine(o.Fp2Ops, o.proj_mul(o.Fp2Ops, o.affine_to_proj(o.Fp2Ops, q), k))
        assert w.b_g2(out.raw, inf) == ref
    # infinity in -> infinity out
    out = ctypes.create_string_buffer(64)
    assert hs.hs_g1_mul(w.g1_b((0, 1)), 1, w.fp_b(5), out) == 1 and w.b_g1(out.raw)[:2] == (0, 1)


def test_glv_decompose(hs):
    """k = k1 + k2 * lambda (mod r) with both halves below 2^127, for every 256-bit k."""
    rng = random.Random(77)
    lam = 0xB3C4D79D41A917585BFC41088D8DAAA78B17EA66B99C90DD
    out = ctypes.create_string_buffer(32)
    edge = [0, 1, 2, lam, lam + 1, o.R_ORDER - 1, o.R_ORDER, o.R_ORDER + 1, o.P - 1, o.P, 2**256 - 1, 2**255,
            2**128 - 1, 2**128, 2**127]
    for k in edge + [rng.randrange(2**256) for _ in range(3000)] + [rng.randrange(o.R_ORDER) for _ in range(1000)]:
        s = hs.hs_glv_decompose(k.to_bytes(32, "little"), out)
        k1 = int.from_bytes(out.raw[:16], "little") * (-1 if s & 1 else 1)
        k2 = int.from_bytes(out.raw[16:], "little") * (-1 if s & 2 else 1)
        assert (k1 + k2 * lam - k) % o.R_ORDER == 0, hex(k)
        assert abs(k1) < 2**127 and abs(k2) < 2**127, hex(k)


def test_gls_decompose_and_mul(hs):
    """4-dimensional GLS on G2: k = sum k_j mu^j (mod r), mu = 6 x^2 = p mod r, |k_j| < 2^67, and the ladder
    equals the oracle's k * Q."""
    rng = random.Random(83)
    mu = 6 * 4965661367192848881 ** 2
    assert mu == o.P % o.R_ORDER
    out = ctypes.create_string_buffer(48)
    edge = [0, 1, 2, mu, mu + 1, mu * mu % o.R_ORDER, o.R_ORDER - 1, o.R_ORDER, o.R_ORDER + 1, o.P - 1, 2**256 - 1, 2**255,
            2**64, 2**128 - 1]
    for k in edge + [rng.randrange(2**256) for _ in range(3000)] + [rng.randrange(o.R_ORDER) for _ in range(1000)]:
        s = hs.hs_gls_decompose(k.to_bytes(32, "little"), out)
        ks = [int.from_bytes(out.raw[12 * j: 12 * j + 12], "little") * (-1 if (s >> j) & 1 else 1) for j in range(4)]
        assert (sum(kj * pow(mu, j, o.R_ORDER) for j, kj in enumerate(ks)) - k) % o.R_ORDER == 0, hex(k)
        assert all(abs(kj) < 2**67 for kj in ks), hex(k)
    q = w.rand_g2(rng)
    out = ctypes.create_string_buffer(128)
    for k in edge[:11] + [rng.randrange(2**256) for _ in range(4)]:
        inf = hs.hs_g2_mul_gls(w.g2_b(q), 0, k.to_bytes(32, "little"), out)
        ref = o.proj_to_affine(o.Fp2Ops, o.proj_mul(o.Fp2Ops, o.affine_to_proj(o.Fp2Ops, q), k % o.R_ORDER))
        assert w.b_g2(out.raw, inf) == ref, hex(k)
    assert hs.hs_g2_mul_gls(w.g2_b(((0, 0), (1, 0))), 1, (5).to_bytes(32, "little"), out) == 1


def test_scalar_mul_glv(hs):
    """GLV ladder == the oracle's k * P on both groups (affine result)."""
    rng = random.Random(78)
    lam = 0xB3C4D79D41A917585BFC41088D8DAAA78B17EA66B99C90DD
    p = w.rand_g1(rng)
    out = ctypes.create_string_buffer(64)
    ks = [0, 1, 2, 15, 16, lam, o.R_ORDER - 1, o.R_ORDER, o.R_ORDER + 1, o.P - 1, 2**256 - 1]
    for k in ks + [rng.randrange(2**256) for _ in range(6)]:
        inf = hs.hs_g1_mul_glv(w.g1_b(p), 0, k.to_bytes(32, "little"), out)
        ref = o.proj_to_affine(o.FpOps, o.proj_mul(o.FpOps, o.affine_to_proj(o.FpOps, p), k % o.R_ORDER))
        assert w.b_g1(out.raw, inf) == ref, hex(k)
    q = w.rand_g2(rng)
    out = ctypes.create_string_buffer(128)
    for k in ks + [rng.randrange(2**256) for _ in range(4)]:
        inf = hs.hs_g2_mul_glv(w.g2_b(q), 0, k.to_bytes(32, "little"), out)
        ref = o.proj_to_affine(o.Fp2Ops, o.proj_mul(o.Fp2Ops, o.affine_to_proj(o.Fp2Ops, q), k % o.R_ORDER))
        assert w.b_g2(out.raw, inf) == ref, hex(k)
    out = ctypes.create_string_buffer(64)
    assert hs.hs_g1_mul_glv(w.g1_b((0, 1)), 1, (5).to_bytes(32, "little"), out) == 1 and w.b_g1(out.raw)[:2] == (0, 1)


def test_fr_arithmetic_and_lagrange(hs):
    """fr.cuh against Python integers mod r; Lagrange coefficients of examples/dkg.rs:216-226."""
    rng = random.Random(79)
    r = o.R_ORDER
    out = ctypes.create_string_buffer(32)
    edge = [0, 1, 2, r - 1, r - 2, (r + 1) // 2, 2**64 - 1]
    vals = edge + [rng.randrange(r) for _ in range(30)]
    for a in vals:
        for b in (vals[0], vals[3], rng.choice(vals), rng.randrange(r)):
            for op, ref in ((0, a * b % r), (1, (a + b) % r), (2, (a - b) % r)):
                hs.hs_fr_op(op, a.to_bytes(32, "little"), b.to_bytes(32, "little"), out)
                assert int.from_bytes(out.raw, "little") == ref, (op, hex(a), hex(b))
    for a in edge + [rng.randrange(r) for _ in range(4)]:
        hs.hs_fr_op(3, a.to_bytes(32, "little"), bytes(32), out)
        assert int.from_bytes(out.raw, "little") == (pow(a, -1, r) if a else 0)  # inv(0) = 0, fp.rs:418-424
    # unreduced 256-bit inputs are reduced on the way in
    hs.hs_fr_op(1, (2**256 - 1).to_bytes(32, "little"), (r + 5).to_bytes(32, "little"), out)
    assert int.from_bytes(out.raw, "little") == (2**256 - 1 + 5) % r
    for ids in ([1, 2, 3], [5, 2, 9, 4, 7], [2**64 - 1, 1, 2**63], [3], [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11]):
        arr = (ctypes.c_uint64 * len(ids))(*ids)
        lam = []
        for i in range(len(ids)):
            hs.hs_lagrange(arr, len(ids), i, out)
            lam.append(int.from_bytes(out.raw, "little"))
        assert lam == o.lagrange_coefficients(ids)
        assert sum(lam) % r == 1  # interpolating the constant polynomial 1


def test_svdw_pair_shared_inversion(hs):
    """The two SvdW maps of one hash share a single inversion; inv0 semantics (u = 1/2 makes tv1 * tv2 = 0,
    svdw.rs:195 + fp.rs:418-424) must survive the sharing on either side and on both."""
    rng = random.Random(80)
    half = (o.P + 1) // 2
    out = ctypes.create_string_buffer(128)
    specials = [0, 1, half, o.P - half, o.P - 1]
    pairs = [(a, b) for a in specials for b in specials] + [(rng.randrange(o.P), rng.randrange(o.P)) for _ in range(6)]
    pairs += [(half, rng.randrange(o.P)), (rng.randrange(o.P), half)]
    for u0, u1 in pairs:
        assert hs.hs_svdw_pair(w.fp_b(u0), w.fp_b(u1), out) == 1
        got = [int.from_bytes(out.raw[32 * i: 32 * i + 32], "little") for i in range(4)]
        assert tuple(got[:2]) == tuple(o.svdw_map_to_point(u0)[:2]), hex(u0)
        assert tuple(got[2:]) == tuple(o.svdw_map_to_point(u1)[:2]), hex(u1)


def test_g1_add_eip196(hs, kats):
    inp = bytes.fromhex(kats["eip196_add"]["input"])
    c = [int.from_bytes(inp[32 * i: 32 * i + 32], "big") for i in range(4)]
    out = ctypes.create_string_buffer(64)
    assert hs.hs_g1_add(w.g1_b(c[0:2]), 0, w.g1_b(c[2:4]), 0, out) == 0
    exp = bytes.fromhex(kats["eip196_add"]["expected"])
    assert w.b_g1(out.raw)[:2] == (int.from_bytes(exp[:32], "big"), int.from_bytes(exp[32:], "big"))
    # P + (-P) = infinity, P + infinity = P
    p = (c[0], c[1], False)
    assert hs.hs_g1_add(w.g1_b(p), 0, w.g1_b(o.g1_affine_neg(p)), 0, out) == 1
    assert hs.hs_g1_add(w.g1_b(p), 0, w.g1_b((0, 1)), 1, out) == 0 and w.b_g1(out.raw) == p


def test_keccak_and_hash_to_curve(hs):
    out = ctypes.create_string_buffer(32)
    for m in (b"", b"abc", bytes(135), bytes(136), bytes(range(256)) * 3):
        hs.hs_keccak256(m, len(m), out)
        assert out.raw == o.keccak256(m)
    dst_prime = o.DST + bytes([len(o.DST)])
    rng = random.Random(5)
    for m in (b"", (20).to_bytes(4, "big"), bytes(32), bytes(rng.randrange(256) for _ in range(200))):
        o64 = ctypes.create_string_buffer(64)
        hs.hs_hash_to_field(m, len(m), dst_prime, len(dst_prime), o64)
        assert [w.b_fp(o64.raw[:32]), w.b_fp(o64.raw[32:])] == o.hash_to_field(m)
        assert hs.hs_hash_to_g1(m, len(m), dst_prime, len(dst_prime), o64) == 0
        assert w.b_g1(o64.raw) == o.proj_to_affine(o.FpOps, o.hash_to_curve_g1(m))


def test_precompute_and_glued(hs):
    """G2Affine::precompute coefficients bit-exact (pairing.rs:676-708) and the mixed fused/precomputed glued
    loop equal to glued_miller_loop (pairing.rs:970-1022)."""
    rng = random.Random(6)
    q = w.rand_g2(rng)
    out = ctypes.create_string_buffer(87 * 192)
    hs.hs_g2_precompute(w.g2_b(q), out)
    ref = o.g2_precompute(q)
    got = [tuple((w.b_fp(out.raw[192 * i + 64 * j: 192 * i + 64 * j + 32]), w.b_fp(out.raw[192 * i + 64 * j + 32: 192 * i + 64 * j + 64]))
                 for j in range(3)) for i in range(87)]
    assert got == [tuple(c) for c in ref]
    f_out = ctypes.create_string_buffer(384)
    for nv, nf in ((1, 1), (1, 3), (0, 1), (1, 0)):
        ps = [w.rand_g1(rng) for _ in range(nv + nf)]
        qv = w.rand_g2(rng)
        qf = [w.rand_g2(rng) for _ in range(nf)]
        for skip in ([0] * (nv + nf), [0] * (nv + nf - 1) + [1]):
            hs.hs_glued(b"".join(w.g1_b(p) for p in ps), bytes(skip), w.g2_b(qv), b"".join(w.g2_b(x) for x in qf) or b"\0",
                        nv, nf, f_out)
            qs = ([qv] if nv else []) + qf
            keep = [i for i in range(nv + nf) if not skip[i]]
            exp = o.glued_miller_loop([o.g2_precompute(qs[i]) for i in keep], [ps[i] for i in keep])
            assert w.b_fp12(f_out.raw) == exp, (nv, nf, skip)


def _twist_points_outside_subgroup(rng, n):
    out = []
    while len(out) < n:
        x = (rng.randrange(o.P), rng.randrange(o.P))
        y = o.fp2_sqrt(o.fp2_add(o.fp2_mul(o.fp2_sqr(x), x), o.FP2_TWIST_CURVE_CONSTANT))
        if y is not None:
            out.append((x, y, False))
    return out


def test_g2_subgroup_check(hs):
    """G2Projective::new subgroup relation (g2.rs:460-525)."""
    rng = random.Random(7)
    good = [o.G2_GEN, w.rand_g2(rng)]
    bad = _twist_points_outside_subgroup(rng, 3)
    for q in good + bad:
        exp = o.g2_projective_new(q[0], q[1])
        assert hs.hs_g2_on_curve(w.g2_b(q)) == 1
        assert hs.hs_g2_in_subgroup(w.g2_b(q)) == (1 if exp == "ok" else 0)
    assert [o.g2_projective_new(q[0], q[1]) for q in bad] == ["NotInSubgroup"] * 3
    off = (good[1][0], o.fp2_add(good[1][1], o.FP2_ONE), False)
    assert hs.hs_g2_on_curve(w.g2_b(off)) == 0 and o.g2_projective_new(off[0], off[1]) == "NotOnCurve"


def test_gt_pow(hs):
    """`Gt * Fr` (gt.rs:188-215) and the bilinearity identity e(P,Q)*s == e(sP,Q) (pairing.rs:1192-1213)."""
    rng = random.Random(8)
    g = o.pairing_affine(o.G1_GEN, o.G2_GEN)
    out = ctypes.create_string_buffer(384)
    mu = o.P % o.R_ORDER  # the exponent is split over the Frobenius eigenvalue (4-dimensional GLS, pairing.cuh)
    for k in (0, 1, 2, o.R_ORDER - 1, mu, mu * mu % o.R_ORDER, mu - 1, rng.randrange(o.R_ORDER), rng.randrange(o.R_ORDER)):
        hs.hs_gt_pow(w.fp12_b(g), w.fp_b(k), out)
        assert w.b_fp12(out.raw) == o.gt_mul(g, k)


def test_expander_generic(hs, kats):
    """XMD over SHA-256 (the reference's RFC 9380 vectors, hasher.rs:367-376) and Keccak-256, several lengths."""
    hs.hs_expand.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
                             ctypes.c_uint32, ctypes.c_char_p]
    t = kats["xmd_sha256_short"]
    dp = t["dst"].encode() + bytes([len(t["dst"])])
    for m, e in t["vectors"]:
        out = ctypes.create_string_buffer(32)
        hs.hs_expand(1, m.encode(), len(m), dp, len(dp), 32, out)
        assert out.raw.hex() == e
    t = kats["xof_shake128_short"]  # XOFExpander::<Shake128>, hasher.rs:345-355,394-410
    dp = t["dst"].encode() + bytes([len(t["dst"])])
    for m, e in t["vectors"]:
        out = ctypes.create_string_buffer(32)
        hs.hs_expand(2, m.encode(), len(m), dp, len(dp), 32, out)
        assert out.raw.hex() == e
    dp = o.DST + bytes([len(o.DST)])
    for hid, name in ((0, "keccak256"), (1, "sha256"), (2, "shake128")):
        for ln in (32, 96, 177, 400):  # 400 > the 168-byte SHAKE128 rate: several squeeze permutations
            out = ctypes.create_string_buffer(ln)
            hs.hs_expand(hid, b"abcdef" * 30, 180, dp, len(dp), ln, out)
            assert out.raw == o.expand_message(b"abcdef" * 30, o.DST, ln, name)


def test_batch_weights_and_short_ladder(hs):
    """The two pieces of random-weight batch verification: r_i = first 8 bytes of Keccak-256(seed || LE64(i)) | 1
    (hash.cuh) and the 64-bit ladder r * P (curve.cuh), against the oracle's Keccak-256 and scalar multiplication."""
    hs.hs_batch_weight.restype = ctypes.c_uint64
    hs.hs_batch_weight.argtypes = [ctypes.c_char_p, ctypes.c_uint64]
    hs.hs_g1_mul_u64.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_char_p]
    rng = random.Random(64)
    seed = bytes(rng.randrange(256) for _ in range(32))
    for idx in (0, 1, 2, 255, 256, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, rng.randrange(1 << 64)):
        exp = int.from_bytes(o.keccak256(seed + idx.to_bytes(8, "little"))[:8], "little") | 1
        assert hs.hs_batch_weight(seed, idx) == exp
    out, inf = ctypes.create_string_buffer(64), ctypes.create_string_buffer(1)
    p = w.rand_g1(rng)
    for k in (1, 2, 15, 16, 17, (1 << 60) - 1, 1 << 60, (1 << 64) - 1, rng.randrange(1 << 64) | 1):
        hs.hs_g1_mul_u64(w.g1_b(p), k, out, inf)
        assert w.b_g1(out.raw, inf.raw[0]) == o.proj_to_affine(o.FpOps, o.proj_mul(o.FpOps, o.affine_to_proj(o.FpOps, p), k))
