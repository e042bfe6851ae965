"""Python big-int restatement of sylow's BN254 hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *small-case* oracle: plain Python integers, no limbs, no Montgomery form.  It
exists to pin values (every sylow value is a canonical residue, so results are fixed by the
mathematics plus the *formulas that choose representatives* -- the Miller-loop line scalings and
the projective group law).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` leg may import it; the product path (``sylow_b200``) never does.

Parity status: PINNED for Fp/Fp2/Fp6 arithmetic, pairing (Gt generator, test_cases KAT,
EIP-197 2-pair vector), G1 add/mul (EIP-196 vectors), SvdW constants, XMD framing (RFC 9380
SHA-256 vectors) -- see tests/test_oracle_golden.py.  UNPINNED in the reference's own tests (the
oracle is sole authority): Keccak-256 hash_to_curve end-to-end values and glued_miller_loop with
infinite points.

Every function cites the reference file:line (under /root/reference/) it restates.
"""
from __future__ import annotations

import hashlib

# ----------------------------------------------------------------------------------------------
# constants                                                     src/fields/fp.rs:51-56,538-542
# ----------------------------------------------------------------------------------------------
P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R_ORDER = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
BLS_X = 4965661367192848881  # src/groups/g2.rs:112
# src/pairing.rs:26-30
ATE_LOOP_COUNT_NAF = [
    1, 0, 1, 0, 0, 0, -1, 0, -1, 0, 0, 0, -1, 0, 1, 0, -1, 0, 0, -1, 0, 0, 0, 0, 0, 1, 0, 0, -1, 0,
    1, 0, 0, -1, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, -1, 0, -1, 0, 0, 1, 0, 0, 0, -1, 0, 0, -1, 0, 1, 0,
    1, 0, 0, 0,
]
DST = b"WARLOCK-CHAOS-V01-CS01-SHA-256"  # src/lib.rs:90
SECURITY_BITS = 128  # src/lib.rs:94


# ----------------------------------------------------------------------------------------------
# Fp                                                             src/fields/fp.rs:304-457,611-662
# ----------------------------------------------------------------------------------------------
def fp_inv(a: int) -> int:
    """inv(0) = 0, src/fields/fp.rs:418-424."""
    return pow(a, P - 2, P) if a % P else 0


def fp_sqrt(a: int):
    """x^((p+1)/4) with post-check, src/fields/fp.rs:611-616.  Returns None if not a square."""
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a % P else None


def fp_is_square(a: int) -> bool:
    """Legendre; true for 0, src/fields/fp.rs:625-631."""
    return pow(a, (P - 1) // 2, P) in (0, 1)


def fp_sgn0(a: int) -> int:
    """Parity of the canonical value, src/fields/fp.rs:636-644."""
    return a & 1


def fp_compute_naf(x: int):
    """Prodinger NAF as (np, nm) bit masks, src/fields/fp.rs:653-662."""
    xh = x >> 1
    x3 = (x + xh) & ((1 << 256) - 1)
    c = xh ^ x3
    return x3 & c, xh & c


# ----------------------------------------------------------------------------------------------
# Fp2 = Fp[u]/(u^2+1)                                            src/fields/fp2.rs
# ----------------------------------------------------------------------------------------------
FP2_ZERO = (0, 0)
FP2_ONE = (1, 0)
TWO_INV = (P + 1) // 2  # src/fields/fp2.rs:18-23


def fp2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def fp2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def fp2_neg(a):
    return (-a[0] % P, -a[1] % P)


def fp2_mul(a, b):
    """Schoolbook, src/fields/fp2.rs:302-305."""
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def fp2_sqr(a):
    """src/fields/fp2.rs:164-171."""
    return ((a[0] + a[1]) * (a[0] - a[1]) % P, 2 * a[0] * a[1] % P)


def fp2_scale(a, k: int):
    """FieldExtension::scale, src/fields/extensions.rs:86-94."""
    return (a[0] * k % P, a[1] * k % P)


def fp2_residue_mul(a):
    """(a+bu)(9+u), src/fields/fp2.rs:99-107."""
    return ((9 * a[0] - a[1]) % P, (a[0] + 9 * a[1]) % P)


def fp2_conj(a):
    """frobenius(1): coefficient is the Fp non-residue -1, src/fields/fp2.rs:119-133."""
    return (a[0], -a[1] % P)


def fp2_frobenius(a, e: int):
    return a if e % 2 == 0 else fp2_conj(a)


def fp2_inv(a):
    """src/fields/fp2.rs:343-361."""
    t = fp_inv((a[0] * a[0] + a[1] * a[1]) % P)
    return (a[0] * t % P, -(a[1] * t) % P)


def fp2_pow(a, e: int):
    r = FP2_ONE
    for i in reversed(range(e.bit_length())):
        r = fp2_sqr(r)
        if (e >> i) & 1:
            r = fp2_mul(r, a)
    return r


XI = (9, 1)
# b' = 3/(9+u), src/fields/fp2.rs:42-55
FP2_TWIST_CURVE_CONSTANT = fp2_mul((3, 0), fp2_inv(XI))

# ----------------------------------------------------------------------------------------------
# Fp6 = Fp2[v]/(v^3 - xi)                                         src/fields/fp6.rs
# ----------------------------------------------------------------------------------------------
FP6_ZERO = (FP2_ZERO, FP2_ZERO, FP2_ZERO)
FP6_ONE = (FP2_ONE, FP2_ZERO, FP2_ZERO)

# Frobenius tables recomputed from their definitions (comments at src/fields/fp6.rs:40-179 and
# src/fields/fp12.rs:29-172); tests/test_oracle_golden.py asserts them against the literals.
FROBENIUS_COEFF_FP6_C1 = [fp2_pow(XI, (P**i - 1) // 3) for i in range(6)]
FROBENIUS_COEFF_FP6_C2 = [fp2_pow(XI, (2 * P**i - 2) // 3) for i in range(6)]
FROBENIUS_COEFF_FP12_C1 = [fp2_pow(XI, (P**i - 1) // 6) for i in range(12)]
# psi constants, src/groups/g2.rs:80-109
EPS_EXP0 = fp2_pow(XI, (P - 1) // 3)
EPS_EXP1 = fp2_pow(XI, (P - 1) // 2)


def fp6_add(a, b):
    return tuple(fp2_add(x, y) for x, y in zip(a, b))


def fp6_sub(a, b):
    return tuple(fp2_sub(x, y) for x, y in zip(a, b))


def fp6_neg(a):
    return tuple(fp2_neg(x) for x in a)


def fp6_mul(a, b):
    """Value-equal to the 36-mul schoolbook of src/fields/fp6.rs:267-368 (written here in the
    compact form the reference quotes in its comment at :274-283)."""
    t0 = fp2_mul(a[0], b[0])
    t1 = fp2_mul(a[1], b[1])
    t2 = fp2_mul(a[2], b[2])
    c0 = fp2_add(
        fp2_residue_mul(fp2_sub(fp2_sub(fp2_mul(fp2_add(a[1], a[2]), fp2_add(b[1], b[2])), t1), t2)), t0
    )
    c1 = fp2_add(
        fp2_sub(fp2_sub(fp2_mul(fp2_add(a[0], a[1]), fp2_add(b[0], b[1])), t0), t1), fp2_residue_mul(t2)
    )
    c2 = fp2_sub(fp2_add(fp2_sub(fp2_mul(fp2_add(a[0], a[2]), fp2_add(b[0], b[2])), t0), t1), t2)
    return (c0, c1, c2)


def fp6_sqr(a):
    """CH-SQR, src/fields/fp6.rs:219-236."""
    t0 = fp2_sqr(a[0])
    cross = fp2_mul(a[0], a[1])
    t1 = fp2_add(cross, cross)
    t2 = fp2_sqr(fp2_add(fp2_sub(a[0], a[1]), a[2]))
    bc = fp2_mul(a[1], a[2])
    s3 = fp2_add(bc, bc)
    s4 = fp2_sqr(a[2])
    return (
        fp2_add(t0, fp2_residue_mul(s3)),
        fp2_add(t1, fp2_residue_mul(s4)),
        fp2_sub(fp2_sub(fp2_add(fp2_add(t1, t2), s3), t0), s4),
    )


def fp6_residue_mul(a):
    """multiplication by v, src/fields/fp6.rs:189-191."""
    return (fp2_residue_mul(a[2]), a[0], a[1])


def fp6_scale(a, k):
    """scale by an Fp2 factor, src/fields/extensions.rs:86-94."""
    return tuple(fp2_mul(x, k) for x in a)


def fp6_frobenius(a, e: int):
    """src/fields/fp6.rs:203-209."""
    return (
        fp2_frobenius(a[0], e),
        fp2_mul(fp2_frobenius(a[1], e), FROBENIUS_COEFF_FP6_C1[e % 6]),
        fp2_mul(fp2_frobenius(a[2], e), FROBENIUS_COEFF_FP6_C2[e % 6]),
    )


def fp6_inv(a):
    """Alg. 17 of eprint 2010/354, src/fields/fp6.rs:400-424."""
    t0 = fp2_sub(fp2_sqr(a[0]), fp2_mul(a[1], fp2_residue_mul(a[2])))
    t1 = fp2_sub(fp2_residue_mul(fp2_sqr(a[2])), fp2_mul(a[0], a[1]))
    t2 = fp2_sub(fp2_sqr(a[1]), fp2_mul(a[0], a[2]))
    inverse = fp2_inv(
        fp2_add(fp2_residue_mul(fp2_add(fp2_mul(a[2], t1), fp2_mul(a[1], t2))), fp2_mul(a[0], t0))
    )
    return (fp2_mul(inverse, t0), fp2_mul(inverse, t1), fp2_mul(inverse, t2))


# ----------------------------------------------------------------------------------------------
# Fp12 = Fp6[w]/(w^2 - v)                                         src/fields/fp12.rs
# ----------------------------------------------------------------------------------------------
FP12_ONE = (FP6_ONE, FP6_ZERO)


def fp12_mul(a, b):
    """Karatsuba, src/fields/fp12.rs:210-239."""
    t0 = fp6_mul(a[0], b[0])
    t1 = fp6_mul(a[1], b[1])
    return (
        fp6_add(fp6_residue_mul(t1), t0),
        fp6_sub(fp6_sub(fp6_mul(fp6_add(a[0], a[1]), fp6_add(b[0], b[1])), t0), t1),
    )


def fp12_sqr(a):
    """complex squaring, src/fields/fp12.rs:536-550."""
    c0 = fp6_sub(a[0], a[1])
    c3 = fp6_sub(a[0], fp6_residue_mul(a[1]))
    c2 = fp6_mul(a[0], a[1])
    c0 = fp6_add(fp6_mul(c0, c3), c2)
    c1 = fp6_add(c2, c2)
    c2 = fp6_residue_mul(c2)
    return (fp6_add(c0, c2), c1)


def fp12_conj(a):
    """unitary_inverse, src/fields/fp12.rs:381-383."""
    return (a[0], fp6_neg(a[1]))


def fp12_inv(a):
    """Alg. 23 of eprint 2010/354, src/fields/fp12.rs:270-287."""
    tmp = fp6_inv(fp6_sub(fp6_sqr(a[0]), fp6_residue_mul(fp6_sqr(a[1]))))
    return (fp6_mul(a[0], tmp), fp6_neg(fp6_mul(a[1], tmp)))


def fp12_frobenius(a, e: int):
    """src/fields/fp12.rs:515-522."""
    return (fp6_frobenius(a[0], e), fp6_scale(fp6_frobenius(a[1], e), FROBENIUS_COEFF_FP12_C1[e % 12]))


def fp12_sparse_mul(f, ell_0, ell_vw, ell_vv):
    """mul_by_024, line-for-line value restatement of src/fields/fp12.rs:426-503."""
    z0, z1, z2 = f[0]
    z3, z4, z5 = f[1]
    x0, x2, x4 = ell_0, ell_vv, ell_vw
    d0 = fp2_mul(z0, x0)
    d2 = fp2_mul(z2, x2)
    d4 = fp2_mul(z4, x4)
    t2 = fp2_add(z0, z4)
    t1 = fp2_add(z0, z2)
    s0 = fp2_add(fp2_add(z1, z3), z5)
    s1 = fp2_mul(z1, x2)
    t3 = fp2_add(s1, d4)
    t4 = fp2_add(fp2_residue_mul(t3), d0)
    r0 = t4
    t3 = fp2_mul(z5, x4)
    s1 = fp2_add(s1, t3)
    t3 = fp2_add(t3, d2)
    t4 = fp2_residue_mul(t3)
    t3 = fp2_mul(z1, x0)
    s1 = fp2_add(s1, t3)
    t4 = fp2_add(t4, t3)
    r1 = t4
    t0 = fp2_add(x0, x2)
    t3 = fp2_sub(fp2_sub(fp2_mul(t1, t0), d0), d2)
    t4 = fp2_mul(z3, x4)
    s1 = fp2_add(s1, t4)
    t3 = fp2_add(t3, t4)
    t0 = fp2_add(z2, z4)
    r2 = t3
    t1 = fp2_add(x2, x4)
    t3 = fp2_sub(fp2_sub(fp2_mul(t0, t1), d2), d4)
    t4 = fp2_residue_mul(t3)
    t3 = fp2_mul(z3, x0)
    s1 = fp2_add(s1, t3)
    t4 = fp2_add(t4, t3)
    r3 = t4
    t3 = fp2_mul(z5, x2)
    s1 = fp2_add(s1, t3)
    t4 = fp2_residue_mul(t3)
    t0 = fp2_add(x0, x4)
    t3 = fp2_sub(fp2_sub(fp2_mul(t2, t0), d0), d4)
    t4 = fp2_add(t4, t3)
    r4 = t4
    t0 = fp2_add(fp2_add(x0, x2), x4)
    t3 = fp2_sub(fp2_mul(s0, t0), s1)
    r5 = t3
    return ((r0, r1, r2), (r3, r4, r5))


def fp12_to_list(a):
    """Tower order c0.c0.c0, c0.c0.c1, ..., c1.c2.c1 (src/fields/fp12.rs:561-574)."""
    return [c for six in a for two in six for c in two]


def fp12_from_list(v):
    return (
        ((v[0], v[1]), (v[2], v[3]), (v[4], v[5])),
        ((v[6], v[7]), (v[8], v[9]), (v[10], v[11])),
    )


# ----------------------------------------------------------------------------------------------
# generic projective group law (complete, j=0)                  src/groups/group.rs:339-386,528-667
# The field is abstracted by a small ops table so the same code serves G1 (Fp) and G2 (Fp2),
# as GroupProjective<D,N,F> does.
# ----------------------------------------------------------------------------------------------
class _FpOps:
    zero, one = 0, 1
    add = staticmethod(lambda a, b: (a + b) % P)
    sub = staticmethod(lambda a, b: (a - b) % P)
    neg = staticmethod(lambda a: -a % P)
    mul = staticmethod(lambda a, b: a * b % P)
    inv = staticmethod(fp_inv)
    is_zero = staticmethod(lambda a: a % P == 0)
    b3 = 9  # F::from(3) * curve_constant (3), src/groups/group.rs:358


class _Fp2Ops:
    zero, one = FP2_ZERO, FP2_ONE
    add = staticmethod(fp2_add)
    sub = staticmethod(fp2_sub)
    neg = staticmethod(fp2_neg)
    mul = staticmethod(fp2_mul)
    inv = staticmethod(fp2_inv)
    is_zero = staticmethod(lambda a: a == FP2_ZERO)
    b3 = fp2_mul((3, 0), FP2_TWIST_CURVE_CONSTANT)


def proj_zero(F):
    """src/groups/group.rs:320-331."""
    return (F.zero, F.one, F.zero)


def proj_double(F, pt):
    """Alg. 9 of eprint 2015/1060, src/groups/group.rs:339-386."""
    x, y, z = pt
    t0 = F.mul(y, y)
    z3 = F.add(t0, t0)
    z3 = F.add(z3, z3)
    z3 = F.add(z3, z3)
    t1 = F.mul(y, z)
    t2 = F.mul(z, z)
    t2 = F.mul(F.b3, t2)
    x3 = F.mul(t2, z3)
    y3 = F.add(t0, t2)
    z3 = F.mul(t1, z3)
    t1 = F.add(t2, t2)
    t2 = F.add(t1, t2)
    t0 = F.sub(t0, t2)
    y3 = F.mul(t0, y3)
    y3 = F.add(x3, y3)
    t1 = F.mul(x, y)
    x3 = F.mul(t0, t1)
    x3 = F.add(x3, x3)
    if F.is_zero(z):
        return proj_zero(F)
    return (x3, y3, z3)


def proj_add(F, a, b):
    """Alg. 7 of eprint 2015/1060, src/groups/group.rs:528-599."""
    x1, y1, z1 = a
    x2, y2, z2 = b
    t0 = F.mul(x1, x2)
    t1 = F.mul(y1, y2)
    t2 = F.mul(z1, z2)
    t3 = F.add(x1, y1)
    t4 = F.add(x2, y2)
    t3 = F.mul(t3, t4)
    t4 = F.add(t0, t1)
    t3 = F.sub(t3, t4)
    t4 = F.add(y1, z1)
    x3 = F.add(y2, z2)
    t4 = F.mul(t4, x3)
    x3 = F.add(t1, t2)
    t4 = F.sub(t4, x3)
    x3 = F.add(x1, z1)
    y3 = F.add(x2, z2)
    x3 = F.mul(x3, y3)
    y3 = F.add(t0, t2)
    y3 = F.sub(x3, y3)
    x3 = F.add(t0, t0)
    t0 = F.add(x3, t0)
    t2 = F.mul(F.b3, t2)
    z3 = F.add(t1, t2)
    t1 = F.sub(t1, t2)
    y3 = F.mul(F.b3, y3)
    x3 = F.mul(t4, y3)
    t2 = F.mul(t3, t1)
    x3 = F.sub(t2, x3)
    y3 = F.mul(y3, t0)
    t1 = F.mul(t1, z3)
    y3 = F.add(t1, y3)
    t0 = F.mul(t0, t3)
    z3 = F.mul(z3, t4)
    z3 = F.add(z3, t0)
    return (x3, y3, z3)


def proj_neg(F, a):
    return (a[0], F.neg(a[1]), a[2])


def proj_mul(F, pt, k: int):
    """NAF double-and-add over 256 digits, scalar is an Fp value (NOT reduced mod r),
    src/groups/group.rs:639-667."""
    np_, nm = fp_compute_naf(k)
    res = proj_zero(F)
    neg = proj_neg(F, pt)
    for i in reversed(range(256)):
        res = proj_double(F, res)
        if (np_ >> i) & 1:
            res = proj_add(F, res, pt)
        elif (nm >> i) & 1:
            res = proj_add(F, res, neg)
    return res


def proj_to_affine(F, pt):
    """Returns (x, y, infinity); infinity is (0, 1, True).  src/groups/group.rs:475-495."""
    inv = F.inv(pt[2])
    if F.is_zero(inv):
        return (F.zero, F.one, True)
    return (F.mul(pt[0], inv), F.mul(pt[1], inv), False)


def affine_to_proj(F, a):
    """src/groups/group.rs:508-517."""
    return (a[0], a[1], F.zero if a[2] else F.one)


def proj_eq(F, a, b) -> bool:
    """cross-multiplied equality, src/groups/group.rs:426-447."""
    az, bz = F.is_zero(a[2]), F.is_zero(b[2])
    if az or bz:
        return az and bz
    return F.mul(a[0], b[2]) == F.mul(b[0], a[2]) and F.mul(a[1], b[2]) == F.mul(b[1], a[2])


FpOps, Fp2Ops = _FpOps, _Fp2Ops

G1_GEN = (1, 2, False)  # src/groups/g1.rs:54-60
# src/groups/g2.rs:47-77 (decimal form: src/sage_reference/g2.sage:5-7)
G2_GEN = (
    (
        10857046999023057135944570762232829481370756359578518086990519993285655852781,
        11559732032986387107991004021392285783925812861821192530917403151452391805634,
    ),
    (
        8495653923123431417604973247489272438418190587263600148770280649306958101930,
        4082367875863433681332203403145435568316851327593401208105741076214120093531,
    ),
    False,
)


def g1_is_on_curve(x: int, y: int) -> bool:
    return (y * y - x * x * x - 3) % P == 0


def g2_is_on_curve(x, y) -> bool:
    return fp2_sub(fp2_sqr(y), fp2_add(fp2_mul(fp2_sqr(x), x), FP2_TWIST_CURVE_CONSTANT)) == FP2_ZERO


def g2_endomorphism(q):
    """psi on an affine point, src/groups/g2.rs:140-152."""
    if q[2]:
        return q
    return (fp2_mul(EPS_EXP0, fp2_conj(q[0])), fp2_mul(EPS_EXP1, fp2_conj(q[1])), False)


def g2_affine_neg(q):
    """src/groups/group.rs:208-216: y = select(-y, one, infinity)."""
    return (q[0], FP2_ONE if q[2] else fp2_neg(q[1]), q[2])


def g1_affine_neg(p):
    return (p[0], 1 if p[2] else -p[1] % P, p[2])


# ----------------------------------------------------------------------------------------------
# pairing                                                         src/pairing.rs
# ----------------------------------------------------------------------------------------------
def g2_doubling_step(r):
    """src/pairing.rs:798-818.  r = [x, y, z] (mutable list of Fp2); returns the Ell triple."""
    x, y, z = r
    a = fp2_scale(fp2_mul(x, y), TWO_INV)
    b = fp2_sqr(y)
    c = fp2_sqr(z)
    d = fp2_add(fp2_add(c, c), c)
    e = fp2_mul(FP2_TWIST_CURVE_CONSTANT, d)
    f = fp2_add(fp2_add(e, e), e)
    g = fp2_scale(fp2_add(b, f), TWO_INV)
    h = fp2_sub(fp2_sqr(fp2_add(y, z)), fp2_add(b, c))
    i = fp2_sub(e, b)
    j = fp2_sqr(x)
    e_sq = fp2_sqr(e)
    r[0] = fp2_mul(a, fp2_sub(b, f))
    r[1] = fp2_sub(fp2_sqr(g), fp2_add(fp2_add(e_sq, e_sq), e_sq))
    r[2] = fp2_mul(b, h)
    return (fp2_residue_mul(i), fp2_neg(h), fp2_add(fp2_add(j, j), j))


def g2_addition_step(r, base):
    """src/pairing.rs:756-772.  base = affine (x, y, inf)."""
    x, y, z = r
    d = fp2_sub(x, fp2_mul(z, base[0]))
    e = fp2_sub(y, fp2_mul(z, base[1]))
    f = fp2_sqr(d)
    g = fp2_sqr(e)
    h = fp2_mul(d, f)
    i = fp2_mul(x, f)
    j = fp2_sub(fp2_add(fp2_mul(z, g), h), fp2_add(i, i))
    r[0] = fp2_mul(d, j)
    r[1] = fp2_sub(fp2_mul(e, fp2_sub(i, j)), fp2_mul(h, y))
    r[2] = fp2_mul(z, h)
    return (
        fp2_residue_mul(fp2_sub(fp2_mul(e, base[0]), fp2_mul(d, base[1]))),
        d,
        fp2_neg(e),
    )


def g2_precompute(q):
    """87 line-coefficient triples, src/pairing.rs:676-708."""
    r = list(affine_to_proj(Fp2Ops, q))
    coeffs = []
    q_neg = g2_affine_neg(q)
    for digit in ATE_LOOP_COUNT_NAF:
        coeffs.append(g2_doubling_step(r))
        if digit == 1:
            coeffs.append(g2_addition_step(r, q))
        elif digit == -1:
            coeffs.append(g2_addition_step(r, q_neg))
    q1 = g2_endomorphism(q)
    q2 = g2_affine_neg(g2_endomorphism(q1))
    coeffs.append(g2_addition_step(r, q1))
    coeffs.append(g2_addition_step(r, q2))
    assert len(coeffs) == 87
    return coeffs


def _ell_eval(c, g1):
    return c[0], fp2_scale(c[1], g1[1]), fp2_scale(c[2], g1[0])


def miller_loop(coeffs, g1):
    """G2PreComputed::miller_loop, src/pairing.rs:590-619."""
    f = FP12_ONE
    idx = 0
    for digit in ATE_LOOP_COUNT_NAF:
        f = fp12_sparse_mul(fp12_sqr(f), *_ell_eval(coeffs[idx], g1))
        idx += 1
        if digit != 0:
            f = fp12_sparse_mul(f, *_ell_eval(coeffs[idx], g1))
            idx += 1
    f = fp12_sparse_mul(f, *_ell_eval(coeffs[idx], g1))
    idx += 1
    f = fp12_sparse_mul(f, *_ell_eval(coeffs[idx], g1))
    return f


def glued_miller_loop(precomps, g1s):
    """src/pairing.rs:970-1022 (zip truncation included)."""
    pairs = list(zip(precomps, g1s))
    f = FP12_ONE
    idx = 0
    for digit in ATE_LOOP_COUNT_NAF:
        f = fp12_sqr(f)
        for c, g1 in pairs:
            f = fp12_sparse_mul(f, *_ell_eval(c[idx], g1))
        idx += 1
        if digit != 0:
            for c, g1 in pairs:
                f = fp12_sparse_mul(f, *_ell_eval(c[idx], g1))
            idx += 1
    for c, g1 in pairs:
        f = fp12_sparse_mul(f, *_ell_eval(c[idx], g1))
    idx += 1
    for c, g1 in pairs:
        f = fp12_sparse_mul(f, *_ell_eval(c[idx], g1))
    return f


def _fp4_square(a, b):
    """src/pairing.rs:274-289."""
    t0 = fp2_sqr(a)
    t1 = fp2_sqr(b)
    c0 = fp2_add(fp2_residue_mul(t1), t0)
    c1 = fp2_sub(fp2_sub(fp2_sqr(fp2_add(a, b)), t0), t1)
    return c0, c1


def cyclotomic_squared(f):
    """Granger-Scott, src/pairing.rs:309-346."""
    z0, z4, z3 = f[0]
    z2, z1, z5 = f[1]
    t0, t1 = _fp4_square(z0, z1)
    z0 = fp2_sub(t0, z0)
    z0 = fp2_add(fp2_add(z0, z0), t0)
    z1 = fp2_add(t1, z1)
    z1 = fp2_add(fp2_add(z1, z1), t1)
    t0, t1 = _fp4_square(z2, z3)
    t2, t3 = _fp4_square(z4, z5)
    z4 = fp2_sub(t0, z4)
    z4 = fp2_add(fp2_add(z4, z4), t0)
    z5 = fp2_add(t1, z5)
    z5 = fp2_add(fp2_add(z5, z5), t1)
    t0 = fp2_residue_mul(t3)
    z2 = fp2_add(t0, z2)
    z2 = fp2_add(fp2_add(z2, z2), t0)
    z3 = fp2_sub(t2, z3)
    z3 = fp2_add(fp2_add(z3, z3), t2)
    return ((z0, z4, z3), (z2, z1, z5))


def cyclotomic_exp(f, e: int):
    """src/pairing.rs:366-378.  The reference iterates 256 bits; leading zero bits only square
    the value 1, so starting at the top set bit is exact."""
    res = FP12_ONE
    for i in reversed(range(e.bit_length())):
        res = cyclotomic_squared(res)
        if (e >> i) & 1:
            res = fp12_mul(res, f)
    return res


def exp_by_neg_z(f):
    """src/pairing.rs:390-392."""
    return fp12_conj(cyclotomic_exp(f, BLS_X))


def final_exponentiation(f):
    """src/pairing.rs:245-492 (easy_part :410-415, hard_part :437-489)."""
    f1 = fp12_conj(f)
    f2 = fp12_inv(f)
    f = fp12_mul(f1, f2)
    inp = fp12_mul(fp12_frobenius(f, 2), f)
    a = exp_by_neg_z(inp)
    b = cyclotomic_squared(a)
    c = cyclotomic_squared(b)
    d = fp12_mul(c, b)
    e = exp_by_neg_z(d)
    f_ = cyclotomic_squared(e)
    g = exp_by_neg_z(f_)
    h = fp12_conj(d)
    i = fp12_conj(g)
    j = fp12_mul(i, e)
    k = fp12_mul(j, h)
    l = fp12_mul(k, b)
    m = fp12_mul(k, e)
    n = fp12_mul(inp, m)
    o = fp12_frobenius(l, 1)
    p_ = fp12_mul(o, n)
    q = fp12_frobenius(k, 2)
    r = fp12_mul(q, p_)
    s = fp12_conj(inp)
    t = fp12_mul(s, l)
    u = fp12_frobenius(t, 3)
    return fp12_mul(u, r)


def pairing_affine(p, q):
    """`pairing` on already-affine inputs (x, y, inf), src/pairing.rs:870-893."""
    either_zero = p[2] or q[2]
    if either_zero:
        return final_exponentiation(FP12_ONE)
    return final_exponentiation(miller_loop(g2_precompute(q), p))


def pairing(p_proj, q_proj):
    """src/pairing.rs:870-893 on projective inputs."""
    return pairing_affine(proj_to_affine(FpOps, p_proj), proj_to_affine(Fp2Ops, q_proj))


def glued_pairing(g1s_affine, g2s_affine):
    """src/pairing.rs:1029-1037 (affine inputs)."""
    return final_exponentiation(glued_miller_loop([g2_precompute(q) for q in g2s_affine], g1s_affine))


# ----------------------------------------------------------------------------------------------
# hash to curve                              src/hasher.rs, src/svdw.rs, src/groups/g1.rs:307-331
# ----------------------------------------------------------------------------------------------
_KECCAK_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_KECCAK_ROT = [
    [0, 36, 3, 41, 18],
    [1, 44, 10, 45, 2],
    [62, 6, 43, 15, 61],
    [28, 55, 25, 21, 56],
    [27, 20, 39, 8, 14],
]
_M64 = (1 << 64) - 1


def _rol(x, n):
    return ((x << n) | (x >> (64 - n))) & _M64 if n else x


def keccak_f1600(a):
    """Keccak-f[1600] on a 5x5 list a[x][y] (FIPS 202 section 3)."""
    for rc in _KECCAK_RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _KECCAK_ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    return a


def keccak256(data: bytes) -> bytes:
    """Legacy Keccak-256 (pad 0x01, rate 136) == sha3::Keccak256 used at src/lib.rs:181,225."""
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i: off + 8 * i + 8], "little")
        a = keccak_f1600(a)
    out = b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))
    return out


def _hash_params(hash_id: str):
    if hash_id == "keccak256":
        return keccak256, 32, 136
    if hash_id == "sha256":
        return (lambda b: hashlib.sha256(b).digest()), 32, 64
    raise ValueError(hash_id)


def expand_message_xmd(msg: bytes, dst: bytes, len_in_bytes: int, hash_id: str = "keccak256",
                       security_param: int = SECURITY_BITS) -> bytes:
    """XMDExpander::new + expand_message, src/hasher.rs:157-172,201-250."""
    h, b_in_bytes, r_in_bytes = _hash_params(hash_id)
    if len(dst) > 255:
        dst = h(b"H2C-OVERSIZE-DST-" + dst)
    ell = (len_in_bytes + b_in_bytes - 1) // b_in_bytes
    dst_prime = dst + bytes([len(dst)])
    if 8 * b_in_bytes < 2 * security_param or ell > 255:
        raise ValueError("ExpandMessage")
    msg_prime = bytes(r_in_bytes) + msg + len_in_bytes.to_bytes(2, "big") + b"\x00" + dst_prime
    b0 = h(msg_prime)
    bvals = [h(b0 + b"\x01" + dst_prime)]
    for i in range(1, ell):
        xored = bytes(x ^ y for x, y in zip(b0, bvals[-1]))
        bvals.append(h(xored + bytes([i + 1]) + dst_prime))
    return b"".join(bvals)[:len_in_bytes]


def expand_message_xof(msg: bytes, dst: bytes, len_in_bytes: int, security_param: int = SECURITY_BITS) -> bytes:
    """XOFExpander::<Shake128>::new + expand_message, src/hasher.rs:274-290,312-329."""
    if len(dst) > 255:
        dst = hashlib.shake_128(b"H2C-OVERSIZE-DST-" + dst).digest((2 * security_param + 7) // 8)
    dst_prime = dst + bytes([len(dst)])
    return hashlib.shake_128(msg + len_in_bytes.to_bytes(2, "big") + dst_prime).digest(len_in_bytes)


def expand_message(msg: bytes, dst: bytes, len_in_bytes: int, hash_id: str = "keccak256") -> bytes:
    if hash_id == "shake128":
        return expand_message_xof(msg, dst, len_in_bytes)
    return expand_message_xmd(msg, dst, len_in_bytes, hash_id)


def hash_to_field(msg: bytes, dst: bytes = DST, count: int = 2, size: int = 48,
                  hash_id: str = "keccak256"):
    """Expander::hash_to_field, src/hasher.rs:84-128 (always two outputs, Q11)."""
    exp = expand_message(msg, dst, count * size, hash_id)
    return [int.from_bytes(exp[size * i: size * (i + 1)], "big") % P for i in range(2)]


# SvdW constants for y^2 = x^3 + 3, src/svdw.rs:123-153 (Z found by find_z_svdw :81-104 is 1)
SVDW_Z = 1
SVDW_A, SVDW_B = 0, 3


def _svdw_g(x):
    return (x * x * x + SVDW_A * x + SVDW_B) % P


SVDW_C1 = _svdw_g(SVDW_Z)
SVDW_C2 = -SVDW_Z * fp_inv(2) % P
_c3 = fp_sqrt(-_svdw_g(SVDW_Z) * (3 * SVDW_Z * SVDW_Z + 4 * SVDW_A) % P)
SVDW_C3 = -_c3 % P if fp_sgn0(_c3) == 1 else _c3
SVDW_C4 = 4 * (-_svdw_g(SVDW_Z)) * fp_inv(3 * SVDW_Z * SVDW_Z + 4 * SVDW_A) % P


def svdw_map_to_point(u: int):
    """SvdW::unchecked_map_to_point, src/svdw.rs:180-262."""
    tv1 = u * u % P
    tv1 = tv1 * SVDW_C1 % P
    tv2 = (1 + tv1) % P
    tv1 = (1 - tv1) % P
    tv3 = tv1 * tv2 % P
    tv3 = fp_inv(tv3)
    tv4 = u * tv1 % P
    tv4 = tv4 * tv3 % P
    tv4 = tv4 * SVDW_C3 % P
    x1 = (SVDW_C2 - tv4) % P
    gx1 = _svdw_g(x1)
    e1 = fp_is_square(gx1)
    x2 = (SVDW_C2 + tv4) % P
    gx2 = _svdw_g(x2)
    e2 = fp_is_square(gx2) and not e1
    x3 = tv2 * tv2 % P
    x3 = x3 * tv3 % P
    x3 = x3 * x3 % P
    x3 = x3 * SVDW_C4 % P
    x3 = (x3 + SVDW_Z) % P
    x = x1 if e1 else x3
    x = x2 if e2 else x
    gx = _svdw_g(x)
    y = fp_sqrt(gx)
    if y is None:
        raise ValueError("SvdWError")
    e3 = fp_sgn0(u) == fp_sgn0(y)
    y = y if e3 else -y % P
    return (x, y)


def hash_to_curve_g1(msg: bytes, dst: bytes = DST, hash_id: str = "keccak256"):
    """G1Projective::hash_to_curve, src/groups/g1.rs:307-331.  Returns the projective sum."""
    u0, u1 = hash_to_field(msg, dst, 2, 48, hash_id)
    a = svdw_map_to_point(u0)
    b = svdw_map_to_point(u1)
    return proj_add(FpOps, (a[0], a[1], 1), (b[0], b[1], 1))


def sign(sk: int, msg: bytes):
    """src/lib.rs:179-187."""
    return proj_mul(FpOps, hash_to_curve_g1(msg), sk)


def verify(pk_proj, msg: bytes, sig_proj) -> bool:
    """src/lib.rs:223-236: two full pairings, raw Fp12 equality."""
    hm = hash_to_curve_g1(msg)
    lhs = pairing(sig_proj, affine_to_proj(Fp2Ops, G2_GEN))
    rhs = pairing(hm, pk_proj)
    return lhs == rhs


def verify_batch(pks_affine, msgs, sigs_affine) -> bool:
    """examples/verify_multiple_messages_same_signer.rs:40-60 generalised to per-message keys:
    prod e(sig_i, G2gen) * e(-H(m_i), pk_i) == 1."""
    g1s, g2s = [], []
    for pk, m, s in zip(pks_affine, msgs, sigs_affine):
        hm = proj_to_affine(FpOps, hash_to_curve_g1(m))
        g1s += [s, g1_affine_neg(hm)]
        g2s += [G2_GEN, pk]
    return glued_pairing(g1s, g2s) == FP12_ONE


# ----------------------------------------------------------------------------------------------
# input validation                       src/groups/g1.rs:111-132, g2.rs:279-297,460-525
# ----------------------------------------------------------------------------------------------
def g1_affine_new(x: int, y: int) -> str:
    """G1Affine::new: 'ok' or 'NotOnCurve' (every curve point is in the r-torsion), g1.rs:111-132."""
    return "ok" if (y * y - x * x * x) % P == 3 else "NotOnCurve"


def g2_proj_endomorphism(pt):
    """G2Projective::endomorphism goes through affine coordinates, g2.rs:207-210."""
    return affine_to_proj(Fp2Ops, g2_endomorphism(proj_to_affine(Fp2Ops, pt)))


def g2_projective_new(x, y, z=FP2_ONE) -> str:
    """G2Projective::new: curve check then the subgroup relation
    (x+1)Q + psi(xQ) + psi^2(xQ) == psi^3(2xQ), g2.rs:460-525."""
    lhs = fp2_mul(fp2_sqr(y), z)
    rhs = fp2_add(fp2_mul(fp2_sqr(x), x), fp2_mul(fp2_mul(fp2_sqr(z), z), FP2_TWIST_CURVE_CONSTANT))
    if not (lhs == rhs or z == FP2_ZERO):
        return "NotOnCurve"
    tmp = (x, y, z)
    a = proj_mul(Fp2Ops, tmp, BLS_X)
    b = g2_proj_endomorphism(a)
    a = proj_add(Fp2Ops, a, tmp)
    rhs_ = g2_proj_endomorphism(b)
    lhs_ = proj_add(Fp2Ops, proj_add(Fp2Ops, rhs_, b), a)
    rhs_ = proj_add(Fp2Ops, proj_double(Fp2Ops, g2_proj_endomorphism(rhs_)), proj_neg(Fp2Ops, lhs_))
    return "ok" if Fp2Ops.is_zero(rhs_[2]) else "NotInSubgroup"


def fp2_sqrt(a):
    """Square root in Fp2 for p = 3 mod 4 (complex method); None if a is not a square.  Test helper for
    building points of E'(Fp2) outside the r-torsion (the reference's own Fp2::sqrt is off the path, Q2)."""
    if a == FP2_ZERO:
        return FP2_ZERO
    a1 = fp2_pow(a, (P - 3) // 4)
    alpha = fp2_mul(fp2_mul(a1, a1), a)
    x0 = fp2_mul(a1, a)
    if alpha == (P - 1, 0):
        r = fp2_mul((0, 1), x0)
    else:
        b = fp2_pow(fp2_add(FP2_ONE, alpha), (P - 1) // 2)
        r = fp2_mul(b, x0)
    return r if fp2_sqr(r) == a else None


def gt_mul(g, k: int):
    """`&Gt * &Fr`: NAF square-and-multiply over 256 digits with double = Fp12 square and neg = conjugate,
    src/groups/gt.rs:188-215 (double :268-270, neg :124-126)."""
    np_, nm = fp_compute_naf(k)
    res = FP12_ONE
    neg = fp12_conj(g)
    for i in reversed(range(256)):
        res = fp12_sqr(res)
        if (np_ >> i) & 1:
            res = fp12_mul(res, g)
        elif (nm >> i) & 1:
            res = fp12_mul(res, neg)
    return res


# ---- threshold aggregation (examples/dkg.rs:190-226, examples/threshold_signing.rs:124-155) -------------------
def lagrange_coefficients(ids):
    """[prod_{j != i} x_j * (x_j - x_i)^-1 mod r for i], the fold of dkg.rs:216-226 (Fr::inv(0) = 0)."""
    out = []
    for i, xi in enumerate(ids):
        acc = 1
        for j, xj in enumerate(ids):
            if j == i:
                continue
            d = (xj - xi) % R_ORDER
            acc = acc * (xj % R_ORDER) % R_ORDER * (pow(d, -1, R_ORDER) if d else 0) % R_ORDER
        out.append(acc)
    return out


def threshold_aggregate(ids, sigs):
    """sum_i lambda_i * sig_i over affine G1 points (x, y, inf) -> affine (dkg.rs:190-206)."""
    acc = proj_zero(FpOps)
    for lam, s in zip(lagrange_coefficients(ids), sigs):
        acc = proj_add(FpOps, acc, proj_mul(FpOps, affine_to_proj(FpOps, s), lam))
    return proj_to_affine(FpOps, acc)
