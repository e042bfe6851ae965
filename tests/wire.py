"""Wire-format helpers shared by the tests (C-ABI layouts of include/sylow_b200.h)."""
import random

from oracle import bn254_py as o


def fp_b(x: int) -> bytes:
    return int(x).to_bytes(32, "little")


def b_fp(b: bytes) -> int:
    return int.from_bytes(b, "little")


def g1_b(p) -> bytes:
    return fp_b(p[0]) + fp_b(p[1])


def g2_b(q) -> bytes:
    return fp_b(q[0][0]) + fp_b(q[0][1]) + fp_b(q[1][0]) + fp_b(q[1][1])


def b_g1(b: bytes, inf=False):
    return (b_fp(b[:32]), b_fp(b[32:64]), bool(inf))


def b_g2(b: bytes, inf=False):
    return ((b_fp(b[:32]), b_fp(b[32:64])), (b_fp(b[64:96]), b_fp(b[96:128])), bool(inf))


def fp12_b(f) -> bytes:
    return b"".join(fp_b(c) for c in o.fp12_to_list(f))


def b_fp12(b: bytes):
    return o.fp12_from_list([b_fp(b[32 * i: 32 * i + 32]) for i in range(12)])


def rand_fp12(rng: random.Random):
    return o.fp12_from_list([rng.randrange(o.P) for _ in range(12)])


def rand_g1(rng: random.Random):
    k = rng.randrange(1, o.R_ORDER)
    return o.proj_to_affine(o.FpOps, o.proj_mul(o.FpOps, o.affine_to_proj(o.FpOps, o.G1_GEN), k))


def rand_g2(rng: random.Random):
    k = rng.randrange(1, o.R_ORDER)
    return o.proj_to_affine(o.Fp2Ops, o.proj_mul(o.Fp2Ops, o.affine_to_proj(o.Fp2Ops, o.G2_GEN), k))
