"""TEST INFRASTRUCTURE - a second, independent pin for the Shallue-van de Woestijne map.

Plain-Python port of the reference's own Sage specification, /root/reference/src/sage_reference/svdw.sage:1-137
(`sgn0` :5-27, `find_z_svdw` :29-47, `generic_svdw.__init__` :62-85, `map_to_point` :87-137).  Sage is not in the image,
so the few Sage facilities the file uses (a prime-field element type, `is_square`, `sqrt`) are restated here on Python
integers; everything else follows the Sage source line by line, INCLUDING the search for Z and the derivation of
c1..c4 - nothing is copied from src/svdw.rs or from oracle/bn254_py.py, which restates the Rust.

Why it exists: the reference holds no value-level known-answer test for its
Keccak-256 hash-to-curve (the 1000-vector Sage JSON is missing from the mount), so oracle/bn254_py.py was the sole
authority for `svdw_map_to_point`.  tests/test_svdw_second_pin.py checks the two restatements against each other on
10^4 random field elements and on the exceptional inputs (u = 0; tv1 * tv2 = 0, where `inv0` returns 0), checks the
derived constants against the literals the reference's Rust tests hold (src/svdw.rs:285-296), and
tests/golden/make_golden.py records map vectors produced by THIS port in reference_kats.json.
"""
from __future__ import annotations

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # BN254 base field (src/fields/fp.rs:51-56)


class F:
    """Element of GF(P) with the handful of operations svdw.sage uses (`F(x)`, + - * / ^, ==, is_square, sqrt)."""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = (v.v if isinstance(v, F) else int(v)) % P

    def __add__(self, o):
        return F(self.v + F(o).v)

    __radd__ = __add__

    def __sub__(self, o):
        return F(self.v - F(o).v)

    def __rsub__(self, o):
        return F(F(o).v - self.v)

    def __mul__(self, o):
        return F(self.v * F(o).v)

    __rmul__ = __mul__

    def __neg__(self):
        return F(-self.v)

    def __truediv__(self, o):
        d = F(o).v
        if d == 0:
            raise ZeroDivisionError("division by zero in GF(p)")
        return F(self.v * pow(d, -1, P))

    def __rtruediv__(self, o):
        return F(o) / self

    def __pow__(self, e):
        return F(pow(self.v, int(e), P))

    def __eq__(self, o):
        return self.v == F(o).v

    def __hash__(self):
        return hash(self.v)

    def is_square(self):
        # Sage: 0 is a square
        return self.v == 0 or pow(self.v, (P - 1) // 2, P) == 1

    def sqrt(self):
        # p = 3 mod 4; either root will do, the map fixes the sign with sgn0 afterwards (svdw.sage:133-135)
        r = pow(self.v, (P + 1) // 4, P)
        if r * r % P != self.v:
            raise ValueError("not a square")
        return F(r)

    def __int__(self):
        return self.v

    def __repr__(self):
        return "F(0x%x)" % self.v


def CMOV(x, y, b):  # svdw.sage:3
    return y if b else x


def sgn0(x):
    """svdw.sage:5-27 for a prime field (degree 1): the parity of the canonical representative."""
    return int(F(x)) % 2


def find_z_svdw(A, B, init_ctr=1):
    """svdw.sage:29-47"""
    g = lambda x: F(x) ** 3 + F(A) * F(x) + F(B)
    h = lambda Z: -(F(3) * Z ** 2 + F(4) * A) / (F(4) * g(Z))
    ctr = init_ctr
    while True:
        for Z_cand in (F(ctr), F(-ctr)):
            if g(Z_cand) == F(0):
                continue
            if h(Z_cand) == F(0):
                continue
            if not h(Z_cand).is_square():
                continue
            if g(Z_cand).is_square() or g(-Z_cand / F(2)).is_square():
                return Z_cand
        ctr += 1


class GenericSvdW:
    """svdw.sage:51-137 (`generic_svdw`) for the curve y^2 = x^3 + A x + B over GF(P)."""

    def __init__(self, A=0, B=3):
        self.A = F(A)
        self.B = F(B)
        self.Z = find_z_svdw(self.A, self.B)
        self.g = lambda x: F(x) ** 3 + self.A * F(x) + self.B
        mgZ = -self.g(self.Z)
        self.c1 = self.g(self.Z)
        self.c2 = F(-self.Z / F(2))
        self.c3 = (mgZ * (3 * self.Z ** 2 + 4 * self.A)).sqrt()
        if sgn0(self.c3) == 1:
            self.c3 = -self.c3
        assert sgn0(self.c3) == 0
        self.c4 = F(4) * mgZ / (3 * self.Z ** 2 + 4 * self.A)
        # values at which the map is undefined (tv1 * tv2 = 0): svdw.sage:80-85
        self.undefs = []
        for zz in (F(1) / mgZ, F(-1) / mgZ):
            if zz.is_square():
                s = zz.sqrt()
                self.undefs += [s, -s]

    def inv0(self, x):  # svdw.sage:56-59
        if F(x) == 0:
            return F(0)
        return F(1) / F(x)

    def map_to_point(self, u):
        """svdw.sage:87-137, statement by statement."""
        u = F(u)
        inv0, c1, c2, c3, c4, A, B, Z = self.inv0, self.c1, self.c2, self.c3, self.c4, self.A, self.B, self.Z
        tv1 = u ** 2
        tv1 = tv1 * c1
        tv2 = 1 + tv1
        tv1 = 1 - tv1
        tv3 = tv1 * tv2
        tv3 = inv0(tv3)
        tv4 = u * tv1
        tv4 = tv4 * tv3
        tv4 = tv4 * c3
        x1 = c2 - tv4
        gx1 = x1 ** 2
        gx1 = gx1 + A
        gx1 = gx1 * x1
        gx1 = gx1 + B
        e1 = gx1.is_square()
        x2 = c2 + tv4
        gx2 = x2 ** 2
        gx2 = gx2 + A
        gx2 = gx2 * x2
        gx2 = gx2 + B
        e2 = gx2.is_square() and not e1
        x3 = tv2 ** 2
        x3 = x3 * tv3
        x3 = x3 ** 2
        x3 = x3 * c4
        x3 = x3 + Z
        x = CMOV(x3, x1, e1)
        x = CMOV(x, x2, e2)
        gx = x ** 2
        gx = gx + A
        gx = gx * x
        gx = gx + B
        y = gx.sqrt()
        e3 = sgn0(u) == sgn0(y)
        y = CMOV(-y, y, e3)
        return (int(x), int(y))


_INSTANCE = None


def bn254_g1_svdw() -> GenericSvdW:
    global _INSTANCE
    if _INSTANCE is None:
        _INSTANCE = GenericSvdW(0, 3)  # y^2 = x^3 + 3 (src/groups/g1.rs: curve constant 3)
    return _INSTANCE
