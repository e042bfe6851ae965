// Two cooperating lanes per Miller loop (north_star: "one pairing per thread or per cooperating warp group").
//
// Same values as pairing.cuh::miller_loop (pairing.rs:590-619 fused with :676-708, steps :756-818) - every
// intermediate is a canonical Montgomery residue, so the split changes who computes a product, never its value.
// Lanes 2k and 2k+1 of a warp share one PairSlot in shared memory (accumulator, running G2 point, line, a four-value
// exchange area) and run the SAME instruction stream: `role` only selects operand addresses, so the warp does not
// diverge inside the Fp2 products.  Each operation is cut into independent Fp2 products, half per lane:
//   Fp12 squaring   (complex form, 2 Fp6 products)          : one Fp6 product per lane
//   doubling step   (4 M + 6 S + 4 Fp products)             : S M S S M + 2 Fp products per lane
//   addition step   (11 M + 2 S + 4 Fp products)            : 6 M + 1 S + 2 Fp products (lane 1: 5 M, one slot idle)
//   sparse product  (13 M)                                  : 7 M on lane 0, 6 M on lane 1
// i.e. 18.45 product slots per lane and doubling iteration against 35 on one thread: 0.95 of a perfect split before
// the exchange (values cross through the slot, a lane-pair __syncwarp() on either side).  What the split buys is
// LATENCY: a batch too small to fill the GPU with one thread per pairing (or the partly filled last wave of a
// large one) finishes in about half the time.  Throughput at full occupancy is the one-thread kernel's business.
#pragma once
#include "pairing.cuh"

#if defined(__CUDA_ARCH__)
#define SY_LANE_SYNC() __syncwarp()
#elif defined(SYLOW_HOSTSIM)
extern "C" void sylow_hostsim_lane_sync();  // tests/hostsim: a two-thread barrier
#define SY_LANE_SYNC() sylow_hostsim_lane_sync()
#else
#define SY_LANE_SYNC() ((void)0)
#endif

namespace sylow {

struct alignas(16) PairSlot {
  Fp12 f;    // Miller accumulator
  G2Proj r;  // running point
  Ell l;     // current line, already scaled: (c0, c1 * yP, c2 * xP)
  Fp2 x[4];  // exchange area
};

// f <- f^2, complex squaring (fp12.rs:536-550): lane 0 forms c2 = a0 a1, lane 1 m = (a0 - a1)(a0 - v a1);
// then f = (m + c2 + v c2, 2 c2).
SY_HD_NOINLINE void lanes_fp12_sqr(PairSlot& s, int role) {
  const Fp6 &a0 = s.f.c0, &a1 = s.f.c1;
  Fp6* xq = reinterpret_cast<Fp6*>(&s.x[0]);
  // lane 0: a0 - a1 (published), lane 1: a0 - v a1 (kept)
  Fp6 q;
  q.c1 = fp2_sub(a0.c1, role ? a1.c0 : a1.c1);
  q.c2 = fp2_sub(a0.c2, role ? a1.c1 : a1.c2);
  if (role) {
    q.c0 = fp2_sub_mul_xi(a0.c0, a1.c2);
  } else {
    q.c0 = fp2_sub(a0.c0, a1.c0);
    *xq = q;
  }
  SY_LANE_SYNC();
  Fp6 t = fp6_mul(role ? *xq : a0, role ? q : a1);
  SY_LANE_SYNC();  // every read of f and of the exchange area is complete
  if (!role) {
    *xq = t;
    s.f.c1 = fp6_dbl(t);
  }
  SY_LANE_SYNC();
  if (role) {
    const Fp6& c2 = *xq;
    s.f.c0 = Fp6{fp2_mul_xi_add(c2.c2, fp2_add(t.c0, c2.c0)), fp2_add(fp2_add(t.c1, c2.c1), c2.c0),
                 fp2_add(fp2_add(t.c2, c2.c2), c2.c1)};
  }
  SY_LANE_SYNC();
}

// pairing.rs:798-818 on s.r; leaves the scaled line in s.l
SY_HD_NOINLINE void lanes_doubling_step(PairSlot& s, int role, const Fp& xp, const Fp& yp) {
  G2Proj& r = s.r;
  Fp2 sq1 = fp2_sqr(role ? r.y : r.z);  // lane 0: c = Z^2, lane 1: b = Y^2
  s.x[role] = sq1;
  Fp2 d3, yz, g, f3, bf;
  if (!role) d3 = fp2_mul3(sq1);
  else yz = fp2_add(r.y, r.z);
  Fp2 m2 = fp2_mul(role ? r.x : SY_TAB(kTwistB)[0], role ? r.y : d3);  // lane 0: e, lane 1: 2a = X Y
  Fp2 sq3 = fp2_sqr(role ? yz : m2);                                    // lane 0: e^2, lane 1: (Y + Z)^2
  SY_LANE_SYNC();                                                       // b and c are visible
  if (!role) {
    f3 = fp2_mul3(m2);
    g = fp2_halve(fp2_add(s.x[1], f3));
    bf = fp2_sub(s.x[1], f3);
  } else {
    s.x[2] = fp2_halve(m2);                            // a
    s.x[3] = fp2_sub(sq3, fp2_add(sq1, s.x[0]));       // h
  }
  Fp2 sq4 = fp2_sqr(role ? r.x : g);  // lane 0: g^2, lane 1: j = X^2
  SY_LANE_SYNC();                     // a and h are visible
  Fp2 m5 = fp2_mul(role ? sq1 : s.x[2], role ? s.x[3] : bf);  // lane 0: X3 = a (b - f), lane 1: Z3 = b h
  Fp2 lin = role ? fp2_mul3(sq4) : fp2_neg(s.x[3]);
  Fp2 lm = fp2_mul_fp(lin, role ? xp : yp);  // lane 0: (-h) yP, lane 1: (3 j) xP
  Fp2 c0;
  if (!role) c0 = fp2_mul_xi(fp2_sub(m2, s.x[1]));  // xi (e - b)
  SY_LANE_SYNC();                                   // every read of r and of the exchange area is complete
  if (!role) {
    r.x = m5;
    r.y = fp2_sub(sq4, fp2_mul3(sq3));  // g^2 - 3 e^2
    s.l.c0 = c0;
    s.l.c1 = lm;
  } else {
    r.z = m5;
    s.l.c2 = lm;
  }
  SY_LANE_SYNC();
}

// pairing.rs:756-772 on s.r with the affine point (qx, qy); leaves the scaled line in s.l
SY_HD_NOINLINE void lanes_addition_step(PairSlot& s, int role, const Fp2& qx, const Fp2& qy, const Fp& xp,
                                        const Fp& yp) {
  G2Proj& r = s.r;
  Fp2 de = fp2_sub(role ? r.y : r.x, fp2_mul(r.z, role ? qy : qx));  // lane 0: d, lane 1: e
  Fp2 fg = fp2_sqr(de);                                             // lane 0: f = d^2, lane 1: g = e^2
  Fp2 uw = fp2_mul(de, role ? qx : qy);                             // lane 0: d qy, lane 1: e qx
  Fp2 hz = fp2_mul(role ? r.z : de, fg);                            // lane 0: h = d f, lane 1: Z g
  s.x[role] = hz;
  if (!role) s.x[3] = uw;
  SY_LANE_SYNC();
  Fp2 ih = fp2_mul(role ? s.x[0] : r.x, role ? r.y : fg);  // lane 0: i = X f, lane 1: h Y
  Fp2 c0, j, imj, z3;
  if (role) {
    c0 = fp2_mul_xi(fp2_sub(uw, s.x[3]));  // xi (e qx - d qy)
  } else {
    j = fp2_sub(fp2_add(s.x[1], hz), fp2_dbl(ih));  // Z g + h - 2 i
    s.x[1] = j;
    s.x[2] = ih;
  }
  SY_LANE_SYNC();
  if (role) imj = fp2_sub(s.x[2], s.x[1]);
  Fp2 m6 = fp2_mul(de, role ? imj : j);  // lane 0: X3 = d j, lane 1: e (i - j)
  if (!role) z3 = fp2_mul(r.z, hz);      // Z3 = Z h (lane 1 has no product left)
  Fp2 lin = role ? fp2_neg(de) : de;
  Fp2 lm = fp2_mul_fp(lin, role ? xp : yp);  // lane 0: d yP, lane 1: (-e) xP
  SY_LANE_SYNC();                            // every read of r is complete
  if (!role) {
    r.x = m6;
    r.z = z3;
    s.l.c1 = lm;
  } else {
    r.y = fp2_sub(m6, ih);
    s.l.c0 = c0;
    s.l.c2 = lm;
  }
  SY_LANE_SYNC();
}

// f <- f * (l.c0 + l.c2 v^2 + l.c1 v w): the 13 products of fp12.rs:426-503, 7 on lane 0 and 6 on lane 1.
SY_HD_NOINLINE void lanes_sparse_mul(PairSlot& s, int role) {
  const Fp2 &z0 = s.f.c0.c0, &z1 = s.f.c0.c1, &z2 = s.f.c0.c2, &z3 = s.f.c1.c0, &z4 = s.f.c1.c1, &z5 = s.f.c1.c2;
  const Fp2 &x0 = s.l.c0, &x4 = s.l.c1, &x2 = s.l.c2;
  // lane 0: (z0 + z2, x0 + x2); lane 1: (z2 + z4, x2 + x4), (z0 + z4, x0 + x4), (z1 + z3 + z5, x0 + x2 + x4)
  Fp2 za = fp2_add(role ? z4 : z0, z2), xa = fp2_add(role ? x4 : x0, x2);
  Fp2 zb, xb, zc, xc;
  if (role) {
    zb = fp2_add(z0, z4);
    xb = fp2_add(x0, x4);
    zc = fp2_add(fp2_add(z1, z3), z5);
    xc = fp2_add(xb, x2);
  }
  Fp2 t0 = fp2_mul(role ? z4 : z0, role ? x4 : x0);  // d0              | d4
  Fp2 t1 = fp2_mul(role ? za : z2, role ? xa : x2);  // d2              | (z2+z4)(x2+x4)
  Fp2 t2 = fp2_mul(role ? z3 : z1, role ? x0 : x2);  // z1 x2           | z3 x0
  Fp2 t3 = fp2_mul(z5, role ? x2 : x4);              // z5 x4           | z5 x2
  Fp2 t4 = fp2_mul(role ? zb : z1, role ? xb : x0);  // z1 x0           | (z0+z4)(x0+x4)
  Fp2 t5 = fp2_mul(role ? zc : za, role ? xc : xa);  // (z0+z2)(x0+x2)  | (z1+z3+z5)(x0+x2+x4)
  Fp2 t6;
  if (!role) {
    t6 = fp2_mul(z3, x4);
    s.x[0] = t0;
    s.x[1] = t1;
    s.x[2] = fp2_add(fp2_add(t2, t3), fp2_add(t4, t6));
  } else {
    s.x[3] = t0;
  }
  SY_LANE_SYNC();  // every read of f is complete; d0, d2, the lane-0 sum and d4 are visible
  if (!role) {
    s.f.c0 = Fp6{fp2_mul_xi_add(fp2_add(t2, s.x[3]), t0), fp2_mul_xi_add(fp2_add(t3, t1), t4),
                 fp2_add(fp2_sub(fp2_sub(t5, t0), t1), t6)};
  } else {
    s.f.c1 = Fp6{fp2_mul_xi_add(fp2_sub(fp2_sub(t1, s.x[1]), t0), t2),
                 fp2_mul_xi_add(t3, fp2_sub(fp2_sub(t4, s.x[0]), t0)),
                 fp2_sub(fp2_sub(fp2_sub(t5, s.x[2]), t2), t3)};
  }
  SY_LANE_SYNC();
}

// Fused precompute + miller_loop of one (P, Q) pair on two lanes; the result is left in s.f (Montgomery form).
// Both lanes pass the same P and Q.
SY_HD_NOINLINE void lanes_miller_loop(PairSlot& s, int role, const Fp& xp, const Fp& yp, const Fp2& qx,
                                      const Fp2& qy) {
  if (!role) {
    s.r = G2Proj{qx, qy, fp2_one()};
    s.f.c0 = fp6_one();
  } else {
    s.f.c1 = fp6_zero();
  }
  SY_LANE_SYNC();
  Fp2 nqy = fp2_neg(qy);
  for (int i = 0; i < 64; i++) {
    SY_LOOP_SYNC();
    if (i != 0) lanes_fp12_sqr(s, role);
    lanes_doubling_step(s, role, xp, yp);
    lanes_sparse_mul(s, role);
    int digit = SY_TAB(kAteNaf)[i];
    if (digit != 0) {
      lanes_addition_step(s, role, qx, digit > 0 ? qy : nqy, xp, yp);
      lanes_sparse_mul(s, role);
    }
  }
  // Q1 = psi(Q), Q2 = -psi(Q1): lane 0 forms the x coordinates, lane 1 the y coordinates
  Fp2 c = fp2_conj(role ? qy : qx);
  const Fp2& eps = role ? SY_TAB(kEpsExp1)[0] : SY_TAB(kEpsExp0)[0];
  Fp2 q1 = fp2_mul(eps, c);
  Fp2 c1 = fp2_conj(q1);
  Fp2 q2 = fp2_mul(eps, c1);
  if (role) q2 = fp2_neg(q2);
  s.x[role] = q1;
  s.x[2 + role] = q2;
  SY_LANE_SYNC();
  Fp2 q1x = s.x[0], q1y = s.x[1], q2x = s.x[2], q2y = s.x[3];
  SY_LANE_SYNC();
  lanes_addition_step(s, role, q1x, q1y, xp, yp);
  lanes_sparse_mul(s, role);
  lanes_addition_step(s, role, q2x, q2y, xp, yp);
  lanes_sparse_mul(s, role);
}

// ------------------------------------------------------------------------ final exponentiation on two lanes
// pairing.rs:245-492 with the same chain as pairing.cuh::final_exponentiation_assign.  Lane 0 owns the c0 half of every
// Fp12 value and lane 1 the c1 half; the running value of the exponentiation loop and a two-Fp6 exchange area sit in
// shared memory (FexpHot), the other eight Fp12 slots of the chain in a per-pair scratch record in global memory
// (FexpCold) that both lanes address.  Per lane: an Fp12 product is 9 Fp2 products (18 on one thread), a cyclotomic
// squaring 5 Fp2 squarings (9), the inversion of the easy part is done by both lanes.
struct alignas(16) FexpHot {
  Fp12 res;
  Fp6 x[2];
};
struct alignas(16) FexpCold {
  Fp12 f, A, C, E, G, tab[3];
};

SY_HD void lanes_copy(Fp12& d, const Fp12& a, int role) {
  if (role) d.c1 = a.c1;
  else d.c0 = a.c0;
  SY_LANE_SYNC();
}
SY_HD void lanes_conj(Fp12& a, int role) {
  if (role) a.c1 = fp6_neg(a.c1);
  SY_LANE_SYNC();
}

// a <- a b (or a conj(b)); b must not alias a.  Lane 0: t0 = a0 b0 and the three diagonal products of
// s = (a0 + a1)(b0 +- b1); lane 1: t1 = +-a1 b1 and the three Karatsuba cross products of s.
SY_HD_NOINLINE void lanes_fp12_mul(Fp6* x, Fp12& a, const Fp12& b, bool conj_b, int role) {
  if (role) x[1] = conj_b ? fp6_sub(b.c0, b.c1) : fp6_add(b.c0, b.c1);
  else x[0] = fp6_add(a.c0, a.c1);
  Fp6 t = fp6_mul(role ? a.c1 : a.c0, role ? b.c1 : b.c0);
  if (role && conj_b) t = fp6_neg(t);
  SY_LANE_SYNC();  // the two sums are visible
  const Fp6 &sa = x[0], &sb = x[1];
  Fp2 ua, ub, va, vb, wa, wb;
  if (role) {
    ua = fp2_add(sa.c1, sa.c2);
    ub = fp2_add(sb.c1, sb.c2);
    va = fp2_add(sa.c0, sa.c1);
    vb = fp2_add(sb.c0, sb.c1);
    wa = fp2_add(sa.c0, sa.c2);
    wb = fp2_add(sb.c0, sb.c2);
  }
  Fp2 p0 = fp2_mul(role ? ua : sa.c0, role ? ub : sb.c0);  // s0 s0' | (s1+s2)(s1'+s2')
  Fp2 p1 = fp2_mul(role ? va : sa.c1, role ? vb : sb.c1);  // s1 s1' | (s0+s1)(s0'+s1')
  Fp2 p2 = fp2_mul(role ? wa : sa.c2, role ? wb : sb.c2);  // s2 s2' | (s0+s2)(s0'+s2')
  SY_LANE_SYNC();                                          // every read of a, b and of the sums is complete
  if (role) {
    x[1] = t;
  } else {
    // what lane 1 must add to (xi P3, P4, P5) - t1 to get s - t0 - t1
    x[0] = Fp6{fp2_sub_mul_xi(fp2_sub(p0, t.c0), fp2_add(p1, p2)),
               fp2_mul_xi_add(p2, fp2_neg(fp2_add(fp2_add(p0, p1), t.c1))),
               fp2_sub(fp2_sub(fp2_sub(p1, p0), p2), t.c2)};
  }
  SY_LANE_SYNC();
  if (role) {
    a.c1 = Fp6{fp2_sub(fp2_mul_xi_add(p0, x[0].c0), t.c0), fp2_sub(fp2_add(p1, x[0].c1), t.c1),
               fp2_sub(fp2_add(p2, x[0].c2), t.c2)};
  } else {
    a.c0 = Fp6{fp2_mul_xi_add(x[1].c2, t.c0), fp2_add(t.c1, x[1].c0), fp2_add(t.c2, x[1].c1)};  // t0 + v t1
  }
  SY_LANE_SYNC();
}

// Granger-Scott squaring (pairing.rs:309-346) in place.  Lane 0: the Fp4 square of (z0, z1) and z4^2, z5^2;
// lane 1: the Fp4 square of (z2, z3) and (z4 + z5)^2.
SY_HD_NOINLINE void lanes_cyclotomic_square(Fp6* x, Fp12& f, int role) {
  Fp2 &z0 = f.c0.c0, &z4 = f.c0.c1, &z3 = f.c0.c2, &z2 = f.c1.c0, &z1 = f.c1.c1, &z5 = f.c1.c2;
  Fp2* xs = reinterpret_cast<Fp2*>(x);
  const Fp2 &pa = role ? z2 : z0, &pb = role ? z3 : z1;
  Fp2 t0 = fp2_sqr(pa);
  Fp2 t1 = fp2_sqr(pb);
  Fp2 t2 = fp2_sqr(fp2_add(pa, pb));
  Fp2 c0 = fp2_mul_xi_add(t1, t0);
  Fp2 c1 = fp2_sub(fp2_sub(t2, t0), t1);
  Fp2 s45, u1, o2;
  if (role) s45 = fp2_add(z4, z5);
  Fp2 u0 = fp2_sqr(role ? s45 : z4);  // z4^2 | (z4 + z5)^2
  if (!role) u1 = fp2_sqr(z5);
  // (z0, z1) <- (3 c0 - 2 z0, 3 c1 + 2 z1) on lane 0; (z4, z5) <- (3 c0 - 2 z4, 3 c1 + 2 z5) on lane 1
  Fp2 o0 = fp2_add(fp2_dbl(fp2_sub(c0, role ? z4 : z0)), c0);
  Fp2 o1 = fp2_add(fp2_dbl(fp2_add(c1, role ? z5 : z1)), c1);
  if (!role) {
    xs[0] = fp2_add(u0, u1);
    Fp2 e0 = fp2_mul_xi_add(u1, u0);  // first half of the Fp4 square of (z4, z5)
    o2 = fp2_add(fp2_dbl(fp2_sub(e0, z3)), e0);
  }
  SY_LANE_SYNC();  // every read of f by the other lane is complete; z4^2 + z5^2 is visible
  if (role) {
    Fp2 e1 = fp2_mul_xi(fp2_sub(u0, xs[0]));  // xi times the second half
    o2 = fp2_add(fp2_dbl(fp2_add(e1, z2)), e1);
    z4 = o0;
    z5 = o1;
    z2 = o2;
  } else {
    z0 = o0;
    z1 = o1;
    z3 = o2;
  }
  SY_LANE_SYNC();
}

// f <- conj(f^x), the width-4 NAF chain of pairing.cuh::exp_by_neg_z_assign
SY_HD_NOINLINE void lanes_exp_by_neg_z(FexpHot& h, Fp12& f, Fp12* tab, int role) {
  Fp12& res = h.res;
  lanes_copy(res, f, role);
  lanes_cyclotomic_square(h.x, res, role);  // f^2
  lanes_copy(tab[0], f, role);
  lanes_fp12_mul(h.x, tab[0], res, false, role);
  lanes_copy(tab[1], tab[0], role);
  lanes_fp12_mul(h.x, tab[1], res, false, role);
  lanes_copy(tab[2], tab[1], role);
  lanes_fp12_mul(h.x, tab[2], res, false, role);
  {
    int i0 = (SY_TAB(kXWnaf4)[0] - 1) >> 1;
    lanes_copy(res, i0 ? tab[i0 - 1] : f, role);
  }
  for (int i = 1; i < SY_XWNAF4_LEN; i++) {
    SY_LOOP_SYNC();
    lanes_cyclotomic_square(h.x, res, role);
    int d = SY_TAB(kXWnaf4)[i];
    if (d != 0) {
      int idx = ((d > 0 ? d : -d) - 1) >> 1;
      lanes_fp12_mul(h.x, res, idx ? tab[idx - 1] : f, d < 0, role);
    }
  }
  lanes_conj(res, role);
  lanes_copy(f, res, role);
}

// r <- frobenius(a, e), e in {1, 2, 3}; r must not alias a (fp6.rs:203-209 / fp12.rs:515-522)
SY_HD_NOINLINE void lanes_frobenius_to(Fp12& r, const Fp12& a, int e, int role) {
  const bool odd = e & 1;
  const Fp2 &c61 = SY_TAB(kFrob6C1)[e], &c62 = SY_TAB(kFrob6C2)[e], &c12 = SY_TAB(kFrob12C1)[e];
  const Fp2 &i1 = role ? a.c1.c1 : a.c0.c1, &i2 = role ? a.c1.c2 : a.c0.c2;
  Fp2 m1 = fp2_mul(odd ? fp2_conj(i1) : i1, c61);
  Fp2 m2 = fp2_mul(odd ? fp2_conj(i2) : i2, c62);
  Fp2 m3 = fp2_mul(role ? m1 : (odd ? fp2_conj(a.c1.c0) : a.c1.c0), c12);
  if (role) {
    r.c1.c1 = m3;
    r.c1.c2 = fp2_mul(m2, c12);
  } else {
    r.c0.c0 = odd ? fp2_conj(a.c0.c0) : a.c0.c0;
    r.c0.c1 = m1;
    r.c0.c2 = m2;
    r.c1.c0 = m3;
  }
  SY_LANE_SYNC();
}

// r <- 1 / a (Alg. 23 of eprint 2010/354, fp12.rs:270-287); r must not alias a.  Both lanes invert the Fp6 norm.
SY_HD_NOINLINE void lanes_fp12_inv(Fp6* x, Fp12& r, const Fp12& a, int role) {
  x[role] = fp6_sqr(role ? a.c1 : a.c0);
  SY_LANE_SYNC();
  Fp6 t = fp6_inv(fp6_sub(x[0], fp6_mul_v(x[1])));
  Fp6 o = fp6_mul(role ? a.c1 : a.c0, t);
  SY_LANE_SYNC();  // the other lane has read both squares
  if (role) r.c1 = fp6_neg(o);
  else r.c0 = o;
  SY_LANE_SYNC();
}

// c.f <- final_exponentiation(c.f) left in c.E  (Montgomery form in and out)
SY_HD_NOINLINE void lanes_final_exponentiation(FexpHot& h, FexpCold& c, int role) {
  Fp12 &f = c.f, &A = c.A, &C = c.C, &E = c.E, &G = c.G;
  SY_LOOP_SYNC();
  lanes_fp12_inv(h.x, A, f, role);
  lanes_conj(f, role);
  lanes_fp12_mul(h.x, f, A, false, role);
  SY_LOOP_SYNC();
  lanes_frobenius_to(A, f, 2, role);
  lanes_fp12_mul(h.x, f, A, false, role);  // inp
  lanes_copy(A, f, role);
  lanes_exp_by_neg_z(h, A, c.tab, role);          // a
  lanes_cyclotomic_square(h.x, A, role);          // b
  lanes_copy(C, A, role);
  lanes_cyclotomic_square(h.x, C, role);          // c
  lanes_fp12_mul(h.x, C, A, false, role);         // d = c b
  lanes_copy(E, C, role);
  lanes_exp_by_neg_z(h, E, c.tab, role);          // e
  lanes_copy(G, E, role);
  lanes_cyclotomic_square(h.x, G, role);          // f
  lanes_exp_by_neg_z(h, G, c.tab, role);          // g
  SY_LOOP_SYNC();
  lanes_conj(G, role);
  lanes_fp12_mul(h.x, G, E, false, role);         // h = conj(g) e
  lanes_fp12_mul(h.x, G, C, true, role);          // k = h conj(d)
  lanes_fp12_mul(h.x, A, G, false, role);         // l = k b
  SY_LOOP_SYNC();
  lanes_fp12_mul(h.x, E, G, false, role);         // m = k e
  lanes_fp12_mul(h.x, E, f, false, role);         // n = inp m
  lanes_frobenius_to(C, A, 1, role);
  lanes_fp12_mul(h.x, E, C, false, role);         // p = frobenius(l, 1) n
  SY_LOOP_SYNC();
  lanes_frobenius_to(C, G, 2, role);
  lanes_fp12_mul(h.x, E, C, false, role);         // r = frobenius(k, 2) p
  lanes_fp12_mul(h.x, A, f, true, role);          // t = conj(inp) l
  lanes_frobenius_to(C, A, 3, role);
  lanes_fp12_mul(h.x, E, C, false, role);         // frobenius(t, 3) r
}

}  // namespace sylow
