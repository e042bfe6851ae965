/* sylow_oracle.c - CPU restatement of sylow's BN254 hot path in plain C.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
 * load or call this file.  The product (libsylow_b200.so) never does.
 *
 * It follows the reference's formulas LITERALLY (that is the point of the CPU baseline): schoolbook
 * Fp2 (fp2.rs:302-305), the 36-product Fp6 schoolbook with its five full multiplications by 9
 * (fp6.rs:317-366), the 87-coefficient precompute followed by the Miller loop (pairing.rs:590-708),
 * the 256-iteration cyclotomic_exp (pairing.rs:366-378) and the 256-iteration NAF scalar
 * multiplication (group.rs:639-667).  What it does NOT reproduce is crypto-bigint's per-operator
 * leave/re-enter of Montgomery form (fp.rs:304-310,387-393): values stay in Montgomery form here, so
 * this baseline is FASTER per core than real sylow - a conservative baseline.
 *
 * Arithmetic: 4 x 64-bit limbs, R = 2^256, unsigned __int128 products.
 * Parity status: pinned against the reference's golden vectors through tests/test_c_oracle.py (Gt
 * generator, pairing test_cases, EIP-196/197, SvdW constants) and against oracle/bn254_py.py.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fp;
typedef struct { fp c0, c1; } fp2;
typedef struct { fp2 c0, c1, c2; } fp6;
typedef struct { fp6 c0, c1; } fp12;

/* fp.rs:51-56 */
static const fp P = {{0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull}};
static const uint64_t INV = 0x87d20782e4866389ull; /* -p^-1 mod 2^64 */
static const fp R2 = {{0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full}};
static const fp R3 = {{0xb1cd6dafda1530dfull, 0x62f210e6a7283db6ull, 0xef7f0b0c0ada0afbull, 0x20fd6e902d592544ull}};

static fp FP_ZERO, FP_ONE, FP_TWO, FP_THREE, FP_FOUR, FP_NINE, TWO_INV;
static fp2 FP2_ZERO_, FP2_ONE_, TWIST_B, EPS_EXP0, EPS_EXP1, G2X, G2Y;
static fp2 FROB6_C1[6], FROB6_C2[6], FROB12_C1[12];
static fp SVDW_C1, SVDW_C2, SVDW_C3, SVDW_C4, SVDW_Z;
static const uint64_t BLS_X = 4965661367192848881ull; /* g2.rs:112 */
/* pairing.rs:26-30 */
static const int8_t ATE_NAF[64] = {1, 0, 1, 0, 0, 0, -1, 0, -1, 0, 0, 0, -1, 0, 1, 0, -1, 0, 0, -1, 0, 0,
                                   0, 0, 0, 1, 0, 0, -1, 0, 1, 0, 0, -1, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0,
                                   -1, 0, -1, 0, 0, 1, 0, 0, 0, -1, 0, 0, -1, 0, 1, 0, 1, 0, 0, 0};

/* ------------------------------------------------------------------------------------------ Fp */
static int fp_geq_p(const fp* a) {
  for (int i = 3; i >= 0; i--) {
    if (a->l[i] > P.l[i]) return 1;
    if (a->l[i] < P.l[i]) return 0;
  }
  return 1;
}
static void fp_sub_p(fp* a) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a->l[i] - P.l[i] - (uint64_t)b;
    a->l[i] = (uint64_t)d;
    b = (d >> 64) & 1;
  }
}
static fp fp_add(fp a, fp b) { /* fp.rs:304-310 */
  fp r;
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a.l[i] + b.l[i];
    r.l[i] = (uint64_t)c;
    c >>= 64;
  }
  if (fp_geq_p(&r)) fp_sub_p(&r);
  return r;
}
static fp fp_sub(fp a, fp b) { /* fp.rs:340-347 */
  fp r;
  u128 bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a.l[i] - b.l[i] - (uint64_t)bw;
    r.l[i] = (uint64_t)d;
    bw = (d >> 64) & 1;
  }
  if (bw) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)r.l[i] + P.l[i];
      r.l[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  return r;
}
static fp fp_neg(fp a) { return fp_sub(FP_ZERO, a); } /* fp.rs:442-449 */
static fp fp_mul(fp a, fp b) {                        /* fp.rs:387-393 (Montgomery CIOS) */
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a.l[j] * b.l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * INV;
    c = (u128)m * P.l[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * P.l[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  fp r = {{t[0], t[1], t[2], t[3]}};
  if (t[4] || fp_geq_p(&r)) fp_sub_p(&r);
  return r;
}
static fp fp_sqr(fp a) { return fp_mul(a, a); } /* fp.rs:620-622 */
static int fp_is_zero(fp a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
static int fp_eq(fp a, fp b) { return memcmp(&a, &b, sizeof(fp)) == 0; }
static fp fp_from_u64(uint64_t v) {
  fp a = {{v, 0, 0, 0}};
  return fp_mul(a, R2);
}
static fp fp_to_mont(fp a) { return fp_mul(a, R2); }
static fp fp_from_mont(fp a) {
  fp one = {{1, 0, 0, 0}};
  return fp_mul(a, one);
}
/* a^e, e = 4 LE words, scanning all 256 bits like crypto-bigint's pow (fp.rs:451-457) */
static fp fp_pow(fp a, const uint64_t e[4]) {
  fp r = FP_ONE;
  for (int w = 3; w >= 0; w--)
    for (int i = 63; i >= 0; i--) {
      r = fp_sqr(r);
      if ((e[w] >> i) & 1) r = fp_mul(r, a);
    }
  return r;
}
static void p_minus(uint64_t k, uint64_t out[4]) { /* p - k for small k */
  u128 b = k;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)P.l[i] - (uint64_t)b;
    out[i] = (uint64_t)d;
    b = (d >> 64) & 1;
  }
}
static void shr(uint64_t a[4], int s) {
  for (int i = 0; i < 4; i++) a[i] = (a[i] >> s) | (i < 3 ? a[i + 1] << (64 - s) : 0);
}
static fp fp_inv(fp a) { /* inv(0) = 0, fp.rs:418-424 */
  uint64_t e[4];
  p_minus(2, e);
  return fp_pow(a, e);
}
static int fp_sqrt(fp a, fp* out) { /* x^((p+1)/4) with post-check, fp.rs:611-616 */
  uint64_t e[4];
  p_minus(0, e);
  e[0] += 1; /* p + 1: p is odd and p.l[0] != 2^64-1 */
  shr(e, 2);
  fp s = fp_pow(a, e);
  *out = s;
  return fp_eq(fp_sqr(s), a);
}
static int fp_is_square(fp a) { /* fp.rs:625-631 */
  uint64_t e[4];
  p_minus(1, e);
  shr(e, 1);
  fp l = fp_pow(a, e);
  return fp_is_zero(l) || fp_eq(l, FP_ONE);
}
static int fp_sgn0(fp a) { return (int)(fp_from_mont(a).l[0] & 1); } /* fp.rs:636-644 */

/* ------------------------------------------------------------------------------------------ Fp2 */
static fp2 fp2_add(fp2 a, fp2 b) { return (fp2){fp_add(a.c0, b.c0), fp_add(a.c1, b.c1)}; }
static fp2 fp2_sub(fp2 a, fp2 b) { return (fp2){fp_sub(a.c0, b.c0), fp_sub(a.c1, b.c1)}; }
static fp2 fp2_neg(fp2 a) { return (fp2){fp_neg(a.c0), fp_neg(a.c1)}; }
static int fp2_is_zero(fp2 a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
static int fp2_eq(fp2 a, fp2 b) { return fp_eq(a.c0, b.c0) && fp_eq(a.c1, b.c1); }
static fp2 fp2_mul(fp2 a, fp2 b) { /* schoolbook, fp2.rs:302-305 */
  return (fp2){fp_sub(fp_mul(a.c0, b.c0), fp_mul(a.c1, b.c1)), fp_add(fp_mul(a.c0, b.c1), fp_mul(a.c1, b.c0))};
}
static fp2 fp2_sqr(fp2 a) { /* fp2.rs:164-171 */
  fp s = fp_add(a.c0, a.c1), d = fp_sub(a.c0, a.c1), c = fp_add(a.c0, a.c0);
  return (fp2){fp_mul(s, d), fp_mul(c, a.c1)};
}
static fp2 fp2_scale(fp2 a, fp k) { return (fp2){fp_mul(a.c0, k), fp_mul(a.c1, k)}; } /* extensions.rs:86-94 */
static fp2 fp2_residue_mul(fp2 a) { /* fp2.rs:99-107: Fp::NINE * x is a full multiplication there */
  return (fp2){fp_sub(fp_mul(FP_NINE, a.c0), a.c1), fp_add(a.c0, fp_mul(FP_NINE, a.c1))};
}
static fp2 fp2_conj(fp2 a) { return (fp2){a.c0, fp_neg(a.c1)}; } /* frobenius(1), fp2.rs:119-133 */
static fp2 fp2_frob(fp2 a, int e) { return (e & 1) ? fp2_conj(a) : a; }
static fp2 fp2_inv(fp2 a) { /* fp2.rs:343-361 */
  fp t = fp_inv(fp_add(fp_sqr(a.c0), fp_sqr(a.c1)));
  return (fp2){fp_mul(a.c0, t), fp_neg(fp_mul(a.c1, t))};
}
static fp2 fp2_pow(fp2 a, const uint64_t e[4]) { /* fp2.rs:62-74 */
  fp2 r = FP2_ONE_;
  for (int w = 3; w >= 0; w--)
    for (int i = 63; i >= 0; i--) {
      r = fp2_mul(r, r);
      if ((e[w] >> i) & 1) r = fp2_mul(r, a);
    }
  return r;
}

/* ------------------------------------------------------------------------------------------ Fp6 */
static fp6 fp6_add(fp6 a, fp6 b) { return (fp6){fp2_add(a.c0, b.c0), fp2_add(a.c1, b.c1), fp2_add(a.c2, b.c2)}; }
static fp6 fp6_sub(fp6 a, fp6 b) { return (fp6){fp2_sub(a.c0, b.c0), fp2_sub(a.c1, b.c1), fp2_sub(a.c2, b.c2)}; }
static fp6 fp6_neg(fp6 a) { return (fp6){fp2_neg(a.c0), fp2_neg(a.c1), fp2_neg(a.c2)}; }
static fp6 fp6_residue_mul(fp6 a) { return (fp6){fp2_residue_mul(a.c2), a.c0, a.c1}; } /* fp6.rs:189-191 */
/* Alg. 5 of eprint 2022/367 exactly as fp6.rs:317-366 writes it */
static fp6 fp6_mul(fp6 s, fp6 o) {
#define M fp_mul
#define A fp_add
#define S fp_sub
  fp a20_m_b21 = S(M(FP_NINE, o.c2.c0), o.c2.c1);
  fp a10_m_b11 = S(M(FP_NINE, o.c1.c0), o.c1.c1);
  fp b21_p_b20 = A(M(FP_NINE, o.c2.c1), o.c2.c0);
  fp b20_m_b21 = S(M(FP_NINE, o.c2.c0), o.c2.c1);
  fp b11_p_b10 = A(M(FP_NINE, o.c1.c1), o.c1.c0);
  fp c00 = S(A(S(A(S(M(s.c0.c0, o.c0.c0), M(s.c0.c1, o.c0.c1)), M(s.c1.c0, a20_m_b21)), M(s.c1.c1, b21_p_b20)),
               M(s.c2.c0, a10_m_b11)), M(s.c2.c1, b11_p_b10));
  fp c01 = A(A(A(A(A(M(s.c0.c0, o.c0.c1), M(s.c0.c1, o.c0.c0)), M(s.c1.c0, b21_p_b20)), M(s.c1.c1, b20_m_b21)),
               M(s.c2.c0, b11_p_b10)), M(s.c2.c1, a10_m_b11));
  fp c10 = S(A(S(A(S(M(s.c0.c0, o.c1.c0), M(s.c0.c1, o.c1.c1)), M(s.c1.c0, o.c0.c0)), M(s.c1.c1, o.c0.c1)),
               M(s.c2.c0, b20_m_b21)), M(s.c2.c1, b21_p_b20));
  fp c11 = A(A(A(A(A(M(s.c0.c0, o.c1.c1), M(s.c0.c1, o.c1.c0)), M(s.c1.c0, o.c0.c1)), M(s.c1.c1, o.c0.c0)),
               M(s.c2.c0, b21_p_b20)), M(s.c2.c1, a20_m_b21));
  fp c20 = S(A(S(A(S(M(s.c0.c0, o.c2.c0), M(s.c0.c1, o.c2.c1)), M(s.c1.c0, o.c1.c0)), M(s.c1.c1, o.c1.c1)),
               M(s.c2.c0, o.c0.c0)), M(s.c2.c1, o.c0.c1));
  fp c21 = A(A(A(A(A(M(s.c0.c0, o.c2.c1), M(s.c0.c1, o.c2.c0)), M(s.c1.c0, o.c1.c1)), M(s.c1.c1, o.c1.c0)),
               M(s.c2.c0, o.c0.c1)), M(s.c2.c1, o.c0.c0));
#undef M
#undef A
#undef S
  return (fp6){{c00, c01}, {c10, c11}, {c20, c21}};
}
static fp6 fp6_sqr(fp6 a) { /* fp6.rs:219-236 */
  fp2 t0 = fp2_sqr(a.c0);
  fp2 cross = fp2_mul(a.c0, a.c1);
  fp2 t1 = fp2_add(cross, cross);
  fp2 t2 = fp2_sqr(fp2_add(fp2_sub(a.c0, a.c1), a.c2));
  fp2 bc = fp2_mul(a.c1, a.c2);
  fp2 s3 = fp2_add(bc, bc);
  fp2 s4 = fp2_sqr(a.c2);
  return (fp6){fp2_add(t0, fp2_residue_mul(s3)), fp2_add(t1, fp2_residue_mul(s4)),
               fp2_sub(fp2_sub(fp2_add(fp2_add(t1, t2), s3), t0), s4)};
}
static fp6 fp6_scale(fp6 a, fp2 k) { return (fp6){fp2_mul(a.c0, k), fp2_mul(a.c1, k), fp2_mul(a.c2, k)}; }
static fp6 fp6_frob(fp6 a, int e) { /* fp6.rs:203-209 */
  return (fp6){fp2_frob(a.c0, e), fp2_mul(fp2_frob(a.c1, e), FROB6_C1[e % 6]), fp2_mul(fp2_frob(a.c2, e), FROB6_C2[e % 6])};
}
static fp6 fp6_inv(fp6 a) { /* fp6.rs:400-424 */
  fp2 t0 = fp2_sub(fp2_sqr(a.c0), fp2_mul(a.c1, fp2_residue_mul(a.c2)));
  fp2 t1 = fp2_sub(fp2_residue_mul(fp2_sqr(a.c2)), fp2_mul(a.c0, a.c1));
  fp2 t2 = fp2_sub(fp2_sqr(a.c1), fp2_mul(a.c0, a.c2));
  fp2 inv = fp2_inv(fp2_add(fp2_residue_mul(fp2_add(fp2_mul(a.c2, t1), fp2_mul(a.c1, t2))), fp2_mul(a.c0, t0)));
  return (fp6){fp2_mul(inv, t0), fp2_mul(inv, t1), fp2_mul(inv, t2)};
}

/* ------------------------------------------------------------------------------------------ Fp12 */
static fp12 FP12_ONE_;
static fp12 fp12_mul(fp12 a, fp12 b) { /* fp12.rs:210-239 */
  fp6 t0 = fp6_mul(a.c0, b.c0), t1 = fp6_mul(a.c1, b.c1);
  return (fp12){fp6_add(fp6_residue_mul(t1), t0),
                fp6_sub(fp6_sub(fp6_mul(fp6_add(a.c0, a.c1), fp6_add(b.c0, b.c1)), t0), t1)};
}
static fp12 fp12_sqr(fp12 a) { /* fp12.rs:536-550 */
  fp6 c0 = fp6_sub(a.c0, a.c1);
  fp6 c3 = fp6_sub(a.c0, fp6_residue_mul(a.c1));
  fp6 c2 = fp6_mul(a.c0, a.c1);
  c0 = fp6_add(fp6_mul(c0, c3), c2);
  fp6 c1 = fp6_add(c2, c2);
  c2 = fp6_residue_mul(c2);
  return (fp12){fp6_add(c0, c2), c1};
}
static fp12 fp12_conj(fp12 a) { return (fp12){a.c0, fp6_neg(a.c1)}; } /* fp12.rs:381-383 */
static fp12 fp12_inv(fp12 a) {                                        /* fp12.rs:270-287 */
  fp6 t = fp6_inv(fp6_sub(fp6_sqr(a.c0), fp6_residue_mul(fp6_sqr(a.c1))));
  return (fp12){fp6_mul(a.c0, t), fp6_neg(fp6_mul(a.c1, t))};
}
static fp12 fp12_frob(fp12 a, int e) { /* fp12.rs:515-522 */
  return (fp12){fp6_frob(a.c0, e), fp6_scale(fp6_frob(a.c1, e), FROB12_C1[e % 12])};
}
static int fp12_eq(fp12 a, fp12 b) { return memcmp(&a, &b, sizeof(fp12)) == 0; }
static fp12 fp12_sparse_mul(fp12 f, fp2 ell_0, fp2 ell_vw, fp2 ell_vv) { /* fp12.rs:426-503 */
  fp2 z0 = f.c0.c0, z1 = f.c0.c1, z2 = f.c0.c2, z3 = f.c1.c0, z4 = f.c1.c1, z5 = f.c1.c2;
  fp2 x0 = ell_0, x2 = ell_vv, x4 = ell_vw;
  fp2 d0 = fp2_mul(z0, x0), d2 = fp2_mul(z2, x2), d4 = fp2_mul(z4, x4);
  fp2 t2 = fp2_add(z0, z4), t1 = fp2_add(z0, z2), s0 = fp2_add(fp2_add(z1, z3), z5);
  fp2 s1 = fp2_mul(z1, x2);
  fp2 t3 = fp2_add(s1, d4);
  fp2 t4 = fp2_add(fp2_residue_mul(t3), d0);
  fp2 r0 = t4;
  t3 = fp2_mul(z5, x4);
  s1 = fp2_add(s1, t3);
  t3 = fp2_add(t3, d2);
  t4 = fp2_residue_mul(t3);
  t3 = fp2_mul(z1, x0);
  s1 = fp2_add(s1, t3);
  t4 = fp2_add(t4, t3);
  fp2 r1 = t4;
  fp2 t0 = fp2_add(x0, x2);
  t3 = fp2_sub(fp2_sub(fp2_mul(t1, t0), d0), d2);
  t4 = fp2_mul(z3, x4);
  s1 = fp2_add(s1, t4);
  t3 = fp2_add(t3, t4);
  t0 = fp2_add(z2, z4);
  fp2 r2 = t3;
  t1 = fp2_add(x2, x4);
  t3 = fp2_sub(fp2_sub(fp2_mul(t0, t1), d2), d4);
  t4 = fp2_residue_mul(t3);
  t3 = fp2_mul(z3, x0);
  s1 = fp2_add(s1, t3);
  t4 = fp2_add(t4, t3);
  fp2 r3 = t4;
  t3 = fp2_mul(z5, x2);
  s1 = fp2_add(s1, t3);
  t4 = fp2_residue_mul(t3);
  t0 = fp2_add(x0, x4);
  t3 = fp2_sub(fp2_sub(fp2_mul(t2, t0), d0), d4);
  t4 = fp2_add(t4, t3);
  fp2 r4 = t4;
  t0 = fp2_add(fp2_add(x0, x2), x4);
  t3 = fp2_sub(fp2_mul(s0, t0), s1);
  return (fp12){{r0, r1, r2}, {r3, r4, t3}};
}

/* ----------------------------------------------------------------------------- groups (group.rs) */
typedef struct { fp x, y, z; } g1p;
typedef struct { fp2 x, y, z; } g2p;
typedef struct { fp x, y; int inf; } g1a;
typedef struct { fp2 x, y; int inf; } g2a;

/* The same generic code instantiated for F = Fp and F = Fp2 (GroupProjective<D,N,F>). */
#define DEFINE_GROUP(T, F, ADD, SUB, MUL, NEG, ISZ, B3EXPR, ZERO, ONE, INVF)                                     \
  static T T##_zero(void) { return (T){ZERO, ONE, ZERO}; }                                                       \
  static T T##_double(T p) { /* group.rs:339-386 */                                                              \
    F t0 = MUL(p.y, p.y);                                                                                        \
    F z3 = ADD(t0, t0);                                                                                          \
    z3 = ADD(z3, z3);                                                                                            \
    z3 = ADD(z3, z3);                                                                                            \
    F t1 = MUL(p.y, p.z);                                                                                        \
    F t2 = MUL(p.z, p.z);                                                                                        \
    t2 = MUL(B3EXPR, t2);                                                                                        \
    F x3 = MUL(t2, z3);                                                                                          \
    F y3 = ADD(t0, t2);                                                                                          \
    z3 = MUL(t1, z3);                                                                                            \
    t1 = ADD(t2, t2);                                                                                            \
    t2 = ADD(t1, t2);                                                                                            \
    t0 = SUB(t0, t2);                                                                                            \
    y3 = MUL(t0, y3);                                                                                            \
    y3 = ADD(x3, y3);                                                                                            \
    t1 = MUL(p.x, p.y);                                                                                          \
    x3 = MUL(t0, t1);                                                                                            \
    x3 = ADD(x3, x3);                                                                                            \
    if (ISZ(p.z)) return T##_zero();                                                                             \
    return (T){x3, y3, z3};                                                                                      \
  }                                                                                                              \
  static T T##_add(T a, T b) { /* group.rs:528-599 */                                                            \
    F t0 = MUL(a.x, b.x), t1 = MUL(a.y, b.y), t2 = MUL(a.z, b.z);                                                \
    F t3 = ADD(a.x, a.y), t4 = ADD(b.x, b.y);                                                                    \
    t3 = MUL(t3, t4);                                                                                            \
    t4 = ADD(t0, t1);                                                                                            \
    t3 = SUB(t3, t4);                                                                                            \
    t4 = ADD(a.y, a.z);                                                                                          \
    F x3 = ADD(b.y, b.z);                                                                                        \
    t4 = MUL(t4, x3);                                                                                            \
    x3 = ADD(t1, t2);                                                                                            \
    t4 = SUB(t4, x3);                                                                                            \
    x3 = ADD(a.x, a.z);                                                                                          \
    F y3 = ADD(b.x, b.z);                                                                                        \
    x3 = MUL(x3, y3);                                                                                            \
    y3 = ADD(t0, t2);                                                                                            \
    y3 = SUB(x3, y3);                                                                                            \
    x3 = ADD(t0, t0);                                                                                            \
    t0 = ADD(x3, t0);                                                                                            \
    t2 = MUL(B3EXPR, t2);                                                                                        \
    F z3 = ADD(t1, t2);                                                                                          \
    t1 = SUB(t1, t2);                                                                                            \
    y3 = MUL(B3EXPR, y3);                                                                                        \
    x3 = MUL(t4, y3);                                                                                            \
    t2 = MUL(t3, t1);                                                                                            \
    x3 = SUB(t2, x3);                                                                                            \
    y3 = MUL(y3, t0);                                                                                            \
    t1 = MUL(t1, z3);                                                                                            \
    y3 = ADD(t1, y3);                                                                                            \
    t0 = MUL(t0, t3);                                                                                            \
    z3 = MUL(z3, t4);                                                                                            \
    z3 = ADD(z3, t0);                                                                                            \
    return (T){x3, y3, z3};                                                                                      \
  }                                                                                                              \
  /* NAF double-and-add over 256 digits (group.rs:639-667, fp.rs:653-662); k = 4 LE words, < 2^255 */           \
  static T T##_mul(T p, const uint64_t k[4]) {                                                                   \
    uint64_t xh[4], x3[4], np[4], nm[4];                                                                         \
    for (int i = 0; i < 4; i++) xh[i] = (k[i] >> 1) | (i < 3 ? k[i + 1] << 63 : 0);                              \
    u128 c = 0;                                                                                                  \
    for (int i = 0; i < 4; i++) {                                                                                \
      c += (u128)k[i] + xh[i];                                                                                   \
      x3[i] = (uint64_t)c;                                                                                       \
      c >>= 64;                                                                                                  \
    }                                                                                                            \
    for (int i = 0; i < 4; i++) {                                                                                \
      uint64_t cc = xh[i] ^ x3[i];                                                                               \
      np[i] = x3[i] & cc;                                                                                        \
      nm[i] = xh[i] & cc;                                                                                        \
    }                                                                                                            \
    T res = T##_zero();                                                                                          \
    T neg = (T){p.x, NEG(p.y), p.z};                                                                             \
    for (int i = 255; i >= 0; i--) {                                                                             \
      res = T##_double(res);                                                                                     \
      if ((np[i >> 6] >> (i & 63)) & 1)                                                                          \
        res = T##_add(res, p);                                                                                   \
      else if ((nm[i >> 6] >> (i & 63)) & 1)                                                                     \
        res = T##_add(res, neg);                                                                                 \
    }                                                                                                            \
    return res;                                                                                                  \
  }

static fp g1_b3(void) { return fp_mul(FP_THREE, FP_THREE); }            /* F::from(3) * curve_constant, group.rs:358 */
static fp2 g2_b3(void) { return fp2_mul((fp2){FP_THREE, FP_ZERO}, TWIST_B); }
DEFINE_GROUP(g1p, fp, fp_add, fp_sub, fp_mul, fp_neg, fp_is_zero, g1_b3(), FP_ZERO, FP_ONE, fp_inv)
DEFINE_GROUP(g2p, fp2, fp2_add, fp2_sub, fp2_mul, fp2_neg, fp2_is_zero, g2_b3(), FP2_ZERO_, FP2_ONE_, fp2_inv)

static g1a g1_to_affine(g1p p) { /* group.rs:475-495 */
  fp zi = fp_inv(p.z);
  if (fp_is_zero(zi)) return (g1a){FP_ZERO, FP_ONE, 1};
  return (g1a){fp_mul(p.x, zi), fp_mul(p.y, zi), 0};
}
static g2a g2_to_affine(g2p p) {
  fp2 zi = fp2_inv(p.z);
  if (fp2_is_zero(zi)) return (g2a){FP2_ZERO_, FP2_ONE_, 1};
  return (g2a){fp2_mul(p.x, zi), fp2_mul(p.y, zi), 0};
}
static g1p g1_from_affine(g1a a) { return (g1p){a.x, a.y, a.inf ? FP_ZERO : FP_ONE}; }
static g2p g2_from_affine(g2a a) { return (g2p){a.x, a.y, a.inf ? FP2_ZERO_ : FP2_ONE_}; }
static g2a g2_endo(g2a q) { /* g2.rs:140-152 */
  if (q.inf) return q;
  return (g2a){fp2_mul(EPS_EXP0, fp2_conj(q.x)), fp2_mul(EPS_EXP1, fp2_conj(q.y)), 0};
}
static g2a g2a_neg(g2a q) { return (g2a){q.x, q.inf ? FP2_ONE_ : fp2_neg(q.y), q.inf}; } /* group.rs:208-216 */
static g1a g1a_neg(g1a p) { return (g1a){p.x, p.inf ? FP_ONE : fp_neg(p.y), p.inf}; }

/* ------------------------------------------------------------------------------ pairing.rs */
typedef struct { fp2 c0, c1, c2; } ell;
typedef struct { ell c[87]; } g2pre;

static ell doubling_step(g2p* r) { /* pairing.rs:798-818 */
  fp2 a = fp2_scale(fp2_mul(r->x, r->y), TWO_INV);
  fp2 b = fp2_sqr(r->y);
  fp2 c = fp2_sqr(r->z);
  fp2 d = fp2_add(fp2_add(c, c), c);
  fp2 e = fp2_mul(TWIST_B, d);
  fp2 f = fp2_add(fp2_add(e, e), e);
  fp2 g = fp2_scale(fp2_add(b, f), TWO_INV);
  fp2 h = fp2_sub(fp2_sqr(fp2_add(r->y, r->z)), fp2_add(b, c));
  fp2 i = fp2_sub(e, b);
  fp2 j = fp2_sqr(r->x);
  fp2 e_sq = fp2_sqr(e);
  r->x = fp2_mul(a, fp2_sub(b, f));
  r->y = fp2_sub(fp2_sqr(g), fp2_add(fp2_add(e_sq, e_sq), e_sq));
  r->z = fp2_mul(b, h);
  return (ell){fp2_residue_mul(i), fp2_neg(h), fp2_add(fp2_add(j, j), j)};
}
static ell addition_step(g2p* r, const g2a* base) { /* pairing.rs:756-772 */
  fp2 d = fp2_sub(r->x, fp2_mul(r->z, base->x));
  fp2 e = fp2_sub(r->y, fp2_mul(r->z, base->y));
  fp2 f = fp2_sqr(d);
  fp2 g = fp2_sqr(e);
  fp2 h = fp2_mul(d, f);
  fp2 i = fp2_mul(r->x, f);
  fp2 j = fp2_sub(fp2_add(fp2_mul(r->z, g), h), fp2_add(i, i));
  fp2 ny = fp2_sub(fp2_mul(e, fp2_sub(i, j)), fp2_mul(h, r->y));
  r->x = fp2_mul(d, j);
  r->y = ny;
  r->z = fp2_mul(r->z, h);
  return (ell){fp2_residue_mul(fp2_sub(fp2_mul(e, base->x), fp2_mul(d, base->y))), d, fp2_neg(e)};
}
static void g2_precompute(const g2a* q, g2pre* out) { /* pairing.rs:676-708 */
  g2p r = g2_from_affine(*q);
  g2a qn = g2a_neg(*q);
  int idx = 0;
  for (int i = 0; i < 64; i++) {
    out->c[idx++] = doubling_step(&r);
    if (ATE_NAF[i] == 1)
      out->c[idx++] = addition_step(&r, q);
    else if (ATE_NAF[i] == -1)
      out->c[idx++] = addition_step(&r, &qn);
  }
  g2a q1 = g2_endo(*q);
  g2a q2 = g2a_neg(g2_endo(q1));
  out->c[idx++] = addition_step(&r, &q1);
  out->c[idx++] = addition_step(&r, &q2);
}
static fp12 line_mul(fp12 f, const ell* c, const g1a* p) {
  return fp12_sparse_mul(f, c->c0, fp2_scale(c->c1, p->y), fp2_scale(c->c2, p->x));
}
static fp12 miller_loop(const g2pre* pre, const g1a* p) { /* pairing.rs:590-619 */
  fp12 f = FP12_ONE_;
  int idx = 0;
  for (int i = 0; i < 64; i++) {
    f = line_mul(fp12_sqr(f), &pre->c[idx++], p);
    if (ATE_NAF[i] != 0) f = line_mul(f, &pre->c[idx++], p);
  }
  f = line_mul(f, &pre->c[idx++], p);
  f = line_mul(f, &pre->c[idx], p);
  return f;
}
/* glued_miller_loop, pairing.rs:970-1022: ONE Fp12 squaring per digit shared by all m pairs */
static fp12 glued_miller_loop(const g2pre* pre, const g1a* p, size_t m) {
  fp12 f = FP12_ONE_;
  int idx = 0;
  for (int i = 0; i < 64; i++) {
    f = fp12_sqr(f);
    for (size_t k = 0; k < m; k++) f = line_mul(f, &pre[k].c[idx], &p[k]);
    idx++;
    if (ATE_NAF[i] != 0) {
      for (size_t k = 0; k < m; k++) f = line_mul(f, &pre[k].c[idx], &p[k]);
      idx++;
    }
  }
  for (size_t k = 0; k < m; k++) f = line_mul(f, &pre[k].c[idx], &p[k]);
  idx++;
  for (size_t k = 0; k < m; k++) f = line_mul(f, &pre[k].c[idx], &p[k]);
  return f;
}
static void fp4_square(fp2 a, fp2 b, fp2* c0, fp2* c1) { /* pairing.rs:274-289 */
  fp2 t0 = fp2_sqr(a), t1 = fp2_sqr(b);
  *c0 = fp2_add(fp2_residue_mul(t1), t0);
  *c1 = fp2_sub(fp2_sub(fp2_sqr(fp2_add(a, b)), t0), t1);
}
static fp12 cyclotomic_squared(fp12 f) { /* pairing.rs:309-346 */
  fp2 z0 = f.c0.c0, z4 = f.c0.c1, z3 = f.c0.c2, z2 = f.c1.c0, z1 = f.c1.c1, z5 = f.c1.c2, t0, t1, t2, t3;
  fp4_square(z0, z1, &t0, &t1);
  z0 = fp2_sub(t0, z0);
  z0 = fp2_add(fp2_add(z0, z0), t0);
  z1 = fp2_add(t1, z1);
  z1 = fp2_add(fp2_add(z1, z1), t1);
  fp4_square(z2, z3, &t0, &t1);
  fp4_square(z4, z5, &t2, &t3);
  z4 = fp2_sub(t0, z4);
  z4 = fp2_add(fp2_add(z4, z4), t0);
  z5 = fp2_add(t1, z5);
  z5 = fp2_add(fp2_add(z5, z5), t1);
  t0 = fp2_residue_mul(t3);
  z2 = fp2_add(t0, z2);
  z2 = fp2_add(fp2_add(z2, z2), t0);
  z3 = fp2_sub(t2, z3);
  z3 = fp2_add(fp2_add(z3, z3), t2);
  return (fp12){{z0, z4, z3}, {z2, z1, z5}};
}
static fp12 cyclotomic_exp(fp12 f, const uint64_t e[4]) { /* pairing.rs:366-378: all 256 bits */
  fp12 res = FP12_ONE_;
  for (int w = 3; w >= 0; w--)
    for (int i = 63; i >= 0; i--) {
      res = cyclotomic_squared(res);
      if ((e[w] >> i) & 1) res = fp12_mul(res, f);
    }
  return res;
}
static fp12 exp_by_neg_z(fp12 f) { /* pairing.rs:390-392 */
  uint64_t e[4] = {BLS_X, 0, 0, 0};
  return fp12_conj(cyclotomic_exp(f, e));
}
static fp12 final_exponentiation(fp12 f0) { /* pairing.rs:245-492 */
  fp12 f = fp12_mul(fp12_conj(f0), fp12_inv(f0));
  fp12 inp = fp12_mul(fp12_frob(f, 2), f);
  fp12 a = exp_by_neg_z(inp);
  fp12 b = cyclotomic_squared(a);
  fp12 c = cyclotomic_squared(b);
  fp12 d = fp12_mul(c, b);
  fp12 e = exp_by_neg_z(d);
  fp12 ff = cyclotomic_squared(e);
  fp12 g = exp_by_neg_z(ff);
  fp12 h = fp12_conj(d);
  fp12 i = fp12_conj(g);
  fp12 j = fp12_mul(i, e);
  fp12 k = fp12_mul(j, h);
  fp12 l = fp12_mul(k, b);
  fp12 m = fp12_mul(k, e);
  fp12 n = fp12_mul(inp, m);
  fp12 o = fp12_frob(l, 1);
  fp12 p = fp12_mul(o, n);
  fp12 q = fp12_frob(k, 2);
  fp12 r = fp12_mul(q, p);
  fp12 s = fp12_conj(inp);
  fp12 t = fp12_mul(s, l);
  fp12 u = fp12_frob(t, 3);
  return fp12_mul(u, r);
}
static fp12 pairing_affine(g1a p, g2a q) { /* pairing.rs:870-893 */
  if (p.inf || q.inf) return final_exponentiation(FP12_ONE_);
  g2pre pre;
  g2_precompute(&q, &pre);
  return final_exponentiation(miller_loop(&pre, &p));
}

/* ------------------------------------------------------------------------------ hashing */
static const uint64_t KRC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
    0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
    0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
    0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
static const int KROT[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
static uint64_t rol(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
static void keccak_f(uint64_t a[5][5]) { /* a[x][y], FIPS 202 */
  for (int rd = 0; rd < 24; rd++) {
    uint64_t c[5], d[5], b[5][5];
    for (int x = 0; x < 5; x++) c[x] = a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4];
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) a[x][y] ^= d[x];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) b[y][(2 * x + 3 * y) % 5] = rol(a[x][y], KROT[x][y]);
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) a[x][y] = b[x][y] ^ (~b[(x + 1) % 5][y] & b[(x + 2) % 5][y]);
    a[0][0] ^= KRC[rd];
  }
}
static void keccak256(const uint8_t* msg, size_t n, uint8_t out[32]) { /* legacy pad 0x01, rate 136 */
  uint64_t a[5][5];
  memset(a, 0, sizeof(a));
  size_t pos = 0;
  for (size_t i = 0; i <= n; i++) {
    uint8_t byte = i < n ? msg[i] : 0x01;
    a[(pos / 8) % 5][(pos / 8) / 5] ^= (uint64_t)byte << (8 * (pos % 8));
    pos++;
    if (i < n && pos == 136) {
      keccak_f(a);
      pos = 0;
    }
  }
  a[16 % 5][16 / 5] ^= 0x8000000000000000ull;
  keccak_f(a);
  for (int i = 0; i < 32; i++) out[i] = (uint8_t)(a[(i / 8) % 5][(i / 8) / 5] >> (8 * (i % 8)));
}
/* XMDExpander<Keccak256>::expand_message for len_in_bytes = 96 (hasher.rs:201-250); dst <= 255 bytes */
static int expand_xmd96(const uint8_t* msg, size_t n, const uint8_t* dst, size_t dn, uint8_t out[96]) {
  if (dn > 255) return -1;
  size_t total = 136 + n + 3 + dn + 1;
  uint8_t* buf = (uint8_t*)malloc(total);
  if (!buf) return -1;
  memset(buf, 0, 136);
  memcpy(buf + 136, msg, n);
  buf[136 + n] = 0;
  buf[136 + n + 1] = 96;
  buf[136 + n + 2] = 0;
  memcpy(buf + 136 + n + 3, dst, dn);
  buf[total - 1] = (uint8_t)dn;
  uint8_t b0[32], bi[32], t[32 + 1 + 256];
  keccak256(buf, total, b0);
  free(buf);
  memset(bi, 0, 32);
  for (int i = 1; i <= 3; i++) {
    for (int j = 0; j < 32; j++) t[j] = b0[j] ^ bi[j];
    t[32] = (uint8_t)i;
    memcpy(t + 33, dst, dn);
    t[33 + dn] = (uint8_t)dn;
    keccak256(t, 34 + dn, bi);
    memcpy(out + 32 * (i - 1), bi, 32);
  }
  return 0;
}
static fp fp_from_be48(const uint8_t* b) { /* hasher.rs:93-111: 48-byte big-endian value mod p */
  fp hi = {{0, 0, 0, 0}}, lo;
  for (int i = 0; i < 2; i++) {
    uint64_t v = 0;
    for (int j = 0; j < 8; j++) v = (v << 8) | b[8 * (1 - i) + j];
    hi.l[i] = v;
  }
  for (int i = 0; i < 4; i++) {
    uint64_t v = 0;
    for (int j = 0; j < 8; j++) v = (v << 8) | b[16 + 8 * (3 - i) + j];
    lo.l[i] = v;
  }
  return fp_add(fp_mul(lo, R2), fp_mul(hi, R3));
}
static fp svdw_g(fp x) { return fp_add(fp_mul(fp_mul(x, x), x), FP_THREE); }
static int svdw_map(fp u, fp* xo, fp* yo) { /* svdw.rs:180-262 */
  fp tv1 = fp_mul(fp_mul(u, u), SVDW_C1);
  fp tv2 = fp_add(FP_ONE, tv1);
  tv1 = fp_sub(FP_ONE, tv1);
  fp tv3 = fp_inv(fp_mul(tv1, tv2));
  fp tv4 = fp_mul(fp_mul(fp_mul(u, tv1), tv3), SVDW_C3);
  fp x1 = fp_sub(SVDW_C2, tv4);
  int e1 = fp_is_square(svdw_g(x1));
  fp x2 = fp_add(SVDW_C2, tv4);
  int e2 = fp_is_square(svdw_g(x2)) & !e1;
  fp x3 = fp_mul(fp_mul(tv2, tv2), tv3);
  x3 = fp_add(fp_mul(fp_mul(x3, x3), SVDW_C4), SVDW_Z);
  fp x = e1 ? x1 : x3;
  x = e2 ? x2 : x;
  fp y;
  if (!fp_sqrt(svdw_g(x), &y)) return -1;
  if (fp_sgn0(u) != fp_sgn0(y)) y = fp_neg(y);
  *xo = x;
  *yo = y;
  return 0;
}
static int hash_to_g1(const uint8_t* msg, size_t n, const uint8_t* dst, size_t dn, g1p* out) { /* g1.rs:307-331 */
  uint8_t uni[96];
  if (expand_xmd96(msg, n, dst, dn, uni)) return -1;
  g1p a, b;
  if (svdw_map(fp_from_be48(uni), &a.x, &a.y)) return -1;
  if (svdw_map(fp_from_be48(uni + 48), &b.x, &b.y)) return -1;
  a.z = FP_ONE;
  b.z = FP_ONE;
  *out = g1p_add(a, b);
  return 0;
}

/* ------------------------------------------------------------------------------ init + wire codec */
static pthread_once_t once = PTHREAD_ONCE_INIT;
static void init_impl(void) {
  memset(&FP_ZERO, 0, sizeof(fp));
  FP_ONE = fp_from_u64(1);
  FP_TWO = fp_from_u64(2);
  FP_THREE = fp_from_u64(3);
  FP_FOUR = fp_from_u64(4);
  FP_NINE = fp_from_u64(9);
  TWO_INV = fp_inv(FP_TWO); /* fp2.rs:18-23 */
  FP2_ZERO_ = (fp2){FP_ZERO, FP_ZERO};
  FP2_ONE_ = (fp2){FP_ONE, FP_ZERO};
  memset(&FP12_ONE_, 0, sizeof(fp12));
  FP12_ONE_.c0.c0.c0 = FP_ONE;
  fp2 xi = {FP_NINE, FP_ONE};
  TWIST_B = fp2_mul((fp2){FP_THREE, FP_ZERO}, fp2_inv(xi)); /* 3/(9+u), fp2.rs:42-55 */
  /* gamma = xi^((p-1)/6); xi^((p^i-1)/6) = prod_{j<i} gamma^(p^j), and x^(p^j) on Fp2 is conj^j */
  uint64_t e[4];
  p_minus(1, e);
  /* divide by 6: p-1 = 6k exactly */
  {
    u128 rem = 0;
    for (int i = 3; i >= 0; i--) {
      u128 cur = (rem << 64) | e[i];
      e[i] = (uint64_t)(cur / 6);
      rem = cur % 6;
    }
  }
  fp2 gamma = fp2_pow(xi, e);
  FROB12_C1[0] = FP2_ONE_;
  for (int i = 1; i < 12; i++) FROB12_C1[i] = fp2_mul(FROB12_C1[i - 1], fp2_frob(gamma, i - 1));
  for (int i = 0; i < 6; i++) {
    FROB6_C1[i] = fp2_sqr(FROB12_C1[i]);                 /* xi^((p^i-1)/3)  (fp6.rs:40-108) */
    FROB6_C2[i] = fp2_sqr(FROB6_C1[i]);                  /* xi^((2p^i-2)/3) (fp6.rs:109-179) */
  }
  EPS_EXP0 = FROB6_C1[1];                                /* xi^((p-1)/3), g2.rs:80-94 */
  EPS_EXP1 = fp2_mul(FROB12_C1[1], FROB6_C1[1]);         /* xi^((p-1)/2), g2.rs:95-109 */
  /* G2 generator, g2.rs:47-77 */
  G2X = (fp2){fp_to_mont((fp){{5106727233969649389ull, 7440829307424791261ull, 4785637993704342649ull, 1729627375292849782ull}}),
              fp_to_mont((fp){{10945020018377822914ull, 17413811393473931026ull, 8241798111626485029ull, 1841571559660931130ull}})};
  G2Y = (fp2){fp_to_mont((fp){{5541340697920699818ull, 16416156555105522555ull, 5380518976772849807ull, 1353435754470862315ull}}),
              fp_to_mont((fp){{6173549831154472795ull, 13567992399387660019ull, 17050234209342075797ull, 650358724130500725ull}})};
  /* SvdW constants for a = 0, b = 3, Z = 1 (svdw.rs:123-153) */
  SVDW_Z = FP_ONE;
  fp gz = svdw_g(SVDW_Z);
  SVDW_C1 = gz;
  SVDW_C2 = fp_neg(fp_mul(SVDW_Z, TWO_INV));
  fp c3;
  fp_sqrt(fp_mul(fp_neg(gz), fp_mul(FP_THREE, fp_sqr(SVDW_Z))), &c3);
  if (fp_sgn0(c3)) c3 = fp_neg(c3);
  SVDW_C3 = c3;
  SVDW_C4 = fp_mul(fp_mul(FP_FOUR, fp_neg(gz)), fp_inv(fp_mul(FP_THREE, fp_sqr(SVDW_Z))));
}
static void init(void) { pthread_once(&once, init_impl); }

static fp rd_fp(const uint8_t* b) {
  fp a;
  memcpy(a.l, b, 32); /* little-endian host */
  return fp_to_mont(a);
}
static void wr_fp(uint8_t* b, fp a) {
  a = fp_from_mont(a);
  memcpy(b, a.l, 32);
}
static fp2 rd_fp2(const uint8_t* b) { return (fp2){rd_fp(b), rd_fp(b + 32)}; }
static void wr_fp2(uint8_t* b, fp2 a) { wr_fp(b, a.c0), wr_fp(b + 32, a.c1); }
static fp12 rd_fp12(const uint8_t* b) {
  fp12 f;
  fp2* c = (fp2*)&f;
  for (int i = 0; i < 6; i++) c[i] = rd_fp2(b + 64 * i);
  return f;
}
static void wr_fp12(uint8_t* b, fp12 f) {
  const fp2* c = (const fp2*)&f;
  for (int i = 0; i < 6; i++) wr_fp2(b + 64 * i, c[i]);
}
static g1a rd_g1(const uint8_t* b, int inf) { return (g1a){rd_fp(b), rd_fp(b + 32), inf}; }
static g2a rd_g2(const uint8_t* b, int inf) { return (g2a){rd_fp2(b), rd_fp2(b + 64), inf}; }

/* ------------------------------------------------------------------------------ exported batch API */
typedef struct {
  int op;
  const uint8_t *a, *a_inf, *b, *b_inf, *c;
  const uint64_t* offs;
  const uint8_t* dst;
  size_t dst_len;
  uint8_t *out, *out_inf;
  size_t lo, hi;
  fp12 partial;
} job;

static void run_item(job* j, size_t i) {
  switch (j->op) {
    case 0: /* pairing */
      wr_fp12(j->out + 384 * i, pairing_affine(rd_g1(j->a + 64 * i, j->a_inf ? j->a_inf[i] : 0),
                                               rd_g2(j->b + 128 * i, j->b_inf ? j->b_inf[i] : 0)));
      break;
    case 1: { /* miller loop value (precompute + miller_loop) */
      g1a p = rd_g1(j->a + 64 * i, 0);
      g2a q = rd_g2(j->b + 128 * i, 0);
      g2pre pre;
      g2_precompute(&q, &pre);
      wr_fp12(j->out + 384 * i, miller_loop(&pre, &p));
      break;
    }
    case 2: wr_fp12(j->out + 384 * i, final_exponentiation(rd_fp12(j->a + 384 * i))); break;
    case 3: { /* g1 mul */
      uint64_t k[4];
      memcpy(k, j->b + 32 * i, 32);
      g1a r = g1_to_affine(g1p_mul(g1_from_affine(rd_g1(j->a + 64 * i, j->a_inf ? j->a_inf[i] : 0)), k));
      wr_fp(j->out + 64 * i, r.x), wr_fp(j->out + 64 * i + 32, r.y);
      if (j->out_inf) j->out_inf[i] = (uint8_t)r.inf;
      break;
    }
    case 4: { /* g2 mul */
      uint64_t k[4];
      memcpy(k, j->b + 32 * i, 32);
      g2a r = g2_to_affine(g2p_mul(g2_from_affine(rd_g2(j->a + 128 * i, j->a_inf ? j->a_inf[i] : 0)), k));
      wr_fp2(j->out + 128 * i, r.x), wr_fp2(j->out + 128 * i + 64, r.y);
      if (j->out_inf) j->out_inf[i] = (uint8_t)r.inf;
      break;
    }
    case 5: { /* hash_to_curve -> affine */
      g1p h;
      int rc = hash_to_g1(j->c + j->offs[i], (size_t)(j->offs[i + 1] - j->offs[i]), j->dst, j->dst_len, &h);
      g1a r = g1_to_affine(h);
      wr_fp(j->out + 64 * i, r.x), wr_fp(j->out + 64 * i + 32, r.y);
      if (j->out_inf) j->out_inf[i] = (uint8_t)(rc ? 2 : r.inf);
      break;
    }
    case 6: { /* verify: two full pairings, lib.rs:223-236.  a = pks, b = sigs, c = msgs */
      g1p h;
      int rc = hash_to_g1(j->c + j->offs[i], (size_t)(j->offs[i + 1] - j->offs[i]), j->dst, j->dst_len, &h);
      g2a gen = {G2X, G2Y, 0};
      fp12 lhs = pairing_affine(rd_g1(j->b + 64 * i, 0), gen);
      fp12 rhs = pairing_affine(g1_to_affine(h), rd_g2(j->a + 128 * i, 0));
      j->out[i] = (uint8_t)(!rc && fp12_eq(lhs, rhs));
      break;
    }
    case 7: { /* sign: sk * H(m), lib.rs:179-187.  a = sks, c = msgs */
      g1p h;
      uint64_t k[4];
      memcpy(k, j->a + 32 * i, 32);
      hash_to_g1(j->c + j->offs[i], (size_t)(j->offs[i + 1] - j->offs[i]), j->dst, j->dst_len, &h);
      g1a r = g1_to_affine(g1p_mul(h, k));
      wr_fp(j->out + 64 * i, r.x), wr_fp(j->out + 64 * i + 32, r.y);
      break;
    }
    case 8: { /* batch-verify Miller values: miller(sig, G2gen) * miller(-H(m), pk) accumulated per thread */
      g1p h;
      hash_to_g1(j->c + j->offs[i], (size_t)(j->offs[i + 1] - j->offs[i]), j->dst, j->dst_len, &h);
      g2a gen = {G2X, G2Y, 0}, pk = rd_g2(j->a + 128 * i, 0);
      g1a sig = rd_g1(j->b + 64 * i, 0), hm = g1a_neg(g1_to_affine(h));
      g2pre pre;
      g2_precompute(&gen, &pre);
      j->partial = fp12_mul(j->partial, miller_loop(&pre, &sig));
      g2_precompute(&pk, &pre);
      j->partial = fp12_mul(j->partial, miller_loop(&pre, &hm));
      break;
    }
    case 9: { /* Miller product partial over pairs */
      g1a p = rd_g1(j->a + 64 * i, 0);
      g2a q = rd_g2(j->b + 128 * i, 0);
      g2pre pre;
      g2_precompute(&q, &pre);
      j->partial = fp12_mul(j->partial, miller_loop(&pre, &p));
      break;
    }
  }
}
/* op 10: the batch form as the reference runs it (examples/verify_multiple_messages_same_signer.rs:40-60):
 * glued_pairing over the pairs (sig_i, G2gen), (-H(m_i), pk_i) - every G2 point precomputed (glued_pairing precomputes
 * the generator once per pair too, pairing.rs:1029-1037), then the shared-squaring loop.  A worker walks its slice in
 * groups of GLUE_GROUP signatures so the coefficient tables (16.7 KB per pair) stay in cache. */
#define GLUE_GROUP 16
static void run_glued_verify(job* j) {
  g2pre* pre = (g2pre*)malloc(sizeof(g2pre) * 2 * GLUE_GROUP);
  g1a pts[2 * GLUE_GROUP];
  g2a gen = {G2X, G2Y, 0};
  for (size_t lo = j->lo; lo < j->hi; lo += GLUE_GROUP) {
    size_t m = j->hi - lo < GLUE_GROUP ? j->hi - lo : GLUE_GROUP;
    for (size_t k = 0; k < m; k++) {
      size_t i = lo + k;
      g1p h;
      hash_to_g1(j->c + j->offs[i], (size_t)(j->offs[i + 1] - j->offs[i]), j->dst, j->dst_len, &h);
      g2a pk = rd_g2(j->a + 128 * i, 0);
      pts[2 * k] = rd_g1(j->b + 64 * i, 0);
      pts[2 * k + 1] = g1a_neg(g1_to_affine(h));
      g2_precompute(&gen, &pre[2 * k]);
      g2_precompute(&pk, &pre[2 * k + 1]);
    }
    j->partial = fp12_mul(j->partial, glued_miller_loop(pre, pts, 2 * m));
  }
  free(pre);
}
static void* worker(void* arg) {
  job* j = (job*)arg;
  if (j->op == 10) {
    run_glued_verify(j);
    return NULL;
  }
  for (size_t i = j->lo; i < j->hi; i++) run_item(j, i);
  return NULL;
}
/* static contiguous chunks, one worker thread per chunk: the rayon par_iter equivalent */
static void run_parallel(job proto, size_t n, int threads, fp12* product) {
  init();
  if (threads < 1) threads = 1;
  if ((size_t)threads > n) threads = n ? (int)n : 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  job* jobs = (job*)malloc(sizeof(job) * threads);
  for (int t = 0; t < threads; t++) {
    jobs[t] = proto;
    jobs[t].lo = n * t / threads;
    jobs[t].hi = n * (t + 1) / threads;
    jobs[t].partial = FP12_ONE_;
    if (threads == 1)
      worker(&jobs[t]);
    else
      pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  fp12 acc = FP12_ONE_;
  for (int t = 0; t < threads; t++) {
    if (threads > 1) pthread_join(th[t], NULL);
    acc = fp12_mul(acc, jobs[t].partial);
  }
  if (product) *product = acc;
  free(th);
  free(jobs);
}

void so_pairing_batch(const uint8_t* g1, const uint8_t* g1_inf, const uint8_t* g2, const uint8_t* g2_inf, size_t n,
                      uint8_t* out, int threads) {
  job j = {0};
  j.op = 0, j.a = g1, j.a_inf = g1_inf, j.b = g2, j.b_inf = g2_inf, j.out = out;
  run_parallel(j, n, threads, NULL);
}
void so_miller_loop_batch(const uint8_t* g1, const uint8_t* g2, size_t n, uint8_t* out, int threads) {
  job j = {0};
  j.op = 1, j.a = g1, j.b = g2, j.out = out;
  run_parallel(j, n, threads, NULL);
}
void so_final_exp_batch(const uint8_t* f, size_t n, uint8_t* out, int threads) {
  job j = {0};
  j.op = 2, j.a = f, j.out = out;
  run_parallel(j, n, threads, NULL);
}
void so_miller_product(const uint8_t* g1, const uint8_t* g2, size_t n, uint8_t out[384], int threads) {
  job j = {0};
  fp12 prod;
  j.op = 9, j.a = g1, j.b = g2;
  run_parallel(j, n, threads, &prod);
  wr_fp12(out, prod);
}
void so_g1_mul_batch(const uint8_t* pts, const uint8_t* inf, const uint8_t* k, size_t n, uint8_t* out, uint8_t* out_inf,
                     int threads) {
  job j = {0};
  j.op = 3, j.a = pts, j.a_inf = inf, j.b = k, j.out = out, j.out_inf = out_inf;
  run_parallel(j, n, threads, NULL);
}
void so_g2_mul_batch(const uint8_t* pts, const uint8_t* inf, const uint8_t* k, size_t n, uint8_t* out, uint8_t* out_inf,
                     int threads) {
  job j = {0};
  j.op = 4, j.a = pts, j.a_inf = inf, j.b = k, j.out = out, j.out_inf = out_inf;
  run_parallel(j, n, threads, NULL);
}
void so_hash_to_g1_batch(const uint8_t* msgs, const uint64_t* offs, size_t n, const uint8_t* dst, size_t dst_len,
                         uint8_t* out, uint8_t* out_inf, int threads) {
  job j = {0};
  j.op = 5, j.c = msgs, j.offs = offs, j.dst = dst, j.dst_len = dst_len, j.out = out, j.out_inf = out_inf;
  run_parallel(j, n, threads, NULL);
}
void so_verify_each(const uint8_t* pks, const uint8_t* msgs, const uint64_t* offs, const uint8_t* sigs, size_t n,
                    const uint8_t* dst, size_t dst_len, uint8_t* ok, int threads) {
  job j = {0};
  j.op = 6, j.a = pks, j.b = sigs, j.c = msgs, j.offs = offs, j.dst = dst, j.dst_len = dst_len, j.out = ok;
  run_parallel(j, n, threads, NULL);
}
void so_sign_batch(const uint8_t* sks, const uint8_t* msgs, const uint64_t* offs, size_t n, const uint8_t* dst,
                   size_t dst_len, uint8_t* out, int threads) {
  job j = {0};
  j.op = 7, j.a = sks, j.c = msgs, j.offs = offs, j.dst = dst, j.dst_len = dst_len, j.out = out;
  run_parallel(j, n, threads, NULL);
}
/* examples/verify_multiple_messages_same_signer.rs:40-60: one final exponentiation for the batch */
int so_verify_batch(const uint8_t* pks, const uint8_t* msgs, const uint64_t* offs, const uint8_t* sigs, size_t n,
                    const uint8_t* dst, size_t dst_len, int threads) {
  job j = {0};
  fp12 prod;
  j.op = 8, j.a = pks, j.b = sigs, j.c = msgs, j.offs = offs, j.dst = dst, j.dst_len = dst_len;
  run_parallel(j, n, threads, &prod);
  return fp12_eq(final_exponentiation(prod), FP12_ONE_);
}
/* the same verdict through glued_miller_loop (shared squarings), the form the reference's example actually runs */
int so_verify_batch_glued(const uint8_t* pks, const uint8_t* msgs, const uint64_t* offs, const uint8_t* sigs, size_t n,
                          const uint8_t* dst, size_t dst_len, int threads) {
  job j = {0};
  fp12 prod;
  j.op = 10, j.a = pks, j.b = sigs, j.c = msgs, j.offs = offs, j.dst = dst, j.dst_len = dst_len;
  run_parallel(j, n, threads, &prod);
  return fp12_eq(final_exponentiation(prod), FP12_ONE_);
}
void so_constants(uint8_t* out /* 32*5 svdw z,c1..c4 ; then 64*(6+6+12) frobenius ; 64*3 twist_b, eps0, eps1 */) {
  init();
  wr_fp(out, SVDW_Z), wr_fp(out + 32, SVDW_C1), wr_fp(out + 64, SVDW_C2), wr_fp(out + 96, SVDW_C3), wr_fp(out + 128, SVDW_C4);
  uint8_t* p = out + 160;
  for (int i = 0; i < 6; i++, p += 64) wr_fp2(p, FROB6_C1[i]);
  for (int i = 0; i < 6; i++, p += 64) wr_fp2(p, FROB6_C2[i]);
  for (int i = 0; i < 12; i++, p += 64) wr_fp2(p, FROB12_C1[i]);
  wr_fp2(p, TWIST_B), wr_fp2(p + 64, EPS_EXP0), wr_fp2(p + 128, EPS_EXP1);
}
