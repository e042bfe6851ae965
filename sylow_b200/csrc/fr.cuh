// Fr = Z/r, the BN254 scalar field: only what the threshold-signature aggregation needs (Lagrange
// coefficients, examples/dkg.rs:216-226, examples/threshold_signing.rs:146-155).  The reference's Fr is
// the same macro-generated type as Fp over the other modulus (src/fields/fp.rs:541-545), inv(0) = 0
// included (:418-424).  This is a cold path (t^2 multiplications per aggregation), so it is plain portable
// C++ on 32-bit limbs - no PTX - and shared by the device build and the host simulation.
#pragma once
#include "constants.cuh"

namespace sylow {

struct Fr {
  uint32_t l[8];  // Montgomery form, R = 2^256, canonical (< r)
};

SY_HD Fr fr_sub_mod_if(const uint32_t* t, uint32_t top) {
  // t (+ top * 2^256) < 2r  ->  canonical
  uint32_t d[8], borrow = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t v = (uint64_t)t[i] - SY_TAB(kFrMod)[i] - borrow;
    d[i] = (uint32_t)v;
    borrow = (uint32_t)(v >> 63);
  }
  bool use_d = top != 0 || borrow == 0;
  Fr r;
  for (int i = 0; i < 8; i++) r.l[i] = use_d ? d[i] : t[i];
  return r;
}
SY_HD Fr fr_add(const Fr& a, const Fr& b) {
  uint32_t t[8], carry = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t v = (uint64_t)a.l[i] + b.l[i] + carry;
    t[i] = (uint32_t)v;
    carry = (uint32_t)(v >> 32);
  }
  return fr_sub_mod_if(t, carry);
}
SY_HD Fr fr_sub(const Fr& a, const Fr& b) {
  uint32_t t[8], borrow = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t v = (uint64_t)a.l[i] - b.l[i] - borrow;
    t[i] = (uint32_t)v;
    borrow = (uint32_t)(v >> 63);
  }
  uint32_t carry = 0, m = borrow ? 0xFFFFFFFFu : 0u;
  Fr r;
  for (int i = 0; i < 8; i++) {
    uint64_t v = (uint64_t)t[i] + (SY_TAB(kFrMod)[i] & m) + carry;
    r.l[i] = (uint32_t)v;
    carry = (uint32_t)(v >> 32);
  }
  return r;
}
// CIOS Montgomery product a b / R mod r
SY_HD_NOINLINE Fr fr_mul(const Fr& a, const Fr& b) {
  uint32_t t[10];
  for (int i = 0; i < 10; i++) t[i] = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < 8; j++) {
      uint64_t v = (uint64_t)a.l[j] * b.l[i] + t[j] + carry;
      t[j] = (uint32_t)v;
      carry = v >> 32;
    }
    uint64_t v = (uint64_t)t[8] + carry;
    t[8] = (uint32_t)v;
    t[9] = (uint32_t)(v >> 32);
    uint32_t m = t[0] * SY_FR_INV;
    carry = ((uint64_t)m * SY_TAB(kFrMod)[0] + t[0]) >> 32;
    for (int j = 1; j < 8; j++) {
      v = (uint64_t)m * SY_TAB(kFrMod)[j] + t[j] + carry;
      t[j - 1] = (uint32_t)v;
      carry = v >> 32;
    }
    v = (uint64_t)t[8] + carry;
    t[7] = (uint32_t)v;
    t[8] = t[9] + (uint32_t)(v >> 32);
  }
  return fr_sub_mod_if(t, t[8]);
}
SY_HD Fr fr_one() {
  Fr r;
  for (int i = 0; i < 8; i++) r.l[i] = SY_TAB(kFrOneM)[i];
  return r;
}
SY_HD bool fr_is_zero(const Fr& a) {
  uint32_t o = 0;
  for (int i = 0; i < 8; i++) o |= a.l[i];
  return o == 0;
}
// any 256-bit integer (LE words) -> Montgomery form of its residue (a * R^2 / R; a R^2 < 2^256 r keeps CIOS in range)
SY_HD Fr fr_from_words(const uint32_t* w) {
  Fr a, r2;
  for (int i = 0; i < 8; i++) {
    a.l[i] = w[i];
    r2.l[i] = SY_TAB(kFrR2)[i];
  }
  return fr_mul(a, r2);
}
SY_HD Fr fr_from_u64(uint64_t x) {  // Fr::from(u64), fp.rs:686-700
  uint32_t w[8] = {(uint32_t)x, (uint32_t)(x >> 32), 0, 0, 0, 0, 0, 0};
  return fr_from_words(w);
}
// Montgomery form -> canonical integer words
SY_HD void fr_to_words(uint32_t* w, const Fr& a) {
  Fr one;
  for (int i = 0; i < 8; i++) one.l[i] = i == 0 ? 1u : 0u;
  Fr r = fr_mul(a, one);
  for (int i = 0; i < 8; i++) w[i] = r.l[i];
}
// a^(r-2); 0 -> 0 like the reference's inv (fp.rs:418-424)
SY_HD_NOINLINE Fr fr_inv(const Fr& a) {
  Fr acc = fr_one();
  for (int i = 253; i >= 0; i--) {
    acc = fr_mul(acc, acc);
    if ((SY_TAB(kFrModM2)[i >> 5] >> (i & 31)) & 1u) acc = fr_mul(acc, a);
  }
  return acc;
}

// Lagrange coefficient at 0 of participant `i` among `ids[0..t)`:
//   prod_{j != i} x_j / (x_j - x_i)   (examples/dkg.rs:216-226)
// computed as (prod x_j) * (prod (x_j - x_i))^-1: the same field element with one inversion.
SY_HD_NOINLINE Fr fr_lagrange_at_zero(const uint64_t* ids, size_t t, size_t i) {
  Fr xi = fr_from_u64(ids[i]);
  Fr num = fr_one(), den = fr_one();
  for (size_t j = 0; j < t; j++) {
    if (j == i) continue;
    Fr xj = fr_from_u64(ids[j]);
    num = fr_mul(num, xj);
    den = fr_mul(den, fr_sub(xj, xi));
  }
  return fr_mul(num, fr_inv(den));
}

}  // namespace sylow
