// RFC 9380 hash-to-curve for BN254 G1 as sylow does it: expand_message_xmd over Keccak-256,
// hash_to_field (count 2, L 48), two Shallue-van de Woestijne maps and one projective addition.
//
// Replaces /root/reference/src/hasher.rs:84-128 (hash_to_field), :201-250 (XMD expand_message),
// /root/reference/src/svdw.rs:180-262 (unchecked_map_to_point) and
// /root/reference/src/groups/g1.rs:307-331 (hash_to_curve).  The digest is legacy Keccak-256
// (pad 0x01), i.e. sha3::Keccak256 as used by src/lib.rs:181,225.
#pragma once
#include "curve.cuh"

namespace sylow {

SY_DEFINE_TABLE(uint64_t, kKeccakRC, 24, 0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull,
                0x8000000080008000ull, 0x000000000000808Bull, 0x0000000080000001ull, 0x8000000080008081ull,
                0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull, 0x0000000080008009ull,
                0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
                0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull,
                0x800000008000000Aull, 0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull,
                0x8000000080008008ull)

SY_HD uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

// Keccak-f[1600], lanes s[x + 5y] (FIPS 202 section 3.2)
SY_HD_NOINLINE void keccak_f1600(uint64_t* st) {
  uint64_t s[25];
#pragma unroll
  for (int i = 0; i < 25; i++) s[i] = st[i];
  for (int round = 0; round < 24; round++) {
    uint64_t c[5], d[5];
#pragma unroll
    for (int x = 0; x < 5; x++) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
#pragma unroll
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
#pragma unroll
    for (int i = 0; i < 25; i++) s[i] ^= d[i % 5];
    // rho + pi
    uint64_t b[25];
    b[0] = s[0];
    b[10] = rotl64(s[1], 1);
    b[20] = rotl64(s[2], 62);
    b[5] = rotl64(s[3], 28);
    b[15] = rotl64(s[4], 27);
    b[16] = rotl64(s[5], 36);
    b[1] = rotl64(s[6], 44);
    b[11] = rotl64(s[7], 6);
    b[21] = rotl64(s[8], 55);
    b[6] = rotl64(s[9], 20);
    b[7] = rotl64(s[10], 3);
    b[17] = rotl64(s[11], 10);
    b[2] = rotl64(s[12], 43);
    b[12] = rotl64(s[13], 25);
    b[22] = rotl64(s[14], 39);
    b[23] = rotl64(s[15], 41);
    b[8] = rotl64(s[16], 45);
    b[18] = rotl64(s[17], 15);
    b[3] = rotl64(s[18], 21);
    b[13] = rotl64(s[19], 8);
    b[14] = rotl64(s[20], 18);
    b[24] = rotl64(s[21], 2);
    b[9] = rotl64(s[22], 61);
    b[19] = rotl64(s[23], 56);
    b[4] = rotl64(s[24], 14);
#pragma unroll
    for (int y = 0; y < 25; y += 5)
#pragma unroll
      for (int x = 0; x < 5; x++) s[y + x] = b[y + x] ^ ((~b[y + (x + 1) % 5]) & b[y + (x + 2) % 5]);
    s[0] ^= SY_TAB(kKeccakRC)[round];
  }
#pragma unroll
  for (int i = 0; i < 25; i++) st[i] = s[i];
}

struct Keccak256 {
  uint64_t s[25];
  int pos;
};
SY_HD void keccak_init(Keccak256& k) {
  for (int i = 0; i < 25; i++) k.s[i] = 0;
  k.pos = 0;
}
SY_HD void keccak_absorb_byte(Keccak256& k, uint8_t b) {
  k.s[k.pos >> 3] ^= (uint64_t)b << ((k.pos & 7) * 8);
  if (++k.pos == 136) {
    keccak_f1600(k.s);
    k.pos = 0;
  }
}
SY_HD void keccak_absorb(Keccak256& k, const uint8_t* p, size_t n) {
  for (size_t i = 0; i < n; i++) keccak_absorb_byte(k, p[i]);
}
SY_HD void keccak_final(Keccak256& k, uint8_t* out32) {
  k.s[k.pos >> 3] ^= (uint64_t)0x01 << ((k.pos & 7) * 8);
  k.s[16] ^= 0x8000000000000000ull;  // byte 135
  keccak_f1600(k.s);
  for (int i = 0; i < 32; i++) out32[i] = (uint8_t)(k.s[i >> 3] >> ((i & 7) * 8));
}

// ---- SHA-256 (FIPS 180-4), for XMDExpander::<Sha256> (the hash of the reference's RFC 9380 vectors, hasher.rs:430-470)
SY_DEFINE_TABLE(uint32_t, kSha256K, 64, 0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4,
                0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
                0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152,
                0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138,
                0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70,
                0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5,
                0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa,
                0xa4506ceb, 0xbef9a3f7, 0xc67178f2)
SY_HD uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
struct Sha256 {
  static constexpr int kBlock = 64;
  uint32_t h[8];
  uint8_t buf[64];
  uint64_t total;
  int pos;
};
SY_HD_NOINLINE void sha256_compress(Sha256& s) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++)
    w[i] = ((uint32_t)s.buf[4 * i] << 24) | ((uint32_t)s.buf[4 * i + 1] << 16) | ((uint32_t)s.buf[4 * i + 2] << 8) | s.buf[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = s.h[0], b = s.h[1], c = s.h[2], d = s.h[3], e = s.h[4], f = s.h[5], g = s.h[6], hh = s.h[7];
  for (int i = 0; i < 64; i++) {
    uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = hh + S1 + ch + SY_TAB(kSha256K)[i] + w[i];
    uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  s.h[0] += a; s.h[1] += b; s.h[2] += c; s.h[3] += d; s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += hh;
}
SY_HD void hash_init(Sha256& s) {
  const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  for (int i = 0; i < 8; i++) s.h[i] = iv[i];
  s.total = 0;
  s.pos = 0;
}
SY_HD void hash_absorb_byte(Sha256& s, uint8_t b) {
  s.buf[s.pos++] = b;
  s.total++;
  if (s.pos == 64) {
    sha256_compress(s);
    s.pos = 0;
  }
}
SY_HD void hash_final(Sha256& s, uint8_t* out32) {
  uint64_t bits = s.total * 8;
  hash_absorb_byte(s, 0x80);
  while (s.pos != 56) hash_absorb_byte(s, 0);
  for (int i = 7; i >= 0; i--) hash_absorb_byte(s, (uint8_t)(bits >> (8 * i)));
  for (int i = 0; i < 8; i++) {
    out32[4 * i] = (uint8_t)(s.h[i] >> 24);
    out32[4 * i + 1] = (uint8_t)(s.h[i] >> 16);
    out32[4 * i + 2] = (uint8_t)(s.h[i] >> 8);
    out32[4 * i + 3] = (uint8_t)s.h[i];
  }
}
// the same three verbs for Keccak-256 so the expander below is generic over the digest (the `D` of XMDExpander<D>)
struct Keccak256H : Keccak256 {
  static constexpr int kBlock = 136;
};
SY_HD void hash_init(Keccak256H& k) { keccak_init(k); }
SY_HD void hash_absorb_byte(Keccak256H& k, uint8_t b) { keccak_absorb_byte(k, b); }
SY_HD void hash_final(Keccak256H& k, uint8_t* out32) { keccak_final(k, out32); }

// XMDExpander::expand_message (hasher.rs:201-250) for any output length up to 255 * 32 bytes; both digests
// have b_in_bytes = 32.  dst_prime = DST' || I2OSP(len(DST'), 1) prepared by the caller.
template <class H>
SY_HD_NOINLINE void expand_message_xmd(const uint8_t* msg, size_t msg_len, const uint8_t* dst_prime, size_t dst_prime_len,
                                       uint32_t len_in_bytes, uint8_t* out) {
  H k;
  uint8_t b0[32], bi[32];
  hash_init(k);
  for (int i = 0; i < H::kBlock; i++) hash_absorb_byte(k, 0);  // Z_pad
  for (size_t i = 0; i < msg_len; i++) hash_absorb_byte(k, msg[i]);
  hash_absorb_byte(k, (uint8_t)(len_in_bytes >> 8));
  hash_absorb_byte(k, (uint8_t)len_in_bytes);
  hash_absorb_byte(k, 0);
  for (size_t i = 0; i < dst_prime_len; i++) hash_absorb_byte(k, dst_prime[i]);
  hash_final(k, b0);
  for (int i = 0; i < 32; i++) bi[i] = 0;
  uint32_t ell = (len_in_bytes + 31) / 32;
  for (uint32_t blk = 1; blk <= ell; blk++) {
    hash_init(k);
    for (int i = 0; i < 32; i++) hash_absorb_byte(k, b0[i] ^ bi[i]);
    hash_absorb_byte(k, (uint8_t)blk);
    for (size_t i = 0; i < dst_prime_len; i++) hash_absorb_byte(k, dst_prime[i]);
    hash_final(k, bi);
    for (uint32_t i = 0; i < 32 && 32 * (blk - 1) + i < len_in_bytes; i++) out[32 * (blk - 1) + i] = bi[i];
  }
}

// ---- SHAKE128 (FIPS 202: rate 168 bytes, domain bits 1111 -> pad byte 0x1f), the XOF of XOFExpander::<Shake128>
// (hasher.rs:258-330 and its RFC 9380 vectors :345-428)
struct Shake128 {
  uint64_t s[25];
  int pos;
};
SY_HD void shake_init(Shake128& k) {
  for (int i = 0; i < 25; i++) k.s[i] = 0;
  k.pos = 0;
}
SY_HD void shake_absorb_byte(Shake128& k, uint8_t b) {
  k.s[k.pos >> 3] ^= (uint64_t)b << ((k.pos & 7) * 8);
  if (++k.pos == 168) {
    keccak_f1600(k.s);
    k.pos = 0;
  }
}
// pad, switch to squeezing and write n output bytes
SY_HD void shake_squeeze(Shake128& k, uint8_t* out, size_t n) {
  k.s[k.pos >> 3] ^= (uint64_t)0x1f << ((k.pos & 7) * 8);
  k.s[20] ^= 0x8000000000000000ull;  // byte 167
  keccak_f1600(k.s);
  int pos = 0;
  for (size_t i = 0; i < n; i++) {
    if (pos == 168) {
      keccak_f1600(k.s);
      pos = 0;
    }
    out[i] = (uint8_t)(k.s[pos >> 3] >> ((pos & 7) * 8));
    pos++;
  }
}
// XOFExpander::expand_message (hasher.rs:312-329): H(msg || I2OSP(len, 2) || DST' || I2OSP(len(DST'), 1), len)
SY_HD_NOINLINE void expand_message_xof(const uint8_t* msg, size_t msg_len, const uint8_t* dst_prime, size_t dst_prime_len,
                                       uint32_t len_in_bytes, uint8_t* out) {
  Shake128 k;
  shake_init(k);
  for (size_t i = 0; i < msg_len; i++) shake_absorb_byte(k, msg[i]);
  shake_absorb_byte(k, (uint8_t)(len_in_bytes >> 8));
  shake_absorb_byte(k, (uint8_t)len_in_bytes);
  for (size_t i = 0; i < dst_prime_len; i++) shake_absorb_byte(k, dst_prime[i]);
  shake_squeeze(k, out, len_in_bytes);
}
// the expander selected by hash_id: 0 XMD-Keccak-256 (sylow's sign/verify), 1 XMD-SHA-256, 2 XOF-SHAKE128
SY_HD void expand_message(int hash_id, const uint8_t* msg, size_t msg_len, const uint8_t* dst_prime, size_t dst_prime_len,
                          uint32_t len_in_bytes, uint8_t* out) {
  if (hash_id == 2)
    expand_message_xof(msg, msg_len, dst_prime, dst_prime_len, len_in_bytes, out);
  else if (hash_id == 1)
    expand_message_xmd<Sha256>(msg, msg_len, dst_prime, dst_prime_len, len_in_bytes, out);
  else
    expand_message_xmd<Keccak256H>(msg, msg_len, dst_prime, dst_prime_len, len_in_bytes, out);
}

// 48 big-endian bytes -> value mod p, in Montgomery form (hasher.rs:93-111)
SY_HD Fp fp_from_be48_mod(const uint8_t* b) {
  Fp hi = fp_zero(), lo;
  for (int i = 0; i < 4; i++) {  // bytes 0..15 -> 128-bit hi
    const uint8_t* q = b + 12 - 4 * i;
    hi.l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
  }
  for (int i = 0; i < 8; i++) {  // bytes 16..47 -> 256-bit lo
    const uint8_t* q = b + 44 - 4 * i;
    lo.l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
  }
  // lo * R + hi * 2^256 * R  (mod p).  lo is an arbitrary 256-bit value (>= p four times out of five), so the
  // constants go first: fp_mul's multiplicand must be < 2p, its multiplier may be any 256-bit value.
  return fp_add(fp_mul(fp_R2(), lo), fp_mul(fp_R3(), hi));
}

// hash_to_field(msg, count = 2, L = 48) (hasher.rs:84-128): expand to 96 bytes, two 48-byte big-endian values mod p.
// hash_id: 0 = XMD Keccak-256 (sylow's sign/verify), 1 = XMD SHA-256, 2 = XOF SHAKE128.
SY_HD_NOINLINE void hash_to_field_xmd(int hash_id, const uint8_t* msg, size_t msg_len, const uint8_t* dst_prime,
                                      size_t dst_prime_len, Fp& u0, Fp& u1) {
  uint8_t uni[96];
  expand_message(hash_id, msg, msg_len, dst_prime, dst_prime_len, 96, uni);
  u0 = fp_from_be48_mod(uni);
  u1 = fp_from_be48_mod(uni + 48);
}
SY_HD void hash_to_field_keccak(const uint8_t* msg, size_t msg_len, const uint8_t* dst_prime, size_t dst_prime_len, Fp& u0,
                                Fp& u1) {
  hash_to_field_xmd(0, msg, msg_len, dst_prime, dst_prime_len, u0, u1);
}

// fp.rs:625-631 (true for 0): the Legendre symbol through the binary-GCD Jacobi iteration (fp.cuh)
SY_HD bool fp_is_square(const Fp& a) { return fp_jacobi(a) >= 0; }
SY_HD bool fp_is_square_fermat(const Fp& a) {  // the reference's own x^((p-1)/2), cross-check for the tests
  Fp l = fp_pow(a, SY_TAB(kPm1h), 252);
  return fp_is_zero(l) | fp_eq(l, fp_one());
}
SY_HD uint32_t fp_sgn0(const Fp& a) { return fp_from_mont(a).l[0] & 1u; }  // fp.rs:636-644

SY_HD Fp svdw_g(const Fp& x) { return fp_add(fp_mul(fp_sqr(x), x), SY_TAB(kFpThree)[0]); }  // x^3 + 3 (a = 0)

// svdw.rs:180-262, split around its inversion so that hash_to_g1 can share ONE Fermat ladder between the two maps.
//   svdw_head: tv1 = 1 - c1 u^2, tv2 = 1 + c1 u^2 and d = tv1 tv2 (the element the reference inverts, :195)
//   svdw_tail: everything after tv3 = inv0(d).  Returns false if the final sqrt check fails (MapError::SvdWError).
SY_HD void svdw_head(const Fp& u, Fp& tv1, Fp& tv2, Fp& d) {
  const Fp one = fp_one();
  Fp t = fp_mul(fp_sqr(u), SY_TAB(kSvdwC1)[0]);
  tv2 = fp_add(one, t);
  tv1 = fp_sub(one, t);
  d = fp_mul(tv1, tv2);
}
SY_HD_NOINLINE bool svdw_tail(const Fp& u, const Fp& tv1, const Fp& tv2, const Fp& tv3, Fp& x, Fp& y) {
  Fp tv4 = fp_mul(fp_mul(fp_mul(u, tv1), tv3), SY_TAB(kSvdwC3)[0]);
  Fp x1 = fp_sub(SY_TAB(kSvdwC2)[0], tv4);
  Fp x2 = fp_add(SY_TAB(kSvdwC2)[0], tv4);
  int j1, j2;
  fp_jacobi2(svdw_g(x1), svdw_g(x2), j1, j2);  // is_square (true for 0, fp.rs:625-631) of both candidates at once
  bool e1 = j1 >= 0;
  bool e2 = (j2 >= 0) & !e1;
  Fp x3 = fp_mul(fp_sqr(tv2), tv3);
  x3 = fp_add(fp_mul(fp_sqr(x3), SY_TAB(kSvdwC4)[0]), SY_TAB(kSvdwZ)[0]);
  x = fp_select(e1, x1, x3);
  x = fp_select(e2, x2, x);
  Fp gx = svdw_g(x);
  y = fp_pow(gx, SY_TAB(kPp1q), 251);  // fp.rs:611-616
  bool ok = fp_eq(fp_sqr(y), gx);
  bool e3 = fp_sgn0(u) == fp_sgn0(y);
  y = fp_select(e3, y, fp_neg(y));
  return ok;
}
SY_HD_NOINLINE bool svdw_map_to_point(const Fp& u, Fp& x, Fp& y) {
  Fp tv1, tv2, d;
  svdw_head(u, tv1, tv2, d);
  return svdw_tail(u, tv1, tv2, fp_inv(d), x, y);
}
// Both maps of one hash with a single inversion (Montgomery's trick): inv0(d0), inv0(d1) from inv(d0' d1') where a
// zero d is replaced by 1 for the product and its inverse forced back to 0 (inv(0) = 0, fp.rs:418-424).
SY_HD_NOINLINE bool svdw_map_pair(const Fp& u0, const Fp& u1, Fp& x0, Fp& y0, Fp& x1, Fp& y1) {
  Fp a1, a2, da, b1, b2, db;
  svdw_head(u0, a1, a2, da);
  svdw_head(u1, b1, b2, db);
  bool za = fp_is_zero(da), zb = fp_is_zero(db);
  Fp ea = fp_select(za, fp_one(), da), eb = fp_select(zb, fp_one(), db);
  Fp t = fp_inv(fp_mul(ea, eb));
  Fp ia = fp_select(za, fp_zero(), fp_mul(t, eb));
  Fp ib = fp_select(zb, fp_zero(), fp_mul(t, ea));
  bool ok = svdw_tail(u0, a1, a2, ia, x0, y0);
  ok &= svdw_tail(u1, b1, b2, ib, x1, y1);
  return ok;
}

// g1.rs:307-331: map both field elements and add (projective result)
SY_HD_NOINLINE bool hash_to_g1(const uint8_t* msg, size_t msg_len, const uint8_t* dst_prime, size_t dst_prime_len,
                               G1Proj& out, int hash_id = 0) {
  Fp u0, u1;
  hash_to_field_xmd(hash_id, msg, msg_len, dst_prime, dst_prime_len, u0, u1);
  G1Proj a, b;
  bool ok = svdw_map_pair(u0, u1, a.x, a.y, b.x, b.y);
  a.z = fp_one();
  b.z = fp_one();
  out = proj_add(a, b);
  return ok;
}

// Batch-verification weights: r_i = first 8 bytes of Keccak-256(seed || LE64(i)), forced odd (never 0).  The seed is
// the caller's secret randomness, drawn after the batch is fixed.
struct WeightSeed {
  uint64_t w[4];
};
SY_HD uint64_t batch_weight(const WeightSeed& seed, uint64_t idx) {
  uint64_t st[25];
  for (int i = 0; i < 25; i++) st[i] = 0;
  for (int i = 0; i < 4; i++) st[i] = seed.w[i];
  st[4] = idx;
  st[5] = 0x01;                   // Keccak padding after the 40 message bytes ...
  st[16] = 0x8000000000000000ull; // ... and the last bit of the 136-byte rate
  keccak_f1600(st);
  return st[0] | 1ull;
}

}  // namespace sylow
