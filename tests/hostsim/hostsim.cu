// TEST-ONLY host simulation of the device headers (sylow_b200/csrc/*.cuh compiled for the CPU).
//
// The product library (libsylow_b200.so) never contains or calls this code.  It exists so that the
// CPU-only test tier (`pytest -m "not gpu"`, no GPU in the build container) can check the tower /
// curve / pairing / hash LOGIC of the device source against the oracle before it runs on a B200.
// The PTX carry-chain primitives (fp_mul/fp_add/fp_sub device paths) are NOT covered here; they are
// covered by the `-m gpu` parity tests.
#include "../../sylow_b200/csrc/wire.cuh"
#include "../../sylow_b200/csrc/hash.cuh"
#include <cstring>

using namespace sylow;

extern "C" {

void hs_fp_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  Fp x = fp_load(a), y = fp_load(b), r;
  switch (op) {
    case 0: r = fp_mul(x, y); break;
    case 1: r = fp_add(x, y); break;
    case 2: r = fp_sub(x, y); break;
    case 3: r = fp_inv(x); break;
    case 4: r = fp_halve(x); break;
    case 5: r = fp_mul9(x); break;
    default: r = fp_neg(x);
  }
  fp_store(out, r);
}

void hs_fp12_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  Fp12 x = fp12_load(a), y = fp12_load(b), r;
  switch (op) {
    case 0: r = fp12_mul(x, y); break;
    case 1: r = fp12_sqr(x); break;
    case 2: r = fp12_inv(x); break;
    case 3: r = fp12_frobenius(x, 1); break;
    case 4: r = fp12_frobenius(x, 2); break;
    case 5: r = fp12_frobenius(x, 3); break;
    case 6: r = cyclotomic_squared(x); break;
    default: r = fp12_sparse_mul(x, y.c0.c0, y.c0.c1, y.c0.c2);
  }
  fp12_store(out, r);
}

void hs_miller_loop(const uint8_t* g1, const uint8_t* g2, uint8_t* out) {
  Fp12 f = miller_loop(fp_load(g1), fp_load(g1 + 32), fp2_load(g2), fp2_load(g2 + 64));
  fp12_store(out, f);
}
void hs_final_exp(const uint8_t* f, uint8_t* out) { fp12_store(out, final_exponentiation(fp12_load(f))); }
void hs_pairing(const uint8_t* g1, const uint8_t* g2, uint8_t* out) {
  Fp12 f = miller_loop(fp_load(g1), fp_load(g1 + 32), fp2_load(g2), fp2_load(g2 + 64));
  fp12_store(out, final_exponentiation(f));
}

int hs_g1_mul(const uint8_t* pt, int inf, const uint8_t* k, uint8_t* out) {
  G1Aff a{fp_load(pt), fp_load(pt + 32), inf != 0};
  Fp kk = fp_load_raw(k);
  G1Aff r = proj_to_affine(proj_scalar_mul(affine_to_proj(a), kk.l));
  fp_store(out, r.x);
  fp_store(out + 32, r.y);
  return r.inf;
}
int hs_g2_mul(const uint8_t* pt, int inf, const uint8_t* k, uint8_t* out) {
  G2Aff a{fp2_load(pt), fp2_load(pt + 64), inf != 0};
  Fp kk = fp_load_raw(k);
  G2Aff r = proj_to_affine(proj_scalar_mul(affine_to_proj(a), kk.l));
  fp2_store(out, r.x);
  fp2_store(out + 64, r.y);
  return r.inf;
}
int hs_g1_add(const uint8_t* p, int pinf, const uint8_t* q, int qinf, uint8_t* out) {
  G1Aff a{fp_load(p), fp_load(p + 32), pinf != 0}, b{fp_load(q), fp_load(q + 32), qinf != 0};
  G1Aff r = proj_to_affine(proj_add(affine_to_proj(a), affine_to_proj(b)));
  fp_store(out, r.x);
  fp_store(out + 32, r.y);
  return r.inf;
}

void hs_keccak256(const uint8_t* msg, size_t n, uint8_t* out32) {
  Keccak256 k;
  keccak_init(k);
  keccak_absorb(k, msg, n);
  keccak_final(k, out32);
}
int hs_hash_to_g1(const uint8_t* msg, size_t n, const uint8_t* dst_prime, size_t dn, uint8_t* out) {
  G1Proj p;
  bool ok = hash_to_g1(msg, n, dst_prime, dn, p);
  G1Aff r = proj_to_affine(p);
  fp_store(out, r.x);
  fp_store(out + 32, r.y);
  return ok ? (r.inf ? 1 : 0) : -1;
}
void hs_hash_to_field(const uint8_t* msg, size_t n, const uint8_t* dst_prime, size_t dn, uint8_t* out64) {
  Fp u0, u1;
  hash_to_field_keccak(msg, n, dst_prime, dn, u0, u1);
  fp_store(out64, u0);
  fp_store(out64 + 32, u1);
}
}
