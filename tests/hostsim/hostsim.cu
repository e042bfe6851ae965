// TEST-ONLY host simulation of the device headers (sylow_b200/csrc/*.cuh compiled for the CPU).
//
// The product library (libsylow_b200.so) never contains or calls this code.  It exists so that the
// CPU-only test tier (`pytest -m "not gpu"`, no GPU in the build container) can check the tower /
// curve / pairing / hash LOGIC of the device source against the oracle before it runs on a B200.
// The PTX carry-chain primitives (fp_mul/fp_add/fp_sub device paths) are NOT covered here; they are
// covered by the `-m gpu` parity tests.
#include "../../sylow_b200/csrc/wire.cuh"
#include "../../sylow_b200/csrc/hash.cuh"
#include "../../sylow_b200/csrc/fr.cuh"
#include "../../sylow_b200/csrc/pairing_lanes.cuh"
#include <atomic>
#include <cstring>
#include <thread>

// Two host threads stand in for the two lanes of pairing_lanes.cuh; SY_LANE_SYNC() is this barrier.
namespace {
struct LaneBarrier {
  std::atomic<int> count{0};
  std::atomic<int> phase{0};
  void wait() {
    int ph = phase.load();
    if (count.fetch_add(1) == 1) {
      count.store(0);
      phase.store(ph + 1);
    } else {
      while (phase.load() == ph) std::this_thread::yield();
    }
  }
};
thread_local LaneBarrier* tl_barrier = nullptr;
}  // namespace
extern "C" void sylow_hostsim_lane_sync() {
  if (tl_barrier) tl_barrier->wait();
}

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
    case 8: r = fp_inv_fermat(x); break;
    case 7: {  // (x / p) + 1 from the Jacobi iteration; +4 if x^((p-1)/2) disagrees
      int j = fp_jacobi(x);
      Fp l = fp_pow(x, SY_TAB(kPm1h), 252);
      int e = fp_is_zero(l) ? 0 : fp_eq(l, fp_one()) ? 1 : -1;
      r = fp_zero();
      r.l[0] = (uint32_t)(j + 1) + (j != e ? 4u : 0u);
      fp_store_raw(out, r);
      return;
    }
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
// the two-lane Miller loop, one host thread per lane
void hs_miller_loop_lanes(const uint8_t* g1, const uint8_t* g2, uint8_t* out) {
  static PairSlot slot;
  LaneBarrier bar;
  Fp xp = fp_load(g1), yp = fp_load(g1 + 32);
  Fp2 qx = fp2_load(g2), qy = fp2_load(g2 + 64);
  auto lane = [&](int role) {
    tl_barrier = &bar;
    lanes_miller_loop(slot, role, xp, yp, qx, qy);
    tl_barrier = nullptr;
  };
  std::thread t1(lane, 1);
  lane(0);
  t1.join();
  fp12_store(out, slot.f);
}
// the two-lane final exponentiation, one host thread per lane
void hs_final_exp_lanes(const uint8_t* f, uint8_t* out) {
  static FexpHot hot;
  static FexpCold cold;
  LaneBarrier bar;
  cold.f = fp12_load(f);
  auto lane = [&](int role) {
    tl_barrier = &bar;
    lanes_final_exponentiation(hot, cold, role);
    tl_barrier = nullptr;
  };
  std::thread t1(lane, 1);
  lane(0);
  t1.join();
  fp12_store(out, cold.E);
}
// G2Affine::precompute: 87 triples, canonical, 16704 bytes
void hs_g2_precompute(const uint8_t* g2, uint8_t* out) {
  Ell c[87];
  g2_precompute(fp2_load(g2), fp2_load(g2 + 64), c);
  for (int i = 0; i < 87; i++) {
    fp2_store(out + 192 * i, c[i].c0);
    fp2_store(out + 192 * i + 64, c[i].c1);
    fp2_store(out + 192 * i + 128, c[i].c2);
  }
}
// glued loop: pair 0 = (g1[0], g2 var) fused, pairs 1..nf = (g1[1+t], fixed[t]) from precomputed tables
void hs_glued(const uint8_t* g1, const uint8_t* g1_skip, const uint8_t* g2var, const uint8_t* g2fixed, int nv, int nf,
              uint8_t* out) {
  MillerG1 p[4];
  for (int i = 0; i < nv + nf; i++) p[i] = MillerG1{fp_load(g1 + 64 * i), fp_load(g1 + 64 * i + 32), g1_skip[i] != 0};
  Fp2 qx = fp2_load(g2var), qy = fp2_load(g2var + 64);
  static Ell tabs[3][87];
  const Ell* tp[3] = {tabs[0], tabs[1], tabs[2]};
  for (int t = 0; t < nf; t++) g2_precompute(fp2_load(g2fixed + 128 * t), fp2_load(g2fixed + 128 * t + 64), tabs[t]);
  Fp12 f;
  if (nv == 1 && nf == 1) f = glued_miller_loop<1, 1>(p, &qx, &qy, tp);
  else if (nv == 1 && nf == 3) f = glued_miller_loop<1, 3>(p, &qx, &qy, tp);
  else if (nv == 0 && nf == 1) f = glued_miller_loop<0, 1>(p, &qx, &qy, tp);
  else f = glued_miller_loop<1, 0>(p, &qx, &qy, tp);
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
int hs_g1_mul_glv(const uint8_t* pt, int inf, const uint8_t* k, uint8_t* out) {
  G1Aff a{fp_load(pt), fp_load(pt + 32), inf != 0};
  Fp kk = fp_load_raw(k);
  G1Aff r = proj_to_affine(proj_scalar_mul_glv(affine_to_proj(a), kk.l));
  fp_store(out, r.x);
  fp_store(out + 32, r.y);
  return r.inf;
}
int hs_g2_mul_glv(const uint8_t* pt, int inf, const uint8_t* k, uint8_t* out) {
  G2Aff a{fp2_load(pt), fp2_load(pt + 64), inf != 0};
  Fp kk = fp_load_raw(k);
  G2Aff r = proj_to_affine(proj_scalar_mul_glv(affine_to_proj(a), kk.l));
  fp2_store(out, r.x);
  fp2_store(out + 64, r.y);
  return r.inf;
}
int hs_g2_mul_gls(const uint8_t* pt, int inf, const uint8_t* k, uint8_t* out) {
  G2Aff a{fp2_load(pt), fp2_load(pt + 64), inf != 0};
  Fp kk = fp_load_raw(k);
  G2Aff r = proj_to_affine(g2_scalar_mul_gls(affine_to_proj(a), kk.l));
  fp2_store(out, r.x);
  fp2_store(out + 64, r.y);
  return r.inf;
}
// out: four magnitudes of 12 bytes LE; return bit j = k_j negative
int hs_gls_decompose(const uint8_t* k, uint8_t* out) {
  Fp kk = fp_load_raw(k);
  uint32_t mag[4][3];
  bool neg[4];
  gls_decompose(kk.l, mag, neg);
  memcpy(out, mag, 48);
  return (neg[0] ? 1 : 0) | (neg[1] ? 2 : 0) | (neg[2] ? 4 : 0) | (neg[3] ? 8 : 0);
}
// out: |k1| (16 bytes LE), |k2| (16 bytes LE); return bit0 = k1 negative, bit1 = k2 negative
int hs_glv_decompose(const uint8_t* k, uint8_t* out) {
  Fp kk = fp_load_raw(k);
  uint32_t k1[4], k2[4];
  bool n1, n2;
  glv_decompose(kk.l, k1, n1, k2, n2);
  memcpy(out, k1, 16);
  memcpy(out + 16, k2, 16);
  return (n1 ? 1 : 0) | (n2 ? 2 : 0);
}
// Fr (fr.cuh): op 0 mul, 1 add, 2 sub, 3 inv on canonical 32-byte LE values
void hs_fr_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  uint32_t wa[8], wb[8], w[8];
  memcpy(wa, a, 32);
  memcpy(wb, b, 32);
  Fr x = fr_from_words(wa), y = fr_from_words(wb), r;
  switch (op) {
    case 0: r = fr_mul(x, y); break;
    case 1: r = fr_add(x, y); break;
    case 2: r = fr_sub(x, y); break;
    default: r = fr_inv(x); break;
  }
  fr_to_words(w, r);
  memcpy(out, w, 32);
}
void hs_lagrange(const uint64_t* ids, size_t t, size_t i, uint8_t* out) {
  uint32_t w[8];
  fr_to_words(w, fr_lagrange_at_zero(ids, t, i));
  memcpy(out, w, 32);
}
// both SvdW maps of a hash through the shared inversion; out = x0 | y0 | x1 | y1
int hs_svdw_pair(const uint8_t* u0, const uint8_t* u1, uint8_t* out) {
  Fp x0, y0, x1, y1;
  bool ok = svdw_map_pair(fp_load(u0), fp_load(u1), x0, y0, x1, y1);
  fp_store(out, x0);
  fp_store(out + 32, y0);
  fp_store(out + 64, x1);
  fp_store(out + 96, y1);
  return ok ? 1 : 0;
}
// batch-verification weight of item idx (hash.cuh) and r * P for a 64-bit r (curve.cuh): the two pieces of the
// random-weight batch verification that are not covered by the other entry points
uint64_t hs_batch_weight(const uint8_t* seed32, uint64_t idx) {
  WeightSeed ws;
  memcpy(ws.w, seed32, 32);
  return batch_weight(ws, idx);
}
void hs_g1_mul_u64(const uint8_t* g1, uint64_t k, uint8_t* out /* 64 B affine */, uint8_t* out_inf) {
  G1Aff a{fp_load(g1), fp_load(g1 + 32), false};
  G1Aff r = proj_to_affine(proj_scalar_mul_u64(affine_to_proj(a), k));
  fp_store(out, r.x);
  fp_store(out + 32, r.y);
  *out_inf = r.inf;
}
// Fp6 product; 192-byte operands
void hs_fp6_mul( const uint8_t* a, const uint8_t* b, uint8_t* out) {
  Fp6 x{fp2_load(a), fp2_load(a + 64), fp2_load(a + 128)}, y{fp2_load(b), fp2_load(b + 64), fp2_load(b + 128)};
  Fp6 r = fp6_mul(x, y);
  fp2_store(out, r.c0);
  fp2_store(out + 64, r.c1);
  fp2_store(out + 128, r.c2);
}
// fp_lin9 on RAW limb values (no Montgomery conversion): out = 9 x + y (+ t) mod p, operands <= p
void hs_lin9(const uint8_t* x, const uint8_t* y, const uint8_t* t, int with_t, uint8_t* out) {
  Fp r = with_t ? fp_lin9(fp_load_raw(x), fp_load_raw(y), fp_load_raw(t)) : fp_lin9(fp_load_raw(x), fp_load_raw(y));
  fp_store_raw(out, r);
}
int hs_g1_add(const uint8_t* p, int pinf, const uint8_t* q, int qinf, uint8_t* out) {
  G1Aff a{fp_load(p), fp_load(p + 32), pinf != 0}, b{fp_load(q), fp_load(q + 32), qinf != 0};
  G1Aff r = proj_to_affine(proj_add(affine_to_proj(a), affine_to_proj(b)));
  fp_store(out, r.x);
  fp_store(out + 32, r.y);
  return r.inf;
}

int hs_g2_in_subgroup(const uint8_t* g2) { return g2_in_subgroup(fp2_load(g2), fp2_load(g2 + 64)) ? 1 : 0; }
int hs_g2_on_curve(const uint8_t* g2) { return g2_on_curve(fp2_load(g2), fp2_load(g2 + 64)) ? 1 : 0; }

void hs_gt_pow(const uint8_t* g, const uint8_t* k, uint8_t* out) {
  Fp kk = fp_load_raw(k);
  fp12_store(out, gt_pow(fp12_load(g), kk.l));
}

void hs_expand(int hash_id, const uint8_t* msg, size_t n, const uint8_t* dst_prime, size_t dn, uint32_t len, uint8_t* out) {
  expand_message(hash_id, msg, n, dst_prime, dn, len, out);
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
