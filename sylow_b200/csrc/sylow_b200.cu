// libsylow_b200.so: kernels + C ABI (include/sylow_b200.h).  sm_100a only, no CPU fallback.
#include "../../include/sylow_b200.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <thread>
#include <cstring>
#include <new>
#include <vector>

#include "hash.cuh"
#include "fr.cuh"
#include "wire.cuh"
#include "kernels_pairing.cuh"

using namespace sylow;

// ------------------------------------------------------------------------------------------------
// kernels.  One item (pair / point / message / check) per thread; control flow is uniform across
// a warp except for the infinity early-outs and the message-length loops of the hash.
// ------------------------------------------------------------------------------------------------
// Scalar multiplication ladder: GLV (curve.cuh) by default; -DSY_GLV=0 restores the plain 4-bit window ladder.
#ifndef SY_GLV
#define SY_GLV 1
#endif
// G2 additionally has the 4-dimensional GLS split (psi); -DSY_GLS=0 falls back to SY_SCALAR_MUL
#ifndef SY_GLS
#define SY_GLS 1
#endif
#if SY_GLV
#define SY_SCALAR_MUL proj_scalar_mul_glv
#else
#define SY_SCALAR_MUL proj_scalar_mul
#endif
#define SY_MUL_THREADS 256
#ifndef SY_G2_THREADS
#define SY_G2_THREADS 256
#endif
#ifndef SY_G2_MINB
#define SY_G2_MINB 2
#endif
#define SY_HASH_THREADS 256
#ifndef SY_G1_MINB
#define SY_G1_MINB 2  // resident 256-thread blocks per SM for the Fp-only kernels (G1 ladder, hash-to-curve)
#endif
#define SY_SMALL_THREADS 128

#ifndef SY_VERIFY_GLUE
#define SY_VERIFY_GLUE 0  // signatures per thread in verify_batch's Miller stage: 0 = by batch size, or 1, 2, 4; SYLOW_B200_VERIFY_GLUE overrides
#endif
struct DstPrime {
  uint8_t b[256];
  uint32_t len;
  int hash_id;  // 0 XMD Keccak-256, 1 XMD SHA-256, 2 XOF SHAKE128
};

// coeffs[i] = G2Affine::precompute(g2[i]): 87 triples (c0, c1, c2), 16704 B per point, Montgomery (raw) or canonical
__global__ void __launch_bounds__(SY_GLUED_THREADS, SY_GLUED_MINB)
k_g2_precompute(const uint8_t* __restrict__ g2, size_t n, uint8_t* __restrict__ coeffs, int raw_out) {
  size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t i = i0 < n ? i0 : n - 1;
  Ell c[87];
  g2_precompute(fp2_load(g2 + i * 128), fp2_load(g2 + i * 128 + 64), c);
  if (i0 >= n) return;
  uint8_t* o = coeffs + i * (87 * 192);
  for (int j = 0; j < 87; j++) {
    if (raw_out) {
      fp2_store_raw(o + 192 * j, c[j].c0);
      fp2_store_raw(o + 192 * j + 64, c[j].c1);
      fp2_store_raw(o + 192 * j + 128, c[j].c2);
    } else {
      fp2_store(o + 192 * j, c[j].c0);
      fp2_store(o + 192 * j + 64, c[j].c1);
      fp2_store(o + 192 * j + 128, c[j].c2);
    }
  }
}

// canonical <-> Montgomery for n Fp values (line-coefficient tables)
__global__ void k_fp_convert(const uint8_t* in, size_t n, uint8_t* out, int to_mont) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (to_mont)
    fp_store_raw(out + i * 32, fp_load(in + i * 32));
  else
    fp_store(out + i * 32, fp_load_raw(in + i * 32));
}

// Item i multiplies NV fused pairs (g1v[i*sv + v], g2v[i*NV + v]) and NF pairs against the fixed tables
// (g1f[i*sf + t], table t): f_out[i] = glued_miller_loop of those NV + NF pairs, Montgomery form.
// The NF tables (87 * 192 B each, Montgomery form) are staged once per block in shared memory; every
// thread of a warp then reads the same triple (broadcast, conflict-free).
template <int NV, int NF>
__global__ void __launch_bounds__(SY_GLUED_THREADS, SY_GLUED_MINB)
k_glued(const uint8_t* __restrict__ g1v, size_t sv, const uint8_t* __restrict__ g1v_inf, const uint8_t* __restrict__ g2v,
        const uint8_t* __restrict__ g2v_inf, const uint8_t* __restrict__ g1f, size_t sf,
        const uint8_t* __restrict__ g1f_inf, const uint8_t* __restrict__ tables, size_t n,
        uint8_t* __restrict__ f_out) {
  extern __shared__ uint4 sy_smem[];
  {
    const uint4* src = reinterpret_cast<const uint4*>(tables);
    for (int w = threadIdx.x; w < NF * 87 * 12; w += blockDim.x) sy_smem[w] = src[w];
    __syncthreads();
  }
  const Ell* tabs[NF > 0 ? NF : 1];
  for (int t = 0; t < NF; t++) tabs[t] = reinterpret_cast<const Ell*>(sy_smem) + 87 * t;
  size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t i = i0 < n ? i0 : n - 1;
  MillerG1 p[NV + NF];
  Fp2 qx[NV > 0 ? NV : 1], qy[NV > 0 ? NV : 1];
  for (int v = 0; v < NV; v++) {
    const uint8_t* a = g1v + (i * sv + v) * 64;
    const uint8_t* q = g2v + (i * NV + v) * 128;
    p[v].x = fp_load(a);
    p[v].y = fp_load(a + 32);
    p[v].skip = (g1v_inf && g1v_inf[i * sv + v]) || (g2v_inf && g2v_inf[i * NV + v]);
    qx[v] = fp2_load(q);
    qy[v] = fp2_load(q + 64);
  }
  for (int t = 0; t < NF; t++) {
    const uint8_t* a = g1f + (i * sf + t) * 64;
    p[NV + t].x = fp_load(a);
    p[NV + t].y = fp_load(a + 32);
    p[NV + t].skip = g1f_inf && g1f_inf[i * sf + t];
  }
  // the accumulator's shared-memory slot sits behind the NF tables (16-byte aligned: 87 * 192 is a multiple of 16)
  Fp12* acc = SY_MILLER_SMEM ? reinterpret_cast<Fp12*>(reinterpret_cast<char*>(sy_smem) + (size_t)NF * 87 * 192 +
                                                       (size_t)threadIdx.x * SY_ACC_STRIDE)
                             : nullptr;
  Fp12 f = glued_miller_loop<NV, NF>(p, qx, qy, tabs, acc);
  if (i0 >= n) return;
  fp12_store_raw(f_out + i * 384, f);
}

// out[t] = in[t] * in[t + T] * in[t + 2T] * ...   (Montgomery form in and out)
__global__ void __launch_bounds__(SY_SMALL_THREADS)
k_fp12_product_strided(const uint8_t* in, size_t n, uint8_t* out, size_t T) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  Fp12 acc = fp12_one();
  bool first = true;
  for (size_t i = t; i < n; i += T) {
    Fp12 v = fp12_load_raw(in + i * 384);
    acc = first ? v : fp12_mul(acc, v);
    first = false;
  }
  fp12_store_raw(out + t * 384, acc);
}

// canonical <-> Montgomery for n Fp12 values
__global__ void k_fp12_convert(const uint8_t* in, size_t n, uint8_t* out, int to_mont) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (to_mont)
    fp12_store_raw(out + i * 384, fp12_load(in + i * 384));
  else
    fp12_store(out + i * 384, fp12_load_raw(in + i * 384));
}

// ok[c] = final_exp(prod_{j<k} f[c*k + j]) == 1
__global__ void __launch_bounds__(SY_FEXP_THREADS, SY_FEXP_MINB)
k_check_products(const uint8_t* f_raw, size_t k, size_t n_checks, uint8_t* ok) {
  size_t c0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t c = c0 < n_checks ? c0 : n_checks - 1;
  Fp12 acc = fp12_one();
  for (size_t j = 0; j < k; j++) {
    Fp12 v = fp12_load_raw(f_raw + (c * k + j) * 384);
    acc = j == 0 ? v : fp12_mul(acc, v);
  }
  final_exponentiation_assign(acc, SY_FEXP_SMEM ? acc_slot() : nullptr);
  bool one = fp12_eq(acc, fp12_one());
  if (c0 >= n_checks) return;
  ok[c] = one ? 1 : 0;
}

// The two halves of k_check_products as kernels of their own, around k_final_exp_lanes for small batches:
// out[c] = prod_{j<k} f[c*k + j] (Montgomery form), and ok[c] = (gt[c] == 1) on canonical values.
__global__ void k_group_products(const uint8_t* f_raw, size_t k, size_t n_checks, uint8_t* out_raw) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_checks) return;
  Fp12 acc = fp12_load_raw(f_raw + c * k * 384);
  for (size_t j = 1; j < k; j++) acc = fp12_mul(acc, fp12_load_raw(f_raw + (c * k + j) * 384));
  fp12_store_raw(out_raw + c * 384, acc);
}
__global__ void k_gt_is_one(const uint8_t* gt, size_t n, uint8_t* ok) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const uint4* p = reinterpret_cast<const uint4*>(gt + c * 384);
  uint32_t d = 0;
  for (int i = 0; i < 24; i++) {
    uint4 v = p[i];
    d |= (i == 0 ? v.x ^ 1u : v.x) | v.y | v.z | v.w;
  }
  ok[c] = d == 0 ? 1 : 0;
}

SY_HD void g1_store_proj(uint8_t* p, const G1Proj& q) {
  fp_store_raw(p, q.x);
  fp_store_raw(p + 32, q.y);
  fp_store_raw(p + 64, q.z);
}
// Projective (Montgomery form) -> affine wire format for a whole batch with Montgomery's trick: thread t owns the
// SY_AFF_K points t, t + T, t + 2T, ... (coalesced), multiplies their Z together, inverts ONCE and unwinds.  A point at
// infinity (Z = 0) enters the product as 1 and is written as (0, 1) + flag, like GroupAffine::from (group.rs:475-495).
// F = Fp: G1, 96-byte projective / 64-byte affine points; F = Fp2: G2, 192 / 128 bytes.
#define SY_AFF_K 8
SY_HD Fp f_load_raw(const uint8_t* p, const Fp*) { return fp_load_raw(p); }
SY_HD Fp2 f_load_raw(const uint8_t* p, const Fp2*) { return fp2_load_raw(p); }
SY_HD void f_store(uint8_t* p, const Fp& v) { fp_store(p, v); }
SY_HD void f_store(uint8_t* p, const Fp2& v) { fp2_store(p, v); }
SY_HD Fp f_select(bool c, const Fp& a, const Fp& b) { return fp_select(c, a, b); }
SY_HD Fp2 f_select(bool c, const Fp2& a, const Fp2& b) { return fp2_select(c, a, b); }
template <class F>
__device__ __forceinline__ void batch_affine(const uint8_t* __restrict__ proj, size_t n, int negate,
                                             uint8_t* __restrict__ out, uint8_t* __restrict__ out_inf) {
  constexpr size_t W = sizeof(F);  // bytes per coordinate in both encodings
  const F* tag = nullptr;
  F one;
  f_set_one(one);
  const size_t T = (n + SY_AFF_K - 1) / SY_AFF_K;
  size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t t = t0 < T ? t0 : T - 1;  // the inversion ladder is block-synchronised: every thread runs it
  F pre[SY_AFF_K];
  F acc = one;
  for (int j = 0; j < SY_AFF_K; j++) {
    size_t i = t + (size_t)j * T;
    F z = i < n ? f_load_raw(proj + i * 3 * W + 2 * W, tag) : one;
    z = f_select(f_is_zero(z), one, z);
    acc = f_mul(acc, z);
    pre[j] = acc;
  }
  F inv = f_inv(acc);
  for (int j = SY_AFF_K - 1; j >= 0; j--) {
    size_t i = t + (size_t)j * T;
    if (i >= n) continue;  // only the last stride can be short; its z entered the product as 1
    F z = f_load_raw(proj + i * 3 * W + 2 * W, tag);
    bool inf = f_is_zero(z);
    z = f_select(inf, one, z);
    F zi = j ? f_mul(inv, pre[j - 1]) : inv;
    inv = f_mul(inv, z);
    if (t0 >= T) continue;
    F x = f_mul(f_load_raw(proj + i * 3 * W, tag), zi), y = f_mul(f_load_raw(proj + i * 3 * W + W, tag), zi);
    if (inf) {
      f_set_zero(x);
      y = one;
    } else if (negate) {
      y = f_neg(y);
    }
    f_store(out + i * 2 * W, x);
    f_store(out + i * 2 * W + W, y);
    if (out_inf) out_inf[i] = inf;
  }
}
__global__ void __launch_bounds__(SY_MUL_THREADS, SY_G1_MINB)
k_g1_batch_affine(const uint8_t* __restrict__ proj, size_t n, int negate, uint8_t* __restrict__ out,
                  uint8_t* __restrict__ out_inf) {
  batch_affine<Fp>(proj, n, negate, out, out_inf);
}
__global__ void __launch_bounds__(SY_MUL_THREADS, 1)
k_g2_batch_affine(const uint8_t* __restrict__ proj, size_t n, uint8_t* __restrict__ out, uint8_t* __restrict__ out_inf) {
  batch_affine<Fp2>(proj, n, 0, out, out_inf);
}

__global__ void __launch_bounds__(SY_MUL_THREADS, SY_G1_MINB)
k_g1_mul(const uint8_t* __restrict__ pts, const uint8_t* __restrict__ pts_inf,
                     const uint8_t* __restrict__ scalars, size_t n, uint8_t* __restrict__ out,
                     uint8_t* __restrict__ out_inf, uint8_t* __restrict__ proj_out) {
  size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t i = i0 < n ? i0 : n - 1;  // all threads run the (block-synchronised) loops; surplus is discarded
  G1Aff a{fp_load(pts + i * 64), fp_load(pts + i * 64 + 32), pts_inf && pts_inf[i]};
  Fp k = fp_load_raw(scalars + i * 32);
  G1Proj q = SY_SCALAR_MUL(affine_to_proj(a), k.l);
  if (proj_out) {  // large batches: k_g1_batch_affine shares one inversion between eight points
    if (i0 < n) g1_store_proj(proj_out + i * 96, q);
    return;
  }
  G1Aff r = proj_to_affine(q);
  if (i0 >= n) return;
  fp_store(out + i * 64, r.x);
  fp_store(out + i * 64 + 32, r.y);
  if (out_inf) out_inf[i] = r.inf;
}

// proj_out[i] = r_i * pts[i] (projective, Montgomery form), r_i the weight of item first_index + i
__global__ void __launch_bounds__(SY_MUL_THREADS, SY_G1_MINB)
k_g1_mul_weight(const uint8_t* __restrict__ pts, const uint8_t* __restrict__ pts_inf, const __grid_constant__ WeightSeed seed,
                uint64_t first_index, size_t n, uint8_t* __restrict__ proj_out) {
  size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t i = i0 < n ? i0 : n - 1;
  G1Aff a{fp_load(pts + i * 64), fp_load(pts + i * 64 + 32), pts_inf && pts_inf[i]};
  G1Proj q = proj_scalar_mul_u64(affine_to_proj(a), batch_weight(seed, first_index + i));
  if (i0 < n) g1_store_proj(proj_out + i * 96, q);
}
// out[i] = the 64-bit weight of item first_index + i (tests)
__global__ void k_batch_weights(const __grid_constant__ WeightSeed seed, uint64_t first_index, size_t n, uint64_t* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = batch_weight(seed, first_index + i);
}

__global__ void __launch_bounds__(SY_G2_THREADS, SY_G2_MINB)
k_g2_mul(const uint8_t* __restrict__ pts, const uint8_t* __restrict__ pts_inf,
                       const uint8_t* __restrict__ scalars, size_t n, uint8_t* __restrict__ out,
                       uint8_t* __restrict__ out_inf, uint8_t* __restrict__ proj_out) {
  size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t i = i0 < n ? i0 : n - 1;
  G2Aff a{fp2_load(pts + i * 128), fp2_load(pts + i * 128 + 64), pts_inf && pts_inf[i]};
  Fp k = fp_load_raw(scalars + i * 32);
#if SY_GLS
  G2Proj q = g2_scalar_mul_gls(affine_to_proj(a), k.l);
#else
  G2Proj q = SY_SCALAR_MUL(affine_to_proj(a), k.l);
#endif
  if (proj_out) {
    if (i0 < n) {
      fp2_store_raw(proj_out + i * 192, q.x);
      fp2_store_raw(proj_out + i * 192 + 64, q.y);
      fp2_store_raw(proj_out + i * 192 + 128, q.z);
    }
    return;
  }
  G2Aff r = proj_to_affine(q);
  if (i0 >= n) return;
  fp2_store(out + i * 128, r.x);
  fp2_store(out + i * 128 + 64, r.y);
  if (out_inf) out_inf[i] = r.inf;
}

// out[i] = affine(+-hash_to_curve(msg_i)); status[i] = 1 if the SvdW sqrt check failed
__global__ void __launch_bounds__(SY_HASH_THREADS, SY_G1_MINB)
k_hash_to_g1(const uint8_t* __restrict__ msgs, const uint64_t* __restrict__ offsets, size_t n,
             const __grid_constant__ DstPrime dst, int negate, uint8_t* __restrict__ out,
             uint8_t* __restrict__ out_inf, int* __restrict__ fail_flag, uint8_t* __restrict__ proj_out) {
  size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t i = i0 < n ? i0 : n - 1;
  uint64_t o0 = offsets[i], o1 = offsets[i + 1];
  G1Proj p;
  bool ok = hash_to_g1(msgs + o0, (size_t)(o1 - o0), dst.b, dst.len, p, dst.hash_id);
  if (proj_out) {  // large batches: affine conversion (and the negation) in k_g1_batch_affine
    if (i0 < n) {
      if (!ok) atomicExch(fail_flag, 1);
      g1_store_proj(proj_out + i * 96, p);
    }
    return;
  }
  G1Aff r = proj_to_affine(p);
  if (i0 >= n) return;
  if (!ok) atomicExch(fail_flag, 1);
  if (negate && !r.inf) r.y = fp_neg(r.y);
  fp_store(out + i * 64, r.x);
  fp_store(out + i * 64 + 32, r.y);
  if (out_inf) out_inf[i] = r.inf;
}


// status[i]: 0 ok, SYLOW_B200_ERR_DECODE (a coordinate >= p), SYLOW_B200_ERR_NOT_ON_CURVE   (g1.rs:111-132)
__global__ void k_g1_validate(const uint8_t* __restrict__ g1, const uint8_t* __restrict__ inf, size_t n,
                              int8_t* __restrict__ status, int keep_errors) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp xr = fp_load_raw(g1 + i * 64), yr = fp_load_raw(g1 + i * 64 + 32);
  int8_t st = keep_errors ? status[i] : 0;
  if (st == 0 && !(inf && inf[i])) {
    if (!fp_raw_is_canonical(xr) || !fp_raw_is_canonical(yr))
      st = SYLOW_B200_ERR_DECODE;
    else if (!g1_on_curve(fp_to_mont(xr), fp_to_mont(yr)))
      st = SYLOW_B200_ERR_NOT_ON_CURVE;
  }
  status[i] = st;
}

// + SYLOW_B200_ERR_NOT_IN_SUBGROUP: (x+1)Q + psi(xQ) + psi^2(xQ) != psi^3(2xQ)   (g2.rs:279-297, :460-525)
__global__ void __launch_bounds__(SY_MUL_THREADS, 1)
k_g2_validate(const uint8_t* __restrict__ g2, const uint8_t* __restrict__ inf, size_t n, int8_t* __restrict__ status,
              int keep_errors) {
  size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t i = i0 < n ? i0 : n - 1;
  Fp2 xr = fp2_load_raw(g2 + i * 128), yr = fp2_load_raw(g2 + i * 128 + 64);
  bool canon = fp_raw_is_canonical(xr.c0) & fp_raw_is_canonical(xr.c1) & fp_raw_is_canonical(yr.c0) &
               fp_raw_is_canonical(yr.c1);
  Fp2 x{fp_to_mont(xr.c0), fp_to_mont(xr.c1)}, y{fp_to_mont(yr.c0), fp_to_mont(yr.c1)};
  bool on = g2_on_curve(x, y);
  bool sub = g2_in_subgroup(x, y);  // every thread runs the (block-synchronised) ladder
  if (i0 >= n) return;
  int8_t st = keep_errors ? status[i] : 0;
  if (st == 0 && !(inf && inf[i])) {
    if (!canon)
      st = SYLOW_B200_ERR_DECODE;
    else if (!on)
      st = SYLOW_B200_ERR_NOT_ON_CURVE;
    else if (!sub)
      st = SYLOW_B200_ERR_NOT_IN_SUBGROUP;
  }
  status[i] = st;
}


// ---- big-endian codecs -------------------------------------------------------------
// 32 big-endian bytes -> Fp limbs (little-endian words), no reduction
__device__ __forceinline__ Fp fp_from_be32(const uint8_t* b, uint32_t top_mask) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint8_t* q = b + 28 - 4 * i;
    r.l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
  }
  r.l[7] &= top_mask;
  return r;
}
__device__ __forceinline__ void fp_to_be32(uint8_t* b, const Fp& v) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint8_t* q = b + 28 - 4 * i;
    q[0] = (uint8_t)(v.l[i] >> 24);
    q[1] = (uint8_t)(v.l[i] >> 16);
    q[2] = (uint8_t)(v.l[i] >> 8);
    q[3] = (uint8_t)v.l[i];
  }
}
__device__ __forceinline__ bool fp_raw_is_u32(const Fp& a, uint32_t v) {
  uint32_t o = a.l[0] ^ v;
#pragma unroll
  for (int i = 1; i < 8; i++) o |= a.l[i];
  return o == 0;
}
// NF big-endian field elements per point (2 for G1, 4 for G2 with the imaginary parts first,
// g2.rs:325-328) -> little-endian wire form + infinity flag + decode status.
//   mode 0: sylow's codec (g1.rs:151-280, g2.rs:319-433): bit 7 of byte 0 flags infinity and must come
//           with x = 0, y = 1;   mode 1: EIP-196/197: the all-zero encoding is the point at infinity
//           (examples/reth_bn128.rs:118-126,187-194) and there is no flag bit.
template <int NF>
__global__ void k_decode_be(const uint8_t* __restrict__ be, size_t stride, size_t n, int mode, uint8_t* __restrict__ out,
                            uint8_t* __restrict__ out_inf, int8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* b = be + i * stride;
  bool flag = mode == 0 && (b[0] >> 7);
  Fp f[NF];
  bool canon = true, all_zero = true;
  for (int j = 0; j < NF; j++) {
    f[j] = fp_from_be32(b + 32 * j, (mode == 0 && j == 0) ? 0x7fffffffu : 0xffffffffu);
    canon &= fp_raw_is_canonical(f[j]);
    all_zero &= fp_raw_is_u32(f[j], 0);
  }
  // wire order: G1 (x, y); G2 (x.c0, x.c1, y.c0, y.c1) from (x.c1, x.c0, y.c1, y.c0)
  Fp w[NF];
  if (NF == 2) {
    w[0] = f[0];
    w[1] = f[1];
  } else {
    w[0] = f[1];
    w[1] = f[0];
    w[2] = f[3];
    w[3] = f[2];
  }
  bool x_zero = true, y_one = true;
  for (int j = 0; j < NF / 2; j++) x_zero &= fp_raw_is_u32(w[j], 0);
  for (int j = 0; j < NF / 2; j++) y_one &= fp_raw_is_u32(w[NF / 2 + j], j == 0 ? 1u : 0u);
  int8_t st = canon ? 0 : SYLOW_B200_ERR_DECODE;
  bool inf = false;
  if (mode == 0) {
    if (flag) {
      inf = true;
      if (st == 0 && !(x_zero && y_one)) st = SYLOW_B200_ERR_DECODE;
    }
  } else {
    inf = all_zero;
  }
  if (inf || st != 0) {  // GroupAffine::zero(): (0, 1, infinity)
    for (int j = 0; j < NF; j++) w[j] = fp_zero();
    w[NF / 2].l[0] = 1;
  }
  for (int j = 0; j < NF; j++) fp_store_raw(out + i * (32 * NF) + 32 * j, w[j]);
  out_inf[i] = inf ? 1 : 0;
  status[i] = st;
}

// wire form -> big-endian (to_be_bytes / to_be_bytes_scrubbed, g1.rs:136-160, g2.rs:319-333)
template <int NF>
__global__ void k_encode_be(const uint8_t* __restrict__ pts, const uint8_t* __restrict__ inf, size_t n, int scrub,
                            uint8_t* __restrict__ be) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool is_inf = inf && inf[i];
  Fp w[NF];
  for (int j = 0; j < NF; j++) w[j] = fp_load_raw(pts + i * (32 * NF) + 32 * j);
  if (is_inf) {
    for (int j = 0; j < NF; j++) w[j] = fp_zero();
    if (!scrub) w[NF / 2].l[0] = 1;
  }
  uint8_t* b = be + i * (32 * NF);
  if (NF == 2) {
    fp_to_be32(b, w[0]);
    fp_to_be32(b + 32, w[1]);
  } else {
    fp_to_be32(b, w[1]);
    fp_to_be32(b + 32, w[0]);
    fp_to_be32(b + 64, w[3]);
    fp_to_be32(b + 96, w[2]);
  }
  if (is_inf && !scrub) b[0] |= 0x80;
}

// per check: the first non-zero pair status wins and forces ok = 0
__global__ void k_fold_status(const int8_t* __restrict__ st_g1, const int8_t* __restrict__ st_g2, size_t k, size_t n_checks,
                              uint8_t* __restrict__ ok, int8_t* __restrict__ status) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_checks) return;
  int8_t st = 0;
  for (size_t j = 0; j < k && st == 0; j++) {
    st = st_g1[c * k + j];
    if (st == 0) st = st_g2[c * k + j];
  }
  if (st != 0) ok[c] = 0;
  status[c] = st;
}


// out[i] = gt[i]^scalars[i]   (`Gt * Fr`, gt.rs:188-215)
__global__ void __launch_bounds__(SY_FEXP_THREADS, SY_FEXP_MINB)
k_gt_pow(const uint8_t* __restrict__ gt, const uint8_t* __restrict__ scalars, size_t n, uint8_t* __restrict__ out) {
  size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t i = i0 < n ? i0 : n - 1;
  Fp12 g = fp12_load(gt + i * 384);
  Fp k = fp_load_raw(scalars + i * 32);
  Fp12 r = gt_pow(g, k.l);
  if (i0 >= n) return;
  fp12_store(out + i * 384, r);
}


// out[i] = expand_message_xmd(msg_i, DST, len_in_bytes)   (Expander::expand_message, hasher.rs:70,201-250)
__global__ void k_expand_message(const uint8_t* __restrict__ msgs, const uint64_t* __restrict__ offsets, size_t n,
                                 const __grid_constant__ DstPrime dst, uint32_t len_in_bytes, uint8_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t o0 = offsets[i], o1 = offsets[i + 1];
  expand_message(dst.hash_id, msgs + o0, (size_t)(o1 - o0), dst.b, dst.len, len_in_bytes, out + i * len_in_bytes);
}
// out[i] = hash_to_field(msg_i, 2, 48): two canonical Fp (Expander::hash_to_field, hasher.rs:84-128)
__global__ void k_hash_to_field(const uint8_t* __restrict__ msgs, const uint64_t* __restrict__ offsets, size_t n,
                                const __grid_constant__ DstPrime dst, uint8_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t o0 = offsets[i], o1 = offsets[i + 1];
  Fp u0, u1;
  hash_to_field_xmd(dst.hash_id, msgs + o0, (size_t)(o1 - o0), dst.b, dst.len, u0, u1);
  fp_store(out + i * 64, u0);
  fp_store(out + i * 64 + 32, u1);
}
// DST longer than 255 bytes: DST' = H("H2C-OVERSIZE-DST-" || DST)   (XMDExpander::new, hasher.rs:157-172)
__global__ void k_oversize_dst(const uint8_t* __restrict__ dst, size_t dst_len, int hash_id, uint8_t* __restrict__ out32) {
  if (blockIdx.x || threadIdx.x) return;
  const char prefix[] = "H2C-OVERSIZE-DST-";
  if (hash_id == 2) {  // XOFExpander::new (hasher.rs:274-281): ceil(2k / 8) = 32 output bytes for k = 128
    Shake128 h;
    shake_init(h);
    for (int i = 0; i < 17; i++) shake_absorb_byte(h, (uint8_t)prefix[i]);
    for (size_t i = 0; i < dst_len; i++) shake_absorb_byte(h, dst[i]);
    shake_squeeze(h, out32, 32);
  } else if (hash_id == 1) {
    Sha256 h;
    hash_init(h);
    for (int i = 0; i < 17; i++) hash_absorb_byte(h, (uint8_t)prefix[i]);
    for (size_t i = 0; i < dst_len; i++) hash_absorb_byte(h, dst[i]);
    hash_final(h, out32);
  } else {
    Keccak256H h;
    hash_init(h);
    for (int i = 0; i < 17; i++) hash_absorb_byte(h, (uint8_t)prefix[i]);
    for (size_t i = 0; i < dst_len; i++) hash_absorb_byte(h, dst[i]);
    hash_final(h, out32);
  }
}


// out[t] = sum of in[t], in[t + T], ...  as projective points in Montgomery form (96 B each).  affine_in = 1: the
// input is the affine wire format (64 B) with an optional infinity flag array.  Complete additions, so
// infinities and repeated points need no special cases.
__global__ void __launch_bounds__(SY_MUL_THREADS, SY_G1_MINB)
k_g1_sum_strided(const uint8_t* __restrict__ in, const uint8_t* __restrict__ in_inf, int affine_in, size_t n,
                 uint8_t* __restrict__ out, size_t T) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  G1Proj acc = proj_zero<Fp>();
  for (size_t i = t; i < n; i += T) {
    G1Proj v;
    if (affine_in) {
      G1Aff a{fp_load(in + i * 64), fp_load(in + i * 64 + 32), in_inf && in_inf[i]};
      v = affine_to_proj(a);
    } else {
      v = G1Proj{fp_load_raw(in + i * 96), fp_load_raw(in + i * 96 + 32), fp_load_raw(in + i * 96 + 64)};
    }
    acc = proj_add(acc, v);
  }
  fp_store_raw(out + t * 96, acc.x);
  fp_store_raw(out + t * 96 + 32, acc.y);
  fp_store_raw(out + t * 96 + 64, acc.z);
}
// one projective point (Montgomery form) -> affine wire format + infinity flag, optionally negated
__global__ void k_g1_finish_sum(const uint8_t* __restrict__ in, int negate, uint8_t* __restrict__ out,
                                uint8_t* __restrict__ out_inf) {
  if (blockIdx.x || threadIdx.x) return;
  G1Proj p{fp_load_raw(in), fp_load_raw(in + 32), fp_load_raw(in + 64)};
  G1Aff r = proj_to_affine(p);
  if (negate && !r.inf) r.y = fp_neg(r.y);
  fp_store(out, r.x);
  fp_store(out + 32, r.y);
  out_inf[0] = r.inf ? 1 : 0;
}

// ---- threshold-signature aggregation (examples/dkg.rs:190-226, threshold_signing.rs:124-155) ----
// lambda[s][i] = prod_{j != i} x_j / (x_j - x_i) in Fr for the participant ids of set s, canonical 32-byte LE.
__global__ void k_lagrange(const uint64_t* __restrict__ ids, size_t n_sets, size_t t, uint8_t* __restrict__ out) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_sets * t) return;
  size_t s = g / t, i = g % t;
  Fr lam = fr_lagrange_at_zero(ids + s * t, t, i);
  uint32_t w[8];
  fr_to_words(w, lam);
  for (int k = 0; k < 8; k++) reinterpret_cast<uint32_t*>(out + g * 32)[k] = w[k];
}
// lambda[s][i] * sig[s][i] as a projective point in Montgomery form (96 B): no per-share inversion
__global__ void __launch_bounds__(SY_MUL_THREADS, SY_G1_MINB)
k_lagrange_mul(const uint64_t* __restrict__ ids, const uint8_t* __restrict__ sigs, const uint8_t* __restrict__ sigs_inf,
               size_t n_sets, size_t t, uint8_t* __restrict__ out) {
  size_t g0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t g = g0 < n_sets * t ? g0 : n_sets * t - 1;  // all threads run the block-synchronised ladder
  size_t s = g / t, i = g % t;
  Fr lam = fr_lagrange_at_zero(ids + s * t, t, i);
  uint32_t w[8];
  fr_to_words(w, lam);
  G1Aff a{fp_load(sigs + g * 64), fp_load(sigs + g * 64 + 32), sigs_inf && sigs_inf[g]};
  G1Proj r = SY_SCALAR_MUL(affine_to_proj(a), w);
  if (g0 != g) return;
  fp_store_raw(out + g * 96, r.x);
  fp_store_raw(out + g * 96 + 32, r.y);
  fp_store_raw(out + g * 96 + 64, r.z);
}
// One warp per set: sum of the set's t projective points, then affine wire format + infinity flag.
__global__ void __launch_bounds__(128) k_g1_segment_sum(const uint8_t* __restrict__ in, size_t n_sets, size_t t,
                                                         uint8_t* __restrict__ out, uint8_t* __restrict__ out_inf) {
  __shared__ G1Proj sh[128];
  size_t s0 = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  size_t s = s0 < n_sets ? s0 : n_sets - 1;
  unsigned lane = threadIdx.x & 31;
  G1Proj acc = proj_zero<Fp>();
  for (size_t i = lane; i < t; i += 32) {
    const uint8_t* p = in + (s * t + i) * 96;
    acc = proj_add(acc, G1Proj{fp_load_raw(p), fp_load_raw(p + 32), fp_load_raw(p + 64)});
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (unsigned off = 16; off > 0; off >>= 1) {
    if (lane < off) acc = proj_add(sh[threadIdx.x], sh[threadIdx.x + off]);
    __syncthreads();
    if (lane < off) sh[threadIdx.x] = acc;
    __syncthreads();
  }
  G1Aff r = proj_to_affine(sh[threadIdx.x & ~31u]);  // every thread: the inversion ladder is block-synchronised
  if (lane == 0 && s0 < n_sets) {
    fp_store(out + s * 64, r.x);
    fp_store(out + s * 64 + 32, r.y);
    out_inf[s] = r.inf ? 1 : 0;
  }
}

SY_HD Fp fp_final_sub_copy(Fp a) {  // p -> 0, anything below p unchanged
  fp_final_sub(a.l);
  return a;
}
// ---- bucket (Pippenger) multi-scalar multiplication on G1 -----------------------------------------
// sum_i k_i P_i = sum_w 2^(c w) sum_d d * B[w][d], B[w][d] = sum of the points whose c-bit digit in window w is d.
//   k_msm_count / k_msm_offsets / k_msm_scatter : counting sort of the point indices by (window, digit)
//   k_msm_bucket_sum   : one thread per bucket adds its points (complete additions: repeated points and infinities
//                        need no special case)
//   k_msm_chunk_reduce : per window, 64 chunks of digits: local running sums S_j = sum (d - d0_j) B_d, T_j = sum B_d
//   k_msm_window_final : R_w = sum_j S_j + d0_j T_j (small double-and-add), then Horner over the windows
SY_HD uint32_t msm_digit(const uint8_t* scalar, int w, int c) {
  // c <= 16 bits starting at bit w * c of a 32-byte little-endian scalar
  int bit = w * c;
  uint64_t v = 0;
  for (int b = 0; b < 5; b++) {
    int byte = (bit >> 3) + b;
    if (byte < 32) v |= (uint64_t)scalar[byte] << (8 * b);
  }
  return (uint32_t)((v >> (bit & 7)) & ((1u << c) - 1u));
}
__global__ void k_msm_count(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ inf, size_t n, int c, int nw,
                            uint32_t* __restrict__ count) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (inf && inf[i])) return;
  for (int w = 0; w < nw; w++) {
    uint32_t d = msm_digit(scalars + i * 32, w, c);
    if (d) atomicAdd(&count[((size_t)w << c) + d], 1u);
  }
}
// Per window (one thread each): exclusive prefix sums of the bucket counts, and the work items of the bucket sums -
// a bucket of `count` points is cut into ceil(count / SY_MSM_ITEM) items so that no thread adds more than SY_MSM_ITEM
// points however skewed the digits are (the top window of 254-bit scalars has a handful of digits; equal scalars put
// everything into one bucket).  Items of window w live at [w * max_items, ...).
#define SY_MSM_ITEM 256
__global__ void k_msm_offsets(const uint32_t* __restrict__ count, int c, int nw, size_t max_items,
                              uint32_t* __restrict__ offset, uint32_t* __restrict__ cursor,
                              uint32_t* __restrict__ item_first, uint32_t* __restrict__ item_start,
                              uint32_t* __restrict__ item_len, uint32_t* __restrict__ items_in_window) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  uint32_t run = 0, items = 0;
  for (uint32_t d = 0; d < (1u << c); d++) {
    size_t k = ((size_t)w << c) + d;
    uint32_t cnt = count[k];
    offset[k] = run;
    cursor[k] = run;
    item_first[k] = items;
    for (uint32_t j = 0; j < cnt; j += SY_MSM_ITEM) {
      item_start[(size_t)w * max_items + items] = run + j;
      item_len[(size_t)w * max_items + items] = cnt - j < SY_MSM_ITEM ? cnt - j : SY_MSM_ITEM;
      items++;
    }
    run += cnt;
  }
  items_in_window[w] = items;
}
__global__ void k_msm_scatter(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ inf, size_t n, int c, int nw,
                              uint32_t* __restrict__ cursor, uint32_t* __restrict__ idx) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (inf && inf[i])) return;
  for (int w = 0; w < nw; w++) {
    uint32_t d = msm_digit(scalars + i * 32, w, c);
    if (d) {
      uint32_t pos = atomicAdd(&cursor[((size_t)w << c) + d], 1u);
      idx[(size_t)w * n + pos] = (uint32_t)i;
    }
  }
}
SY_HD G1Proj g1_load_proj(const uint8_t* p) { return G1Proj{fp_load_raw(p), fp_load_raw(p + 32), fp_load_raw(p + 64)}; }
__global__ void __launch_bounds__(SY_MUL_THREADS, SY_G1_MINB)
k_msm_item_sum(const uint8_t* __restrict__ pts, size_t n, int nw, size_t max_items,
               const uint32_t* __restrict__ item_start, const uint32_t* __restrict__ item_len,
               const uint32_t* __restrict__ items_in_window, const uint32_t* __restrict__ idx,
               uint8_t* __restrict__ item_sum) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nw * max_items) return;
  size_t w = t / max_items;
  if (t - w * max_items >= items_in_window[w]) return;
  G1Proj acc = proj_zero<Fp>();
  const uint32_t* list = idx + w * n + item_start[t];
  for (uint32_t j = 0; j < item_len[t]; j++) {
    const uint8_t* p = pts + (size_t)list[j] * 64;  // Montgomery form (converted once by k_fp_convert)
    acc = proj_add(acc, G1Proj{fp_load_raw(p), fp_load_raw(p + 32), fp_one()});
  }
  g1_store_proj(item_sum + t * 96, acc);
}
__global__ void __launch_bounds__(SY_MUL_THREADS, SY_G1_MINB)
k_msm_bucket_sum(int c, int nw, size_t max_items, const uint32_t* __restrict__ count,
                 const uint32_t* __restrict__ item_first, const uint8_t* __restrict__ item_sum,
                 uint8_t* __restrict__ buckets) {
  size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= ((size_t)nw << c)) return;
  size_t w = b >> c;
  uint32_t items = (count[b] + SY_MSM_ITEM - 1) / SY_MSM_ITEM;
  G1Proj acc = proj_zero<Fp>();
  for (uint32_t j = 0; j < items; j++) acc = proj_add(acc, g1_load_proj(item_sum + (w * max_items + item_first[b] + j) * 96));
  g1_store_proj(buckets + b * 96, acc);
}
// small multiple m * P, m < 2^16, double-and-add from the top bit
SY_HD G1Proj g1_mul_small(const G1Proj& p, uint32_t m) {
  G1Proj acc = proj_zero<Fp>();
  for (int b = 15; b >= 0; b--) {
    acc = proj_double(acc);
    if ((m >> b) & 1u) acc = proj_add(acc, p);
  }
  return acc;
}
#define SY_MSM_CHUNKS 64
__global__ void __launch_bounds__(64) k_msm_chunk_reduce(const uint8_t* __restrict__ buckets, int c, int nw,
                                                          uint8_t* __restrict__ partial) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nw * SY_MSM_CHUNKS) return;
  int w = t / SY_MSM_CHUNKS, j = t % SY_MSM_CHUNKS;
  uint32_t per = ((1u << c) + SY_MSM_CHUNKS - 1) / SY_MSM_CHUNKS;
  uint32_t d0 = (uint32_t)j * per, d1 = d0 + per < (1u << c) ? d0 + per : (1u << c);
  // digits d0 + 1 .. d1 - 1 and d1 itself belong to the next chunk's base: use the half-open range (d0, d1] mapped
  // to local weights 1 .. d1 - d0; digit 0 never holds points, the top digit 2^c - 1 is < 2^c = last d1
  G1Proj run = proj_zero<Fp>(), sum = proj_zero<Fp>();
  for (uint32_t d = d1; d > d0; d--) {
    if (d < (1u << c)) run = proj_add(run, g1_load_proj(buckets + (((size_t)w << c) + d) * 96));
    sum = proj_add(sum, run);
  }
  // sum = sum_{d in (d0, d1]} (d - d0) B_d, run = sum B_d; the window value needs d B_d = (d - d0) B_d + d0 B_d
  G1Proj r = proj_add(sum, g1_mul_small(run, d0));
  g1_store_proj(partial + (size_t)t * 96, r);
}
// R_w = sum of the 64 chunk values of window w (one thread per window), written over the window's first partial
__global__ void k_msm_window_sum(uint8_t* __restrict__ partial, int nw) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  G1Proj r = proj_zero<Fp>();
  for (int j = 0; j < SY_MSM_CHUNKS; j++) r = proj_add(r, g1_load_proj(partial + ((size_t)w * SY_MSM_CHUNKS + j) * 96));
  g1_store_proj(partial + (size_t)w * SY_MSM_CHUNKS * 96, r);
}
// Horner over the windows: ((R_{nw-1} 2^c + R_{nw-2}) 2^c + ...) - 256 doublings, the only serial part
__global__ void k_msm_window_final(const uint8_t* __restrict__ partial, int c, int nw, uint8_t* __restrict__ out_proj) {
  if (blockIdx.x || threadIdx.x) return;
  G1Proj acc = proj_zero<Fp>();
  for (int w = nw - 1; w >= 0; w--) {
    for (int b = 0; b < c; b++) acc = proj_double(acc);
    acc = proj_add(acc, g1_load_proj(partial + (size_t)w * SY_MSM_CHUNKS * 96));
  }
  g1_store_proj(out_proj, acc);
}

__global__ void k_fp_op(int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
  size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t i = i0 < n ? i0 : n - 1;
  Fp x = fp_load(a + i * 32), y = fp_load(b + i * 32), r;
  switch (op) {
    case 0: r = fp_mul(x, y); break;
    case 1: r = fp_add(x, y); break;
    case 2: r = fp_sub(x, y); break;
    case 3: r = fp_inv(x); break;
    case 4: r = fp_halve(x); break;
    case 6: r = fp_inv_fermat(x); break;  // the Fermat ladder the binary-GCD inversion replaced (cross-check)
    case 7: {  // quadratic character: (x / p) + 1 in {0, 1, 2} from the Jacobi iteration, +4 if the Fermat form disagrees
      int j = fp_jacobi(x);
      Fp l = fp_pow(x, SY_TAB(kPm1h), 252);
      int e = fp_is_zero(l) ? 0 : fp_eq(l, fp_one()) ? 1 : -1;
      r = fp_zero();
      r.l[0] = (uint32_t)(j + 1) + (j != e ? 4u : 0u);
      if (i0 < n) fp_store_raw(out + i * 32, r);
      return;
    }
    case 8:  // raw limbs in and out: 9 a + b and 9 a + b + b mod p for operands <= p (fp_lin9)
    case 9: {
      Fp xr = fp_load_raw(a + i * 32), yr = fp_load_raw(b + i * 32);
      r = op == 8 ? fp_lin9(xr, yr) : fp_lin9(xr, yr, fp_final_sub_copy(yr));
      if (i0 < n) fp_store_raw(out + i * 32, r);
      return;
    }
    default: r = fp_neg(x);
  }
  if (i0 >= n) return;
  fp_store(out + i * 32, r);
}

__global__ void __launch_bounds__(SY_SMALL_THREADS)
k_fp12_op(int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
  size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t i = i0 < n ? i0 : n - 1;
  Fp12 x = fp12_load(a + i * 384), y = fp12_load(b + i * 384), r;
  switch (op) {
    case 0: r = fp12_mul(x, y); break;
    case 1: r = fp12_sqr(x); break;
    case 2: r = fp12_inv(x); break;
    case 3: r = fp12_frobenius(x, 1); break;
    case 4: r = fp12_frobenius(x, 2); break;
    case 5: r = fp12_frobenius(x, 3); break;
    case 6: r = cyclotomic_squared(x); break;
    case 7: r = fp12_sparse_mul(x, y.c0.c0, y.c0.c1, y.c0.c2); break;
    // the tower's lower levels on their own (operands in the leading coefficients, the rest of the output is 0)
    case 8: r = Fp12{Fp6{fp2_mul(x.c0.c0, y.c0.c0), fp2_zero(), fp2_zero()}, fp6_zero()}; break;
    case 9: r = Fp12{Fp6{fp2_sqr(x.c0.c0), fp2_zero(), fp2_zero()}, fp6_zero()}; break;
    case 10: r = Fp12{fp6_mul(x.c0, y.c0), fp6_zero()}; break;
    case 11: r = Fp12{fp6_sqr(x.c0), fp6_zero()}; break;
    case 12: r = Fp12{Fp6{fp2_inv(x.c0.c0), fp2_zero(), fp2_zero()}, fp6_zero()}; break;
    case 13: r = Fp12{Fp6{fp2_mul_xi(x.c0.c0), fp2_mul_xi_add(x.c0.c0, y.c0.c0), fp2_sub_mul_xi(y.c0.c0, x.c0.c0)}, fp6_zero()}; break;
    case 14: r = Fp12{fp6_inv(x.c0), fp6_zero()}; break;
    case 15: r = Fp12{Fp6{fp2_mul_fp(x.c0.c0, y.c0.c0.c0), fp2_halve(x.c0.c0), fp2_conj(x.c0.c0)}, fp6_zero()}; break;
    default: r = fp12_one();
  }
  if (i0 >= n) return;
  fp12_store(out + i * 384, r);
}


// Register-resident Montgomery-multiplication throughput probe (the roofline denominator).
template <int CHAINS>
__global__ void k_imad_probe(int iters, const uint32_t* src, uint32_t* sink) {
  Fp x[CHAINS], y;
  for (int i = 0; i < 8; i++) y.l[i] = src[i];  // run-time multiplier: no constant folding
  for (int c = 0; c < CHAINS; c++) {
    x[c] = fp_one();
    x[c].l[0] += threadIdx.x + c;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = fp_mul(x[c], y);
  }
  uint32_t acc = 0;
  for (int c = 0; c < CHAINS; c++)
    for (int i = 0; i < 8; i++) acc ^= x[c].l[i];
  if (acc == 0x12345678u) sink[0] = acc;  // practically never true; keeps the loop alive
}

// Instruction-fetch probe: the same dependent Montgomery multiplications as k_imad_probe<1>, but UNROLL copies of
// the multiplication in the loop body (UNROLL * ~3 KB of straight-line code per iteration), to see how the
// achieved rate depends on the instruction footprint that has to stream through the instruction caches.
template <int UNROLL>
__global__ void k_ifetch_probe(int iters, const uint32_t* src, uint32_t* sink) {
  Fp x = fp_one(), y;
  for (int i = 0; i < 8; i++) y.l[i] = src[i];
  x.l[0] += threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < UNROLL; c++) x = fp_mul(x, y);
  }
  uint32_t acc = 0;
  for (int i = 0; i < 8; i++) acc ^= x.l[i];
  if (acc == 0x12345678u) sink[0] = acc;
}

// Tower-level throughput probe: `iters` dependent applications of one tower operation per thread on
// thread-private operands, launched like the pairing kernels (256 threads, one block per SM), to see at
// which level of the tower the limb-product rate falls below the back-to-back fp_mul rate.
template <int OP>
__global__ void __launch_bounds__(256, 1) k_tower_probe(int iters, const uint32_t* src, uint32_t* sink) {
  Fp12 a, b;
  Fp* pa = reinterpret_cast<Fp*>(&a);
  Fp* pb = reinterpret_cast<Fp*>(&b);
  for (int j = 0; j < 12; j++)
    for (int i = 0; i < 8; i++) {
      pa[j].l[i] = (src[i] + threadIdx.x * 977u + j) & (i == 7 ? 0x1FFFFFFFu : 0xFFFFFFFFu);
      pb[j].l[i] = (src[8 + i] + threadIdx.x * 131u + j * 7u) & (i == 7 ? 0x1FFFFFFFu : 0xFFFFFFFFu);
    }
  G2Proj r{a.c0.c0, a.c0.c1, a.c0.c2};
  for (int it = 0; it < iters; it++) {
    if (OP == 0) a.c0.c0 = fp2_mul(a.c0.c0, b.c0.c0);
    if (OP == 1) a.c0.c0 = fp2_sqr(a.c0.c0);
    if (OP == 2) a.c0 = fp6_mul(a.c0, b.c0);
    if (OP == 3) a = fp12_mul(a, b);
    if (OP == 4) a = fp12_sqr(a);
    if (OP == 5) a = fp12_sparse_mul(a, b.c0.c0, b.c0.c1, b.c0.c2);
    if (OP == 6) a = cyclotomic_squared(a);
    if (OP == 7) {
      Ell e = g2_doubling_step(r);
      a.c0.c0 = fp2_add(a.c0.c0, e.c0);
      a.c0.c1 = fp2_add(a.c0.c1, e.c1);
      a.c0.c2 = fp2_add(a.c0.c2, e.c2);
    }
    if (OP == 8) a.c0.c0 = fp2_add(fp2_sub(a.c0.c0, b.c0.c0), a.c0.c1);
    if (OP == 9) a.c0.c0.c0 = fp_add(fp_sub(a.c0.c0.c0, b.c0.c0.c0), a.c0.c0.c1);
  }
  uint32_t acc = 0;
  for (int j = 0; j < 12; j++)
    for (int i = 0; i < 8; i++) acc ^= pa[j].l[i];
  acc ^= r.x.c0.l[0] ^ r.y.c0.l[1] ^ r.z.c1.l[2];
  if (acc == 0x12345678u) sink[0] = acc;
}

// Pipe-overlap probe: one fp_mul (IMAD.WIDE pipe) and eight fp additions (ALU pipe) per iteration.
//   MODE 0: the additions are independent of the product (same basic block: the scheduler may interleave them)
//   MODE 1: strictly alternating phases (the additions consume the product, the next product consumes the sum)
//   MODE 2: MODE 1 with the odd warps skewed by one addition phase
//   MODE 3: MODE 1 with the odd warps skewed by half a multiplication
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_overlap_probe(int iters, const uint32_t* src, uint32_t* sink) {
  Fp x = fp_one(), y, u, v;
  for (int i = 0; i < 8; i++) {
    y.l[i] = src[i] & (i == 7 ? 0x1FFFFFFFu : 0xFFFFFFFFu);
    u.l[i] = (src[8 + i] + threadIdx.x) & (i == 7 ? 0x0FFFFFFFu : 0xFFFFFFFFu);
    v.l[i] = (src[16 + i] ^ threadIdx.x) & (i == 7 ? 0x0FFFFFFFu : 0xFFFFFFFFu);
  }
  x.l[0] += threadIdx.x;
  if (MODE == 2 && ((threadIdx.x >> 5) & 1)) {
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
      u = fp_add(fp_sub(u, v), y);
      v = fp_sub(fp_add(v, u), y);
    }
  }
  if (MODE == 3 && ((threadIdx.x >> 5) & 1)) {
    uint32_t t = 0;
#pragma unroll 1
    for (int k = 0; k < 40; k++) t = t * x.l[1] + u.l[k & 7];
    u.l[0] ^= t & 1u;
  }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {
      x = fp_mul(x, y);
#pragma unroll
      for (int k = 0; k < 2; k++) {
        u = fp_add(fp_sub(u, v), y);
        v = fp_sub(fp_add(v, u), y);
      }
    } else {
      x = fp_mul(u, y);
      u = fp_add(fp_sub(x, v), y);
      v = fp_sub(fp_add(v, u), y);
      u = fp_add(fp_sub(u, v), y);
      v = fp_sub(fp_add(v, u), x);
      u = fp_add(u, v);
      u.l[7] &= 0x1FFFFFFFu;
    }
  }
  uint32_t acc = 0;
  for (int i = 0; i < 8; i++) acc ^= x.l[i] ^ u.l[i] ^ v.l[i];
  if (acc == 0x12345678u) sink[0] = acc;
}

// Raw pipe probes.  MODE 0: independent mad.wide.u32 (IMAD.WIDE.U32); MODE 1: independent 32-bit
// mad.lo.u32 (IMAD); MODE 2: mad.lo.cc/madc.hi.cc carry chains of 4 pairs (IMAD.WIDE.U32.X with
// predicate carries, the exact instruction form fp_mul uses).  64 multiply-adds per loop iteration.
template <int MODE>
__global__ void k_pipe_probe(int iters, const uint32_t* src, uint32_t* sink) {
  uint32_t a = src[0] | 1u, b = src[1] | 3u;
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < (MODE == 1 ? 4 : 8); rep++) {
      if (MODE == 0) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
          uint64_t acc = ((uint64_t)r[2 * c + 1] << 32) | r[2 * c];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(r[2 * c]));
          r[2 * c] = (uint32_t)acc;
          r[2 * c + 1] = (uint32_t)(acc >> 32);
        }
      } else if (MODE == 1) {
#pragma unroll
        for (int c = 0; c < 16; c++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[c]) : "r"(a), "r"(b));
      } else {
#pragma unroll
        for (int c = 0; c < 2; c++) {
          uint32_t* q = r + 8 * c;
          asm volatile(
              "mad.lo.cc.u32 %0, %8, %9, %0;\n\t"
              "madc.hi.cc.u32 %1, %8, %9, %1;\n\t"
              "madc.lo.cc.u32 %2, %8, %9, %2;\n\t"
              "madc.hi.cc.u32 %3, %8, %9, %3;\n\t"
              "madc.lo.cc.u32 %4, %8, %9, %4;\n\t"
              "madc.hi.cc.u32 %5, %8, %9, %5;\n\t"
              "madc.lo.cc.u32 %6, %8, %9, %6;\n\t"
              "madc.hi.u32 %7, %8, %9, %7;"
              : "+r"(q[0]), "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7])
              : "r"(a), "r"(b));
        }
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) acc ^= r[i];
  if (acc == 0x12345678u) sink[0] = acc;
}

// FP64 FMA probe (16 independent chains, 64 DFMA per loop iteration): how fast the OTHER multiplier on the SM
// is.  Recorded for DESIGN.md section 7 (double-precision limb products as a possible second pipe).
__global__ void k_dfma_probe(int iters, const uint32_t* src, uint32_t* sink) {
  double a = 1.0 + (double)(src[0] & 1023) * 1e-9, b = 1e-7 + (double)(src[1] & 1023) * 1e-12;
  double r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = (double)(threadIdx.x + i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 4; rep++) {
#pragma unroll
      for (int c = 0; c < 16; c++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(r[c]) : "d"(a), "d"(b));
    }
  }
  double acc = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) acc += r[i];
  if (acc == 0.12345) sink[0] = 1;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct DevBuf {
  uint8_t* p = nullptr;
  size_t cap = 0;
};

struct sylow_b200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_in = nullptr, copy_out = nullptr;  // host<->device copies of the chunked host-pointer calls
  cudaStream_t stream2 = nullptr;                      // second compute stream: odd chunks (their tails overlap)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;    // fork/join of the chunk-interleaved device paths
  int sms = 148;
  int last_cuda = 0;
  uint64_t launches = 0;
  DevBuf in_a, in_b, in_c, in_d, flag_a, flag_b, out, scratch0, scratch1, scratch2;
  int* d_fail = nullptr;
  uint8_t* d_gen_table = nullptr;  // G2PreComputed of the G2 generator, Montgomery form (16704 B)
  DevBuf tables, sum0, sum1, proj, msm;
  DevBuf check_stage[2];  // group products / Gt values of launch_check_products' two-lane path, per compute stream
  DevBuf fexp_lanes[2];  // FexpCold records of k_final_exp_lanes, one buffer per compute stream (stream, stream2)
  unsigned glued_attr_mask = 0;
  // multi-device parent (sylow_b200_create_multi): owns one single-device context per GPU and nothing else
  std::vector<sylow_b200_ctx*> children;
};

static int fail_cuda(sylow_b200_ctx* ctx, cudaError_t e) {
  if (ctx) ctx->last_cuda = (int)e;
  return SYLOW_B200_ERR_CUDA;
}
#define CK(call)                                   \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) return fail_cuda(ctx, e__); \
  } while (0)
#define CKS(call)                  \
  do {                             \
    int s__ = (call);              \
    if (s__ != 0) return s__;      \
  } while (0)

#define ENTER(ctx)                                          \
  if (!(ctx)) return SYLOW_B200_ERR_ARG;                    \
  if (!(ctx)->children.empty()) (ctx) = (ctx)->children[0]; \
  CK(cudaSetDevice((ctx)->device));
// `_dev` entry points work on one device's memory: they need a single-device context (sylow_b200_device_ctx)
#define ENTER_DEV(ctx) \
  if ((ctx) && !(ctx)->children.empty()) return SYLOW_B200_ERR_ARG;

static int reserve(sylow_b200_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return 0;
  if (b.p) CK(cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t cap = bytes + (bytes >> 3) + 256;
  cudaError_t e = cudaMalloc(&b.p, cap);
  if (e != cudaSuccess) {
    ctx->last_cuda = (int)e;
    return e == cudaErrorMemoryAllocation ? SYLOW_B200_ERR_NOMEM : SYLOW_B200_ERR_CUDA;
  }
  b.cap = cap;
  return 0;
}

static constexpr size_t sy_gcd(size_t a, size_t b) { return b ? sy_gcd(b, a % b) : a; }
static constexpr size_t sy_lcm(size_t a, size_t b) { return a / sy_gcd(a, b) * b; }
static inline unsigned nblocks(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }
static inline cudaStream_t pick(sylow_b200_ctx* ctx, void* stream) {
  return stream ? (cudaStream_t)stream : ctx->stream;
}
#define LAUNCHED(ctx)        \
  do {                       \
    (ctx)->launches++;       \
    CK(cudaGetLastError());  \
  } while (0)

// Launch shapes that do not waste the last wave.  Both pairing kernels run one item per thread for several
// milliseconds, so a batch that is not a whole number of waves (k_miller keeps 256 threads resident per SM,
// k_final_exp 384) ends with a partly filled wave that takes as long as a full one - at 2^17 items per GPU (a 2^20
// batch sharded over 8) that is 3.46 and 2.31 waves.  The whole waves are launched as usual; a remainder of less than
// 0.6 wave is launched separately as one (k_final_exp) or two (k_miller) SMALL blocks per SM, so that every SM works
// on it at low occupancy - a lone warp per scheduler runs 1.5x (Miller) to 2x (final exponentiation) faster than one of
// two or three co-resident warps (profiles/r01_overlap_probe.md), and the remainder finishes in that much less time.
struct WaveSplit {
  size_t n_main;       // items in whole waves (launched with the kernel's normal block size)
  size_t n_tail;       // remainder
  unsigned tail_threads, tail_blocks;
};
// (Cutting the last 1 + r waves into two rounds of equal, reduced occupancy instead was measured and is not better:
// profiles/r02g_policy_sweep.jsonl, tail_split 2.)
static WaveSplit wave_split(const sylow_b200_ctx* ctx, size_t n, int threads, int blocks_per_sm) {
  WaveSplit w{n, 0, 0, 0};
  const size_t wave = (size_t)ctx->sms * threads * blocks_per_sm;
  const size_t r = n % wave;
  const char* v = getenv("SYLOW_B200_TAIL_SPLIT");  // 0 disables (measurement)
  const int mode = v ? atoi(v) : 1;
  if (!mode || r == 0 || r * 10 >= wave * 6) return w;
  const size_t slots = (size_t)ctx->sms * blocks_per_sm;
  unsigned t = (unsigned)(((r + slots - 1) / slots + 31) / 32 * 32);
  if (t > (unsigned)threads) t = threads;
  w.n_main = n - r;
  w.n_tail = r;
  w.tail_threads = t;
  w.tail_blocks = (unsigned)((r + t - 1) / t);
  return w;
}
// Two lanes per item (k_miller_lanes, k_final_exp_lanes, csrc/pairing_lanes.cuh): the same values in 0.55-0.6 of the
// time per item when the items alone cannot fill the GPU, 0.88 of the one-thread kernels' throughput when they can
// (profiles/r02f_lanes_kbench.jsonl, r02g_policy_sweep.jsonl).  Used for batches of at most one two-lane wave
// (148 x 128 items) and for the remainder of a larger batch after its whole one-thread waves when that remainder fits one
// two-lane wave: 1 .. 9 472 pairings 8.5 -> 5.0 ms, 2^17 pairings (a 2^20 batch over 8 GPUs) 46.3 -> 44.3 ms.
// SYLOW_B200_LANES: 0 never, 1 automatic (default), 2 always (measurement and the parity tests).
static int lanes_mode() {  // read on every call (the tests flip it inside one process)
  const char* v = getenv("SYLOW_B200_LANES");
  return v ? atoi(v) : 1;
}
static int launch_miller_lanes(sylow_b200_ctx* ctx, const uint8_t* g1, const uint8_t* g1_inf, const uint8_t* g2,
                               const uint8_t* g2_inf, size_t n, uint8_t* f_out, int raw_out, cudaStream_t s) {
  // spread a small batch over every SM: pairs per block = n / (2 blocks x SMs), in whole warps (16 pairs), at most 64
  const size_t slots = (size_t)ctx->sms * SY_LANES_MINB;
  size_t pairs = ((n + slots - 1) / slots + 15) / 16 * 16;
  if (pairs > SY_LANES_THREADS / 2) pairs = SY_LANES_THREADS / 2;
  const unsigned threads = (unsigned)pairs * 2;
  k_miller_lanes<<<nblocks(n, (int)pairs), threads, SY_LANES_SMEM_BYTES(threads), s>>>(g1, g1_inf, g2, g2_inf, n, f_out,
                                                                                      raw_out);
  LAUNCHED(ctx);
  return 0;
}
static int launch_miller(sylow_b200_ctx* ctx, const uint8_t* g1, const uint8_t* g1_inf, const uint8_t* g2,
                         const uint8_t* g2_inf, size_t n, uint8_t* f_out, int raw_out, cudaStream_t s) {
  const int lanes = lanes_mode();
  const size_t lane_wave = (size_t)ctx->sms * (SY_LANES_THREADS / 2) * SY_LANES_MINB;
  if (lanes == 2 || (lanes == 1 && n <= lane_wave))
    return launch_miller_lanes(ctx, g1, g1_inf, g2, g2_inf, n, f_out, raw_out, s);
  WaveSplit w = wave_split(ctx, n, SY_MILLER_THREADS, SY_MILLER_MINB);
  const size_t wave1 = (size_t)ctx->sms * SY_MILLER_THREADS * SY_MILLER_MINB;
  if (lanes == 1 && n % wave1 && n % wave1 <= lane_wave) {
    const size_t r = n % wave1, o = n - r;
    k_miller<<<nblocks(o, SY_MILLER_THREADS), SY_MILLER_THREADS, SY_MILLER_SMEM_BYTES(SY_MILLER_THREADS), s>>>(
        g1, g1_inf, g2, g2_inf, 1, o, f_out, raw_out);
    LAUNCHED(ctx);
    return launch_miller_lanes(ctx, g1 + o * 64, g1_inf ? g1_inf + o : nullptr, g2 + o * 128,
                               g2_inf ? g2_inf + o : nullptr, r, f_out + o * 384, raw_out, s);
  }
  if (w.n_main) {
    k_miller<<<nblocks(w.n_main, SY_MILLER_THREADS), SY_MILLER_THREADS, SY_MILLER_SMEM_BYTES(SY_MILLER_THREADS), s>>>(
        g1, g1_inf, g2, g2_inf, 1, w.n_main, f_out, raw_out);
    LAUNCHED(ctx);
  }
  if (w.n_tail) {
    const size_t o = w.n_main;
    k_miller<<<w.tail_blocks, w.tail_threads, SY_MILLER_SMEM_BYTES(w.tail_threads), s>>>(
        g1 + o * 64, g1_inf ? g1_inf + o : nullptr, g2 + o * 128, g2_inf ? g2_inf + o : nullptr, 1, w.n_tail,
        f_out + o * 384, raw_out);
    LAUNCHED(ctx);
  }
  return 0;
}
static int launch_final_exp_lanes(sylow_b200_ctx* ctx, const uint8_t* f, int raw_in, size_t n, uint8_t* gt_out,
                                  cudaStream_t s) {
  const size_t slots = (size_t)ctx->sms * SY_LANES_MINB;
  size_t pairs = ((n + slots - 1) / slots + 15) / 16 * 16;
  if (pairs > SY_LANES_THREADS / 2) pairs = SY_LANES_THREADS / 2;
  const unsigned threads = (unsigned)pairs * 2, blocks = nblocks(n, (int)pairs);
  DevBuf& scratch = ctx->fexp_lanes[s == ctx->stream2 ? 1 : 0];
  CKS(reserve(ctx, scratch, SY_FLANES_SCRATCH_BYTES((size_t)blocks * pairs)));
  k_final_exp_lanes<<<blocks, threads, SY_FLANES_SMEM_BYTES(threads), s>>>(f, raw_in, n, gt_out, scratch.p);
  LAUNCHED(ctx);
  return 0;
}
static int launch_final_exp(sylow_b200_ctx* ctx, const uint8_t* f, int raw_in, size_t n, uint8_t* gt_out, cudaStream_t s) {
  const int lanes = lanes_mode();
  const size_t lane_wave = (size_t)ctx->sms * (SY_LANES_THREADS / 2) * SY_LANES_MINB;
  if (lanes == 2 || (lanes == 1 && n <= lane_wave)) return launch_final_exp_lanes(ctx, f, raw_in, n, gt_out, s);
  const size_t wave1 = (size_t)ctx->sms * SY_FEXP_THREADS * SY_FEXP_MINB;
  if (lanes == 1 && n % wave1 && n % wave1 <= lane_wave) {
    const size_t r = n % wave1, o = n - r;
    k_final_exp<<<nblocks(o, SY_FEXP_THREADS), SY_FEXP_THREADS, SY_FEXP_SMEM_BYTES(SY_FEXP_THREADS), s>>>(f, raw_in, o,
                                                                                                       gt_out);
    LAUNCHED(ctx);
    return launch_final_exp_lanes(ctx, f + o * 384, raw_in, r, gt_out + o * 384, s);
  }
  const WaveSplit w = wave_split(ctx, n, SY_FEXP_THREADS, SY_FEXP_MINB);
  if (w.n_main) {
    k_final_exp<<<nblocks(w.n_main, SY_FEXP_THREADS), SY_FEXP_THREADS, SY_FEXP_SMEM_BYTES(SY_FEXP_THREADS), s>>>(
        f, raw_in, w.n_main, gt_out);
    LAUNCHED(ctx);
  }
  if (w.n_tail) {
    const size_t o = w.n_main;
    k_final_exp<<<w.tail_blocks, w.tail_threads, SY_FEXP_SMEM_BYTES(w.tail_threads), s>>>(f + o * 384, raw_in, w.n_tail,
                                                                                         gt_out + o * 384);
    LAUNCHED(ctx);
  }
  return 0;
}

// ok[c] = final_exp(prod_{j<k} f[c*k + j]) == 1 for the Montgomery-form Miller values at f.  Batches that fit one
// two-lane wave go through k_final_exp_lanes (group products, exponentiation, comparison as three launches: about half
// the latency); larger ones through the fused one-thread kernel, in blocks small enough to reach every SM.
static int launch_check_products(sylow_b200_ctx* ctx, const uint8_t* f, size_t k, size_t n_checks, uint8_t* ok,
                                 cudaStream_t s) {
  const size_t lane_wave = (size_t)ctx->sms * (SY_LANES_THREADS / 2) * SY_LANES_MINB;
  if (lanes_mode() != 0 && n_checks <= lane_wave) {
    DevBuf& stage = ctx->check_stage[s == ctx->stream2 ? 1 : 0];
    CKS(reserve(ctx, stage, n_checks * 384));
    if (k > 1) {
      k_group_products<<<nblocks(n_checks, 32), 32, 0, s>>>(f, k, n_checks, stage.p);
      LAUNCHED(ctx);
    }
    CKS(launch_final_exp_lanes(ctx, k > 1 ? stage.p : f, 1, n_checks, stage.p, s));
    k_gt_is_one<<<nblocks(n_checks, SY_SMALL_THREADS), SY_SMALL_THREADS, 0, s>>>(stage.p, n_checks, ok);
    LAUNCHED(ctx);
    return 0;
  }
  unsigned t = (unsigned)(((n_checks + ctx->sms - 1) / ctx->sms + 31) / 32 * 32);
  if (t > SY_FEXP_THREADS) t = SY_FEXP_THREADS;
  k_check_products<<<nblocks(n_checks, (int)t), t, SY_FEXP_SMEM_BYTES(t), s>>>(f, k, n_checks, ok);
  LAUNCHED(ctx);
  return 0;
}

// defined further down (needs the table constants): one glued loop of NV fused + NF table pairs per item
template <int NV, int NF>
static int launch_glued(sylow_b200_ctx* ctx, const uint8_t* g1v, size_t sv, const uint8_t* g1v_inf, const uint8_t* g2v,
                        const uint8_t* g2v_inf, const uint8_t* g1f, size_t sf, const uint8_t* g1f_inf,
                        const uint8_t* tables, size_t n, uint8_t* f_out, cudaStream_t s);

extern "C" {

int sylow_b200_create(sylow_b200_ctx** out, int device_id) {
  if (!out) return SYLOW_B200_ERR_ARG;
  *out = nullptr;
  sylow_b200_ctx* ctx = new (std::nothrow) sylow_b200_ctx();
  if (!ctx) return SYLOW_B200_ERR_NOMEM;
  ctx->device = device_id;
  cudaError_t e = cudaSetDevice(device_id);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, device_id);
  // opt-in for more than 48 KB of dynamic shared memory (the Miller loop's accumulators); per device
  if (e == cudaSuccess && SY_MILLER_SMEM)
    e = cudaFuncSetAttribute(k_miller, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)SY_MILLER_SMEM_BYTES(SY_MILLER_THREADS));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_miller_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)SY_LANES_SMEM_BYTES(SY_LANES_THREADS));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_final_exp_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)SY_FLANES_SMEM_BYTES(SY_LANES_THREADS));
  if (e == cudaSuccess && SY_FEXP_SMEM)
    e = cudaFuncSetAttribute(k_final_exp, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)SY_FEXP_SMEM_BYTES(SY_FEXP_THREADS));
  if (e == cudaSuccess && SY_FEXP_SMEM)
    e = cudaFuncSetAttribute(k_check_products, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)SY_FEXP_SMEM_BYTES(SY_FEXP_THREADS));
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_fail, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(ctx->d_fail, 0, sizeof(int));
  if (e != cudaSuccess) {
    delete ctx;
    return SYLOW_B200_ERR_CUDA;
  }
  *out = ctx;
  return 0;
}

int sylow_b200_destroy(sylow_b200_ctx* ctx) {
  if (!ctx) return SYLOW_B200_ERR_ARG;
  if (!ctx->children.empty()) {
    for (sylow_b200_ctx* c : ctx->children) sylow_b200_destroy(c);
    delete ctx;
    return 0;
  }
  cudaSetDevice(ctx->device);
  DevBuf* bufs[] = {&ctx->in_a, &ctx->in_b, &ctx->in_c, &ctx->in_d, &ctx->flag_a,
                    &ctx->flag_b, &ctx->out, &ctx->scratch0, &ctx->scratch1, &ctx->scratch2};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  if (ctx->d_fail) cudaFree(ctx->d_fail);
  if (ctx->d_gen_table) cudaFree(ctx->d_gen_table);
  if (ctx->tables.p) cudaFree(ctx->tables.p);
  if (ctx->sum0.p) cudaFree(ctx->sum0.p);
  if (ctx->sum1.p) cudaFree(ctx->sum1.p);
  if (ctx->proj.p) cudaFree(ctx->proj.p);
  if (ctx->msm.p) cudaFree(ctx->msm.p);
  for (DevBuf& b : ctx->fexp_lanes)
    if (b.p) cudaFree(b.p);
  for (DevBuf& b : ctx->check_stage)
    if (b.p) cudaFree(b.p);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
  if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  delete ctx;
  return 0;
}

int sylow_b200_create_multi(sylow_b200_ctx** out, const int* device_ids, int n_devices) {
  if (!out) return SYLOW_B200_ERR_ARG;
  *out = nullptr;
  if (!device_ids || n_devices < 1) return SYLOW_B200_ERR_ARG;
  sylow_b200_ctx* parent = new (std::nothrow) sylow_b200_ctx();
  if (!parent) return SYLOW_B200_ERR_NOMEM;
  parent->device = device_ids[0];
  for (int i = 0; i < n_devices; i++) {
    sylow_b200_ctx* c = nullptr;
    int st = sylow_b200_create(&c, device_ids[i]);
    if (st == 0) {
      try {
        parent->children.push_back(c);
      } catch (...) {
        sylow_b200_destroy(c);
        st = SYLOW_B200_ERR_NOMEM;
      }
    }
    if (st != 0) {
      for (sylow_b200_ctx* d : parent->children) sylow_b200_destroy(d);
      delete parent;
      return st;
    }
  }
  *out = parent;
  return 0;
}
int sylow_b200_device_count(const sylow_b200_ctx* ctx) {
  if (!ctx) return 0;
  return ctx->children.empty() ? 1 : (int)ctx->children.size();
}
sylow_b200_ctx* sylow_b200_device_ctx(sylow_b200_ctx* ctx, int i) {
  if (!ctx || i < 0) return nullptr;
  if (ctx->children.empty()) return i == 0 ? ctx : nullptr;
  return i < (int)ctx->children.size() ? ctx->children[i] : nullptr;
}

const char* sylow_b200_strerror(int status) {
  switch (status) {
    case SYLOW_B200_OK: return "ok";
    case SYLOW_B200_ERR_ARG: return "invalid argument";
    case SYLOW_B200_ERR_CUDA: return "CUDA error";
    case SYLOW_B200_ERR_NOT_ON_CURVE: return "point not on curve";
    case SYLOW_B200_ERR_NOT_IN_SUBGROUP: return "point not in the r-torsion subgroup";
    case SYLOW_B200_ERR_CANNOT_HASH: return "cannot hash to group";
    case SYLOW_B200_ERR_DECODE: return "coordinate is not a canonical field element";
    case SYLOW_B200_ERR_NOMEM: return "out of memory";
    default: return "unknown status";
  }
}
int sylow_b200_last_cuda_error(const sylow_b200_ctx* ctx) {
  if (!ctx) return 0;
  for (const sylow_b200_ctx* c : ctx->children)
    if (c->last_cuda) return c->last_cuda;
  return ctx->last_cuda;
}
uint64_t sylow_b200_launch_count(const sylow_b200_ctx* ctx) {
  if (!ctx) return 0;
  uint64_t t = ctx->launches;
  for (const sylow_b200_ctx* c : ctx->children) t += c->launches;
  return t;
}

// ------------------------------------------------------------------------------- device variants
int sylow_b200_miller_loop_batch_dev(sylow_b200_ctx* ctx, const uint8_t* g1, const uint8_t* g1_inf, const uint8_t* g2,
                                     const uint8_t* g2_inf, size_t n, uint8_t* f_out, void* stream) {
  ENTER_DEV(ctx);
  if (!ctx || (n && (!g1 || !g2 || !f_out))) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  return launch_miller(ctx, g1, g1_inf, g2, g2_inf, n, f_out, 0, pick(ctx, stream));
}

int sylow_b200_final_exp_batch_dev(sylow_b200_ctx* ctx, const uint8_t* f, size_t n, uint8_t* gt_out, void* stream) {
  ENTER_DEV(ctx);
  if (!ctx || (n && (!f || !gt_out))) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  return launch_final_exp(ctx, f, 0, n, gt_out, pick(ctx, stream));
}

// Miller loops then final exponentiations of one slice on one stream.  The Miller values (Montgomery form) are staged
// in gt_out itself; the final exponentiation is in place.
static int pairing_dev_single(sylow_b200_ctx* ctx, const uint8_t* g1, const uint8_t* g1_inf, const uint8_t* g2,
                              const uint8_t* g2_inf, size_t n, uint8_t* gt_out, cudaStream_t s) {
  CKS(launch_miller(ctx, g1, g1_inf, g2, g2_inf, n, gt_out, 1, s));
  return launch_final_exp(ctx, gt_out, 1, n, gt_out, s);
}

// Optional slicing of the device path over the caller's stream and the context's second stream (SYLOW_B200_PAIR_CHUNK =
// slice size in pairs).  Measured (profiles/r02_pair_chunk_sweep.jsonl): the partly filled last wave of one kernel does
// NOT usefully overlap the other stream's kernels - a k_final_exp block needs a whole SM's registers and a k_miller
// block half of them, so the two never share an SM - and slicing only adds wave tails.  Off by default; the tails are
// handled by the low-occupancy tail launches above instead.
static size_t pairing_chunk(const sylow_b200_ctx* ctx, size_t n) {
  static const long env = [] {
    const char* v = getenv("SYLOW_B200_PAIR_CHUNK");
    return v ? atol(v) : 0L;
  }();
  if (env > 0) return (size_t)env;
  (void)ctx;
  return n;  // default: no slicing (the tail launches of launch_miller / launch_final_exp do the work; see below)
}

int sylow_b200_pairing_batch_dev(sylow_b200_ctx* ctx, const uint8_t* g1, const uint8_t* g1_inf, const uint8_t* g2,
                                 const uint8_t* g2_inf, size_t n, uint8_t* gt_out, void* stream) {
  ENTER_DEV(ctx);
  if (!ctx || (n && (!g1 || !g2 || !gt_out))) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  cudaStream_t s = pick(ctx, stream);
  const size_t chunk = pairing_chunk(ctx, n);
  if (n <= chunk + chunk / 2 || s == ctx->stream2) return pairing_dev_single(ctx, g1, g1_inf, g2, g2_inf, n, gt_out, s);
  CK(cudaEventRecord(ctx->ev_fork, s));
  CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
  size_t c = 0;
  for (size_t off = 0; off < n; c++) {
    size_t m = n - off <= chunk + chunk / 2 ? n - off : chunk;
    CKS(pairing_dev_single(ctx, g1 + off * 64, g1_inf ? g1_inf + off : nullptr, g2 + off * 128,
                           g2_inf ? g2_inf + off : nullptr, m, gt_out + off * 384, (c & 1) ? ctx->stream2 : s));
    off += m;
  }
  CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
  CK(cudaStreamWaitEvent(s, ctx->ev_join, 0));
  return 0;
}

// reduce n Montgomery-form Fp12 values at `buf` (clobbered, with `tmp` as ping-pong space of at least
// ceil(n/4) values) to one value written to d_out in canonical (raw_out = 0) or Montgomery form.
static int product_reduce(sylow_b200_ctx* ctx, uint8_t* buf, uint8_t* tmp, size_t n, uint8_t* d_out, int raw_out,
                          cudaStream_t s) {
  uint8_t* cur = buf;
  uint8_t* nxt = tmp;
  while (n > 1) {
    size_t T = n / 4;
    if (T < 1) T = 1;
    if (T > 148 * 256) T = 148 * 256;
    k_fp12_product_strided<<<nblocks(T, SY_SMALL_THREADS), SY_SMALL_THREADS, 0, s>>>(cur, n, nxt, T);
    LAUNCHED(ctx);
    uint8_t* t = cur;
    cur = nxt;
    nxt = t;
    n = T;
  }
  if (raw_out) {
    CK(cudaMemcpyAsync(d_out, cur, 384, cudaMemcpyDeviceToDevice, s));
  } else {
    k_fp12_convert<<<1, 32, 0, s>>>(cur, 1, d_out, 0);
    LAUNCHED(ctx);
  }
  return 0;
}

static const uint8_t kOneCanonical[384] = {1};

int sylow_b200_miller_product_dev(sylow_b200_ctx* ctx, const uint8_t* g1, const uint8_t* g1_inf, const uint8_t* g2,
                                  const uint8_t* g2_inf, size_t n, uint8_t* f_out, void* stream) {
  ENTER_DEV(ctx);
  if (!ctx || !f_out || (n && (!g1 || !g2))) return SYLOW_B200_ERR_ARG;
  cudaStream_t s = pick(ctx, stream);
  if (!n) {
    CK(cudaMemcpyAsync(f_out, kOneCanonical, 384, cudaMemcpyHostToDevice, s));
    return 0;
  }
  CKS(reserve(ctx, ctx->scratch0, n * 384));
  CKS(reserve(ctx, ctx->scratch1, (n / 4 + 1) * 384));
  CKS(launch_miller(ctx, g1, g1_inf, g2, g2_inf, n, ctx->scratch0.p, 1, s));
  return product_reduce(ctx, ctx->scratch0.p, ctx->scratch1.p, n, f_out, 0, s);
}

int sylow_b200_pairing_check_batch_dev(sylow_b200_ctx* ctx, const uint8_t* g1, const uint8_t* g1_inf,
                                       const uint8_t* g2, const uint8_t* g2_inf, size_t k, size_t n_checks,
                                       uint8_t* ok_out, void* stream) {
  ENTER_DEV(ctx);
  if (!ctx || (n_checks && !ok_out)) return SYLOW_B200_ERR_ARG;
  if (!n_checks) return 0;
  size_t n = k * n_checks;
  if (n && (!g1 || !g2)) return SYLOW_B200_ERR_ARG;
  cudaStream_t s = pick(ctx, stream);
  size_t k_prod = k;  // Miller values per check left for k_check_products to multiply
  if (n) {
    CKS(reserve(ctx, ctx->scratch0, n * 384));
    // 2- and 4-pair checks (ecPairing, Groth16 without fixed tables) as ONE glued loop per check: one Fp12 squaring per
    // digit for all pairs of the check, like glued_miller_loop (pairing.rs:970-1022); needs enough checks to fill the GPU
    const size_t wave = (size_t)ctx->sms * SY_GLUED_THREADS * SY_GLUED_MINB;
    if (k == 2 && n_checks >= wave) {
      CKS((launch_glued<2, 0>(ctx, g1, 2, g1_inf, g2, g2_inf, nullptr, 0, nullptr, nullptr, n_checks, ctx->scratch0.p, s)));
      k_prod = 1;
    } else if (k == 4 && n_checks >= wave / 2) {
      CKS((launch_glued<4, 0>(ctx, g1, 4, g1_inf, g2, g2_inf, nullptr, 0, nullptr, nullptr, n_checks, ctx->scratch0.p, s)));
      k_prod = 1;
    } else {
      CKS(launch_miller(ctx, g1, g1_inf, g2, g2_inf, n, ctx->scratch0.p, 1, s));
    }
  }
  return launch_check_products(ctx, ctx->scratch0.p, k_prod, n_checks, ok_out, s);
}

// batches from this size on convert to affine coordinates with one inversion per eight points
#define SY_AFF_MIN_BATCH 32768
static int g1_batch_affine(sylow_b200_ctx* ctx, const uint8_t* d_proj, size_t n, int negate, uint8_t* d_out,
                           uint8_t* d_out_inf, cudaStream_t s) {
  size_t T = (n + SY_AFF_K - 1) / SY_AFF_K;
  k_g1_batch_affine<<<nblocks(T, SY_MUL_THREADS), SY_MUL_THREADS, 0, s>>>(d_proj, n, negate, d_out, d_out_inf);
  LAUNCHED(ctx);
  return 0;
}

int sylow_b200_g1_mul_batch_dev(sylow_b200_ctx* ctx, const uint8_t* pts, const uint8_t* pts_inf,
                                const uint8_t* scalars, size_t n, uint8_t* out, uint8_t* out_inf, void* stream) {
  ENTER_DEV(ctx);
  if (!ctx || (n && (!pts || !scalars || !out))) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  uint8_t* proj = nullptr;
  if (n >= SY_AFF_MIN_BATCH) {
    CKS(reserve(ctx, ctx->proj, n * 96));
    proj = ctx->proj.p;
  }
  k_g1_mul<<<nblocks(n, SY_MUL_THREADS), SY_MUL_THREADS, 0, pick(ctx, stream)>>>(pts, pts_inf, scalars, n,
                                                                                        out, out_inf, proj);
  LAUNCHED(ctx);
  if (proj) CKS(g1_batch_affine(ctx, proj, n, 0, out, out_inf, pick(ctx, stream)));
  return 0;
}
int sylow_b200_g2_mul_batch_dev(sylow_b200_ctx* ctx, const uint8_t* pts, const uint8_t* pts_inf,
                                const uint8_t* scalars, size_t n, uint8_t* out, uint8_t* out_inf, void* stream) {
  ENTER_DEV(ctx);
  if (!ctx || (n && (!pts || !scalars || !out))) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  uint8_t* proj = nullptr;
  if (n >= SY_AFF_MIN_BATCH) {
    CKS(reserve(ctx, ctx->proj, n * 192));
    proj = ctx->proj.p;
  }
  k_g2_mul<<<nblocks(n, SY_G2_THREADS), SY_G2_THREADS, 0, pick(ctx, stream)>>>(pts, pts_inf, scalars,
                                                                                          n, out, out_inf, proj);
  LAUNCHED(ctx);
  if (proj) {
    size_t T = (n + SY_AFF_K - 1) / SY_AFF_K;
    k_g2_batch_affine<<<nblocks(T, SY_MUL_THREADS), SY_MUL_THREADS, 0, pick(ctx, stream)>>>(proj, n, out, out_inf);
    LAUNCHED(ctx);
  }
  return 0;
}

static int make_dst_prime(sylow_b200_ctx* ctx, const uint8_t* dst, size_t dst_len, int hash_id, DstPrime& dp) {
  if (hash_id != SYLOW_B200_HASH_KECCAK256 && hash_id != SYLOW_B200_HASH_SHA256 && hash_id != SYLOW_B200_HASH_SHAKE128)
    return SYLOW_B200_ERR_ARG;
  if (dst_len && !dst) return SYLOW_B200_ERR_ARG;
  memset(dp.b, 0, sizeof(dp.b));
  dp.hash_id = hash_id;
  if (dst_len > 255) {
    // oversize DST (hasher.rs:158-163): hashed on the device, like every other digest of this library
    if (!ctx) return SYLOW_B200_ERR_ARG;
    uint8_t* d_tmp = nullptr;
    CK(cudaMalloc(&d_tmp, dst_len + 32));
    cudaError_t e = cudaMemcpyAsync(d_tmp + 32, dst, dst_len, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
      k_oversize_dst<<<1, 32, 0, ctx->stream>>>(d_tmp + 32, dst_len, hash_id, d_tmp);
      ctx->launches++;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(dp.b, d_tmp, 32, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_tmp);
    if (e != cudaSuccess) return fail_cuda(ctx, e);
    dst_len = 32;
  } else if (dst_len) {
    memcpy(dp.b, dst, dst_len);
  }
  dp.b[dst_len] = (uint8_t)dst_len;  // DST' || I2OSP(len(DST'), 1)   (hasher.rs:205-209)
  dp.len = (uint32_t)dst_len + 1;
  return 0;
}

static int hash_launch(sylow_b200_ctx* ctx, const uint8_t* d_msgs, const uint64_t* d_offsets, size_t n,
                       const DstPrime& dp, int negate, uint8_t* d_out, uint8_t* d_out_inf, cudaStream_t s) {
  uint8_t* proj = nullptr;
  if (n >= SY_AFF_MIN_BATCH) {
    CKS(reserve(ctx, ctx->proj, n * 96));
    proj = ctx->proj.p;
  }
  CK(cudaMemsetAsync(ctx->d_fail, 0, sizeof(int), s));  // the flag reports THIS call's hashes only
  k_hash_to_g1<<<nblocks(n, SY_HASH_THREADS), SY_HASH_THREADS, 0, s>>>(d_msgs, d_offsets, n, dp, negate, d_out,
                                                                     d_out_inf, ctx->d_fail, proj);
  LAUNCHED(ctx);
  if (proj) CKS(g1_batch_affine(ctx, proj, n, negate, d_out, d_out_inf, s));
  return 0;
}

int sylow_b200_hash_to_g1_batch_dev(sylow_b200_ctx* ctx, const uint8_t* d_msgs, const uint64_t* d_offsets, size_t n,
                                    const uint8_t* dst, size_t dst_len, int hash_id, uint8_t* d_out,
                                    uint8_t* d_out_inf, void* stream) {
  ENTER_DEV(ctx);
  if (!ctx || (n && (!d_offsets || !d_out))) return SYLOW_B200_ERR_ARG;
  DstPrime dp;
  CKS(make_dst_prime(ctx, dst, dst_len, hash_id, dp));
  if (!n) return 0;
  return hash_launch(ctx, d_msgs, d_offsets, n, dp, 0, d_out, d_out_inf, pick(ctx, stream));
}

// G2 generator in wire form, uploaded once per context on first use
static const uint64_t kG2GenWords[16] = {
    // x.c0, x.c1, y.c0, y.c1 as 4 LE u64 words each (g2.rs:47-77)
    5106727233969649389ull,  7440829307424791261ull,  4785637993704342649ull,  1729627375292849782ull,
    10945020018377822914ull, 17413811393473931026ull, 8241798111626485029ull,  1841571559660931130ull,
    5541340697920699818ull,  16416156555105522555ull, 5380518976772849807ull,  1353435754470862315ull,
    6173549831154472795ull,  13567992399387660019ull, 17050234209342075797ull, 650358724130500725ull};

#define SY_TABLE_BYTES (87 * 192)

}  // extern "C" (templates need C++ linkage)

template <int NV, int NF>
static int launch_glued(sylow_b200_ctx* ctx, const uint8_t* g1v, size_t sv, const uint8_t* g1v_inf, const uint8_t* g2v,
                        const uint8_t* g2v_inf, const uint8_t* g1f, size_t sf, const uint8_t* g1f_inf,
                        const uint8_t* tables, size_t n, uint8_t* f_out, cudaStream_t s) {
  // the opt-in for > 48 KB of dynamic shared memory is per device: remember it per context
  size_t smem = (size_t)NF * SY_TABLE_BYTES + SY_MILLER_SMEM_BYTES(SY_GLUED_THREADS);
  unsigned bit = 1u << (NV * 4 + NF);
  if (!(ctx->glued_attr_mask & bit)) {
    CK(cudaFuncSetAttribute(k_glued<NV, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->glued_attr_mask |= bit;
  }
  k_glued<NV, NF><<<nblocks(n, SY_GLUED_THREADS), SY_GLUED_THREADS, smem, s>>>(g1v, sv, g1v_inf, g2v, g2v_inf, g1f, sf,
                                                                              g1f_inf, tables, n, f_out);
  LAUNCHED(ctx);
  return 0;
}

extern "C" {

// G2PreComputed of the generator, computed on the device once per context
static int ensure_gen_table(sylow_b200_ctx* ctx, cudaStream_t s) {
  if (ctx->d_gen_table) return 0;
  // the field is set only after the table has been computed: a failure on the way leaves it null (and frees both
  // buffers), so the next call builds it again instead of using an uninitialised table
  uint8_t *d_gen = nullptr, *d_tab = nullptr;
  cudaError_t e = cudaMalloc(&d_gen, 128);
  if (e == cudaSuccess) e = cudaMalloc(&d_tab, SY_TABLE_BYTES);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_gen, kG2GenWords, 128, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) {
    k_g2_precompute<<<1, 32, 0, s>>>(d_gen, 1, d_tab, 1);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (d_gen) cudaFree(d_gen);
  if (e != cudaSuccess) {
    if (d_tab) cudaFree(d_tab);
    return fail_cuda(ctx, e);
  }
  ctx->d_gen_table = d_tab;
  return 0;
}

// f[i] = miller(sig_i, G2gen) * miller(-H(m_i), pk_i) for every signature, Montgomery form, in scratch0:
// one 2-pair glued loop per thread (shared squaring; the generator's lines come from shared memory).  An infinite
// signature or key makes its pair contribute 1, like pairing() does (pairing.rs:876-886).
static int verify_miller_values(sylow_b200_ctx* ctx, const uint8_t* d_pks, const uint8_t* d_pks_inf, const uint8_t* d_msgs,
                                const uint64_t* d_offsets, const uint8_t* d_sigs, const uint8_t* d_sigs_inf, size_t n,
                                const DstPrime& dp, cudaStream_t s) {
  CKS(ensure_gen_table(ctx, s));
  CKS(reserve(ctx, ctx->scratch0, n * 384));
  CKS(reserve(ctx, ctx->scratch2, n * 65 + 64));
  uint8_t* d_hm = ctx->scratch2.p;
  uint8_t* d_hm_inf = d_hm + n * 64;
  CKS(hash_launch(ctx, d_msgs, d_offsets, n, dp, 1, d_hm, d_hm_inf, s));
  return launch_glued<1, 1>(ctx, d_hm, 1, d_hm_inf, d_pks, d_pks_inf, d_sigs, 1, d_sigs_inf, ctx->d_gen_table, n,
                            ctx->scratch0.p, s);
}

// sum of n affine G1 points (wire format, optional infinity flags) -> one affine point + flag at d_out / d_out_inf
// (device).  Tree of strided partial sums in projective coordinates, then one inversion.  Launched <<<1, 1>>>
// at the end because fp_pow re-converges with __syncthreads().
static int g1_sum_reduce(sylow_b200_ctx* ctx, const uint8_t* d_pts, const uint8_t* d_inf, size_t n, int negate,
                         uint8_t* d_out, uint8_t* d_out_inf, cudaStream_t s, int affine_in = 1) {
  size_t T = n / 8;
  if (T < 1) T = 1;
  if (T > 148 * 512) T = 148 * 512;
  CKS(reserve(ctx, ctx->sum0, T * 96));
  CKS(reserve(ctx, ctx->sum1, (T / 4 + 1) * 96));
  k_g1_sum_strided<<<nblocks(T, SY_MUL_THREADS), SY_MUL_THREADS, 0, s>>>(d_pts, d_inf, affine_in, n, ctx->sum0.p, T);
  LAUNCHED(ctx);
  uint8_t* cur = ctx->sum0.p;
  uint8_t* nxt = ctx->sum1.p;
  size_t m = T;
  while (m > 1) {
    size_t T2 = m / 4;
    if (T2 < 1) T2 = 1;
    k_g1_sum_strided<<<nblocks(T2, SY_MUL_THREADS), SY_MUL_THREADS, 0, s>>>(cur, nullptr, 0, m, nxt, T2);
    LAUNCHED(ctx);
    uint8_t* t = cur;
    cur = nxt;
    nxt = t;
    m = T2;
  }
  k_g1_finish_sum<<<1, 1, 0, s>>>(cur, negate, d_out, d_out_inf);
  LAUNCHED(ctx);
  return 0;
}

static bool load_seed(const uint8_t* weight_seed, WeightSeed& ws) {
  if (!weight_seed) return false;
  memcpy(ws.w, weight_seed, 32);
  return true;
}

// `pairs_ready` (may be null): an event after which the keys and signatures are in device memory - the host entry point
// copies them on its copy stream while the hash-to-curve kernel, which only needs the messages, already runs
static int verify_partial_core(sylow_b200_ctx* ctx, const uint8_t* d_pks, const uint8_t* d_pks_inf, const uint8_t* d_msgs,
                               const uint64_t* d_offsets, const uint8_t* d_sigs, const uint8_t* d_sigs_inf, size_t n,
                               const uint8_t* dst, size_t dst_len, int hash_id, const uint8_t* weight_seed,
                               uint64_t first_index, uint8_t* d_f_out, void* stream, cudaEvent_t pairs_ready) {
  if (!ctx || !d_f_out || (n && (!d_pks || !d_offsets || !d_sigs))) return SYLOW_B200_ERR_ARG;
  DstPrime dp;
  CKS(make_dst_prime(ctx, dst, dst_len, hash_id, dp));
  cudaStream_t s = pick(ctx, stream);
  if (!n) {
    CK(cudaMemcpyAsync(d_f_out, kOneCanonical, 384, cudaMemcpyHostToDevice, s));
    return 0;
  }
  // prod_i e(sig_i, G2gen) = e(sum_i sig_i, G2gen): the n signature pairs collapse into ONE Miller loop against
  // the generator (bilinearity; the Gt value after the final exponentiation is the same), so the slice costs
  // n fused Miller loops (-H(m_i), pk_i), n point additions and one extra loop.  With weights the same identity
  // holds for r_i sig_i and r_i H(m_i): two 64-bit ladders per signature on top.
  WeightSeed ws;
  const bool weighted = load_seed(weight_seed, ws);
  CKS(ensure_gen_table(ctx, s));
  CKS(reserve(ctx, ctx->scratch0, (n + 1) * 384));
  CKS(reserve(ctx, ctx->scratch1, ((n + 1) / 4 + 1) * 384));
  CKS(reserve(ctx, ctx->scratch2, n * 65 + 256));
  uint8_t* d_hm = ctx->scratch2.p;
  uint8_t* d_hm_inf = d_hm + n * 64;
  uint8_t* d_sum = d_hm + ((n * 65 + 63) / 64) * 64;  // 64 B point + flag
  uint8_t* d_sum_inf = d_sum + 64;
  CKS(hash_launch(ctx, d_msgs, d_offsets, n, dp, 1, d_hm, d_hm_inf, s));
  if (pairs_ready) CK(cudaStreamWaitEvent(s, pairs_ready, 0));
  if (weighted) {
    CKS(reserve(ctx, ctx->proj, n * 96));
    k_g1_mul_weight<<<nblocks(n, SY_MUL_THREADS), SY_MUL_THREADS, 0, s>>>(d_hm, d_hm_inf, ws, first_index, n, ctx->proj.p);
    LAUNCHED(ctx);
    CKS(g1_batch_affine(ctx, ctx->proj.p, n, 0, d_hm, d_hm_inf, s));
  }
  // The n Miller loops (-r_i H(m_i), pk_i) only ever meet as a product, so SY_VERIFY_GLUE signatures share one thread
  // and ONE Fp12 squaring per loop digit - what glued_miller_loop does for the whole batch in the reference
  // (pairing.rs:970-1022).  A thread's 2 / 4 pairs save 63 x 36 of the 8 444 multiplications of every pair after its
  // first.  The n % GLUE leftover pairs go through k_miller.
  static const int glue_env = [] {
    const char* v = getenv("SYLOW_B200_VERIFY_GLUE");
    int g = v ? atoi(v) : SY_VERIFY_GLUE;
    return g == 1 || g == 2 || g == 4 ? g : 0;
  }();
  int glue = glue_env;
  if (!glue) {
    // fewest wave-times: whole waves of n / g threads (256 resident per SM) times the multiplications of g glued pairs.
    // Measured at 2^20 signatures (profiles/r02_verify_glue.jsonl): 192.4 / 171.2 / 163.3 ms for 1 / 2 / 4; at 2^13 the
    // order reverses (9.3 / 12.4 / 18.4 ms: a quarter of the threads, each four times as long).
    const size_t wave = (size_t)ctx->sms * SY_GLUED_THREADS * SY_GLUED_MINB;
    const size_t cost[3] = {8444, 2 * 8444 - 2268, 4 * 8444 - 3 * 2268};
    size_t best = ~(size_t)0;
    for (int j = 0; j < 3; j++) {
      const size_t g = (size_t)1 << j, waves = (n / g + wave - 1) / wave;
      if (n >= g && waves * cost[j] < best) {
        best = waves * cost[j];
        glue = (int)g;
      }
    }
  }
  size_t n_f = n;  // Miller values in scratch0 before the signature pair's value is appended
  if (glue > 1 && n >= (size_t)glue) {
    const size_t items = n / glue, done = items * glue;
    if (glue == 2)
      CKS((launch_glued<2, 0>(ctx, d_hm, 2, d_hm_inf, d_pks, d_pks_inf, nullptr, 0, nullptr, nullptr, items, ctx->scratch0.p, s)));
    else
      CKS((launch_glued<4, 0>(ctx, d_hm, 4, d_hm_inf, d_pks, d_pks_inf, nullptr, 0, nullptr, nullptr, items, ctx->scratch0.p, s)));
    if (done < n)
      CKS(launch_miller(ctx, d_hm + done * 64, d_hm_inf + done, d_pks + done * 128, d_pks_inf ? d_pks_inf + done : nullptr,
                        n - done, ctx->scratch0.p + items * 384, 1, s));
    n_f = items + (n - done);
  } else {
    CKS(launch_miller(ctx, d_hm, d_hm_inf, d_pks, d_pks_inf, n, ctx->scratch0.p, 1, s));
  }
  if (weighted) {
    k_g1_mul_weight<<<nblocks(n, SY_MUL_THREADS), SY_MUL_THREADS, 0, s>>>(d_sigs, d_sigs_inf, ws, first_index, n,
                                                                       ctx->proj.p);
    LAUNCHED(ctx);
    CKS(g1_sum_reduce(ctx, ctx->proj.p, nullptr, n, 0, d_sum, d_sum_inf, s, 0));
  } else {
    CKS(g1_sum_reduce(ctx, d_sigs, d_sigs_inf, n, 0, d_sum, d_sum_inf, s));
  }
  CKS((launch_glued<0, 1>(ctx, nullptr, 0, nullptr, nullptr, nullptr, d_sum, 1, d_sum_inf, ctx->d_gen_table, 1,
                          ctx->scratch0.p + n_f * 384, s)));
  return product_reduce(ctx, ctx->scratch0.p, ctx->scratch1.p, n_f + 1, d_f_out, 0, s);
}

int sylow_b200_verify_batch_partial_dev(sylow_b200_ctx* ctx, const uint8_t* d_pks, const uint8_t* d_pks_inf,
                                        const uint8_t* d_msgs, const uint64_t* d_offsets, const uint8_t* d_sigs,
                                        const uint8_t* d_sigs_inf, size_t n, const uint8_t* dst, size_t dst_len,
                                        int hash_id, const uint8_t* weight_seed, uint64_t first_index, uint8_t* d_f_out,
                                        void* stream) {
  ENTER_DEV(ctx);
  return verify_partial_core(ctx, d_pks, d_pks_inf, d_msgs, d_offsets, d_sigs, d_sigs_inf, n, dst, dst_len, hash_id,
                             weight_seed, first_index, d_f_out, stream, nullptr);
}

// *failed = 1 if a hash-to-curve of the last hashing `_dev` call enqueued on `stream` hit SvdW's failing square-root
// check (GroupError::CannotHashToGroup); synchronises the stream and clears the flag.
int sylow_b200_hash_failed_dev(sylow_b200_ctx* ctx, void* stream, int* failed) {
  ENTER_DEV(ctx);
  if (!ctx || !failed) return SYLOW_B200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = pick(ctx, stream);
  int f = 0;
  CK(cudaMemcpyAsync(&f, ctx->d_fail, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (f) CK(cudaMemsetAsync(ctx->d_fail, 0, sizeof(int), s));
  *failed = f;
  return 0;
}

// ------------------------------------------------------------------------------- host variants
static int to_dev(sylow_b200_ctx* ctx, DevBuf& b, const void* h, size_t bytes, const uint8_t** d) {
  *d = nullptr;
  if (!h || !bytes) return 0;
  CKS(reserve(ctx, b, bytes));
  CK(cudaMemcpyAsync(b.p, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
  *d = b.p;
  return 0;
}
static int finish(sylow_b200_ctx* ctx) {
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}


// Host-pointer pairing / Miller-loop batch.  Large batches are cut into chunks of whole waves of both kernels (SMs x
// 1536 pairs: six Miller-loop waves, four final-exponentiation waves) and pipelined over three streams: the
// host->device copy of chunk c+1 and the device->host copy of chunk c-1 run under the kernels of chunk c, so the
// PCIe time of the 576 bytes per pairing disappears from the call.
static int pairing_host(sylow_b200_ctx* ctx, const uint8_t* g1, const uint8_t* g1_inf, const uint8_t* g2,
                        const uint8_t* g2_inf, size_t n, uint8_t* out, bool final_exp) {
  const uint8_t *d1i, *d2i;
  CKS(reserve(ctx, ctx->in_a, n * 64));
  CKS(reserve(ctx, ctx->in_b, n * 128));
  CKS(reserve(ctx, ctx->out, n * 384));
  CKS(to_dev(ctx, ctx->flag_a, g1_inf, n, &d1i));
  CKS(to_dev(ctx, ctx->flag_b, g2_inf, n, &d2i));
  if (d1i || d2i) CK(cudaStreamSynchronize(ctx->stream));  // the flags are read from both compute streams
  // whole waves of both kernels: a multiple of the threads each of them keeps resident per SM
  const size_t chunk =
      (size_t)ctx->sms * sy_lcm(sy_lcm(SY_MILLER_THREADS * SY_MILLER_MINB, SY_FEXP_THREADS * SY_FEXP_MINB), 768) * 2;
  const size_t n_chunks = n <= 2 * chunk ? 1 : (n + chunk - 1) / chunk;
  const size_t step = n_chunks == 1 ? n : chunk;
  std::vector<cudaEvent_t> ev(2 * n_chunks, nullptr);
  int rc = 0;
  cudaError_t e = cudaSuccess;
  for (size_t i = 0; i < ev.size() && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
  auto d2h = [&](size_t c) {
    size_t off = c * step, m = off + step < n ? step : n - off;
    cudaError_t r = cudaStreamWaitEvent(ctx->copy_out, ev[2 * c + 1], 0);
    if (r == cudaSuccess)
      r = cudaMemcpyAsync(out + off * 384, ctx->out.p + off * 384, m * 384, cudaMemcpyDeviceToHost, ctx->copy_out);
    return r;
  };
  for (size_t c = 0; c < n_chunks && e == cudaSuccess && rc == 0; c++) {
    size_t off = c * step, m = off + step < n ? step : n - off;
    e = cudaMemcpyAsync(ctx->in_a.p + off * 64, g1 + off * 64, m * 64, cudaMemcpyHostToDevice, ctx->copy_in);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(ctx->in_b.p + off * 128, g2 + off * 128, m * 128, cudaMemcpyHostToDevice, ctx->copy_in);
    if (e == cudaSuccess) e = cudaEventRecord(ev[2 * c], ctx->copy_in);
    cudaStream_t cs = (c & 1) ? ctx->stream2 : ctx->stream;
    if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, ev[2 * c], 0);
    if (e != cudaSuccess) break;
    const uint8_t* f1 = d1i ? d1i + off : nullptr;
    const uint8_t* f2 = d2i ? d2i + off : nullptr;
    uint8_t* o = ctx->out.p + off * 384;
    if (final_exp)
      rc = pairing_dev_single(ctx, ctx->in_a.p + off * 64, f1, ctx->in_b.p + off * 128, f2, m, o, cs);
    else
      rc = sylow_b200_miller_loop_batch_dev(ctx, ctx->in_a.p + off * 64, f1, ctx->in_b.p + off * 128, f2, m, o, cs);
    if (rc) break;
    e = cudaEventRecord(ev[2 * c + 1], cs);
    // the device->host copy of the PREVIOUS chunk is issued after this chunk's kernels are queued: with pageable
    // host memory that copy blocks the calling thread, and the GPU then still has a chunk of work in front of it
    if (e == cudaSuccess && c > 0) e = d2h(c - 1);
  }
  if (e == cudaSuccess && rc == 0) e = d2h(n_chunks - 1);
  cudaError_t e2 = cudaStreamSynchronize(ctx->copy_out);
  cudaError_t e3 = cudaStreamSynchronize(ctx->stream);
  cudaError_t e4 = cudaStreamSynchronize(ctx->copy_in);
  cudaError_t e5 = cudaStreamSynchronize(ctx->stream2);
  for (cudaEvent_t v : ev)
    if (v) cudaEventDestroy(v);
  if (rc) return rc;
  if (e == cudaSuccess) e = e2;
  if (e == cudaSuccess) e = e3;
  if (e == cudaSuccess) e = e4;
  if (e == cudaSuccess) e = e5;
  if (e != cudaSuccess) return fail_cuda(ctx, e);
  return 0;
}

}  // extern "C" (templates need C++ linkage)

// ------------------------------------------------------------------------------- multi-device dispatch
// A context made by sylow_b200_create_multi shards the batch as contiguous slices [g n / G, (g + 1) n / G), one
// host thread per GPU driving that GPU's own single-device context.  Per-item outputs land in the
// caller's buffers directly; product forms combine the G 384-byte partials on the first device.  Nothing throws
// across the ABI: thread creation failures come back as SYLOW_B200_ERR_NOMEM.
template <class Fn>
static int multi_slices(sylow_b200_ctx* ctx, size_t n, Fn fn) {
  const size_t G = ctx->children.size();
  std::vector<int> rc(G, 0);
  try {
    std::vector<std::thread> th;
    th.reserve(G);
    for (size_t g = 0; g < G; g++) {
      size_t off = g * n / G, m = (g + 1) * n / G - off;
      if (!m) continue;
      sylow_b200_ctx* c = ctx->children[g];
      th.emplace_back([&rc, &fn, c, g, off, m] { rc[g] = fn(c, g, off, m); });
    }
    for (std::thread& t : th) t.join();
  } catch (...) {
    return SYLOW_B200_ERR_NOMEM;
  }
  for (int r : rc)
    if (r) return r;
  return 0;
}
static inline bool is_multi(const sylow_b200_ctx* ctx) { return ctx && ctx->children.size() > 1; }
static inline const uint8_t* adv(const uint8_t* p, size_t bytes) { return p ? p + bytes : nullptr; }
static inline uint8_t* adv(uint8_t* p, size_t bytes) { return p ? p + bytes : nullptr; }
// messages of the slice [off, off + m): the byte range and offsets rebased to 0
struct MsgSlice {
  const uint8_t* msgs;
  std::vector<uint64_t> offs;
  MsgSlice(const uint8_t* all, const uint64_t* offsets, size_t off, size_t m) : msgs(all ? all + offsets[off] : nullptr), offs(m + 1) {
    for (size_t i = 0; i <= m; i++) offs[i] = offsets[off + i] - offsets[off];
  }
};

extern "C" {

int sylow_b200_pairing_batch(sylow_b200_ctx* ctx, const uint8_t* g1, const uint8_t* g1_inf, const uint8_t* g2,
                             const uint8_t* g2_inf, size_t n, uint8_t* gt_out) {
  if (is_multi(ctx) && g1 && g2 && gt_out)
    return multi_slices(ctx, n, [=](sylow_b200_ctx* c, size_t, size_t off, size_t m) {
      return sylow_b200_pairing_batch(c, g1 + off * 64, adv(g1_inf, off), g2 + off * 128, adv(g2_inf, off), m,
                                      gt_out + off * 384);
    });
  ENTER(ctx);
  if (n && (!g1 || !g2 || !gt_out)) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  return pairing_host(ctx, g1, g1_inf, g2, g2_inf, n, gt_out, true);
}

int sylow_b200_miller_loop_batch(sylow_b200_ctx* ctx, const uint8_t* g1, const uint8_t* g1_inf, const uint8_t* g2,
                                 const uint8_t* g2_inf, size_t n, uint8_t* f_out) {
  if (is_multi(ctx) && g1 && g2 && f_out)
    return multi_slices(ctx, n, [=](sylow_b200_ctx* c, size_t, size_t off, size_t m) {
      return sylow_b200_miller_loop_batch(c, g1 + off * 64, adv(g1_inf, off), g2 + off * 128, adv(g2_inf, off), m,
                                          f_out + off * 384);
    });
  ENTER(ctx);
  if (n && (!g1 || !g2 || !f_out)) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  return pairing_host(ctx, g1, g1_inf, g2, g2_inf, n, f_out, false);
}

int sylow_b200_miller_product(sylow_b200_ctx* ctx, const uint8_t* g1, const uint8_t* g1_inf, const uint8_t* g2,
                              const uint8_t* g2_inf, size_t n, uint8_t f_out[384]) {
  if (is_multi(ctx) && g1 && g2 && f_out && n) {
    const size_t G = ctx->children.size();
    std::vector<uint8_t> part(G * 384, 0);
    for (size_t g = 0; g < G; g++) part[g * 384] = 1;  // empty slices contribute 1
    uint8_t* pp = part.data();
    CKS(multi_slices(ctx, n, [=](sylow_b200_ctx* c, size_t g, size_t off, size_t m) {
      return sylow_b200_miller_product(c, g1 + off * 64, adv(g1_inf, off), g2 + off * 128, adv(g2_inf, off), m,
                                       pp + g * 384);
    }));
    return sylow_b200_fp12_product(ctx->children[0], pp, G, f_out);
  }
  ENTER(ctx);
  if (!f_out || (n && (!g1 || !g2))) return SYLOW_B200_ERR_ARG;
  const uint8_t *d1, *d1i, *d2, *d2i;
  CKS(to_dev(ctx, ctx->in_a, g1, n * 64, &d1));
  CKS(to_dev(ctx, ctx->flag_a, g1_inf, n, &d1i));
  CKS(to_dev(ctx, ctx->in_b, g2, n * 128, &d2));
  CKS(to_dev(ctx, ctx->flag_b, g2_inf, n, &d2i));
  CKS(reserve(ctx, ctx->out, 384));
  CKS(sylow_b200_miller_product_dev(ctx, d1, d1i, d2, d2i, n, ctx->out.p, nullptr));
  CK(cudaMemcpyAsync(f_out, ctx->out.p, 384, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

int sylow_b200_final_exp_batch(sylow_b200_ctx* ctx, const uint8_t* f, size_t n, uint8_t* gt_out) {
  if (is_multi(ctx) && f && gt_out && n >= 4096)
    return multi_slices(ctx, n, [=](sylow_b200_ctx* c, size_t, size_t off, size_t m) {
      return sylow_b200_final_exp_batch(c, f + off * 384, m, gt_out + off * 384);
    });
  ENTER(ctx);
  if (n && (!f || !gt_out)) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  const uint8_t* df;
  CKS(to_dev(ctx, ctx->in_a, f, n * 384, &df));
  CKS(reserve(ctx, ctx->out, n * 384));
  CKS(sylow_b200_final_exp_batch_dev(ctx, df, n, ctx->out.p, nullptr));
  CK(cudaMemcpyAsync(gt_out, ctx->out.p, n * 384, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

int sylow_b200_fp12_product(sylow_b200_ctx* ctx, const uint8_t* f, size_t n, uint8_t out[384]) {
  ENTER(ctx);
  if (!out || (n && !f)) return SYLOW_B200_ERR_ARG;
  if (!n) {
    memcpy(out, kOneCanonical, 384);
    return 0;
  }
  const uint8_t* df;
  CKS(to_dev(ctx, ctx->in_a, f, n * 384, &df));
  CKS(reserve(ctx, ctx->scratch0, n * 384));
  CKS(reserve(ctx, ctx->scratch1, (n / 4 + 1) * 384));
  CKS(reserve(ctx, ctx->out, 384));
  k_fp12_convert<<<nblocks(n, 128), 128, 0, ctx->stream>>>(df, n, ctx->scratch0.p, 1);
  LAUNCHED(ctx);
  CKS(product_reduce(ctx, ctx->scratch0.p, ctx->scratch1.p, n, ctx->out.p, 0, ctx->stream));
  CK(cudaMemcpyAsync(out, ctx->out.p, 384, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

int sylow_b200_pairing_check_batch(sylow_b200_ctx* ctx, const uint8_t* g1, const uint8_t* g1_inf, const uint8_t* g2,
                                   const uint8_t* g2_inf, size_t k, size_t n_checks, uint8_t* ok_out) {
  if (is_multi(ctx) && g1 && g2 && ok_out)
    return multi_slices(ctx, n_checks, [=](sylow_b200_ctx* c, size_t, size_t off, size_t m) {
      return sylow_b200_pairing_check_batch(c, g1 + off * k * 64, adv(g1_inf, off * k), g2 + off * k * 128,
                                            adv(g2_inf, off * k), k, m, ok_out + off);
    });
  ENTER(ctx);
  if (n_checks && !ok_out) return SYLOW_B200_ERR_ARG;
  if (!n_checks) return 0;
  size_t n = k * n_checks;
  if (n && (!g1 || !g2)) return SYLOW_B200_ERR_ARG;
  const uint8_t *d1, *d1i, *d2, *d2i;
  CKS(to_dev(ctx, ctx->in_a, g1, n * 64, &d1));
  CKS(to_dev(ctx, ctx->flag_a, g1_inf, n, &d1i));
  CKS(to_dev(ctx, ctx->in_b, g2, n * 128, &d2));
  CKS(to_dev(ctx, ctx->flag_b, g2_inf, n, &d2i));
  CKS(reserve(ctx, ctx->out, n_checks));
  CKS(sylow_b200_pairing_check_batch_dev(ctx, d1, d1i, d2, d2i, k, n_checks, ctx->out.p, nullptr));
  CK(cudaMemcpyAsync(ok_out, ctx->out.p, n_checks, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

static int mul_host(sylow_b200_ctx* ctx, int g2, const uint8_t* pts, const uint8_t* pts_inf, const uint8_t* scalars,
                    size_t n, uint8_t* out, uint8_t* out_inf) {
  if (is_multi(ctx) && pts && scalars && out) {
    const size_t pb = g2 ? 128 : 64;
    return multi_slices(ctx, n, [=](sylow_b200_ctx* c, size_t, size_t off, size_t m) {
      return mul_host(c, g2, pts + off * pb, adv(pts_inf, off), scalars + off * 32, m, out + off * pb, adv(out_inf, off));
    });
  }
  ENTER(ctx);
  if (n && (!pts || !scalars || !out)) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  size_t pb = g2 ? 128 : 64;
  const uint8_t *dp, *dpi, *dk;
  CKS(to_dev(ctx, ctx->in_a, pts, n * pb, &dp));
  CKS(to_dev(ctx, ctx->flag_a, pts_inf, n, &dpi));
  CKS(to_dev(ctx, ctx->in_b, scalars, n * 32, &dk));
  CKS(reserve(ctx, ctx->out, n * pb));
  CKS(reserve(ctx, ctx->flag_b, n));
  if (g2)
    CKS(sylow_b200_g2_mul_batch_dev(ctx, dp, dpi, dk, n, ctx->out.p, ctx->flag_b.p, nullptr));
  else
    CKS(sylow_b200_g1_mul_batch_dev(ctx, dp, dpi, dk, n, ctx->out.p, ctx->flag_b.p, nullptr));
  CK(cudaMemcpyAsync(out, ctx->out.p, n * pb, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_inf) CK(cudaMemcpyAsync(out_inf, ctx->flag_b.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}
int sylow_b200_g1_mul_batch(sylow_b200_ctx* ctx, const uint8_t* pts, const uint8_t* pts_inf, const uint8_t* scalars,
                            size_t n, uint8_t* out, uint8_t* out_inf) {
  return mul_host(ctx, 0, pts, pts_inf, scalars, n, out, out_inf);
}
int sylow_b200_g2_mul_batch(sylow_b200_ctx* ctx, const uint8_t* pts, const uint8_t* pts_inf, const uint8_t* scalars,
                            size_t n, uint8_t* out, uint8_t* out_inf) {
  return mul_host(ctx, 1, pts, pts_inf, scalars, n, out, out_inf);
}

static int check_hash_fail(sylow_b200_ctx* ctx) {
  int f = 0;
  CK(cudaMemcpyAsync(&f, ctx->d_fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (f) {
    CK(cudaMemsetAsync(ctx->d_fail, 0, sizeof(int), ctx->stream));
    return SYLOW_B200_ERR_CANNOT_HASH;
  }
  return 0;
}

static int msgs_to_dev(sylow_b200_ctx* ctx, const uint8_t* msgs, const uint64_t* offsets, size_t n,
                       const uint8_t** d_msgs, const uint64_t** d_off) {
  if (!offsets) return SYLOW_B200_ERR_ARG;
  if (offsets[0] != 0) return SYLOW_B200_ERR_ARG;
  for (size_t i = 0; i < n; i++)
    if (offsets[i + 1] < offsets[i]) return SYLOW_B200_ERR_ARG;
  size_t total = (size_t)offsets[n];
  if (total && !msgs) return SYLOW_B200_ERR_ARG;
  const uint8_t* dm;
  const uint8_t* dof;
  CKS(to_dev(ctx, ctx->in_c, msgs, total, &dm));
  if (!dm) {  // all messages empty: still need a valid pointer
    CKS(reserve(ctx, ctx->in_c, 16));
    dm = ctx->in_c.p;
  }
  CKS(to_dev(ctx, ctx->in_d, offsets, (n + 1) * sizeof(uint64_t), &dof));
  *d_msgs = dm;
  *d_off = reinterpret_cast<const uint64_t*>(dof);
  return 0;
}

int sylow_b200_hash_to_g1_batch(sylow_b200_ctx* ctx, const uint8_t* msgs, const uint64_t* offsets, size_t n,
                                const uint8_t* dst, size_t dst_len, int hash_id, uint8_t* out, uint8_t* out_inf) {
  if (is_multi(ctx) && offsets && out && n)
    return multi_slices(ctx, n, [=](sylow_b200_ctx* c, size_t, size_t off, size_t m) {
      MsgSlice ms(msgs, offsets, off, m);
      return sylow_b200_hash_to_g1_batch(c, ms.msgs, ms.offs.data(), m, dst, dst_len, hash_id, out + off * 64,
                                         adv(out_inf, off));
    });
  ENTER(ctx);
  DstPrime dp;
  CKS(make_dst_prime(ctx, dst, dst_len, hash_id, dp));
  if (!n) return 0;
  if (!out) return SYLOW_B200_ERR_ARG;
  const uint8_t* dm;
  const uint64_t* dof;
  CKS(msgs_to_dev(ctx, msgs, offsets, n, &dm, &dof));
  CKS(reserve(ctx, ctx->out, n * 64));
  CKS(reserve(ctx, ctx->flag_b, n));
  CKS(hash_launch(ctx, dm, dof, n, dp, 0, ctx->out.p, ctx->flag_b.p, ctx->stream));
  CK(cudaMemcpyAsync(out, ctx->out.p, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_inf) CK(cudaMemcpyAsync(out_inf, ctx->flag_b.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  return check_hash_fail(ctx);
}

int sylow_b200_sign_batch(sylow_b200_ctx* ctx, const uint8_t* sks, const uint8_t* msgs, const uint64_t* offsets,
                          size_t n, const uint8_t* dst, size_t dst_len, int hash_id, uint8_t* sigs_out,
                          uint8_t* sigs_out_inf) {
  if (is_multi(ctx) && offsets && sks && sigs_out && n)
    return multi_slices(ctx, n, [=](sylow_b200_ctx* c, size_t, size_t off, size_t m) {
      MsgSlice ms(msgs, offsets, off, m);
      return sylow_b200_sign_batch(c, sks + off * 32, ms.msgs, ms.offs.data(), m, dst, dst_len, hash_id,
                                   sigs_out + off * 64, adv(sigs_out_inf, off));
    });
  ENTER(ctx);
  DstPrime dp;
  CKS(make_dst_prime(ctx, dst, dst_len, hash_id, dp));
  if (!n) return 0;
  if (!sks || !sigs_out) return SYLOW_B200_ERR_ARG;
  const uint8_t *dm, *dk;
  const uint64_t* dof;
  CKS(msgs_to_dev(ctx, msgs, offsets, n, &dm, &dof));
  CKS(to_dev(ctx, ctx->in_b, sks, n * 32, &dk));
  CKS(reserve(ctx, ctx->scratch2, n * 65));
  CKS(reserve(ctx, ctx->out, n * 64));
  CKS(reserve(ctx, ctx->flag_b, n));
  CKS(hash_launch(ctx, dm, dof, n, dp, 0, ctx->scratch2.p, ctx->scratch2.p + n * 64, ctx->stream));
  CKS(sylow_b200_g1_mul_batch_dev(ctx, ctx->scratch2.p, ctx->scratch2.p + n * 64, dk, n, ctx->out.p, ctx->flag_b.p,
                                  nullptr));
  CK(cudaMemcpyAsync(sigs_out, ctx->out.p, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
  if (sigs_out_inf) CK(cudaMemcpyAsync(sigs_out_inf, ctx->flag_b.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  return check_hash_fail(ctx);
}

int sylow_b200_verify_batch_partial(sylow_b200_ctx* ctx, const uint8_t* pks, const uint8_t* pks_inf, const uint8_t* msgs,
                                    const uint64_t* offsets, const uint8_t* sigs, const uint8_t* sigs_inf, size_t n,
                                    const uint8_t* dst, size_t dst_len, int hash_id, const uint8_t* weight_seed,
                                    uint64_t first_index, uint8_t f_out[384]) {
  if (is_multi(ctx) && offsets && pks && sigs && f_out && n) {
    const size_t G = ctx->children.size();
    std::vector<uint8_t> part(G * 384, 0);
    for (size_t g = 0; g < G; g++) part[g * 384] = 1;
    uint8_t* pp = part.data();
    CKS(multi_slices(ctx, n, [=](sylow_b200_ctx* c, size_t g, size_t off, size_t m) {
      MsgSlice ms(msgs, offsets, off, m);
      return sylow_b200_verify_batch_partial(c, pks + off * 128, adv(pks_inf, off), ms.msgs, ms.offs.data(),
                                             sigs + off * 64, adv(sigs_inf, off), m, dst, dst_len, hash_id,
                                             weight_seed, first_index + off, pp + g * 384);
    }));
    return sylow_b200_fp12_product(ctx->children[0], pp, G, f_out);
  }
  ENTER(ctx);
  if (!f_out) return SYLOW_B200_ERR_ARG;
  DstPrime dp;
  CKS(make_dst_prime(ctx, dst, dst_len, hash_id, dp));
  if (!n) {
    memcpy(f_out, kOneCanonical, 384);
    return 0;
  }
  if (!pks || !sigs) return SYLOW_B200_ERR_ARG;
  const uint8_t *dm, *dpk, *dsg, *dpki, *dsgi;
  const uint64_t* dof;
  // messages first, on the compute stream: the hash-to-curve kernel starts as soon as they are there, while the keys
  // and signatures (six times the bytes) still travel on the copy stream
  CKS(msgs_to_dev(ctx, msgs, offsets, n, &dm, &dof));
  CKS(reserve(ctx, ctx->in_a, n * 128));
  CKS(reserve(ctx, ctx->in_b, n * 64));
  if (pks_inf) CKS(reserve(ctx, ctx->flag_a, n));
  if (sigs_inf) CKS(reserve(ctx, ctx->flag_b, n));
  CKS(reserve(ctx, ctx->out, 384));
  dpk = ctx->in_a.p;
  dsg = ctx->in_b.p;
  dpki = pks_inf ? ctx->flag_a.p : nullptr;
  dsgi = sigs_inf ? ctx->flag_b.p : nullptr;
  cudaError_t e = cudaMemcpyAsync(ctx->in_a.p, pks, n * 128, cudaMemcpyHostToDevice, ctx->copy_in);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->in_b.p, sigs, n * 64, cudaMemcpyHostToDevice, ctx->copy_in);
  if (e == cudaSuccess && pks_inf) e = cudaMemcpyAsync(ctx->flag_a.p, pks_inf, n, cudaMemcpyHostToDevice, ctx->copy_in);
  if (e == cudaSuccess && sigs_inf) e = cudaMemcpyAsync(ctx->flag_b.p, sigs_inf, n, cudaMemcpyHostToDevice, ctx->copy_in);
  if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_fork, ctx->copy_in);
  int st_core = e == cudaSuccess ? verify_partial_core(ctx, dpk, dpki, dm, dof, dsg, dsgi, n, dst, dst_len, hash_id,
                                                       weight_seed, first_index, ctx->out.p, nullptr, ctx->ev_fork)
                                 : fail_cuda(ctx, e);
  if (st_core) {
    cudaStreamSynchronize(ctx->copy_in);  // the caller's buffers must not be in flight when we return
    cudaStreamSynchronize(ctx->stream);
    return st_core;
  }
  CK(cudaMemcpyAsync(f_out, ctx->out.p, 384, cudaMemcpyDeviceToHost, ctx->stream));
  return check_hash_fail(ctx);
}

int sylow_b200_verify_batch_finish(sylow_b200_ctx* ctx, const uint8_t* partials, size_t n_partials, int* ok) {
  ENTER(ctx);
  if (!ok || (n_partials && !partials)) return SYLOW_B200_ERR_ARG;
  uint8_t prod[384], gt[384];
  CKS(sylow_b200_fp12_product(ctx, partials, n_partials, prod));
  CKS(sylow_b200_final_exp_batch(ctx, prod, 1, gt));
  *ok = memcmp(gt, kOneCanonical, 384) == 0;
  return 0;
}

int sylow_b200_verify_batch(sylow_b200_ctx* ctx, const uint8_t* pks, const uint8_t* pks_inf, const uint8_t* msgs,
                            const uint64_t* offsets, const uint8_t* sigs, const uint8_t* sigs_inf, size_t n,
                            const uint8_t* dst, size_t dst_len, int hash_id, const uint8_t* weight_seed, int* ok) {
  if (!ok) return SYLOW_B200_ERR_ARG;
  uint8_t f[384];
  CKS(sylow_b200_verify_batch_partial(ctx, pks, pks_inf, msgs, offsets, sigs, sigs_inf, n, dst, dst_len, hash_id,
                                      weight_seed, 0, f));
  return sylow_b200_verify_batch_finish(ctx, f, 1, ok);
}

int sylow_b200_verify_each(sylow_b200_ctx* ctx, const uint8_t* pks, const uint8_t* pks_inf, const uint8_t* msgs,
                           const uint64_t* offsets, const uint8_t* sigs, const uint8_t* sigs_inf, size_t n,
                           const uint8_t* dst, size_t dst_len, int hash_id, uint8_t* ok_out) {
  if (is_multi(ctx) && offsets && pks && sigs && ok_out && n)
    return multi_slices(ctx, n, [=](sylow_b200_ctx* c, size_t, size_t off, size_t m) {
      MsgSlice ms(msgs, offsets, off, m);
      return sylow_b200_verify_each(c, pks + off * 128, adv(pks_inf, off), ms.msgs, ms.offs.data(), sigs + off * 64,
                                    adv(sigs_inf, off), m, dst, dst_len, hash_id, ok_out + off);
    });
  ENTER(ctx);
  DstPrime dp;
  CKS(make_dst_prime(ctx, dst, dst_len, hash_id, dp));
  if (!n) return 0;
  if (!pks || !sigs || !ok_out) return SYLOW_B200_ERR_ARG;
  const uint8_t *dm, *dpk, *dsg, *dpki, *dsgi;
  const uint64_t* dof;
  CKS(msgs_to_dev(ctx, msgs, offsets, n, &dm, &dof));
  CKS(to_dev(ctx, ctx->in_a, pks, n * 128, &dpk));
  CKS(to_dev(ctx, ctx->in_b, sigs, n * 64, &dsg));
  CKS(to_dev(ctx, ctx->flag_a, pks_inf, n, &dpki));
  CKS(to_dev(ctx, ctx->flag_b, sigs_inf, n, &dsgi));
  CKS(verify_miller_values(ctx, dpk, dpki, dm, dof, dsg, dsgi, n, dp, ctx->stream));
  CKS(reserve(ctx, ctx->out, n));
  CKS(launch_check_products(ctx, ctx->scratch0.p, 1, n, ctx->out.p, ctx->stream));
  CK(cudaMemcpyAsync(ok_out, ctx->out.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  return check_hash_fail(ctx);
}

// ------------------------------------------------------------------------------- precomputed G2
int sylow_b200_g2_precompute(sylow_b200_ctx* ctx, const uint8_t* g2, size_t n, uint8_t* coeffs_out) {
  ENTER(ctx);
  if (n && (!g2 || !coeffs_out)) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  const uint8_t* d2;
  CKS(to_dev(ctx, ctx->in_b, g2, n * 128, &d2));
  CKS(reserve(ctx, ctx->out, n * SY_TABLE_BYTES));
  k_g2_precompute<<<nblocks(n, SY_GLUED_THREADS), SY_GLUED_THREADS, 0, ctx->stream>>>(d2, n, ctx->out.p, 0);
  LAUNCHED(ctx);
  CK(cudaMemcpyAsync(coeffs_out, ctx->out.p, n * SY_TABLE_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

// uploads k canonical tables and converts them to Montgomery form in ctx->tables
static int tables_to_dev(sylow_b200_ctx* ctx, const uint8_t* coeffs, size_t k) {
  const uint8_t* dc;
  CKS(to_dev(ctx, ctx->in_d, coeffs, k * SY_TABLE_BYTES, &dc));
  CKS(reserve(ctx, ctx->tables, k * SY_TABLE_BYTES));
  size_t nfp = k * SY_TABLE_BYTES / 32;
  k_fp_convert<<<nblocks(nfp, 128), 128, 0, ctx->stream>>>(dc, nfp, ctx->tables.p, 1);
  LAUNCHED(ctx);
  return 0;
}

int sylow_b200_miller_loop_precomputed(sylow_b200_ctx* ctx, const uint8_t* coeffs, const uint8_t* g1,
                                       const uint8_t* g1_inf, size_t n, uint8_t* f_out) {
  ENTER(ctx);
  if (!coeffs || (n && (!g1 || !f_out))) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  const uint8_t *d1, *d1i;
  CKS(tables_to_dev(ctx, coeffs, 1));
  CKS(to_dev(ctx, ctx->in_a, g1, n * 64, &d1));
  CKS(to_dev(ctx, ctx->flag_a, g1_inf, n, &d1i));
  CKS(reserve(ctx, ctx->out, n * 384));
  CKS((launch_glued<0, 1>(ctx, nullptr, 0, nullptr, nullptr, nullptr, d1, 1, d1i, ctx->tables.p, n, ctx->out.p,
                          ctx->stream)));
  k_fp12_convert<<<nblocks(n, 128), 128, 0, ctx->stream>>>(ctx->out.p, n, ctx->out.p, 0);
  LAUNCHED(ctx);
  CK(cudaMemcpyAsync(f_out, ctx->out.p, n * 384, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

int sylow_b200_pairing_check_fixed_batch_dev(sylow_b200_ctx* ctx, const uint8_t* d_g1, const uint8_t* d_g1_inf,
                                             const uint8_t* d_g2_var, const uint8_t* d_g2_var_inf, size_t k_var,
                                             const uint8_t* d_tables, size_t k_fixed, size_t n_checks,
                                             uint8_t* d_ok_out, void* stream) {
  ENTER_DEV(ctx);
  if (!ctx || (n_checks && (!d_g1 || !d_ok_out || !d_tables || (k_var && !d_g2_var)))) return SYLOW_B200_ERR_ARG;
  if (!n_checks) return 0;
  cudaStream_t s = pick(ctx, stream);
  size_t k = k_var + k_fixed;
  CKS(reserve(ctx, ctx->scratch0, n_checks * 384));
  const uint8_t* g1f = d_g1 + k_var * 64;
  const uint8_t* g1f_inf = d_g1_inf ? d_g1_inf + k_var : nullptr;
  int st;
  if (k_var == 1 && k_fixed == 1)
    st = launch_glued<1, 1>(ctx, d_g1, k, d_g1_inf, d_g2_var, d_g2_var_inf, g1f, k, g1f_inf, d_tables, n_checks, ctx->scratch0.p, s);
  else if (k_var == 1 && k_fixed == 3)
    st = launch_glued<1, 3>(ctx, d_g1, k, d_g1_inf, d_g2_var, d_g2_var_inf, g1f, k, g1f_inf, d_tables, n_checks, ctx->scratch0.p, s);
  else if (k_var == 0 && k_fixed == 1)
    st = launch_glued<0, 1>(ctx, d_g1, k, d_g1_inf, d_g2_var, d_g2_var_inf, g1f, k, g1f_inf, d_tables, n_checks, ctx->scratch0.p, s);
  else
    return SYLOW_B200_ERR_ARG;  // other shapes: use sylow_b200_pairing_check_batch
  CKS(st);
  return launch_check_products(ctx, ctx->scratch0.p, 1, n_checks, d_ok_out, s);
}

int sylow_b200_tables_to_device(sylow_b200_ctx* ctx, const uint8_t* coeffs, size_t k, uint8_t* d_tables_out, void* stream) {
  ENTER_DEV(ctx);
  if (!ctx || !coeffs || !d_tables_out || !k) return SYLOW_B200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CKS(tables_to_dev(ctx, coeffs, k));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpyAsync(d_tables_out, ctx->tables.p, k * SY_TABLE_BYTES, cudaMemcpyDeviceToDevice, pick(ctx, stream)));
  CK(cudaStreamSynchronize(pick(ctx, stream)));  // ctx->tables is reused by the next call
  return 0;
}

int sylow_b200_pairing_check_fixed_batch(sylow_b200_ctx* ctx, const uint8_t* g1, const uint8_t* g1_inf,
                                         const uint8_t* g2_var, const uint8_t* g2_var_inf, size_t k_var,
                                         const uint8_t* coeffs_fixed, size_t k_fixed, size_t n_checks, uint8_t* ok_out) {
  ENTER(ctx);
  if (n_checks && (!g1 || !ok_out || !coeffs_fixed || (k_var && !g2_var))) return SYLOW_B200_ERR_ARG;
  if (!((k_var == 1 && (k_fixed == 1 || k_fixed == 3)) || (k_var == 0 && k_fixed == 1))) return SYLOW_B200_ERR_ARG;
  if (!n_checks) return 0;
  size_t k = k_var + k_fixed;
  const uint8_t *d1, *d1i, *d2, *d2i;
  CKS(tables_to_dev(ctx, coeffs_fixed, k_fixed));
  CKS(to_dev(ctx, ctx->in_a, g1, n_checks * k * 64, &d1));
  CKS(to_dev(ctx, ctx->flag_a, g1_inf, n_checks * k, &d1i));
  CKS(to_dev(ctx, ctx->in_b, g2_var, n_checks * k_var * 128, &d2));
  CKS(to_dev(ctx, ctx->flag_b, g2_var_inf, n_checks * k_var, &d2i));
  CKS(reserve(ctx, ctx->out, n_checks));
  CKS(sylow_b200_pairing_check_fixed_batch_dev(ctx, d1, d1i, d2, d2i, k_var, ctx->tables.p, k_fixed, n_checks, ctx->out.p,
                                               nullptr));
  CK(cudaMemcpyAsync(ok_out, ctx->out.p, n_checks, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

// ------------------------------------------------------------------------------- validation
int sylow_b200_g1_validate_batch(sylow_b200_ctx* ctx, const uint8_t* g1, const uint8_t* g1_inf, size_t n,
                                 int8_t* status_out) {
  ENTER(ctx);
  if (n && (!g1 || !status_out)) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  const uint8_t *d1, *d1i;
  CKS(to_dev(ctx, ctx->in_a, g1, n * 64, &d1));
  CKS(to_dev(ctx, ctx->flag_a, g1_inf, n, &d1i));
  CKS(reserve(ctx, ctx->out, n));
  k_g1_validate<<<nblocks(n, 128), 128, 0, ctx->stream>>>(d1, d1i, n, reinterpret_cast<int8_t*>(ctx->out.p), 0);
  LAUNCHED(ctx);
  CK(cudaMemcpyAsync(status_out, ctx->out.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

int sylow_b200_g2_validate_batch(sylow_b200_ctx* ctx, const uint8_t* g2, const uint8_t* g2_inf, size_t n,
                                 int8_t* status_out) {
  if (is_multi(ctx) && g2 && status_out)
    return multi_slices(ctx, n, [=](sylow_b200_ctx* c, size_t, size_t off, size_t m) {
      return sylow_b200_g2_validate_batch(c, g2 + off * 128, adv(g2_inf, off), m, status_out + off);
    });
  ENTER(ctx);
  if (n && (!g2 || !status_out)) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  const uint8_t *d2, *d2i;
  CKS(to_dev(ctx, ctx->in_b, g2, n * 128, &d2));
  CKS(to_dev(ctx, ctx->flag_b, g2_inf, n, &d2i));
  CKS(reserve(ctx, ctx->out, n));
  k_g2_validate<<<nblocks(n, SY_MUL_THREADS), SY_MUL_THREADS, 0, ctx->stream>>>(d2, d2i, n,
                                                                               reinterpret_cast<int8_t*>(ctx->out.p), 0);
  LAUNCHED(ctx);
  CK(cudaMemcpyAsync(status_out, ctx->out.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

// ------------------------------------------------------------------------------- big-endian codecs
static int from_be(sylow_b200_ctx* ctx, int g2, const uint8_t* be, size_t n, int mode, int validate, uint8_t* out,
                   uint8_t* inf_out, int8_t* status_out) {
  ENTER(ctx);
  if (n && (!be || !out || !inf_out || !status_out)) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  size_t w = g2 ? 128 : 64;
  const uint8_t* dbe;
  CKS(to_dev(ctx, ctx->in_c, be, n * w, &dbe));
  CKS(reserve(ctx, ctx->in_a, n * w));
  CKS(reserve(ctx, ctx->flag_a, n));
  CKS(reserve(ctx, ctx->out, n));
  int8_t* st = reinterpret_cast<int8_t*>(ctx->out.p);
  if (g2) {
    k_decode_be<4><<<nblocks(n, 128), 128, 0, ctx->stream>>>(dbe, 128, n, mode, ctx->in_a.p, ctx->flag_a.p, st);
    LAUNCHED(ctx);
    if (validate) {
      k_g2_validate<<<nblocks(n, SY_MUL_THREADS), SY_MUL_THREADS, 0, ctx->stream>>>(ctx->in_a.p, ctx->flag_a.p, n, st, 1);
      LAUNCHED(ctx);
    }
  } else {
    k_decode_be<2><<<nblocks(n, 128), 128, 0, ctx->stream>>>(dbe, 64, n, mode, ctx->in_a.p, ctx->flag_a.p, st);
    LAUNCHED(ctx);
    if (validate) {
      k_g1_validate<<<nblocks(n, 128), 128, 0, ctx->stream>>>(ctx->in_a.p, ctx->flag_a.p, n, st, 1);
      LAUNCHED(ctx);
    }
  }
  CK(cudaMemcpyAsync(out, ctx->in_a.p, n * w, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(inf_out, ctx->flag_a.p, n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(status_out, st, n, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}
int sylow_b200_g1_from_be_bytes_batch(sylow_b200_ctx* ctx, const uint8_t* be, size_t n, int eip_mode, uint8_t* g1_out,
                                      uint8_t* inf_out, int8_t* status_out) {
  return from_be(ctx, 0, be, n, eip_mode ? 1 : 0, 1, g1_out, inf_out, status_out);
}
int sylow_b200_g2_from_be_bytes_batch(sylow_b200_ctx* ctx, const uint8_t* be, size_t n, int eip_mode, uint8_t* g2_out,
                                      uint8_t* inf_out, int8_t* status_out) {
  return from_be(ctx, 1, be, n, eip_mode ? 1 : 0, 1, g2_out, inf_out, status_out);
}
static int to_be(sylow_b200_ctx* ctx, int g2, const uint8_t* pts, const uint8_t* inf, size_t n, int scrub, uint8_t* be_out) {
  ENTER(ctx);
  if (n && (!pts || !be_out)) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  size_t w = g2 ? 128 : 64;
  const uint8_t *dp, *di;
  CKS(to_dev(ctx, ctx->in_a, pts, n * w, &dp));
  CKS(to_dev(ctx, ctx->flag_a, inf, n, &di));
  CKS(reserve(ctx, ctx->out, n * w));
  if (g2)
    k_encode_be<4><<<nblocks(n, 128), 128, 0, ctx->stream>>>(dp, di, n, scrub, ctx->out.p);
  else
    k_encode_be<2><<<nblocks(n, 128), 128, 0, ctx->stream>>>(dp, di, n, scrub, ctx->out.p);
  LAUNCHED(ctx);
  CK(cudaMemcpyAsync(be_out, ctx->out.p, n * w, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}
int sylow_b200_g1_to_be_bytes_batch(sylow_b200_ctx* ctx, const uint8_t* g1, const uint8_t* g1_inf, size_t n, int scrubbed,
                                    uint8_t* be_out) {
  return to_be(ctx, 0, g1, g1_inf, n, scrubbed, be_out);
}
int sylow_b200_g2_to_be_bytes_batch(sylow_b200_ctx* ctx, const uint8_t* g2, const uint8_t* g2_inf, size_t n, int scrubbed,
                                    uint8_t* be_out) {
  return to_be(ctx, 1, g2, g2_inf, n, scrubbed, be_out);
}

int sylow_b200_eip197_pairing_check_batch(sylow_b200_ctx* ctx, const uint8_t* input, size_t k, size_t n_checks,
                                          uint8_t* ok_out, int8_t* status_out) {
  ENTER(ctx);
  if (n_checks && (!ok_out || !status_out || (k && !input))) return SYLOW_B200_ERR_ARG;
  if (!n_checks) return 0;
  size_t n = k * n_checks;
  CKS(reserve(ctx, ctx->out, 2 * n_checks + 64));
  uint8_t* d_ok = ctx->out.p;
  int8_t* d_st = reinterpret_cast<int8_t*>(ctx->out.p + ((n_checks + 15) / 16) * 16);
  if (n) {
    const uint8_t* din;
    CKS(to_dev(ctx, ctx->in_c, input, n * 192, &din));
    CKS(reserve(ctx, ctx->in_a, n * 64));
    CKS(reserve(ctx, ctx->in_b, n * 128));
    CKS(reserve(ctx, ctx->flag_a, n));
    CKS(reserve(ctx, ctx->flag_b, n));
    CKS(reserve(ctx, ctx->scratch2, 2 * n + 64));
    int8_t* st1 = reinterpret_cast<int8_t*>(ctx->scratch2.p);
    int8_t* st2 = st1 + n;
    k_decode_be<2><<<nblocks(n, 128), 128, 0, ctx->stream>>>(din, 192, n, 1, ctx->in_a.p, ctx->flag_a.p, st1);
    LAUNCHED(ctx);
    k_decode_be<4><<<nblocks(n, 128), 128, 0, ctx->stream>>>(din + 64, 192, n, 1, ctx->in_b.p, ctx->flag_b.p, st2);
    LAUNCHED(ctx);
    k_g1_validate<<<nblocks(n, 128), 128, 0, ctx->stream>>>(ctx->in_a.p, ctx->flag_a.p, n, st1, 1);
    LAUNCHED(ctx);
    k_g2_validate<<<nblocks(n, SY_MUL_THREADS), SY_MUL_THREADS, 0, ctx->stream>>>(ctx->in_b.p, ctx->flag_b.p, n, st2, 1);
    LAUNCHED(ctx);
    CKS(sylow_b200_pairing_check_batch_dev(ctx, ctx->in_a.p, ctx->flag_a.p, ctx->in_b.p, ctx->flag_b.p, k, n_checks, d_ok,
                                           nullptr));
    k_fold_status<<<nblocks(n_checks, 128), 128, 0, ctx->stream>>>(st1, st2, k, n_checks, d_ok, d_st);
    LAUNCHED(ctx);
  } else {
    CK(cudaMemsetAsync(d_ok, 1, n_checks, ctx->stream));  // empty input: success (reth_bn128.rs:173-175)
    CK(cudaMemsetAsync(d_st, 0, n_checks, ctx->stream));
  }
  CK(cudaMemcpyAsync(ok_out, d_ok, n_checks, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(status_out, d_st, n_checks, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

int sylow_b200_gt_mul_batch(sylow_b200_ctx* ctx, const uint8_t* gt, const uint8_t* scalars, size_t n, uint8_t* out) {
  ENTER(ctx);
  if (n && (!gt || !scalars || !out)) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  const uint8_t *dg, *dk;
  CKS(to_dev(ctx, ctx->in_a, gt, n * 384, &dg));
  CKS(to_dev(ctx, ctx->in_b, scalars, n * 32, &dk));
  CKS(reserve(ctx, ctx->out, n * 384));
  k_gt_pow<<<nblocks(n, SY_FEXP_THREADS), SY_FEXP_THREADS, 0, ctx->stream>>>(dg, dk, n, ctx->out.p);
  LAUNCHED(ctx);
  CK(cudaMemcpyAsync(out, ctx->out.p, n * 384, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

// ------------------------------------------------------------------------------- Expander trait
int sylow_b200_expand_message_batch(sylow_b200_ctx* ctx, const uint8_t* msgs, const uint64_t* offsets, size_t n,
                                    const uint8_t* dst, size_t dst_len, int hash_id, size_t len_in_bytes, uint8_t* out) {
  ENTER(ctx);
  DstPrime dp;
  CKS(make_dst_prime(ctx, dst, dst_len, hash_id, dp));
  // XMD: ell = ceil(len / 32) <= 255; both: len < 2^16 (hasher.rs:211-216, i2osp(len, 2))
  if (len_in_bytes == 0 || len_in_bytes > 65535) return SYLOW_B200_ERR_ARG;
  if (hash_id != SYLOW_B200_HASH_SHAKE128 && (len_in_bytes + 31) / 32 > 255) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  if (!out) return SYLOW_B200_ERR_ARG;
  const uint8_t* dm;
  const uint64_t* dof;
  CKS(msgs_to_dev(ctx, msgs, offsets, n, &dm, &dof));
  CKS(reserve(ctx, ctx->out, n * len_in_bytes));
  k_expand_message<<<nblocks(n, 128), 128, 0, ctx->stream>>>(dm, dof, n, dp, (uint32_t)len_in_bytes, ctx->out.p);
  LAUNCHED(ctx);
  CK(cudaMemcpyAsync(out, ctx->out.p, n * len_in_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

int sylow_b200_hash_to_field_batch(sylow_b200_ctx* ctx, const uint8_t* msgs, const uint64_t* offsets, size_t n,
                                   const uint8_t* dst, size_t dst_len, int hash_id, uint8_t* out) {
  ENTER(ctx);
  DstPrime dp;
  CKS(make_dst_prime(ctx, dst, dst_len, hash_id, dp));
  if (!n) return 0;
  if (!out) return SYLOW_B200_ERR_ARG;
  const uint8_t* dm;
  const uint64_t* dof;
  CKS(msgs_to_dev(ctx, msgs, offsets, n, &dm, &dof));
  CKS(reserve(ctx, ctx->out, n * 64));
  k_hash_to_field<<<nblocks(n, 128), 128, 0, ctx->stream>>>(dm, dof, n, dp, ctx->out.p);
  LAUNCHED(ctx);
  CK(cudaMemcpyAsync(out, ctx->out.p, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

// ------------------------------------------------------------------------------- sums / same-signer batches
int sylow_b200_g1_sum(sylow_b200_ctx* ctx, const uint8_t* pts, const uint8_t* pts_inf, size_t n, uint8_t out[64],
                      uint8_t* out_inf) {
  ENTER(ctx);
  if (!out || !out_inf || (n && !pts)) return SYLOW_B200_ERR_ARG;
  if (!n) {
    memset(out, 0, 64);
    out[32] = 1;  // GroupAffine::zero(): (0, 1, infinity)
    *out_inf = 1;
    return 0;
  }
  const uint8_t *dp, *di;
  CKS(to_dev(ctx, ctx->in_a, pts, n * 64, &dp));
  CKS(to_dev(ctx, ctx->flag_a, pts_inf, n, &di));
  CKS(reserve(ctx, ctx->out, 128));
  CKS(g1_sum_reduce(ctx, dp, di, n, 0, ctx->out.p, ctx->out.p + 64, ctx->stream));
  CK(cudaMemcpyAsync(out, ctx->out.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(out_inf, ctx->out.p + 64, 1, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

#define SY_MSM_BUCKET_MIN 131072
// Bucket MSM on device buffers: d_out = 64-byte affine point, d_out_inf its flag.  window_bits = 0 picks c from n.
static int g1_msm_bucket_dev(sylow_b200_ctx* ctx, const uint8_t* d_pts, const uint8_t* d_inf, const uint8_t* d_scalars,
                             size_t n, int window_bits, uint8_t* d_out, uint8_t* d_out_inf, cudaStream_t s) {
  int c = window_bits;
  if (c <= 0) {
    int lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    c = lg - 10 < 8 ? 8 : lg - 10;  // measured optimum (tools/msm_time.py): 8 up to 2^18 points, 10 at 2^20
  }
  if (c < 4) c = 4;
  if (c > (window_bits > 0 ? 16 : 13)) c = window_bits > 0 ? 16 : 13;  // the automatic choice stops at 13 (64 digits per chunk)
  const int nw = (256 + c - 1) / c;
  const size_t nb = (size_t)nw << c;
  // scratch: counts | offsets | cursors | first item (nb u32 each), item start | length (nw * max_items u32 each), items
  // per window, index lists (nw * n u32), Montgomery points (n * 64), item sums, buckets (nb * 96), chunk partials, result
  const size_t max_items = ((size_t)1 << c) + n / SY_MSM_ITEM + 1;
  const size_t ni = (size_t)nw * max_items;
  size_t o_cnt = 0, o_off = o_cnt + nb * 4, o_cur = o_off + nb * 4, o_ifi = o_cur + nb * 4, o_ist = o_ifi + nb * 4;
  size_t o_iln = o_ist + ni * 4, o_iw = o_iln + ni * 4, o_idx = o_iw + 256;
  size_t o_pts = (o_idx + (size_t)nw * n * 4 + 255) & ~(size_t)255, o_isum = o_pts + n * 64, o_bkt = o_isum + ni * 96;
  size_t o_par = o_bkt + nb * 96, o_res = o_par + (size_t)nw * SY_MSM_CHUNKS * 96, total = o_res + 96;
  CKS(reserve(ctx, ctx->msm, total));
  uint8_t* base = ctx->msm.p;
  uint32_t* cnt = reinterpret_cast<uint32_t*>(base + o_cnt);
  uint32_t* off = reinterpret_cast<uint32_t*>(base + o_off);
  uint32_t* cur = reinterpret_cast<uint32_t*>(base + o_cur);
  uint32_t* ifi = reinterpret_cast<uint32_t*>(base + o_ifi);
  uint32_t* ist = reinterpret_cast<uint32_t*>(base + o_ist);
  uint32_t* iln = reinterpret_cast<uint32_t*>(base + o_iln);
  uint32_t* iw = reinterpret_cast<uint32_t*>(base + o_iw);
  uint32_t* idx = reinterpret_cast<uint32_t*>(base + o_idx);
  CK(cudaMemsetAsync(cnt, 0, nb * 4, s));
  k_msm_count<<<nblocks(n, 256), 256, 0, s>>>(d_scalars, d_inf, n, c, nw, cnt);
  LAUNCHED(ctx);
  k_msm_offsets<<<nblocks(nw, 32), 32, 0, s>>>(cnt, c, nw, max_items, off, cur, ifi, ist, iln, iw);
  LAUNCHED(ctx);
  k_msm_scatter<<<nblocks(n, 256), 256, 0, s>>>(d_scalars, d_inf, n, c, nw, cur, idx);
  LAUNCHED(ctx);
  k_fp_convert<<<nblocks(2 * n, 128), 128, 0, s>>>(d_pts, 2 * n, base + o_pts, 1);
  LAUNCHED(ctx);
  k_msm_item_sum<<<nblocks(ni, SY_MUL_THREADS), SY_MUL_THREADS, 0, s>>>(base + o_pts, n, nw, max_items, ist, iln, iw, idx,
                                                                       base + o_isum);
  LAUNCHED(ctx);
  k_msm_bucket_sum<<<nblocks(nb, SY_MUL_THREADS), SY_MUL_THREADS, 0, s>>>(c, nw, max_items, cnt, ifi, base + o_isum,
                                                                         base + o_bkt);
  LAUNCHED(ctx);
  k_msm_chunk_reduce<<<nblocks((size_t)nw * SY_MSM_CHUNKS, 64), 64, 0, s>>>(base + o_bkt, c, nw, base + o_par);
  LAUNCHED(ctx);
  k_msm_window_sum<<<nblocks(nw, 32), 32, 0, s>>>(base + o_par, nw);
  LAUNCHED(ctx);
  k_msm_window_final<<<1, 1, 0, s>>>(base + o_par, c, nw, base + o_res);
  LAUNCHED(ctx);
  k_g1_finish_sum<<<1, 1, 0, s>>>(base + o_res, 0, d_out, d_out_inf);
  LAUNCHED(ctx);
  return 0;
}

int sylow_b200_g1_msm_bucket(sylow_b200_ctx* ctx, const uint8_t* pts, const uint8_t* pts_inf, const uint8_t* scalars,
                             size_t n, int window_bits, uint8_t out[64], uint8_t* out_inf) {
  ENTER(ctx);
  if (!out || !out_inf || (n && (!pts || !scalars)) || window_bits < 0 || window_bits > 16) return SYLOW_B200_ERR_ARG;
  if (!n) return sylow_b200_g1_sum(ctx, nullptr, nullptr, 0, out, out_inf);
  if (n >= ((size_t)1 << 31)) return SYLOW_B200_ERR_ARG;
  const uint8_t *dp, *di, *dk;
  CKS(to_dev(ctx, ctx->in_a, pts, n * 64, &dp));
  CKS(to_dev(ctx, ctx->flag_a, pts_inf, n, &di));
  CKS(to_dev(ctx, ctx->in_b, scalars, n * 32, &dk));
  CKS(reserve(ctx, ctx->out, 128));
  CKS(g1_msm_bucket_dev(ctx, dp, di, dk, n, window_bits, ctx->out.p, ctx->out.p + 64, ctx->stream));
  CK(cudaMemcpyAsync(out, ctx->out.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(out_inf, ctx->out.p + 64, 1, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

int sylow_b200_g1_msm(sylow_b200_ctx* ctx, const uint8_t* pts, const uint8_t* pts_inf, const uint8_t* scalars, size_t n,
                      uint8_t out[64], uint8_t* out_inf) {
  ENTER(ctx);
  if (!out || !out_inf || (n && (!pts || !scalars))) return SYLOW_B200_ERR_ARG;
  if (!n) return sylow_b200_g1_sum(ctx, nullptr, nullptr, 0, out, out_inf);
  // from 2^17 points on the bucket method beats n GLV ladders + tree sum (tools/msm_time.py)
  if (n >= SY_MSM_BUCKET_MIN && n < ((size_t)1 << 31))
    return sylow_b200_g1_msm_bucket(ctx, pts, pts_inf, scalars, n, 0, out, out_inf);
  const uint8_t *dp, *di, *dk;
  CKS(to_dev(ctx, ctx->in_a, pts, n * 64, &dp));
  CKS(to_dev(ctx, ctx->flag_a, pts_inf, n, &di));
  CKS(to_dev(ctx, ctx->in_b, scalars, n * 32, &dk));
  CKS(reserve(ctx, ctx->scratch2, n * 65));
  CKS(reserve(ctx, ctx->out, 128));
  CKS(sylow_b200_g1_mul_batch_dev(ctx, dp, di, dk, n, ctx->scratch2.p, ctx->scratch2.p + n * 64, nullptr));
  CKS(g1_sum_reduce(ctx, ctx->scratch2.p, ctx->scratch2.p + n * 64, n, 0, ctx->out.p, ctx->out.p + 64, ctx->stream));
  CK(cudaMemcpyAsync(out, ctx->out.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(out_inf, ctx->out.p + 64, 1, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

int sylow_b200_lagrange_coefficients_batch(sylow_b200_ctx* ctx, const uint64_t* ids, size_t n_sets, size_t t,
                                           uint8_t* out) {
  ENTER(ctx);
  if ((n_sets && t) && (!ids || !out)) return SYLOW_B200_ERR_ARG;
  size_t n = n_sets * t;
  if (!n) return 0;
  const uint8_t* d_ids;
  CKS(to_dev(ctx, ctx->in_a, reinterpret_cast<const uint8_t*>(ids), n * 8, &d_ids));
  CKS(reserve(ctx, ctx->out, n * 32));
  k_lagrange<<<nblocks(n, 128), 128, 0, ctx->stream>>>(reinterpret_cast<const uint64_t*>(d_ids), n_sets, t, ctx->out.p);
  LAUNCHED(ctx);
  CK(cudaMemcpyAsync(out, ctx->out.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

int sylow_b200_threshold_aggregate_batch(sylow_b200_ctx* ctx, const uint64_t* ids, const uint8_t* sigs,
                                         const uint8_t* sigs_inf, size_t n_sets, size_t t, uint8_t* out,
                                         uint8_t* out_inf) {
  ENTER(ctx);
  if (n_sets && (!out || !out_inf || (t && (!ids || !sigs)))) return SYLOW_B200_ERR_ARG;
  if (!n_sets) return 0;
  if (!t) {  // the empty sum: GroupProjective::default() = infinity, encoded (0, 1) + flag
    for (size_t s = 0; s < n_sets; s++) {
      memset(out + s * 64, 0, 64);
      out[s * 64 + 32] = 1;
      out_inf[s] = 1;
    }
    return 0;
  }
  size_t n = n_sets * t;
  const uint8_t *d_ids, *d_sigs, *d_inf;
  CKS(to_dev(ctx, ctx->in_b, reinterpret_cast<const uint8_t*>(ids), n * 8, &d_ids));
  CKS(to_dev(ctx, ctx->in_a, sigs, n * 64, &d_sigs));
  CKS(to_dev(ctx, ctx->flag_a, sigs_inf, n, &d_inf));
  CKS(reserve(ctx, ctx->scratch2, n * 96));
  CKS(reserve(ctx, ctx->out, n_sets * 65));
  k_lagrange_mul<<<nblocks(n, SY_MUL_THREADS), SY_MUL_THREADS, 0, ctx->stream>>>(
      reinterpret_cast<const uint64_t*>(d_ids), d_sigs, d_inf, n_sets, t, ctx->scratch2.p);
  LAUNCHED(ctx);
  k_g1_segment_sum<<<nblocks(n_sets, 4), 128, 0, ctx->stream>>>(ctx->scratch2.p, n_sets, t, ctx->out.p,
                                                               ctx->out.p + n_sets * 64);
  LAUNCHED(ctx);
  CK(cudaMemcpyAsync(out, ctx->out.p, n_sets * 64, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(out_inf, ctx->out.p + n_sets * 64, n_sets, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

int sylow_b200_verify_batch_same_signer(sylow_b200_ctx* ctx, const uint8_t* pk, int pk_inf, const uint8_t* msgs,
                                        const uint64_t* offsets, const uint8_t* sigs, const uint8_t* sigs_inf, size_t n,
                                        const uint8_t* dst, size_t dst_len, int hash_id, const uint8_t* weight_seed,
                                        int* ok) {
  ENTER(ctx);
  if (!ok) return SYLOW_B200_ERR_ARG;
  DstPrime dp;
  CKS(make_dst_prime(ctx, dst, dst_len, hash_id, dp));
  if (!n) {
    *ok = 1;
    return 0;
  }
  if (!pk || !sigs) return SYLOW_B200_ERR_ARG;
  WeightSeed ws;
  const bool weighted = load_seed(weight_seed, ws);
  const uint8_t *dm, *dpk, *dsg, *dsgi;
  const uint64_t* dof;
  CKS(msgs_to_dev(ctx, msgs, offsets, n, &dm, &dof));
  CKS(to_dev(ctx, ctx->in_a, pk, 128, &dpk));
  CKS(to_dev(ctx, ctx->in_b, sigs, n * 64, &dsg));
  CKS(to_dev(ctx, ctx->flag_b, sigs_inf, n, &dsgi));
  // e(sum r_i sig_i, G2gen) * e(-sum r_i H(m_i), pk) == 1 (r_i = 1 without a seed): two Miller loops for the whole batch
  CKS(reserve(ctx, ctx->scratch2, n * 65 + 64));
  CKS(reserve(ctx, ctx->out, 1024));
  uint8_t* d_pairs = ctx->out.p;        // two G1 points (128 B) + two flags at +128 + the G2 pair at +256
  uint8_t* d_flags = ctx->out.p + 128;
  uint8_t* d_g2 = ctx->out.p + 256;     // G2gen || pk
  uint8_t* d_g2_inf = ctx->out.p + 520; // 0, pk_inf
  uint8_t* d_ok = ctx->out.p + 512;
  uint8_t* d_hm = ctx->scratch2.p;
  cudaStream_t st_ = ctx->stream;
  CKS(hash_launch(ctx, dm, dof, n, dp, 0, d_hm, d_hm + n * 64, st_));
  if (weighted) {
    CKS(reserve(ctx, ctx->proj, n * 96));
    k_g1_mul_weight<<<nblocks(n, SY_MUL_THREADS), SY_MUL_THREADS, 0, st_>>>(dsg, dsgi, ws, 0, n, ctx->proj.p);
    LAUNCHED(ctx);
    CKS(g1_sum_reduce(ctx, ctx->proj.p, nullptr, n, 0, d_pairs, d_flags, st_, 0));
    k_g1_mul_weight<<<nblocks(n, SY_MUL_THREADS), SY_MUL_THREADS, 0, st_>>>(d_hm, d_hm + n * 64, ws, 0, n, ctx->proj.p);
    LAUNCHED(ctx);
    CKS(g1_sum_reduce(ctx, ctx->proj.p, nullptr, n, 1, d_pairs + 64, d_flags + 1, st_, 0));
  } else {
    CKS(g1_sum_reduce(ctx, dsg, dsgi, n, 0, d_pairs, d_flags, st_));
    CKS(g1_sum_reduce(ctx, d_hm, d_hm + n * 64, n, 1, d_pairs + 64, d_flags + 1, st_));
  }
  const uint8_t h_g2_inf[2] = {0, (uint8_t)(pk_inf ? 1 : 0)};
  CK(cudaMemcpyAsync(d_g2, kG2GenWords, 128, cudaMemcpyHostToDevice, st_));
  CK(cudaMemcpyAsync(d_g2 + 128, dpk, 128, cudaMemcpyDeviceToDevice, st_));
  CK(cudaMemcpyAsync(d_g2_inf, h_g2_inf, 2, cudaMemcpyHostToDevice, st_));
  CKS(sylow_b200_pairing_check_batch_dev(ctx, d_pairs, d_flags, d_g2, d_g2_inf, 2, 1, d_ok, nullptr));
  uint8_t h_ok = 0;
  CK(cudaMemcpyAsync(&h_ok, d_ok, 1, cudaMemcpyDeviceToHost, st_));
  int st = check_hash_fail(ctx);
  *ok = h_ok;
  return st;
}

// ------------------------------------------------------------------------------- diagnostics
int sylow_b200_batch_weights(sylow_b200_ctx* ctx, const uint8_t* weight_seed, uint64_t first_index, size_t n,
                             uint64_t* out) {
  ENTER(ctx);
  if (!weight_seed || (n && !out)) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  WeightSeed ws;
  load_seed(weight_seed, ws);
  CKS(reserve(ctx, ctx->out, n * 8));
  k_batch_weights<<<nblocks(n, 128), 128, 0, ctx->stream>>>(ws, first_index, n, reinterpret_cast<uint64_t*>(ctx->out.p));
  LAUNCHED(ctx);
  CK(cudaMemcpyAsync(out, ctx->out.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

int sylow_b200_fp_op_batch(sylow_b200_ctx* ctx, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
  ENTER(ctx);
  if (n && (!a || !b || !out)) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  const uint8_t *da, *db;
  CKS(to_dev(ctx, ctx->in_a, a, n * 32, &da));
  CKS(to_dev(ctx, ctx->in_b, b, n * 32, &db));
  CKS(reserve(ctx, ctx->out, n * 32));
  k_fp_op<<<nblocks(n, 128), 128, 0, ctx->stream>>>(op, da, db, n, ctx->out.p);
  LAUNCHED(ctx);
  CK(cudaMemcpyAsync(out, ctx->out.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}

int sylow_b200_fp12_op_batch(sylow_b200_ctx* ctx, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
  ENTER(ctx);
  if (n && (!a || !b || !out)) return SYLOW_B200_ERR_ARG;
  if (!n) return 0;
  const uint8_t *da, *db;
  CKS(to_dev(ctx, ctx->in_a, a, n * 384, &da));
  CKS(to_dev(ctx, ctx->in_b, b, n * 384, &db));
  CKS(reserve(ctx, ctx->out, n * 384));
  k_fp12_op<<<nblocks(n, SY_SMALL_THREADS), SY_SMALL_THREADS, 0, ctx->stream>>>(op, da, db, n, ctx->out.p);
  LAUNCHED(ctx);
  CK(cudaMemcpyAsync(out, ctx->out.p, n * 384, cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx);
}


int sylow_b200_imad_probe(sylow_b200_ctx* ctx, int variant, int blocks, int threads, int iters, float* ms_out,
                          double* ops_out) {
  ENTER(ctx);
  if (!ms_out || !ops_out || blocks <= 0 || threads <= 0 || threads > 1024 || iters <= 0) return SYLOW_B200_ERR_ARG;
  CKS(reserve(ctx, ctx->out, 256));
  CK(cudaMemsetAsync(ctx->out.p, 0x5a, 256, ctx->stream));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const uint32_t* src = reinterpret_cast<const uint32_t*>(ctx->out.p);
  uint32_t* sink = reinterpret_cast<uint32_t*>(ctx->out.p) + 32;
  double per_thread_iter = 0;
  for (int rep = 0; rep < 2; rep++) {  // the first pass warms up
    CK(cudaEventRecord(e0, ctx->stream));
    switch (variant) {
      case 0: k_imad_probe<1><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 1; break;
      case 1: k_imad_probe<2><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 2; break;
      case 2: k_imad_probe<4><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 4; break;
      case 10: k_pipe_probe<0><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 64; break;
      case 11: k_pipe_probe<1><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 64; break;
      case 12: k_pipe_probe<2><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 64; break;
      case 13: k_dfma_probe<<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 64; break;
      case 20: k_ifetch_probe<4><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 4; break;
      case 21: k_ifetch_probe<16><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 16; break;
      case 22: k_ifetch_probe<64><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 64; break;
      case 23: k_ifetch_probe<256><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 256; break;
      case 30: k_tower_probe<0><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 3; break;
      case 31: k_tower_probe<1><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 2; break;
      case 32: k_tower_probe<2><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 18; break;
      case 33: k_tower_probe<3><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 54; break;
      case 34: k_tower_probe<4><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 36; break;
      case 35: k_tower_probe<5><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 39; break;
      case 36: k_tower_probe<6><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 18; break;
      case 37: k_tower_probe<7><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 26; break;
      case 38: k_tower_probe<8><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 4; break;
      case 39: k_tower_probe<9><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 2; break;
      case 40: k_overlap_probe<0><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 1; break;
      case 41: k_overlap_probe<1><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 1; break;
      case 42: k_overlap_probe<2><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 1; break;
      case 43: k_overlap_probe<3><<<blocks, threads, 0, ctx->stream>>>(iters, src, sink); per_thread_iter = 1; break;
      default: return SYLOW_B200_ERR_ARG;
    }
    LAUNCHED(ctx);
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaEventSynchronize(e1));
  }
  CK(cudaEventElapsedTime(ms_out, e0, e1));
  *ops_out = (double)blocks * threads * (double)iters * per_thread_iter;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}

}  // extern "C"
