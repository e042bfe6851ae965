//! `src/batch.rs` of a sylow fork: the batched entry points BASELINE.json's north_star asks for
//! (`pairing_batch`, `verify_batch`, `g1_mul_batch`, ...), implemented on top of `sylow-cuda-sys`.
//!
//! It must live INSIDE sylow because `Gt(pub(crate) Fp12)`, `MillerLoopResult(pub(crate) Fp12)` and the
//! fields of `GroupAffine` are crate-private (src/groups/gt.rs:115, src/pairing.rs:72,
//! src/groups/group.rs:174-181).  `lib.rs` adds:
//!
//! ```ignore
//! mod batch;
//! pub use crate::batch::{final_exponentiation_batch, g1_mul_batch, g2_mul_batch, glued_miller_loop_batch,
//!                        glued_pairing_batch, pairing_batch, pairing_check_batch, sign_batch,
//!                        threshold_aggregate_batch, verify_batch, verify_each, BatchError, Engine};
//! ```
//!
//! All slicing over GPUs, the per-GPU host threads and the combination of the 384-byte Miller partial products live
//! behind the C ABI (`sylow_b200_create_multi`): every function here is ONE call on ONE context, so this file, the
//! C++ header and the Python binding cannot drift apart.
//!
//! NOT COMPILED in the build image (no Rust toolchain).  `tests/test_rust_boundary.py` checks statically that
//! every name in the `pub use` list above is defined here, that every `sys::` symbol exists in
//! `sylow-cuda-sys/src/lib.rs` and that every call passes as many arguments as the declaration takes; the C ABI
//! underneath is what the parity tests exercise (Python ctypes + include/sylow_b200.hpp).
use crate::fields::fp::Fp;
use crate::fields::fp12::Fp12;
use crate::fields::fp2::Fp2;
use crate::fields::fp6::Fp6;
use crate::groups::g1::{G1Affine, G1Projective};
use crate::groups::g2::{G2Affine, G2Projective};
use crate::groups::group::GroupError;
use crate::groups::gt::Gt;
use crate::pairing::MillerLoopResult;
use crypto_bigint::U256;
use subtle::Choice;
use sylow_cuda_sys as sys;

/// Errors of the batched entry points: sylow's own `GroupError` where the library reports one, and the failures a
/// CPU library cannot have.
#[derive(Debug)]
pub enum BatchError {
    /// NotOnCurve / NotInSubgroup / CannotHashToGroup / DecodeError, as the reference's typed API would raise them
    Group(GroupError),
    /// mismatched batch lengths, unsupported shape
    InvalidArgument,
    /// device or host allocation failed
    OutOfMemory,
    /// a CUDA runtime call failed; `code` is the `cudaError_t` (`sylow_b200_last_cuda_error`)
    Cuda { code: i32 },
}

impl From<GroupError> for BatchError {
    fn from(e: GroupError) -> Self {
        BatchError::Group(e)
    }
}

/// One context over the listed GPUs.  `Engine::new(&[0, 1, .., 7])` for an 8xB200 box; a batch is cut into
/// contiguous slices inside the library and results are bit-identical to a single GPU's.
pub struct Engine {
    ctx: *mut sys::SylowB200Ctx,
}

// A context is used by one call at a time; `&mut self` on every entry point enforces it.
unsafe impl Send for Engine {}

impl Engine {
    pub fn new(devices: &[i32]) -> Result<Self, BatchError> {
        let mut ctx = core::ptr::null_mut();
        let st = unsafe { sys::sylow_b200_create_multi(&mut ctx, devices.as_ptr(), devices.len() as i32) };
        if st != sys::SYLOW_B200_OK {
            return Err(map_status(st, 0));
        }
        Ok(Engine { ctx })
    }
    pub fn device_count(&self) -> usize {
        unsafe { sys::sylow_b200_device_count(self.ctx) as usize }
    }
    fn check(&self, st: i32) -> Result<(), BatchError> {
        if st == sys::SYLOW_B200_OK {
            return Ok(());
        }
        Err(map_status(st, unsafe { sys::sylow_b200_last_cuda_error(self.ctx) }))
    }
}

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { sys::sylow_b200_destroy(self.ctx) };
    }
}

fn map_status(st: i32, cuda: i32) -> BatchError {
    match st {
        sys::SYLOW_B200_ERR_NOT_ON_CURVE => BatchError::Group(GroupError::NotOnCurve),
        sys::SYLOW_B200_ERR_NOT_IN_SUBGROUP => BatchError::Group(GroupError::NotInSubgroup),
        sys::SYLOW_B200_ERR_CANNOT_HASH => BatchError::Group(GroupError::CannotHashToGroup),
        sys::SYLOW_B200_ERR_DECODE => BatchError::Group(GroupError::DecodeError),
        sys::SYLOW_B200_ERR_NOMEM => BatchError::OutOfMemory,
        sys::SYLOW_B200_ERR_CUDA => BatchError::Cuda { code: cuda },
        _ => BatchError::InvalidArgument,
    }
}

fn same_len(a: usize, b: usize) -> Result<usize, BatchError> {
    if a == b {
        Ok(a)
    } else {
        Err(BatchError::InvalidArgument)
    }
}

// ---- marshalling: Fp <-> 32 little-endian bytes of `value().to_words()` (src/fields/fp.rs:232-234) ----
fn put_fp(out: &mut Vec<u8>, x: &Fp) {
    for w in x.value().to_words() {
        out.extend_from_slice(&w.to_le_bytes());
    }
}
fn get_fp(b: &[u8]) -> Fp {
    let mut w = [0u64; 4];
    for (i, c) in b.chunks_exact(8).enumerate() {
        w[i] = u64::from_le_bytes(c.try_into().unwrap());
    }
    Fp::new(U256::from_words(w))
}
fn get_fp2(b: &[u8]) -> Fp2 {
    Fp2::new(&[get_fp(&b[0..32]), get_fp(&b[32..64])])
}
fn get_fp12(b: &[u8]) -> Fp12 {
    let c: Vec<Fp2> = b.chunks_exact(64).map(get_fp2).collect();
    Fp12::new(&[Fp6::new(&[c[0], c[1], c[2]]), Fp6::new(&[c[3], c[4], c[5]])])
}
fn put_fp12(out: &mut Vec<u8>, f: &Fp12) {
    // tower order c0.c0.c0, c0.c0.c1, ..., c1.c2.c1 (src/fields/fp12.rs:561-574)
    for c6 in f.0.iter() {
        for c2 in c6.0.iter() {
            for c in c2.0.iter() {
                put_fp(out, c);
            }
        }
    }
}
fn get_g1(b: &[u8], inf: u8) -> G1Affine {
    G1Affine { x: get_fp(&b[..32]), y: get_fp(&b[32..64]), infinity: Choice::from(inf) }
}
fn get_g2(b: &[u8], inf: u8) -> G2Affine {
    G2Affine { x: get_fp2(&b[..64]), y: get_fp2(&b[64..128]), infinity: Choice::from(inf) }
}
fn marshal_g1(p: &[G1Projective]) -> (Vec<u8>, Vec<u8>) {
    let (mut v, mut inf) = (Vec::with_capacity(p.len() * 64), Vec::with_capacity(p.len()));
    for q in p {
        let a = G1Affine::from(q); // src/groups/group.rs:475-495
        put_fp(&mut v, &a.x);
        put_fp(&mut v, &a.y);
        inf.push(a.is_zero() as u8);
    }
    (v, inf)
}
fn marshal_g2(p: &[G2Projective]) -> (Vec<u8>, Vec<u8>) {
    let (mut v, mut inf) = (Vec::with_capacity(p.len() * 128), Vec::with_capacity(p.len()));
    for q in p {
        let a = G2Affine::from(q);
        for c in [a.x.0[0], a.x.0[1], a.y.0[0], a.y.0[1]] {
            put_fp(&mut v, &c);
        }
        inf.push(a.is_zero() as u8);
    }
    (v, inf)
}
fn marshal_scalars(k: &[Fp]) -> Vec<u8> {
    let mut v = Vec::with_capacity(k.len() * 32);
    k.iter().for_each(|x| put_fp(&mut v, x));
    v
}
fn pack_msgs(msgs: &[&[u8]]) -> (Vec<u8>, Vec<u64>) {
    let mut offs = Vec::with_capacity(msgs.len() + 1);
    let mut buf = Vec::new();
    offs.push(0u64);
    for m in msgs {
        buf.extend_from_slice(m);
        offs.push(buf.len() as u64);
    }
    if buf.is_empty() {
        buf.push(0); // a valid pointer for an all-empty batch
    }
    (buf, offs)
}

/// `p.iter().zip(q).map(|(p, q)| pairing(p, q))` on the GPUs (src/pairing.rs:870-893).
pub fn pairing_batch(e: &mut Engine, p: &[G1Projective], q: &[G2Projective]) -> Result<Vec<Gt>, BatchError> {
    let n = same_len(p.len(), q.len())?;
    let (g1, g1i) = marshal_g1(p);
    let (g2, g2i) = marshal_g2(q);
    let mut out = vec![0u8; n * 384];
    e.check(unsafe {
        sys::sylow_b200_pairing_batch(e.ctx, g1.as_ptr(), g1i.as_ptr(), g2.as_ptr(), g2i.as_ptr(), n, out.as_mut_ptr())
    })?;
    Ok(out.chunks_exact(384).map(|b| Gt(get_fp12(b))).collect())
}

/// `glued_miller_loop` over the batch (src/pairing.rs:970-1022): the product of the per-pair Miller values, equal
/// bit for bit to the shared-squaring loop.  Infinite pairs contribute 1.
pub fn glued_miller_loop_batch(e: &mut Engine, p: &[G1Projective], q: &[G2Projective]) -> Result<MillerLoopResult, BatchError> {
    let n = p.len().min(q.len()); // zip truncation, like the reference (:975)
    let (g1, g1i) = marshal_g1(&p[..n]);
    let (g2, g2i) = marshal_g2(&q[..n]);
    let mut out = [0u8; 384];
    e.check(unsafe {
        sys::sylow_b200_miller_product(e.ctx, g1.as_ptr(), g1i.as_ptr(), g2.as_ptr(), g2i.as_ptr(), n, out.as_mut_ptr())
    })?;
    Ok(MillerLoopResult(get_fp12(&out)))
}

/// `MillerLoopResult::final_exponentiation` (src/pairing.rs:245-492) over a batch.
pub fn final_exponentiation_batch(e: &mut Engine, f: &[MillerLoopResult]) -> Result<Vec<Gt>, BatchError> {
    let n = f.len();
    let mut inp = Vec::with_capacity(n * 384);
    f.iter().for_each(|x| put_fp12(&mut inp, &x.0));
    let mut out = vec![0u8; n * 384];
    e.check(unsafe { sys::sylow_b200_final_exp_batch(e.ctx, inp.as_ptr(), n, out.as_mut_ptr()) })?;
    Ok(out.chunks_exact(384).map(|b| Gt(get_fp12(b))).collect())
}

/// `glued_pairing(g1s, g2s)` (src/pairing.rs:1029-1037): one Miller product, one final exponentiation.
pub fn glued_pairing_batch(e: &mut Engine, p: &[G1Projective], q: &[G2Projective]) -> Result<Gt, BatchError> {
    let f = glued_miller_loop_batch(e, p, q)?;
    Ok(final_exponentiation_batch(e, &[f])?.remove(0))
}

/// `p.len() / pairs_per_check` independent checks `glued_pairing(..) == Gt::identity()` (the ecPairing / Groth16
/// shape, examples/reth_bn128.rs:211-214).
pub fn pairing_check_batch(e: &mut Engine, p: &[G1Projective], q: &[G2Projective], pairs_per_check: usize) -> Result<Vec<bool>, BatchError> {
    let n = same_len(p.len(), q.len())?;
    if pairs_per_check == 0 || n % pairs_per_check != 0 {
        return Err(BatchError::InvalidArgument);
    }
    let (g1, g1i) = marshal_g1(p);
    let (g2, g2i) = marshal_g2(q);
    let mut ok = vec![0u8; n / pairs_per_check];
    e.check(unsafe {
        sys::sylow_b200_pairing_check_batch(e.ctx, g1.as_ptr(), g1i.as_ptr(), g2.as_ptr(), g2i.as_ptr(), pairs_per_check,
                                            n / pairs_per_check, ok.as_mut_ptr())
    })?;
    Ok(ok.into_iter().map(|b| b != 0).collect())
}

/// `scalars[i] * pts[i]` (src/groups/group.rs:639-667), returned affine.
pub fn g1_mul_batch(e: &mut Engine, pts: &[G1Projective], scalars: &[Fp]) -> Result<Vec<G1Affine>, BatchError> {
    let n = same_len(pts.len(), scalars.len())?;
    let (g1, g1i) = marshal_g1(pts);
    let ks = marshal_scalars(scalars);
    let (mut out, mut inf) = (vec![0u8; n * 64], vec![0u8; n]);
    e.check(unsafe {
        sys::sylow_b200_g1_mul_batch(e.ctx, g1.as_ptr(), g1i.as_ptr(), ks.as_ptr(), n, out.as_mut_ptr(), inf.as_mut_ptr())
    })?;
    Ok(out.chunks_exact(64).zip(&inf).map(|(b, &i)| get_g1(b, i)).collect())
}

/// The same on the twist.
pub fn g2_mul_batch(e: &mut Engine, pts: &[G2Projective], scalars: &[Fp]) -> Result<Vec<G2Affine>, BatchError> {
    let n = same_len(pts.len(), scalars.len())?;
    let (g2, g2i) = marshal_g2(pts);
    let ks = marshal_scalars(scalars);
    let (mut out, mut inf) = (vec![0u8; n * 128], vec![0u8; n]);
    e.check(unsafe {
        sys::sylow_b200_g2_mul_batch(e.ctx, g2.as_ptr(), g2i.as_ptr(), ks.as_ptr(), n, out.as_mut_ptr(), inf.as_mut_ptr())
    })?;
    Ok(out.chunks_exact(128).zip(&inf).map(|(b, &i)| get_g2(b, i)).collect())
}

/// `sign(&sk[i], msgs[i])` (src/lib.rs:179-187): hash-to-curve and one scalar multiplication per message.
pub fn sign_batch(e: &mut Engine, sks: &[Fp], msgs: &[&[u8]]) -> Result<Vec<G1Affine>, BatchError> {
    let n = same_len(sks.len(), msgs.len())?;
    let ks = marshal_scalars(sks);
    let (buf, offs) = pack_msgs(msgs);
    let (mut out, mut inf) = (vec![0u8; n * 64], vec![0u8; n]);
    e.check(unsafe {
        sys::sylow_b200_sign_batch(e.ctx, ks.as_ptr(), buf.as_ptr(), offs.as_ptr(), n, crate::DST.as_ptr(), crate::DST.len(),
                                   sys::SYLOW_B200_HASH_KECCAK256, out.as_mut_ptr(), inf.as_mut_ptr())
    })?;
    Ok(out.chunks_exact(64).zip(&inf).map(|(b, &i)| get_g1(b, i)).collect())
}

/// `verify(&pks[i], msgs[i], &sigs[i])` per signature (src/lib.rs:223-236), identity keys and signatures included
/// (the library applies `pairing`'s infinity rule, src/pairing.rs:876-886).
pub fn verify_each(e: &mut Engine, pks: &[G2Projective], msgs: &[&[u8]], sigs: &[G1Projective]) -> Result<Vec<bool>, BatchError> {
    let n = same_len(pks.len(), same_len(msgs.len(), sigs.len())?)?;
    let (pk, pki) = marshal_g2(pks);
    let (sg, sgi) = marshal_g1(sigs);
    let (buf, offs) = pack_msgs(msgs);
    let mut ok = vec![0u8; n];
    e.check(unsafe {
        sys::sylow_b200_verify_each(e.ctx, pk.as_ptr(), pki.as_ptr(), buf.as_ptr(), offs.as_ptr(), sg.as_ptr(), sgi.as_ptr(), n,
                                    crate::DST.as_ptr(), crate::DST.len(), sys::SYLOW_B200_HASH_KECCAK256, ok.as_mut_ptr())
    })?;
    Ok(ok.into_iter().map(|b| b != 0).collect())
}

/// prod e(r_i sig_i, G2gen) * e(-r_i H(m_i), pk_i) == 1 with ONE final exponentiation
/// (examples/verify_multiple_messages_same_signer.rs:40-60, generalised to per-message keys).
/// `weight_seed = None` is the reference example's unweighted product (aggregate verification: cancelling errors
/// pass); `Some(seed)` with 32 secret random bytes drawn after the batch is fixed is batch verification with random
/// 64-bit weights, sound for every signature.
pub fn verify_batch(e: &mut Engine, pks: &[G2Projective], msgs: &[&[u8]], sigs: &[G1Projective], weight_seed: Option<&[u8; 32]>) -> Result<bool, BatchError> {
    let n = same_len(pks.len(), same_len(msgs.len(), sigs.len())?)?;
    let (pk, pki) = marshal_g2(pks);
    let (sg, sgi) = marshal_g1(sigs);
    let (buf, offs) = pack_msgs(msgs);
    let seed = weight_seed.map_or(core::ptr::null(), |s| s.as_ptr());
    let mut ok = 0i32;
    e.check(unsafe {
        sys::sylow_b200_verify_batch(e.ctx, pk.as_ptr(), pki.as_ptr(), buf.as_ptr(), offs.as_ptr(), sg.as_ptr(), sgi.as_ptr(), n,
                                     crate::DST.as_ptr(), crate::DST.len(), sys::SYLOW_B200_HASH_KECCAK256, seed, &mut ok)
    })?;
    Ok(ok != 0)
}

/// Threshold aggregation (examples/dkg.rs:190-226): `ids.len() / t` independent sets of `t` partial signatures;
/// `out[s] = sum_i lambda_i * sigs[s*t + i]` with the Lagrange coefficients at 0 computed on the device.
pub fn threshold_aggregate_batch(e: &mut Engine, ids: &[u64], sigs: &[G1Projective], t: usize) -> Result<Vec<G1Affine>, BatchError> {
    if t == 0 || ids.len() != sigs.len() || ids.len() % t != 0 {
        return Err(BatchError::InvalidArgument);
    }
    let n_sets = ids.len() / t;
    let (sg, sgi) = marshal_g1(sigs);
    let (mut out, mut inf) = (vec![0u8; n_sets * 64], vec![0u8; n_sets]);
    e.check(unsafe {
        sys::sylow_b200_threshold_aggregate_batch(e.ctx, ids.as_ptr(), sg.as_ptr(), sgi.as_ptr(), n_sets, t, out.as_mut_ptr(),
                                                  inf.as_mut_ptr())
    })?;
    Ok(out.chunks_exact(64).zip(&inf).map(|(b, &i)| get_g1(b, i)).collect())
}
