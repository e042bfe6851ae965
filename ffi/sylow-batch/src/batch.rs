//! `src/batch.rs` of a sylow fork: the batched entry points BASELINE.json's north_star asks for
//! (`pairing_batch`, `verify_batch`, `g1_mul_batch`, ...), implemented on top of `sylow-cuda-sys`.
//!
//! It must live INSIDE sylow because `Gt(pub(crate) Fp12)`, `MillerLoopResult(pub(crate) Fp12)` and the
//! fields of `GroupAffine` are crate-private (src/groups/gt.rs:115, src/pairing.rs:72,
//! src/groups/group.rs:174-181).  `lib.rs` adds:
//!
//! ```ignore
//! mod batch;
//! pub use crate::batch::{g1_mul_batch, g2_mul_batch, pairing_batch, pairing_check_batch, sign_batch,
//!                        threshold_aggregate_batch, verify_batch, verify_each, Engine};
//! ```
//!
//! NOT COMPILED in the build image (no Rust toolchain); the C ABI underneath is what the parity tests
//! exercise (Python ctypes + include/sylow_b200.hpp).
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

/// One context per GPU.  `Engine::new(&[0, 1, .., 7])` for an 8xB200 box.
pub struct Engine {
    ctxs: Vec<*mut sys::SylowB200Ctx>,
}

// A context is used by one call at a time; `&mut self` on every method enforces it.
unsafe impl Send for Engine {}

fn status(code: i32) -> Result<(), GroupError> {
    match code {
        sys::SYLOW_B200_OK => Ok(()),
        sys::SYLOW_B200_ERR_NOT_ON_CURVE => Err(GroupError::NotOnCurve),
        sys::SYLOW_B200_ERR_NOT_IN_SUBGROUP => Err(GroupError::NotInSubgroup),
        sys::SYLOW_B200_ERR_CANNOT_HASH => Err(GroupError::CannotHashToGroup),
        _ => Err(GroupError::DecodeError),
    }
}

impl Engine {
    pub fn new(devices: &[i32]) -> Result<Self, GroupError> {
        let mut ctxs = Vec::with_capacity(devices.len());
        for &d in devices {
            let mut c = core::ptr::null_mut();
            status(unsafe { sys::sylow_b200_create(&mut c, d) })?;
            ctxs.push(c);
        }
        Ok(Engine { ctxs })
    }
    /// Contiguous slice `[g*n/G, (g+1)*n/G)` of a batch of `n` for GPU `g`.
    fn slice(&self, g: usize, n: usize) -> core::ops::Range<usize> {
        let k = self.ctxs.len();
        (g * n / k)..((g + 1) * n / k)
    }
}

impl Drop for Engine {
    fn drop(&mut self) {
        for &c in &self.ctxs {
            unsafe { sys::sylow_b200_destroy(c) };
        }
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
fn pack_msgs(msgs: &[&[u8]]) -> (Vec<u8>, Vec<u64>) {
    let mut offs = Vec::with_capacity(msgs.len() + 1);
    let mut buf = Vec::new();
    offs.push(0u64);
    for m in msgs {
        buf.extend_from_slice(m);
        offs.push(buf.len() as u64);
    }
    (buf, offs)
}

/// `p.iter().zip(q).map(|(p, q)| pairing(p, q))` on the GPUs (src/pairing.rs:870-893).
pub fn pairing_batch(e: &mut Engine, p: &[G1Projective], q: &[G2Projective]) -> Result<Vec<Gt>, GroupError> {
    assert_eq!(p.len(), q.len());
    let n = p.len();
    let (g1, g1i) = marshal_g1(p);
    let (g2, g2i) = marshal_g2(q);
    let mut out = vec![0u8; n * 384];
    // one scoped thread per GPU, each on its contiguous slice; no data-path collective
    std::thread::scope(|s| {
        let mut rest: &mut [u8] = &mut out;
        let mut hs = Vec::new();
        for (g, &c) in e.ctxs.iter().enumerate() {
            let r = e.slice(g, n);
            let (mine, tail) = rest.split_at_mut(r.len() * 384);
            rest = tail;
            let (g1, g1i, g2, g2i) = (&g1, &g1i, &g2, &g2i);
            let c = c as usize;
            hs.push(s.spawn(move || unsafe {
                sys::sylow_b200_pairing_batch(c as *mut _, g1[r.start * 64..].as_ptr(), g1i[r.start..].as_ptr(),
                                              g2[r.start * 128..].as_ptr(), g2i[r.start..].as_ptr(), r.len(),
                                              mine.as_mut_ptr())
            }));
        }
        hs.into_iter().try_for_each(|h| status(h.join().unwrap()))
    })?;
    Ok(out.chunks_exact(384).map(|b| Gt(get_fp12(b))).collect())
}

/// `glued_miller_loop` over the batch (src/pairing.rs:970-1022): per-GPU partial products are combined
/// with `sylow_b200_fp12_product` (7 Fp12 multiplications for 8 GPUs).
pub fn miller_product(e: &mut Engine, p: &[G1Projective], q: &[G2Projective]) -> Result<MillerLoopResult, GroupError> {
    let n = p.len().min(q.len()); // zip truncation, like the reference (:975)
    let (g1, g1i) = marshal_g1(&p[..n]);
    let (g2, g2i) = marshal_g2(&q[..n]);
    let k = e.ctxs.len();
    let mut partials = vec![0u8; k * 384];
    for (g, &c) in e.ctxs.iter().enumerate() {
        let r = e.slice(g, n);
        status(unsafe {
            sys::sylow_b200_miller_product(c, g1[r.start * 64..].as_ptr(), g1i[r.start..].as_ptr(),
                                           g2[r.start * 128..].as_ptr(), g2i[r.start..].as_ptr(), r.len(),
                                           partials[g * 384..].as_mut_ptr())
        })?;
    }
    let mut out = [0u8; 384];
    status(unsafe { sys::sylow_b200_fp12_product(e.ctxs[0], partials.as_ptr(), k, out.as_mut_ptr()) })?;
    Ok(MillerLoopResult(get_fp12(&out)))
}

/// `scalars[i] * pts[i]` (src/groups/group.rs:639-667), returned affine.
pub fn g1_mul_batch(e: &mut Engine, pts: &[G1Projective], scalars: &[Fp]) -> Result<Vec<G1Affine>, GroupError> {
    assert_eq!(pts.len(), scalars.len());
    let n = pts.len();
    let (g1, g1i) = marshal_g1(pts);
    let mut ks = Vec::with_capacity(n * 32);
    scalars.iter().for_each(|k| put_fp(&mut ks, k));
    let (mut out, mut inf) = (vec![0u8; n * 64], vec![0u8; n]);
    status(unsafe {
        sys::sylow_b200_g1_mul_batch(e.ctxs[0], g1.as_ptr(), g1i.as_ptr(), ks.as_ptr(), n, out.as_mut_ptr(),
                                     inf.as_mut_ptr())
    })?;
    Ok(out.chunks_exact(64).zip(&inf)
        .map(|(b, &i)| G1Affine { x: get_fp(&b[..32]), y: get_fp(&b[32..]), infinity: Choice::from(i) })
        .collect())
}

/// Threshold aggregation (examples/dkg.rs:190-226): `ids.len() / t` independent sets of `t` partial signatures;
/// `out[s] = sum_i lambda_i * sigs[s*t + i]` with the Lagrange coefficients at 0 computed on the device.
pub fn threshold_aggregate_batch(e: &mut Engine, ids: &[u64], sigs: &[G1Projective], t: usize) -> Result<Vec<G1Affine>, GroupError> {
    assert!(t > 0 && ids.len() == sigs.len() && ids.len() % t == 0);
    let n_sets = ids.len() / t;
    let (sg, sgi) = marshal_g1(sigs);
    let (mut out, mut inf) = (vec![0u8; n_sets * 64], vec![0u8; n_sets]);
    status(unsafe {
        sys::sylow_b200_threshold_aggregate_batch(e.ctxs[0], ids.as_ptr(), sg.as_ptr(), sgi.as_ptr(), n_sets, t,
                                                  out.as_mut_ptr(), inf.as_mut_ptr())
    })?;
    Ok(out.chunks_exact(64).zip(&inf)
        .map(|(b, &i)| G1Affine { x: get_fp(&b[..32]), y: get_fp(&b[32..]), infinity: Choice::from(i) })
        .collect())
}

/// prod e(sig_i, G2gen) * e(-H(m_i), pk_i) == 1 with ONE final exponentiation
/// (examples/verify_multiple_messages_same_signer.rs:40-60, generalised to per-message keys).
pub fn verify_batch(e: &mut Engine, pks: &[G2Projective], msgs: &[&[u8]], sigs: &[G1Projective]) -> Result<bool, GroupError> {
    assert!(pks.len() == msgs.len() && msgs.len() == sigs.len());
    let n = msgs.len();
    let (pk, _) = marshal_g2(pks);
    let (sg, _) = marshal_g1(sigs);
    let k = e.ctxs.len();
    let mut partials = vec![0u8; k * 384];
    for (g, &c) in e.ctxs.iter().enumerate() {
        let r = e.slice(g, n);
        let (buf, offs) = pack_msgs(&msgs[r.clone()]);
        status(unsafe {
            sys::sylow_b200_verify_batch_partial(c, pk[r.start * 128..].as_ptr(), buf.as_ptr(), offs.as_ptr(),
                                                 sg[r.start * 64..].as_ptr(), r.len(), crate::DST.as_ptr(),
                                                 crate::DST.len(), sys::SYLOW_B200_HASH_KECCAK256,
                                                 partials[g * 384..].as_mut_ptr())
        })?;
    }
    let mut ok = 0i32;
    status(unsafe { sys::sylow_b200_verify_batch_finish(e.ctxs[0], partials.as_ptr(), k, &mut ok) })?;
    Ok(ok != 0)
}
