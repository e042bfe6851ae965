//! Raw bindings to `libsylow_b200.so` (include/sylow_b200.h).  GENERATED from the header by the
//! snippet in INTEGRATION.md section 2; every function is `unsafe` because it takes raw pointers.
//!
//! Not compiled in the build image (no Rust toolchain there); the same C ABI is exercised by the Python
//! ctypes harness and the C++ header.
#![no_std]
#![allow(non_camel_case_types)]

/// Opaque context: one per CUDA device.
#[repr(C)]
pub struct SylowB200Ctx {
    _private: [u8; 0],
}

pub const SYLOW_B200_OK: i32 = 0;
pub const SYLOW_B200_ERR_ARG: i32 = -1;
pub const SYLOW_B200_ERR_CUDA: i32 = -2;
pub const SYLOW_B200_ERR_NOT_ON_CURVE: i32 = -3;
pub const SYLOW_B200_ERR_NOT_IN_SUBGROUP: i32 = -4;
pub const SYLOW_B200_ERR_CANNOT_HASH: i32 = -5;
pub const SYLOW_B200_ERR_DECODE: i32 = -6;
pub const SYLOW_B200_ERR_NOMEM: i32 = -7;
pub const SYLOW_B200_HASH_KECCAK256: i32 = 0;
pub const SYLOW_B200_HASH_SHA256: i32 = 1;

extern "C" {
    pub fn sylow_b200_create(out: *mut *mut SylowB200Ctx, device_id: i32) -> i32;
    pub fn sylow_b200_destroy(ctx: *mut SylowB200Ctx) -> i32;
    pub fn sylow_b200_strerror(status: i32) -> *const core::ffi::c_char;
    pub fn sylow_b200_last_cuda_error(ctx: *const SylowB200Ctx) -> i32;
    pub fn sylow_b200_launch_count(ctx: *const SylowB200Ctx) -> u64;
    pub fn sylow_b200_pairing_batch(ctx: *mut SylowB200Ctx, g1: *const u8, g1_inf: *const u8, g2: *const u8, g2_inf: *const u8, n: usize, gt_out: *mut u8) -> i32;
    pub fn sylow_b200_miller_loop_batch(ctx: *mut SylowB200Ctx, g1: *const u8, g1_inf: *const u8, g2: *const u8, g2_inf: *const u8, n: usize, f_out: *mut u8) -> i32;
    pub fn sylow_b200_miller_product(ctx: *mut SylowB200Ctx, g1: *const u8, g1_inf: *const u8, g2: *const u8, g2_inf: *const u8, n: usize, f_out: *mut u8) -> i32;
    pub fn sylow_b200_final_exp_batch(ctx: *mut SylowB200Ctx, f: *const u8, n: usize, gt_out: *mut u8) -> i32;
    pub fn sylow_b200_fp12_product(ctx: *mut SylowB200Ctx, f: *const u8, n: usize, out: *mut u8) -> i32;
    pub fn sylow_b200_pairing_check_batch(ctx: *mut SylowB200Ctx, g1: *const u8, g1_inf: *const u8, g2: *const u8, g2_inf: *const u8, pairs_per_check: usize, n_checks: usize, ok_out: *mut u8) -> i32;
    pub fn sylow_b200_g2_precompute(ctx: *mut SylowB200Ctx, g2: *const u8, n: usize, coeffs_out: *mut u8) -> i32;
    pub fn sylow_b200_miller_loop_precomputed(ctx: *mut SylowB200Ctx, coeffs: *const u8, g1: *const u8, g1_inf: *const u8, n: usize, f_out: *mut u8) -> i32;
    pub fn sylow_b200_pairing_check_fixed_batch(ctx: *mut SylowB200Ctx, g1: *const u8, g1_inf: *const u8, g2_var: *const u8, g2_var_inf: *const u8, k_var: usize, coeffs_fixed: *const u8, k_fixed: usize, n_checks: usize, ok_out: *mut u8) -> i32;
    pub fn sylow_b200_g1_validate_batch(ctx: *mut SylowB200Ctx, g1: *const u8, g1_inf: *const u8, n: usize, status_out: *mut i8) -> i32;
    pub fn sylow_b200_g2_validate_batch(ctx: *mut SylowB200Ctx, g2: *const u8, g2_inf: *const u8, n: usize, status_out: *mut i8) -> i32;
    pub fn sylow_b200_g1_from_be_bytes_batch(ctx: *mut SylowB200Ctx, be: *const u8, n: usize, eip_mode: i32, g1_out: *mut u8, inf_out: *mut u8, status_out: *mut i8) -> i32;
    pub fn sylow_b200_g2_from_be_bytes_batch(ctx: *mut SylowB200Ctx, be: *const u8, n: usize, eip_mode: i32, g2_out: *mut u8, inf_out: *mut u8, status_out: *mut i8) -> i32;
    pub fn sylow_b200_g1_to_be_bytes_batch(ctx: *mut SylowB200Ctx, g1: *const u8, g1_inf: *const u8, n: usize, scrubbed: i32, be_out: *mut u8) -> i32;
    pub fn sylow_b200_g2_to_be_bytes_batch(ctx: *mut SylowB200Ctx, g2: *const u8, g2_inf: *const u8, n: usize, scrubbed: i32, be_out: *mut u8) -> i32;
    pub fn sylow_b200_eip197_pairing_check_batch(ctx: *mut SylowB200Ctx, input: *const u8, k: usize, n_checks: usize, ok_out: *mut u8, status_out: *mut i8) -> i32;
    pub fn sylow_b200_g1_mul_batch(ctx: *mut SylowB200Ctx, pts: *const u8, pts_inf: *const u8, scalars: *const u8, n: usize, out: *mut u8, out_inf: *mut u8) -> i32;
    pub fn sylow_b200_g2_mul_batch(ctx: *mut SylowB200Ctx, pts: *const u8, pts_inf: *const u8, scalars: *const u8, n: usize, out: *mut u8, out_inf: *mut u8) -> i32;
    pub fn sylow_b200_gt_mul_batch(ctx: *mut SylowB200Ctx, gt: *const u8, scalars: *const u8, n: usize, out: *mut u8) -> i32;
    pub fn sylow_b200_hash_to_g1_batch(ctx: *mut SylowB200Ctx, msgs: *const u8, offsets: *const u64, n: usize, dst: *const u8, dst_len: usize, hash_id: i32, out: *mut u8, out_inf: *mut u8) -> i32;
    pub fn sylow_b200_expand_message_batch(ctx: *mut SylowB200Ctx, msgs: *const u8, offsets: *const u64, n: usize, dst: *const u8, dst_len: usize, hash_id: i32, len_in_bytes: usize, out: *mut u8) -> i32;
    pub fn sylow_b200_hash_to_field_batch(ctx: *mut SylowB200Ctx, msgs: *const u8, offsets: *const u64, n: usize, dst: *const u8, dst_len: usize, hash_id: i32, out: *mut u8) -> i32;
    pub fn sylow_b200_sign_batch(ctx: *mut SylowB200Ctx, sks: *const u8, msgs: *const u8, offsets: *const u64, n: usize, dst: *const u8, dst_len: usize, hash_id: i32, sigs_out: *mut u8) -> i32;
    pub fn sylow_b200_verify_each(ctx: *mut SylowB200Ctx, pks: *const u8, msgs: *const u8, offsets: *const u64, sigs: *const u8, n: usize, dst: *const u8, dst_len: usize, hash_id: i32, ok_out: *mut u8) -> i32;
    pub fn sylow_b200_verify_batch_partial(ctx: *mut SylowB200Ctx, pks: *const u8, msgs: *const u8, offsets: *const u64, sigs: *const u8, n: usize, dst: *const u8, dst_len: usize, hash_id: i32, f_out: *mut u8) -> i32;
    pub fn sylow_b200_verify_batch_finish(ctx: *mut SylowB200Ctx, partials: *const u8, n_partials: usize, ok: *mut i32) -> i32;
    pub fn sylow_b200_verify_batch(ctx: *mut SylowB200Ctx, pks: *const u8, msgs: *const u8, offsets: *const u64, sigs: *const u8, n: usize, dst: *const u8, dst_len: usize, hash_id: i32, ok: *mut i32) -> i32;
    pub fn sylow_b200_pairing_batch_dev(ctx: *mut SylowB200Ctx, d_g1: *const u8, d_g1_inf: *const u8, d_g2: *const u8, d_g2_inf: *const u8, n: usize, d_gt_out: *mut u8, stream: *mut core::ffi::c_void) -> i32;
    pub fn sylow_b200_miller_loop_batch_dev(ctx: *mut SylowB200Ctx, d_g1: *const u8, d_g1_inf: *const u8, d_g2: *const u8, d_g2_inf: *const u8, n: usize, d_f_out: *mut u8, stream: *mut core::ffi::c_void) -> i32;
    pub fn sylow_b200_miller_product_dev(ctx: *mut SylowB200Ctx, d_g1: *const u8, d_g1_inf: *const u8, d_g2: *const u8, d_g2_inf: *const u8, n: usize, d_f_out: *mut u8, stream: *mut core::ffi::c_void) -> i32;
    pub fn sylow_b200_final_exp_batch_dev(ctx: *mut SylowB200Ctx, d_f: *const u8, n: usize, d_gt_out: *mut u8, stream: *mut core::ffi::c_void) -> i32;
    pub fn sylow_b200_pairing_check_batch_dev(ctx: *mut SylowB200Ctx, d_g1: *const u8, d_g1_inf: *const u8, d_g2: *const u8, d_g2_inf: *const u8, pairs_per_check: usize, n_checks: usize, d_ok_out: *mut u8, stream: *mut core::ffi::c_void) -> i32;
    pub fn sylow_b200_pairing_check_fixed_batch_dev(ctx: *mut SylowB200Ctx, d_g1: *const u8, d_g1_inf: *const u8, d_g2_var: *const u8, d_g2_var_inf: *const u8, k_var: usize, d_tables: *const u8, k_fixed: usize, n_checks: usize, d_ok_out: *mut u8, stream: *mut core::ffi::c_void) -> i32;
    pub fn sylow_b200_tables_to_device(ctx: *mut SylowB200Ctx, coeffs: *const u8, k: usize, d_tables_out: *mut u8, stream: *mut core::ffi::c_void) -> i32;
    pub fn sylow_b200_g1_mul_batch_dev(ctx: *mut SylowB200Ctx, d_pts: *const u8, d_pts_inf: *const u8, d_scalars: *const u8, n: usize, d_out: *mut u8, d_out_inf: *mut u8, stream: *mut core::ffi::c_void) -> i32;
    pub fn sylow_b200_g2_mul_batch_dev(ctx: *mut SylowB200Ctx, d_pts: *const u8, d_pts_inf: *const u8, d_scalars: *const u8, n: usize, d_out: *mut u8, d_out_inf: *mut u8, stream: *mut core::ffi::c_void) -> i32;
    pub fn sylow_b200_hash_to_g1_batch_dev(ctx: *mut SylowB200Ctx, d_msgs: *const u8, d_offsets: *const u64, n: usize, dst: *const u8, dst_len: usize, hash_id: i32, d_out: *mut u8, d_out_inf: *mut u8, stream: *mut core::ffi::c_void) -> i32;
    pub fn sylow_b200_verify_batch_partial_dev(ctx: *mut SylowB200Ctx, d_pks: *const u8, d_msgs: *const u8, d_offsets: *const u64, d_sigs: *const u8, n: usize, dst: *const u8, dst_len: usize, hash_id: i32, d_f_out: *mut u8, stream: *mut core::ffi::c_void) -> i32;
    pub fn sylow_b200_fp_op_batch(ctx: *mut SylowB200Ctx, op: i32, a: *const u8, b: *const u8, n: usize, out: *mut u8) -> i32;
    pub fn sylow_b200_fp12_op_batch(ctx: *mut SylowB200Ctx, op: i32, a: *const u8, b: *const u8, n: usize, out: *mut u8) -> i32;
    pub fn sylow_b200_imad_probe(ctx: *mut SylowB200Ctx, variant: i32, blocks: i32, threads: i32, iters: i32, ms_out: *mut f32, ops_out: *mut f64) -> i32;
}
