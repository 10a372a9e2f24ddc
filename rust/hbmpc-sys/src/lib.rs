//! Raw bindings to `libhbmpc_b200.so` -- one declaration per entry point of `include/hbmpc_b200.h`, same order, same argument
//! meaning (read the header for the contracts; each entry point names the reference function it replaces).
//! Elements are `U256`-style canonical little-endian limbs (`[u64; 4]` per `ark_bls12_381::Fr`, one `u64` per Goldilocks element);
//! every data pointer may be a host or a device pointer; return values are `ShareErrorCode` numbers (0 = success) plus
//! `HBMPC_NO_DEVICE` (100) / `HBMPC_CUDA_ERROR` (101).  A context is thread-compatible: one caller at a time.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct hbmpc_ctx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct hbmpc_group {
    _p: [u8; 0],
}

pub const HBMPC_SUCCESS: c_int = 0;
pub const HBMPC_INSUFFICIENT_SHARES: c_int = 1;
pub const HBMPC_DEGREE_MISMATCH: c_int = 2;
pub const HBMPC_ID_MISMATCH: c_int = 3;
pub const HBMPC_INVALID_INPUT: c_int = 4;
pub const HBMPC_TYPE_MISMATCH: c_int = 5;
pub const HBMPC_NO_SUITABLE_DOMAIN: c_int = 6;
pub const HBMPC_POLYNOMIAL_OPERATION_ERROR: c_int = 7;
pub const HBMPC_DECODING_ERROR: c_int = 8;
pub const HBMPC_NO_DEVICE: c_int = 100;
pub const HBMPC_CUDA_ERROR: c_int = 101;

extern "C" {
    // ---- context
    pub fn hbmpc_ctx_create(device: c_int, out: *mut *mut hbmpc_ctx) -> c_int;
    pub fn hbmpc_ctx_destroy(ctx: *mut hbmpc_ctx);
    pub fn hbmpc_ctx_set_stream(ctx: *mut hbmpc_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn hbmpc_ctx_set_async(ctx: *mut hbmpc_ctx, on: c_int) -> c_int;
    pub fn hbmpc_ctx_synchronize(ctx: *mut hbmpc_ctx) -> c_int;
    pub fn hbmpc_ctx_launch_count(ctx: *const hbmpc_ctx) -> u64;
    pub fn hbmpc_ctx_device(ctx: *const hbmpc_ctx) -> c_int;
    pub fn hbmpc_last_error(ctx: *const hbmpc_ctx) -> *const c_char;

    // ---- K1 / K2: RobustShare::compute_shares, make_vandermonde + apply_vandermonde
    pub fn hbmpc_compute_shares_batch(ctx: *mut hbmpc_ctx, n: usize, d: usize, b: usize, coeffs: *const u64, shares: *mut u64) -> c_int;
    pub fn hbmpc_share_secrets_batch(ctx: *mut hbmpc_ctx, seed32: *const u8, n: usize, d: usize, b: usize, secrets: *const u64,
                                     shares: *mut u64, coeffs_out: *mut u64) -> c_int;
    pub fn hbmpc_apply_vandermonde_batch(ctx: *mut hbmpc_ctx, n: usize, cols: usize, b: usize, input: *const u64, out: *mut u64,
                                         recipient_major: c_int) -> c_int;
    pub fn hbmpc_apply_vandermonde_msgs(ctx: *mut hbmpc_ctx, n: usize, cols: usize, b: usize, input: *const u64,
                                        recipient_out: *const *mut u64) -> c_int;
    pub fn hbmpc_apply_matrix_batch(ctx: *mut hbmpc_ctx, rows: usize, cols: usize, matrix: *const u64, b: usize, input: *const u64,
                                    out: *mut u64, recipient_major: c_int) -> c_int;

    // ---- K3 / K4: batch_recover_secret, RobustShare::recover_secret
    pub fn hbmpc_batch_recover(ctx: *mut hbmpc_ctx, n: usize, d: usize, t: usize, s: usize, sender_ids: *const usize, b: usize,
                               evals: *const u64, coeffs: *mut u64, path: *mut i32, flags: *mut u64) -> c_int;
    pub fn hbmpc_batch_recover_secrets(ctx: *mut hbmpc_ctx, n: usize, d: usize, t: usize, s: usize, sender_ids: *const usize, b: usize,
                                       evals: *const u64, secrets: *mut u64, path: *mut i32) -> c_int;
    pub fn hbmpc_batch_recover_msgs(ctx: *mut hbmpc_ctx, n: usize, d: usize, t: usize, s: usize, sender_ids: *const usize, b: usize,
                                    sender_evals: *const *const u64, coeffs: *mut u64, path: *mut i32, flags: *mut u64) -> c_int;
    pub fn hbmpc_batch_recover_secrets_msgs(ctx: *mut hbmpc_ctx, n: usize, d: usize, t: usize, s: usize, sender_ids: *const usize,
                                            b: usize, sender_evals: *const *const u64, secrets: *mut u64, path: *mut i32) -> c_int;
    pub fn hbmpc_robust_interpolate_batch(ctx: *mut hbmpc_ctx, n: usize, d: usize, t: usize, s: usize, ids: *const usize, b: usize,
                                          shares: *const u64, coeffs: *mut u64, secrets: *mut u64, path: *mut i32, flags: *mut u64) -> c_int;
    pub fn hbmpc_nonrobust_recover_batch(ctx: *mut hbmpc_ctx, n: usize, deg: usize, s: usize, ids: *const usize, b: usize,
                                         shares: *const u64, sender_major: c_int, coeffs: *mut u64, secrets: *mut u64,
                                         status: *mut i32) -> c_int;

    // ---- K5, wire records, randomness
    pub fn hbmpc_elementwise(ctx: *mut hbmpc_ctx, op: c_int, count: usize, a: *const u64, b: *const u64, out: *mut u64) -> c_int;
    /// `op`: 0 triple mask (in = {a, b, r_2t}), 1 Beaver mask (in = {a, x, b, y}; two outputs), 2 Beaver finalise
    /// (in = {c, x, y, a-x, b-y}); `inputs` / `outputs` are arrays of 3|4|5 and 1|2|1 pointers to `count` values each.
    pub fn hbmpc_share_algebra_fused(ctx: *mut hbmpc_ctx, op: c_int, count: usize, inputs: *const *const u64, outputs: *const *mut u64) -> c_int;
    pub fn hbmpc_unpack_share_records(ctx: *mut hbmpc_ctx, count: usize, records: *const c_void, values: *mut u64, ids: *mut u64,
                                      degrees: *mut u64) -> c_int;
    pub fn hbmpc_pack_share_records(ctx: *mut hbmpc_ctx, count: usize, values: *const u64, per_id: usize, degree: usize,
                                    records: *mut c_void) -> c_int;
    pub fn hbmpc_sample_fr_batch(ctx: *mut hbmpc_ctx, seed32: *const u8, count: usize, out: *mut u64) -> c_int;
    pub fn hbmpc_sample_polynomials(ctx: *mut hbmpc_ctx, seed32: *const u8, b: usize, d: usize, secrets: *const u64,
                                    coeffs: *mut u64) -> c_int;

    // ---- Goldilocks (common/math/goldilocks.rs): one u64 per element
    pub fn hbmpc_gl_compute_shares_batch(ctx: *mut hbmpc_ctx, n: usize, d: usize, b: usize, coeffs: *const u64, shares: *mut u64) -> c_int;
    pub fn hbmpc_gl_apply_vandermonde_batch(ctx: *mut hbmpc_ctx, n: usize, cols: usize, b: usize, input: *const u64, out: *mut u64,
                                            recipient_major: c_int) -> c_int;
    pub fn hbmpc_gl_batch_recover(ctx: *mut hbmpc_ctx, n: usize, d: usize, t: usize, s: usize, sender_ids: *const usize, b: usize,
                                  evals: *const u64, coeffs: *mut u64, secrets: *mut u64, path: *mut i32) -> c_int;
    pub fn hbmpc_gl_nonrobust_recover_batch(ctx: *mut hbmpc_ctx, n: usize, deg: usize, s: usize, ids: *const usize, b: usize,
                                            shares: *const u64, sender_major: c_int, coeffs: *mut u64, secrets: *mut u64,
                                            status: *mut i32) -> c_int;
    pub fn hbmpc_gl_elementwise(ctx: *mut hbmpc_ctx, op: c_int, count: usize, a: *const u64, b: *const u64, out: *mut u64) -> c_int;

    // ---- one process, several GPUs
    pub fn hbmpc_group_create(devices: *const c_int, n_devices: usize, out: *mut *mut hbmpc_group) -> c_int;
    pub fn hbmpc_group_destroy(grp: *mut hbmpc_group);
    pub fn hbmpc_group_size(grp: *const hbmpc_group) -> usize;
    pub fn hbmpc_group_ctx(grp: *mut hbmpc_group, i: usize) -> *mut hbmpc_ctx;
    pub fn hbmpc_group_shard_range(grp: *const hbmpc_group, b: usize, i: usize, lo: *mut usize, hi: *mut usize);
    pub fn hbmpc_group_compute_shares_batch(grp: *mut hbmpc_group, n: usize, d: usize, b: usize, coeffs: *const u64, shares: *mut u64) -> c_int;
    pub fn hbmpc_group_apply_vandermonde_batch(grp: *mut hbmpc_group, n: usize, cols: usize, b: usize, input: *const u64, out: *mut u64,
                                               recipient_major: c_int) -> c_int;
    pub fn hbmpc_group_batch_recover(grp: *mut hbmpc_group, n: usize, d: usize, t: usize, s: usize, sender_ids: *const usize, b: usize,
                                     evals: *const u64, coeffs: *mut u64, path: *mut i32, flags: *mut u64) -> c_int;
    pub fn hbmpc_group_batch_recover_secrets(grp: *mut hbmpc_group, n: usize, d: usize, t: usize, s: usize, sender_ids: *const usize,
                                             b: usize, evals: *const u64, secrets: *mut u64, path: *mut i32) -> c_int;
    pub fn hbmpc_group_robust_interpolate_batch(grp: *mut hbmpc_group, n: usize, d: usize, t: usize, s: usize, ids: *const usize,
                                                b: usize, shares: *const u64, coeffs: *mut u64, secrets: *mut u64, path: *mut i32,
                                                flags: *mut u64) -> c_int;

    // ---- probes (roofline denominators)
    pub fn hbmpc_measure_imad_peak(ctx: *mut hbmpc_ctx, variant: c_int, giga_inst_per_s: *mut f64, elapsed_ms: *mut f64) -> c_int;
    pub fn hbmpc_measure_wide_chains(ctx: *mut hbmpc_ctx, chains: c_int, warps_per_smsp: c_int, giga_inst_per_s: *mut f64) -> c_int;
    pub fn hbmpc_measure_mont_mul(ctx: *mut hbmpc_ctx, ilp: c_int, warps_per_smsp: c_int, giga_products_per_s: *mut f64) -> c_int;
}
