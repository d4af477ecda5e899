//! Raw bindings of include/kzg_b200.h (the subset the drop-in needs).
#![allow(non_camel_case_types)]
use std::os::raw::c_char;

#[repr(C)] pub struct kzg_ctx { _p: [u8; 0] }
#[repr(C)] pub struct kzg_setup { _p: [u8; 0] }

pub const KZG_OK: i32 = 0;
pub const KZG_E_DEGREE_TOO_HIGH: i32 = 1;
pub const KZG_E_CONSTANT_POLY: i32 = 2;
pub const KZG_E_REMAINDER_NONZERO: i32 = 3;

extern "C" {
    pub fn kzg_ctx_create(out: *mut *mut kzg_ctx, device_id: i32) -> i32;
    pub fn kzg_ctx_destroy(ctx: *mut kzg_ctx);
    pub fn kzg_last_error(ctx: *const kzg_ctx) -> *const c_char;
    pub fn kzg_setup_generate(ctx: *mut kzg_ctx, secret_be: *const u8, first: u64, n: u64, out: *mut *mut kzg_setup) -> i32;
    pub fn kzg_setup_import_g1(ctx: *mut kzg_ctx, g1_c48: *const u8, n: u64, out: *mut *mut kzg_setup) -> i32;
    pub fn kzg_setup_export_g1(ctx: *mut kzg_ctx, s: *const kzg_setup, first: u64, count: u64, out_c48: *mut u8) -> i32;
    pub fn kzg_setup_len(s: *const kzg_setup) -> u64;
    pub fn kzg_setup_generate_g2(ctx: *mut kzg_ctx, s: *mut kzg_setup, secret_be: *const u8, count: u64) -> i32;
    pub fn kzg_setup_import_g2(ctx: *mut kzg_ctx, s: *mut kzg_setup, g2_c96: *const u8, count: u64) -> i32;
    pub fn kzg_setup_export_g2(ctx: *mut kzg_ctx, s: *const kzg_setup, first: u64, count: u64, out_c96: *mut u8) -> i32;
    pub fn kzg_setup_g2_len(s: *const kzg_setup) -> u64;
    pub fn kzg_setup_destroy(s: *mut kzg_setup);
    pub fn kzg_commit(ctx: *mut kzg_ctx, s: *const kzg_setup, coeffs_le32: *const u8, n: u64, out_c48: *mut u8) -> i32;
    pub fn kzg_evaluate(ctx: *mut kzg_ctx, coeffs_le32: *const u8, n: u64, z_le: *const u8, y_le: *mut u8) -> i32;
    pub fn kzg_open(ctx: *mut kzg_ctx, s: *const kzg_setup, coeffs_le32: *const u8, n: u64, z_le: *const u8,
                    y_expected_le: *const u8, y_out_le: *mut u8, proof_c48: *mut u8) -> i32;
    pub fn kzg_commit_batch(ctx: *mut kzg_ctx, s: *const kzg_setup, coeffs_le32: *const u8, blobs: u64, n: u64, out_c48: *mut u8) -> i32;
    pub fn kzg_open_batch(ctx: *mut kzg_ctx, s: *const kzg_setup, coeffs_le32: *const u8, blobs: u64, n: u64,
                          z_le32: *const u8, y_le32: *mut u8, proofs_c48: *mut u8) -> i32;
    pub fn kzg_commit_open_batch(ctx: *mut kzg_ctx, s: *const kzg_setup, coeffs_le32: *const u8, blobs: u64, n: u64,
                                 z_le32: *const u8, commits_c48: *mut u8, y_le32: *mut u8, proofs_c48: *mut u8) -> i32;
    // precomputed tables (once per setup) and several commitments in flight (ticket t runs on lane t mod `lanes`)
    pub fn kzg_ctx_set_option(ctx: *mut kzg_ctx, name: *const c_char, value: i64) -> i32;
    pub fn kzg_setup_prepare_msm(ctx: *mut kzg_ctx, s: *mut kzg_setup, window_bits: i32) -> i32;
    pub fn kzg_setup_prepare_fixed(ctx: *mut kzg_ctx, s: *mut kzg_setup, n_points: u64, window_bits: i32) -> i32;
    pub fn kzg_msm_g1(ctx: *mut kzg_ctx, s: *const kzg_setup, scalars_le32: *const u8, n: u64, out_c48: *mut u8) -> i32;
    pub fn kzg_msm_g1_submit(ctx: *mut kzg_ctx, s: *const kzg_setup, scalars_le32: *const u8, n: u64, ticket: *mut i32) -> i32;
    pub fn kzg_msm_g1_wait(ctx: *mut kzg_ctx, ticket: i32, out_c48: *mut u8) -> i32;
    // multi-GPU (one process per GPU): peer mailboxes over CUDA IPC, collective sharded MSM
    pub fn kzg_p2p_init(ctx: *mut kzg_ctx, world: i32, rank: i32, handle_out: *mut u8) -> i32;
    pub fn kzg_p2p_connect(ctx: *mut kzg_ctx, handles: *const u8) -> i32;
    pub fn kzg_msm_g1_sharded_submit(ctx: *mut kzg_ctx, s: *const kzg_setup, scalars_le32: *const u8, n: u64, ticket: *mut i32) -> i32;
    // host-side G2 / pairing (only needed when blst is dropped altogether)
    pub fn kzg_g2_from_scalar(k_le: *const u8, out: *mut u8);
    pub fn kzg_g2_sub(a: *const u8, b: *const u8, out: *mut u8) -> i32;
    pub fn kzg_pairing_equal(a1: *const u8, b1: *const u8, a2: *const u8, b2: *const u8, equal: *mut i32) -> i32;
    pub fn kzg_verify_proof(proof: *const u8, commitment: *const u8, z_le: *const u8, y_le: *const u8,
                            g2_tau: *const u8, ok: *mut i32) -> i32;
}
