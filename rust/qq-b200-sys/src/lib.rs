//! Raw + safe bindings of libqq_b200.so (include/qq_b200.h).  NOT BUILT IN THIS ENVIRONMENT (no cargo/rustc).
//! All `unsafe` lives here so that quisquislib can keep `#![deny(unsafe_code)]` (reference src/lib.rs:4).
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct QqCtx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct QqPrepared {
    _private: [u8; 0],
}

#[repr(C)]
pub struct QqMulti {
    _private: [u8; 0],
}

extern "C" {
    pub fn qq_init(ctx: *mut *mut QqCtx, device: c_int) -> c_int;
    pub fn qq_destroy(ctx: *mut QqCtx);
    pub fn qq_last_error(ctx: *const QqCtx) -> *const c_char;
    pub fn qq_update_public_key_batch(ctx: *mut QqCtx, pk: *const u8, r: *const u8, out_pk: *mut u8, status: *mut u8, n: usize) -> c_int;
    pub fn qq_verify_public_key_update_batch(ctx: *mut QqCtx, upd: *const u8, pk: *const u8, r: *const u8, status: *mut u8, n: usize) -> c_int;
    pub fn qq_generate_commitment_batch(ctx: *mut QqCtx, pk: *const u8, r: *const u8, v: *const u8, out: *mut u8, status: *mut u8, n: usize) -> c_int;
    pub fn qq_add_commitments_batch(ctx: *mut QqCtx, a: *const u8, b: *const u8, negate_b: c_int, out: *mut u8, status: *mut u8, n: usize) -> c_int;
    pub fn qq_mul_commitment_batch(ctx: *mut QqCtx, comm: *const u8, s: *const u8, out: *mut u8, status: *mut u8, n: usize) -> c_int;
    pub fn qq_update_account_batch(ctx: *mut QqCtx, acc: *const u8, bl: *const u8, u: *const u8, c: *const u8, out: *mut u8, status: *mut u8, n: usize) -> c_int;
    pub fn qq_verify_account_batch(ctx: *mut QqCtx, acc: *const u8, sk: *const u8, bl: *const u8, status: *mut u8, n: usize) -> c_int;
    pub fn qq_delta_epsilon_batch(ctx: *mut QqCtx, acc: *const u8, bl: *const u8, r: *const u8, base_pk: *const u8, delta: *mut u8, eps: *mut u8, status: *mut u8, n: usize) -> c_int;
    pub fn qq_delta_identity_check(ctx: *mut QqCtx, acc: *const u8, n: usize, verdict: *mut u8) -> c_int;
    pub fn qq_fixed_base_batch(ctx: *mut QqCtx, which: c_int, s: *const u8, out: *mut u8, status: *mut u8, n: usize) -> c_int;
    pub fn qq_fixed_base_i64_batch(ctx: *mut QqCtx, which: c_int, v: *const i64, out: *mut u8, n: usize) -> c_int;
    pub fn qq_fixed_base_set_window(ctx: *mut QqCtx, which: c_int, window_bits: c_int) -> c_int;
    pub fn qq_fixed_base_window(ctx: *const QqCtx, which: c_int) -> c_int;
    pub fn qq_msm(ctx: *mut QqCtx, scalars: *const u8, points: *const u8, n: usize, out: *mut u8, status: *mut u8) -> c_int;
    pub fn qq_msm_partial(ctx: *mut QqCtx, scalars: *const u8, points: *const u8, n: usize, out_xyzt: *mut u8, status: *mut u8) -> c_int;
    pub fn qq_msm_points_prepare(ctx: *mut QqCtx, points: *const u8, n: usize, out: *mut *mut QqPrepared) -> c_int;
    pub fn qq_msm_points_free(ctx: *mut QqCtx, p: *mut QqPrepared);
    pub fn qq_msm_points_count(p: *const QqPrepared) -> usize;
    pub fn qq_msm_prepared(ctx: *mut QqCtx, scalars: *const u8, points: *const QqPrepared, n: usize, out: *mut u8, status: *mut u8) -> c_int;
    pub fn qq_verify_update_account_dlog_batch(ctx: *mut QqCtx, transcript_label: *const c_char, verifier_label: *const c_char, input_accounts: *const u8, delta_accounts: *const u8, z: *const u8, x: *const u8, n: usize, nproofs: usize, status: *mut u8) -> c_int;
    pub fn qq_verify_delta_compact_batch(ctx: *mut QqCtx, transcript_label: *const c_char, verifier_label: *const c_char, delta_accounts: *const u8, epsilon_accounts: *const u8, zv: *const u8, zr1: *const u8, zr2: *const u8, x: *const u8, n: usize, nproofs: usize, status: *mut u8) -> c_int;
    pub fn qq_decommit_batch(ctx: *mut QqCtx, comm: *const u8, sk: *const u8, out: *mut u8, status: *mut u8, n: usize) -> c_int;
    pub fn qq_decommit_value_batch(ctx: *mut QqCtx, comm: *const u8, sk: *const u8, search_bits: c_int, out_values: *mut u64, status: *mut u8, n: usize) -> c_int;
    pub fn qq_from_uniform_bytes_batch(ctx: *mut QqCtx, uniform64: *const u8, out: *mut u8, n: usize) -> c_int;
    pub fn qq_vector_pedersen_gens(ctx: *mut QqCtx, capacity: usize, out_h: *mut u8, out_g: *mut u8) -> c_int;
    pub fn qq_bulletproof_gens(ctx: *mut QqCtx, gens_capacity: usize, party_capacity: usize, out_g: *mut u8, out_h: *mut u8) -> c_int;
    pub fn qq_points_sum(ctx: *mut QqCtx, xyzt: *const u8, k: usize, out: *mut u8, is_identity: *mut u8) -> c_int;
    // batched verifiers (host Merlin transcripts + GPU MSM batches); layouts in include/qq_b200.h
    pub fn qq_verify_account_sigma_batch(ctx: *mut QqCtx, transcript_label: *const c_char, verifier_label: *const c_char, delta_accounts: *const u8, epsilon_accounts: *const u8, base_pk: *const u8, zv: *const u8, zsk: *const u8, zr: *const u8, x: *const u8, n: usize, nproofs: usize, status: *mut u8) -> c_int;
    pub fn qq_verify_zero_balance_batch(ctx: *mut QqCtx, transcript_label: *const c_char, verifier_label: *const c_char, accounts: *const u8, z: *const u8, x: *const u8, n: usize, nproofs: usize, vector_form: c_int, status: *mut u8) -> c_int;
    pub fn qq_verify_destroy_account_batch(ctx: *mut QqCtx, transcript_label: *const c_char, verifier_label: *const c_char, accounts: *const u8, z: *const u8, x: *const u8, n: usize, nproofs: usize, status: *mut u8) -> c_int;
    pub fn qq_verify_same_value_compact_batch(ctx: *mut QqCtx, enc_accounts: *const u8, commitments: *const u8, zv: *const u8, zr: *const u8, x: *const u8, nproofs: usize, status: *mut u8) -> c_int;
    pub fn qq_verify_update_account_dark_tx_batch(ctx: *mut QqCtx, transcript_label: *const c_char, verifier_label: *const c_char, delta_accounts: *const u8, output_accounts: *const u8, z: *const u8, x: *const u8, n: usize, nproofs: usize, status: *mut u8) -> c_int;
    pub fn qq_verify_ddh_batch(ctx: *mut QqCtx, transcript_label: *const c_char, verifier_label: *const c_char, g: *const u8, h: *const u8, g_dash: *const u8, h_dash: *const u8, challenge: *const u8, z: *const u8, nproofs: usize, status: *mut u8) -> c_int;
    pub fn qq_verify_svp_batch(ctx: *mut QqCtx, transcript_label: *const c_char, verifier_label: *const c_char, commitment_a: *const u8, b: *const u8, proof: *const u8, nproofs: usize, status: *mut u8) -> c_int;
    pub fn qq_verify_hadamard_batch(ctx: *mut QqCtx, transcript_label: *const c_char, verifier_label: *const c_char, omega: *const u8, commit_a: *const u8, commit_b: *const u8, commit_c: *const u8, proof: *const u8, nproofs: usize, status: *mut u8, detail: *mut u8) -> c_int;
    pub fn qq_verify_product_batch(ctx: *mut QqCtx, transcript_label: *const c_char, verifier_label: *const c_char, c_prod_a: *const u8, statement: *const u8, proof: *const u8, nproofs: usize, status: *mut u8, detail: *mut u8) -> c_int;
    pub fn qq_verify_shuffle_batch(ctx: *mut QqCtx, transcript_label: *const c_char, verifier_label: *const c_char, shuffle_input: *const u8, shuffle_output: *const u8, statement: *const u8, proof: *const u8, nproofs: usize, status: *mut u8, stage: *mut u8, detail: *mut u8) -> c_int;
    pub fn qq_verify_range_proof_batch(ctx: *mut QqCtx, transcript_label: *const c_char, verifier_label: *const c_char, transcript_state: *const u8, domain_label: *const c_char, commitments: *const u8, proofs: *const u8, n_bits: usize, m: usize, chain: usize, nproofs: usize, status: *mut u8) -> c_int;
    // round 2: wire format, secret mode, prover commitments, multi-device handle, device-side combine
    pub fn qq_shuffle_proofs_from_bincode(input: *const u8, in_len: usize, nproofs: usize, out_proofs: *mut u8, consumed: *mut usize) -> c_int;
    pub fn qq_shuffle_statements_from_bincode(input: *const u8, in_len: usize, nproofs: usize, out_statements: *mut u8, consumed: *mut usize) -> c_int;
    pub fn qq_shuffle_proofs_to_bincode(proofs: *const u8, nproofs: usize, out: *mut u8, out_cap: usize, written: *mut usize) -> c_int;
    pub fn qq_shuffle_statements_to_bincode(statements: *const u8, nproofs: usize, out: *mut u8, out_cap: usize, written: *mut usize) -> c_int;
    pub fn qq_accounts_from_bincode(input: *const u8, in_len: usize, out_accounts: *mut u8, cap_accounts: usize, n_accounts: *mut usize, consumed: *mut usize) -> c_int;
    pub fn qq_sigma_proof_from_bincode(input: *const u8, in_len: usize, variant: *mut c_int, out_scalars: *mut u8, cap_scalars: usize, lens: *mut usize, out_x: *mut u8, consumed: *mut usize) -> c_int;
    pub fn qq_set_secret_mode(ctx: *mut QqCtx, on: c_int) -> c_int;
    pub fn qq_secret_mode(ctx: *const QqCtx) -> c_int;
    pub fn qq_sigma_commit_batch(ctx: *mut QqCtx, points: *const u8, r: *const u8, v: *const u8, out_points: *mut u8, status: *mut u8, n: usize) -> c_int;
    pub fn qq_init_multi(out: *mut *mut QqMulti, devices: *const c_int, ndev: c_int) -> c_int;
    pub fn qq_destroy_multi(m: *mut QqMulti);
    pub fn qq_multi_device_count(m: *const QqMulti) -> c_int;
    pub fn qq_multi_ctx(m: *mut QqMulti, index: c_int) -> *mut QqCtx;
    pub fn qq_multi_update_account_batch(m: *mut QqMulti, acc: *const u8, bl: *const u8, u: *const u8, c: *const u8, out_acc: *mut u8, status: *mut u8, n: usize) -> c_int;
    pub fn qq_multi_msm(m: *mut QqMulti, scalars: *const u8, points: *const u8, n: usize, out_point: *mut u8, status: *mut u8) -> c_int;
    pub fn qq_multi_verify_shuffle_batch(m: *mut QqMulti, transcript_label: *const c_char, verifier_label: *const c_char, shuffle_input: *const u8, shuffle_output: *const u8, statement: *const u8, proof: *const u8, nproofs: usize, status: *mut u8, stage: *mut u8, detail: *mut u8) -> c_int;
    pub fn qq_points_sum_dev(ctx: *mut QqCtx, records: *const u8, k: usize, stride: usize, out_point: *mut u8, is_identity: *mut u8, status: *mut u8) -> c_int;
    pub fn qq_warp_ops_selftest(ctx: *mut QqCtx, p_xyzt: *const u8, q_xyzt: *const u8, n: usize, out_dbl: *mut u8, out_add: *mut u8) -> c_int;
    pub fn qq_msm_set_shifted(ctx: *mut QqCtx, budget_bytes: usize, use_it: c_int) -> c_int;
    pub fn qq_transcript_state_bytes() -> usize;
    pub fn qq_verify_set_transcripts(ctx: *mut QqCtx, on_device: c_int) -> c_int;
    pub fn qq_verify_set_aggregation(ctx: *mut QqCtx, on: c_int) -> c_int;
    pub fn qq_msm_set_overlap(ctx: *mut QqCtx, split_min: std::os::raw::c_long, tail_pct: c_int, sort_blocks_per_sm: c_int) -> c_int;
    pub fn qq_transcript_capture(ctx: *mut QqCtx, states_out: *mut u8, capacity_states: usize) -> c_int;
    pub fn qq_msm_segmented(ctx: *mut QqCtx, scalars: *const u8, points: *const u8, offsets: *const u32, m: usize, out: *mut u8, status: *mut u8) -> c_int;
    pub fn qq_msm_grouped(ctx: *mut QqCtx, scalars: *const u8, points: *const u8, offsets: *const u32, m: usize, out: *mut u8, status: *mut u8) -> c_int;
}

pub const ST_BAD_POINT: u8 = 1;

/// Owning handle of one GPU context.
pub struct Gpu(*mut QqCtx);

impl Gpu {
    pub fn new(device: i32) -> Result<Gpu, i32> {
        let mut p: *mut QqCtx = std::ptr::null_mut();
        let rc = unsafe { qq_init(&mut p, device) };
        if rc == 0 { Ok(Gpu(p)) } else { Err(rc) }
    }
    /// `Account::update_account` over n accounts (128 B each); panics where the reference's `.unwrap()` would.
    pub fn update_account(&self, acc: &[u8], bl: &[u8], u: &[u8], c: &[u8]) -> Vec<u8> {
        let n = bl.len() / 32;
        assert!(acc.len() == n * 128 && u.len() == n * 32 && c.len() == n * 32);
        let mut out = vec![0u8; n * 128];
        let mut st = vec![0u8; n];
        let rc = unsafe { qq_update_account_batch(self.0, acc.as_ptr(), bl.as_ptr(), u.as_ptr(), c.as_ptr(), out.as_mut_ptr(), st.as_mut_ptr(), n) };
        assert_eq!(rc, 0, "qq_update_account_batch failed");
        // every status is handled: an undecodable point is the reference's `.unwrap()` panic; a non-canonical scalar cannot
        // be built through dalek's `Scalar` API, so it is a caller bug here - never return partially valid output
        for &s in &st {
            match s {
                0 => {}
                ST_BAD_POINT => panic!("called `Option::unwrap()` on a `None` value"),
                ST_BAD_SCALAR => panic!("qq_update_account_batch: non-canonical scalar"),
                other => panic!("qq_update_account_batch: unexpected status {}", other),
            }
        }
        out
    }
    /// `Verifier::multiscalar_multiplication`: None if any point fails to decompress.
    pub fn msm(&self, scalars: &[u8], points: &[u8]) -> Option<[u8; 32]> {
        let n = scalars.len() / 32;
        assert!(points.len() == n * 32);
        let (mut out, mut st) = ([0u8; 32], 0u8);
        let rc = unsafe { qq_msm(self.0, scalars.as_ptr(), points.as_ptr(), n, out.as_mut_ptr(), &mut st) };
        assert_eq!(rc, 0, "qq_msm failed");
        match st {
            0 => Some(out),
            ST_BAD_POINT => None, // optional_multiscalar_mul's None
            ST_BAD_SCALAR => panic!("qq_msm: non-canonical scalar (the reference never builds one)"),
            other => panic!("qq_msm: unexpected status {}", other),
        }
    }
}
impl Drop for Gpu {
    fn drop(&mut self) { unsafe { qq_destroy(self.0) } }
}
