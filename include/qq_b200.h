/* qq_b200.h -- C ABI of libqq_b200.so: the B200 (sm_100a) implementation of quisquis-rust's data-parallel hot path
 * (batched Ristretto255 scalar multiplication behind account updates / ElGamal commitments, and multiscalar
 * multiplication under proof verification).
 *
 * The reference has no FFI boundary (`#![deny(unsafe_code)]`, reference src/lib.rs:4); the seam is its public Rust API.
 * Every entry point below names the reference function it batches (file:line under /root/reference).  A Rust
 * `cuda/` sys crate binds these symbols; tests and bench.py bind them with ctypes.
 *
 * Conventions
 *   - All arrays are caller-owned and tightly packed (AoS exactly as the reference serialises its types):
 *       CompressedRistretto 32 B; Scalar 32 B little-endian; RistrettoPublicKey = gr||grsk 64 B
 *       (src/ristretto/keys.rs:113-120); ElGamalCommitment = c||d 64 B (src/elgamal/elgamal.rs:135-142);
 *       Account = pk||comm 128 B (src/accounts/accounts.rs:47-53).
 *   - Functions without a suffix take HOST pointers: the library stages host->device, runs the kernels and copies
 *     results back before returning (synchronous).  `_dev` variants take DEVICE pointers on the ctx's GPU and only
 *     enqueue + synchronise the kernels.
 *   - Return value: QQ_OK or a negative QQ_ERR_* (API-level failure: bad argument, CUDA error, out of memory).
 *   - status[i] (one byte per element): see QQ_ST_*.  Where the reference would panic (`.unwrap()` on a failed
 *     decompress, e.g. src/ristretto/keys.rs:278-279, src/elgamal/elgamal.rs:47,50,66-67) or return
 *     Err("Error::Decompression Failed") / None, status[i] = QQ_ST_BAD_POINT and the outputs for element i are zero.
 *   - One qq_ctx per GPU and per calling thread (or external locking).  There is NO CPU fallback: qq_init fails
 *     when no sm_100 device is usable.
 *   - The batched verifiers (qq_verify_*_batch) run their Fiat-Shamir transcripts on std::thread workers (one per host
 *     core, created and joined inside the call); qq_verify_shuffle_batch and qq_verify_range_proof_batch also drive the GPU
 *     from one internal worker thread while the caller's thread prepares the next slice.  The calls stay synchronous.
 */
#ifndef QQ_B200_H
#define QQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QQ_OK 0
#define QQ_ERR_ARG (-1)
#define QQ_ERR_CUDA (-2)
#define QQ_ERR_NOMEM (-3)
#define QQ_ERR_NODEVICE (-4)
#define QQ_ERR_INTERNAL (-5) /* an internal failure (a C++ exception caught at the boundary); qq_last_error has the text */

#define QQ_ST_OK 0
#define QQ_ST_BAD_POINT 1   /* a compressed point failed RFC 9496 decoding */
#define QQ_ST_BAD_SCALAR 2  /* a scalar was not canonical (>= l) */
#define QQ_ST_KEYPAIR 3     /* Err("Invalid Account::Keypair Verification Failed") / `false` */
#define QQ_ST_COMMIT 4      /* Err("Invalid Account::Commitment Verification Failed") / identity check failed */
#define QQ_ST_PROOF 6       /* a sigma-protocol challenge did not match: Err("DLOG Proof Verify: Failed") and its siblings */
#define QQ_ST_NOT_FOUND 5   /* decommit_value: no v below 2^search_bits (the reference would keep searching up to 2^64) */
#define QQ_ST_PANIC 7       /* an undecodable point where the reference `unwrap()`s (panics) although a neighbouring check of the same function returns Err */

#define QQ_BASE_B 0 /* Ristretto basepoint, BASE_PK_BTC_COMPRESSED[0] (src/ristretto/constants.rs:13-16) */
#define QQ_BASE_H 1 /* Pedersen H,         BASE_PK_BTC_COMPRESSED[1] (src/ristretto/constants.rs:17-20) */

typedef struct qq_ctx qq_ctx;

/* ---- context ------------------------------------------------------------------------------------------------- */
int qq_init(qq_ctx** ctx, int device);
void qq_destroy(qq_ctx* ctx);
const char* qq_last_error(const qq_ctx* ctx);
int qq_device_sm_count(const qq_ctx* ctx);
/* number of kernels launched by this ctx since creation (bench.py reports it as gpu_launches) */
uint64_t qq_launch_count(const qq_ctx* ctx);
/* milliseconds spent in kernels during the most recent call, measured with CUDA events on the ctx stream */
float qq_last_kernel_ms(const qq_ctx* ctx);
/* per-kernel-family milliseconds of the most recent call: [0] decompress [1] variable-base [2] fixed-base
 * [3] finish/compress [4] msm-bucket [5] msm-reduce [6] transcript kernels; returns number of entries written */
int qq_last_kernel_breakdown(const qq_ctx* ctx, float* ms, int cap);
/* CUDA events on the ctx stream (the stream every kernel of this ctx is launched on), so callers can time a
 * sequence of _dev calls on the device: record slot a, run, record slot b, read elapsed.  slots 0..7 */
int qq_event_record(qq_ctx* ctx, int slot);
int qq_event_elapsed_ms(qq_ctx* ctx, int slot_a, int slot_b, float* ms);
/* device memory helpers for the _dev entry points (thin cudaMalloc/cudaMemcpy wrappers so callers need no CUDA) */
int qq_dev_alloc(qq_ctx* ctx, void** dptr, size_t bytes);
int qq_dev_free(qq_ctx* ctx, void* dptr);
int qq_dev_upload(qq_ctx* ctx, void* dptr, const void* host, size_t bytes);
int qq_dev_download(qq_ctx* ctx, void* host, const void* dptr, size_t bytes);
/* integer-pipe micro-benchmark: thread-level IMAD.WIDE-class ops per second on this device (roofline denominator) */
int qq_measure_imad_peak(qq_ctx* ctx, double* wide_ops_per_s, double* lo_ops_per_s);

/* ---- RistrettoPublicKey ----------------------------------------------------------------------------------------
 * update_public_key(p, r) = (r*gr, r*grsk)              reference src/ristretto/keys.rs:146-148 (Mul :266-282) */
int qq_update_public_key_batch(qq_ctx* ctx, const uint8_t* pk, const uint8_t* r, uint8_t* out_pk, uint8_t* status,
                               size_t n);
int qq_update_public_key_batch_dev(qq_ctx* ctx, const uint8_t* pk, const uint8_t* r, uint8_t* out_pk,
                                   uint8_t* status, size_t n);
/* verify_public_key_update(u, p, r): status 0 = true, QQ_ST_KEYPAIR = false   src/ristretto/keys.rs:161-169 */
int qq_verify_public_key_update_batch(qq_ctx* ctx, const uint8_t* updated_pk, const uint8_t* pk, const uint8_t* r,
                                      uint8_t* status, size_t n);

/* ---- ElGamalCommitment -----------------------------------------------------------------------------------------
 * generate_commitment(p, r, v) = (r*gr, v*B + r*grsk)                         src/elgamal/elgamal.rs:41-53 */
int qq_generate_commitment_batch(qq_ctx* ctx, const uint8_t* pk, const uint8_t* r, const uint8_t* v,
                                 uint8_t* out_comm, uint8_t* status, size_t n);
int qq_generate_commitment_batch_dev(qq_ctx* ctx, const uint8_t* pk, const uint8_t* r, const uint8_t* v,
                                     uint8_t* out_comm, uint8_t* status, size_t n);
/* add_commitments(a, b)                                                       src/elgamal/elgamal.rs:65-69
 * impl Sub (negate = 1)                                                       src/elgamal/elgamal.rs:201-218 */
int qq_add_commitments_batch(qq_ctx* ctx, const uint8_t* a, const uint8_t* b, int negate_b, uint8_t* out_comm,
                             uint8_t* status, size_t n);
/* impl Mul<&Scalar> for &ElGamalCommitment                                    src/elgamal/elgamal.rs:220-236 */
int qq_mul_commitment_batch(qq_ctx* ctx, const uint8_t* comm, const uint8_t* s, uint8_t* out_comm, uint8_t* status,
                            size_t n);

/* ---- Account ---------------------------------------------------------------------------------------------------
 * update_account(a, bl, u, c): pk' = u*pk ; comm' = generate_commitment(OLD pk, c, bl) + a.comm
 *                                                                             src/accounts/accounts.rs:143-154 */
int qq_update_account_batch(qq_ctx* ctx, const uint8_t* acc, const uint8_t* bl, const uint8_t* u, const uint8_t* c,
                            uint8_t* out_acc, uint8_t* status, size_t n);
int qq_update_account_batch_dev(qq_ctx* ctx, const uint8_t* acc, const uint8_t* bl, const uint8_t* u,
                                const uint8_t* c, uint8_t* out_acc, uint8_t* status, size_t n);
/* verify_account(sk, bl): status 0 / QQ_ST_KEYPAIR / QQ_ST_COMMIT / QQ_ST_BAD_POINT
 *                                                                             src/accounts/accounts.rs:81-84 */
int qq_verify_account_batch(qq_ctx* ctx, const uint8_t* acc, const uint8_t* sk, const uint8_t* bl, uint8_t* status,
                            size_t n);
int qq_verify_account_batch_dev(qq_ctx* ctx, const uint8_t* acc, const uint8_t* sk, const uint8_t* bl,
                                uint8_t* status, size_t n);
/* create_delta_and_epsilon_accounts(a, bl, base_pk) with the random scalars r supplied by the caller (the reference
 * draws them from OsRng inside the function, src/accounts/accounts.rs:203,318-326):
 *   delta_i   = (a_i.pk,  generate_commitment(a_i.pk,  r_i, bl_i))
 *   epsilon_i = (base_pk, generate_commitment(base_pk, r_i, bl_i))           src/accounts/accounts.rs:198-220
 * base_pk must be BASE_PK_BTC_COMPRESSED (B, H) = RistrettoPublicKey::generate_base_pk() (src/ristretto/keys.rs:171-177),
 * the only base key the reference defines; the fixed-base tables are built for it and any other value returns
 * QQ_ERR_ARG. */
int qq_delta_epsilon_batch(qq_ctx* ctx, const uint8_t* acc, const uint8_t* bl, const uint8_t* r,
                           const uint8_t* base_pk, uint8_t* out_delta, uint8_t* out_epsilon, uint8_t* status,
                           size_t n);
/* Verifier::verify_delta_identity_check: sum of c and sum of d over all n accounts must be the identity.
 * *verdict = QQ_ST_OK / QQ_ST_COMMIT / QQ_ST_BAD_POINT                        src/accounts/verifier.rs:566-581 */
int qq_delta_identity_check(qq_ctx* ctx, const uint8_t* acc, size_t n, uint8_t* verdict);

/* ---- fixed-base multiplication: out_i = s_i * Base, Base in {QQ_BASE_B, QQ_BASE_H}
 * `&Scalar * &RISTRETTO_BASEPOINT_TABLE` (src/elgamal/elgamal.rs:49,85; src/ristretto/keys.rs:102) and the B/H legs of
 * PedersenGens::commit used throughout src/shuffle */
int qq_fixed_base_batch(qq_ctx* ctx, int which, const uint8_t* s, uint8_t* out_points, uint8_t* status, size_t n);
int qq_fixed_base_batch_dev(qq_ctx* ctx, int which, const uint8_t* s, uint8_t* out_points, uint8_t* status,
                            size_t n);
/* Window width of the large fixed-base table of base `which` (dalek's counterpart is the fixed radix-16
 * RistrettoBasepointTable, constants.rs / edwards.rs::EdwardsBasepointTable; the reference uses it through
 * src/elgamal/elgamal.rs:49).  Tables hold ceil(255 / W) x (2^(W-1) + 1) affine points of 96 bytes in device memory:
 * W = 16 -> 50 MB (default at qq_init, L2 resident), 22 -> 2.4 GB, 24 -> 8.9 GB, 26 -> 32 GB.  A wider window means
 * fewer additions per scalar and identical results.  window_bits = 0 frees the table (batches then use the
 * shared-memory 6-bit table only); otherwise 8 <= window_bits <= 28.  Rebuilds synchronously. */
int qq_fixed_base_set_window(qq_ctx* ctx, int which, int window_bits);
/* Tuning: variable-base calls of at most max_scalar_mults scalar multiplications (9-account anonymity sets, a block's
 * worth of transactions) run four lanes per multiplication (k_varbase_coop: latency of 2 instead of 8 field products
 * per group operation); larger calls run one thread per point.  < 0 restores the default (160 per SM), 0 disables (and
 * with it the other small-batch paths, for A/B measurements). */
int qq_varbase_set_coop_limit(qq_ctx* ctx, long max_scalar_mults);
/* out_i = enc(v_i * Base) for signed 64-bit values: what `&Scalar::from(v as u64) * &RISTRETTO_BASEPOINT_TABLE` (and its
 * negation) computes for balances (src/elgamal/elgamal.rs:285-300, src/accounts/accounts.rs:419-429).  Only the windows
 * that a 64-bit magnitude can touch are walked (3 at W = 22).  Requires a large-window table (window_bits != 0). */
int qq_fixed_base_i64_batch(qq_ctx* ctx, int which, const int64_t* v, uint8_t* out_points, size_t n);
int qq_fixed_base_i64_batch_dev(qq_ctx* ctx, int which, const int64_t* v, uint8_t* out_points, size_t n);
int qq_fixed_base_window(const qq_ctx* ctx, int which);

/* ---- secret scalars ----------------------------------------------------------------------------------
 * The wallet side of the path handles secrets: u, c, bl of update_account, r of update_public_key / generate_commitment /
 * create_delta_and_epsilon_accounts, sk of verify_account / decommit, and the blindings of the provers (src/accounts/prover.rs).
 * curve25519-dalek reads its window tables with LookupTable::select (every entry touched, masked) for them.
 * qq_set_secret_mode(ctx, 1): the variable-base kernels scan all nine entries of a window table and keep one with masks,
 * fixed-base batches use the shared-memory 6-bit table with the same masked scan of a window's 33 entries (the large-window
 * tables in L2 / HBM and the four-lane small-batch kernels, whose reads are addressed by the digit, are not used), 64-bit
 * values are expanded to scalars first.  Control flow was already independent of the scalars in either mode.  Entry points
 * covered: qq_update_public_key_batch, qq_mul_commitment_batch, qq_generate_commitment_batch, qq_update_account_batch,
 * qq_delta_epsilon_batch, qq_verify_account_batch, qq_decommit_batch, qq_fixed_base_batch, qq_fixed_base_i64_batch,
 * qq_sigma_commit_batch (and their _dev forms).  Outputs are byte-identical in both modes.  The verifiers and the MSMs work on
 * public data and ignore the mode.  Default off (QQ_SECRET_MODE=1 in the environment turns it on at qq_init). */
int qq_set_secret_mode(qq_ctx* ctx, int on);
int qq_secret_mode(const qq_ctx* ctx);
/* Prover-side sigma-protocol commitments: out_i = enc(r_i * dec(P_i)) when v == NULL, enc(v_i * B + r_i * dec(P_i)) otherwise -
 * the e / f maps of Prover::verify_delta_compact_prover (src/accounts/prover.rs:164-207), verify_account_prover (:415-454),
 * zero_balance_account_vector_prover (:629-640), destroy_account_prover (:742-753), verify_update_account_dark_tx_prover
 * (:885-921).  points, r, v, out_points: n x 32 B.  status as for qq_generate_commitment_batch. */
int qq_sigma_commit_batch(qq_ctx* ctx, const uint8_t* points, const uint8_t* r, const uint8_t* v, uint8_t* out_points,
                          uint8_t* status, size_t n);
int qq_sigma_commit_batch_dev(qq_ctx* ctx, const uint8_t* points, const uint8_t* r, const uint8_t* v, uint8_t* out_points,
                              uint8_t* status, size_t n);

/* ---- multiscalar multiplication --------------------------------------------------------------------------------
 * Verifier::multiscalar_multiplication = RistrettoPoint::optional_multiscalar_mul over compressed points
 * (src/accounts/verifier.rs:91-99) and the Bulletproofs verification mega-MSM reached from
 * src/accounts/verifier.rs:517,548.  out = sum_i s_i * decompress(P_i) compressed; *status = QQ_ST_BAD_POINT when any
 * point fails to decode (the reference returns None), QQ_ST_BAD_SCALAR for a non-canonical scalar. */
int qq_msm(qq_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t* out_point,
           uint8_t* status);
int qq_msm_dev(qq_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t* out_point,
               uint8_t* status);
/* same sum, but returns the uncompressed partial result as 4 x 32 canonical bytes (X, Y, Z, T) so that per-GPU
 * partial sums can be gathered (NCCL all-gather, 128 B per rank) and combined with qq_points_sum. */
int qq_msm_partial(qq_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t* out_xyzt,
                   uint8_t* status);
int qq_msm_partial_dev(qq_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t* out_xyzt,
                       uint8_t* status);
/* Point sets that are reused across many MSMs -- the Bulletproofs generators G_i, H_i, which the reference rebuilds
 * with BulletproofGens::new(64, 16) on every range-proof verification (src/accounts/verifier.rs:510,540), the Pedersen
 * bases of src/pedersen/vectorpedersen.rs:45-75 -- can be decompressed ONCE into the library's device-resident MSM
 * form (96 bytes per point + validity).  qq_msm_prepared then skips the per-point inverse square root, which is 62 %
 * of the work of qq_msm.  An invalid point in the set makes every MSM that touches it return QQ_ST_BAD_POINT. */
typedef struct qq_prepared qq_prepared;
int qq_msm_points_prepare(qq_ctx* ctx, const uint8_t* points, size_t n, qq_prepared** out);
int qq_msm_points_prepare_dev(qq_ctx* ctx, const uint8_t* points, size_t n, qq_prepared** out);
void qq_msm_points_free(qq_ctx* ctx, qq_prepared* p);
size_t qq_msm_points_count(const qq_prepared* p);
/* out = sum_{i<n} s_i * P_i over the first n points of the prepared set (n <= qq_msm_points_count) */
int qq_msm_prepared(qq_ctx* ctx, const uint8_t* scalars, const qq_prepared* points, size_t n, uint8_t* out_point,
                    uint8_t* status);
int qq_msm_prepared_dev(qq_ctx* ctx, const uint8_t* scalars, const qq_prepared* points, size_t n, uint8_t* out_point,
                        uint8_t* status);
/* Shifted form of prepared point sets: qq_msm_points_prepare also stores 2^(c k) P_i for every window k (c fixed by the size of
 * the set; ceil(256 / c) x 96 B per point) when that fits budget_bytes (default 112 MB = sets up to 2^16 points, whose shifted
 * form stays L2-resident; 0 = never; sets below 1 024 points never).  qq_msm_prepared then drops every digit of every scalar
 * into ONE set of 2^(c-1) buckets: no per-window reductions and no chain of 240 doublings at the end (measured 0.44 against
 * 0.62 ms for 2^10 .. 2^12 points - the Bulletproofs generator set -, 0.61 against 0.76 ms at 2^16; beyond L2 the larger gather
 * footprint loses: 3.3 against 2.35 ms at 2^20).  use_it = 0: ignore shifted forms that exist (A/B measurements).  Results
 * are identical. */
int qq_msm_set_shifted(qq_ctx* ctx, size_t budget_bytes, int use_it);
size_t qq_msm_points_shifted_bytes(const qq_prepared* p);
/* Tuning of the large MSM (compressed points, n >= split_min): the last tail_pct % of the points are decompressed by a
 * second kernel while the counting sort of the digits runs beside it on a high-priority stream with sort_blocks_per_sm
 * blocks per SM.  Defaults 2^17, 30, 3 (measured: 2^20 points 4.18 -> 3.99 ms, 2^24 58.0 -> 54.4 ms); tail_pct 0 turns the
 * overlap off.  Results do not depend on it. */
int qq_msm_set_overlap(qq_ctx* ctx, long split_min, int tail_pct, int sort_blocks_per_sm);
/* sum of k extended points given as k x 128 B (X,Y,Z,T canonical) -> compressed; *is_identity set to 1/0 */
int qq_points_sum(qq_ctx* ctx, const uint8_t* xyzt, size_t k, uint8_t* out_point, uint8_t* is_identity);
/* Parity hook for the warp-cooperative group operations (csrc/ge_warp.cuh: one point per warp, one 32-bit limb per lane; used
 * by the window Horner chain of the large MSM): item j = two extended points as 4 x 8 little-endian 32-bit limbs each (ANY
 * limb values, i.e. field elements in the saturated form, not necessarily points of the curve - the formulas are polynomial);
 * out_dbl[j] = the doubling formula applied to p[j], out_add[j] = the addition formula applied to p[j], q[j], as 128 B of
 * limbs (values mod p).  Follows the same formulas as curve25519-dalek's curve_models (ProjectivePoint::double,
 * EdwardsPoint + ProjectiveNielsPoint), which the reference reaches through every point operation (src/ristretto/keys.rs:277-281). */
int qq_warp_ops_selftest(qq_ctx* ctx, const uint8_t* p_xyzt, const uint8_t* q_xyzt, size_t n, uint8_t* out_dbl, uint8_t* out_add);
/* many small MSMs (2..9 terms each in the reference, src/accounts/verifier.rs:165-880, src/shuffle/*):
 * instance j covers terms offsets[j] .. offsets[j+1]-1 ; out m x 32 B ; status m */
int qq_msm_segmented(qq_ctx* ctx, const uint8_t* scalars, const uint8_t* points, const uint32_t* offsets, size_t m,
                     uint8_t* out_points, uint8_t* status);
/* m independent LARGE MSMs (hundreds to millions of terms each; m <= 4096) in ONE Pippenger pass: MSM j covers terms
 * offsets[j] .. offsets[j+1]-1 (offsets[0] = 0), out m x 32 B, status m (`optional_multiscalar_mul` semantics per MSM: a bad
 * term fails its own MSM only).  The windows of MSM j are the windows [j K, (j + 1) K) of one bucket array, so sorting, bucket
 * accumulation and the running-sum reduction are one launch each for all MSMs and the m Horner chains run side by side.
 * This is the per-proof form of Bulletproofs verification (RangeProof::verify_multiple is ONE MSM per proof,
 * src/accounts/verifier.rs:517) for a batch of proofs, and what the batched verifiers use to locate the failing proofs when a
 * randomised aggregate does not verify. */
int qq_msm_grouped(qq_ctx* ctx, const uint8_t* scalars, const uint8_t* points, const uint32_t* offsets, size_t m,
                   uint8_t* out_points, uint8_t* status);

/* ---- sigma-protocol verification ----------------------------------------------------------------------------------
 * Verifier::verify_update_account_verifier (src/accounts/verifier.rs:223-292) for `nproofs` independent proofs over n
 * accounts each.  input_accounts / delta_accounts: nproofs x n x 128 B (updated_input_accounts, updated_delta_accounts),
 * z: nproofs x n x 32 B, x: nproofs x 32 B.  Every proof is checked under Transcript::new(transcript_label) followed
 * by Verifier::new(verifier_label, ..) (the reference's test uses b"UpdateAccount", b"DLOGProof", verifier.rs:1049-1072).
 * status[p]: QQ_ST_OK = Ok(()), QQ_ST_PROOF = Err("DLOG Proof Verify: Failed"), QQ_ST_BAD_POINT where the reference
 * panics / returns that Err on an undecodable point, QQ_ST_BAD_SCALAR for a non-canonical z or x.
 * The 2 n commitments of every proof are computed on the GPU as 3-term MSMs; the Fiat-Shamir transcript (Merlin,
 * STROBE-128) runs on the host. */
int qq_verify_update_account_dlog_batch(qq_ctx* ctx, const char* transcript_label, const char* verifier_label,
                                        const uint8_t* input_accounts, const uint8_t* delta_accounts, const uint8_t* z,
                                        const uint8_t* x, size_t n, size_t nproofs, uint8_t* status);

/* Verifier::verify_delta_compact_verifier (src/accounts/verifier.rs:138-209; the reference's test uses the labels
 * b"DeltaCompact", b"DLEQProof", verifier.rs:978-1002): delta and epsilon accounts commit to the same values.
 * delta_accounts / epsilon_accounts: nproofs x n x 128 B; zv, zr1, zr2: nproofs x n x 32 B; x: nproofs x 32 B.
 * status[p]: QQ_ST_OK = Ok(()), QQ_ST_PROOF = Err("Dleq Proof Verify: Failed"), QQ_ST_BAD_POINT =
 * Err("Delta Compact Proof Verify: Failed"), QQ_ST_BAD_SCALAR for a non-canonical response or challenge. */
int qq_verify_delta_compact_batch(qq_ctx* ctx, const char* transcript_label, const char* verifier_label,
                                  const uint8_t* delta_accounts, const uint8_t* epsilon_accounts, const uint8_t* zv,
                                  const uint8_t* zr1, const uint8_t* zr2, const uint8_t* x, size_t n, size_t nproofs,
                                  uint8_t* status);

/* Verifier::verify_account_verifier_bulletproof (src/accounts/verifier.rs:396-470) = the sigma-protocol part of
 * Verifier::verify_account_verifier (:305-381, whose R1CS range proof is outside this path): the senders know their
 * secret keys and delta / epsilon accounts hold the same balance.  delta_accounts (updated_delta_account_sender),
 * epsilon_accounts (account_epsilon_sender): nproofs x n x 128 B; base_pk: 64 B; zv, zsk, zr: nproofs x n x 32 B;
 * x: nproofs x 32 B.  status[p]: QQ_ST_OK = Ok(()), QQ_ST_PROOF = Err("sender account verification failed"),
 * QQ_ST_BAD_POINT = Err("Account Verify: Failed"), QQ_ST_BAD_SCALAR for a non-canonical response or challenge. */
int qq_verify_account_sigma_batch(qq_ctx* ctx, const char* transcript_label, const char* verifier_label,
                                  const uint8_t* delta_accounts, const uint8_t* epsilon_accounts, const uint8_t* base_pk,
                                  const uint8_t* zv, const uint8_t* zsk, const uint8_t* zr, const uint8_t* x, size_t n,
                                  size_t nproofs, uint8_t* status);

/* Verifier::zero_balance_account_vector_verifier (src/accounts/verifier.rs:593-634) when vector_form != 0,
 * Verifier::zero_balance_account_verifier (:647-680) when vector_form == 0 (n must be 1).  accounts: nproofs x n x 128 B,
 * z: nproofs x n x 32 B, x: nproofs x 32 B.  The vector form keeps the reference verifier's domain separator
 * b"ZeroBalanceAccounVectorProof" (:605; the prover writes b"ZeroBalanceAccountVectorProof", prover.rs:613, so the
 * reference's verifier rejects its own prover's proofs - and so does this one).  status[p]: QQ_ST_OK,
 * QQ_ST_PROOF = Err("Zero balance account verification failed"), QQ_ST_BAD_POINT = Err("Zero balance Account Verify: Failed"). */
int qq_verify_zero_balance_batch(qq_ctx* ctx, const char* transcript_label, const char* verifier_label,
                                 const uint8_t* accounts, const uint8_t* z, const uint8_t* x, size_t n, size_t nproofs,
                                 int vector_form, uint8_t* status);

/* Verifier::destroy_account_verifier (src/accounts/verifier.rs:693-735).  accounts: nproofs x n x 128 B, z: nproofs x n
 * x 32 B, x: nproofs x 32 B.  status[p]: QQ_ST_OK, QQ_ST_PROOF = Err("Destroy account verification failed"),
 * QQ_ST_BAD_POINT = Err("Destroy Account Verify: Failed"). */
int qq_verify_destroy_account_batch(qq_ctx* ctx, const char* transcript_label, const char* verifier_label,
                                    const uint8_t* accounts, const uint8_t* z, const uint8_t* x, size_t n, size_t nproofs,
                                    uint8_t* status);

/* Verifier::verify_same_value_compact_verifier (src/accounts/verifier.rs:747-806), one proof per element: enc_accounts
 * nproofs x 128 B, commitments (Pedersen, PedersenGens::default()) nproofs x 32 B, zv / zr / x: nproofs x 32 B
 * (SigmaProof::Dleq(zv[0], zr[0], _, x)).  The transcript labels are fixed inside the reference function.
 * status[p]: QQ_ST_OK, QQ_ST_PROOF = Err("Same Value Proof Verify: Failed"), QQ_ST_BAD_POINT =
 * Err("Delta Compact Proof Verify: Failed"). */
int qq_verify_same_value_compact_batch(qq_ctx* ctx, const uint8_t* enc_accounts, const uint8_t* commitments,
                                       const uint8_t* zv, const uint8_t* zr, const uint8_t* x, size_t nproofs,
                                       uint8_t* status);

/* Verifier::verify_update_account_dark_tx_verifier (src/accounts/verifier.rs:818-917; the reference's test uses the
 * labels b"UpdateAccount", b"DLOGProof", :1075-1111).  delta_accounts (delta_updated_accounts), output_accounts:
 * nproofs x n x 128 B; z: nproofs x 2 x 32 B (z_vector[0], z_vector[1]); x: nproofs x 32 B.  status[p]: QQ_ST_OK,
 * QQ_ST_PROOF = Err("Update Output Challenge : DLOG Proof Verify: Failed"), QQ_ST_BAD_POINT = Err("Update Account: DLOG
 * Proof Verify: Failed") (undecodable key), QQ_ST_PANIC where the reference panics (undecodable commitment in
 * `d.comm - i.comm`, :862-866). */
int qq_verify_update_account_dark_tx_batch(qq_ctx* ctx, const char* transcript_label, const char* verifier_label,
                                           const uint8_t* delta_accounts, const uint8_t* output_accounts, const uint8_t* z,
                                           const uint8_t* x, size_t n, size_t nproofs, uint8_t* status);

/* ---- leaf arguments of the Bayer-Groth shuffle proof (src/shuffle, ROWS = COLUMNS = 3) ---------------------------------
 * Per proof: Merlin transcript and scalar algebra on the host threads, every group equation as one MSM that must be the
 * identity, all MSMs of all proofs in one GPU batch.  transcript_label / verifier_label: the labels the caller used for
 * Transcript::new / Verifier::new (the sub-proofs of a ShuffleProof share one running transcript in the reference; these
 * entry points verify stand-alone arguments, as the reference's own unit tests do).
 *
 * DDHProof::verify_ddh_proof (src/shuffle/ddh.rs:109-142): all arrays nproofs x 32 B.  status[p]: QQ_ST_OK, QQ_ST_PROOF or
 * QQ_ST_BAD_POINT (both Err("DDH Proof Verify: Failed")), QQ_ST_BAD_SCALAR. */
int qq_verify_ddh_batch(qq_ctx* ctx, const char* transcript_label, const char* verifier_label, const uint8_t* g,
                        const uint8_t* h, const uint8_t* g_dash, const uint8_t* h_dash, const uint8_t* challenge,
                        const uint8_t* z, size_t nproofs, uint8_t* status);
/* SVPProof::verify (src/shuffle/singlevalueproduct.rs:175-257).  commitment_a, b (SVPStatement): nproofs x 32 B; proof:
 * nproofs x 352 B = commitment_d | commitment_delta_small | commitment_delta_capital | a_twildle[3] | b_twildle[3] |
 * r_twildle | s_twildle.  status[p]: QQ_ST_OK, QQ_ST_PROOF = Err("SingleValue Product Proof Verify: Failed"),
 * QQ_ST_BAD_POINT = Err("SingleValue Product Proof Verify: Decompression Failed"), QQ_ST_BAD_SCALAR. */
int qq_verify_svp_batch(qq_ctx* ctx, const char* transcript_label, const char* verifier_label, const uint8_t* commitment_a,
                        const uint8_t* b, const uint8_t* proof, size_t nproofs, uint8_t* status);
/* HadamardProof::verify (src/shuffle/hadamard.rs:249-389).  omega (HadamardStatement), commit_a, commit_b, commit_c:
 * nproofs x 96 B; proof: nproofs x 640 B = commitment_a_0 | commitment_b_0 | commitment_c_0 | commitment_delta[4] | a_bar[3]
 * | b_bar[3] | c_bar[3] | r_bar | s_bar | t_bar | rho_bar.  status[p]: QQ_ST_OK, QQ_ST_BAD_POINT = Err("HadamardProof Verify:
 * Decompression Failed"), QQ_ST_BAD_SCALAR, QQ_ST_PROOF with detail[p] (detail may be NULL): 1 = "Omega values are not
 * unique", 2 = "A_bar , B_bar, C_bar check failed", 3 = "Delta Commitment check failed". */
int qq_verify_hadamard_batch(qq_ctx* ctx, const char* transcript_label, const char* verifier_label, const uint8_t* omega,
                             const uint8_t* commit_a, const uint8_t* commit_b, const uint8_t* commit_c, const uint8_t* proof,
                             size_t nproofs, uint8_t* status, uint8_t* detail);
/* ProductProof::verify (src/shuffle/product.rs:170-195: MultiHadamardProof::verify :325-389, ZeroProof::verify :508-600,
 * SVPProof::verify) on one running transcript.  c_prod_A: nproofs x 96 B (the three column commitments); statement: nproofs x
 * 192 B = c_b | zero_statement.c_A[3] | svp commitment_a | svp b; proof: nproofs x 1024 B = c_B[3] | c_A_0 | c_B_m | c_D[7] |
 * a_vec[3] | b_vec[3] | r | s | t | SVPProof (352 B).  status[p]: QQ_ST_OK, QQ_ST_BAD_SCALAR, QQ_ST_BAD_POINT (detail 10 =
 * "Multihadamard Proof Verify: Failed", 11 = "ZeroProof Verify: Decompression Failed" / "ZeroProof Verify: Failed", 12 =
 * the SVP's), QQ_ST_PROOF with detail 1 = "Multihadamard Product Proof Verify: c_B_1 == c_A_1 Failed", 2 = "... c_B_m == c_b
 * Failed", 3..6 = "Zero Argument Proof Verify: c_d_(m+1) == com(0,0) / com(a_bar, r) / com(b_bar, s) / com(a_bar * b_bar, t)
 * ... Failed", 7 = "SingleValue Product Proof Verify: Failed".  detail may be NULL. */
int qq_verify_product_batch(qq_ctx* ctx, const char* transcript_label, const char* verifier_label, const uint8_t* c_prod_A,
                            const uint8_t* statement, const uint8_t* proof, size_t nproofs, uint8_t* status, uint8_t* detail);
/* ShuffleProof::verify (src/shuffle/shuffle.rs:547-712): the whole shuffle argument over 9 accounts, nproofs independent
 * proofs in two GPU round trips (Hadamard + product arguments + key aggregates, then DDH check + the two multi-exponentiation
 * arguments).  shuffle_input / shuffle_output: nproofs x 9 x 128 B.  statement: nproofs x 352 B = HadamardStatement omega[3] |
 * ProductStatement (192 B, as above) | DDHStatement G_dash | H_dash.  proof: nproofs x 3776 B = c_A[3] | c_tau[3] | c_B[3] |
 * c_B_dash[3] | HadamardProof (640 B) | ProductProof (1024 B) | multi_exponen_pk | multi_exponen_commit (MultiexpoProof, 832 B
 * each: c_A_0 | c_B_k[6] | E_k_0[6] | E_k_1[6] | a_vec[3] | r | b | s | t) | DDHProof challenge | z - the structs' fields in
 * declaration order, vectors without their bincode length prefix.  status[p]: QQ_ST_OK = Ok(()), QQ_ST_PROOF (a check
 * failed), QQ_ST_BAD_POINT (an Err caused by an undecodable point), QQ_ST_BAD_SCALAR.  stage[p] / detail[p] (may be NULL): 1
 * Hadamard (detail as qq_verify_hadamard_batch), 2 "Shuffle Proof Verify:prod pf i .. N (yi + x^i -z) failed", 3 "ShuffleProof
 * Verify: Decompression Failed" (c_A / c_B), 4 product argument (detail as qq_verify_product_batch), 5 the same message for an
 * input key, 6 "DDH Proof Verify: Failed", 7 / 8 "Multi-exponentiation Pubkey / Commitment Argument: ..." with detail 1 "Verify
 * com(0,0) == c_B_m Failed", 2 "Verify Em == C Failed", 3 "a Scalar vector Verification Failed", 4 "Scalar b Verification Failed",
 * 5 "E_K Verification Failed". */
int qq_verify_shuffle_batch(qq_ctx* ctx, const char* transcript_label, const char* verifier_label, const uint8_t* shuffle_input,
                            const uint8_t* shuffle_output, const uint8_t* statement, const uint8_t* proof, size_t nproofs,
                            uint8_t* status, uint8_t* stage, uint8_t* detail);
/* Where the batched verifiers (qq_verify_shuffle_batch, qq_verify_range_proof_batch, the sigma verifiers qq_verify_*_batch of
 * src/accounts/verifier.rs and qq_verify_ddh_batch) run the per-proof Fiat-Shamir transcripts (Merlin,
 * src/accounts/transcript.rs:55-82) and the Z/l algebra: on_device != 0 (default) in transcript kernels, one GPU thread per
 * proof - the proof bytes are the only upload, job lists, transcripts and verdicts stay in device memory; on_device == 0 on
 * the host threads, job lists uploaded per batch (the round-1 arrangement, kept for A/B measurements; the sigma verifiers also
 * take it for batches of at most 4 proofs, where a serial GPU transcript costs more than it saves).  Verdicts are identical. */
int qq_verify_set_transcripts(qq_ctx* ctx, int on_device);
/* Aggregate form of the device-resident shuffle verifier (default on).  on: per proof only G, H, g_r, h_r (the MSM results
 * the transcript absorbs) are evaluated on their own; the other 28 group equations of every proof are multiplied by fresh
 * random 128-bit weights in the transcript kernels and decided together by ONE Pippenger MSM over the whole batch (146 terms
 * per proof + the six fixed generators).  Proofs that fail a scalar-level check, and the whole batch when the aggregate is
 * not the identity, are verified in the exact form (on == 0: every equation an MSM of its own), which alone reports
 * (status, stage, detail); an invalid proof is accepted with probability <= 2^-128.  Verdicts of valid proofs are identical.
 * When the aggregate is not the identity, the weighted sums of groups of 64 proofs are evaluated by one grouped MSM
 * (qq_msm_grouped's machinery) and only the proofs of failing groups go to the exact form.
 * qq_verify_range_proof_batch follows the same switch: on (default) one weighted MSM over the whole batch, narrowed by grouped
 * MSMs (64 groups, failing groups cut in eight, ...) when it fails; off: every transcript's own verification MSM (the
 * reference's per-proof form, 2 n m + 2 + T terms each) as one group of a grouped MSM, 1 024 transcripts per pass. */
int qq_verify_set_aggregation(qq_ctx* ctx, int on);

/* ---- Bulletproofs range proofs (BASELINE configs[3]) -----------------------------------------------------------------
 * RangeProof::verify_multiple / verify_single of the `bulletproofs` crate, as called by
 *   Verifier::verify_non_negative_sender_receiver_bulletproof_batch_verifier   src/accounts/verifier.rs:504-523
 *       (domain_label "AggregateBulletProof", n_bits 64, m = epsilon_account.len() <= 16 a power of two, chain 1)
 *   Verifier::verify_non_negative_sender_receiver_bulletproof_vector_verifier  src/accounts/verifier.rs:534-555
 *       (same label, m = 1, chain = proof_vector.len(): the proofs of one call run on ONE transcript, in order)
 * for nproofs independent transcripts.  commitments: nproofs x chain x m x 32 B (the reference passes acc.comm.d);
 * proofs: nproofs x chain x (9 + 2 lg(n_bits m)) x 32 B, each RangeProof::to_bytes() = A | S | T_1 | T_2 | t_x |
 * t_x_blinding | e_blinding | L_0 | R_0 | .. | a | b.  Generators: PedersenGens::default(), BulletproofGens::new(64, 16).
 * The transcript of proof p is Transcript::new(transcript_label) + Verifier::new(verifier_label) (verifier_label NULL: a bare
 * merlin transcript, for RangeProof::verify_multiple outside the reference's Verifier), or - when
 * transcript_state is not NULL - the qq_transcript_state_bytes() bytes at transcript_state + p * that size (a transcript an
 * earlier verification on the same Verifier left behind, see qq_transcript_capture); then domain_sep(domain_label) unless
 * domain_label is NULL.
 * Host threads run the transcripts; the GPU folds the generator scalars of all proofs with random weights and evaluates ONE
 * aggregated MSM (2 n m + 2 shared terms + 4 + 2 lg(n m) + m terms per proof); a failing aggregate is bisected, so the
 * verdict is per transcript.  status[p]: QQ_ST_OK = Ok(()), QQ_ST_PROOF = Err("Bulletproof verification failed"),
 * QQ_ST_BAD_POINT = the same Err caused by an undecodable point, QQ_ST_BAD_SCALAR = a non-canonical scalar (the crate's
 * RangeProof::from_bytes FormatError). */
int qq_verify_range_proof_batch(qq_ctx* ctx, const char* transcript_label, const char* verifier_label,
                                const uint8_t* transcript_state, const char* domain_label, const uint8_t* commitments,
                                const uint8_t* proofs, size_t n_bits, size_t m, size_t chain, size_t nproofs, uint8_t* status);
/* Size of one serialised verifier transcript: 200 bytes STROBE-128 state | pos | pos_begin | cur_flags | tag 0xa5 | 4 zero
 * bytes of merlin::Transcript.  qq_verify_range_proof_batch validates tag and positions; an entry that fails (all-zero: the
 * sigma check of that proof ended before the capture) gets QQ_ST_PROOF. */
size_t qq_transcript_state_bytes(void);
/* The NEXT entry point called on this ctx, if it is a sigma-proof verification that keeps a transcript
 * (qq_verify_account_sigma_batch, qq_verify_zero_balance_batch, qq_verify_destroy_account_batch,
 * qq_verify_same_value_compact_batch, qq_verify_update_account_dark_tx_batch, qq_verify_ddh_batch), also writes the transcript
 * of each proof, as it stands after the challenge, to states_out: min(capacity_states, nproofs) entries of
 * qq_transcript_state_bytes(), zero-filled for proofs whose check ended before the challenge.  The reference keeps one
 * running transcript per Verifier across verify_account_verifier_bulletproof and the range proof (verifier.rs:1603-1628).
 * One-shot: ANY next entry point (whatever it returns) disarms it; NULL cancels. */
int qq_transcript_capture(qq_ctx* ctx, uint8_t* states_out, size_t capacity_states);

/* ---- several GPUs behind one handle -------------------------------------------------------------------
 * qq_multi owns one qq_ctx and one worker thread per listed device (the first is the root).  Batches of independent units are
 * cut into contiguous slices, one per device, no data-path exchange (results land in the caller's arrays at their positions).
 * qq_multi_msm cuts ONE multiscalar multiplication (Verifier::multiscalar_multiplication, src/accounts/verifier.rs:91-99; the
 * Bulletproofs mega-MSM) into slices of (scalar, point) pairs: every device runs Pippenger on its slice, the 144-byte partial
 * results are pulled into the root's memory with peer copies (NVLink / NVSwitch) and added and encoded there by one kernel -
 * they never visit the host.  Same outputs and status codes as the single-device entry points.  One caller thread at a time. */
typedef struct qq_multi qq_multi;
int qq_init_multi(qq_multi** out, const int* devices, int ndev);
void qq_destroy_multi(qq_multi* m);
int qq_multi_device_count(const qq_multi* m);
qq_ctx* qq_multi_ctx(qq_multi* m, int index);      /* the per-device context (tuning knobs, qq_dev_alloc for the _dev form) */
const char* qq_multi_last_error(const qq_multi* m);
int qq_multi_update_account_batch(qq_multi* m, const uint8_t* acc, const uint8_t* bl, const uint8_t* u, const uint8_t* c,
                                  uint8_t* out_acc, uint8_t* status, size_t n);
int qq_multi_generate_commitment_batch(qq_multi* m, const uint8_t* pk, const uint8_t* r, const uint8_t* v, uint8_t* out_comm,
                                       uint8_t* status, size_t n);
int qq_multi_verify_shuffle_batch(qq_multi* m, const char* transcript_label, const char* verifier_label, const uint8_t* shuffle_input,
                                  const uint8_t* shuffle_output, const uint8_t* statement, const uint8_t* proof, size_t nproofs,
                                  uint8_t* status, uint8_t* stage, uint8_t* detail);
int qq_multi_verify_range_proof_batch(qq_multi* m, const char* transcript_label, const char* verifier_label,
                                      const uint8_t* transcript_state, const char* domain_label, const uint8_t* commitments,
                                      const uint8_t* proofs, size_t n_bits, size_t m_values, size_t chain, size_t nproofs, uint8_t* status);
int qq_multi_msm(qq_multi* m, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t* out_point, uint8_t* status);
/* the same with the slices already resident: scalars_per_device[d] / points_per_device[d] are DEVICE pointers on device d
 * (count_per_device[d] x 32 B each) */
int qq_multi_msm_dev(qq_multi* m, const uint8_t* const* scalars_per_device, const uint8_t* const* points_per_device,
                     const size_t* count_per_device, uint8_t* out_point, uint8_t* status);
/* sum of k partial MSM results that sit in DEVICE memory as records of `stride` bytes (a multiple of 16, >= 128): 128 B canonical
 * X | Y | Z | T as qq_msm_partial_dev writes them, and - when stride > 128 - the status byte at offset 128 (e.g. the output of an
 * NCCL all-gather of those records).  out_point: compressed sum (zeros when a status is set); *is_identity, *status may be NULL. */
int qq_points_sum_dev(qq_ctx* ctx, const uint8_t* records, size_t k, size_t stride, uint8_t* out_point, uint8_t* is_identity,
                      uint8_t* status);

/* ---- wire format ---------------------------------------------------------------------------------
 * bincode 1.x (Cargo.toml:29; default configuration: little endian, fixed-width integers, Vec = u64 length + elements, enum =
 * u32 variant index) encodings of the reference's serde-derived types <-> the flattened layouts above, so that batches can be
 * fed straight from serialised transactions.  Host-only helpers without a context.  QQ_ERR_ARG: truncated input, or a Vec
 * whose length is not the one ShuffleProof::verify indexes (ROWS = COLUMNS = 3; the reference would panic / return Err).
 * *consumed / *written (may be NULL) receive the number of bincode bytes read / needed.
 *   ShuffleProof      src/shuffle/shuffle.rs:164-184 (+ hadamard.rs:34-57, product.rs:39-92, singlevalueproduct.rs:33-48,
 *                     multiexponential.rs:37-56, ddh.rs:27-32): 3 920 B each -> 3 776 B each
 *   ShuffleStatement  src/shuffle/shuffle.rs:153-161: 360 B each -> 352 B each */
int qq_shuffle_proofs_from_bincode(const uint8_t* in, size_t in_len, size_t nproofs, uint8_t* out_proofs, size_t* consumed);
int qq_shuffle_statements_from_bincode(const uint8_t* in, size_t in_len, size_t nproofs, uint8_t* out_statements, size_t* consumed);
int qq_shuffle_proofs_to_bincode(const uint8_t* proofs, size_t nproofs, uint8_t* out, size_t out_cap, size_t* written);
int qq_shuffle_statements_to_bincode(const uint8_t* statements, size_t nproofs, uint8_t* out, size_t out_cap, size_t* written);
/* Vec<Account> (src/accounts/accounts.rs:47-53; e.g. the shuffle's input / output account lists): u64 length + 128 B per
 * account.  *n_accounts is set even when cap_accounts is too small (QQ_ERR_ARG then). */
int qq_accounts_from_bincode(const uint8_t* in, size_t in_len, uint8_t* out_accounts, size_t cap_accounts, size_t* n_accounts,
                             size_t* consumed);
/* SigmaProof (src/accounts/prover.rs:20-26): *variant 0 = Dlog(z, x), lens[0] = |z|; 1 = Dleq(zv, zr1, zr2, x), lens[0..3).
 * out_scalars receives the vectors back to back (at most cap_scalars entries of 32 B), out_x the challenge. */
int qq_sigma_proof_from_bincode(const uint8_t* in, size_t in_len, int* variant, uint8_t* out_scalars, size_t cap_scalars,
                                size_t lens[3], uint8_t out_x[32], size_t* consumed);

/* ---- decommit ------------------------------------------------------------------------------------------------------
 * ElGamalCommitment::decommit(sk) = enc(d - sk*c) = enc(v*B)                    src/elgamal/elgamal.rs:106-108 */
int qq_decommit_batch(qq_ctx* ctx, const uint8_t* comm, const uint8_t* sk, uint8_t* out_points, uint8_t* status, size_t n);
/* ElGamalCommitment::decommit_value(sk) (src/elgamal/elgamal.rs:119-122, brute_force_decrypt :169-182; the reference's
 * tests recover 160000 and 16734, elgamal.rs:293-303, accounts.rs:584-595): the smallest v < 2^search_bits with
 * v*B = d - sk*c, by baby-step / giant-step on the GPU (1 <= search_bits <= 48).  status 0 and out_values[i] = v, or
 * QQ_ST_NOT_FOUND / QQ_ST_BAD_POINT / QQ_ST_BAD_SCALAR with out_values[i] = 0. */
int qq_decommit_value_batch(qq_ctx* ctx, const uint8_t* comm, const uint8_t* sk, int search_bits, uint64_t* out_values,
                            uint8_t* status, size_t n);

/* ---- hash-to-group and generator derivation ---------------------------------------------------------------------
 * RistrettoPoint::from_uniform_bytes over n blocks of 64 uniform bytes (the tail of hash_from_bytes::<Sha3_512>,
 * src/pedersen/vectorpedersen.rs:49-51,66-70): out_i = enc(elligator(lo_i) + elligator(hi_i)). */
int qq_from_uniform_bytes_batch(qq_ctx* ctx, const uint8_t* uniform64, uint8_t* out_points, size_t n);
/* VectorPedersenGens::new(capacity) (src/pedersen/vectorpedersen.rs:45-75): out_h = H (32 B), out_g = G_vec
 * ((capacity - 1) x 32 B: B followed by the SHA3-512 hash chain started at H).  capacity >= 2. */
int qq_vector_pedersen_gens(qq_ctx* ctx, size_t capacity, uint8_t* out_h, uint8_t* out_g);
/* bulletproofs::BulletproofGens::new(gens_capacity, party_capacity), rebuilt by the reference on every range-proof
 * verification (src/accounts/verifier.rs:510,540): out_g, out_h = party_capacity x gens_capacity x 32 B, party-major.
 * Feed them to qq_msm_points_prepare once and reuse the handle. */
int qq_bulletproof_gens(qq_ctx* ctx, size_t gens_capacity, size_t party_capacity, uint8_t* out_g, uint8_t* out_h);

#ifdef __cplusplus
}
#endif
#endif /* QQ_B200_H */
