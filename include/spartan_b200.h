/*
 * spartan_b200 — C ABI of the B200-native Spartan prover (libspartan_b200.so).
 *
 * Drop-in boundary for the prover hot path of microsoft/Spartan (libspartan 0.9.0).  The reference has no FFI of its own
 * (pure Rust, single crate); these are the entry points a `libspartan-sys` shim would bind, each naming the Rust item it
 * replaces (paths relative to /root/reference).  Conventions:
 *   - every function returns an int status: SP_OK or an SP_ERR_* code; sp_last_error() gives the message.  No exception or
 *     panic crosses the boundary (the reference panics on misuse, e.g. src/r1csproof.rs:156).
 *   - scalars cross as `uint64_t[4]` little-endian limbs in MONTGOMERY form = the in-memory layout of `Scalar`
 *     (src/scalar/ristretto255.rs:195-199), so a Rust `&[Scalar]` is passed as-is (`as_ptr() as *const u64`).
 *     Where the reference API takes canonical bytes (`[u8;32]`, src/lib.rs:64,121) so does this ABI.
 *   - group elements cross only as 32-byte ristretto255 encodings (`CompressedRistretto::as_bytes`).
 *   - handles are opaque and own device memory; one sp_ctx per host thread / per GPU.
 *   - there is NO CPU fallback: sp_ctx_create fails with SP_ERR_NO_DEVICE when no CUDA device is present.
 */
#ifndef SPARTAN_B200_H
#define SPARTAN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SP_OK 0
#define SP_ERR_NO_DEVICE 1
#define SP_ERR_CUDA 2
#define SP_ERR_INVALID_ARG 3
#define SP_ERR_INVALID_INDEX 4     /* R1CSError::InvalidIndex            src/errors.rs:28-41 */
#define SP_ERR_INVALID_SCALAR 5    /* R1CSError::InvalidScalar */
#define SP_ERR_INVALID_INPUTS 6    /* R1CSError::InvalidNumberOfInputs */
#define SP_ERR_INTERNAL 7
#define SP_ERR_VERIFY 9            /* ProofVerifyError::InternalError: the proof was rejected                    src/errors.rs:5-12 */
#define SP_ERR_DECOMPRESS 10       /* ProofVerifyError::DecompressionError: a point of the proof does not decompress */
#define SP_ERR_INVALID_POINT 8     /* CompressedRistretto::decompress() == None (ProofVerifyError::DecompressionError, src/errors.rs:10) */

typedef struct sp_ctx sp_ctx;
typedef struct sp_poly sp_poly;           /* DensePolynomial.Z on the device          src/dense_mlpoly.rs:18-22 */
typedef struct sp_gens sp_gens;           /* MultiCommitGens + fixed-base tables      src/commitments.rs:8-12 */
typedef struct sp_points sp_points;       /* a caller-supplied &[GroupElement] on the device   src/group.rs:8-9 */
typedef struct sp_instance sp_instance;   /* Instance (R1CSShape + digest)            src/lib.rs:111-114 */
typedef struct sp_nizk_gens sp_nizk_gens; /* NIZKGens                                 src/lib.rs:468-486 */
typedef struct sp_snark_gens sp_snark_gens;     /* SNARKGens                          src/lib.rs:277-309 */
typedef struct sp_snark_encoding sp_snark_encoding; /* (ComputationCommitment, ComputationDecommitment)  src/lib.rs:44-55 */

/* ---- context */
int sp_device_count(void);
int sp_ctx_create(int device, sp_ctx** out);
void sp_ctx_destroy(sp_ctx* ctx);
const char* sp_last_error(const sp_ctx* ctx);  /* ctx may be NULL: message of the last failed sp_ctx_create */
unsigned long long sp_kernel_launches(void);   /* kernels launched by this library since load */

/* ---- one proof over several GPUs (the reference is single-process, the loops being sharded are dense_mlpoly.rs:215-223,
 * sumcheck.rs:290-357 / :460-469 / :625-652, dense_mlpoly.rs:165-177, product_tree.rs:18-56).  One process per GPU, one context per process.
 * Every rank calls sp_comm_export, the hosts exchange the sp_comm_handle_bytes()-byte handles by any transport (MPI, a socket, torch.distributed),
 * every rank calls sp_comm_connect with all handles in rank order (world = 2, 4 or 8 GPUs of one NVLink domain).  From then on sp_nizk_prove* and
 * sp_snark_prove* must be called by EVERY rank with identical arguments: each call is one sharded proof, every rank returns the same proof bytes —
 * the bytes the single-GPU prover returns.  Operator-level entry points, encode and the verifiers stay single-GPU. */
size_t sp_comm_handle_bytes(void);
int sp_comm_export(sp_ctx* ctx, uint8_t* handle_out);
int sp_comm_connect(sp_ctx* ctx, int rank, int world, const uint8_t* handles /* world x sp_comm_handle_bytes() */);
int sp_comm_info(const sp_ctx* ctx, int* rank, int* world);
int sp_comm_set_enabled(sp_ctx* ctx, int enabled);   /* 0: the next prove calls run on this rank's GPU alone (all ranks must switch together) */
/* self-test of the host helper threads that share the single-point commitments of a ZK sumcheck round (engine.hpp: HostPool; SP_HOST_THREADS=0 disables
 * them): runs `iterations` small jobs sets, returns the number of jobs that did not run exactly once; *helpers = number of helper threads.  No GPU needed. */
int sp_host_pool_selftest(int iterations, int* helpers);
/* 1 (default): SNARK::prove commits to the dereferenced SPARK values on a background stream underneath the latency-bound rounds that precede the
 * commitment in the transcript (same bytes, shorter proof); 0: every kernel on the context's stream, one after the other (per-kernel profiling) */
int sp_ctx_set_overlap(sp_ctx* ctx, int enabled);
/* phase timings of the last prove call, the labels of the reference's `profile` feature (src/timer.rs; src/r1csproof.rs:152-298) */
int sp_timings(sp_ctx* ctx, char* buf, size_t buflen);
/* bytes moved host->device / device->host by this library since load (bench.py e2e leg) */
void sp_io_bytes(unsigned long long* h2d_bytes, unsigned long long* d2h_bytes);
/* CUDA-event timing of every kernel family on the launching stream (bench.py roofline leg).  sp_prof_enable(1) clears and starts;
 * sp_prof_report writes "name:launches:total_ms:total_algorithmic_bytes;..." after synchronising the device */
void sp_prof_enable(int on);
int sp_prof_report(char* buf, size_t buflen);
/* device-side stopwatch on the context's stream (CUDA events) */
int sp_timer_start(sp_ctx* ctx);
int sp_timer_stop_ms(sp_ctx* ctx, float* ms);

/* ---- scalar field helpers (host side, for harnesses) */
int sp_scalar_from_bytes(const uint8_t canonical[32], uint64_t out_mont[4]);        /* Scalar::from_bytes       ristretto255.rs:391 */
void sp_scalar_to_bytes(const uint64_t mont[4], uint8_t out_canonical[32]);         /* Scalar::to_bytes         ristretto255.rs:419 */
void sp_scalar_from_bytes_wide(const uint8_t wide[64], uint64_t out_mont[4]);       /* Scalar::from_bytes_wide  ristretto255.rs:435 */
void sp_scalar_mul(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);      /* Mul for Scalar           ristretto255.rs:690 */
void sp_scalar_add(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);      /* ristretto255.rs:736 */
void sp_scalar_sub(const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);      /* ristretto255.rs:718 */
int sp_scalar_invert(const uint64_t a[4], uint64_t out[4]);                         /* Scalar::invert           ristretto255.rs:541 */

/* ---- dense multilinear polynomials on the device */
int sp_poly_upload(sp_ctx* ctx, const uint64_t* limbs_mont, size_t len, sp_poly** out);   /* DensePolynomial::new   dense_mlpoly.rs:121 */
int sp_poly_download(sp_ctx* ctx, const sp_poly* p, uint64_t* limbs_mont, size_t len);
size_t sp_poly_len(const sp_poly* p);
void sp_poly_free(sp_poly* p);
/* bound_poly_var_top on k polynomials of equal length                                      dense_mlpoly.rs:215-223 */
int sp_fold_top(sp_ctx* ctx, sp_poly* const* polys, int k, const uint64_t r_mont[4]);
/* one sumcheck round evaluation: kind 0 = A*B (sumcheck.rs:460-469), 1 = A*B*C (sumcheck.rs:204-228), 2 = A*(B*C-D) (sumcheck.rs:625-652).
 * out = [e(0), e(2), e(3)] as Montgomery limbs (e(3) = 0 for kind 0) */
int sp_sumcheck_eval(sp_ctx* ctx, int kind, sp_poly* const* polys, uint64_t out[3][4]);
/* fused: bind the top variable of every table to r, then evaluate the next round on the folded tables */
int sp_sumcheck_fold_eval(sp_ctx* ctx, int kind, sp_poly* const* polys, const uint64_t r_mont[4], uint64_t out[3][4]);
/* the same two calls on a connected context, on CYCLIC SHARDS (rank r holds entries r, r+W, ...): partial sums are exchanged over NVLink inside
 * the kernel ("scalar-add all-reduce"), every rank gets the evaluations of the whole tables; shards of >= 8192 entries for fold_eval */
int sp_sumcheck_eval_sharded(sp_ctx* ctx, int kind, sp_poly* const* polys, uint64_t out[3][4]);
int sp_sumcheck_fold_eval_sharded(sp_ctx* ctx, int kind, sp_poly* const* polys, const uint64_t r[4], uint64_t out[3][4]);
/* SumcheckInstanceProof::prove_cubic_batched evaluation loops (sumcheck.rs:290-357): ninst instances of comb = A*B*C (product_tree.rs:283-286).
 * out = ninst x [e0, e2, e3] Montgomery limbs.  The same sp_poly may be passed as C of several instances (poly_C_par); A and B must be distinct. */
int sp_sumcheck_batched_eval(sp_ctx* ctx, int ninst, sp_poly* const* A, sp_poly* const* B, sp_poly* const* C, uint64_t* out /* ninst*3*4 */);
/* bound_poly_var_top(r) on every table (each shared C once) fused with the next round's evaluations */
int sp_sumcheck_batched_fold_eval(sp_ctx* ctx, int ninst, sp_poly* const* A, sp_poly* const* B, sp_poly* const* C, const uint64_t r_mont[4],
                                  uint64_t* out /* ninst*3*4 */);
int sp_eq_evals(sp_ctx* ctx, const uint64_t* r_mont, size_t ell, sp_poly** out);           /* EqPolynomial::evals   dense_mlpoly.rs:68-84 */
int sp_poly_evaluate(sp_ctx* ctx, const sp_poly* p, const uint64_t* r_mont, size_t ell, uint64_t out[4]); /* DensePolynomial::evaluate :236 */
int sp_poly_bound_rows(sp_ctx* ctx, const sp_poly* p, const uint64_t* L_mont, size_t L_size, sp_poly** out); /* DensePolynomial::bound :206 */
int sp_dot(sp_ctx* ctx, const sp_poly* a, const sp_poly* b, uint64_t out[4]);              /* compute_dotproduct   nizk/mod.rs:435 */

/* ---- Pedersen generators and commitments */
/* MultiCommitGens::new(n, label): G[0..n), h = G[n]  (n+1 points drawn)                    commitments.rs:15-33 */
int sp_gens_create(sp_ctx* ctx, const uint8_t* label, size_t label_len, size_t n, sp_gens** out);
/* the caller's own MultiCommitGens (e.g. after `scale`, commitments.rs:43-49): n+1 encodings, G[0..n) then h; SP_ERR_INVALID_POINT if one fails */
int sp_gens_upload(sp_ctx* ctx, const uint8_t* compressed32, size_t n, sp_gens** out);
void sp_gens_free(sp_gens* g);
int sp_gens_export(sp_ctx* ctx, const sp_gens* g, uint8_t* out32 /* (n+1)*32: G then h */);
/* GroupElement::vartime_multiscalar_mul(scalars, G[0..n)) compressed                       group.rs:98-117 */
int sp_msm(sp_ctx* ctx, const sp_gens* g, const uint64_t* scalars_mont, size_t n, uint8_t out32[32]);
/* DensePolynomial::commit_inner: L rows of R scalars; out[i] = (MSM(row_i, G) + blinds[i]*h).compress(); blinds may be NULL (zero)
 *                                                                                          dense_mlpoly.rs:148-177 */
int sp_commit_rows(sp_ctx* ctx, const sp_gens* g, const sp_poly* p, size_t L, size_t R, const uint64_t* blinds_mont, uint8_t* out32);
int sp_point_decompress_check(sp_ctx* ctx, const uint8_t* in32, size_t n, int* ok);        /* CompressedRistretto::decompress */
int sp_point_roundtrip(sp_ctx* ctx, const uint8_t* in32, size_t n, uint8_t* out32);        /* decompress().compress() on the device */

/* ---- variable-base MSM on caller-supplied points (bucket method; no precomputed tables — for point sets that are used once or are too
 * large for window tables, e.g. the 2^24-point standalone MSM of BASELINE.json).                                   group.rs:98-117 */
/* points cross as n 32-byte encodings; SP_ERR_INVALID_POINT if any fails to decompress */
int sp_points_upload(sp_ctx* ctx, const uint8_t* compressed32, size_t n, sp_points** out);
/* MultiCommitGens::new(n, label).G (the h generator is not kept)                                                   commitments.rs:15-33 */
int sp_points_derive(sp_ctx* ctx, const uint8_t* label, size_t label_len, size_t n, sp_points** out);
size_t sp_points_len(const sp_points* p);
void sp_points_free(sp_points* p);
int sp_points_export(sp_ctx* ctx, const sp_points* p, size_t offset, size_t n, uint8_t* out32);
/* GroupElement::vartime_multiscalar_mul(scalars[0..n), points[offset..offset+n)).compress() */
int sp_msm_var(sp_ctx* ctx, const sp_points* p, size_t offset, const uint64_t* scalars_mont, size_t n, uint8_t out32[32]);
int sp_msm_var_resident(sp_ctx* ctx, const sp_points* p, size_t offset, const sp_poly* scalars, uint8_t out32[32]);
/* every rank of a connected context passes its slice of the points and of the scalars; point-add all-reduce of the W partial sums; every rank
 * returns the encoding of the whole sum (group.rs:98-117 on a vector split by index range) */
int sp_msm_var_sharded(sp_ctx* ctx, const sp_points* p, size_t offset, const sp_poly* scalars, uint8_t out32[32]);

/* ---- inner-product argument, operator level: BulletReductionProof::prove (src/nizk/bullet.rs:32-132) for a host that keeps the transcript and
 * the orchestration.  sp_ipa_begin copies a_vec / b_vec; the generator vector G is `gens` (G[0..n)).  Each round: sp_ipa_round_LR returns the
 * compressed L, R of bullet.rs:83-97 (c_L, c_R are formed inside; Q and H are the caller's points), the host appends them and derives u
 * (bullet.rs:99-103), sp_ipa_fold applies bullet.rs:105-108.  After lg n rounds sp_ipa_finish returns a[0], b[0] and G[0] (bullet.rs:113-122). */
typedef struct sp_ipa sp_ipa;
int sp_ipa_begin(sp_ctx* ctx, const sp_gens* gens, const sp_poly* a_vec, const sp_poly* b_vec, sp_ipa** out);
void sp_ipa_free(sp_ipa* h);
int sp_ipa_round_LR(sp_ctx* ctx, sp_ipa* h, const uint8_t Q32[32], const uint8_t H32[32], const uint64_t blind_L_mont[4], const uint64_t blind_R_mont[4],
                    uint8_t L32[32], uint8_t R32[32]);
int sp_ipa_fold(sp_ctx* ctx, sp_ipa* h, const uint64_t u_mont[4], const uint64_t u_inv_mont[4]);
int sp_ipa_finish(sp_ctx* ctx, sp_ipa* h, uint64_t a_hat_mont[4], uint64_t b_hat_mont[4], uint8_t G_hat32[32]);

/* ---- instances                                                                           lib.rs:111-274 */
/* Instance::new: entries are (row, col, canonical 32-byte value); rows < num_cons, cols < num_vars + 1 + num_inputs */
int sp_instance_create(sp_ctx* ctx, size_t num_cons, size_t num_vars, size_t num_inputs, const uint64_t* A_row, const uint64_t* A_col, const uint8_t* A_val32,
                       size_t nA, const uint64_t* B_row, const uint64_t* B_col, const uint8_t* B_val32, size_t nB, const uint64_t* C_row, const uint64_t* C_col,
                       const uint8_t* C_val32, size_t nC, sp_instance** out);
/* Instance::produce_synthetic_r1cs with the OsRng replaced by the seeded SHAKE256 generator of DESIGN.md.
 * vars_out: num_vars*4 limbs, inputs_out: num_inputs*4 limbs (Montgomery) */
int sp_instance_synthetic(sp_ctx* ctx, size_t num_cons, size_t num_vars, size_t num_inputs, uint64_t seed, sp_instance** out, uint64_t* vars_out,
                          uint64_t* inputs_out);
void sp_instance_free(sp_instance* inst);
int sp_instance_dims(const sp_instance* inst, size_t* num_cons, size_t* num_vars, size_t* num_inputs);
/* R1CSShape::get_digest (r1cs.rs:154-158): NIZK::prove / verify absorb zlib(bincode(shape)) (lib.rs:514).  The reference compresses with flate2's
 * miniz_oxide backend at level 6; this library computes the same stream with its own restatement of miniz's compressor (csrc/deflate.cpp, bit-identical
 * to C miniz level 6) on first use.  A caller that wants its own compressor's bytes absorbed instead passes them with sp_instance_set_digest. */
int sp_instance_set_digest(sp_instance* inst, const uint8_t* digest, size_t len);
int sp_instance_digest(const sp_instance* inst, uint8_t** out, size_t* len);   /* free with sp_free */
/* bincode(R1CSShape) — the bytes the reference deflates to obtain the digest (r1cs.rs:154-158) */
int sp_instance_bincode(const sp_instance* inst, uint8_t** out, size_t* len);
/* COO export in the reference's entry order: matrix 0/1/2 = A/B/C; vals are Montgomery limbs */
int sp_instance_nnz(const sp_instance* inst, int matrix, size_t* nnz);
int sp_instance_export(const sp_instance* inst, int matrix, uint64_t* row, uint64_t* col, uint64_t* val_mont);
int sp_instance_is_sat(sp_ctx* ctx, const sp_instance* inst, const uint64_t* vars_mont, size_t nvars, const uint64_t* inputs_mont, size_t ninputs, int* sat);

/* ---- NIZK                                                                                lib.rs:468-591 */
int sp_nizk_gens_create(sp_ctx* ctx, size_t num_cons, size_t num_vars, size_t num_inputs, sp_nizk_gens** out);
void sp_nizk_gens_free(sp_nizk_gens* g);
/* NIZK::prove(inst, vars, inputs, gens, &mut Transcript::new(transcript_label)); tape_seed = the scalar RandomTape::new draws from OsRng
 * (random.rs:13-15).  vars are HOST Montgomery limbs and are copied to the device inside the call.  *proof = bincode::serialize(&NIZK),
 * released with sp_free. */
int sp_nizk_prove(sp_ctx* ctx, const sp_instance* inst, const uint64_t* vars_mont, size_t nvars, const uint64_t* inputs_mont, size_t ninputs,
                  const sp_nizk_gens* gens, const uint8_t* transcript_label, size_t label_len, const uint64_t tape_seed_mont[4], uint8_t** proof,
                  size_t* proof_len);
/* same with the assignment already resident on the device (bench "value" leg) */
int sp_nizk_prove_resident(sp_ctx* ctx, const sp_instance* inst, const sp_poly* vars, const uint64_t* inputs_mont, size_t ninputs, const sp_nizk_gens* gens,
                           const uint8_t* transcript_label, size_t label_len, const uint64_t tape_seed_mont[4], uint8_t** proof, size_t* proof_len);

/* ---- SNARK                                                                               lib.rs:277-465 */
int sp_snark_gens_create(sp_ctx* ctx, size_t num_cons, size_t num_vars, size_t num_inputs, size_t num_nz_entries, sp_snark_gens** out);
void sp_snark_gens_free(sp_snark_gens* g);
int sp_snark_encode(sp_ctx* ctx, const sp_instance* inst, const sp_snark_gens* gens, sp_snark_encoding** out);   /* SNARK::encode lib.rs:325 */
void sp_snark_encoding_free(sp_snark_encoding* e);
/* bincode::serialize(&ComputationCommitment) */
int sp_snark_commitment_bytes(const sp_snark_encoding* e, uint8_t** out, size_t* len);
int sp_snark_prove(sp_ctx* ctx, const sp_instance* inst, const sp_snark_encoding* enc, const uint64_t* vars_mont, size_t nvars, const uint64_t* inputs_mont,
                   size_t ninputs, const sp_snark_gens* gens, const uint8_t* transcript_label, size_t label_len, const uint64_t tape_seed_mont[4],
                   uint8_t** proof, size_t* proof_len);
int sp_snark_prove_resident(sp_ctx* ctx, const sp_instance* inst, const sp_snark_encoding* enc, const sp_poly* vars, const uint64_t* inputs_mont,
                            size_t ninputs, const sp_snark_gens* gens, const uint8_t* transcript_label, size_t label_len,
                            const uint64_t tape_seed_mont[4], uint8_t** proof, size_t* proof_len);

/* bincode(ComputationCommitment) (as written by sp_snark_commitment_bytes or by the reference) -> a commitment-only handle for sp_snark_verify */
int sp_snark_commitment_load(sp_ctx* ctx, const uint8_t* bytes, size_t len, sp_snark_encoding** out);

/* ---- verifiers: SP_OK = accepted; SP_ERR_VERIFY / SP_ERR_DECOMPRESS = rejected (sp_last_error names the failing check).  `proof` is
 * bincode::serialize(&NIZK) / (&SNARK), as produced by sp_*_prove or by the reference. */
/* NIZK::verify(&self, &Instance, &InputsAssignment, &mut Transcript, &NIZKGens)                         lib.rs:549-591 */
int sp_nizk_verify(sp_ctx* ctx, const sp_instance* inst, const uint64_t* inputs_mont, size_t ninputs, const sp_nizk_gens* gens, const uint8_t* label,
                   size_t label_len, const uint8_t* proof, size_t proof_len);
/* SNARK::verify(&self, &ComputationCommitment, &InputsAssignment, &mut Transcript, &SNARKGens)          lib.rs:423-465
 * (`comm`: a handle from sp_snark_encode or from sp_snark_commitment_load; only its commitment part is read) */
int sp_snark_verify(sp_ctx* ctx, const sp_snark_encoding* comm, const uint64_t* inputs_mont, size_t ninputs, const sp_snark_gens* gens, const uint8_t* label,
                    size_t label_len, const uint8_t* proof, size_t proof_len);

/* ---- the caller-owned transcript.  The reference's prove / verify take `transcript: &mut Transcript` (src/lib.rs:339-347, :423-429, :501-508,
 * :549-555): the caller may have absorbed its own data before the call and may keep using the transcript afterwards.  The *_t entry points take
 * merlin's whole STROBE-128 state instead of a label — SP_TRANSCRIPT_STATE_BYTES = 200 bytes of Keccak state, then pos, pos_begin, cur_flags
 * (merlin::strobe::Strobe128) — and leave it exactly as the reference's call leaves the caller's transcript.  sp_*_prove / sp_*_verify with a
 * label are the shorthand for a transcript the caller has just created with Transcript::new(label) and does not reuse (benches/snark.rs:56). */
#define SP_TRANSCRIPT_STATE_BYTES 203
int sp_nizk_prove_t(sp_ctx* ctx, const sp_instance* inst, const uint64_t* vars_mont, size_t nvars, const uint64_t* inputs_mont, size_t ninputs,
                    const sp_nizk_gens* gens, uint8_t* strobe_state /* in, out */, const uint64_t tape_seed[4], uint8_t** proof, size_t* proof_len);
int sp_snark_prove_t(sp_ctx* ctx, const sp_instance* inst, const sp_snark_encoding* enc, const uint64_t* vars_mont, size_t nvars, const uint64_t* inputs_mont,
                     size_t ninputs, const sp_snark_gens* gens, uint8_t* strobe_state /* in, out */, const uint64_t tape_seed[4], uint8_t** proof, size_t* proof_len);
int sp_nizk_verify_t(sp_ctx* ctx, const sp_instance* inst, const uint64_t* inputs_mont, size_t ninputs, const sp_nizk_gens* gens, uint8_t* strobe_state /* in, out */,
                     const uint8_t* proof, size_t proof_len);
int sp_snark_verify_t(sp_ctx* ctx, const sp_snark_encoding* comm, const uint64_t* inputs_mont, size_t ninputs, const sp_snark_gens* gens,
                      uint8_t* strobe_state /* in, out */, const uint8_t* proof, size_t proof_len);
/* merlin::Transcript::{new, append_message, challenge_bytes} on such a state buffer (transcript.rs:13-30), for hosts without merlin and for the tests */
int sp_transcript_new(const uint8_t* label, size_t label_len, uint8_t* strobe_state /* out */);
int sp_transcript_append_message(uint8_t* strobe_state, const uint8_t* label, size_t label_len, const uint8_t* msg, size_t msg_len);
int sp_transcript_challenge_bytes(uint8_t* strobe_state, const uint8_t* label, size_t label_len, uint8_t* out, size_t n);

void sp_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
