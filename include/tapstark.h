/*
 * tapstark.h -- C ABI of libtapstark_b200.so: the B200-native (sm_100a) prover commitment hot path of
 * TapSTARK (bitlayer-org/tap-stark).
 *
 * The reference is a Rust workspace with no FFI boundary; the replaceable units are three Plonky3-derived
 * traits.  Each entry point below names the reference interface it stands in for (paths relative to the
 * reference checkout).  A Rust shim (tap-stark_b200/rust/) implements those traits by
 * calling this ABI; statuses != TS_OK are turned into panics there, matching the reference's
 * assert!/expect convention.
 *
 * Conventions
 *   - BabyBear elements are u32 in MONTGOMERY form (x * 2^32 mod p, p = 0x78000001): the in-memory form of
 *     p3-baby-bear's `BabyBear` at Plonky3 rev 72b2fc16 [MEM], so a Rust Vec<BabyBear> is passed as-is.
 *   - BabyBear^4 (BinomialExtensionField<BabyBear,4>, basic/src/field/mod.rs:53-64) = 4 consecutive u32,
 *     low-degree coefficient first.  An extension matrix h x w is a base matrix h x 4w.
 *   - Matrices are row-major (p3-matrix RowMajorMatrix), rows = height, width in u32 elements.
 *   - "committed order" = rows bit-reversed, i.e. what
 *     `dft.coset_lde_batch(..).bit_reverse_rows().to_row_major_matrix()` yields
 *     (fri/src/two_adic_pcs.rs:237-240).
 *   - Digests / commitments are 32 bytes = [[u8;4];8] as observed by the challenger
 *     (basic/src/challenger/mod.rs:197-223).
 *   - A ts_ctx is used by one host thread at a time.  All work is enqueued on the ctx stream; entry points
 *     that return host-visible results synchronise that stream before returning.
 *   - No CPU fallback exists: without a CUDA device ts_ctx_create fails.
 */
#ifndef TAPSTARK_H
#define TAPSTARK_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TS_OK 0
#define TS_ERR_CUDA 1
#define TS_ERR_ARG 2
#define TS_ERR_NOT_CONSTANT 3 /* fri/src/prover.rs:130-134 final layer not constant */
#define TS_ERR_NO_WITNESS 4   /* basic/src/challenger/mod.rs:101 "failed to find witness" */

#define TS_P 0x78000001u
#define TS_GENERATOR_MONTY 0x0fffffbeu /* 31 in Montgomery form: Val::generator(), two_adic_pcs.rs:235 */

typedef struct ts_ctx ts_ctx;
typedef struct ts_matrix ts_matrix; /* device-resident RowMajorMatrix<BabyBear> */
typedef struct ts_tree ts_tree;     /* BFMmcs::ProverData: digest layers + the committed matrices */
typedef struct ts_challenger ts_challenger;

/* ---------------------------------------------------------------- context */
/* stream: a cudaStream_t to enqueue on (e.g. torch's current stream), or NULL for a private stream. */
int ts_ctx_create(int device, void *stream, ts_ctx **out);
void ts_ctx_destroy(ts_ctx *ctx);
const char *ts_last_error(const ts_ctx *ctx);
int ts_ctx_synchronize(ts_ctx *ctx);
/* 1 if this library was built for a real GPU (always, for the shipped library). */
int ts_is_device_build(void);

/* Per-kernel-class device timers (CUDA events on the ctx stream), for bench.py's roofline.
 * kind: see TS_K_*.  Returns accumulated ms and launch count since the last reset. */
#define TS_K_NTT_PASS 0  /* in-place DIF digit pass (ntt.cuh: ntt_pass_kernel) */
#define TS_K_LDE_MID 1   /* fused inverse-last-digit + coset scale + forward-first-digit (lde_mid_kernel) */
#define TS_K_HASH_LEAVES 2
#define TS_K_TREE 3
#define TS_K_FOLD 4
#define TS_K_MISC 5
#define TS_K_COUNT 6
int ts_ctx_set_profiling(ts_ctx *ctx, int on);
int ts_ctx_reset_stats(ts_ctx *ctx);
int ts_ctx_get_stats(ts_ctx *ctx, int kind, double *ms, uint64_t *launches);
uint64_t ts_ctx_total_launches(const ts_ctx *ctx);
/* Freed matrices/trees go to a stream-ordered cache instead of cudaFree (which would synchronise the device
 * every step); ts_ctx_trim returns the cache to the driver. */
int ts_ctx_trim(ts_ctx *ctx);

/* ---------------------------------------------------------------- matrices */
int ts_matrix_alloc(ts_ctx *ctx, size_t rows, size_t width, ts_matrix **out);
/* copies a host RowMajorMatrix (Montgomery u32) to the device (pinned staging is the caller's business) */
int ts_matrix_from_host(ts_ctx *ctx, const uint32_t *host, size_t rows, size_t width, ts_matrix **out);
/* copies from a device pointer (device-to-device) */
int ts_matrix_from_device(ts_ctx *ctx, const uint32_t *dev, size_t rows, size_t width, ts_matrix **out);
/* wraps caller-owned device memory without copying; the caller keeps it alive */
int ts_matrix_wrap_device(ts_ctx *ctx, uint32_t *dev, size_t rows, size_t width, ts_matrix **out);
/* rows [row0, row0+nrows) -> host.  Serves BFMmcs::get_matrices (basic/src/mmcs/bf_mmcs.rs:52) and
 * Pcs::get_evaluations_on_domain (fri/src/two_adic_pcs.rs:247-258). */
int ts_matrix_download(ts_ctx *ctx, const ts_matrix *m, size_t row0, size_t nrows, uint32_t *host);
uint32_t *ts_matrix_device_ptr(const ts_matrix *m);
size_t ts_matrix_rows(const ts_matrix *m);
size_t ts_matrix_width(const ts_matrix *m);
void ts_matrix_free(ts_matrix *m);
/* in-place canonical <-> Montgomery on the device (helpers for hosts that hold canonical values) */
int ts_matrix_to_monty(ts_ctx *ctx, ts_matrix *m);
int ts_matrix_from_monty(ts_ctx *ctx, ts_matrix *m);
/* out = bit-reversed-rows copy of m (p3-matrix bit_reverse_rows().to_row_major_matrix()) */
int ts_matrix_bit_reverse_rows(ts_ctx *ctx, const ts_matrix *m, ts_matrix **out);

/* ---------------------------------------------------------------- p3_dft::TwoAdicSubgroupDft<BabyBear>
 * Replaces the `Dft` parameter of TwoAdicFriPcs (fri/src/two_adic_pcs.rs:207, used at :237-240).
 * `evals`/`coeffs` are consumed logically (the trait takes the matrix by value) but NOT freed or modified.
 * Outputs are new device matrices owned by the caller.
 *   ts_coset_lde_batch : the PCS hot call.  out = coset_lde_batch(evals, added_bits, shift) in COMMITTED
 *                        (bit-reversed) row order, (rows<<added_bits) x width.
 *   natural_order != 0 : undo the bit reversal (plain trait semantics, one extra permutation pass).      */
int ts_coset_lde_batch(ts_ctx *ctx, const ts_matrix *evals, unsigned added_bits, uint32_t shift_monty,
                       int natural_order, ts_matrix **out);
int ts_dft_batch(ts_ctx *ctx, const ts_matrix *coeffs, ts_matrix **out);                       /* natural -> natural */
int ts_idft_batch(ts_ctx *ctx, const ts_matrix *evals, ts_matrix **out);                       /* natural -> natural */
int ts_coset_dft_batch(ts_ctx *ctx, const ts_matrix *coeffs, uint32_t shift_monty, ts_matrix **out);
int ts_lde_batch(ts_ctx *ctx, const ts_matrix *evals, unsigned added_bits, ts_matrix **out);   /* shift = 1, natural */
/* Host-buffer form of the hot call (what a drop-in `impl TwoAdicSubgroupDft for GpuDft` does):
 * H2D, LDE, D2H.  out_host must hold (rows<<added_bits)*width u32; committed order unless natural_order. */
int ts_coset_lde_batch_host(ts_ctx *ctx, const uint32_t *evals_host, size_t rows, size_t width,
                            unsigned added_bits, uint32_t shift_monty, int natural_order, uint32_t *out_host);

/* Host-buffer forms of the plain trait methods (dft_batch / idft_batch / coset_dft_batch of
 * p3_dft::TwoAdicSubgroupDft; natural order in, natural order out; fri/src/fold_even_odd.rs:75-81 calls `dft`):
 * H2D, transform, D2H inside the call.  shift_monty is used by TS_COSET_DFT only. */
#define TS_DFT 0
#define TS_IDFT 1
#define TS_COSET_DFT 2
int ts_dft_batch_host(ts_ctx *ctx, int kind, const uint32_t *mat_host, size_t rows, size_t width, uint32_t shift_monty,
                      uint32_t *out_host);
/* Page-lock / unlock a caller-owned host buffer (a Rust Vec<BabyBear> the prover keeps for the whole proof): the *_host
 * entry points then copy its column windows straight out of the caller's rows at the PCIe rate.  Optional: a pageable
 * source is gathered by host threads into page-locked bounce slots inside the call (measured 132 ms against 91 ms for a
 * 2^22 x 256 commit; page-locking the buffer per call would cost ~1 s). */
int ts_host_register(ts_ctx *ctx, const void *host, size_t bytes);
int ts_host_unregister(ts_ctx *ctx, const void *host);

/* ---------------------------------------------------------------- basic::mmcs::bf_mmcs::BFMmcs<T>
 * (basic/src/mmcs/bf_mmcs.rs:17-68).  Blake3 row hash + 2-to-1 Blake3 tree over device-resident matrices.
 *   TS_LAYOUT_P3_INJECT : [MEM] p3-merkle-tree FieldMerkleTreeMmcs<.., SerializingHasher32<Blake3>,
 *                         CompressionFunctionFromHasher<u8,Blake3,2,32>, 32> (north_star's "Blake3 Merkle").
 *   TS_LAYOUT_PADDED    : one leaf layer whose leaf i concatenates, largest matrix first, row
 *                         i >> (log_max - log_h) of every matrix -- the reference's padding_matrix
 *                         (basic/src/tcs/mod.rs:339-378).
 * Heights must be powers of two.  Matrices are BORROWED unless take_ownership != 0 (then the tree frees them). */
#define TS_LAYOUT_P3_INJECT 0
#define TS_LAYOUT_PADDED 1
int ts_mmcs_commit(ts_ctx *ctx, ts_matrix *const *mats, size_t n_mats, int layout, int take_ownership,
                   uint8_t root[32], ts_tree **out);
size_t ts_tree_num_matrices(const ts_tree *t);
ts_matrix *ts_tree_matrix(const ts_tree *t, size_t i); /* BFMmcs::get_matrices */
size_t ts_tree_depth(const ts_tree *t);                /* siblings in an opening */
size_t ts_tree_max_height(const ts_tree *t);           /* BFMmcs::get_max_height */
/* the 32-byte root, device to device on the context's stream (no host synchronisation): a rank's sub-root goes
 * straight into the all-gather buffer; ts_mmcs_commit with root == NULL skips the host copy of the root. */
int ts_tree_root_copy(ts_ctx *ctx, const ts_tree *t, uint8_t *dst_device);
/* BFMmcs::open_batch(query_index): rows_out receives every matrix's opened row (Montgomery) concatenated
 * in caller order (sum of widths u32); path_out receives depth x 32 bytes. */
int ts_mmcs_open_batch(ts_ctx *ctx, const ts_tree *t, size_t index, uint32_t *rows_out, uint8_t *path_out);
/* BFMmcs::verify_batch; host-side (the verifier is out of scope for the GPU). returns TS_OK / TS_ERR_ARG */
int ts_mmcs_verify_batch(const size_t *heights, const size_t *widths, size_t n_mats, int layout, size_t index,
                         const uint32_t *rows_monty, const uint8_t *path, size_t depth, const uint8_t root[32]);
/* copies digest layer `layer` (0 = leaves) to host, 32 bytes per node */
int ts_tree_layer(ts_ctx *ctx, const ts_tree *t, size_t layer, uint8_t *out, size_t *n_nodes);
void ts_tree_free(ts_tree *t);

/* ---------------------------------------------------------------- BfChallenger<F, U32, Blake3Permutation, 16>
 * (basic/src/challenger/mod.rs, chan_field.rs).  Host-side state, 64-byte sponge.  Sampled field values
 * are returned CANONICAL (they are transcript integers), ext = 4 coefficients. */
int ts_challenger_new(ts_challenger **out);
int ts_challenger_clone(const ts_challenger *c, ts_challenger **out);
void ts_challenger_free(ts_challenger *c);
void ts_challenger_observe(ts_challenger *c, const uint8_t word[4]);
void ts_challenger_observe_digest(ts_challenger *c, const uint8_t digest[32]);
uint32_t ts_challenger_sample_base(ts_challenger *c);
void ts_challenger_sample_ext(ts_challenger *c, uint32_t out[4]);
size_t ts_challenger_sample_bits(ts_challenger *c, unsigned bits, int ext);
int ts_challenger_check_witness(ts_challenger *c, unsigned bits, uint32_t witness, int ext);
/* deterministic: the smallest valid witness in 0..4096 (the reference's rayon find_any is racy) */
int ts_challenger_grind(ts_challenger *c, unsigned bits, int ext, uint32_t *witness);

/* ---------------------------------------------------------------- FRI fold
 * TwoAdicFriGenericConfig::fold_matrix (fri/src/two_adic_pcs.rs:116-147) == fold_even_odd
 * (fri/src/fold_even_odd.rs:20-52).  in: 2h elements viewed as h x 2; out: h elements.
 * beta is Montgomery form (1 u32 for base, 4 for ext).  Pointers are DEVICE pointers. */
int ts_fri_fold_base(ts_ctx *ctx, const uint32_t *in_dev, size_t h, uint32_t beta_monty, uint32_t *out_dev);
int ts_fri_fold_ext(ts_ctx *ctx, const uint32_t *in_dev, size_t h, const uint32_t beta_monty[4],
                    uint32_t *out_dev);
/* host-buffer form (what the fri crate's fold_even_odd(Vec<F>, beta) -> Vec<F> binds to) */
int ts_fri_fold_ext_host(ts_ctx *ctx, const uint32_t *in_host, size_t h, const uint32_t beta_monty[4],
                         uint32_t *out_host);

/* ---------------------------------------------------------------- FRI commit phase
 * bf_commit_phase (fri/src/prover.rs:93-141): commit layer -> observe -> sample beta -> fold -> add the
 * next input of equal length, until `blowup` values remain; they must all be equal.
 * inputs: device matrices of EF elements (width 4, one row per element), lengths strictly descending
 * (two_adic_pcs.rs:389).  Outputs: commits (rounds x 32 bytes), trees[round] (prover data for the query
 * phase, each owning its layer -- round 0 holds a clone of inputs[0] like the reference's folded.clone();
 * pass NULL to discard and skip the clone), final_poly (canonical, 4 u32), *rounds. */
int ts_fri_commit_phase(ts_ctx *ctx, ts_matrix *const *inputs, size_t n_inputs, unsigned log_blowup,
                        ts_challenger *chal, uint8_t *commits, ts_tree **trees, uint32_t final_poly[4],
                        size_t *rounds);

/* ---------------------------------------------------------------- Pcs (basic/src/bf_pcs.rs:19-88)
 * TwoAdicFriPcs::commit (fri/src/two_adic_pcs.rs:227-245): for every (domain, evals):
 * shift = generator / domain.shift; LDE; bit-reverse rows; one MMCS commit over all LDEs.
 * domain_shifts_monty[i] is domain.shift (Montgomery); evals borrowed; the tree owns the LDE matrices. */
int ts_pcs_commit(ts_ctx *ctx, ts_matrix *const *evals, const uint32_t *domain_shifts_monty, size_t n,
                  unsigned log_blowup, int layout, uint8_t root[32], ts_tree **out);
/* host-buffer form: evals_host[i] is rows[i] x widths[i] */
int ts_pcs_commit_host(ts_ctx *ctx, const uint32_t *const *evals_host, const size_t *rows, const size_t *widths,
                       const uint32_t *domain_shifts_monty, size_t n, unsigned log_blowup, int layout,
                       uint8_t root[32], ts_tree **out);
/* Pcs::get_evaluations_on_domain (two_adic_pcs.rs:247-258): first `domain_size` committed rows of matrix
 * idx, re-bit-reversed -> host (domain.shift must be the generator). */
int ts_pcs_get_evaluations_on_domain(ts_ctx *ctx, const ts_tree *t, size_t idx, size_t domain_size,
                                     uint32_t *out_host);
/* sum_i alpha^i * column_i of a committed matrix -> EF vector (rows x 4): the hot inner loop of
 * TwoAdicFriPcs::open (`mat.dot_ext_powers(alpha)`, two_adic_pcs.rs:375).  alpha Montgomery. */
int ts_dot_ext_powers(ts_ctx *ctx, const ts_matrix *m, const uint32_t alpha_monty[4], ts_matrix **out);


/* Pcs::open (basic/src/bf_pcs.rs:61-74) = TwoAdicFriPcs::open (fri/src/two_adic_pcs.rs:260-419) followed by bf_prove
 * (fri/src/prover.rs:19-90), whole, on the device-resident commitments -- what a Rust `impl Pcs for GpuTwoAdicFriPcs`
 * calls; no host language has to re-implement the orchestration.
 *   rounds[r]        prover data of the r-th ts_pcs_commit being opened
 *   n_points[]       for every matrix of every round (round-major, matrices in commit order) its number of points
 *   points_monty     4 u32 (Montgomery BabyBear^4) per point, in the same order
 *   chal             the caller's challenger: sampled (alpha, betas, query indices) and observed (layer roots,
 *                    proof-of-work witness) exactly as the reference does, so the caller's transcript continues
 * Output: *out_bytes (free with ts_bytes_free) holds the postcard encoding (serde; the format the reference's own
 * commented-out test uses, uni-stark/tests/mul_air.rs:133) of the pair `(OpenedValues, FriProof)`:
 *   OpenedValues = Vec<Vec<Vec<Vec<Challenge>>>>                 round, matrix, point, column  (two_adic_pcs.rs:325-386)
 *   FriProof { commit_phase_commits: Vec<[u8;32]>, query_proofs: Vec<BfQueryProof>, final_poly, pow_witness }
 *   BfQueryProof { input_proof: Vec<BatchOpening { opened_values: Vec<Vec<Val>>, opening_proof: Vec<[u8;32]> }>,
 *                  commit_phase_openings: Vec<(Vec<Vec<Challenge>>, Vec<[u8;32]>)> }             (fri/src/proof.rs:13-33)
 * with unsigned integers and lengths as LEB128 varints, BabyBear as its canonical u32 ([MEM] p3-baby-bear's Serialize;
 * parity unpinned -- no serialized vector exists in the reference), BabyBear^4 as 4 of them, digests as 32 raw bytes.
 * The Merkle opening proof (sibling digests, leaf level first) stands where the reference has its Taproot
 * CommitedProof. */
int ts_pcs_open(ts_ctx *ctx, const ts_tree *const *rounds, size_t n_rounds, const size_t *n_points,
                const uint32_t *points_monty, unsigned log_blowup, unsigned num_queries, unsigned proof_of_work_bits,
                ts_challenger *chal, uint8_t **out_bytes, size_t *out_len);
void ts_bytes_free(uint8_t *bytes);

/* ---------------------------------------------------------------- TapTree commitment (f2, first slice)
 * The commitment TapTreeMmcs really computes (basic/src/mmcs/taptree_mmcs.rs:101-114 -> basic/src/tcs/mod.rs:238-282): one
 * Bitcoin script per leaf (tcs/mod.rs:197-225), BIP-341 TapLeaf hashes, a TapBranch tree of lexicographically sorted pairs
 * and the leaf permutation that sorting causes (basic/src/tcs/builder.rs:38-93).
 * Every leaf of a tree carries the same bit-commitment locking scripts around pushed integers, so the host passes the
 * template once:  script(i) = seg[0] P(i) seg[1] P(x_1) ... P(x_{n_push-1}) seg[n_push],  P = minimal script-number push,
 * x_k = canonical value of word push_word[k-1] of row i of `leaf_rows` (Montgomery on the device; for extension elements
 * the host lists the limbs in the reversed order of tcs/mod.rs:214).  segs = the n_push + 1 segments concatenated,
 * seg_offsets = n_push + 2 offsets.  The locking-script bytes themselves come from the external `bitcomm` crate (PARITY
 * UNPINNED; oracle/taptree.py restates them from the in-tree twin under scripts/src/bit_comm).  The reference repeats this
 * commit num_queries times with fresh bit-commitments (tcs/mod.rs:284-292): call once per template.
 * Outputs: root (32 bytes, as TapNodeHash serialises), the tree handle; leaf_indices[m] = position of Merkle leaf m among
 * the TapTree's leaves (CompleteTaptree's reverse_idx_dict); ts_taptree_level downloads one level (0 = leaf hashes). */
typedef struct ts_taptree ts_taptree;
int ts_taptree_commit(ts_ctx *ctx, const ts_matrix *leaf_rows, const uint8_t *segs, const size_t *seg_offsets,
                      const uint32_t *push_word, size_t n_push, uint8_t root[32], ts_taptree **out);
int ts_taptree_leaf_indices(ts_ctx *ctx, const ts_taptree *t, uint32_t *out_host);
int ts_taptree_level(ts_ctx *ctx, const ts_taptree *t, unsigned level, uint8_t *out_host);
void ts_taptree_free(ts_taptree *t);
/* The rows the leaves commit to when several matrices of different heights share one tree: PolyTCS::padding_matrix
 * (basic/src/tcs/mod.rs:341-383) -- matrices tallest first (stable), a matrix of height h repeats each of its rows over
 * hmax / h consecutive leaves, rows concatenated.  Heights are powers of two.  *out: hmax x (sum of widths), Montgomery. */
int ts_padded_leaf_rows(ts_ctx *ctx, const ts_matrix *const *mats, size_t n_mats, ts_matrix **out);
/* CommitedData::query_proof (basic/src/tcs/mod.rs:141-146) without the script bytes: the TaprootMerkleBranch of Merkle leaf
 * `index` (log2(n_leaves) sibling hashes of 32 bytes, leaf level first; verify_inclusion, complete_taptree.rs:64-73, folds
 * them with sorted pairs) and its position among the TapTree's leaves (leaf_indices[index]; may be NULL). */
int ts_taptree_open(ts_ctx *ctx, const ts_taptree *t, size_t index, uint8_t *path_out, uint32_t *position_out);

/* ---------------------------------------------------------------- quotient values (f3)
 * quotient_values of uni_stark::prove (uni-stark/src/prover.rs:122-194) on the device-resident trace LDE:
 * for every point of the quotient domain g*H_m, m = 2^(log_n + log_quotient_degree) <= rows of the LDE, run the
 * AIR's constraint program on (row i, row i + m/n) of get_evaluations_on_domain's view (two_adic_pcs.rs:247-258),
 * fold the constraints with alpha (ProverConstraintFolder::assert_zero, uni-stark/src/folder.rs:60-64), divide by
 * Z_H.  Output: 2^log_quotient_degree chunk matrices of (m >> log_quotient_degree) x 4 (flatten_to_base +
 * split_evals, prover.rs:79-80), chunk k living on the coset g*w_m^k*H_n -- the input of the second ts_pcs_commit.
 * Program: 4 words per instruction {op, dst, a, b}; operand = kind << 28 | index with kind 0 register (< 64),
 * 1 local column, 2 next column, 3 public value, 4 constant, 5 selector (0 is_first_row, 1 is_last_row,
 * 2 is_transition); op 0 ADD, 1 SUB, 2 MUL, 3 NEG, 4 ASSERT_ZERO(a).  Constants, public values and alpha are
 * Montgomery form.  A malformed program is TS_ERR_ARG. */
int ts_quotient_values(ts_ctx *ctx, const ts_matrix *trace_lde, unsigned log_n, unsigned log_quotient_degree,
                       const uint32_t *program, size_t n_instr, const uint32_t *consts_monty, size_t n_consts,
                       const uint32_t *public_values_monty, size_t n_public, const uint32_t alpha_monty[4],
                       ts_matrix **chunks_out);

/* ---------------------------------------------------------------- reduced openings: TwoAdicFriPcs::open (f1)
 * The pass between the commitments and FRI (fri/src/two_adic_pcs.rs:312-389), on the device-resident LDE. */
/* compute_inverse_denominators (:677-720): out[X] = 1 / (x_X - z), x_X = g * w_h^bitrev(X), h = 2^log_h, as an
 * h x 4 matrix.  Smaller heights use its first rows (bit-reversed cosets nest). */
int ts_inv_denoms(ts_ctx *ctx, unsigned log_h, const uint32_t z_monty[4], ts_matrix **out);
/* interpolate_coset (:358-369): ys[c] = p_c(z) from the low coset = the first n committed rows of `lde`;
 * inv_denoms from ts_inv_denoms for the same z.  ys_out_monty: width x 4 u32 (host). */
int ts_interpolate_low_coset(ts_ctx *ctx, const ts_matrix *lde, size_t n, const uint32_t z_monty[4],
                             const ts_matrix *inv_denoms, uint32_t *ys_out_monty);
/* "reduce rows" (:371-381): ro[X] += alpha_pow_offset * (dot[X] - reduced_ys) * inv_denoms[X] */
int ts_reduce_opening_acc(ts_ctx *ctx, const ts_matrix *dot, const ts_matrix *inv_denoms,
                          const uint32_t alpha_pow_offset_monty[4], const uint32_t reduced_ys_monty[4], ts_matrix *ro);
int ts_matrix_zero(ts_ctx *ctx, ts_matrix *m);
/* Synthetic inputs (the reference's tests draw theirs from `RowMajorMatrix::rand(&mut rng, ..)`, fri/tests/pcs.rs,
 * fri/src/fold_even_odd.rs:70; that ChaCha stream cannot be reproduced without the rand crates):
 * fills the rows x width device buffer `dev` with columns [col0, col0+width) of the rows x total_width matrix whose
 * element (r, c) is SplitMix64((seed << 40) + r*total_width + c) mod p -- a pure function of (seed, r, c), so every rank of
 * a column-sharded run and the CPU oracle (oracle.splitmix_matrix) hold the same trace.  monty != 0: Montgomery form. */
int ts_fill_splitmix(ts_ctx *ctx, uint32_t *dev, size_t rows, size_t width, uint64_t seed, size_t col0,
                     size_t total_width, int monty);

/* ---------------------------------------------------------------- sharded building blocks (one process per GPU)
 * The path shards as SURVEY 8(e): LDE by columns (no communication), one all-to-all to re-shard by rows,
 * hashing / alpha-reduction / folding on contiguous row ranges; sub-roots are all-gathered (32 bytes per rank)
 * and the top log2(G) tree levels are hashed on the host.  Collectives live in the host layer
 * (tap-stark_b200/parallel.py, torch.distributed NCCL); these entry points are the per-rank pieces. */
/* LDE written into caller-owned memory (e.g. the send buffer of the all-to-all). */
int ts_coset_lde_batch_into(ts_ctx *ctx, const ts_matrix *evals, unsigned added_bits, uint32_t shift_monty,
                            ts_matrix *out);
/* Fused LDE + re-shard: the committed LDE of this rank's columns with every output row written by the LAST butterfly
 * pass straight to the rank that owns its row range -- committed row r goes to owner_ptrs[r / (N / n_owners)] at local
 * row r % (N / n_owners), dst_pitch words per row, columns [0, width).  Remote owners are peer-mapped device memory
 * (ts_ipc_open), i.e. the stores travel over NVLink as part of the kernel and no all-to-all follows; the caller
 * orders "all ranks have written" before anyone reads (one tiny collective on the stream).  Needs the position-major
 * two-digit path (rows >= 2^18, width % 4 == 0); otherwise TS_ERR_ARG and the caller uses ts_coset_lde_batch_into +
 * all-to-all. */
int ts_coset_lde_batch_scatter(ts_ctx *ctx, const ts_matrix *evals, unsigned added_bits, uint32_t shift_monty,
                               uint32_t *const *owner_ptrs, size_t n_owners, size_t dst_pitch);
/* device allocations shareable across the processes of one node, and their CUDA IPC handles (64 bytes) */
int ts_device_malloc(ts_ctx *ctx, size_t bytes, void **out);
int ts_device_free(ts_ctx *ctx, void *p);
int ts_ipc_get_handle(ts_ctx *ctx, void *dev_ptr, uint8_t handle[64]);
int ts_ipc_open(ts_ctx *ctx, const uint8_t handle[64], void **out);
int ts_ipc_close(ts_ctx *ctx, void *p);
/* Copy-engine transfers beside the kernels (no SM is taken from the LDE, unlike NCCL's copy kernels): dst/src may be this
 * device's memory, a peer's IPC-mapped buffer (NVLink) or PINNED host memory (e.g. a column window of the host's
 * row-major trace: ts_copy2d_async with src_pitch = full row bytes).  The copy runs on transfer stream `lane` (0..7:
 * independent queues, e.g. lane 0 host staging, lanes 1..7 peer copies spread over the copy engines); after_main != 0 orders it after everything queued on the
 * context stream so far.  ts_copy_join makes the context stream wait for all transfers issued on that lane so far.
 * Across ranks the caller still needs one collective before reading what PEERS wrote (e.g. a 4-byte all-reduce after
 * ts_copy_join). */
int ts_copy_async(ts_ctx *ctx, int lane, void *dst, const void *src, size_t bytes, int after_main);
int ts_copy2d_async(ts_ctx *ctx, int lane, void *dst, size_t dst_pitch, const void *src, size_t src_pitch,
                    size_t width_bytes, size_t rows, int after_main);
int ts_copy_join(ts_ctx *ctx, int lane);
/* alpha^0 .. alpha^(count-1) (+16 zero entries) as a device matrix, computed once per opening */
int ts_alpha_powers(ts_ctx *ctx, const uint32_t alpha_monty[4], size_t count, ts_matrix **out);
/* acc (+)= sum_c alpha^(first_power + c) * m[:, c]; a row shard whose columns arrive as several blocks calls it
 * once per block with first_power = the block's first global column.  No host synchronisation. */
int ts_dot_ext_powers_acc(ts_ctx *ctx, const ts_matrix *m, const ts_matrix *alpha_powers, size_t first_power,
                          ts_matrix *acc, int accumulate);
/* acc = sum over the concatenated columns of blocks[0..n_blocks) (equal row counts): one pass over all blocks
 * when they have equal power-of-two widths (the column blocks of the all-to-all), else one pass per block. */
int ts_dot_ext_powers_blocks(ts_ctx *ctx, ts_matrix *const *blocks, size_t n_blocks, const ts_matrix *alpha_powers,
                             ts_matrix *acc);
/* ALL row-sharded commit-phase rounds of a step in one call, the sub-root exchange inside the kernels: every rank owns a
 * MAILBOX (ts_fri_mailbox_words() u32 of zero-filled device memory, e.g. from ts_device_malloc) that its peers map with
 * ts_ipc_open; mailboxes[s] is rank s's mailbox as seen from this process (mailboxes[rank] = its own).  Per round: leaf
 * hash + subtree, the sub-root is stored into every rank's mailbox (NVLink), the sponge step starts when all sub-roots of
 * the round have arrived, the fold reads beta from device memory.  epoch: any value that differs from the previous call's
 * (a step counter), the same on every rank.  cur_dev: this rank's rows of the layer (len_global / world extension
 * elements); out_dev receives its rows after n_rounds folds.  The host challenger replays the rounds at the end, as in
 * ts_fri_chain_end; commits_out = n_rounds x 32 bytes. */
int ts_fri_mailbox_words(void);
int ts_fri_commit_phase_sharded(ts_ctx *ctx, const uint32_t *cur_dev, size_t len_global, size_t rank, size_t world,
                                size_t n_rounds, uint32_t *const *mailboxes, uint32_t epoch, ts_challenger *chal,
                                uint8_t *commits_out, uint32_t *out_dev);
/* Incremental form of ts_mmcs_commit (BFMmcs::commit, basic/src/mmcs/bf_mmcs.rs:23) for ONE matrix whose rows reach a rank
 * as equal-width column blocks over time (the re-shard of a column-sharded LDE): the blocks are registered up front
 * (borrowed, in column order), every window [block_begin, block_end) -- whole 64-byte Blake3 blocks of the row -- is
 * absorbed as soon as its blocks have arrived, in order, and finish builds the tree.  The result equals ts_mmcs_commit
 * over the same blocks.  Rows of at most 256 columns (one Blake3 chunk). */
int ts_mmcs_commit_begin(ts_ctx *ctx, ts_matrix *const *blocks, size_t n_blocks, ts_tree **out);
int ts_mmcs_commit_window(ts_ctx *ctx, ts_tree *tree, size_t block_begin, size_t block_end);
int ts_mmcs_commit_finish(ts_ctx *ctx, ts_tree *tree, uint8_t *root_or_null);
/* Chained commit-phase rounds of a ROW-SHARDED layer (fri/src/prover.rs:112-126 per round): nothing returns to the host
 * between rounds.  After whole digests have been observed the BfChallenger state is (root, previous squeeze), so one small
 * kernel per round combines the ranks' sub-roots into the layer root (top log2(n_sub) levels), advances the sponge
 * (h = Blake3(root || h_prev)) and leaves beta/2 in device memory, where the fold of that round reads it.
 *   begin : device state initialised from the host challenger (TS_ERR_ARG if it holds partially observed input)
 *   step  : sub_roots_dev = n_sub x 32 bytes in rank order (the caller's all-gather), round = 0, 1, ...
 *   fold  : ts_fri_fold_ext_shard with this round's beta taken from the chain
 *   end   : ONE read-back of all layer roots; the host challenger replays observe/sample per round; commits_out
 *           (n_rounds x 32 bytes) may be NULL; frees the chain. */
int ts_fri_chain_begin(ts_ctx *ctx, const ts_challenger *chal, size_t max_rounds, uint32_t **chain_dev);
int ts_fri_chain_step(ts_ctx *ctx, uint32_t *chain_dev, const uint8_t *sub_roots_dev, size_t n_sub, size_t round);
int ts_fri_fold_ext_shard_chain(ts_ctx *ctx, const uint32_t *in_dev, size_t h_global, size_t first, size_t h_local,
                                const uint32_t *chain_dev, const uint32_t *addend_dev, uint32_t *out_dev);
int ts_fri_chain_end(ts_ctx *ctx, uint32_t *chain_dev, ts_challenger *chal, size_t n_rounds, uint8_t *commits_out);
/* fold rows [first, first + h_local) of a layer of h_global rows (fold_matrix on a contiguous row range);
 * addend_dev (may be NULL) is the matching slice of the next FRI input. */
int ts_fri_fold_ext_shard(ts_ctx *ctx, const uint32_t *in_dev, size_t h_global, size_t first, size_t h_local,
                          const uint32_t beta_monty[4], const uint32_t *addend_dev, uint32_t *out_dev);
/* The same fold that also emits the leaf digests of the NEXT commit-phase round (fri/src/prover.rs:112-113: the folded vector
 * viewed as rows of two extension elements, hashed as canonical little-endian words): next_digests_dev = (h_local / 2) x 8 words.
 * The layer is then not read a second time by a leaf-hash pass.  h_global >= 512; first and h_local multiples of 512. */
int ts_fri_fold_hash_shard(ts_ctx *ctx, const uint32_t *in_dev, size_t h_global, size_t first, size_t h_local,
                           const uint32_t beta_monty[4], const uint32_t *addend_dev, uint32_t *out_dev,
                           uint32_t *next_digests_dev);
/* plain Blake3 on the host: compress the G sub-roots into the root (and what the verifier side uses) */
void ts_blake3_host(const uint8_t *in, size_t len, uint8_t out[32]);

#ifdef __cplusplus
}
#endif
#endif
