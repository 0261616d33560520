/*
 * tapstark_oracle.h -- CPU ORACLE for the TapSTARK prover commitment hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by the product
 * library (tap-stark_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 * It restates, in plain C over CANONICAL u32 field values, the algorithms of the reference
 * (bitlayer-org/tap-stark, /root/reference) for the path
 *     coset LDE -> bit-reversed rows -> row-hash + tree commitment -> FRI commit phase (fold)
 * Each function cites the reference file:line it follows.  Where the arithmetic lives in an
 * un-vendored dependency (Plonky3 @72b2fc162738df459619488a98bb06eaf64e5b4a: p3-dft, p3-field,
 * p3-baby-bear, p3-matrix, p3-merkle-tree, p3-symmetric; blake3 1.5) the published algorithm is
 * restated and marked [MEM].
 *
 * PARITY PINNING (see tests/test_oracle_pins.py):
 *   pinned by reference KATs : Blake3 (scripts/src/hashes/blake3.rs:537-587), challenger golden
 *                              1103171332 (script_expr/src/challenger_expr.rs:278-296), fold property
 *                              (fri/src/fold_even_odd.rs:65-95), padded leaf layout (basic/src/tcs/mod.rs:594-602)
 *   parity unpinned          : LDE known-answer values (no KAT in the reference; pinned only against the
 *                              mathematical definition, or_naive_dft), Blake3-Merkle roots (construction absent
 *                              from the reference; pinned against the `blake3` PyPI package).
 */
#ifndef TAPSTARK_ORACLE_H
#define TAPSTARK_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define OR_P 0x78000001u /* basic/src/field/mod.rs:45 */

/* ---- BabyBear, canonical representation --------------------------------------------------- */
uint32_t or_bb_add(uint32_t a, uint32_t b);
uint32_t or_bb_sub(uint32_t a, uint32_t b);
uint32_t or_bb_mul(uint32_t a, uint32_t b);
uint32_t or_bb_pow(uint32_t a, uint64_t e);
uint32_t or_bb_inv(uint32_t a);
uint32_t or_two_adic_generator(unsigned bits); /* [MEM] p3-baby-bear: 0x1a427a41^(2^(27-bits)) */
uint32_t or_to_monty(uint32_t x);              /* x * 2^32 mod p  ([MEM] in-memory form of p3 BabyBear) */
uint32_t or_from_monty(uint32_t x);
void or_to_monty_vec(uint32_t *v, size_t n);
void or_from_monty_vec(uint32_t *v, size_t n);

/* ---- BabyBear^4 = F_p[x]/(x^4-11), coefficients low->high (basic/src/field/mod.rs:53-64) ---- */
void or_ef_add(const uint32_t a[4], const uint32_t b[4], uint32_t o[4]);
void or_ef_sub(const uint32_t a[4], const uint32_t b[4], uint32_t o[4]);
void or_ef_mul(const uint32_t a[4], const uint32_t b[4], uint32_t o[4]);
void or_ef_inv(const uint32_t a[4], uint32_t o[4]);

/* ---- DFT family ([MEM] p3-dft TwoAdicSubgroupDft; matrices row-major h x w) ---------------- */
void or_naive_dft(const uint32_t *in, uint32_t *out, unsigned log_n, size_t w); /* definition: out[i]=sum_k in[k] w^(ik) */
void or_dft_batch(uint32_t *mat, unsigned log_n, size_t w);                     /* natural -> natural */
void or_idft_batch(uint32_t *mat, unsigned log_n, size_t w);
void or_coset_dft_batch(uint32_t *mat, unsigned log_n, size_t w, uint32_t shift);
/* out is (n<<added_bits) x w, natural order: out[i] = p(shift * w_N^i) */
void or_coset_lde_batch(const uint32_t *in, unsigned log_n, size_t w, unsigned added_bits,
                        uint32_t shift, uint32_t *out);
void or_bit_reverse_rows(uint32_t *mat, unsigned log_h, size_t w);
/* fri/src/two_adic_pcs.rs:235-240: coset_lde_batch(evals, log_blowup, shift).bit_reverse_rows() */
void or_pcs_lde_committed(const uint32_t *in, unsigned log_n, size_t w, unsigned added_bits,
                          uint32_t shift, uint32_t *out);

/* ---- Blake3 (plain hash mode, 32-byte output) ---------------------------------------------- */
void or_blake3(const uint8_t *in, size_t len, uint8_t out[32]);

/* ---- MMCS: Blake3 row hash + binary tree ---------------------------------------------------- */
#define OR_LAYOUT_P3_INJECT 0 /* [MEM] p3-merkle-tree FieldMerkleTree: shorter matrices injected per layer */
#define OR_LAYOUT_PADDED 1    /* basic/src/tcs/mod.rs:339-378 padding_matrix: one leaf layer, rows repeated */
typedef struct or_tree or_tree;
/* mats[i] is heights[i] x widths[i] canonical u32 (EF matrices passed flattened: width*4). Borrowed. */
or_tree *or_mmcs_commit(const uint32_t *const *mats, const size_t *heights, const size_t *widths,
                        size_t k, int layout, uint8_t root[32]);
size_t or_tree_depth(const or_tree *t);       /* number of sibling digests in a path */
size_t or_tree_num_layers(const or_tree *t);
const uint8_t *or_tree_layer(const or_tree *t, size_t layer, size_t *len_digests);
/* opened rows are written concatenated in the caller's matrix order; path = depth x 32 bytes */
void or_mmcs_open_batch(const or_tree *t, size_t index, uint32_t *rows_out, uint8_t *path_out);
int or_mmcs_verify_batch(const size_t *heights, const size_t *widths, size_t k, int layout,
                         size_t index, const uint32_t *rows, const uint8_t *path, size_t depth,
                         const uint8_t root[32]);
void or_tree_free(or_tree *t);
/* leaf value list of the padded layout (for the reference's comment vectors, tcs/mod.rs:594-602) */
size_t or_padded_leaf(const uint32_t *const *mats, const size_t *heights, const size_t *widths,
                      size_t k, size_t leaf, uint32_t *out);

/* ---- SHA-256 (FIPS 180-4) and the BIP-341 tagged hashes of the TapTree commitment (basic/src/tcs/builder.rs:26,64 through
 * rust-bitcoin [MEM]): TapLeaf = H_tag("TapLeaf", 0xc0 || compact_size(len) || script), TapBranch = H_tag("TapBranch", min || max).
 * Pinned to hashlib in tests/test_taptree.py. */
void or_sha256(const uint8_t *data, size_t len, uint8_t out[32]);
void or_tap_leaf_hash(const uint8_t *script, size_t len, uint8_t out[32]);
void or_tap_branch_hash(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);

/* ---- FRI fold (fri/src/two_adic_pcs.rs:116-147 == fri/src/fold_even_odd.rs:20-52) ---------- */
void or_fold_matrix_bb(const uint32_t *in, unsigned log_h, uint32_t beta, uint32_t *out);
void or_fold_matrix_ef(const uint32_t *in, unsigned log_h, const uint32_t beta[4], uint32_t *out);
/* fri/src/two_adic_pcs.rs:87-114 fold_row (verifier side, used for self-consistency) */
void or_fold_row_ef(size_t index, unsigned log_height, const uint32_t beta[4],
                    const uint32_t e0[4], const uint32_t e1[4], uint32_t out[4]);

/* fri/src/two_adic_pcs.rs:375 mat.dot_ext_powers(alpha): out[r] (EF) = sum_c alpha^c * m[r][c] */
void or_dot_ext_powers(const uint32_t *m, size_t rows, size_t w, const uint32_t alpha[4], uint32_t *out);

/* ---- BfChallenger with Blake3Permutation (basic/src/challenger/mod.rs) ---------------------- */
typedef struct {
    uint8_t state[16][4];
    uint8_t in_buf[16][4];
    int n_in;
    uint8_t out_buf[16][4];
    int n_out;
    int fake_perm; /* 1: fri/tests/fri.rs:37-48 TestPermutation (reverse the 16 words) */
} or_challenger;
void or_chal_init(or_challenger *c, int fake_perm);
void or_chal_observe(or_challenger *c, const uint8_t v[4]);
void or_chal_observe_digest(or_challenger *c, const uint8_t d[32]); /* 8 words, in order */
uint32_t or_chal_sample_bb(or_challenger *c);
void or_chal_sample_ef(or_challenger *c, uint32_t out[4]);
size_t or_chal_sample_bits(or_challenger *c, unsigned bits, int ext);
int or_chal_check_witness(or_challenger *c, unsigned bits, uint32_t witness, int ext);
uint32_t or_chal_grind(or_challenger *c, unsigned bits, int ext); /* smallest valid witness */

/* ---- FRI commit phase (fri/src/prover.rs:93-141), EF codewords, Blake3-Merkle MMCS ---------- */
/* inputs[i] has lens[i] EF elements (4 u32 each), lens strictly sorted descending by the caller.
 * commits: rounds x 32 bytes; layers_out (optional): if non-NULL receives malloc'ed copy of each folded
 * layer *as committed* (round r: lens0>>r EF elements).  Returns number of rounds, or -1 if the
 * final values are not all equal (prover.rs:130-134). */
int or_fri_commit_phase(const uint32_t *const *inputs, const size_t *lens, size_t n_inputs,
                        unsigned log_blowup, or_challenger *chal, uint8_t *commits,
                        uint32_t final_poly[4], uint32_t **layers_out, uint32_t *betas_out);

int or_num_threads(void);
void or_set_num_threads(int n);
void or_splitmix_fill(uint32_t *out, size_t count, uint64_t seed);
void or_merkle_root_from_leaves(const uint8_t *leaves, size_t n, uint8_t root[32]);

#ifdef __cplusplus
}
#endif
#endif
