/*
 * tapstark_oracle.c -- CPU ORACLE (test infrastructure, see tapstark_oracle.h).
 *
 * Plain C restatement of the reference's algorithms on canonical u32 BabyBear values.
 * Never linked into the product library.  Build: see oracle/Makefile.
 */
#include "tapstark_oracle.h"

#include <assert.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define P OR_P

int or_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* SURVEY 8(d) synthetic trace: element i = SplitMix64((seed << 40) + i) mod p, canonical (same as oracle.py splitmix_matrix;
 * here so that a 2^22 x 256 trace does not go through several 8 GiB numpy temporaries). */
void or_splitmix_fill(uint32_t *out, size_t count, uint64_t seed) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < count; i++) {
        uint64_t z = (seed << 40) + (uint64_t)i + 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        out[i] = (uint32_t)(z % P);
    }
}

/* bench.py's CPU arm sets the thread count explicitly: under torch.distributed.run the environment carries
 * OMP_NUM_THREADS=1, which would silently time the port on one core. */
void or_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ============================== BabyBear ============================================= */
/* basic/src/field/mod.rs:43-51 (MOD = 0x78000001); arithmetic is [MEM] p3-baby-bear, exact mod p. */
uint32_t or_bb_add(uint32_t a, uint32_t b) {
    uint32_t s = a + b; /* < 2p < 2^32 */
    return s >= P ? s - P : s;
}
uint32_t or_bb_sub(uint32_t a, uint32_t b) { return a >= b ? a - b : a + P - b; }
uint32_t or_bb_mul(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) % P); }
uint32_t or_bb_pow(uint32_t a, uint64_t e) {
    uint32_t r = 1;
    while (e) {
        if (e & 1) r = or_bb_mul(r, a);
        a = or_bb_mul(a, a);
        e >>= 1;
    }
    return r;
}
uint32_t or_bb_inv(uint32_t a) { return or_bb_pow(a, (uint64_t)P - 2); }
/* [MEM] p3-baby-bear TwoAdicField: generator of the 2^27 subgroup is 0x1a427a41 (= 31^15) */
uint32_t or_two_adic_generator(unsigned bits) {
    assert(bits <= 27);
    uint32_t g = 0x1a427a41u;
    for (unsigned i = bits; i < 27; i++) g = or_bb_mul(g, g);
    return g;
}
uint32_t or_to_monty(uint32_t x) { return (uint32_t)((((uint64_t)x) << 32) % P); }
uint32_t or_from_monty(uint32_t x) {
    /* multiply by 2^-32 mod p */
    static uint32_t rinv = 0;
    if (!rinv) rinv = or_bb_inv(or_to_monty(1));
    return or_bb_mul(x, rinv);
}
void or_to_monty_vec(uint32_t *v, size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) v[i] = or_to_monty(v[i]);
}
void or_from_monty_vec(uint32_t *v, size_t n) {
    uint32_t rinv = or_bb_inv(or_to_monty(1));
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) v[i] = or_bb_mul(v[i], rinv);
}

/* ============================== BabyBear^4 =========================================== */
/* basic/src/field/mod.rs:53-64: BinomialExtensionField<BabyBear,4>; [MEM] W = 11 (x^4 = 11). */
#define EF_W 11u
void or_ef_add(const uint32_t a[4], const uint32_t b[4], uint32_t o[4]) {
    for (int i = 0; i < 4; i++) o[i] = or_bb_add(a[i], b[i]);
}
void or_ef_sub(const uint32_t a[4], const uint32_t b[4], uint32_t o[4]) {
    for (int i = 0; i < 4; i++) o[i] = or_bb_sub(a[i], b[i]);
}
void or_ef_mul(const uint32_t a[4], const uint32_t b[4], uint32_t o[4]) {
    uint32_t r[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) r[i + j] = or_bb_add(r[i + j], or_bb_mul(a[i], b[j]));
    uint32_t t[4];
    for (int i = 0; i < 4; i++) t[i] = r[i];
    for (int i = 4; i < 7; i++) t[i - 4] = or_bb_add(t[i - 4], or_bb_mul(EF_W, r[i]));
    memcpy(o, t, sizeof t);
}
void or_ef_inv(const uint32_t a[4], uint32_t o[4]) {
    /* generic inversion a^(p^4-2); p^4-2 is a 124-bit exponent */
    unsigned __int128 e = (unsigned __int128)P * P;
    e = e * P * P; /* p^4 < 2^124 */
    e -= 2;
    uint32_t r[4] = {1, 0, 0, 0}, b[4];
    memcpy(b, a, sizeof b);
    while (e) {
        if (e & 1) or_ef_mul(r, b, r);
        or_ef_mul(b, b, b);
        e >>= 1;
    }
    memcpy(o, r, sizeof r);
}

/* ============================== bit reversal ========================================= */
static inline size_t brev(size_t x, unsigned bits) {
    size_t r = 0;
    for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}
/* [MEM] p3-matrix reverse_matrix_index_bits / bit_reverse_rows: row r <-> row bitrev(r) */
void or_bit_reverse_rows(uint32_t *mat, unsigned log_h, size_t w) {
    size_t h = (size_t)1 << log_h;
    uint32_t *tmp = (uint32_t *)malloc(w * sizeof(uint32_t));
    for (size_t i = 0; i < h; i++) {
        size_t j = brev(i, log_h);
        if (i < j) {
            memcpy(tmp, mat + i * w, w * 4);
            memcpy(mat + i * w, mat + j * w, w * 4);
            memcpy(mat + j * w, tmp, w * 4);
        }
    }
    free(tmp);
}

/* ============================== DFT family =========================================== */
/* Definition used throughout ([MEM] p3-dft convention, consistent with fri/src/fold_even_odd.rs:75-94):
 *   dft(a)[i] = sum_k a[k] * w_n^(i k),  w_n = two_adic_generator(log n), natural order in and out. */
void or_naive_dft(const uint32_t *in, uint32_t *out, unsigned log_n, size_t w) {
    size_t n = (size_t)1 << log_n;
    uint32_t g = or_two_adic_generator(log_n);
    uint32_t *pw = (uint32_t *)malloc(n * 4);
    pw[0] = 1;
    for (size_t i = 1; i < n; i++) pw[i] = or_bb_mul(pw[i - 1], g);
    for (size_t i = 0; i < n; i++)
        for (size_t c = 0; c < w; c++) {
            uint32_t acc = 0;
            for (size_t k = 0; k < n; k++)
                acc = or_bb_add(acc, or_bb_mul(in[k * w + c], pw[(i * k) & (n - 1)]));
            out[i * w + c] = acc;
        }
    free(pw);
}

/* [MEM] p3-dft Radix2Dit::dft_batch: bit-reverse rows, then log n decimation-in-time layers applied
 * to whole rows.  Result: natural-order DFT of every column. */
static void dit_layers(uint32_t *mat, unsigned log_n, size_t w, uint32_t root) {
    size_t n = (size_t)1 << log_n;
    if (log_n == 0) return;
    uint32_t *tw = (uint32_t *)malloc((n / 2) * 4);
    tw[0] = 1;
    for (size_t i = 1; i < n / 2; i++) tw[i] = or_bb_mul(tw[i - 1], root);
    or_bit_reverse_rows(mat, log_n, w);
    for (unsigned layer = 0; layer < log_n; layer++) {
        size_t half = (size_t)1 << layer;
        size_t stride = n >> (layer + 1); /* twiddle index stride: w_{2half}^j = root^(j*stride) */
#pragma omp parallel for schedule(static)
        for (size_t b = 0; b < n / 2; b++) {
            size_t blk = b / half, j = b % half;
            uint32_t t = tw[j * stride];
            uint32_t *ra = mat + (blk * 2 * half + j) * w;
            uint32_t *rb = ra + half * w;
            for (size_t c = 0; c < w; c++) {
                uint32_t x = ra[c], y = or_bb_mul(rb[c], t);
                ra[c] = or_bb_add(x, y);
                rb[c] = or_bb_sub(x, y);
            }
        }
    }
    free(tw);
}
void or_dft_batch(uint32_t *mat, unsigned log_n, size_t w) {
    dit_layers(mat, log_n, w, or_two_adic_generator(log_n));
}
/* [MEM] TwoAdicSubgroupDft::idft_batch default: dft, reverse rows 1.., scale by 1/n  ==  DFT with w^-1, /n */
void or_idft_batch(uint32_t *mat, unsigned log_n, size_t w) {
    size_t n = (size_t)1 << log_n;
    dit_layers(mat, log_n, w, or_bb_inv(or_two_adic_generator(log_n)));
    uint32_t ninv = or_bb_inv((uint32_t)(n % P));
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n * w; i++) mat[i] = or_bb_mul(mat[i], ninv);
}
/* [MEM] coset_dft_batch default: multiply coefficient row k by shift^k, then dft_batch */
void or_coset_dft_batch(uint32_t *mat, unsigned log_n, size_t w, uint32_t shift) {
    size_t n = (size_t)1 << log_n;
    uint32_t *pw = (uint32_t *)malloc(n * 4);
    pw[0] = 1;
    for (size_t i = 1; i < n; i++) pw[i] = or_bb_mul(pw[i - 1], shift);
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < n; k++)
        for (size_t c = 0; c < w; c++) mat[k * w + c] = or_bb_mul(mat[k * w + c], pw[k]);
    free(pw);
    or_dft_batch(mat, log_n, w);
}
/* [MEM] coset_lde_batch default: idft_batch -> zero-pad to n<<added_bits rows -> coset_dft_batch(shift) */
void or_coset_lde_batch(const uint32_t *in, unsigned log_n, size_t w, unsigned added_bits,
                        uint32_t shift, uint32_t *out) {
    size_t n = (size_t)1 << log_n, N = n << added_bits;
    memcpy(out, in, n * w * 4);
    or_idft_batch(out, log_n, w);
    memset(out + n * w, 0, (N - n) * w * 4);
    or_coset_dft_batch(out, log_n + added_bits, w, shift);
}
void or_pcs_lde_committed(const uint32_t *in, unsigned log_n, size_t w, unsigned added_bits,
                          uint32_t shift, uint32_t *out) {
    or_coset_lde_batch(in, log_n, w, added_bits, shift, out);
    or_bit_reverse_rows(out, log_n + added_bits, w);
}

/* ============================== Blake3 =============================================== */
/* BLAKE3 spec (plain hash, no key), restated from the published reference algorithm.  blake3 crate 1.5
 * is what basic/src/challenger/mod.rs:35-39 calls; KATs: scripts/src/hashes/blake3.rs:537-587. */
static const uint32_t B3_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                  0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t B3_PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
enum { B3_CHUNK_START = 1, B3_CHUNK_END = 2, B3_PARENT = 4, B3_ROOT = 8 };

static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static inline void b3_g(uint32_t *s, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
    s[a] = s[a] + s[b] + mx;
    s[d] = rotr32(s[d] ^ s[a], 16);
    s[c] = s[c] + s[d];
    s[b] = rotr32(s[b] ^ s[c], 12);
    s[a] = s[a] + s[b] + my;
    s[d] = rotr32(s[d] ^ s[a], 8);
    s[c] = s[c] + s[d];
    s[b] = rotr32(s[b] ^ s[c], 7);
}
static void b3_compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter,
                        uint32_t block_len, uint32_t flags, uint32_t out[16]) {
    uint32_t s[16], m[16], t[16];
    for (int i = 0; i < 8; i++) s[i] = cv[i];
    for (int i = 0; i < 4; i++) s[8 + i] = B3_IV[i];
    s[12] = (uint32_t)counter;
    s[13] = (uint32_t)(counter >> 32);
    s[14] = block_len;
    s[15] = flags;
    memcpy(m, block, sizeof m);
    for (int r = 0; r < 7; r++) {
        b3_g(s, 0, 4, 8, 12, m[0], m[1]);
        b3_g(s, 1, 5, 9, 13, m[2], m[3]);
        b3_g(s, 2, 6, 10, 14, m[4], m[5]);
        b3_g(s, 3, 7, 11, 15, m[6], m[7]);
        b3_g(s, 0, 5, 10, 15, m[8], m[9]);
        b3_g(s, 1, 6, 11, 12, m[10], m[11]);
        b3_g(s, 2, 7, 8, 13, m[12], m[13]);
        b3_g(s, 3, 4, 9, 14, m[14], m[15]);
        for (int i = 0; i < 16; i++) t[i] = m[B3_PERM[i]];
        memcpy(m, t, sizeof m);
    }
    for (int i = 0; i < 8; i++) {
        out[i] = s[i] ^ s[i + 8];
        out[i + 8] = s[i + 8] ^ cv[i];
    }
}
static inline uint32_t ld32le(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline void st32le(uint8_t *p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
/* an "output node" awaiting its final flags */
typedef struct {
    uint32_t cv[8];
    uint32_t block[16];
    uint64_t counter;
    uint32_t block_len, flags;
} b3_output;
static void b3_output_cv(const b3_output *o, uint32_t cv[8]) {
    uint32_t out[16];
    b3_compress(o->cv, o->block, o->counter, o->block_len, o->flags, out);
    memcpy(cv, out, 32);
}
/* process one chunk (<=1024 bytes) up to, but not including, its last block */
static b3_output b3_chunk(const uint8_t *in, size_t len, uint64_t chunk_counter) {
    b3_output o;
    memcpy(o.cv, B3_IV, 32);
    o.counter = chunk_counter;
    size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
    uint32_t out[16];
    for (size_t b = 0; b < nblocks; b++) {
        size_t off = b * 64, bl = len - off < 64 ? len - off : 64;
        uint8_t buf[64];
        memset(buf, 0, 64);
        memcpy(buf, in + off, bl);
        for (int i = 0; i < 16; i++) o.block[i] = ld32le(buf + 4 * i);
        o.block_len = (uint32_t)bl;
        o.flags = (b == 0 ? B3_CHUNK_START : 0);
        if (b + 1 == nblocks) {
            o.flags |= B3_CHUNK_END;
            break;
        }
        b3_compress(o.cv, o.block, o.counter, o.block_len, o.flags, out);
        memcpy(o.cv, out, 32);
    }
    return o;
}
static b3_output b3_parent(const uint32_t l[8], const uint32_t r[8]) {
    b3_output o;
    memcpy(o.cv, B3_IV, 32);
    memcpy(o.block, l, 32);
    memcpy(o.block + 8, r, 32);
    o.counter = 0;
    o.block_len = 64;
    o.flags = B3_PARENT;
    return o;
}
void or_blake3(const uint8_t *in, size_t len, uint8_t out[32]) {
    uint32_t stack[54][8];
    int sp = 0;
    uint64_t chunk = 0;
    size_t off = 0;
    /* all chunks but the last are finished and merged into the CV stack */
    while (len - off > 1024) {
        b3_output o = b3_chunk(in + off, 1024, chunk);
        uint32_t cv[8];
        b3_output_cv(&o, cv);
        uint64_t total = chunk + 1;
        while ((total & 1) == 0) {
            b3_output p = b3_parent(stack[--sp], cv);
            b3_output_cv(&p, cv);
            total >>= 1;
        }
        memcpy(stack[sp++], cv, 32);
        chunk++;
        off += 1024;
    }
    b3_output o = b3_chunk(in + off, len - off, chunk);
    while (sp > 0) {
        uint32_t cv[8];
        b3_output_cv(&o, cv);
        o = b3_parent(stack[--sp], cv);
    }
    uint32_t w[16];
    b3_compress(o.cv, o.block, 0, o.block_len, o.flags | B3_ROOT, w);
    for (int i = 0; i < 8; i++) st32le(out + 4 * i, w[i]);
}

/* ============================== MMCS ================================================= */
/* Hash of a list of row slices: [MEM] p3-symmetric SerializingHasher32<Blake3>::hash_iter_slices --
 * every element as canonical u32, little-endian bytes, concatenated, one plain Blake3. */
static void hash_rows(const uint32_t *const *rows, const size_t *widths, size_t k, uint8_t out[32]) {
    size_t total = 0;
    for (size_t i = 0; i < k; i++) total += widths[i];
    uint8_t stackbuf[4096];
    uint8_t *buf = total * 4 <= sizeof stackbuf ? stackbuf : (uint8_t *)malloc(total * 4);
    size_t o = 0;
    for (size_t i = 0; i < k; i++)
        for (size_t c = 0; c < widths[i]; c++, o += 4) st32le(buf + o, rows[i][c]);
    or_blake3(buf, total * 4, out);
    if (buf != stackbuf) free(buf);
}
/* [MEM] CompressionFunctionFromHasher<u8,Blake3,2,32>: blake3(left || right), plain hash */
static void compress2(const uint8_t l[32], const uint8_t r[32], uint8_t out[32]) {
    uint8_t buf[64];
    memcpy(buf, l, 32);
    memcpy(buf + 32, r, 32);
    or_blake3(buf, 64, out);
}
static unsigned log2_strict(size_t x) {
    unsigned l = 0;
    while (((size_t)1 << l) < x) l++;
    assert(((size_t)1 << l) == x && "heights must be powers of two");
    return l;
}

struct or_tree {
    size_t k;
    const uint32_t **mats;
    size_t *heights, *widths;
    size_t *order; /* indices sorted by height desc, stable */
    int layout;
    size_t n_layers;
    uint8_t **layers; /* layers[0] = leaf digests (max_height), last = root */
    size_t *layer_len;
};

static void sort_order(const size_t *heights, size_t k, size_t *order) {
    /* stable insertion sort by height descending: sorted_by_key(Reverse(height)),
     * basic/src/tcs/mod.rs:344 and [MEM] FieldMerkleTree::new */
    for (size_t i = 0; i < k; i++) order[i] = i;
    for (size_t i = 1; i < k; i++) {
        size_t x = order[i], j = i;
        while (j > 0 && heights[order[j - 1]] < heights[x]) {
            order[j] = order[j - 1];
            j--;
        }
        order[j] = x;
    }
}

size_t or_padded_leaf(const uint32_t *const *mats, const size_t *heights, const size_t *widths,
                      size_t k, size_t leaf, uint32_t *out) {
    /* basic/src/tcs/mod.rs:339-378: leaf i gets, for every matrix largest first, row i >> (log_max-log_h) */
    size_t *order = (size_t *)malloc(k * sizeof(size_t));
    sort_order(heights, k, order);
    unsigned lmax = log2_strict(heights[order[0]]);
    size_t o = 0;
    for (size_t q = 0; q < k; q++) {
        size_t m = order[q];
        size_t row = leaf >> (lmax - log2_strict(heights[m]));
        for (size_t c = 0; c < widths[m]; c++) out[o++] = mats[m][row * widths[m] + c];
    }
    free(order);
    return o;
}

or_tree *or_mmcs_commit(const uint32_t *const *mats, const size_t *heights, const size_t *widths,
                        size_t k, int layout, uint8_t root[32]) {
    assert(k > 0);
    or_tree *t = (or_tree *)calloc(1, sizeof *t);
    t->k = k;
    t->layout = layout;
    t->mats = (const uint32_t **)malloc(k * sizeof(void *));
    t->heights = (size_t *)malloc(k * sizeof(size_t));
    t->widths = (size_t *)malloc(k * sizeof(size_t));
    t->order = (size_t *)malloc(k * sizeof(size_t));
    memcpy(t->mats, mats, k * sizeof(void *));
    memcpy(t->heights, heights, k * sizeof(size_t));
    memcpy(t->widths, widths, k * sizeof(size_t));
    sort_order(heights, k, t->order);
    size_t hmax = heights[t->order[0]];
    unsigned lmax = log2_strict(hmax);
    t->n_layers = lmax + 1;
    t->layers = (uint8_t **)malloc(t->n_layers * sizeof(void *));
    t->layer_len = (size_t *)malloc(t->n_layers * sizeof(size_t));
    for (size_t l = 0; l < t->n_layers; l++) {
        t->layer_len[l] = hmax >> l;
        t->layers[l] = (uint8_t *)malloc(32 * t->layer_len[l]);
    }
    /* leaf layer */
    size_t n_first = 0; /* number of (sorted) matrices hashed into the leaf layer */
    if (layout == OR_LAYOUT_PADDED) n_first = k;
    else
        while (n_first < k && heights[t->order[n_first]] == hmax) n_first++;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < hmax; i++) {
        const uint32_t *rows[64];
        size_t ws[64];
        assert(n_first <= 64);
        for (size_t q = 0; q < n_first; q++) {
            size_t m = t->order[q];
            size_t row = i >> (lmax - log2_strict(heights[m]));
            rows[q] = mats[m] + row * widths[m];
            ws[q] = widths[m];
        }
        hash_rows(rows, ws, n_first, t->layers[0] + 32 * i);
    }
    /* upper layers; [MEM] FieldMerkleTree compress_and_inject */
    size_t next_mat = n_first;
    for (size_t l = 1; l < t->n_layers; l++) {
        size_t len = t->layer_len[l];
        size_t inj0 = next_mat;
        if (layout == OR_LAYOUT_P3_INJECT)
            while (next_mat < k && heights[t->order[next_mat]] == len) next_mat++;
        size_t n_inj = next_mat - inj0;
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < len; i++) {
            uint8_t d[32];
            compress2(t->layers[l - 1] + 64 * i, t->layers[l - 1] + 64 * i + 32, d);
            if (n_inj) {
                const uint32_t *rows[64];
                size_t ws[64];
                uint8_t rd[32];
                for (size_t q = 0; q < n_inj; q++) {
                    size_t m = t->order[inj0 + q];
                    rows[q] = mats[m] + i * widths[m];
                    ws[q] = widths[m];
                }
                hash_rows(rows, ws, n_inj, rd);
                compress2(d, rd, t->layers[l] + 32 * i);
            } else
                memcpy(t->layers[l] + 32 * i, d, 32);
        }
    }
    assert(layout == OR_LAYOUT_PADDED || next_mat == k);
    memcpy(root, t->layers[t->n_layers - 1], 32);
    return t;
}
size_t or_tree_depth(const or_tree *t) { return t->n_layers - 1; }
size_t or_tree_num_layers(const or_tree *t) { return t->n_layers; }
const uint8_t *or_tree_layer(const or_tree *t, size_t layer, size_t *len) {
    if (len) *len = t->layer_len[layer];
    return t->layers[layer];
}
/* Root of the plain 2-to-1 tree over n = 2^k given leaf digests (parent = blake3(left || right), [MEM]
 * CompressionFunctionFromHasher): lets a test rebuild the top of a commitment from a leaf layer that was produced
 * elsewhere (the device) without re-hashing 16 GiB of rows on the CPU. */
void or_merkle_root_from_leaves(const uint8_t *leaves, size_t n, uint8_t root[32]) {
    uint8_t *cur = (uint8_t *)malloc(32 * n);
    memcpy(cur, leaves, 32 * n);
    while (n > 1) {
        const size_t half = n / 2;
        uint8_t *nxt = (uint8_t *)malloc(32 * half);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < half; i++) compress2(cur + 64 * i, cur + 64 * i + 32, nxt + 32 * i);
        free(cur);
        cur = nxt;
        n = half;
    }
    memcpy(root, cur, 32);
    free(cur);
}

void or_mmcs_open_batch(const or_tree *t, size_t index, uint32_t *rows_out, uint8_t *path_out) {
    /* [MEM] FieldMerkleTreeMmcs::open_batch; same index semantics as basic/src/mmcs/bf_mmcs.rs:11-16 */
    unsigned lmax = log2_strict(t->heights[t->order[0]]);
    size_t o = 0;
    for (size_t m = 0; m < t->k; m++) {
        size_t row = index >> (lmax - log2_strict(t->heights[m]));
        memcpy(rows_out + o, t->mats[m] + row * t->widths[m], t->widths[m] * 4);
        o += t->widths[m];
    }
    for (size_t l = 0; l + 1 < t->n_layers; l++)
        memcpy(path_out + 32 * l, t->layers[l] + 32 * ((index >> l) ^ 1), 32);
}
int or_mmcs_verify_batch(const size_t *heights, const size_t *widths, size_t k, int layout,
                         size_t index, const uint32_t *rows, const uint8_t *path, size_t depth,
                         const uint8_t root[32]) {
    size_t *order = (size_t *)malloc(k * sizeof(size_t));
    size_t *offs = (size_t *)malloc(k * sizeof(size_t));
    sort_order(heights, k, order);
    size_t o = 0;
    for (size_t m = 0; m < k; m++) {
        offs[m] = o;
        o += widths[m];
    }
    size_t hmax = heights[order[0]];
    if (((size_t)1 << depth) != hmax) {
        free(order); free(offs);
        return 0;
    }
    const uint32_t *rp[64];
    size_t ws[64];
    size_t n_first = 0;
    if (layout == OR_LAYOUT_PADDED) n_first = k;
    else
        while (n_first < k && heights[order[n_first]] == hmax) n_first++;
    for (size_t q = 0; q < n_first; q++) {
        rp[q] = rows + offs[order[q]];
        ws[q] = widths[order[q]];
    }
    uint8_t cur[32];
    hash_rows(rp, ws, n_first, cur);
    size_t next_mat = n_first, len = hmax;
    for (size_t l = 0; l < depth; l++) {
        const uint8_t *sib = path + 32 * l;
        uint8_t d[32];
        if (index & 1) compress2(sib, cur, d);
        else compress2(cur, sib, d);
        index >>= 1;
        len >>= 1;
        size_t inj0 = next_mat;
        if (layout == OR_LAYOUT_P3_INJECT)
            while (next_mat < k && heights[order[next_mat]] == len) next_mat++;
        if (next_mat > inj0) {
            uint8_t rd[32];
            for (size_t q = inj0; q < next_mat; q++) {
                rp[q - inj0] = rows + offs[order[q]];
                ws[q - inj0] = widths[order[q]];
            }
            hash_rows(rp, ws, next_mat - inj0, rd);
            compress2(d, rd, cur);
        } else
            memcpy(cur, d, 32);
    }
    free(order); free(offs);
    return memcmp(cur, root, 32) == 0;
}
void or_tree_free(or_tree *t) {
    if (!t) return;
    for (size_t l = 0; l < t->n_layers; l++) free(t->layers[l]);
    free(t->layers); free(t->layer_len); free(t->mats); free(t->heights); free(t->widths); free(t->order);
    free(t);
}

/* ============================== FRI fold ============================================= */
/* fri/src/two_adic_pcs.rs:116-147, literally: powers[j] = (beta/2) * g_inv^j, bit-reversed, then
 * out[i] = (1/2 + powers[i]) * lo + (1/2 - powers[i]) * hi. */
void or_fold_matrix_bb(const uint32_t *in, unsigned log_h, uint32_t beta, uint32_t *out) {
    size_t h = (size_t)1 << log_h;
    uint32_t g_inv = or_bb_inv(or_two_adic_generator(log_h + 1));
    uint32_t one_half = or_bb_inv(2);
    uint32_t half_beta = or_bb_mul(beta, one_half);
    uint32_t *powers = (uint32_t *)malloc(h * 4);
    uint32_t cur = half_beta;
    for (size_t j = 0; j < h; j++) {
        powers[brev(j, log_h)] = cur; /* shifted_powers + reverse_slice_index_bits */
        cur = or_bb_mul(cur, g_inv);
    }
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < h; i++) {
        uint32_t lo = in[2 * i], hi = in[2 * i + 1];
        out[i] = or_bb_add(or_bb_mul(or_bb_add(one_half, powers[i]), lo),
                           or_bb_mul(or_bb_sub(one_half, powers[i]), hi));
    }
    free(powers);
}
void or_fold_matrix_ef(const uint32_t *in, unsigned log_h, const uint32_t beta[4], uint32_t *out) {
    size_t h = (size_t)1 << log_h;
    uint32_t g_inv = or_bb_inv(or_two_adic_generator(log_h + 1));
    uint32_t one_half[4] = {or_bb_inv(2), 0, 0, 0};
    uint32_t half_beta[4];
    or_ef_mul(beta, one_half, half_beta);
    uint32_t *powers = (uint32_t *)malloc(h * 16);
    uint32_t cur[4];
    memcpy(cur, half_beta, 16);
    for (size_t j = 0; j < h; j++) {
        memcpy(powers + 4 * brev(j, log_h), cur, 16);
        for (int c = 0; c < 4; c++) cur[c] = or_bb_mul(cur[c], g_inv);
    }
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < h; i++) {
        const uint32_t *lo = in + 8 * i, *hi = in + 8 * i + 4;
        uint32_t a[4], b[4], x[4], y[4];
        or_ef_add(one_half, powers + 4 * i, a);
        or_ef_sub(one_half, powers + 4 * i, b);
        or_ef_mul(a, lo, x);
        or_ef_mul(b, hi, y);
        or_ef_add(x, y, out + 4 * i);
    }
    free(powers);
}
/* fri/src/two_adic_pcs.rs:87-114 */
void or_fold_row_ef(size_t index, unsigned log_height, const uint32_t beta[4],
                    const uint32_t e0[4], const uint32_t e1[4], uint32_t out[4]) {
    uint32_t start = or_bb_pow(or_two_adic_generator(log_height + 1), brev(index, log_height));
    uint32_t xs0[4] = {start, 0, 0, 0};
    uint32_t xs1[4] = {or_bb_mul(start, or_two_adic_generator(1)), 0, 0, 0};
    uint32_t num[4], den[4], deninv[4], d[4], t[4];
    or_ef_sub(beta, xs0, num);
    or_ef_sub(e1, e0, d);
    or_ef_sub(xs1, xs0, den);
    or_ef_inv(den, deninv);
    or_ef_mul(num, d, t);
    or_ef_mul(t, deninv, t);
    or_ef_add(e0, t, out);
}

/* ============================== dot_ext_powers ======================================= */
/* fri/src/two_adic_pcs.rs:375 `mat.dot_ext_powers(alpha)` ([MEM] p3-matrix): out[r] = sum_c alpha^c * m[r][c] */
void or_dot_ext_powers(const uint32_t *m, size_t rows, size_t w, const uint32_t alpha[4], uint32_t *out) {
    uint32_t *apow = (uint32_t *)malloc(w * 16);
    uint32_t cur[4] = {1, 0, 0, 0};
    for (size_t c = 0; c < w; c++) {
        memcpy(apow + 4 * c, cur, 16);
        or_ef_mul(cur, alpha, cur);
    }
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < rows; r++) {
        uint64_t acc[4] = {0, 0, 0, 0};
        for (size_t c = 0; c < w; c++) {
            const uint64_t v = m[r * w + c];
            for (int k = 0; k < 4; k++) acc[k] = (acc[k] + v * apow[4 * c + k]) % P;
        }
        for (int k = 0; k < 4; k++) out[4 * r + k] = (uint32_t)acc[k];
    }
    free(apow);
}

/* ============================== Challenger =========================================== */
/* basic/src/challenger/mod.rs:22-49 (Blake3Permutation), :151-174 (duplexing), :183-194 (observe),
 * :261-313 (sample), :341-348 (sample_bits), :95-114 (grind / check_witness);
 * basic/src/challenger/chan_field.rs:12-18 (from_pf: u32 LE % p), :35-42 (mod_p = 1 << 12). */
void or_chal_init(or_challenger *c, int fake_perm) {
    memset(c, 0, sizeof *c);
    c->fake_perm = fake_perm;
}
static void chal_permute(or_challenger *c) {
    if (c->fake_perm) { /* fri/tests/fri.rs:37-48 */
        for (int i = 0; i < 8; i++) {
            uint8_t t[4];
            memcpy(t, c->state[i], 4);
            memcpy(c->state[i], c->state[15 - i], 4);
            memcpy(c->state[15 - i], t, 4);
        }
        return;
    }
    uint8_t h[32];
    or_blake3(&c->state[0][0], 64, h);
    memset(c->state, 0, 32);
    memcpy(&c->state[8][0], h, 32);
}
static void chal_duplex(or_challenger *c) {
    for (int i = 0; i < c->n_in; i++) memcpy(c->state[i], c->in_buf[i], 4);
    c->n_in = 0;
    chal_permute(c);
    c->n_out = 8;
    for (int i = 0; i < 8; i++) memcpy(c->out_buf[i], c->state[8 + i], 4);
}
void or_chal_observe(or_challenger *c, const uint8_t v[4]) {
    c->n_out = 0;
    memcpy(c->in_buf[c->n_in++], v, 4);
    if (c->n_in == 8) chal_duplex(c);
}
void or_chal_observe_digest(or_challenger *c, const uint8_t d[32]) {
    for (int i = 0; i < 8; i++) or_chal_observe(c, d + 4 * i);
}
uint32_t or_chal_sample_bb(or_challenger *c) {
    if (c->n_in != 0 || c->n_out == 0) chal_duplex(c);
    const uint8_t *w = c->out_buf[--c->n_out]; /* Vec::pop: last word first */
    return ld32le(w) % P;
}
void or_chal_sample_ef(or_challenger *c, uint32_t out[4]) {
    for (int i = 0; i < 4; i++) out[i] = or_chal_sample_bb(c);
}
size_t or_chal_sample_bits(or_challenger *c, unsigned bits, int ext) {
    uint32_t v;
    if (ext) {
        uint32_t e[4];
        or_chal_sample_ef(c, e);
        v = e[0];
    } else
        v = or_chal_sample_bb(c);
    return (size_t)((uint64_t)v >> (32 - bits));
}
int or_chal_check_witness(or_challenger *c, unsigned bits, uint32_t witness, int ext) {
    uint8_t w[4], z[4] = {0, 0, 0, 0};
    st32le(w, witness);
    or_chal_observe(c, w);
    for (int i = 0; i < 7; i++) or_chal_observe(c, z);
    return or_chal_sample_bits(c, bits, ext) == 0;
}
uint32_t or_chal_grind(or_challenger *c, unsigned bits, int ext) {
    /* reference: rayon find_any over 0..4096 (non-deterministic); oracle fixes the smallest witness */
    for (uint32_t w = 0; w < 4096; w++) {
        or_challenger t = *c;
        if (or_chal_check_witness(&t, bits, w, ext)) {
            int ok = or_chal_check_witness(c, bits, w, ext);
            assert(ok);
            (void)ok;
            return w;
        }
    }
    return 0xFFFFFFFFu;
}

/* ============================== FRI commit phase ===================================== */
int or_fri_commit_phase(const uint32_t *const *inputs, const size_t *lens, size_t n_inputs,
                        unsigned log_blowup, or_challenger *chal, uint8_t *commits,
                        uint32_t final_poly[4], uint32_t **layers_out, uint32_t *betas_out) {
    size_t len = lens[0], next_in = 1;
    uint32_t *folded = (uint32_t *)malloc(len * 16);
    memcpy(folded, inputs[0], len * 16);
    int rounds = 0;
    while (len > ((size_t)1 << log_blowup)) {
        /* leaves = RowMajorMatrix::new(folded.clone(), 2); commit_matrix (prover.rs:112-113) */
        size_t h = len / 2, w = 8;
        const uint32_t *m = folded;
        uint8_t root[32];
        or_tree *t = or_mmcs_commit(&m, &h, &w, 1, OR_LAYOUT_P3_INJECT, root);
        or_tree_free(t);
        memcpy(commits + 32 * rounds, root, 32);
        if (layers_out) {
            layers_out[rounds] = (uint32_t *)malloc(len * 16);
            memcpy(layers_out[rounds], folded, len * 16);
        }
        or_chal_observe_digest(chal, root); /* prover.rs:114 */
        uint32_t beta[4];
        or_chal_sample_ef(chal, beta); /* prover.rs:116 */
        if (betas_out) memcpy(betas_out + 4 * rounds, beta, 16);
        uint32_t *nf = (uint32_t *)malloc(h * 16);
        or_fold_matrix_ef(folded, log2_strict(h), beta, nf); /* prover.rs:119 */
        free(folded);
        folded = nf;
        len = h;
        rounds++;
        if (next_in < n_inputs && lens[next_in] == len) { /* prover.rs:124-126 */
            for (size_t i = 0; i < len * 4; i++) folded[i] = or_bb_add(folded[i], inputs[next_in][i]);
            next_in++;
        }
    }
    memcpy(final_poly, folded, 16);
    int ok = 1;
    for (size_t i = 0; i < len; i++)
        if (memcmp(folded + 4 * i, final_poly, 16) != 0) ok = 0; /* prover.rs:130-134 */
    free(folded);
    return ok ? rounds : -1;
}

/* ---- SHA-256 (FIPS 180-4) + BIP-341 tagged hashes ------------------------------------------------------------------ */
static const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
static uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void sha_block(uint32_t h[8], const uint8_t b[64]) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)b[4 * i] << 24) | ((uint32_t)b[4 * i + 1] << 16) | ((uint32_t)b[4 * i + 2] << 8) | b[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        const uint32_t s0 = sha_rotr(w[i - 15], 7) ^ sha_rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        const uint32_t s1 = sha_rotr(w[i - 2], 17) ^ sha_rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t v[8];
    memcpy(v, h, 32);
    for (int i = 0; i < 64; i++) {
        const uint32_t S1 = sha_rotr(v[4], 6) ^ sha_rotr(v[4], 11) ^ sha_rotr(v[4], 25), ch = (v[4] & v[5]) ^ (~v[4] & v[6]);
        const uint32_t t1 = v[7] + S1 + ch + SHA_K[i] + w[i];
        const uint32_t S0 = sha_rotr(v[0], 2) ^ sha_rotr(v[0], 13) ^ sha_rotr(v[0], 22), mj = (v[0] & v[1]) ^ (v[0] & v[2]) ^ (v[1] & v[2]);
        v[7] = v[6], v[6] = v[5], v[5] = v[4], v[4] = v[3] + t1, v[3] = v[2], v[2] = v[1], v[1] = v[0], v[0] = t1 + S0 + mj;
    }
    for (int i = 0; i < 8; i++) h[i] += v[i];
}
/* SHA-256 of prefix (may be empty) followed by data */
static void sha256_two(const uint8_t *pre, size_t pre_len, const uint8_t *data, size_t len, uint8_t out[32]) {
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint8_t buf[64];
    size_t fill = 0;
    const uint64_t total = (uint64_t)pre_len + len;
    for (int part = 0; part < 2; part++) {
        const uint8_t *p = part ? data : pre;
        const size_t n = part ? len : pre_len;
        for (size_t i = 0; i < n; i++) {
            buf[fill++] = p[i];
            if (fill == 64) sha_block(h, buf), fill = 0;
        }
    }
    buf[fill++] = 0x80;
    if (fill > 56) {
        memset(buf + fill, 0, 64 - fill);
        sha_block(h, buf);
        fill = 0;
    }
    memset(buf + fill, 0, 56 - fill);
    for (int i = 0; i < 8; i++) buf[56 + i] = (uint8_t)((total * 8) >> (56 - 8 * i));
    sha_block(h, buf);
    for (int i = 0; i < 8; i++) out[4 * i] = (uint8_t)(h[i] >> 24), out[4 * i + 1] = (uint8_t)(h[i] >> 16), out[4 * i + 2] = (uint8_t)(h[i] >> 8), out[4 * i + 3] = (uint8_t)h[i];
}
void or_sha256(const uint8_t *data, size_t len, uint8_t out[32]) { sha256_two(NULL, 0, data, len, out); }
static void tagged_hash(const char *tag, const uint8_t *pre, size_t pre_len, const uint8_t *data, size_t len, uint8_t out[32]) {
    uint8_t t[32];
    or_sha256((const uint8_t *)tag, strlen(tag), t);
    uint8_t *m = (uint8_t *)malloc(64 + pre_len + len + 1);
    memcpy(m, t, 32), memcpy(m + 32, t, 32);
    if (pre_len) memcpy(m + 64, pre, pre_len);
    if (len) memcpy(m + 64 + pre_len, data, len);
    or_sha256(m, 64 + pre_len + len, out);
    free(m);
}
void or_tap_leaf_hash(const uint8_t *script, size_t len, uint8_t out[32]) {
    uint8_t pre[6] = {0xc0};
    size_t n = 1;
    if (len < 0xfd) pre[n++] = (uint8_t)len;
    else if (len <= 0xffff) pre[n++] = 0xfd, pre[n++] = (uint8_t)len, pre[n++] = (uint8_t)(len >> 8);
    else pre[n++] = 0xfe, pre[n++] = (uint8_t)len, pre[n++] = (uint8_t)(len >> 8), pre[n++] = (uint8_t)(len >> 16), pre[n++] = (uint8_t)(len >> 24);
    tagged_hash("TapLeaf", pre, n, script, len, out);
}
void or_tap_branch_hash(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
    const int a_first = memcmp(a, b, 32) <= 0;
    tagged_hash("TapBranch", a_first ? a : b, 32, a_first ? b : a, 32, out);
}
