// sha256.cuh -- the commitment the reference really computes, first slice: BIP-341 tagged SHA-256 leaf
// hashes over the TapTree leaf SCRIPTS and the sorted-pair TapBranch tree with its leaf permutation.
//
//   leaf script   basic/src/tcs/mod.rs:197-225: every leaf of one tree carries the SAME bit-commitment locking scripts
//                 (`use_bcs.clone()`, :252-256) around pushed integers that differ per leaf (its index and its row of
//                 evaluations).  The host hands over that template once: constant byte segments seg[0..n_push] and, per
//                 push, which word of the leaf's row it carries;  script(i) = seg[0] P(i) seg[1] P(x_1) ... seg[n_push],
//                 P = minimal script-number push.  One thread streams its leaf's bytes through SHA-256 -- nothing of the
//                 ~0.7 KB x (width + 1) script is ever materialised (the reference builds every script as a Vec, clones it
//                 once per evaluation, :211-218, and hashes it on one core).
//   TapLeaf       H_TapLeaf(0xc0 || compact_size(len) || script): `NodeInfo::new_leaf_with_ver(script, TapScript)`,
//                 basic/src/tcs/builder.rs:26 [MEM rust-bitcoin]; the 64-byte tag block is a precomputed midstate.
//   TapBranch     H_TapBranch(min(a,b) || max(a,b)), byte-lexicographic: `NodeInfo::combine_with_order`, builder.rs:64;
//                 the kernel records which pairs swapped, and build_tree's index bookkeeping (:66-85) collapses to
//                 position(m) = m ^ mask(m), bit l of mask(m) = "pair m >> (l + 1) of level l swapped" (swapping the two
//                 halves of an aligned range flips exactly bit l of every position in it).
// Digests are kept as the 8 big-endian state words, so byte-lexicographic order is word-wise numeric order.
#pragma once
#include "field.cuh"

namespace sha {

#ifdef TS_EMULATE
static const uint32_t K256[64] = {
#else
__device__ __constant__ uint32_t K256[64] = {
#endif
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

TS_D uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

// one compression: state += F(state, 16 big-endian message words)
TS_D void compress(uint32_t (&st)[8], uint32_t (&w)[16]) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
    TS_UNROLL16
    for (int i = 0; i < 64; i++) {
        uint32_t wi;
        if (i < 16) {
            wi = w[i & 15];
        } else {
            const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            const uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3), s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
            wi = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
            w[i & 15] = wi;
        }
        const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g);
        const uint32_t t1 = h + S1 + ch + K256[i] + wi;
        const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
        const uint32_t t2 = S0 + mj;
        h = g, g = f, f = e, e = d + t1, d = c, c = b, b = a, a = t1 + t2;
    }
    st[0] += a, st[1] += b, st[2] += c, st[3] += d, st[4] += e, st[5] += f, st[6] += g, st[7] += h;
}

// BIP-341 tag midstates: SHA-256 state after the block SHA256(tag) || SHA256(tag)
constexpr uint32_t TAPLEAF_MID[8] = {0x9ce0e4e6u, 0x7c116c39u, 0x38b3caf2u, 0xc30f5089u, 0xd3f3936cu, 0x47636e60u, 0x7db33eeau, 0xddc6f0c9u};
constexpr uint32_t TAPBRANCH_MID[8] = {0x23a865a9u, 0xb8a40da7u, 0x977c1e04u, 0xc49e246fu, 0xb5be1376u, 0x9d24c9b7u, 0xb583b5d4u, 0xa8d226d2u};

constexpr int LEAF_NT = 128;
constexpr int MAX_PUSH = 1 + 4 * 512;  // index + up to 512 extension elements or 2048 base elements per leaf

struct LeafParams {
    const uint32_t *rows;      // n_leaves x width words, Montgomery BabyBear
    uint32_t width;
    size_t n_leaves;
    const uint8_t *segs;       // concatenated constant segments
    const uint32_t *seg_off;   // n_push + 2 offsets into segs
    const uint32_t *push_word; // n_push - 1: word of the row carried by push k + 1 (push 0 is the leaf index)
    uint32_t n_push;
    uint32_t const_len;        // total bytes of all segments
    uint32_t midstate[8];      // SHA-256 state after the 64-byte block SHA256("TapLeaf") || SHA256("TapLeaf")
    uint32_t *out;             // n_leaves x 8 state words
};

// minimal script-number push of a non-negative value < 2^32 (the reference pushes indices and canonical field values):
// OP_0, OP_1..OP_16, or length byte + little-endian magnitude with a zero byte appended when the top bit is set
TS_D uint32_t push_len(uint32_t v) {
    if (v <= 16) return 1;
    uint32_t n = v < 0x100u ? 1 : v < 0x10000u ? 2 : v < 0x1000000u ? 3 : 4;
    if ((v >> (8 * n - 1)) & 1u) n++;
    return 1 + n;
}

struct Stream {  // SHA-256 over a byte stream; the 16-word block buffer lives in shared memory, interleaved by thread
    uint32_t st[8];
    uint32_t *buf;  // buf[w * LEAF_NT]
    uint32_t cur, n;
    TS_D void flush_block() {
        uint32_t w[16];
        TS_UNROLL
        for (int i = 0; i < 16; i++) w[i] = buf[i * LEAF_NT];
        compress(st, w);
    }
    TS_D void byte(uint32_t b) {
        cur = (cur << 8) | b;
        n++;
        if ((n & 3u) == 0) {
            buf[(((n - 1) >> 2) & 15u) * LEAF_NT] = cur;
            if ((n & 63u) == 0) flush_block();
        }
    }
    TS_D void push(uint32_t v) {
        if (v == 0) {
            byte(0x00);
        } else if (v <= 16) {
            byte(0x50 + v);
        } else {
            const uint32_t len = push_len(v) - 1;
            byte(len);
            for (uint32_t i = 0; i < len; i++) byte(i < 4 ? (v >> (8 * i)) & 0xffu : 0u);
        }
    }
    TS_D void finish(uint64_t total_bytes) {
        byte(0x80);
        while ((n & 63u) != 56u) byte(0);
        const uint64_t bits = total_bytes * 8;
        for (int i = 7; i >= 0; i--) byte((uint32_t)(bits >> (8 * i)) & 0xffu);
    }
};

__global__ void __launch_bounds__(LEAF_NT) taptree_leaf_kernel(LeafParams p) {
    TS_DYN_SMEM(uint32_t, sm);  // 16 x LEAF_NT words
    const size_t leaf = (size_t)blockIdx.x * LEAF_NT + threadIdx.x;
    if (leaf >= p.n_leaves) return;
    const uint32_t *row = p.rows + leaf * p.width;
    // script length first: it is part of the hashed prefix (compact size)
    uint32_t slen = p.const_len + push_len((uint32_t)leaf);
    for (uint32_t k = 1; k < p.n_push; k++) slen += push_len(bb::from_monty(row[p.push_word[k - 1]]));
    Stream s;
    TS_UNROLL
    for (int i = 0; i < 8; i++) s.st[i] = p.midstate[i];
    s.buf = sm + threadIdx.x;
    s.cur = 0;
    s.n = 64;  // the tag block is already absorbed
    s.byte(0xc0);  // leaf version TapScript
    uint32_t pre = 1;
    if (slen < 0xfd) {
        s.byte(slen);
        pre += 1;
    } else if (slen <= 0xffff) {
        s.byte(0xfd), s.byte(slen & 0xff), s.byte(slen >> 8);
        pre += 3;
    } else {
        s.byte(0xfe), s.byte(slen & 0xff), s.byte((slen >> 8) & 0xff), s.byte((slen >> 16) & 0xff), s.byte(slen >> 24);
        pre += 5;
    }
    for (uint32_t k = 0; k <= p.n_push; k++) {
        for (uint32_t o = p.seg_off[k]; o < p.seg_off[k + 1]; o++) s.byte(p.segs[o]);
        if (k < p.n_push) s.push(k == 0 ? (uint32_t)leaf : bb::from_monty(row[p.push_word[k - 1]]));
    }
    s.finish((uint64_t)64 + pre + slen);
    TS_UNROLL
    for (int i = 0; i < 8; i++) p.out[leaf * 8 + i] = s.st[i];
}

// one tree level: out[t] = H_TapBranch(min(in[2t], in[2t+1]) || max(..)), swapped[t] = 1 when the right node sorted first
__global__ void __launch_bounds__(256) taptree_branch_kernel(const uint32_t *in, size_t n_pairs, uint32_t *out, uint8_t *swapped) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    uint32_t a[8], b[8];
    TS_UNROLL
    for (int i = 0; i < 8; i++) {
        a[i] = in[(2 * t) * 8 + i];
        b[i] = in[(2 * t + 1) * 8 + i];
    }
    bool left_first = true;  // a <= b
    bool decided = false;
    TS_UNROLL
    for (int i = 0; i < 8; i++)
        if (!decided && a[i] != b[i]) {
            left_first = a[i] < b[i];
            decided = true;
        }
    constexpr uint32_t mid[8] = {0x23a865a9u, 0xb8a40da7u, 0x977c1e04u, 0xc49e246fu, 0xb5be1376u, 0x9d24c9b7u, 0xb583b5d4u, 0xa8d226d2u};  // TAPBRANCH_MID
    uint32_t st[8], w[16];
    TS_UNROLL
    for (int i = 0; i < 8; i++) {
        st[i] = mid[i];
        w[i] = left_first ? a[i] : b[i];
        w[8 + i] = left_first ? b[i] : a[i];
    }
    compress(st, w);
    TS_UNROLL
    for (int i = 0; i < 16; i++) w[i] = 0;
    w[0] = 0x80000000u;
    w[15] = 128 * 8;  // tag block + the two hashes
    compress(st, w);
    TS_UNROLL
    for (int i = 0; i < 8; i++) out[t * 8 + i] = st[i];
    swapped[t] = left_first ? 0 : 1;
}

// leaf_indices[m] = m ^ mask(m): where the Merkle leaf m ended up among the TapTree's leaves (reverse_idx_dict, builder.rs:95-101)
__global__ void taptree_perm_kernel(const uint8_t *swapped, uint32_t log_n, uint32_t *leaf_indices) {
    const size_t n = (size_t)1 << log_n;
    for (size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x; m < n; m += (size_t)gridDim.x * blockDim.x) {
        uint32_t mask = 0;
        size_t off = 0;  // level l's flags start after the n/2 + n/4 + ... flags of the levels below
        for (uint32_t l = 0; l < log_n; l++) {
            mask |= (uint32_t)swapped[off + (m >> (l + 1))] << l;
            off += n >> (l + 1);
        }
        leaf_indices[m] = (uint32_t)m ^ mask;
    }
}

}  // namespace sha
