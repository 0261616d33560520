// host_side.h -- the host-resident parts of the path: the Fiat-Shamir challenger and the verifier-side
// Merkle check.  In the reference these run on the host too (a few hundred bytes per FRI round), so this is
// the product's host logic, not a fallback for device work.
//
//   BfChallenger<F, U32, Blake3Permutation, 16>   basic/src/challenger/mod.rs:22-49,95-114,151-194,261-348
//   ChallengeField::from_pf (u32 LE mod p)         basic/src/challenger/chan_field.rs:12-18
//   PermutationField::mod_p (= 1 << 12)            basic/src/challenger/chan_field.rs:35-42
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace hostb3 {
static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                               0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
enum : uint32_t { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };
inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
inline void g(uint32_t *s, int a, int b, int c, int d, uint32_t x, uint32_t y) {
    s[a] += s[b] + x; s[d] = rotr(s[d] ^ s[a], 16);
    s[c] += s[d];     s[b] = rotr(s[b] ^ s[c], 12);
    s[a] += s[b] + y; s[d] = rotr(s[d] ^ s[a], 8);
    s[c] += s[d];     s[b] = rotr(s[b] ^ s[c], 7);
}
// full 16-word output of the compression function
inline void compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter, uint32_t len,
                     uint32_t flags, uint32_t out[16]) {
    static const int perm[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
    uint32_t s[16], m[16], t[16];
    for (int i = 0; i < 8; i++) s[i] = cv[i];
    for (int i = 0; i < 4; i++) s[8 + i] = IV[i];
    s[12] = (uint32_t)counter; s[13] = (uint32_t)(counter >> 32); s[14] = len; s[15] = flags;
    memcpy(m, block, 64);
    for (int r = 0; r < 7; r++) {
        g(s, 0, 4, 8, 12, m[0], m[1]);   g(s, 1, 5, 9, 13, m[2], m[3]);
        g(s, 2, 6, 10, 14, m[4], m[5]);  g(s, 3, 7, 11, 15, m[6], m[7]);
        g(s, 0, 5, 10, 15, m[8], m[9]);  g(s, 1, 6, 11, 12, m[10], m[11]);
        g(s, 2, 7, 8, 13, m[12], m[13]); g(s, 3, 4, 9, 14, m[14], m[15]);
        for (int i = 0; i < 16; i++) t[i] = m[perm[i]];
        memcpy(m, t, 64);
    }
    for (int i = 0; i < 8; i++) { out[i] = s[i] ^ s[i + 8]; out[i + 8] = s[i + 8] ^ cv[i]; }
}
struct Node { uint32_t cv[8], block[16]; uint64_t counter; uint32_t len, flags; };
inline void node_cv(const Node &n, uint32_t cv[8]) {
    uint32_t o[16];
    compress(n.cv, n.block, n.counter, n.len, n.flags, o);
    memcpy(cv, o, 32);
}
inline Node chunk_node(const uint8_t *in, size_t len, uint64_t counter) {
    Node n;
    memcpy(n.cv, IV, 32);
    n.counter = counter;
    size_t nb = len == 0 ? 1 : (len + 63) / 64;
    for (size_t b = 0; b < nb; b++) {
        size_t off = b * 64, bl = len - off < 64 ? len - off : 64;
        uint8_t buf[64] = {0};
        memcpy(buf, in + off, bl);
        memcpy(n.block, buf, 64);  // little-endian host
        n.len = (uint32_t)bl;
        n.flags = b == 0 ? CHUNK_START : 0;
        if (b + 1 == nb) { n.flags |= CHUNK_END; break; }
        uint32_t o[16];
        compress(n.cv, n.block, n.counter, n.len, n.flags, o);
        memcpy(n.cv, o, 32);
    }
    return n;
}
inline Node parent_node(const uint32_t l[8], const uint32_t r[8]) {
    Node n;
    memcpy(n.cv, IV, 32);
    memcpy(n.block, l, 32);
    memcpy(n.block + 8, r, 32);
    n.counter = 0; n.len = 64; n.flags = PARENT;
    return n;
}
inline void hash(const uint8_t *in, size_t len, uint8_t out[32]) {
    uint32_t stack[54][8];
    int sp = 0;
    uint64_t chunk = 0;
    size_t off = 0;
    while (len - off > 1024) {
        Node n = chunk_node(in + off, 1024, chunk);
        uint32_t cv[8];
        node_cv(n, cv);
        for (uint64_t total = chunk + 1; (total & 1) == 0; total >>= 1) {
            Node p = parent_node(stack[--sp], cv);
            node_cv(p, cv);
        }
        memcpy(stack[sp++], cv, 32);
        chunk++;
        off += 1024;
    }
    Node n = chunk_node(in + off, len - off, chunk);
    while (sp > 0) {
        uint32_t cv[8];
        node_cv(n, cv);
        n = parent_node(stack[--sp], cv);
    }
    uint32_t o[16];
    compress(n.cv, n.block, 0, n.len, n.flags | ROOT, o);
    memcpy(out, o, 32);
}
}  // namespace hostb3

struct ts_challenger {
    uint8_t state[16][4];
    std::vector<uint32_t> in_buf;   // pending observed words (LE u32)
    std::vector<uint32_t> out_buf;  // squeezed words, popped from the back
    ts_challenger() { memset(state, 0, sizeof state); }

    void duplex() {
        for (size_t i = 0; i < in_buf.size(); i++) memcpy(state[i], &in_buf[i], 4);
        in_buf.clear();
        uint8_t h[32];
        hostb3::hash(&state[0][0], 64, h);  // Blake3Permutation: mod.rs:34-48
        memset(state, 0, 32);
        memcpy(&state[8][0], h, 32);
        out_buf.resize(8);
        memcpy(out_buf.data(), &state[8][0], 32);
    }
    void observe(const uint8_t w[4]) {
        out_buf.clear();
        uint32_t v;
        memcpy(&v, w, 4);
        in_buf.push_back(v);
        if (in_buf.size() == 8) duplex();
    }
    uint32_t sample_base() {
        if (!in_buf.empty() || out_buf.empty()) duplex();
        uint32_t v = out_buf.back();
        out_buf.pop_back();
        return v % 0x78000001u;
    }
    size_t sample_bits(unsigned bits, bool ext) {
        uint32_t v = sample_base();
        if (ext) for (int i = 0; i < 3; i++) (void)sample_base();
        return (size_t)((uint64_t)v >> (32 - bits));
    }
    bool check_witness(unsigned bits, uint32_t witness, bool ext) {
        uint8_t w[4], z[4] = {0, 0, 0, 0};
        memcpy(w, &witness, 4);
        observe(w);
        for (int i = 0; i < 7; i++) observe(z);
        return sample_bits(bits, ext) == 0;
    }
};
