// hash.cuh -- Blake3 row hashing and the 2-to-1 Blake3 tree (sm_100a).
//
// Stands in for BFMmcs::commit (basic/src/mmcs/bf_mmcs.rs:23).  Construction ([MEM] Plonky3
// FieldMerkleTreeMmcs<Val, u8, SerializingHasher32<Blake3>, CompressionFunctionFromHasher<u8,Blake3,2,32>, 32>,
// the shape sketched in the commented-out uni-stark/tests/mul_air.rs:284-287):
//   leaf digest  = Blake3( canonical-u32 little-endian bytes of the row(s) )      (plain hash mode)
//   parent       = Blake3( left || right )                                        (plain 64-byte hash)
// Blake3 itself is the function the reference uses in basic/src/challenger/mod.rs:35-39 and pins with
// KATs in scripts/src/hashes/blake3.rs:537-587.
//
// The compression function runs entirely in registers, one leaf per thread.  Message words are staged
// through shared memory: the CTA reads row segments coalesced (consecutive threads -> consecutive words of
// one row), converts Montgomery -> canonical once, and each thread then reads its own leaf's words with a
// conflict-free stride (odd pitch).
#pragma once
#include "field.cuh"

namespace b3 {

enum : uint32_t { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

TS_D uint32_t rotr(uint32_t x, int n) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}

// Blake3 is bound by the alu pipe (LOP3 + SHF: 8 per G, and ptxas also puts the 6 additions there as IADD3/IADD:
// ncu showed alu 91.5 % / fma 16 % active).  The additions are therefore issued as IMAD x * 1 + y on the otherwise
// idle fma pipe; the multiplier 1 comes from constant memory so that ptxas cannot fold it back into an IADD.
#if !defined(TS_EMULATE) && !defined(TS_B3_NO_FMA_ADDS)
__constant__ uint32_t c_one = 1;
TS_D uint32_t fadd(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(c_one), "r"(b));
    return r;
#else
    return a + b;
#endif
}
#else
TS_D uint32_t fadd(uint32_t a, uint32_t b) { return a + b; }
#endif

// A rotation on the fma pipe: x * 2^(32-n) as a 64-bit product holds the bits shifted out in its high word, and high + low is the
// rotated word (no common bits).  Two fma instructions (IMAD.WIDE, IMAD) instead of one alu instruction (SHF): it pays only for
// as many rotations as it takes to level the two pipes -- per G 8 alu + 6 fma instructions, a warp instruction occupying its
// 16-lane pipe for two cycles, so moving ONE rotation in every second G gives alu 7.5 / fma 7 / issue 14.5 per G.
// TS_B3_ROT_FMA = how many of the 8 G of a round do that with their rotation by 16.  Measured on the 2^24 x 1 KiB leaf hash
// (profiles/r02/README.md): 0: 10.02 ms, 1: 9.96, **2: 9.61**, 3: 10.66, 4: 9.98 -- two, as the pipe arithmetic says.
#ifndef TS_B3_ROT_FMA
#define TS_B3_ROT_FMA 2
#endif
#if !defined(TS_EMULATE) && !defined(TS_B3_NO_FMA_ADDS)
TS_D uint32_t rot16_fma(uint32_t x) {
#if defined(__CUDA_ARCH__)
    uint32_t lo, hi, r;
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, 65536;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(x));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(hi), "r"(c_one), "r"(lo));
    return r;
#else
    return rotr(x, 16);
#endif
}
#else
TS_D uint32_t rot16_fma(uint32_t x) { return rotr(x, 16); }
#endif

#define TS_B3_G_(a, b, c, d, mx, my, ROT16) \
    a = fadd(fadd(a, b), (mx));             \
    d = ROT16(d ^ a);                       \
    c = fadd(c, d);                         \
    b = rotr(b ^ c, 12);                    \
    a = fadd(fadd(a, b), (my));             \
    d = rotr(d ^ a, 8);                     \
    c = fadd(c, d);                         \
    b = rotr(b ^ c, 7);
TS_D uint32_t rot16_alu(uint32_t x) { return rotr(x, 16); }
#define TS_B3_G(a, b, c, d, mx, my) TS_B3_G_(a, b, c, d, mx, my, rot16_alu)
// the g-th G of a round (g = 0..7): the first TS_B3_ROT_FMA of the sequence 0, 2, 4, 6, 1, 3, 5, 7 rotate on the fma pipe
#define TS_B3_GSEL(g, a, b, c, d, mx, my)                                          \
    if ((((g) & 1) ? 4 + ((g) >> 1) : ((g) >> 1)) < TS_B3_ROT_FMA) {                \
        TS_B3_G_(a, b, c, d, mx, my, rot16_fma)                                     \
    } else {                                                                        \
        TS_B3_G_(a, b, c, d, mx, my, rot16_alu)                                     \
    }

// message schedule: word index used at (round, slot), i.e. the permutation applied r times
struct Sched {
    int s[7][16];
    constexpr Sched() : s{} {
        const int perm[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
        for (int i = 0; i < 16; i++) s[0][i] = i;
        for (int r = 1; r < 7; r++)
            for (int i = 0; i < 16; i++) s[r][i] = s[r - 1][perm[i]];
    }
};

// cv <- first 8 words of compress(cv, m, counter, block_len, flags)
TS_D void compress(uint32_t (&cv)[8], const uint32_t (&m)[16], uint32_t counter, uint32_t block_len,
                   uint32_t flags) {
    constexpr Sched S{};
    uint32_t s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
    uint32_t s8 = 0x6A09E667u, s9 = 0xBB67AE85u, s10 = 0x3C6EF372u, s11 = 0xA54FF53Au;
    uint32_t s12 = counter, s13 = 0u, s14 = block_len, s15 = flags;
    TS_UNROLL
    for (int r = 0; r < 7; r++) {
        TS_B3_GSEL(0, s0, s4, s8, s12, m[S.s[r][0]], m[S.s[r][1]])
        TS_B3_GSEL(1, s1, s5, s9, s13, m[S.s[r][2]], m[S.s[r][3]])
        TS_B3_GSEL(2, s2, s6, s10, s14, m[S.s[r][4]], m[S.s[r][5]])
        TS_B3_GSEL(3, s3, s7, s11, s15, m[S.s[r][6]], m[S.s[r][7]])
        TS_B3_GSEL(4, s0, s5, s10, s15, m[S.s[r][8]], m[S.s[r][9]])
        TS_B3_GSEL(5, s1, s6, s11, s12, m[S.s[r][10]], m[S.s[r][11]])
        TS_B3_GSEL(6, s2, s7, s8, s13, m[S.s[r][12]], m[S.s[r][13]])
        TS_B3_GSEL(7, s3, s4, s9, s14, m[S.s[r][14]], m[S.s[r][15]])
    }
    cv[0] = s0 ^ s8; cv[1] = s1 ^ s9; cv[2] = s2 ^ s10; cv[3] = s3 ^ s11;
    cv[4] = s4 ^ s12; cv[5] = s5 ^ s13; cv[6] = s6 ^ s14; cv[7] = s7 ^ s15;
}
TS_D void iv(uint32_t (&cv)[8]) {
    cv[0] = 0x6A09E667u; cv[1] = 0xBB67AE85u; cv[2] = 0x3C6EF372u; cv[3] = 0xA54FF53Au;
    cv[4] = 0x510E527Fu; cv[5] = 0x9B05688Cu; cv[6] = 0x1F83D9ABu; cv[7] = 0x5BE0CD19u;
}
// out = Blake3 parent-style compression of (l || r) with the given flags and IV key
TS_D void compress_pair(const uint32_t (&l)[8], const uint32_t (&r)[8], uint32_t flags, uint32_t (&out)[8]) {
    uint32_t m[16];
    TS_UNROLL
    for (int i = 0; i < 8; i++) { m[i] = l[i]; m[8 + i] = r[i]; }
    iv(out);
    compress(out, m, 0, 64, flags);
}

// ---- leaf hashing ---------------------------------------------------------------------------------
constexpr int MAX_SEG = 32;
struct Segments {
    const uint32_t *ptr[MAX_SEG];
    uint32_t width[MAX_SEG];
    uint32_t shift[MAX_SEG];  // row = leaf >> shift
    int n;
    uint32_t total_words;
};
constexpr int LEAVES_PER_CTA = 128;
constexpr int PIECE_WORDS = 64;          // words staged per leaf per step (4 Blake3 blocks)
constexpr int PITCH = PIECE_WORDS + 1;   // odd pitch: thread l reading word j hits bank (l + j) mod 32
constexpr int MAX_STACK = 6;             // up to 64 chunks = 64 KiB per leaf

// digests[leaf][8] = Blake3(concat_s row_s[leaf >> shift_s] as canonical LE u32).  monty: inputs are
// Montgomery form and are converted while staging.
__global__ void __launch_bounds__(LEAVES_PER_CTA) hash_leaves_kernel(Segments sg, size_t n_leaves, int monty,
                                                                     uint32_t *digests) {
    TS_DYN_SMEM(uint32_t, stage);
    const int tid = threadIdx.x;
    const size_t leaf0 = (size_t)blockIdx.x * LEAVES_PER_CTA;
    const size_t leaf = leaf0 + tid;
    const uint32_t total_blocks = sg.total_words == 0 ? 1u : (sg.total_words + 15u) / 16u;
    uint32_t cv[8];
    uint32_t stack[MAX_STACK][8];
    int sp = 0;
    b3::iv(cv);
    for (uint32_t c0 = 0; c0 < total_blocks * 16u; c0 += PIECE_WORDS) {
        // cooperative, coalesced staging of words [c0, c0+PIECE) of LEAVES_PER_CTA leaves
        __syncthreads();
        for (int it = tid; it < LEAVES_PER_CTA * PIECE_WORDS; it += LEAVES_PER_CTA) {
            const int l = it / PIECE_WORDS, j = it % PIECE_WORDS;
            const uint32_t c = c0 + j;
            uint32_t v = 0;
            if (c < sg.total_words && leaf0 + l < n_leaves) {
                uint32_t off = c;
                int s = 0;
                while (off >= sg.width[s]) { off -= sg.width[s]; s++; }
                v = sg.ptr[s][(size_t)((leaf0 + l) >> sg.shift[s]) * sg.width[s] + off];
                if (monty) v = bb::from_monty(v);
            }
            stage[l * PITCH + j] = v;
        }
        __syncthreads();
        if (leaf < n_leaves) {
            for (int bi = 0; bi < PIECE_WORDS / 16; bi++) {
                const uint32_t blk = c0 / 16 + bi;
                if (blk >= total_blocks) break;
                uint32_t m[16];
                TS_UNROLL
                for (int i = 0; i < 16; i++) m[i] = stage[tid * PITCH + bi * 16 + i];
                const uint32_t in_chunk = blk & 15u, chunk = blk >> 4;
                const bool last = (blk + 1 == total_blocks);
                uint32_t flags = (in_chunk == 0 ? CHUNK_START : 0u) | ((in_chunk == 15 || last) ? CHUNK_END : 0u);
                const uint32_t rem = sg.total_words - blk * 16u;  // >= 1 unless the message is empty
                const uint32_t block_len = sg.total_words == 0 ? 0u : (rem >= 16u ? 64u : rem * 4u);
                if (last && sp == 0) flags |= ROOT;
                compress(cv, m, chunk, block_len, flags);
                if (last) {
                    // fold the CV stack: parent(stack[top], cv) ..., ROOT on the final one
                    while (sp > 0) {
                        uint32_t out[8];
                        sp--;
                        compress_pair(stack[sp], cv, PARENT | (sp == 0 ? ROOT : 0u), out);
                        TS_UNROLL
                        for (int i = 0; i < 8; i++) cv[i] = out[i];
                    }
                } else if (in_chunk == 15) {
                    // chunk finished and more input follows: merge completed subtrees (trailing zeros rule)
                    uint32_t total_chunks = chunk + 1;
                    while ((total_chunks & 1u) == 0) {
                        uint32_t out[8];
                        sp--;
                        compress_pair(stack[sp], cv, PARENT, out);
                        TS_UNROLL
                        for (int i = 0; i < 8; i++) cv[i] = out[i];
                        total_chunks >>= 1;
                    }
                    TS_UNROLL
                    for (int i = 0; i < 8; i++) stack[sp][i] = cv[i];
                    sp++;
                    b3::iv(cv);
                }
            }
        }
    }
    if (leaf < n_leaves) {
        TS_UNROLL
        for (int i = 0; i < 8; i++) digests[leaf * 8 + i] = cv[i];
    }
}

// Per-block state machine shared by the leaf kernels: absorb message block `blk` (0-based) of a leaf of
// total_words u32 words.  After the last block, cv holds the Blake3 digest.
TS_D void absorb_block(uint32_t (&cv)[8], uint32_t (*stack)[8], int &sp, const uint32_t (&m)[16], uint32_t blk,
                       uint32_t total_blocks, uint32_t total_words) {
    const uint32_t in_chunk = blk & 15u, chunk = blk >> 4;
    const bool last = (blk + 1 == total_blocks);
    uint32_t flags = (in_chunk == 0 ? CHUNK_START : 0u) | ((in_chunk == 15 || last) ? CHUNK_END : 0u);
    const uint32_t rem = total_words - blk * 16u;
    const uint32_t block_len = total_words == 0 ? 0u : (rem >= 16u ? 64u : rem * 4u);
    if (last && sp == 0) flags |= ROOT;
    compress(cv, m, chunk, block_len, flags);
    if (last) {
        while (sp > 0) {
            uint32_t out[8];
            sp--;
            compress_pair(stack[sp], cv, PARENT | (sp == 0 ? ROOT : 0u), out);
            TS_UNROLL
            for (int i = 0; i < 8; i++) cv[i] = out[i];
        }
    } else if (in_chunk == 15) {
        uint32_t total_chunks = chunk + 1;
        while ((total_chunks & 1u) == 0) {
            uint32_t out[8];
            sp--;
            compress_pair(stack[sp], cv, PARENT, out);
            TS_UNROLL
            for (int i = 0; i < 8; i++) cv[i] = out[i];
            total_chunks >>= 1;
        }
        TS_UNROLL
        for (int i = 0; i < 8; i++) stack[sp][i] = cv[i];
        sp++;
        b3::iv(cv);
    }
}

// Hot path: ONE matrix, one row per leaf, width % 4 == 0 (the committed LDE).  A warp owns 32 consecutive
// leaves.  Per 64-byte block: 4 coalesced LDG.128 per lane (8 rows x 64 B per instruction) are issued one block
// AHEAD and stay in flight during the compression; the words are converted Montgomery -> canonical once,
// transposed through a warp-private, XOR-swizzled 2 KiB shared buffer (conflict-free STS.128 and LDS.128), and
// each lane compresses its own row in registers.  No block-wide barrier.
// Several matrices of EQUAL power-of-two width (the column blocks a rank receives from the all-to-all) are
// hashed as one row: word cw of the leaf lives in segment cw >> log_seg_w.
struct FastSegs {
    const uint32_t *ptr[MAX_SEG];
    uint32_t seg_w;      // width of every segment
    int log_seg_w;       // log2(seg_w) when n > 1 (unused for n == 1)
    int n;
};
constexpr int FAST_WARPS = 8;
// [blk_begin, blk_end) is the window of 64-byte blocks absorbed by this launch: the host->device pipeline hashes a
// row incrementally as its column chunks arrive, the chaining value parked in `digests` between launches (rows of
// at most one Blake3 chunk, so there is no subtree stack to carry).  A full hash is the window [0, total_blocks).
#ifndef TS_B3_MINBLOCKS
#define TS_B3_MINBLOCKS 0  // 0 = no register bound: 80 registers, 3 CTAs per SM; bounding it to 64 (4 CTAs) measured slower (10.05 ms)
#endif
#if TS_B3_MINBLOCKS > 0
#define TS_B3_BOUNDS __launch_bounds__(FAST_WARPS * 32, TS_B3_MINBLOCKS)
#else
#define TS_B3_BOUNDS __launch_bounds__(FAST_WARPS * 32)
#endif
__global__ void TS_B3_BOUNDS hash_rows_fast_kernel(FastSegs fs, uint32_t width, size_t n_leaves,
                                                                        int monty, uint32_t *digests, uint32_t blk_begin,
                                                                        uint32_t blk_end) {
    TS_DYN_SMEM(uint32_t, sm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *ws = sm + warp * 512;
    const size_t leaf0 = ((size_t)blockIdx.x * FAST_WARPS + warp) * 32;
    if (leaf0 >= n_leaves) return;  // warp-uniform
    const uint32_t total_blocks = (width + 15u) / 16u;
    const uint32_t sub = lane & 3, r0 = lane >> 2;
    uint4 pf[4];
    uint32_t cv[8], stack[MAX_STACK][8];
    int sp = 0;
    if (blk_begin == 0) {
        iv(cv);
    } else if (leaf0 + lane < n_leaves) {
        const uint4 *st = reinterpret_cast<const uint4 *>(digests + (leaf0 + lane) * 8);
        const uint4 a = st[0], b = st[1];
        cv[0] = a.x; cv[1] = a.y; cv[2] = a.z; cv[3] = a.w;
        cv[4] = b.x; cv[5] = b.y; cv[6] = b.z; cv[7] = b.w;
    } else {
        iv(cv);
    }
#define TS_FETCH(blk_)                                                                                        \
    TS_UNROLL                                                                                                 \
    for (int k = 0; k < 4; k++) {                                                                             \
        const uint32_t row = r0 + 8 * k, cw = 16u * (blk_) + 4u * sub;                                        \
        const uint32_t sgi = fs.n > 1 ? (cw >> fs.log_seg_w) : 0u;                                            \
        const uint32_t off = fs.n > 1 ? (cw & (fs.seg_w - 1u)) : cw;                                          \
        pf[k] = (leaf0 + row < n_leaves && cw < width)                                                        \
                    ? *reinterpret_cast<const uint4 *>(fs.ptr[sgi] + (leaf0 + row) * fs.seg_w + off)         \
                    : make_uint4(0, 0, 0, 0);                                                                 \
    }
    TS_FETCH(blk_begin)
    for (uint32_t blk = blk_begin; blk < blk_end; blk++) {
        TS_UNROLL
        for (int k = 0; k < 4; k++) {
            const uint32_t row = r0 + 8 * k;
            uint4 v = pf[k];
            if (monty) {
                v.x = bb::from_monty(v.x); v.y = bb::from_monty(v.y);
                v.z = bb::from_monty(v.z); v.w = bb::from_monty(v.w);
            }
            *reinterpret_cast<uint4 *>(ws + row * 16 + 4 * (sub ^ ((row >> 1) & 3u))) = v;
        }
        __syncwarp();
        if (blk + 1 < blk_end) { TS_FETCH(blk + 1) }
        uint32_t m[16];
        TS_UNROLL
        for (int j = 0; j < 4; j++) {
            const uint4 t = *reinterpret_cast<const uint4 *>(ws + lane * 16 + 4 * ((uint32_t)j ^ ((lane >> 1) & 3u)));
            m[4 * j + 0] = t.x; m[4 * j + 1] = t.y; m[4 * j + 2] = t.z; m[4 * j + 3] = t.w;
        }
        __syncwarp();
        absorb_block(cv, stack, sp, m, blk, total_blocks, width);
    }
#undef TS_FETCH
    if (leaf0 + lane < n_leaves) {
        uint4 *o = reinterpret_cast<uint4 *>(digests + (leaf0 + lane) * 8);
        o[0] = make_uint4(cv[0], cv[1], cv[2], cv[3]);
        o[1] = make_uint4(cv[4], cv[5], cv[6], cv[7]);
    }
}

// Fast path for narrow single-segment leaves (<= 16 words, e.g. the FRI layers: 2 ext elements = 8 words):
// one thread reads its own contiguous leaf, one compression.
__global__ void __launch_bounds__(256) hash_leaves_small_kernel(const uint32_t *rows, uint32_t width, size_t n_leaves,
                                                                int monty, uint32_t *digests) {
    const size_t leaf = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n_leaves) return;
    uint32_t m[16];
    TS_UNROLL
    for (int i = 0; i < 16; i++) {
        uint32_t v = 0;
        if ((uint32_t)i < width) {
            v = rows[leaf * width + i];
            if (monty) v = bb::from_monty(v);
        }
        m[i] = v;
    }
    uint32_t cv[8];
    iv(cv);
    compress(cv, m, 0, width * 4u, CHUNK_START | CHUNK_END | ROOT);
    TS_UNROLL
    for (int i = 0; i < 8; i++) digests[leaf * 8 + i] = cv[i];
}

// ---- tree ---------------------------------------------------------------------------------------
// One CTA consumes 2*T child digests and produces up to `levels` layers above them (T, T/2, ... nodes),
// every layer written to its own global array (openings need all of them).  Upper layers are reduced out
// of shared memory.
constexpr int TREE_T = 128;
constexpr int TREE_MAX_LEVELS = 8;  // 2*128 = 2^8 children per CTA
struct TreeLevels {
    uint32_t *out[TREE_MAX_LEVELS];
    int levels;
};
constexpr int NODE_PITCH = 9;
__global__ void __launch_bounds__(TREE_T) tree_reduce_kernel(const uint32_t *children, size_t n_children,
                                                             TreeLevels lv) {
    TS_DYN_SMEM(uint32_t, nodes);  // TREE_T * NODE_PITCH words
    const int tid = threadIdx.x;
    const size_t parent0 = (size_t)blockIdx.x * TREE_T;
    size_t n_par = n_children / 2;
    uint32_t l[8], r[8], o[8];
    {
        const size_t pidx = parent0 + tid;
        if (pidx < n_par) {
            TS_UNROLL
            for (int i = 0; i < 8; i++) { l[i] = children[pidx * 16 + i]; r[i] = children[pidx * 16 + 8 + i]; }
            compress_pair(l, r, CHUNK_START | CHUNK_END | ROOT, o);
            TS_UNROLL
            for (int i = 0; i < 8; i++) { lv.out[0][pidx * 8 + i] = o[i]; nodes[tid * NODE_PITCH + i] = o[i]; }
        }
    }
    int width = TREE_T;
    for (int level = 1; level < lv.levels; level++) {
        __syncthreads();
        width >>= 1;
        n_par >>= 1;
        const size_t pidx = (parent0 >> level) + tid;
        const bool act = tid < width && pidx < n_par;
        if (act) {
            TS_UNROLL
            for (int i = 0; i < 8; i++) {
                l[i] = nodes[(2 * tid) * NODE_PITCH + i];
                r[i] = nodes[(2 * tid + 1) * NODE_PITCH + i];
            }
            compress_pair(l, r, CHUNK_START | CHUNK_END | ROOT, o);
        }
        __syncthreads();
        if (act) {
            TS_UNROLL
            for (int i = 0; i < 8; i++) { lv.out[level][pidx * 8 + i] = o[i]; nodes[tid * NODE_PITCH + i] = o[i]; }
        }
    }
}

// Big layers: one thread reduces 8 consecutive children to 4 + 2 + 1 parents held in registers -- three levels per
// launch, no barrier, every lane busy (tree_reduce_kernel halves its active threads per level and waits at a barrier
// between levels: profiles/r01, 6 of 10 stall cycles).  n_children must be a multiple of 8.
__global__ void __launch_bounds__(128) tree_reduce3_kernel(const uint32_t *__restrict__ children, size_t n_children,
                                                           uint32_t *__restrict__ out0, uint32_t *__restrict__ out1,
                                                           uint32_t *__restrict__ out2) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_children / 8) return;
    const uint4 *in = reinterpret_cast<const uint4 *>(children) + t * 16;  // 8 digests = 16 uint4
    uint32_t p[4][8];
    TS_UNROLL
    for (int k = 0; k < 4; k++) {
        const uint4 a0 = in[4 * k], a1 = in[4 * k + 1], b0 = in[4 * k + 2], b1 = in[4 * k + 3];
        const uint32_t l[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const uint32_t r[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        compress_pair(l, r, CHUNK_START | CHUNK_END | ROOT, p[k]);
        uint4 *o = reinterpret_cast<uint4 *>(out0) + (4 * t + k) * 2;
        o[0] = make_uint4(p[k][0], p[k][1], p[k][2], p[k][3]);
        o[1] = make_uint4(p[k][4], p[k][5], p[k][6], p[k][7]);
    }
    uint32_t q[2][8];
    TS_UNROLL
    for (int j = 0; j < 2; j++) {
        compress_pair(p[2 * j], p[2 * j + 1], CHUNK_START | CHUNK_END | ROOT, q[j]);
        uint4 *o = reinterpret_cast<uint4 *>(out1) + (2 * t + j) * 2;
        o[0] = make_uint4(q[j][0], q[j][1], q[j][2], q[j][3]);
        o[1] = make_uint4(q[j][4], q[j][5], q[j][6], q[j][7]);
    }
    uint32_t r8[8];
    compress_pair(q[0], q[1], CHUNK_START | CHUNK_END | ROOT, r8);
    uint4 *o = reinterpret_cast<uint4 *>(out2) + t * 2;
    o[0] = make_uint4(r8[0], r8[1], r8[2], r8[3]);
    o[1] = make_uint4(r8[4], r8[5], r8[6], r8[7]);
}

// padding_matrix (basic/src/tcs/mod.rs:341-383) materialised: out[leaf] = concat_s row_s[leaf >> shift_s], the row of field words
// a TapTree leaf commits to (matrices tallest first; a shorter matrix repeats each row over 2^shift consecutive leaves).
__global__ void __launch_bounds__(256) padded_rows_kernel(Segments sg, size_t n_leaves, uint32_t *__restrict__ out) {
    const size_t total = n_leaves * sg.total_words;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t leaf = i / sg.total_words;
        uint32_t off = (uint32_t)(i % sg.total_words);
        int s = 0;
        while (off >= sg.width[s]) { off -= sg.width[s]; s++; }
        out[i] = sg.ptr[s][(leaf >> sg.shift[s]) * sg.width[s] + off];
    }
}

// P3 injection layer: out[i] = H( H(prev[2i] || prev[2i+1]) || rows_digest[i] )
__global__ void compress_inject_kernel(const uint32_t *children, const uint32_t *rows_digest, size_t n_par,
                                       uint32_t *out) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_par) return;
    uint32_t l[8], r[8], o[8], o2[8];
    TS_UNROLL
    for (int i = 0; i < 8; i++) { l[i] = children[p * 16 + i]; r[i] = children[p * 16 + 8 + i]; }
    compress_pair(l, r, CHUNK_START | CHUNK_END | ROOT, o);
    TS_UNROLL
    for (int i = 0; i < 8; i++) r[i] = rows_digest[p * 8 + i];
    compress_pair(o, r, CHUNK_START | CHUNK_END | ROOT, o2);
    TS_UNROLL
    for (int i = 0; i < 8; i++) out[p * 8 + i] = o2[i];
}

}  // namespace b3
