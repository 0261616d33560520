// fold.cuh -- FRI commit-phase even/odd fold and the reduced-opening column combination (sm_100a).
//
// fold: TwoAdicFriGenericConfig::fold_matrix (fri/src/two_adic_pcs.rs:116-147) == fold_even_odd
// (fri/src/fold_even_odd.rs:20-52):
//     out[i] = (1/2 + pw_i) * lo + (1/2 - pw_i) * hi,   pw_i = (beta/2) * g_inv^bitrev(i),
//     g_inv = two_adic_generator(log h + 1)^-1, (lo, hi) = row i of the h x 2 view.
// Computed as out[i] = (lo+hi)/2 + (beta/2) * [ g_inv^bitrev(i) * (lo-hi) ]  (same field element, exact).
// The reference materialises and bit-reverse-permutes the `powers` vector every call; here
// g_inv^bitrev(i) = T_lo[i mod 256] * T_hi(i div 256): T_lo is a 256-entry size-independent table
// (w_512^-bitrev8(t)), T_hi one scalar per CTA chunk, so the kernel streams 32 B in / 16 B out per element
// with no other global traffic.
#pragma once
#include "field.cuh"
#include "hash.cuh"

namespace fold {

constexpr int FOLD_T = 256;

TS_D uint32_t brev_bits(uint32_t x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

struct InvRootPows {
    uint32_t v[28];  // v[k] = (w_{2h}^-1)^(2^k), Montgomery
};
TS_D uint32_t pow_from_table(const InvRootPows &rp, uint32_t e) {
    uint32_t acc = bb::MONTY_ONE;
    for (int k = 0; e; k++, e >>= 1)
        if (e & 1) acc = bb::mmul(acc, rp.v[k]);
    return acc;
}

// delta table: D.v[t] = g_inv^(bitrev(c+1) - bitrev(c)) for a chunk index c with t trailing one bits, so a
// CTA walking consecutive 256-element chunks updates its per-chunk scalar with ONE multiply.
struct ChunkDeltas {
    uint32_t v[28];
};

// EF fold.  in: 2h EF (8 u32 per output row), out: h EF.  addend (optional): the next FRI input of length h
// (fri/src/prover.rs:124-126), added after folding.  tlo[t] = (w_512^-1)^bitrev8(t) (Montgomery), used when
// log_h >= 8; smaller layers take the direct power.
// Row-sharded use (multi-GPU): the launch folds rows [first, first + h_local) of a layer of 2^log_h rows;
// in/out/addend point at the shard.  first and h_local are multiples of 256 when log_h >= 8.
__global__ void __launch_bounds__(FOLD_T) fold_ext_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out,
                                                          const uint4 *__restrict__ addend, int log_h, size_t first,
                                                          size_t h_local, ef::E4 half_beta, InvRootPows rp,
                                                          ChunkDeltas dl, const uint32_t *__restrict__ tlo,
                                                          const uint32_t *__restrict__ half_beta_dev) {
    // half_beta_dev != nullptr: beta/2 (4 Montgomery words) was left in device memory by the sponge step of this round
    // (fri_tail.cuh: sponge_step_kernel), so no host round trip separates the layer's commitment from its fold
    if (half_beta_dev) {
        TS_UNROLL
        for (int i = 0; i < 4; i++) half_beta.c[i] = half_beta_dev[i];
    }
    const ef::E4Const hb = ef::prepare(half_beta);
    const size_t h = h_local;
    size_t c0 = 0, c1 = 1;
    uint32_t t_lo = bb::MONTY_ONE, t_hi = bb::MONTY_ONE;
    const bool chunked = log_h >= 8;
    const size_t chunk_first = first >> 8;
    if (chunked) {
        const size_t chunks = h >> 8, cpb = (chunks + gridDim.x - 1) / gridDim.x;
        c0 = (size_t)blockIdx.x * cpb;
        c1 = c0 + cpb < chunks ? c0 + cpb : chunks;
        if (c0 >= c1) return;
        t_lo = tlo[threadIdx.x];
        t_hi = pow_from_table(rp, brev_bits((uint32_t)(c0 + chunk_first), log_h - 8));
    } else if (blockIdx.x > 0) {
        return;
    }
    for (size_t c = c0; c < c1; c++) {
        const size_t i = (c << 8) + threadIdx.x;
        if (i < h) {
            const uint32_t s = chunked ? bb::mmul(t_lo, t_hi)
                                       : pow_from_table(rp, brev_bits((uint32_t)(i + first), log_h));
            const uint4 a = in[2 * i], b = in[2 * i + 1];
            ef::E4 lo{{a.x, a.y, a.z, a.w}}, hi{{b.x, b.y, b.z, b.w}};
            const ef::E4 sum = ef::half(ef::add(lo, hi));
            const ef::E4 dif = ef::scale(ef::sub(lo, hi), s);
            ef::E4 r = ef::add(sum, ef::mul(dif, hb));
            if (addend) {
                const uint4 x = addend[i];
                r = ef::add(r, ef::E4{{x.x, x.y, x.z, x.w}});
            }
            out[i] = make_uint4(r.c[0], r.c[1], r.c[2], r.c[3]);
        }
        // trailing ones of the global chunk index -> which delta moves bitrev(c) to bitrev(c+1)
        int t = 0;
        for (size_t cc = c + chunk_first; cc & 1; cc >>= 1) t++;
        t_hi = bb::mmul(t_hi, dl.v[t < 27 ? t : 27]);
    }
}

// Fold + the NEXT round's leaf hashes in one pass (fri/src/prover.rs:110-126): the next round commits to the folded vector viewed
// as rows of two extension elements (prover.rs:112), i.e. leaf j = Blake3(out[2j] || out[2j+1]) as canonical little-endian words.
// A thread folds the two neighbouring rows 2j and 2j+1 and hashes what it just produced, so the layer is not read a second time
// by a leaf-hash kernel (and one dependent launch per round disappears).  Twiddles: g_inv^bitrev(2j) = T9[j mod 256] * T_hi(j div
// 256) with T9[t] = (w_1024^-1)^bitrev8(t) raised per thread once, g_inv^bitrev(2j+1) = g_inv^bitrev(2j) * g_inv^(h/2).
// Requires log_h >= 9; dl holds the chunk deltas of (log_h - 9)-bit chunk indices.  digests: h_local/2 x 8 words.
// Row-sharded use as fold_ext_kernel: rows [first, first + h_local) of the layer, both multiples of 512; in/out/addend/digests
// point at the shard.
__global__ void __launch_bounds__(FOLD_T) fold_hash_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out,
                                                           const uint4 *__restrict__ addend, int log_h, size_t first,
                                                           size_t h_local, ef::E4 half_beta, InvRootPows rp, ChunkDeltas dl,
                                                           const uint32_t *__restrict__ half_beta_dev,
                                                           uint32_t *__restrict__ digests) {
    if (half_beta_dev) {
        TS_UNROLL
        for (int i = 0; i < 4; i++) half_beta.c[i] = half_beta_dev[i];
    }
    const ef::E4Const hb = ef::prepare(half_beta);
    const size_t chunks = h_local >> 9, cpb = (chunks + gridDim.x - 1) / gridDim.x, chunk_first = first >> 9;
    const size_t c0 = (size_t)blockIdx.x * cpb, c1 = c0 + cpb < chunks ? c0 + cpb : chunks;
    if (c0 >= c1) return;
    const uint32_t t_lo = pow_from_table(rp, brev_bits(threadIdx.x, 8) << (log_h - 9));
    const uint32_t w4_inv = rp.v[log_h - 1];  // g_inv^(h/2)
    uint32_t t_hi = pow_from_table(rp, brev_bits((uint32_t)(c0 + chunk_first), log_h - 9));
    for (size_t c = c0; c < c1; c++) {
        const size_t j = (c << 8) + threadIdx.x;
        const uint32_t s0 = bb::mmul(t_lo, t_hi), s1 = bb::mmul(s0, w4_inv);
        uint32_t m[16];
        TS_UNROLL
        for (int k = 0; k < 2; k++) {
            const uint4 a = in[4 * j + 2 * k], b = in[4 * j + 2 * k + 1];
            ef::E4 lo{{a.x, a.y, a.z, a.w}}, hi{{b.x, b.y, b.z, b.w}};
            const ef::E4 sum = ef::half(ef::add(lo, hi));
            const ef::E4 dif = ef::scale(ef::sub(lo, hi), k ? s1 : s0);
            ef::E4 r = ef::add(sum, ef::mul(dif, hb));
            if (addend) {
                const uint4 x = addend[2 * j + k];
                r = ef::add(r, ef::E4{{x.x, x.y, x.z, x.w}});
            }
            out[2 * j + k] = make_uint4(r.c[0], r.c[1], r.c[2], r.c[3]);
            TS_UNROLL
            for (int i = 0; i < 4; i++) m[4 * k + i] = bb::from_monty(r.c[i]);
        }
        TS_UNROLL
        for (int i = 8; i < 16; i++) m[i] = 0;
        uint32_t cv[8];
        b3::iv(cv);
        b3::compress(cv, m, 0, 32u, b3::CHUNK_START | b3::CHUNK_END | b3::ROOT);
        uint4 *d = reinterpret_cast<uint4 *>(digests) + 2 * j;
        d[0] = make_uint4(cv[0], cv[1], cv[2], cv[3]);
        d[1] = make_uint4(cv[4], cv[5], cv[6], cv[7]);
        int t = 0;
        for (size_t cc = c + chunk_first; cc & 1; cc >>= 1) t++;
        t_hi = bb::mmul(t_hi, dl.v[t < 27 ? t : 27]);
    }
}

// Base-field fold (the reference's own fold test and fri/tests/fri.rs run FRI over BabyBear itself).
__global__ void __launch_bounds__(FOLD_T) fold_base_kernel(const uint2 *__restrict__ in, uint32_t *__restrict__ out,
                                                           int log_h, uint32_t half_beta, InvRootPows rp) {
    const size_t h = (size_t)1 << log_h;
    for (size_t i = (size_t)blockIdx.x * FOLD_T + threadIdx.x; i < h; i += (size_t)gridDim.x * FOLD_T) {
        const uint32_t s = pow_from_table(rp, brev_bits((uint32_t)i, log_h));
        const uint2 v = in[i];
        const uint32_t sum = bb::half(bb::add(v.x, v.y));
        const uint32_t dif = bb::mmul(bb::sub(v.x, v.y), s);
        out[i] = bb::add(sum, bb::mmul(dif, half_beta));
    }
}

// dot_ext_powers: out[r] = sum_c alpha^c * m[r][c]  (fri/src/two_adic_pcs.rs:375, `mat.dot_ext_powers(alpha)`).
// alpha powers (Montgomery EF) come precomputed in `apow` (width entries).  One thread per row segment:
// a CTA stages rows coalesced through shared memory, then thread t reduces row t.
constexpr int DOT_ROWS = 64;
constexpr int DOT_COLS = 64;
__global__ void __launch_bounds__(256) dot_ext_powers_kernel(const uint32_t *__restrict__ m, size_t rows, uint32_t width,
                                                             const uint4 *__restrict__ apow, uint4 *__restrict__ out,
                                                             int accumulate) {
    TS_DYN_SMEM(uint32_t, sm);  // DOT_ROWS x (DOT_COLS+1) data, then 4 partial sets
    uint32_t *tile = sm;
    const int tid = threadIdx.x;
    const size_t row0 = (size_t)blockIdx.x * DOT_ROWS;
    const int r = tid & (DOT_ROWS - 1), part = tid / DOT_ROWS;  // 4 column parts per row
    uint32_t acc[4] = {0, 0, 0, 0};
    for (uint32_t c0 = 0; c0 < width; c0 += DOT_COLS) {
        __syncthreads();
        for (int it = tid; it < DOT_ROWS * DOT_COLS; it += 256) {
            const int rr = it / DOT_COLS, cc = it % DOT_COLS;
            uint32_t v = 0;
            if (row0 + rr < rows && c0 + cc < width) v = m[(row0 + rr) * width + c0 + cc];
            tile[rr * (DOT_COLS + 1) + cc] = v;
        }
        __syncthreads();
        for (int cc = part; cc < DOT_COLS && c0 + cc < width; cc += 4) {
            const uint32_t v = tile[r * (DOT_COLS + 1) + cc];
            const uint4 a = __ldg(apow + c0 + cc);
            acc[0] = bb::add(acc[0], bb::mmul(v, a.x));
            acc[1] = bb::add(acc[1], bb::mmul(v, a.y));
            acc[2] = bb::add(acc[2], bb::mmul(v, a.z));
            acc[3] = bb::add(acc[3], bb::mmul(v, a.w));
        }
    }
    __syncthreads();
    uint32_t *red = sm;  // reuse: [part][row][4]
    for (int k = 0; k < 4; k++) red[(part * DOT_ROWS + r) * 4 + k] = acc[k];
    __syncthreads();
    if (part == 0 && row0 + r < rows) {
        uint32_t o[4];
        for (int k = 0; k < 4; k++) {
            uint32_t s = red[r * 4 + k];
            for (int pp = 1; pp < 4; pp++) s = bb::add(s, red[(pp * DOT_ROWS + r) * 4 + k]);
            o[k] = s;
        }
        if (accumulate) {
            const uint4 prev = out[row0 + r];
            o[0] = bb::add(o[0], prev.x); o[1] = bb::add(o[1], prev.y);
            o[2] = bb::add(o[2], prev.z); o[3] = bb::add(o[3], prev.w);
        }
        out[row0 + r] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// t < 2 p 2^32  ->  t mod (p 2^32) congruent, < p 2^32
TS_D uint64_t fold64(uint64_t t) {
    const uint32_t hi = (uint32_t)(t >> 32);
    return ((uint64_t)bb::umin32(hi, hi - bb::P) << 32) | (uint32_t)t;
}

// Rows of 4 or 8 words (quotient chunks, fri/src/two_adic_pcs.rs:375 on a BabyBear^4 column flattened to base-field columns):
// one thread per row, the row in one or two 16-byte loads, grid-stride.  The warp-transposing kernel below moves 16 of every 64
// staged bytes for such rows (measured 1.4 TB/s on 2^23 x 4).
__global__ void __launch_bounds__(256) dot_rows_small_kernel(const uint4 *__restrict__ m, size_t rows, uint32_t quads,
                                                             const uint4 *__restrict__ apow, uint4 *__restrict__ out,
                                                             int accumulate) {
    uint4 a[8];
    TS_UNROLL
    for (int i = 0; i < 8; i++) a[i] = (uint32_t)i < 4 * quads ? apow[i] : make_uint4(0, 0, 0, 0);
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
        uint64_t acc[4] = {0, 0, 0, 0};
        TS_UNROLL
        for (int q = 0; q < 2; q++) {
            if ((uint32_t)q < quads) {
                const uint4 v = m[r * quads + q];
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                TS_UNROLL
                for (int i = 0; i < 4; i += 2) {
                    const uint4 a0 = a[4 * q + i], a1 = a[4 * q + i + 1];
                    acc[0] = fold64(bb::madw(w[i + 1], a1.x, bb::madw(w[i], a0.x, acc[0])));
                    acc[1] = fold64(bb::madw(w[i + 1], a1.y, bb::madw(w[i], a0.y, acc[1])));
                    acc[2] = fold64(bb::madw(w[i + 1], a1.z, bb::madw(w[i], a0.z, acc[2])));
                    acc[3] = fold64(bb::madw(w[i + 1], a1.w, bb::madw(w[i], a0.w, acc[3])));
                }
            }
        }
        uint32_t o[4];
        TS_UNROLL
        for (int k = 0; k < 4; k++) o[k] = bb::redc(acc[k]);
        if (accumulate) {
            const uint4 prev = out[r];
            o[0] = bb::add(o[0], prev.x); o[1] = bb::add(o[1], prev.y);
            o[2] = bb::add(o[2], prev.z); o[3] = bb::add(o[3], prev.w);
        }
        out[r] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// Hot path of dot_ext_powers for matrices with width % 4 == 0 (the committed LDE): a warp owns 32 rows; per block
// of 16 columns the lanes issue 4 coalesced LDG.128 each (one block ahead), transpose through a warp-private
// XOR-swizzled 2 KiB buffer, and every lane accumulates its own row in four 64-bit sums (IMAD.WIDE), folded
// below p 2^32 by a min on the high word every two products and Montgomery-reduced once per row.
// apow must be padded with zeros to a multiple of 16 entries.
constexpr int DOT_FAST_WARPS = 8;
constexpr size_t DOT_SMEM_BYTES = (size_t)DOT_FAST_WARPS * 1024 * 4;  // 4 KiB transpose buffer per warp
#ifndef TS_DOT_MINBLOCKS
#define TS_DOT_MINBLOCKS 3
#endif
// Several column blocks of EQUAL power-of-two width (what a rank holds after the all-to-all) are read as one row:
// word cw of the row lives in block cw >> log_seg_w (same convention as b3::FastSegs).
constexpr int DOT_MAX_SEG = 32;
struct DotSegs {
    const uint32_t *ptr[DOT_MAX_SEG];
    uint32_t seg_w;
    int log_seg_w;
    int n;
};
__global__ void __launch_bounds__(DOT_FAST_WARPS * 32, TS_DOT_MINBLOCKS) dot_rows_fast_kernel(DotSegs sg, size_t rows, uint32_t width,
                                                                          const uint4 *__restrict__ apow,
                                                                          uint4 *__restrict__ out, int accumulate) {
    TS_DYN_SMEM(uint32_t, sm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *ws = sm + warp * 1024;
    const size_t row0 = ((size_t)blockIdx.x * DOT_FAST_WARPS + warp) * 32;
    if (row0 >= rows) return;  // warp-uniform
    // one iteration = 32 columns: 8 lanes cover 128 contiguous bytes of a row (4 rows per LDG.128), 8 loads per lane
    // in flight one iteration ahead -- 128-byte DRAM bursts and twice the bytes in flight of the 64-byte version
    const uint32_t total_iters = (width + 31u) / 32u;
    const uint32_t sub = lane & 7, r0 = lane >> 3;
    uint4 pf[8];
    // 64-bit running sums, kept below p * 2^32 (the Montgomery reduction's input range) by one conditional subtraction
    // of p * 2^32 -- a min on the high word -- after every two products (2 p^2 + p 2^32 < 2 p 2^32 < 2^64)
    uint64_t acc[4] = {0, 0, 0, 0};
#define TS_DOT_FETCH(it_)                                                                            \
    TS_UNROLL                                                                                        \
    for (int k = 0; k < 8; k++) {                                                                    \
        const uint32_t row = r0 + 4 * k, cw = 32u * (it_) + 4u * sub;                                \
        const uint32_t sgi = sg.n > 1 ? (cw >> sg.log_seg_w) : 0u;                                   \
        const uint32_t off = sg.n > 1 ? (cw & (sg.seg_w - 1u)) : cw;                                 \
        pf[k] = (row0 + row < rows && cw < width)                                                    \
                    ? *reinterpret_cast<const uint4 *>(sg.ptr[sgi] + (row0 + row) * sg.seg_w + off)  \
                    : make_uint4(0, 0, 0, 0);                                                        \
    }
    TS_DOT_FETCH(0u)
    for (uint32_t it = 0; it < total_iters; it++) {
        TS_UNROLL
        for (int k = 0; k < 8; k++) {
            const uint32_t row = r0 + 4 * k;
            *reinterpret_cast<uint4 *>(ws + row * 32 + 4 * (sub ^ (row & 7u))) = pf[k];
        }
        __syncwarp();
        if (it + 1 < total_iters) { TS_DOT_FETCH(it + 1) }
        TS_UNROLL
        for (int half = 0; half < 2; half++) {
            const uint32_t c0 = 32u * it + 16u * half;
            if (c0 >= width) break;  // warp-uniform: apow is padded to 16, not 32
            uint32_t v[16];
            TS_UNROLL
            for (int j = 0; j < 4; j++) {
                const uint4 t = *reinterpret_cast<const uint4 *>(ws + lane * 32 + 4 * ((uint32_t)(4 * half + j) ^ (lane & 7u)));
                v[4 * j + 0] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
            }
            TS_UNROLL
            for (int i = 0; i < 16; i += 2) {
                const uint4 a0 = __ldg(apow + c0 + i), a1 = __ldg(apow + c0 + i + 1);
                acc[0] = fold64(acc[0] + (uint64_t)v[i] * a0.x + (uint64_t)v[i + 1] * a1.x);
                acc[1] = fold64(acc[1] + (uint64_t)v[i] * a0.y + (uint64_t)v[i + 1] * a1.y);
                acc[2] = fold64(acc[2] + (uint64_t)v[i] * a0.z + (uint64_t)v[i + 1] * a1.z);
                acc[3] = fold64(acc[3] + (uint64_t)v[i] * a0.w + (uint64_t)v[i + 1] * a1.w);
            }
        }
        __syncwarp();
    }
#undef TS_DOT_FETCH
    if (row0 + lane < rows) {
        uint32_t r[4] = {bb::redc(acc[0]), bb::redc(acc[1]), bb::redc(acc[2]), bb::redc(acc[3])};
        if (accumulate) {
            const uint4 prev = out[row0 + lane];
            r[0] = bb::add(r[0], prev.x); r[1] = bb::add(r[1], prev.y);
            r[2] = bb::add(r[2], prev.z); r[3] = bb::add(r[3], prev.w);
        }
        out[row0 + lane] = make_uint4(r[0], r[1], r[2], r[3]);
    }
}

}  // namespace fold
