// ntt_v4.cuh -- fourth-generation digit pass: ONE kernel template for all four launches of a two-digit coset LDE.
//
//     P1  inverse, top digit      caller's row-major matrix -> blocked scratch        (strided rows, POST twiddles)
//     P2  inverse, low digit      scratch in place                                     (contiguous rows)
//     P3  forward, top digit      scratch -> blocked N x w intermediate, once per coset (MID: bit-reversed row load, coset
//                                                                                      powers on the way in, POST twiddles)
//     P4  forward, low digit      intermediate -> caller's row-major LDE, committed order (contiguous rows)
//
// Round 1 fused the work of P2 and P3 into `lde_mid` (ntt_pm.cuh): the coefficient tile stayed on chip across the 2^b
// cosets, which costs 128 KiB of shared memory per CTA and therefore ONE resident CTA per SM (issue-active 44 %, 42 % slower
// per butterfly than the digit passes, profiles/r01/v15_ncu_full.md).  Here P3 re-reads the coefficient tile once per coset
// -- the 2^b CTAs of a tile are adjacent in launch order, so three of the four reads are L2 hits -- and every launch runs at
// three CTAs per SM.
//
// Same tile, swizzle and inner rounds as ntt_pm.cuh (2^D positions x 2^(14-D) columns, position-major 16-byte units,
// unit(h, p) = hx(h) ^ sigma(p); sigma is exactly the 128-byte TMA swizzle of a plane with 16-byte rows).  What changed
// is the code around the first and last round, where the SASS of the round-1 pass spent 20 and 19.5 instructions per
// element against 12 and 9 of arithmetic (profiles/r02/sass/README.md):
//   * per-thread column base pointers and compile-time trip counts: a global access is  base + c * step  (no per-access
//     column predicate, no 64-bit multiply per element);
//   * inter-digit (POST) twiddles come from a table laid out [tile][position]: two LDG.128 per radix-4 item instead of
//     four BREV + IMAD + scattered 8-byte gathers out of the 2^m-entry power table; for P3 the per-(coset, tile) scalar
//     (s w_N^j)^Kc / n is folded into that table, so the way in costs one multiplication per element instead of two;
//   * coset powers (PRE) are laid out [coset][group][register]: LDG.128 pairs.
#pragma once
#include "ntt_pm.cuh"

namespace ntt4 {

using ntt::brev_bits;
using ntt::brev_c;
using nttf::FastTables;
using nttp::dft_v4;
using nttp::dnq;
using nttp::Geo;
using nttp::hx;
using nttp::io_blk;
using nttp::ld4;
using nttp::rtw;
using nttp::sigma;
using nttp::st4;
using nttp::V4;
using nttp::vmul;

struct PassParams {
    const uint32_t *src;
    uint32_t *dst;
    uint32_t src_pitch, dst_pitch, ncols;  // row strides in words (row-major side), columns of the window
    size_t src_slice = 0, dst_slice = 0;   // != 0: blocked layout [col / 8][row][8] with this many words per column group
    int lo_bits = 0, hi_bits = 0;          // pass mode: row = (hi << (lo_bits + D)) | (q << lo_bits) | lo
    uint32_t n_col_slices = 1;
    int cs_shift = -1;                     // log2(n_col_slices) when that is a power of two
    FastTables t;
    const uint2 *post = nullptr;  // pass mode: [2^lo_bits][2^D]; MID: [2^b][2^klo_bits][2^D]; nullptr: no inter-digit twiddle
    const uint2 *pre = nullptr;   // MID: [2^b][256][2^(D-8)] coset powers of the first round's operands
    int klo_bits = 0, b = 0, m = 0;  // MID: tile Kc of coset j reads rows [brev(Kc) << D, +2^D), writes rows (brev(j) << m) + Kc + (p << klo_bits)
};

// byte offset of (row 0, col) in either layout; the row term is added separately:  row * row_pitch_bytes
TS_D size_t col_off_bytes(uint32_t col, size_t slice) {
    return slice ? ((size_t)(col >> 3) * slice + (col & 7u)) * 4 : (size_t)col * 4;
}
TS_D size_t row_pitch_bytes(uint32_t pitch, size_t slice) { return slice ? 32 : (size_t)pitch * 4; }

TS_D uint4 ldg16(const char *p) { return nttf::ldg_tile(reinterpret_cast<const uint32_t *>(p)); }

// ---- first round straight from global memory ----------------------------------------------------------------------------
// item = (g, h): positions g + 256 c (c < R) of quad h.  BREV: position q holds source row brev_D(q) of a contiguous tile,
// i.e. the R operands of an item are the R consecutive rows  brev_8(g) * R + brev(c).
// CSRC: the tile is 2^D consecutive rows of a blocked matrix (32 bytes per row): every operand address is the item's base
// plus a compile-time offset.  Threads whose columns lie beyond the matrix skip the round: the tile words they would have
// written are only ever read back by items of the same (invalid) quad, whose results are never stored.
template <int D, bool INV, bool BREV, bool PRE, bool CSRC, int NT, int NQv = dnq(D)>
TS_D void r1_ld(uint4 *tile, const char *src_h, bool valid, size_t row_base, int log_stride, size_t pitch_b, const FastTables &t,
                const uint2 *pre, int tid) {
    using G = Geo<D, NQv>;
    constexpr int LOGR = G::LOGR1, R = 1 << LOGR;
    constexpr int ITEMS = 256 * G::NQ;
    static_assert(ITEMS % NT == 0 && NT % G::NQ == 0, "r1_ld: thread count");
    static_assert(!BREV || CSRC, "bit-reversed loads read a contiguous blocked tile");
    if (!valid) return;
    const uint32_t h = tid & (G::NQ - 1);
    const size_t step = CSRC ? 0 : (((size_t)256 << log_stride) * pitch_b);
    TS_NOUNROLL
    for (int k = 0; k < ITEMS / NT; k++) {
        const uint32_t g = (uint32_t)(tid + k * NT) / G::NQ;
        V4 x[R];
        if (CSRC) {
            const char *ptr = src_h + (row_base + (BREV ? (size_t)brev_bits(g, 8) * R : (size_t)g)) * 32;
            TS_UNROLL
            for (int c = 0; c < R; c++) {
                const uint4 v = ldg16(ptr + (BREV ? brev_c(c, LOGR) * 32 : c * 256 * 32));
                x[c] = V4{{v.x, v.y, v.z, v.w}};
            }
        } else {
            const char *ptr = src_h + (row_base + ((size_t)g << log_stride)) * pitch_b;
            TS_UNROLL
            for (int c = 0; c < R; c++) {
                const uint4 v = ldg16(ptr + (size_t)c * step);
                x[c] = V4{{v.x, v.y, v.z, v.w}};
            }
        }
        if (PRE) {
            const uint4 *pq = reinterpret_cast<const uint4 *>(pre + (size_t)g * R);
            TS_UNROLL
            for (int c = 0; c < R; c += 2) {
                const uint4 w = __ldg(pq + c / 2);
                x[c] = vmul(x[c], make_uint2(w.x, w.y));
                x[c + 1] = vmul(x[c + 1], make_uint2(w.z, w.w));
            }
        }
        dft_v4<LOGR, INV, true>(x);
        TS_UNROLL
        for (int i = 1; i < R; i++) x[i] = vmul(x[i], rtw<INV, false, D, 8>(t, nullptr, g, (uint32_t)brev_c(i, LOGR), i));
        const uint32_t b0 = hx<D>(h) ^ sigma(g);
        TS_UNROLL
        for (int c = 0; c < R; c++) st4(tile + b0 + 256 * c, x[c]);
    }
}

// ---- last round straight to global memory -------------------------------------------------------------------------------
// item = (blk, h): positions 4 blk + c of quad h.  POST: position p is multiplied by post[p] (16-byte pairs, 2 LDG.128).
template <int D, bool INV, bool POST, int NT, int NQv = dnq(D)>
TS_D void r4_st(const uint4 *tile, char *dst_h, bool valid, size_t row_base, int log_stride, size_t pitch_b, const uint2 *post,
                int tid) {
    using G = Geo<D, NQv>;
    constexpr int ITEMS = (G::L / 4) * G::NQ;
    static_assert(ITEMS % NT == 0 && NT % G::NQ == 0, "r4_st: thread count");
    if (!valid) return;
    const uint32_t h = tid & (G::NQ - 1);
    const size_t step = ((size_t)1 << log_stride) * pitch_b;
    TS_NOUNROLL
    for (int k = 0; k < ITEMS / NT; k++) {
        const uint32_t blk = io_blk<G::NQ>((uint32_t)(tid + k * NT) / G::NQ);
        const uint32_t base = hx<D>(h) ^ (blk << 2) ^ ((blk >> 1) & 3u) ^ (((blk >> 3) & 1u) << 2);
        uint4 pa, pb;
        if (POST) {
            const uint4 *pp = reinterpret_cast<const uint4 *>(post + 4 * blk);
            pa = __ldg(pp);
            pb = __ldg(pp + 1);
        }
        V4 x[4];
        TS_UNROLL
        for (int c = 0; c < 4; c++) x[c] = ld4(tile + (base ^ (uint32_t)c));
        dft_v4<2, INV, POST, 0>(x);
        if (POST) {
            x[0] = vmul(x[0], make_uint2(pa.x, pa.y));
            x[1] = vmul(x[1], make_uint2(pa.z, pa.w));
            x[2] = vmul(x[2], make_uint2(pb.x, pb.y));
            x[3] = vmul(x[3], make_uint2(pb.z, pb.w));
        }
        char *ptr = dst_h + (row_base + ((size_t)(4 * blk) << log_stride)) * pitch_b;
        TS_UNROLL
        for (int c = 0; c < 4; c++) st4(reinterpret_cast<uint4 *>(ptr + (size_t)c * step), x[c]);
    }
}

#ifndef TS_V4_MINBLOCKS
#define TS_V4_MINBLOCKS 3
#endif
constexpr int V4_NT = 256;

template <int D, bool INV, bool MID, bool CSRC>
__global__ void __launch_bounds__(V4_NT, TS_V4_MINBLOCKS) pass_kernel(PassParams p) {
    TS_DYN_SMEM(uint4, tile);
    using G = Geo<D>;
    static_assert(!MID || CSRC, "MID reads a contiguous blocked tile");
    const int tid = threadIdx.x;
    uint32_t cs, tile_id;
    if (p.cs_shift >= 0) {  // column slices are the fastest-varying part of blockIdx (neighbouring CTAs share rows, tables)
        cs = blockIdx.x & ((1u << p.cs_shift) - 1);
        tile_id = blockIdx.x >> p.cs_shift;
    } else {
        cs = blockIdx.x % p.n_col_slices;
        tile_id = blockIdx.x / p.n_col_slices;
    }
    const uint32_t col = (cs << (14 - D)) + 4 * (tid & (G::NQ - 1));
    const bool valid = col < p.ncols;
    const uint32_t colc = valid ? col : 0;
    const char *src_h = reinterpret_cast<const char *>(p.src) + col_off_bytes(colc, p.src_slice);
    char *dst_h = reinterpret_cast<char *>(p.dst) + col_off_bytes(colc, p.dst_slice);
    const size_t sp = row_pitch_bytes(p.src_pitch, p.src_slice), dp = row_pitch_bytes(p.dst_pitch, p.dst_slice);
    size_t src_base, dst_base;
    int dst_log_stride;
    const uint2 *post = p.post;
    if (MID) {
        const uint32_t j = tile_id & ((1u << p.b) - 1), Kc = tile_id >> p.b;
        src_base = (size_t)brev_bits(Kc, p.klo_bits) << D;
        dst_base = ((size_t)brev_bits(j, p.b) << p.m) + Kc;
        dst_log_stride = p.klo_bits;
        if (post) post += (((size_t)j << p.klo_bits) + Kc) << D;
        r1_ld<D, INV, true, true, true, V4_NT>(tile, src_h, valid, src_base, 0, sp, p.t, p.pre + ((size_t)j << D), tid);
    } else {
        const uint32_t lo = tile_id & ((1u << p.lo_bits) - 1), hi = tile_id >> p.lo_bits;
        src_base = dst_base = ((size_t)hi << (p.lo_bits + D)) + lo;
        dst_log_stride = p.lo_bits;
        if (post) post += (size_t)lo << D;
        r1_ld<D, INV, false, false, CSRC, V4_NT>(tile, src_h, valid, src_base, p.lo_bits, sp, p.t, nullptr, tid);
    }
    __syncthreads();
    nttp::dif_r2<D, INV, false, V4_NT>(tile, p.t, nullptr, tid);
    __syncthreads();
    nttp::dif_r3<D, INV, false, V4_NT>(tile, p.t, nullptr, tid);
    __syncthreads();
    if (post) r4_st<D, INV, true, V4_NT>(tile, dst_h, valid, dst_base, dst_log_stride, dp, post, tid);
    else r4_st<D, INV, false, V4_NT>(tile, dst_h, valid, dst_base, dst_log_stride, dp, nullptr, tid);
}

// ---- table builders (run once per shape, cached in the context) -----------------------------------------------------------
// post[lo][p] = w_{2^m}^(+-lo * brev_D(p)) out of the 2^big_log power table
__global__ void fill_post_table_kernel(uint2 *dst, const uint2 *tw_big, int big_log, int m, int D, int lo_bits, int inverse) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ((size_t)1 << (lo_bits + D))) return;
    const uint32_t lo = (uint32_t)(e >> D), p = (uint32_t)e & ((1u << D) - 1);
    uint32_t idx = (lo * brev_bits(p, D)) << (big_log - m);
    if (inverse) idx = ((1u << big_log) - idx) & ((1u << big_log) - 1);
    dst[e] = tw_big[idx];
}
// post[j][Kc][p] = w_{2^m}^(Kc * brev_D(p)) * lane[j][Kc]   (canonical product with its Shoup companion)
__global__ void fill_mid_post_table_kernel(uint2 *dst, const uint2 *tw_big, const uint2 *lane, int big_log, int m, int D,
                                           int klo_bits, int b) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ((size_t)1 << (b + klo_bits + D))) return;
    const uint32_t p = (uint32_t)e & ((1u << D) - 1);
    const size_t jk = e >> D;  // (j << klo_bits) + Kc
    const uint32_t Kc = (uint32_t)jk & ((1u << klo_bits) - 1);
    const uint32_t w0 = tw_big[(size_t)(Kc * brev_bits(p, D)) << (big_log - m)].x;
    const uint32_t w = (uint32_t)(((uint64_t)w0 * lane[jk].x) % bb::P);
    dst[e] = make_uint2(w, bb::shoup_prime(w));
}
// pre4[j][g][c] = pre[j][g + 256 c]
__global__ void fill_pre_table_kernel(uint2 *dst, const uint2 *pre, int D, int b) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (1u << (b + D))) return;
    const int LOGR = D - 8;
    const uint32_t j = e >> D, r = e & ((1u << D) - 1), g = r >> LOGR, c = r & ((1u << LOGR) - 1);
    dst[e] = pre[((size_t)j << D) + g + 256u * c];
}

}  // namespace ntt4
