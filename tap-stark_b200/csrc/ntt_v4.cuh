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

// =====================================================================================================================
// TMA variant of the contiguous-source passes (P2, P4): the 64 KiB tile is staged by `cp.async.bulk.tensor` (one elected
// thread, one bulk tensor copy per 4-column plane, completion on an mbarrier) instead of 16 LDG.128 per thread.  The
// source is a blocked matrix [col / 8][row][8 columns]; as a tensor it is {8 columns, 256 rows, rows / 256, column groups}
// and the box {4, 256, 2^D / 256, 1} is exactly one plane of the position-major tile: row p of the plane lands at 16-byte
// unit p (profiles/tools/tma_probe.cu checks this on the GPU).  The hardware swizzle modes cannot produce the tile's
// sigma(p): with a 16- or 32-byte inner box CU_TENSOR_MAP_SWIZZLE_128B faults (the box row is narrower than the swizzle
// span; same probe), so the plane is staged UNSWIZZLED and the first round, which reads 32 consecutive units per warp
// instruction (conflict free as they are), writes its results into the swizzled layout the other rounds use.
// Opt-in (TS_TMA=1): measured against the LDG form in profiles/r02/README.md.  Device builds only.
#ifndef TS_EMULATE
#include <cuda.h>

namespace ntt4 {

TS_D uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
TS_D void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
TS_D void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
TS_D void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
TS_D void tma_load_4d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// first round out of the TMA-staged tile: plane h holds row p at unit (h << D) + p.  Lanes walk 32 consecutive groups of one
// plane and the results go to the swizzled layout of the other rounds, unit hx(h) ^ sigma(g) + 256 c: that is the staged
// unit of group g' = sigma(g) ^ s3(h), which differs from g in its low three bits only -- an item of the SAME warp reads it
// in this iteration, hence the __syncwarp between the loads and the stores.
template <int D, bool INV, int NT, int NQv = dnq(D)>
TS_D void r1_sm(uint4 *tile, const FastTables &t, int tid) {
    using G = Geo<D, NQv>;
    constexpr int LOGR = G::LOGR1, R = 1 << LOGR;
    constexpr int ITEMS = 256 * G::NQ;
    static_assert(ITEMS % NT == 0 && NT % 32 == 0 && 256 % NT == 0 || NT % 256 == 0, "r1_sm: thread count");
    TS_NOUNROLL
    for (int k = 0; k < ITEMS / NT; k++) {
        const uint32_t item = (uint32_t)(tid + k * NT), g = item & 255, h = item >> 8;
        const uint32_t src0 = (h << D) + g;
        V4 x[R];
        TS_UNROLL
        for (int c = 0; c < R; c++) x[c] = ld4(tile + src0 + 256 * c);
        __syncwarp();
        dft_v4<LOGR, INV, true>(x);
        TS_UNROLL
        for (int i = 1; i < R; i++) x[i] = vmul(x[i], rtw<INV, false, D, 8>(t, nullptr, g, (uint32_t)brev_c(i, LOGR), i));
        const uint32_t b0 = hx<D>(h) ^ sigma(g);
        TS_UNROLL
        for (int c = 0; c < R; c++) st4(tile + b0 + 256 * c, x[c]);
    }
}

// ---- warp-shuffle exchange between the last two rounds (TS_SHFL=1) ---------------------------------------------------------
// R3 (radix 8, stride 4) leaves a thread with the 8 positions 32 blk + g + 4 c of one quad; R4 (radix 4, stride 1) wants the
// 4 positions 4 blk4 + c' with blk4 = 8 blk + c and c' = g.  The four lanes g = 0..3 of a group (lane bits 1..2; bit 0 is the
// quad) therefore hold an 8 x 4 block that has to be transposed: two butterfly exchanges (__shfl_xor by 4, then by 2), after
// which lane t owns the R4 items c = 2t and 2t + 1.  No shared-memory round trip and no barrier between R3 and R4 --
// at the price of 32 SHFL + 64 SEL per 32 elements where the shared-memory form issues 16 LDS/STS.128
// (profiles/r02/README.md has the measured comparison).
TS_D V4 shfl4(const V4 &a, int mask) {
    V4 r;
    TS_UNROLL
    for (int k = 0; k < 4; k++) r.v[k] = __shfl_xor_sync(0xffffffffu, a.v[k], mask);
    return r;
}
TS_D V4 sel4(bool p, const V4 &a, const V4 &b) {
    V4 r;
    TS_UNROLL
    for (int k = 0; k < 4; k++) r.v[k] = p ? a.v[k] : b.v[k];
    return r;
}
template <int D, bool INV, bool POST, int NT, int NQv = dnq(D)>
TS_D void r34_shfl_st(uint4 *tile, char *dst_h, bool valid, size_t row_base, int log_stride, size_t pitch_b, const FastTables &t,
                      const uint2 *post, int tid) {
    using G = Geo<D, NQv>;
    static_assert(G::NQ == 2, "r34_shfl_st: lane bit 0 = quad, bits 1..2 = g (D = 11 tiles)");
    constexpr int ITEMS = (G::L / 8) * G::NQ;
    static_assert(ITEMS % NT == 0, "r34_shfl_st: thread count");
    const size_t step = ((size_t)1 << log_stride) * pitch_b;
    TS_NOUNROLL
    for (int k = 0; k < ITEMS / NT; k++) {
        const uint32_t item = (uint32_t)(tid + k * NT);
        const uint32_t h = item & 1u, g = (item >> 1) & 3u, blk = item >> 3;  // lanes: quad, then g, then block
        const uint32_t base = hx<D>(h) ^ ((blk << 5) | g) ^ ((blk & 1u) << 2);
        V4 x[8];
        TS_UNROLL
        for (int c = 0; c < 8; c++) x[c] = ld4(tile + (base ^ (uint32_t)((4 * c) ^ ((c >> 1) & 3))));
        dft_v4<3, INV, true>(x);
        TS_UNROLL
        for (int i = 1; i < 8; i++) x[i] = vmul(x[i], rtw<INV, false, 5, 2>(t, nullptr, g, (uint32_t)brev_c(i, 3), i));
        // the shared-memory form stores register c to position 32 blk + g + 4 c: "digit c" below is register c
        // exchange A (g bit 1): keep digits [0,4) or [4,8), receive the partner's same digits
        const bool hiA = (g & 2u) != 0;
        V4 own[4], oth[4];
        TS_UNROLL
        for (int i = 0; i < 4; i++) {
            const V4 send = sel4(hiA, x[i], x[i + 4]);
            own[i] = sel4(hiA, x[i + 4], x[i]);
            oth[i] = shfl4(send, 4);
        }
        // now own[i] / oth[i]: digit c = 4 hiA + i from lane g and from lane g ^ 2
        const bool hiB = (g & 1u) != 0;
        V4 y[2][4];  // y[kk][g'] : R4 item digit c = 2 t + kk, position c' = g'
        TS_UNROLL
        for (int kk = 0; kk < 2; kk++) {
            // keep digits {2 hiB, 2 hiB + 1} of my half, send the other two
            const V4 keep_own = sel4(hiB, own[2 + kk], own[kk]), keep_oth = sel4(hiB, oth[2 + kk], oth[kk]);
            const V4 send_own = sel4(hiB, own[kk], own[2 + kk]), send_oth = sel4(hiB, oth[kk], oth[2 + kk]);
            const V4 recv_own = shfl4(send_own, 2), recv_oth = shfl4(send_oth, 2);
            // sources: keep_own from lane g, keep_oth from g ^ 2, recv_own from g ^ 1, recv_oth from g ^ 3.  Position g' wants the
            // value of lane g' = g ^ m: low bit of m picks kept / received, high bit picks own / other (two select levels)
            const V4 a_same = keep_own, a_flip = recv_own, b_same = keep_oth, b_flip = recv_oth;
            TS_UNROLL
            for (int gp = 0; gp < 4; gp++) {
                const bool m0 = ((gp & 1) != 0) != hiB, m1 = ((gp & 2) != 0) != hiA;
                y[kk][gp] = sel4(m1, sel4(m0, b_flip, b_same), sel4(m0, a_flip, a_same));
            }
        }
        const uint32_t t4 = (g & 1u) * 1u + (g & 2u);  // lane's item pair index t = 2 hiA + hiB  (digits 4 hiA + 2 hiB + kk)
        TS_UNROLL
        for (int kk = 0; kk < 2; kk++) {
            const uint32_t blk4 = 8 * blk + 2 * t4 + kk;
            V4 z[4];
            TS_UNROLL
            for (int cp = 0; cp < 4; cp++) z[cp] = y[kk][cp];
            dft_v4<2, INV, POST, 0>(z);
            if (POST) {
                const uint4 *pp = reinterpret_cast<const uint4 *>(post + 4 * blk4);
                const uint4 pa = __ldg(pp), pb = __ldg(pp + 1);
                z[0] = vmul(z[0], make_uint2(pa.x, pa.y));
                z[1] = vmul(z[1], make_uint2(pa.z, pa.w));
                z[2] = vmul(z[2], make_uint2(pb.x, pb.y));
                z[3] = vmul(z[3], make_uint2(pb.z, pb.w));
            }
            if (valid) {
                char *ptr = dst_h + (row_base + ((size_t)(4 * blk4) << log_stride)) * pitch_b;
                TS_UNROLL
                for (int cp = 0; cp < 4; cp++) st4(reinterpret_cast<uint4 *>(ptr + (size_t)cp * step), z[cp]);
            }
        }
    }
}

template <int D, bool INV, bool MID, bool CSRC>
__global__ void __launch_bounds__(V4_NT, 2) pass_shfl_kernel(PassParams p) {
    TS_DYN_SMEM(uint4, tile);
    using G = Geo<D>;
    const int tid = threadIdx.x;
    uint32_t cs, tile_id;
    if (p.cs_shift >= 0) {
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
    size_t dst_base;
    int dst_log_stride;
    const uint2 *post = p.post;
    if (MID) {
        const uint32_t j = tile_id & ((1u << p.b) - 1), Kc = tile_id >> p.b;
        dst_base = ((size_t)brev_bits(j, p.b) << p.m) + Kc;
        dst_log_stride = p.klo_bits;
        if (post) post += (((size_t)j << p.klo_bits) + Kc) << D;
        r1_ld<D, INV, true, true, true, V4_NT>(tile, src_h, valid, (size_t)brev_bits(Kc, p.klo_bits) << D, 0, sp, p.t,
                                               p.pre + ((size_t)j << D), tid);
    } else {
        const uint32_t lo = tile_id & ((1u << p.lo_bits) - 1), hi = tile_id >> p.lo_bits;
        dst_base = ((size_t)hi << (p.lo_bits + D)) + lo;
        dst_log_stride = p.lo_bits;
        if (post) post += (size_t)lo << D;
        r1_ld<D, INV, false, false, CSRC, V4_NT>(tile, src_h, valid, dst_base, p.lo_bits, sp, p.t, nullptr, tid);
    }
    __syncthreads();
    nttp::dif_r2<D, INV, false, V4_NT>(tile, p.t, nullptr, tid);
    __syncthreads();
    // the quad a lane works on in the fused last rounds is lane bit 0 = tid & 1 = the quad of its column pointers
    if (post) r34_shfl_st<D, INV, true, V4_NT>(tile, dst_h, valid, dst_base, dst_log_stride, dp, p.t, post, tid);
    else r34_shfl_st<D, INV, false, V4_NT>(tile, dst_h, valid, dst_base, dst_log_stride, dp, p.t, nullptr, tid);
}

template <int D, bool INV>
__global__ void __launch_bounds__(V4_NT, TS_V4_MINBLOCKS) pass_tma_kernel(PassParams p, const __grid_constant__ CUtensorMap src_map) {
    extern __shared__ unsigned char tma_raw_[];
    using G = Geo<D>;
    // bulk tensor copies want a 128-byte aligned destination; 1024 keeps the planes on bank-row boundaries
    // (offset added to the __shared__ base, so the compiler keeps LDS/STS instead of generic loads)
    uint4 *tile = reinterpret_cast<uint4 *>(tma_raw_ + ((1024u - (smem_u32(tma_raw_) & 1023u)) & 1023u));
    uint64_t *bar = reinterpret_cast<uint64_t *>(tile + G::UNITS);
    const int tid = threadIdx.x;
    uint32_t cs, tile_id;
    if (p.cs_shift >= 0) {
        cs = blockIdx.x & ((1u << p.cs_shift) - 1);
        tile_id = blockIdx.x >> p.cs_shift;
    } else {
        cs = blockIdx.x % p.n_col_slices;
        tile_id = blockIdx.x / p.n_col_slices;
    }
    const uint32_t col0 = cs << (14 - D);
    const uint32_t col = col0 + 4 * (tid & (G::NQ - 1));
    const bool valid = col < p.ncols;
    char *dst_h = reinterpret_cast<char *>(p.dst) + col_off_bytes(valid ? col : 0, p.dst_slice);
    const size_t dp = row_pitch_bytes(p.dst_pitch, p.dst_slice);
    const uint32_t lo = tile_id & ((1u << p.lo_bits) - 1), hi = tile_id >> p.lo_bits;  // contiguous source: lo_bits == 0
    const size_t row_base = ((size_t)hi << (p.lo_bits + D)) + lo;
    if (tid == 0) {
        mbar_init(bar, 1);
        uint32_t planes = 0;
        for (int h = 0; h < G::NQ; h++) planes += (col0 + 4 * h < p.ncols) ? 1u : 0u;
        mbar_expect_tx(bar, planes * (uint32_t)(G::L * 16));
        for (int h = 0; h < G::NQ; h++) {
            const uint32_t ch = col0 + 4 * h;
            if (ch < p.ncols) tma_load_4d(tile + ((size_t)h << D), &src_map, bar, (int)(ch & 7u), 0, (int)(row_base >> 8), (int)(ch >> 3));
        }
    }
    __syncthreads();  // the barrier is initialised before anyone polls it
    mbar_wait(bar, 0);
    r1_sm<D, INV, V4_NT>(tile, p.t, tid);
    __syncthreads();
    nttp::dif_r2<D, INV, false, V4_NT>(tile, p.t, nullptr, tid);
    __syncthreads();
    nttp::dif_r3<D, INV, false, V4_NT>(tile, p.t, nullptr, tid);
    __syncthreads();
    r4_st<D, INV, false, V4_NT>(tile, dst_h, valid, row_base, p.lo_bits, dp, nullptr, tid);
}

}  // namespace ntt4
#endif  // !TS_EMULATE
