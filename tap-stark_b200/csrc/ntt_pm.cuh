// ntt_pm.cuh -- "position-major" digit kernels: the third and fastest NTT path (wide matrices, D in {9,10,11}).
//
// Same mathematics and tile shape as ntt_fast.cuh (2^D positions x 2^(14-D) lanes = 64 KiB, lanes = adjacent
// columns), but the shared-memory tile is laid out position-major in 16-byte units:
//
//     unit(h, p) = (h * L) ^ s3(h) ^ sigma(p)      h = lane / 4   ("quad": 4 adjacent columns, one uint4)
//     sigma(p)   = p ^ ((p >> 3) & 3) ^ (((p >> 5) & 1) << 2),   s3(h) = 3-bit reversal of h (staggers tile I/O)
//
// and a work item is (butterfly group, quad): a thread holds R positions x 4 columns in registers.  Consequences
// (profiles/r01: in ntt_fast tile I/O took 56 % of the time with 31 % of the instructions, and the rounds stalled
// on the shared-memory instruction queue):
//   * every shared access is LDS.128 / STS.128, every global access LDG.128 / STG.128 straight to/from the tile
//     (0.5 instr per element for I/O instead of ~8), twiddles are fetched once per 4 columns;
//   * rounds are radix 8 (R positions x 4 columns = 32 data registers): 2^D = 2^(D-8) * 8 * 8 * 4;
//   * sigma makes all four round shapes (strides 256, 32, 4, 1) and the tile I/O conflict free at 16-byte
//     granularity, and element addresses are  base + const  or  base ^ const.
#pragma once
#include "ntt_fast.cuh"

namespace nttp {

using ntt::brev_bits;
using ntt::brev_c;
using nttf::btw;
using nttf::FastTables;

TS_D uint32_t sigma(uint32_t p) { return p ^ ((p >> 3) & 3u) ^ (((p >> 5) & 1u) << 2); }
// quad offset: everything below combines with XOR only (or adds constants into zero bit fields), so element
// addresses are  hx ^ const  /  hx' + const
template <int D>
TS_D uint32_t hx(uint32_t h) {
    return (h << D) ^ (((h & 1u) << 2) | (h & 2u) | ((h >> 2) & 1u));
}

// 4-wide field helpers (one uint4 = the same position in 4 adjacent columns)
struct V4 {
    uint32_t v[4];
};
TS_D V4 ld4(const uint4 *p) {
    const uint4 t = *p;
    return V4{{t.x, t.y, t.z, t.w}};
}
TS_D void st4(uint4 *p, const V4 &a) { *p = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]); }
TS_D V4 vadd(const V4 &a, const V4 &b) {
    V4 r;
    TS_UNROLL
    for (int k = 0; k < 4; k++) r.v[k] = bb::add(a.v[k], b.v[k]);
    return r;
}
TS_D V4 vsub(const V4 &a, const V4 &b) {
    V4 r;
    TS_UNROLL
    for (int k = 0; k < 4; k++) r.v[k] = bb::sub(a.v[k], b.v[k]);
    return r;
}
// (a - b) * w, w a Shoup pair
TS_D V4 vsubmul(const V4 &a, const V4 &b, uint32_t w, uint32_t wp) {
    V4 r;
    TS_UNROLL
    for (int k = 0; k < 4; k++) r.v[k] = bb::shoup(a.v[k] - b.v[k] + bb::P, w, wp);
    return r;
}
TS_D V4 vmul(const V4 &a, uint2 w) {
    V4 r;
    TS_UNROLL
    for (int k = 0; k < 4; k++) r.v[k] = bb::shoup(a.v[k], w.x, w.y);
    return r;
}
TS_D V4 vmul2(const V4 &a, uint2 w1, uint2 w2) {
    V4 r;
    TS_UNROLL
    for (int k = 0; k < 4; k++) r.v[k] = bb::shoup(bb::shoup_lazy(a.v[k], w1.x, w1.y), w2.x, w2.y);
    return r;
}

// unreduced sum / difference for operands of a Shoup multiplication (which takes any u32): a + b < 2p, a - b + p < 2p
TS_D V4 vadd_raw(const V4 &a, const V4 &b) {
    V4 r;
    TS_UNROLL
    for (int k = 0; k < 4; k++) r.v[k] = a.v[k] + b.v[k];
    return r;
}
TS_D V4 vsub_raw(const V4 &a, const V4 &b) {
    V4 r;
    TS_UNROLL
    for (int k = 0; k < 4; k++) r.v[k] = a.v[k] - b.v[k] + bb::P;
    return r;
}

// radix-2^LOGR DIF DFT on R quads, compile-time twiddles; natural in, register i holds frequency brev(i).
// LAZY: the caller multiplies every output register >= LAZY_FROM by a twiddle next, so the last stage leaves those
// outputs unreduced in [0, 2p) (one instruction per sum / difference instead of two).
template <int LOGR, bool INV, bool LAZY = false, int LAZY_FROM = 1>
TS_D void dft_v4(V4 (&x)[1 << LOGR]) {
    constexpr bb::InnerTw<LOGR, INV> T{};
    constexpr int R = 1 << LOGR;
    TS_UNROLL
    for (int s = 0; s < LOGR; s++) {
        const int half = R >> (s + 1);
        const bool last = LAZY && s == LOGR - 1;
        TS_UNROLL
        for (int blk = 0; blk < R; blk += 2 * half) {
            TS_UNROLL
            for (int j = 0; j < half; j++) {
                const V4 a = x[blk + j], b = x[blk + j + half];
                x[blk + j] = (last && blk + j >= LAZY_FROM) ? vadd_raw(a, b) : vadd(a, b);
                const int e = j << s;
                x[blk + j + half] = e == 0 ? ((last && blk + j + half >= LAZY_FROM) ? vsub_raw(a, b) : vsub(a, b))
                                           : vsubmul(a, b, T.w[e], T.wp[e]);
            }
        }
    }
}

// inter-round twiddle w^(+-g*u) of a round with 2^LOGS points in 2^LOGG groups.  SM: shared [i][g] table (8-column
// lde_mid).  Otherwise the product-indexed global table (FastTables::rt): the lanes of a warp hold consecutive g, so
// the load is one coalesced L1 hit instead of a 32-line gather out of a plain power table indexed by g*u.
template <bool INV, bool SM, int LOGS, int LOGG>
TS_D uint2 rtw(const FastTables &t, const uint2 *sm_tab, uint32_t g, uint32_t u, uint32_t i) {
    if (SM) return sm_tab[(i << LOGG) + g];
    constexpr int OFF = LOGS == 8 ? nttf::RT_R2_OFF : LOGS == 5 ? nttf::RT_R3_OFF : nttf::RT_R1_OFF(LOGS);
    return __ldg(t.rt[INV ? 1 : 0] + OFF + ((uint32_t)brev_c((int)u, LOGS - LOGG) << LOGG) + g);
}

// default tile: 2^D positions x 2^(14-D) columns = 64 KiB (NQv quads of 4 columns per position)
__host__ __device__ constexpr int dnq(int D) { return (1 << (14 - D)) / 4; }
template <int D, int NQv = dnq(D)>
struct Geo {
    static constexpr int L = 1 << D;
    static constexpr int NQ = NQv;                // quads per position
    static constexpr int K = 4 * NQ;              // lanes (columns) per tile
    static constexpr int UNITS = NQ * L;          // 16-byte units per tile
    static constexpr int LOGR1 = D - 8;           // first-round radix (2, 4 or 8)
};

// ---- DIF rounds (forward direction of the data flow; INV only selects the roots) ---------------------------
// R1: radix 2^(D-8), stride 256.  PRE (lde_mid): position c of the item is multiplied by pw[c] and by `lw`.
template <int D, bool INV, bool PRE, bool SM, int NT, int NQv = dnq(D)>
TS_D void dif_r1(const uint4 *src, uint4 *dst, const FastTables &t, const uint2 *sm_tab, const uint2 *pre_tab, uint2 lw,
                 int tid) {
    using G = Geo<D, NQv>;
    constexpr int LOGR = G::LOGR1, R = 1 << LOGR;
    for (int item = tid; item < 256 * G::NQ; item += NT) {
        const uint32_t g = item & 255, h = item >> 8;
        const uint32_t b0 = hx<D>(h) ^ sigma(g);  // sigma(g + 256 c) = sigma(g) + 256 c
        V4 x[R];
        TS_UNROLL
        for (int c = 0; c < R; c++) x[c] = ld4(src + b0 + 256 * c);
        if (PRE) {
            TS_UNROLL
            for (int c = 0; c < R; c++) x[c] = vmul2(x[c], __ldg(pre_tab + g + 256 * c), lw);
        }
        dft_v4<LOGR, INV, true>(x);
        TS_UNROLL
        for (int i = 1; i < R; i++) x[i] = vmul(x[i], rtw<INV, SM, D, 8>(t, sm_tab, g, (uint32_t)brev_c(i, LOGR), i));
        TS_UNROLL
        for (int c = 0; c < R; c++) st4(dst + b0 + 256 * c, x[c]);
    }
}
// R2: radix 8 inside blocks of 256, stride 32: p = 256 blk + g + 32 c
template <int D, bool INV, bool SM, int NT, int NQv = dnq(D)>
TS_D void dif_r2(uint4 *tile, const FastTables &t, const uint2 *sm_tab, int tid) {
    using G = Geo<D, NQv>;
    constexpr int ITEMS = (G::L / 8) * G::NQ;
    for (int item = tid; item < ITEMS; item += NT) {
        const uint32_t grp = item & (G::L / 8 - 1), h = item / (G::L / 8);
        const uint32_t blk = grp >> 5, g = grp & 31;
        const uint32_t b0 = hx<D>(h) ^ ((blk << 8) | (g ^ ((g >> 3) & 3u))), b1 = b0 ^ 4u;  // bit 2 ^= c & 1
        V4 x[8];
        TS_UNROLL
        for (int c = 0; c < 8; c++) x[c] = ld4(tile + ((c & 1) ? b1 : b0) + 32 * c);
        dft_v4<3, INV, true>(x);
        TS_UNROLL
        for (int i = 1; i < 8; i++) x[i] = vmul(x[i], rtw<INV, SM, 8, 5>(t, sm_tab, g, (uint32_t)brev_c(i, 3), i));
        TS_UNROLL
        for (int c = 0; c < 8; c++) st4(tile + ((c & 1) ? b1 : b0) + 32 * c, x[c]);
    }
}
// R3: radix 8 inside blocks of 32, stride 4: p = 32 blk + g + 4 c
template <int D, bool INV, bool SM, int NT, int NQv = dnq(D)>
TS_D void dif_r3(uint4 *tile, const FastTables &t, const uint2 *sm_tab, int tid) {
    using G = Geo<D, NQv>;
    constexpr int ITEMS = (G::L / 8) * G::NQ;
    for (int item = tid; item < ITEMS; item += NT) {
        const uint32_t grp = item & (G::L / 8 - 1), h = item / (G::L / 8);
        const uint32_t blk = grp >> 2, g = grp & 3;
        const uint32_t base = hx<D>(h) ^ ((blk << 5) | g) ^ ((blk & 1u) << 2);
        V4 x[8];
        TS_UNROLL
        for (int c = 0; c < 8; c++) x[c] = ld4(tile + (base ^ (uint32_t)((4 * c) ^ ((c >> 1) & 3))));
        dft_v4<3, INV, true>(x);
        TS_UNROLL
        for (int i = 1; i < 8; i++) x[i] = vmul(x[i], rtw<INV, SM, 5, 2>(t, sm_tab, g, (uint32_t)brev_c(i, 3), i));
        TS_UNROLL
        for (int c = 0; c < 8; c++) st4(tile + (base ^ (uint32_t)((4 * c) ^ ((c >> 1) & 3))), x[c]);
    }
}
// R4: radix 4 on consecutive positions p = 4 blk + c.  POST: inter-digit twiddle w^(+-lo*brev_D(p)) per position.
template <int D, bool INV, bool POST, int NT, int NQv = dnq(D)>
TS_D void dif_r4(uint4 *tile, const FastTables &t, uint32_t lo, int tw_shift, int tid) {
    using G = Geo<D, NQv>;
    constexpr int ITEMS = (G::L / 4) * G::NQ;
    for (int item = tid; item < ITEMS; item += NT) {
        const uint32_t blk = item & (G::L / 4 - 1), h = item / (G::L / 4);
        const uint32_t base = hx<D>(h) ^ (blk << 2) ^ ((blk >> 1) & 3u) ^ (((blk >> 3) & 1u) << 2);
        uint2 pt[4];
        if (POST) {
            TS_UNROLL
            for (int c = 0; c < 4; c++) pt[c] = btw<INV>(t, (lo * brev_bits(4 * blk + c, D)) << tw_shift);
        }
        V4 x[4];
        TS_UNROLL
        for (int c = 0; c < 4; c++) x[c] = ld4(tile + (base ^ (uint32_t)c));
        dft_v4<2, INV>(x);
        if (POST) {
            TS_UNROLL
            for (int c = 0; c < 4; c++) x[c] = vmul(x[c], pt[c]);
        }
        TS_UNROLL
        for (int c = 0; c < 4; c++) st4(tile + (base ^ (uint32_t)c), x[c]);
    }
}

// ---- DIT rounds (inverse sub-transform of lde_mid: bit-reversed positions in, natural out) ------------------
// block i of a round holds the sub-sequence with digit brev(i); digit c is multiplied by w^-(c j) before the DFT.
template <int D, int NT, int NQv = dnq(D)>
TS_D void dit_r4(uint4 *tile, int tid) {  // radix 4, consecutive positions, no twiddles
    using G = Geo<D, NQv>;
    constexpr int ITEMS = (G::L / 4) * G::NQ;
    for (int item = tid; item < ITEMS; item += NT) {
        const uint32_t blk = item & (G::L / 4 - 1), h = item / (G::L / 4);
        const uint32_t base = hx<D>(h) ^ (blk << 2) ^ ((blk >> 1) & 3u) ^ (((blk >> 3) & 1u) << 2);
        V4 v[4];
        TS_UNROLL
        for (int c = 0; c < 4; c++) v[c] = ld4(tile + (base ^ (uint32_t)brev_c(c, 2)));
        dft_v4<2, true>(v);
        TS_UNROLL
        for (int u = 0; u < 4; u++) st4(tile + (base ^ (uint32_t)u), v[brev_c(u, 2)]);
    }
}
template <int D, bool SM, int NT, int NQv = dnq(D)>
TS_D void dit_r3(uint4 *tile, const FastTables &t, const uint2 *sm_tab, int tid) {  // radix 8, stride 4
    using G = Geo<D, NQv>;
    constexpr int ITEMS = (G::L / 8) * G::NQ;
    for (int item = tid; item < ITEMS; item += NT) {
        const uint32_t grp = item & (G::L / 8 - 1), h = item / (G::L / 8);
        const uint32_t blk = grp >> 2, j = grp & 3;
        const uint32_t base = hx<D>(h) ^ ((blk << 5) | j) ^ ((blk & 1u) << 2);
        V4 v[8];
        TS_UNROLL
        for (int c = 0; c < 8; c++) {
            const int i = brev_c(c, 3);
            v[c] = ld4(tile + (base ^ (uint32_t)((4 * i) ^ ((i >> 1) & 3))));
        }
        TS_UNROLL
        for (int c = 1; c < 8; c++) v[c] = vmul(v[c], rtw<true, SM, 5, 2>(t, sm_tab, j, (uint32_t)c, c));
        dft_v4<3, true>(v);
        TS_UNROLL
        for (int u = 0; u < 8; u++) st4(tile + (base ^ (uint32_t)((4 * u) ^ ((u >> 1) & 3))), v[brev_c(u, 3)]);
    }
}
template <int D, bool SM, int NT, int NQv = dnq(D)>
TS_D void dit_r2(uint4 *tile, const FastTables &t, const uint2 *sm_tab, int tid) {  // radix 8, stride 32
    using G = Geo<D, NQv>;
    constexpr int ITEMS = (G::L / 8) * G::NQ;
    for (int item = tid; item < ITEMS; item += NT) {
        const uint32_t grp = item & (G::L / 8 - 1), h = item / (G::L / 8);
        const uint32_t blk = grp >> 5, j = grp & 31;
        const uint32_t b0 = hx<D>(h) ^ ((blk << 8) | (j ^ ((j >> 3) & 3u))), b1 = b0 ^ 4u;
        V4 v[8];
        TS_UNROLL
        for (int c = 0; c < 8; c++) {
            const int i = brev_c(c, 3);
            v[c] = ld4(tile + ((i & 1) ? b1 : b0) + 32 * i);
        }
        TS_UNROLL
        for (int c = 1; c < 8; c++) v[c] = vmul(v[c], rtw<true, SM, 8, 5>(t, sm_tab, j, (uint32_t)c, c));
        dft_v4<3, true>(v);
        TS_UNROLL
        for (int u = 0; u < 8; u++) st4(tile + ((u & 1) ? b1 : b0) + 32 * u, v[brev_c(u, 3)]);
    }
}
template <int D, bool SM, int NT, int NQv = dnq(D)>
TS_D void dit_r1(uint4 *tile, const FastTables &t, const uint2 *sm_tab, int tid) {  // radix 2^(D-8), stride 256
    using G = Geo<D, NQv>;
    constexpr int LOGR = G::LOGR1, R = 1 << LOGR;
    for (int item = tid; item < 256 * G::NQ; item += NT) {
        const uint32_t j = item & 255, h = item >> 8;
        const uint32_t b0 = hx<D>(h) ^ sigma(j);
        V4 v[R];
        TS_UNROLL
        for (int c = 0; c < R; c++) v[c] = ld4(tile + b0 + 256 * brev_c(c, LOGR));
        TS_UNROLL
        for (int c = 1; c < R; c++) v[c] = vmul(v[c], rtw<true, SM, D, 8>(t, sm_tab, j, (uint32_t)c, c));
        dft_v4<LOGR, true>(v);
        TS_UNROLL
        for (int u = 0; u < R; u++) st4(tile + b0 + 256 * u, v[brev_c(u, LOGR)]);
    }
}

// ---- tile I/O: one uint4 per (position, quad), straight between global memory and the tile -----------------
// Matrix layouts.  slice == 0: row-major, word (row, col) at row * pitch + col.  slice != 0: "blocked" internal
// layout [col / 8][row][8 columns] with `slice` words between column groups.  profiles/tools/tile_copy.cu: with
// row-major intermediates the strided digit steps 2 MiB between its 32-byte segments (one TLB page each,
// 2.0 TB/s); blocked intermediates step 64 KiB (3.7 TB/s) and make the contiguous digit's tile a single 64 KiB run
// (6.7 TB/s).  Only the first read and the last write of an LDE touch the caller's row-major matrices.
TS_D size_t word_off(size_t row, uint32_t col, uint32_t pitch, size_t slice) {
    return slice ? (size_t)(col >> 3) * slice + row * 8 + (col & 7u) : row * pitch + col;
}

// BREV: tile position q receives source row brev_D(q)
template <int D, bool BREV, int NT, int NQv = dnq(D)>
TS_D void load_tile(uint4 *tile, const uint32_t *src, size_t row_base, size_t row_stride, uint32_t pitch, size_t slice,
                    uint32_t ncols, uint32_t col0, int tid) {
    using G = Geo<D, NQv>;
    constexpr int TOTAL = G::L * G::NQ, U = TOTAL / NT;
    static_assert(TOTAL % NT == 0, "tile");
    uint4 val[U];
    TS_UNROLL
    for (int u = 0; u < U; u++) {
        const int it = u * NT + tid;
        const uint32_t h = it & (G::NQ - 1), q = it / G::NQ;
        const uint32_t r = BREV ? brev_bits(q, D) : q;
        const uint32_t col = col0 + 4 * h;
        val[u] = col < ncols ? nttf::ldg_tile(src + word_off(row_base + (size_t)r * row_stride, col, pitch, slice))
                             : make_uint4(0, 0, 0, 0);
    }
    TS_UNROLL
    for (int u = 0; u < U; u++) {
        const int it = u * NT + tid;
        const uint32_t h = it & (G::NQ - 1), q = it / G::NQ;
        tile[hx<D>(h) ^ sigma(q)] = val[u];
    }
}
template <int D, int NT, int NQv = dnq(D)>
TS_D void store_tile(const uint4 *tile, uint32_t *dst, size_t row_base, size_t row_stride, uint32_t pitch, size_t slice,
                     uint32_t ncols, uint32_t col0, int tid) {
    using G = Geo<D, NQv>;
    constexpr int TOTAL = G::L * G::NQ;
    for (int it = tid; it < TOTAL; it += NT) {
        const uint32_t h = it & (G::NQ - 1), q = it / G::NQ;
        const uint32_t col = col0 + 4 * h;
        if (col < ncols)
            *reinterpret_cast<uint4 *>(dst + word_off(row_base + (size_t)q * row_stride, col, pitch, slice)) =
                tile[hx<D>(h) ^ sigma(q)];
    }
}

// store of the last forward pass of a column-sharded LDE: every row goes to the rank that owns its row range.  Lanes walk
// (quad, position) in order, so with contiguous rows a warp writes one contiguous 512-byte run -- full NVLink packets.
template <int D, int NT, int NQv = dnq(D)>
TS_D void store_tile_peers(const uint4 *tile, const nttf::FastPassParams &p, size_t row_base, size_t row_stride, uint32_t col0,
                           int tid) {
    using G = Geo<D, NQv>;
    constexpr int TOTAL = G::L * G::NQ;
    const size_t mask = ((size_t)1 << p.peer_log_rows) - 1;
    for (int it = tid; it < TOTAL; it += NT) {
        const uint32_t h = it & (G::NQ - 1), q = it / G::NQ;
        const uint32_t col = col0 + 4 * h;
        if (col < p.ncols) {
            const size_t row = row_base + (size_t)q * row_stride;
            uint32_t *base = p.peer[row >> p.peer_log_rows];
            *reinterpret_cast<uint4 *>(base + (row & mask) * p.dst_pitch + col) = tile[hx<D>(h) ^ sigma(q)];
        }
    }
}

// ---- rounds fused with tile I/O ------------------------------------------------------------------------------
// The first DIF round can take its operands straight from global memory and the last one can write its results
// straight back: two of the five shared-memory round trips and two of the five barriers of a tile disappear.
// Item order: NQ adjacent lanes cover one row (a full 32..128-byte segment per instruction); io_blk() permutes the
// radix-4 blocks of a warp so that every quarter-warp still hits eight distinct 16-byte bank groups.
template <int NQ>
TS_D uint32_t io_blk(uint32_t j) {
    if (NQ == 2) return (j & ~7u) | ((j & 3u) << 1) | ((j >> 2) & 1u);
    if (NQ == 4) return (j & ~3u) | ((j & 1u) << 1) | ((j >> 1) & 1u);
    return j;
}
template <int D, bool INV, int NT, int NQv = dnq(D)>
TS_D void dif_r1_ld(uint4 *tile, const uint32_t *src, size_t row_base, size_t row_stride, uint32_t pitch, size_t slice,
                    uint32_t ncols, uint32_t col0, const FastTables &t, int tid) {
    using G = Geo<D, NQv>;
    constexpr int LOGR = G::LOGR1, R = 1 << LOGR;
    TS_UNROLL2
    for (int item = tid; item < 256 * G::NQ; item += NT) {
        const uint32_t h = item & (G::NQ - 1), g = item / G::NQ;
        const uint32_t col = col0 + 4 * h;
        V4 x[R];
        TS_UNROLL
        for (int c = 0; c < R; c++) {
            const uint4 v = col < ncols
                                ? nttf::ldg_tile(src + word_off(row_base + (size_t)(g + 256 * c) * row_stride, col, pitch, slice))
                                : make_uint4(0, 0, 0, 0);
            x[c] = V4{{v.x, v.y, v.z, v.w}};
        }
        dft_v4<LOGR, INV, true>(x);
        TS_UNROLL
        for (int i = 1; i < R; i++) x[i] = vmul(x[i], rtw<INV, false, D, 8>(t, nullptr, g, (uint32_t)brev_c(i, LOGR), i));
        const uint32_t b0 = hx<D>(h) ^ sigma(g);
        TS_UNROLL
        for (int c = 0; c < R; c++) st4(tile + b0 + 256 * c, x[c]);
    }
}
// PSM: the post twiddles of this CTA's `lo` were staged in shared memory as two planes of uint4,
// post[blk] = (w(4 blk), w(4 blk + 1)), post[L/4 + blk] = (w(4 blk + 2), w(4 blk + 3))  (lde_mid: same for all cosets)
template <int D, bool INV, bool POST, int NT, int NQv = dnq(D), bool PSM = false>
TS_D void dif_r4_st(const uint4 *tile, uint32_t *dst, size_t row_base, size_t row_stride, uint32_t pitch, size_t slice,
                    uint32_t ncols, uint32_t col0, const FastTables &t, uint32_t lo, int tw_shift, int tid,
                    const uint4 *post = nullptr) {
    using G = Geo<D, NQv>;
    constexpr int ITEMS = (G::L / 4) * G::NQ;
    for (int item = tid; item < ITEMS; item += NT) {
        const uint32_t h = item & (G::NQ - 1), blk = io_blk<G::NQ>(item / G::NQ);
        const uint32_t base = hx<D>(h) ^ (blk << 2) ^ ((blk >> 1) & 3u) ^ (((blk >> 3) & 1u) << 2);
        uint2 pt[4];
        if (POST && PSM) {
            const uint4 a = post[blk], b = post[G::L / 4 + blk];
            pt[0] = make_uint2(a.x, a.y); pt[1] = make_uint2(a.z, a.w);
            pt[2] = make_uint2(b.x, b.y); pt[3] = make_uint2(b.z, b.w);
        } else if (POST) {
            TS_UNROLL
            for (int c = 0; c < 4; c++) pt[c] = btw<INV>(t, (lo * brev_bits(4 * blk + c, D)) << tw_shift);
        }
        V4 x[4];
        TS_UNROLL
        for (int c = 0; c < 4; c++) x[c] = ld4(tile + (base ^ (uint32_t)c));
        dft_v4<2, INV>(x);
        if (POST) {
            TS_UNROLL
            for (int c = 0; c < 4; c++) x[c] = vmul(x[c], pt[c]);
        }
        const uint32_t col = col0 + 4 * h;
        if (col < ncols) {
            TS_UNROLL
            for (int c = 0; c < 4; c++)
                st4(reinterpret_cast<uint4 *>(dst + word_off(row_base + (size_t)(4 * blk + c) * row_stride, col, pitch, slice)), x[c]);
        }
    }
}

// first DIT round of lde_mid straight from global memory: tile position q holds source row brev_D(q), so the four
// operands of block blk are rows (c << (D-2)) | brev_(D-2)(blk), c = 0..3
template <int D, int NT, int NQv = dnq(D)>
TS_D void dit_r4_ld(uint4 *tile, const uint32_t *src, size_t row_base, uint32_t pitch, size_t slice, uint32_t ncols,
                    uint32_t col0, int tid) {
    using G = Geo<D, NQv>;
    constexpr int ITEMS = (G::L / 4) * G::NQ;
    for (int item = tid; item < ITEMS; item += NT) {
        const uint32_t h = item & (G::NQ - 1), blk = io_blk<G::NQ>(item / G::NQ);
        const uint32_t base = hx<D>(h) ^ (blk << 2) ^ ((blk >> 1) & 3u) ^ (((blk >> 3) & 1u) << 2);
        const uint32_t col = col0 + 4 * h, r0 = brev_bits(blk, D - 2);
        V4 v[4];
        TS_UNROLL
        for (int c = 0; c < 4; c++) {
            const uint4 t = col < ncols ? nttf::ldg_tile(src + word_off(row_base + r0 + ((size_t)c << (D - 2)), col, pitch, slice))
                                        : make_uint4(0, 0, 0, 0);
            v[c] = V4{{t.x, t.y, t.z, t.w}};
        }
        dft_v4<2, true>(v);
        TS_UNROLL
        for (int u = 0; u < 4; u++) st4(tile + (base ^ (uint32_t)u), v[brev_c(u, 2)]);
    }
}

// ---- cp.async (LDGSTS) tile prefetch: 16 bytes straight into the position-major tile ---------------------------
TS_D void cp_async16(uint4 *smem_dst, const uint32_t *gsrc, bool valid) {
#if defined(__CUDA_ARCH__)
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? 16 : 0;  // src-size 0: zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
#else
    *smem_dst = valid ? *reinterpret_cast<const uint4 *>(gsrc) : make_uint4(0, 0, 0, 0);
#endif
}
TS_D void cp_async_commit() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
TS_D void cp_async_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

template <int D, bool BREV, int NT, int NQv = dnq(D)>
TS_D void prefetch_tile(uint4 *tile, const uint32_t *src, size_t row_base, size_t row_stride, uint32_t pitch, size_t slice,
                        uint32_t ncols, uint32_t col0, int tid) {
    using G = Geo<D, NQv>;
    constexpr int TOTAL = G::L * G::NQ;
    TS_UNROLL
    for (int u = 0; u < TOTAL / NT; u++) {
        const int it = u * NT + tid;
        const uint32_t h = it & (G::NQ - 1), q = it / G::NQ;
        const uint32_t r = BREV ? brev_bits(q, D) : q;
        const uint32_t col = col0 + 4 * h;
        const bool ok = col < ncols;
        cp_async16(tile + (hx<D>(h) ^ sigma(q)), src + word_off(row_base + (size_t)r * row_stride, ok ? col : 0, pitch, slice), ok);
    }
    cp_async_commit();
}


// ---- kernels ------------------------------------------------------------------------------------------------
#ifndef TS_PM_PASS_MINBLOCKS
#define TS_PM_PASS_MINBLOCKS 3
#endif
#ifndef TS_PM_PASS_NT
#define TS_PM_PASS_NT 256
#endif
constexpr int PM_PASS_NT = TS_PM_PASS_NT;

template <int D, bool INV>
__global__ void __launch_bounds__(PM_PASS_NT, TS_PM_PASS_MINBLOCKS) ntt_pass_pm_kernel(nttf::FastPassParams p) {
    TS_DYN_SMEM(uint4, tile);
    const int tid = threadIdx.x;
    const uint32_t cs = blockIdx.x % p.n_col_slices, tile_id = blockIdx.x / p.n_col_slices;
    const uint32_t lo = tile_id & ((1u << p.lo_bits) - 1), hi = tile_id >> p.lo_bits;
    const uint32_t col0 = cs << (14 - D);
#ifdef TS_EXP_SAMETILE  // timing experiment only: every CTA works on the same (L2-resident) tile
    const size_t row_base = 0, row_stride = (size_t)1 << p.lo_bits;
#else
    const size_t row_base = ((size_t)hi << (p.lo_bits + D)) + lo, row_stride = (size_t)1 << p.lo_bits;
#endif
#ifdef TS_EXP_NOCOMPUTE  // timing experiment only: the tile I/O without the transform
    load_tile<D, false, PM_PASS_NT>(tile, p.src, row_base, row_stride, p.src_pitch, p.src_slice, p.ncols, col0, tid);
    __syncthreads();
    store_tile<D, PM_PASS_NT>(tile, p.dst, row_base, row_stride, p.dst_pitch, p.dst_slice, p.ncols, col0, tid);
    return;
#endif
#ifndef TS_PM_NO_FUSED_IO
    dif_r1_ld<D, INV, PM_PASS_NT>(tile, p.src, row_base, row_stride, p.src_pitch, p.src_slice, p.ncols, col0, p.t, tid);
    __syncthreads();
    dif_r2<D, INV, false, PM_PASS_NT>(tile, p.t, nullptr, tid);
    __syncthreads();
    dif_r3<D, INV, false, PM_PASS_NT>(tile, p.t, nullptr, tid);
    __syncthreads();
    if (p.peer_log_rows >= 0) {  // scatter to the row owners (ts_coset_lde_batch_scatter); only the contiguous last digit
        dif_r4<D, INV, false, PM_PASS_NT>(tile, p.t, 0, 0, tid);
        __syncthreads();
        store_tile_peers<D, PM_PASS_NT>(tile, p, row_base, row_stride, col0, tid);
    } else if (p.lo_bits > 0)
        dif_r4_st<D, INV, true, PM_PASS_NT>(tile, p.dst, row_base, row_stride, p.dst_pitch, p.dst_slice, p.ncols, col0, p.t, lo,
                                            p.tw_shift, tid);
    else
        dif_r4_st<D, INV, false, PM_PASS_NT>(tile, p.dst, row_base, row_stride, p.dst_pitch, p.dst_slice, p.ncols, col0, p.t, 0, 0,
                                             tid);
#else
    load_tile<D, false, PM_PASS_NT>(tile, p.src, row_base, row_stride, p.src_pitch, p.src_slice, p.ncols, col0, tid);
    __syncthreads();
    dif_r1<D, INV, false, false, PM_PASS_NT>(tile, tile, p.t, nullptr, nullptr, make_uint2(0, 0), tid);
    __syncthreads();
    dif_r2<D, INV, false, PM_PASS_NT>(tile, p.t, nullptr, tid);
    __syncthreads();
    dif_r3<D, INV, false, PM_PASS_NT>(tile, p.t, nullptr, tid);
    __syncthreads();
    if (p.lo_bits > 0) dif_r4<D, INV, true, PM_PASS_NT>(tile, p.t, lo, p.tw_shift, tid);
    else dif_r4<D, INV, false, PM_PASS_NT>(tile, p.t, 0, 0, tid);
    __syncthreads();
    store_tile<D, PM_PASS_NT>(tile, p.dst, row_base, row_stride, p.dst_pitch, p.dst_slice, p.ncols, col0, tid);
#endif
}

#ifndef TS_PM_MID_NT
#define TS_PM_MID_NT 512
#endif
constexpr int PM_MID_NT = TS_PM_MID_NT;

template <int D>
__global__ void __launch_bounds__(PM_MID_NT, 1) lde_mid_pm_kernel(nttf::FastMidParams p) {
    TS_DYN_SMEM(uint4, smem);
    using G = Geo<D>;
    constexpr int R1 = 1 << G::LOGR1;
    uint4 *A = smem, *W = smem + G::UNITS;
    // shared inter-round twiddle tables [register i][group g] (conflict-free LDS, see nttf::stw)
    uint2 *f1 = reinterpret_cast<uint2 *>(W + G::UNITS);  // fwd R1: w_L^(g brev(i)),    R1 x 256
    uint2 *f2 = f1 + R1 * 256;                            // fwd R2: w_256^(g brev3(i)), 8 x 32
    uint2 *f3 = f2 + 256;                                 // fwd R3: w_32^(g brev3(i)),  8 x 4
    uint2 *i1 = f3 + 32;                                  // inv (DIT) tables: w^-(c j)
    uint2 *i2 = i1 + R1 * 256;
    uint2 *i3 = i2 + 256;
    uint4 *post = reinterpret_cast<uint4 *>(i3 + 32);    // post twiddles w_n^(Kc brev(p)), two planes (dif_r4_st)
    const int tid = threadIdx.x;
    const int m = D + p.klo_bits;
    const FastTables &t = p.t;
    for (int e = tid; e < R1 * 256; e += PM_MID_NT) {
        const uint32_t i = e >> 8, g = e & 255;
        f1[e] = __ldg(t.tw_small + ((g * brev_bits(i, G::LOGR1)) << (t.small_log - D)));
        i1[e] = __ldg(t.tw_small_inv + ((g * i) << (t.small_log - D)));
    }
    for (int e = tid; e < 256; e += PM_MID_NT) {
        const uint32_t i = e >> 5, g = e & 31;
        f2[e] = __ldg(t.tw_small + ((g * brev_bits(i, 3)) << (t.small_log - 8)));
        i2[e] = __ldg(t.tw_small_inv + ((g * i) << (t.small_log - 8)));
    }
    for (int e = tid; e < 32; e += PM_MID_NT) {
        const uint32_t i = e >> 2, g = e & 3;
        f3[e] = __ldg(t.tw_small + ((g * brev_bits(i, 3)) << (t.small_log - 5)));
        i3[e] = __ldg(t.tw_small_inv + ((g * i) << (t.small_log - 5)));
    }
    // One CTA per SM walks a CONTIGUOUS range of tiles (tile = Kc * n_col_slices + column slice): the round tables
    // above are filled once per CTA instead of once per tile (profiles/r01: 4 % of the stall samples sat on those fills,
    // and with one resident CTA every tile boundary is an idle SM), and the post-twiddle table is regathered only when
    // Kc changes, i.e. once per n_col_slices tiles.
    const uint32_t n_tiles = p.n_tiles ? p.n_tiles : gridDim.x;
    const uint32_t per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
    const uint32_t tile_begin = blockIdx.x * per_cta, tile_end = tile_begin + per_cta < n_tiles ? tile_begin + per_cta : n_tiles;
    uint32_t post_Kc = 0xffffffffu;
    bool prefetched = false;
    for (uint32_t tile_id = tile_begin; tile_id < tile_end; tile_id++) {
#ifdef TS_EXP_SAMETILE
    const uint32_t cs = tile_id % p.n_col_slices, Kc = 0;
#else
    const uint32_t cs = tile_id % p.n_col_slices, Kc = tile_id / p.n_col_slices;
#endif
    const uint32_t col0 = cs << (14 - D);
    __syncthreads();  // the previous tile's last round has finished with W and the post table; the tables above are visible
    // the inter-digit twiddles depend on (Kc, position) only: gathered once per Kc (2^D scattered 8-byte reads of the
    // big table) instead of once per coset, and under the latency of the tile load that follows
    if (p.klo_bits > 0 && Kc != post_Kc) {
        for (int e = tid; e < G::L / 2; e += PM_MID_NT) {
            const uint32_t blk = e & (G::L / 4 - 1), c0 = 2 * (e / (G::L / 4));
            const uint2 w0 = btw<false>(t, (Kc * brev_bits(4 * blk + c0, D)) << p.tw_shift);
            const uint2 w1 = btw<false>(t, (Kc * brev_bits(4 * blk + c0 + 1, D)) << p.tw_shift);
            post[e] = make_uint4(w0.x, w0.y, w1.x, w1.y);
        }
        post_Kc = Kc;
    }
    // inverse sub-transform on the lowest digit: rows brev(Kc)*L + x, loaded into bit-reversed positions
#if defined(TS_PM_NO_FUSED_IO) || defined(TS_EXP_NOCOMPUTE)
    load_tile<D, true, PM_MID_NT>(A, p.src, (size_t)brev_bits(Kc, p.klo_bits) << D, 1, p.src_pitch, p.src_slice, p.ncols, col0, tid);
    __syncthreads();
#endif
#ifdef TS_EXP_NOCOMPUTE
    for (uint32_t j = 0; j < (1u << p.b); j++)
        store_tile<D, PM_MID_NT>(A, p.dst, ((size_t)brev_bits(j, p.b) << m) + Kc, (size_t)1 << p.klo_bits, p.dst_pitch,
                                 p.dst_slice, p.ncols, col0, tid);
    continue;
#endif
#ifdef TS_PM_NO_FUSED_IO
    dit_r4<D, PM_MID_NT>(A, tid);
#else
    if (prefetched) {  // this tile's rows were copied into A (cp.async) behind the previous tile's last coset
        cp_async_wait<0>();
        __syncthreads();
        dit_r4<D, PM_MID_NT>(A, tid);
    } else {
        dit_r4_ld<D, PM_MID_NT>(A, p.src, (size_t)brev_bits(Kc, p.klo_bits) << D, p.src_pitch, p.src_slice, p.ncols, col0, tid);
    }
#endif
    __syncthreads();
    dit_r3<D, true, PM_MID_NT>(A, t, i3, tid);
    __syncthreads();
    dit_r2<D, true, PM_MID_NT>(A, t, i2, tid);
    __syncthreads();
    dit_r1<D, true, PM_MID_NT>(A, t, i1, tid);
    for (uint32_t j = 0; j < (1u << p.b); j++) {
        const uint2 lw = __ldg(p.lane_tab + ((size_t)j << p.klo_bits) + Kc);
        __syncthreads();  // orders the last DIT round / the previous coset's store before W is rewritten
        dif_r1<D, false, true, true, PM_MID_NT>(A, W, t, f1, p.pre_tab + ((size_t)j << D), lw, tid);
        __syncthreads();
#if !defined(TS_PM_NO_FUSED_IO) && defined(TS_PM_MID_PREFETCH)  // measured slower (19.6 vs 19.1 ms): opt-in build flag
        if (j + 1 == (1u << p.b) && tile_id + 1 < tile_end) {
            // A is dead from here on: start the next tile's (bit-reversed) row copy so that it lands during R2..R4
            const uint32_t nt = tile_id + 1;
#ifdef TS_EXP_SAMETILE
            const uint32_t ncs = nt % p.n_col_slices, nKc = 0;
#else
            const uint32_t ncs = nt % p.n_col_slices, nKc = nt / p.n_col_slices;
#endif
            prefetch_tile<D, true, PM_MID_NT>(A, p.src, (size_t)brev_bits(nKc, p.klo_bits) << D, 1, p.src_pitch, p.src_slice, p.ncols,
                                              ncs << (14 - D), tid);
            prefetched = true;
        }
#endif
        dif_r2<D, false, true, PM_MID_NT>(W, t, f2, tid);
        __syncthreads();
        dif_r3<D, false, true, PM_MID_NT>(W, t, f3, tid);
        __syncthreads();
        const size_t row_base = ((size_t)brev_bits(j, p.b) << m) + Kc, row_stride = (size_t)1 << p.klo_bits;
#ifdef TS_PM_NO_FUSED_IO
        if (p.klo_bits > 0) dif_r4<D, false, true, PM_MID_NT>(W, t, Kc, p.tw_shift, tid);
        else dif_r4<D, false, false, PM_MID_NT>(W, t, 0, 0, tid);
        __syncthreads();
        store_tile<D, PM_MID_NT>(W, p.dst, row_base, row_stride, p.dst_pitch, p.dst_slice, p.ncols, col0, tid);
#else
        if (p.klo_bits > 0)
            dif_r4_st<D, false, true, PM_MID_NT, dnq(D), true>(W, p.dst, row_base, row_stride, p.dst_pitch, p.dst_slice, p.ncols,
                                                               col0, t, Kc, p.tw_shift, tid, post);
        else
            dif_r4_st<D, false, false, PM_MID_NT>(W, p.dst, row_base, row_stride, p.dst_pitch, p.dst_slice, p.ncols, col0, t, 0, 0,
                                                  tid);
#endif
    }
    }  // tile loop
}


}  // namespace nttp

// =====================================================================================================================
// Persistent, software-pipelined variants.  profiles/tools/tile_copy.cu measured the memory-system ceiling of
// the 2048 x 32-byte tile pattern at 4.0 TB/s (contiguous rows) and 2.0 TB/s (rows 2 MiB apart): the digit passes
// are within 2x of their memory floor, so the remaining lever is to run memory and arithmetic CONCURRENTLY inside
// one CTA instead of relying on 2-3 co-resident CTAs drifting out of phase: one CTA per SM loops over tiles,
// cp.async (LDGSTS, 16 B, straight into the position-major tile) prefetches tile i+1 while tile i is transformed,
// and the STG.128 of tile i drain while tile i+1 is transformed.
namespace nttp {

#ifndef TS_PM2_NT
#define TS_PM2_NT 512
#endif
constexpr int PM2_NT = TS_PM2_NT;

struct PersistPassParams {
    nttf::FastPassParams p;
    uint32_t n_tiles;  // (2^(hi_bits + lo_bits)) * n_col_slices
};

template <int D, bool INV>
__global__ void __launch_bounds__(PM2_NT, 1) ntt_pass_pm2_kernel(PersistPassParams pp) {
    TS_DYN_SMEM(uint4, smem);
    using G = Geo<D>;
    const nttf::FastPassParams &p = pp.p;
    const int tid = threadIdx.x;
    auto buf = [&](int i) { return smem + (i ? G::UNITS : 0); };
    auto geom = [&](uint32_t t, uint32_t &lo, uint32_t &col0, size_t &row_base) {
        const uint32_t cs = t % p.n_col_slices, tile_id = t / p.n_col_slices;
        lo = tile_id & ((1u << p.lo_bits) - 1);
        const uint32_t hi = tile_id >> p.lo_bits;
        col0 = cs << (14 - D);
        row_base = ((size_t)hi << (p.lo_bits + D)) + lo;
    };
    const size_t row_stride = (size_t)1 << p.lo_bits;
    uint32_t t = blockIdx.x, lo, col0;
    size_t row_base;
    int cur = 0;
    if (t < pp.n_tiles) {
        geom(t, lo, col0, row_base);
        prefetch_tile<D, false, PM2_NT>(buf(0), p.src, row_base, row_stride, p.src_pitch, p.src_slice, p.ncols, col0, tid);
    }
    for (; t < pp.n_tiles; t += gridDim.x) {
        geom(t, lo, col0, row_base);
        const uint32_t tn = t + gridDim.x;
        if (tn < pp.n_tiles) {
            uint32_t lo2, col2;
            size_t rb2;
            geom(tn, lo2, col2, rb2);
            prefetch_tile<D, false, PM2_NT>(buf(cur ^ 1), p.src, rb2, row_stride, p.src_pitch, p.src_slice, p.ncols, col2, tid);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        uint4 *tile = buf(cur);
        dif_r1<D, INV, false, false, PM2_NT>(tile, tile, p.t, nullptr, nullptr, make_uint2(0, 0), tid);
        __syncthreads();
        dif_r2<D, INV, false, PM2_NT>(tile, p.t, nullptr, tid);
        __syncthreads();
        dif_r3<D, INV, false, PM2_NT>(tile, p.t, nullptr, tid);
        __syncthreads();
        if (p.lo_bits > 0) dif_r4<D, INV, true, PM2_NT>(tile, p.t, lo, p.tw_shift, tid);
        else dif_r4<D, INV, false, PM2_NT>(tile, p.t, 0, 0, tid);
        __syncthreads();
        store_tile<D, PM2_NT>(tile, p.dst, row_base, row_stride, p.dst_pitch, p.dst_slice, p.ncols, col0, tid);
        __syncthreads();  // buf[cur] is the prefetch target of the next iteration
        cur ^= 1;
    }
}

struct PersistMidParams {
    nttf::FastMidParams p;
    uint32_t n_tiles;  // 2^klo_bits * n_col_slices
};

template <int D>
__global__ void __launch_bounds__(PM2_NT, 1) lde_mid_pm2_kernel(PersistMidParams pp) {
    TS_DYN_SMEM(uint4, smem);
    using G = Geo<D>;
    constexpr int R1 = 1 << G::LOGR1;
    const nttf::FastMidParams &p = pp.p;
    uint4 *W = smem;                                   // work tile of the current coset
    auto Abuf = [&](int i) { return smem + (i ? 2 * G::UNITS : G::UNITS); };  // coefficient tile / prefetch target
    uint2 *f1 = reinterpret_cast<uint2 *>(smem + 3 * G::UNITS);  // fwd R1 [i][g], R1 x 256
    uint2 *f2 = f1 + R1 * 256;                                   // fwd R2, 8 x 32
    uint2 *f3 = f2 + 256;                                        // fwd R3, 8 x 4
    uint2 *i2 = f3 + 32;                                         // inverse (DIT) R2 / R3; DIT R1 reads global
    uint2 *i3 = i2 + 256;
    const int tid = threadIdx.x;
    const FastTables &t = p.t;
    for (int e = tid; e < R1 * 256; e += PM2_NT) {
        const uint32_t i = e >> 8, g = e & 255;
        f1[e] = __ldg(t.tw_small + ((g * brev_bits(i, G::LOGR1)) << (t.small_log - D)));
    }
    for (int e = tid; e < 256; e += PM2_NT) {
        const uint32_t i = e >> 5, g = e & 31;
        f2[e] = __ldg(t.tw_small + ((g * brev_bits(i, 3)) << (t.small_log - 8)));
        i2[e] = __ldg(t.tw_small_inv + ((g * i) << (t.small_log - 8)));
    }
    for (int e = tid; e < 32; e += PM2_NT) {
        const uint32_t i = e >> 2, g = e & 3;
        f3[e] = __ldg(t.tw_small + ((g * brev_bits(i, 3)) << (t.small_log - 5)));
        i3[e] = __ldg(t.tw_small_inv + ((g * i) << (t.small_log - 5)));
    }
    const int m = D + p.klo_bits;
    uint32_t tl = blockIdx.x;
    int cur = 0;
    if (tl < pp.n_tiles) {
        const uint32_t cs = tl % p.n_col_slices, Kc = tl / p.n_col_slices;
        prefetch_tile<D, true, PM2_NT>(Abuf(0), p.src, (size_t)brev_bits(Kc, p.klo_bits) << D, 1, p.src_pitch, p.src_slice,
                                       p.ncols, cs << (14 - D), tid);
    }
    for (; tl < pp.n_tiles; tl += gridDim.x) {
        const uint32_t cs = tl % p.n_col_slices, Kc = tl / p.n_col_slices, col0 = cs << (14 - D);
        const uint32_t tn = tl + gridDim.x;
        if (tn < pp.n_tiles) {
            const uint32_t cs2 = tn % p.n_col_slices, K2 = tn / p.n_col_slices;
            prefetch_tile<D, true, PM2_NT>(Abuf(cur ^ 1), p.src, (size_t)brev_bits(K2, p.klo_bits) << D, 1, p.src_pitch,
                                           p.src_slice, p.ncols, cs2 << (14 - D), tid);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        uint4 *A = Abuf(cur);
        dit_r4<D, PM2_NT>(A, tid);
        __syncthreads();
        dit_r3<D, true, PM2_NT>(A, t, i3, tid);
        __syncthreads();
        dit_r2<D, true, PM2_NT>(A, t, i2, tid);
        __syncthreads();
        dit_r1<D, false, PM2_NT>(A, t, nullptr, tid);
        for (uint32_t j = 0; j < (1u << p.b); j++) {
            const uint2 lw = __ldg(p.lane_tab + ((size_t)j << p.klo_bits) + Kc);
            __syncthreads();
            dif_r1<D, false, true, true, PM2_NT>(A, W, t, f1, p.pre_tab + ((size_t)j << D), lw, tid);
            __syncthreads();
            dif_r2<D, false, true, PM2_NT>(W, t, f2, tid);
            __syncthreads();
            dif_r3<D, false, true, PM2_NT>(W, t, f3, tid);
            __syncthreads();
            if (p.klo_bits > 0) dif_r4<D, false, true, PM2_NT>(W, t, Kc, p.tw_shift, tid);
            else dif_r4<D, false, false, PM2_NT>(W, t, 0, 0, tid);
            __syncthreads();
            store_tile<D, PM2_NT>(W, p.dst, ((size_t)brev_bits(j, p.b) << m) + Kc, (size_t)1 << p.klo_bits, p.dst_pitch,
                                  p.dst_slice, p.ncols, col0, tid);
        }
        __syncthreads();  // Abuf[cur] becomes the prefetch target two iterations from now; W is rewritten next
        cur ^= 1;
    }
}

}  // namespace nttp
