// ntt_fast.cuh -- fast-path digit kernels for wide matrices (width % 4 == 0, width >= lanes) and digit sizes
// D in {9,10,11}: the shapes of BASELINE configs 2-4.  Same mathematics as ntt.cuh (which stays the generic
// path); what changes is everything around the butterflies (profiles/r01: v1 spent 30 instr per element-round
// and stalled 2/3 of the time on twiddle loads):
//   * D, the lane count K = 2^(14-D) and the round schedule are compile-time: rounds M8 (radix 2^(D-8), stride
//     256), M4 (radix 16, stride 16), M0 (radix 16, stride 1);
//   * tile word index = (lane << D) | (swz(p) ^ ((lane & 7) << 2)), no padding: every element address of a
//     round is  base ^ const  or  base + const  (0-1 ALU op per element instead of ~8), still conflict free;
//   * a thread owns the SAME butterfly group in every lane, so its inter-round, coset and inter-digit twiddles
//     are loaded ONCE per tile into registers and reused across the K lanes (no loads in the inner loops);
//   * tile load/store moves uint4 (4 adjacent columns) per thread.
#pragma once
#include "ntt.cuh"

namespace nttf {

using ntt::brev_bits;
using ntt::brev_c;
using ntt::dft_regs;
using ntt::SMALL_LOG;

struct FastTables {
    const uint2 *tw_small;      // w_{2^small_log}^e   (global: 4096 entries; lde_mid: a shared-memory copy of 2^D)
    const uint2 *tw_small_inv;  // w_{2^small_log}^-e
    const uint2 *tw_big;        // w_{2^big_log}^e
    int big_log;
    int small_log;
    // product-indexed inter-round tables in global memory (ntt_pm.cuh rtw): rt[dir] holds, back to back,
    // R1 tables for D = 9, 10, 11 ([row][g], 256 groups), the R2 table (8 x 32) and the R3 table (8 x 4);
    // row r, group g = root^(g * brev(r)), root = w (dir 0) or w^-1 (dir 1) of the round's sub-transform size.
    const uint2 *rt[2];
};
__host__ __device__ constexpr int RT_R1_OFF(int D) { return D == 9 ? 0 : D == 10 ? 2 * 256 : 6 * 256; }
constexpr int RT_R2_OFF = 14 * 256, RT_R3_OFF = 15 * 256, RT_ENTRIES = 15 * 256 + 32;
// dst[(row << logg) + g] = src[(g * brev(row)) << (small_log - logs)] for a round of 2^logs points in 2^logg groups
__global__ void fill_round_table_kernel(uint2 *dst, const uint2 *src, int logs, int logg, int small_log) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (1u << logs)) return;
    const uint32_t row = e >> logg, g = e & ((1u << logg) - 1);
    dst[e] = src[(g * ntt::brev_bits(row, logs - logg)) << (small_log - logs)];
}

// Inter-round twiddle of register/digit `i` for butterfly group `g` in a round whose sub-transform has 2^LOGS
// points and 2^LOGG groups: w_{2^LOGS}^(+-g*u).  SM = false: read-only global load from the 2^small_log table.
// SM = true (lde_mid): t.tw_small[_inv] point at shared-memory tables laid out [i][g] by fill_round_tables(), so
// the 32 threads of a warp (consecutive g) read consecutive entries -- conflict free -- whereas indexing a
// plain power table by g*u would serialise up to 32-way.
template <bool INV, bool SM, int LOGS, int LOGG>
TS_D uint2 stw(const FastTables &t, const uint2 *sm_tab, uint32_t g, uint32_t u, uint32_t i) {
    if (SM) return sm_tab[(i << LOGG) + g];
    return __ldg((INV ? t.tw_small_inv : t.tw_small) + ((g * u) << (t.small_log - LOGS)));
}
template <bool INV>
TS_D uint2 btw(const FastTables &t, uint32_t e) {
    if (INV) e = ((1u << t.big_log) - e) & ((1u << t.big_log) - 1);
    return __ldg(t.tw_big + e);
}

// 16-byte global load of a tile segment.  A tile row segment is only 4*K bytes (32 B for D = 11) of a row that the
// neighbouring column-slice CTAs read next, so L2 is asked to fetch the whole 128-byte line: DRAM then sees one
// 128 B burst instead of four scattered 32 B sector reads.
TS_D uint4 ldg_tile(const uint32_t *p) {
#if defined(__CUDA_ARCH__) && !defined(TS_NO_L2_128B)
    uint4 v;
    asm volatile("ld.global.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#else
    return *reinterpret_cast<const uint4 *>(p);
#endif
}

template <int D>
TS_D uint32_t phys(uint32_t lane, uint32_t p) {
    return (lane << D) | (ntt::swz(p) ^ ((lane & 7u) << 2));
}

// ---- DIF rounds ----------------------------------------------------------------------------------------
// M8: radix R = 2^(D-8), elements p = g + 256 c, g = tid & 255.  Optional prescale (lde_mid first round):
// element c of every lane is multiplied by pos[c] and by the lane scalar.
template <int D, bool INV, bool PRE, int NT, bool SM>
TS_D void dif_m8(const uint32_t *src, uint32_t *dst, const FastTables &t, const uint2 *sm_tab, const uint2 *pw,
                 const uint2 *lane_w, int tid) {
    constexpr int LOGR = D - 8, R = 1 << LOGR, K = 1 << (14 - D);
    const uint32_t g = tid & 255, sg = g ^ (g >> 4);
    uint2 tw[R];
    TS_UNROLL
    for (int i = 1; i < R; i++) tw[i] = stw<INV, SM, D, 8>(t, sm_tab, g, (uint32_t)brev_c(i, LOGR), i);
    TS_UNROLL2
    for (int lane = tid >> 8; lane < K; lane += NT / 256) {
        const uint32_t b0 = ((uint32_t)lane << D) | (sg ^ ((lane & 7u) << 2)), b1 = b0 ^ 16u;
        uint32_t x[R];
        TS_UNROLL
        for (int c = 0; c < R; c++) x[c] = src[((c & 1) ? b1 : b0) + 256 * c];
        if (PRE) {
            const uint2 lw = lane_w[lane];
            TS_UNROLL
            for (int c = 0; c < R; c++) x[c] = bb::shoup(bb::shoup_lazy(x[c], pw[c].x, pw[c].y), lw.x, lw.y);
        }
        dft_regs<LOGR, INV>(x);
        TS_UNROLL
        for (int i = 1; i < R; i++) x[i] = bb::shoup(x[i], tw[i].x, tw[i].y);
        TS_UNROLL
        for (int c = 0; c < R; c++) dst[((c & 1) ? b1 : b0) + 256 * c] = x[c];
    }
}
// M4: radix 16 inside blocks of 256: p = 256 blk + g + 16 c
template <int D, bool INV, int NT, bool SM>
TS_D void dif_m4(uint32_t *tile, const FastTables &t, const uint2 *sm_tab, int tid) {
    constexpr int G = 1 << (D - 4), K = 1 << (14 - D);
    const uint32_t grp = tid & (G - 1), blk = grp >> 4, g = grp & 15;
    const uint32_t pb = (blk << 8) | ((blk & 1u) << 4) | g;
    uint2 tw[16];
    TS_UNROLL
    for (int i = 1; i < 16; i++) tw[i] = stw<INV, SM, 8, 4>(t, sm_tab, g, (uint32_t)brev_c(i, 4), i);
    TS_UNROLL2
    for (int lane = tid >> (D - 4); lane < K; lane += (NT >> (D - 4)) > 0 ? (NT >> (D - 4)) : 1) {
        const uint32_t base = ((uint32_t)lane << D) | (pb ^ ((lane & 7u) << 2));
        uint32_t x[16];
        TS_UNROLL
        for (int c = 0; c < 16; c++) x[c] = tile[base ^ (17u * c)];
        dft_regs<4, INV>(x);
        TS_UNROLL
        for (int i = 1; i < 16; i++) x[i] = bb::shoup(x[i], tw[i].x, tw[i].y);
        TS_UNROLL
        for (int c = 0; c < 16; c++) tile[base ^ (17u * c)] = x[c];
    }
}
// M0: radix 16 on 16 consecutive positions p = 16 g' + i.  POST: multiply position q by the inter-digit
// twiddle w_{n'}^(+-lo*brev_D(q)) (index pre-shifted by the caller: e = (lo * brev) << tw_shift).
// the 16 inter-digit twiddles of this thread's M0 group (positions 16 gq + i)
template <int D, bool INV>
TS_D void load_post_tw(uint2 (&pt)[16], const FastTables &t, uint32_t lo, int tw_shift, int tid) {
    const uint32_t gq = tid & ((1 << (D - 4)) - 1);
    TS_UNROLL
    for (int i = 0; i < 16; i++) pt[i] = btw<INV>(t, (lo * brev_bits(16u * gq + i, D)) << tw_shift);
}
template <int D, bool INV, bool POST, int NT>
TS_D void dif_m0(uint32_t *tile, const uint2 (&pt)[16], int tid) {
    constexpr int G = 1 << (D - 4), K = 1 << (14 - D);
    const uint32_t gq = tid & (G - 1);
    const uint32_t pb = (gq << 4) ^ (gq & 15u) ^ (((gq >> 4) & 1u) << 4);
    TS_UNROLL2
    for (int lane = tid >> (D - 4); lane < K; lane += (NT >> (D - 4)) > 0 ? (NT >> (D - 4)) : 1) {
        const uint32_t base = ((uint32_t)lane << D) | (pb ^ ((lane & 7u) << 2));
        uint32_t x[16];
        TS_UNROLL
        for (int c = 0; c < 16; c++) x[c] = tile[base ^ (uint32_t)c];
        dft_regs<4, INV>(x);
        if (POST) {
            TS_UNROLL
            for (int i = 0; i < 16; i++) x[i] = bb::shoup(x[i], pt[i].x, pt[i].y);
        }
        TS_UNROLL
        for (int c = 0; c < 16; c++) tile[base ^ (uint32_t)c] = x[c];
    }
}

// ---- DIT rounds (inverse sub-transform of lde_mid: bit-reversed positions in, natural out) ----------------
template <int D, int NT>
TS_D void dit_m0(uint32_t *tile, int tid) {
    constexpr int G = 1 << (D - 4), K = 1 << (14 - D);
    const uint32_t gq = tid & (G - 1);
    const uint32_t pb = (gq << 4) ^ (gq & 15u) ^ (((gq >> 4) & 1u) << 4);
    TS_UNROLL2
    for (int lane = tid >> (D - 4); lane < K; lane += (NT >> (D - 4)) > 0 ? (NT >> (D - 4)) : 1) {
        const uint32_t base = ((uint32_t)lane << D) | (pb ^ ((lane & 7u) << 2));
        uint32_t v[16];
        TS_UNROLL
        for (int c = 0; c < 16; c++) v[c] = tile[base ^ (uint32_t)brev_c(c, 4)];  // block i holds digit brev(i)
        dft_regs<4, true>(v);
        TS_UNROLL
        for (int u = 0; u < 16; u++) tile[base ^ (uint32_t)u] = v[brev_c(u, 4)];
    }
}
template <int D, int NT, bool SM>
TS_D void dit_m4(uint32_t *tile, const FastTables &t, const uint2 *sm_tab, int tid) {
    constexpr int G = 1 << (D - 4), K = 1 << (14 - D);
    const uint32_t grp = tid & (G - 1), blk = grp >> 4, j = grp & 15;
    const uint32_t pb = (blk << 8) | ((blk & 1u) << 4) | j;
    uint2 tw[16];  // for digit c: w_256^-(c j)
    TS_UNROLL
    for (int c = 1; c < 16; c++) tw[c] = stw<true, SM, 8, 4>(t, sm_tab, j, (uint32_t)c, c);
    TS_UNROLL2
    for (int lane = tid >> (D - 4); lane < K; lane += (NT >> (D - 4)) > 0 ? (NT >> (D - 4)) : 1) {
        const uint32_t base = ((uint32_t)lane << D) | (pb ^ ((lane & 7u) << 2));
        uint32_t v[16];
        TS_UNROLL
        for (int c = 0; c < 16; c++) v[c] = tile[base ^ (17u * brev_c(c, 4))];
        TS_UNROLL
        for (int c = 1; c < 16; c++) v[c] = bb::shoup(v[c], tw[c].x, tw[c].y);
        dft_regs<4, true>(v);
        TS_UNROLL
        for (int u = 0; u < 16; u++) tile[base ^ (17u * u)] = v[brev_c(u, 4)];
    }
}
template <int D, int NT, bool SM>
TS_D void dit_m8(uint32_t *tile, const FastTables &t, const uint2 *sm_tab, int tid) {
    constexpr int LOGR = D - 8, R = 1 << LOGR, K = 1 << (14 - D);
    const uint32_t j = tid & 255, sj = j ^ (j >> 4);
    uint2 tw[R];
    TS_UNROLL
    for (int c = 1; c < R; c++) tw[c] = stw<true, SM, D, 8>(t, sm_tab, j, (uint32_t)c, c);
    TS_UNROLL2
    for (int lane = tid >> 8; lane < K; lane += NT / 256) {
        const uint32_t b0 = ((uint32_t)lane << D) | (sj ^ ((lane & 7u) << 2)), b1 = b0 ^ 16u;
        uint32_t v[R];
        TS_UNROLL
        for (int c = 0; c < R; c++) {
            const int i = brev_c(c, LOGR);
            v[c] = tile[((i & 1) ? b1 : b0) + 256 * i];
        }
        TS_UNROLL
        for (int c = 1; c < R; c++) v[c] = bb::shoup(v[c], tw[c].x, tw[c].y);
        dft_regs<LOGR, true>(v);
        TS_UNROLL
        for (int u = 0; u < R; u++) tile[((u & 1) ? b1 : b0) + 256 * u] = v[brev_c(u, LOGR)];
    }
}

// ---- tile I/O: uint4 = 4 adjacent columns of one row ------------------------------------------------------
// BREV: tile position q holds source row index brev_D(q) (the DIT input order of lde_mid)
template <int D, bool BREV, int NT>
TS_D void load_tile(uint32_t *tile, const uint32_t *src, size_t row_base, size_t row_stride, uint32_t pitch,
                    uint32_t ncols, uint32_t col0, int tid) {
    constexpr int LOGV = 12 - D, NV = 1 << LOGV;  // vectors per position (K/4)
    constexpr int TOTAL = (1 << D) * NV;          // = 4096 uint4 per tile
    constexpr int U = TOTAL / NT;                 // every load of the tile is issued before the first use:
    static_assert(TOTAL % NT == 0, "tile");       // ONE memory round trip per tile (profiles/r01: 4 batches of 4
    uint4 val[U];                                 // loads made this phase 35 % of the kernel)
    TS_UNROLL
    for (int u = 0; u < U; u++) {
        const int it = u * NT + tid;
        const uint32_t v = it & (NV - 1), q = it >> LOGV;
        const uint32_t r = BREV ? brev_bits(q, D) : q;
        const uint32_t col = col0 + 4 * v;
        val[u] = col < ncols ? ldg_tile(src + (row_base + (size_t)r * row_stride) * pitch + col) : make_uint4(0, 0, 0, 0);
    }
    TS_UNROLL
    for (int u = 0; u < U; u++) {
        const int it = u * NT + tid;
        const uint32_t v = it & (NV - 1), q = it >> LOGV;
        tile[phys<D>(4 * v + 0, q)] = val[u].x;
        tile[phys<D>(4 * v + 1, q)] = val[u].y;
        tile[phys<D>(4 * v + 2, q)] = val[u].z;
        tile[phys<D>(4 * v + 3, q)] = val[u].w;
    }
}
template <int D, int NT>
TS_D void store_tile(const uint32_t *tile, uint32_t *dst, size_t row_base, size_t row_stride, uint32_t pitch,
                     uint32_t ncols, uint32_t col0, int tid) {
    constexpr int LOGV = 12 - D, NV = 1 << LOGV;
    constexpr int TOTAL = (1 << D) * NV;
    for (int it = tid; it < TOTAL; it += NT) {
        const uint32_t v = it & (NV - 1), q = it >> LOGV;
        const uint32_t col = col0 + 4 * v;
        if (col < ncols) {
            uint4 o;
            o.x = tile[phys<D>(4 * v + 0, q)];
            o.y = tile[phys<D>(4 * v + 1, q)];
            o.z = tile[phys<D>(4 * v + 2, q)];
            o.w = tile[phys<D>(4 * v + 3, q)];
            *reinterpret_cast<uint4 *>(dst + (row_base + (size_t)q * row_stride) * pitch + col) = o;
        }
    }
}

// ---- kernels ------------------------------------------------------------------------------------------------
// src/dst point at the first column of the window being transformed (a column chunk of a wider matrix is a
// window: pitch = row stride in u32, ncols = window width), so chunks can be pipelined against H2D copies.
struct FastPassParams {
    const uint32_t *src;
    uint32_t *dst;
    uint32_t src_pitch, dst_pitch, ncols;
    size_t src_slice = 0, dst_slice = 0;  // != 0: "blocked" layout [col / 8][row][8] with this slice stride (ntt_pm only)
    int lo_bits, hi_bits;
    uint32_t n_col_slices;
    int tw_shift;  // big_log - (lo_bits + D)
    FastTables t;
    // peer_log_rows >= 0 (last forward pass of a column-sharded LDE, ntt_pm only): output row r is written to
    // peer[r >> peer_log_rows] at row r & (2^peer_log_rows - 1), dst_pitch words per row -- the owner of that row range;
    // remote entries are peer-mapped device memory (NVLink stores), so the re-shard needs no separate all-to-all
    uint32_t *peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int peer_log_rows = -1;
};
#ifndef TS_PASS_MINBLOCKS
#define TS_PASS_MINBLOCKS 2
#endif
#ifndef TS_MID_NT
#define TS_MID_NT 512
#endif
constexpr int PASS_NT = 256;

// In-place digit pass, one tile = (hi, lo) x column slice: row = (hi << (lo_bits+D)) | (q << lo_bits) | lo
template <int D, bool INV>
__global__ void __launch_bounds__(PASS_NT, TS_PASS_MINBLOCKS) ntt_pass_fast_kernel(FastPassParams p) {
    TS_DYN_SMEM(uint32_t, tile);
    const int tid = threadIdx.x;
    const uint32_t cs = blockIdx.x % p.n_col_slices, tile_id = blockIdx.x / p.n_col_slices;
    const uint32_t lo = tile_id & ((1u << p.lo_bits) - 1), hi = tile_id >> p.lo_bits;
    const uint32_t col0 = cs << (14 - D);
    const size_t row_base = ((size_t)hi << (p.lo_bits + D)) + lo, row_stride = (size_t)1 << p.lo_bits;
    load_tile<D, false, PASS_NT>(tile, p.src, row_base, row_stride, p.src_pitch, p.ncols, col0, tid);
    __syncthreads();
    dif_m8<D, INV, false, PASS_NT, false>(tile, tile, p.t, nullptr, nullptr, nullptr, tid);
    __syncthreads();
    dif_m4<D, INV, PASS_NT, false>(tile, p.t, nullptr, tid);
    uint2 pt[16];  // 3 CTAs per SM overlap this fetch; holding it across rounds would spill at 80 registers
    if (p.lo_bits > 0) load_post_tw<D, INV>(pt, p.t, lo, p.tw_shift, tid);
    __syncthreads();
    if (p.lo_bits > 0) dif_m0<D, INV, true, PASS_NT>(tile, pt, tid);
    else dif_m0<D, INV, false, PASS_NT>(tile, pt, tid);
    __syncthreads();
    store_tile<D, PASS_NT>(tile, p.dst, row_base, row_stride, p.dst_pitch, p.ncols, col0, tid);
}

struct FastMidParams {
    const uint32_t *src;
    uint32_t *dst;
    uint32_t src_pitch, dst_pitch, ncols;
    size_t src_slice = 0, dst_slice = 0;
    int klo_bits, b;
    uint32_t n_col_slices;
    int tw_shift;  // big_log - m
    FastTables t;
    const uint2 *pre_tab;   // [2^b][2^D]
    const uint2 *lane_tab;  // [2^b][2^klo_bits]
    uint32_t n_tiles = 0;   // ntt_pm lde_mid: tiles in total; a CTA walks a contiguous range of them (0: one tile per CTA)
};
constexpr int MID_NT = TS_MID_NT;

template <int D>
__global__ void __launch_bounds__(MID_NT, 1) lde_mid_fast_kernel(FastMidParams p) {
    TS_DYN_SMEM(uint32_t, smem);
    constexpr int L = 1 << D, K = 1 << (14 - D), R8 = 1 << (D - 8);
    uint32_t *A = smem, *W = smem + K * L;
    // shared-memory inter-round twiddle tables, laid out [register i][group g] (see stw):
    uint2 *f8 = reinterpret_cast<uint2 *>(W + K * L);  // forward M8: w_L^(g brev(i)),   R8 x 256
    uint2 *f4 = f8 + R8 * 256;                         // forward M4: w_256^(g brev4(i)), 16 x 16
    uint2 *i8 = f4 + 256;                              // inverse (DIT) M8: w_L^-(c j),  R8 x 256
    uint2 *i4 = i8 + R8 * 256;                         // inverse (DIT) M4: w_256^-(c j), 16 x 16
    uint2 *lane_w = i4 + 256;
    const int tid = threadIdx.x;
    const uint32_t cs = blockIdx.x % p.n_col_slices, Kc = blockIdx.x / p.n_col_slices;
    const uint32_t col0 = cs << (14 - D);
    const int m = D + p.klo_bits;
    // inter-digit twiddles w_n^(Kc * brev(q)) do not depend on the coset: fetched once, in flight during the
    // whole inverse sub-transform
    uint2 pt[16];
    if (p.klo_bits > 0) load_post_tw<D, false>(pt, p.t, Kc, p.tw_shift, tid);
    for (int e = tid; e < R8 * 256; e += MID_NT) {
        const uint32_t i = e >> 8, g = e & 255;
        f8[e] = __ldg(p.t.tw_small + ((g * (uint32_t)brev_bits(i, D - 8)) << (p.t.small_log - D)));
        i8[e] = __ldg(p.t.tw_small_inv + ((g * i) << (p.t.small_log - D)));
    }
    for (int e = tid; e < 256; e += MID_NT) {
        const uint32_t i = e >> 4, g = e & 15;
        f4[e] = __ldg(p.t.tw_small + ((g * (uint32_t)brev_bits(i, 4)) << (p.t.small_log - 8)));
        i4[e] = __ldg(p.t.tw_small_inv + ((g * i) << (p.t.small_log - 8)));
    }
    const FastTables &ts = p.t;
    // inverse sub-transform on the lowest digit: rows brev(Kc)*L + x, loaded into bit-reversed positions
    load_tile<D, true, MID_NT>(A, p.src, (size_t)brev_bits(Kc, p.klo_bits) << D, 1, p.src_pitch, p.ncols, col0, tid);
    __syncthreads();
    dit_m0<D, MID_NT>(A, tid);
    __syncthreads();
    dit_m4<D, MID_NT, true>(A, ts, i4, tid);
    __syncthreads();
    // coset prescale factors of this thread's first-round elements (k_hi = g + 256 c), fetched one coset ahead
    uint2 pw[R8], pw_next[R8];
    {
        const uint32_t g = tid & 255;
        TS_UNROLL
        for (int c = 0; c < R8; c++) pw_next[c] = __ldg(p.pre_tab + g + 256 * c);
    }
    dit_m8<D, MID_NT, true>(A, ts, i8, tid);
    for (uint32_t j = 0; j < (1u << p.b); j++) {
        if (tid < K) lane_w[tid] = p.lane_tab[((size_t)j << p.klo_bits) + Kc];
        TS_UNROLL
        for (int c = 0; c < R8; c++) pw[c] = pw_next[c];
        if (j + 1 < (1u << p.b)) {
            const uint32_t g = tid & 255;
            TS_UNROLL
            for (int c = 0; c < R8; c++) pw_next[c] = __ldg(p.pre_tab + ((size_t)(j + 1) << D) + g + 256 * c);
        }
        __syncthreads();  // also orders the last DIT round / the previous coset's store before W is rewritten
        dif_m8<D, false, true, MID_NT, true>(A, W, ts, f8, pw, lane_w, tid);
        __syncthreads();
        dif_m4<D, false, MID_NT, true>(W, ts, f4, tid);
        __syncthreads();
        if (p.klo_bits > 0) dif_m0<D, false, true, MID_NT>(W, pt, tid);
        else dif_m0<D, false, false, MID_NT>(W, pt, tid);
        __syncthreads();
        store_tile<D, MID_NT>(W, p.dst, ((size_t)brev_bits(j, p.b) << m) + Kc, (size_t)1 << p.klo_bits, p.dst_pitch,
                              p.ncols, col0, tid);
    }
}

}  // namespace nttf
