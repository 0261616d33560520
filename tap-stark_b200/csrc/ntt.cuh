// ntt.cuh -- batched BabyBear NTT digit passes and the fused LDE middle kernel (sm_100a).
//
// What it replaces: p3_dft::TwoAdicSubgroupDft::coset_lde_batch as called by TwoAdicFriPcs::commit
// (fri/src/two_adic_pcs.rs:237-240) followed by bit_reverse_rows().to_row_major_matrix().
//
// Formulation (DESIGN.md "LDE"): a size-2^m transform is plain in-place radix-2 decimation-in-frequency,
// with its m stages grouped into "digits" of d <= 11 bits.  One kernel launch = one digit for every column:
// a CTA stages a tile [2^d positions] x [K lanes] in shared memory (lanes = adjacent rows x adjacent
// columns, so every global access is a contiguous >= 32-byte segment of a row-major matrix), runs the 2^d
// point sub-transform as radix-16 register rounds, applies the inter-digit twiddle on the way out.
// Natural-order input therefore yields BIT-REVERSED output with no permutation pass - exactly the
// "committed order" the PCS wants.
//
// The LDE is: inverse passes over the top digits (ntt_pass_kernel<INV>), then lde_mid_kernel which finishes
// the inverse transform on the lowest digit, and - with the coefficients still on chip - multiplies by the
// coset powers and runs the top digit of the forward transform once per coset (2^added_bits of them),
// then forward passes over the remaining digits (ntt_pass_kernel<FWD>) in place on the output.
#pragma once
#include "field.cuh"

namespace ntt {

constexpr int SMALL_LOG = 12;   // tw_small holds w_{4096}^e
constexpr int MAX_DIGIT = 11;   // tile = 2^11 positions x 8 lanes x 4 B = 64 KiB

// XOR swizzle of tile positions: makes both "32 consecutive positions" and "32 threads x stride 16"
// (the stride-1 radix-16 round) bank-conflict free.
TS_D uint32_t swz(uint32_t p) { return p ^ ((p >> 4) & 15u) ^ (((p >> 8) & 1u) << 4); }

__host__ __device__ constexpr int brev_c(int i, int bits) {
    int r = 0;
    for (int k = 0; k < bits; k++) r |= ((i >> k) & 1) << (bits - 1 - k);
    return r;
}
TS_D uint32_t brev_bits(uint32_t x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

// In-register radix-2^LOGR DFT, decimation in frequency, compile-time twiddles.
// Natural-order input, output register i holds frequency brev(i).
template <int LOGR, bool INV>
TS_D void dft_regs(uint32_t (&x)[1 << LOGR]) {
    constexpr bb::InnerTw<LOGR, INV> T{};
    constexpr int R = 1 << LOGR;
    TS_UNROLL
    for (int s = 0; s < LOGR; s++) {
        const int half = R >> (s + 1);
        TS_UNROLL
        for (int blk = 0; blk < R; blk += 2 * half) {
            TS_UNROLL
            for (int j = 0; j < half; j++) {
                const uint32_t a = x[blk + j], b = x[blk + j + half];
                x[blk + j] = bb::add(a, b);
                const int e = j << s;  // w_{2 half}^j = w_R^(j 2^s)
                if (e == 0) {
                    x[blk + j + half] = bb::sub(a, b);
                } else {
                    x[blk + j + half] = bb::shoup(a - b + bb::P, T.w[e], T.wp[e]);
                }
            }
        }
    }
}

template <bool INV>
TS_D uint2 small_tw(const uint2 *__restrict__ tw_small, uint32_t e) {
    if (INV) e = ((1u << SMALL_LOG) - e) & ((1u << SMALL_LOG) - 1);
    return __ldg(tw_small + e);
}

// Optional prescale applied while loading the first forward round of lde_mid_kernel.
struct Prescale {
    const uint2 *pos_tab;   // indexed by tile position (k_hi)
    const uint2 *lane_tab;  // indexed by lane
};

// One DIF round: sub-blocks of size 2^logS are split into 2^LOGR sub-blocks of size M = 2^(logS-LOGR).
// src may differ from dst (first forward round of lde_mid reads the coefficient tile, writes the work tile).
template <int LOGR, bool INV, bool PRE>
TS_D void round_dif(const uint32_t *src, uint32_t *dst, int LS, int logK, int logL, int logS,
                    const uint2 *__restrict__ tw_small, Prescale pre, int tid, int nt) {
    constexpr int R = 1 << LOGR;
    const int logM = logS - LOGR;
    const int logG = logL - LOGR;  // groups per lane
    const int total = 1 << (logK + logG);
    for (int it = tid; it < total; it += nt) {
        const int lane = it >> logG, grp = it & ((1 << logG) - 1);
        const uint32_t blk = grp >> logM, g = grp & ((1u << logM) - 1);
        const uint32_t base = (blk << logS) + g;
        const uint32_t *s = src + lane * LS;
        uint32_t *t = dst + lane * LS;
        uint32_t x[R];
        TS_UNROLL
        for (int c = 0; c < R; c++) x[c] = s[swz(base + ((uint32_t)c << logM))];
        if (PRE) {
            const uint2 lw = pre.lane_tab[lane];
            TS_UNROLL
            for (int c = 0; c < R; c++) {
                const uint2 pw = __ldg(pre.pos_tab + base + ((uint32_t)c << logM));
                x[c] = bb::shoup(bb::shoup_lazy(x[c], pw.x, pw.y), lw.x, lw.y);
            }
        }
        dft_regs<LOGR, INV>(x);
        if (logM > 0) {
            TS_UNROLL
            for (int i = 1; i < R; i++) {
                const uint32_t u = brev_c(i, LOGR);
                const uint2 w = small_tw<INV>(tw_small, (g * u) << (SMALL_LOG - logS));
                x[i] = bb::shoup(x[i], w.x, w.y);
            }
        }
        TS_UNROLL
        for (int i = 0; i < R; i++) t[swz(base + ((uint32_t)i << logM))] = x[i];
    }
}

// One DIT round: 2^LOGR adjacent blocks of size B = 2^logB (bit-reversed digit order) -> one block of B*R.
template <int LOGR, bool INV>
TS_D void round_dit(uint32_t *tile, int LS, int logK, int logL, int logB, const uint2 *__restrict__ tw_small,
                    int tid, int nt) {
    constexpr int R = 1 << LOGR;
    const int logG = logL - LOGR;
    const int total = 1 << (logK + logG);
    for (int it = tid; it < total; it += nt) {
        const int lane = it >> logG, grp = it & ((1 << logG) - 1);
        const uint32_t blk = grp >> logB, j = grp & ((1u << logB) - 1);
        const uint32_t base = (blk << (logB + LOGR)) + j;
        uint32_t *t = tile + lane * LS;
        uint32_t x[R], v[R];
        TS_UNROLL
        for (int i = 0; i < R; i++) x[i] = t[swz(base + ((uint32_t)i << logB))];
        if (logB > 0) {
            TS_UNROLL
            for (int i = 1; i < R; i++) {
                const uint32_t c = brev_c(i, LOGR);
                const uint2 w = small_tw<INV>(tw_small, (c * j) << (SMALL_LOG - (logB + LOGR)));
                x[i] = bb::shoup(x[i], w.x, w.y);
            }
        }
        TS_UNROLL
        for (int c = 0; c < R; c++) v[c] = x[brev_c(c, LOGR)];
        dft_regs<LOGR, INV>(v);
        TS_UNROLL
        for (int u = 0; u < R; u++) t[swz(base + ((uint32_t)u << logB))] = v[brev_c(u, LOGR)];
    }
}

template <bool INV, bool PRE>
TS_D void dif_round_dispatch(int logr, const uint32_t *src, uint32_t *dst, int LS, int logK, int logL, int logS,
                             const uint2 *tw_small, Prescale pre, int tid, int nt) {
    switch (logr) {
        case 1: round_dif<1, INV, PRE>(src, dst, LS, logK, logL, logS, tw_small, pre, tid, nt); break;
        case 2: round_dif<2, INV, PRE>(src, dst, LS, logK, logL, logS, tw_small, pre, tid, nt); break;
        case 3: round_dif<3, INV, PRE>(src, dst, LS, logK, logL, logS, tw_small, pre, tid, nt); break;
        default: round_dif<4, INV, PRE>(src, dst, LS, logK, logL, logS, tw_small, pre, tid, nt); break;
    }
}
TS_HD int first_radix(int d) { return d <= 4 ? d : ((d & 3) ? (d & 3) : 4); }

// All DIF rounds of a 2^d-point sub-transform: [remainder radix, 16, 16, ...]; the last (stride-1) round is
// radix 16 whenever d >= 4.  Ends with a barrier.  If PRE, the first round reads `first_src` with prescale.
template <bool INV, bool PRE>
TS_D void dif_rounds(const uint32_t *first_src, uint32_t *tile, int LS, int logK, int d, const uint2 *tw_small,
                     Prescale pre, int tid, int nt) {
    int logS = d;
    const int r0 = first_radix(d);
    if (d > 0) {
        dif_round_dispatch<INV, PRE>(r0, first_src, tile, LS, logK, d, logS, tw_small, pre, tid, nt);
        __syncthreads();
        logS -= r0;
    }
    while (logS > 0) {
        round_dif<4, INV, false>(tile, tile, LS, logK, d, logS, tw_small, pre, tid, nt);
        __syncthreads();
        logS -= 4;
    }
}
// All DIT rounds (mirror schedule): [16, 16, ..., remainder radix].  Ends with a barrier.
template <bool INV>
TS_D void dit_rounds(uint32_t *tile, int LS, int logK, int d, const uint2 *tw_small, int tid, int nt) {
    const int r0 = first_radix(d);
    int logB = 0;
    while (logB + r0 < d) {
        round_dit<4, INV>(tile, LS, logK, d, logB, tw_small, tid, nt);
        __syncthreads();
        logB += 4;
    }
    if (d > 0) {
        switch (r0) {
            case 1: round_dit<1, INV>(tile, LS, logK, d, logB, tw_small, tid, nt); break;
            case 2: round_dit<2, INV>(tile, LS, logK, d, logB, tw_small, tid, nt); break;
            case 3: round_dit<3, INV>(tile, LS, logK, d, logB, tw_small, tid, nt); break;
            default: round_dit<4, INV>(tile, LS, logK, d, logB, tw_small, tid, nt); break;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Generic digit pass.  Row index of (tile position q, lane row bt):
//     row = (hi << (lo_bits + d)) | (q << lo_bits) | lo
// lanes batch over adjacent `lo` (batch_lo) or adjacent `hi`.
struct PassParams {
    const uint32_t *src;
    uint32_t *dst;
    uint32_t width;
    int d, lo_bits, hi_bits;
    int logBt, logCt, pad;
    int batch_lo;
    uint32_t n_col_slices;
    int post_tw;       // multiply position q by w_{2^(lo_bits+d)}^(+-lo*brev_d(q))
    int tw_shift;      // big_log - (lo_bits + d)
    int big_log;
    const uint2 *tw_small, *tw_big;
    int has_scale;
    uint2 scale;
};

template <bool INV>
__global__ void __launch_bounds__(256) ntt_pass_kernel(PassParams p) {
    TS_DYN_SMEM(uint32_t, tile);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int L = 1 << p.d, LS = L + p.pad;
    const int logK = p.logBt + p.logCt, Kmask = (1 << logK) - 1;
    const uint32_t cs = blockIdx.x % p.n_col_slices, tile_id = blockIdx.x / p.n_col_slices;
    uint32_t hi0, lo0;
    if (p.batch_lo) {
        const int lb = p.lo_bits - p.logBt;  // lo blocks = 2^lb
        hi0 = tile_id >> lb;
        lo0 = (tile_id & ((1u << lb) - 1)) << p.logBt;
    } else {
        hi0 = tile_id << p.logBt;
        lo0 = 0;
    }
    const uint32_t col0 = cs << p.logCt;
    const int total = L << logK;
    for (int it = tid; it < total; it += nt) {
        const int lane = it & Kmask;
        const uint32_t q = it >> logK;
        const uint32_t bt = lane >> p.logCt, ct = lane & ((1 << p.logCt) - 1);
        const uint32_t hi = hi0 + (p.batch_lo ? 0 : bt), lo = lo0 + (p.batch_lo ? bt : 0);
        const bool valid = (col0 + ct < p.width) && (hi < (1u << p.hi_bits));
        const size_t row = ((size_t)hi << (p.lo_bits + p.d)) + ((size_t)q << p.lo_bits) + lo;
        tile[lane * LS + swz(q)] = valid ? p.src[row * p.width + col0 + ct] : 0u;
    }
    __syncthreads();
    dif_rounds<INV, false>(tile, tile, LS, logK, p.d, p.tw_small, Prescale{nullptr, nullptr}, tid, nt);
    for (int it = tid; it < total; it += nt) {
        const int lane = it & Kmask;
        const uint32_t q = it >> logK;
        const uint32_t bt = lane >> p.logCt, ct = lane & ((1 << p.logCt) - 1);
        const uint32_t hi = hi0 + (p.batch_lo ? 0 : bt), lo = lo0 + (p.batch_lo ? bt : 0);
        const bool valid = (col0 + ct < p.width) && (hi < (1u << p.hi_bits));
        if (!valid) continue;
        uint32_t v = tile[lane * LS + swz(q)];
        if (p.post_tw) {
            uint32_t e = (lo * brev_bits(q, p.d)) << p.tw_shift;
            if (INV) e = ((1u << p.big_log) - e) & ((1u << p.big_log) - 1);
            const uint2 w = __ldg(p.tw_big + e);
            v = bb::shoup(v, w.x, w.y);
        }
        if (p.has_scale) v = bb::shoup(v, p.scale.x, p.scale.y);
        const size_t row = ((size_t)hi << (p.lo_bits + p.d)) + ((size_t)q << p.lo_bits) + lo;
        p.dst[row * p.width + col0 + ct] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// LDE middle kernel.  Source rows: chunk brev(K) holds, for coefficient residue K (mod n_lo), the
// partially inverse-transformed values y[x], x in [0,L).  Per tile (lanes = adjacent K x adjacent columns):
//   A <- y in bit-reversed position;  DIT inverse rounds  ->  A[k_hi] = n * a[K + n_lo k_hi]
//   for every coset j:  W[k_hi] = A[k_hi] * (sigma_j^n_lo)^k_hi * (sigma_j^K / n);  DIF forward rounds;
//                       position q -> row (brev_b(j) << m) | (q << klo_bits) | K, times w_n^(K brev_d(q)).
struct MidParams {
    const uint32_t *src;
    uint32_t *dst;
    uint32_t width;
    int d, klo_bits, b;
    int logBt, logCt, pad;
    uint32_t n_col_slices;
    int tw_shift, big_log;  // big_log - m
    const uint2 *tw_small, *tw_big;
    const uint2 *pre_tab;   // [2^b][2^d]
    const uint2 *lane_tab;  // [2^b][2^klo_bits]
};

__global__ void __launch_bounds__(512) lde_mid_kernel(MidParams p) {
    TS_DYN_SMEM(uint32_t, smem);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int L = 1 << p.d, LS = L + p.pad;
    const int logK = p.logBt + p.logCt, K = 1 << logK, Kmask = K - 1;
    uint32_t *A = smem, *W = smem + K * LS;
    uint2 *lane_w = reinterpret_cast<uint2 *>(W + K * LS);  // K entries: per-lane scalar of the current coset
    const uint32_t cs = blockIdx.x % p.n_col_slices, tile_id = blockIdx.x / p.n_col_slices;
    const uint32_t K0 = tile_id << p.logBt, col0 = cs << p.logCt;
    const int m = p.d + p.klo_bits;
    const int total = L << logK;
    for (int it = tid; it < total; it += nt) {
        const int lane = it & Kmask;
        const uint32_t pos = it >> logK;
        const uint32_t bt = lane >> p.logCt, ct = lane & ((1 << p.logCt) - 1);
        const uint32_t Kc = K0 + bt;
        const bool valid = (col0 + ct < p.width) && (Kc < (1u << p.klo_bits));
        const size_t row = ((size_t)brev_bits(Kc, p.klo_bits) << p.d) + brev_bits(pos, p.d);
        A[lane * LS + swz(pos)] = valid ? p.src[row * p.width + col0 + ct] : 0u;
    }
    __syncthreads();
    dit_rounds<true>(A, LS, logK, p.d, p.tw_small, tid, nt);
    for (uint32_t j = 0; j < (1u << p.b); j++) {
        if (tid < K) {
            const uint32_t Kc = K0 + ((uint32_t)tid >> p.logCt);
            lane_w[tid] = (Kc < (1u << p.klo_bits)) ? p.lane_tab[((size_t)j << p.klo_bits) + Kc] : make_uint2(0, 0);
        }
        __syncthreads();
        Prescale pre{p.pre_tab + ((size_t)j << p.d), lane_w};
        if (p.d > 0) {
            dif_rounds<false, true>(A, W, LS, logK, p.d, p.tw_small, pre, tid, nt);
        } else {  // L == 1: only the lane scalar
            if (tid < K) W[tid * LS] = bb::shoup(A[tid * LS], lane_w[tid].x, lane_w[tid].y);
            __syncthreads();
        }
        const size_t coset_base = (size_t)brev_bits(j, p.b) << m;
        for (int it = tid; it < total; it += nt) {
            const int lane = it & Kmask;
            const uint32_t q = it >> logK;
            const uint32_t bt = lane >> p.logCt, ct = lane & ((1 << p.logCt) - 1);
            const uint32_t Kc = K0 + bt;
            const bool valid = (col0 + ct < p.width) && (Kc < (1u << p.klo_bits));
            if (!valid) continue;
            uint32_t v = W[lane * LS + swz(q)];
            if (p.klo_bits > 0) {
                const uint32_t e = (Kc * brev_bits(q, p.d)) << p.tw_shift;
                const uint2 w = __ldg(p.tw_big + e);
                v = bb::shoup(v, w.x, w.y);
            }
            const size_t row = coset_base + ((size_t)q << p.klo_bits) + Kc;
            p.dst[row * p.width + col0 + ct] = v;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Table generators (device side, once per (size, shift); cached by the context).

// tw[e] = w_{2^log}^e as Shoup pair, e in [0, 2^log).  root_pows[k] = w^(2^k) in Montgomery form.
struct RootPows {
    uint32_t v[28];
};
__global__ void gen_twiddles_kernel(uint2 *out, int log, RootPows rp) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (1u << log)) return;
    uint32_t acc = bb::MONTY_ONE;
    for (int k = 0; k < log; k++)
        if ((e >> k) & 1) acc = bb::mmul(acc, rp.v[k]);
    const uint32_t w = bb::from_monty(acc);
    out[e] = make_uint2(w, bb::shoup_prime(w));
}

TS_D uint32_t mpow(uint32_t base_monty, uint32_t e) {
    uint32_t acc = bb::MONTY_ONE;
    while (e) {
        if (e & 1) acc = bb::mmul(acc, base_monty);
        base_monty = bb::mmul(base_monty, base_monty);
        e >>= 1;
    }
    return acc;
}
// pre_tab[j][k_hi] = (sigma_j^n_lo)^k_hi ; lane_tab[j][K] = sigma_j^K * n_inv ; sigma_j = shift * w_N^j
__global__ void gen_coset_tables_kernel(uint2 *pre_tab, uint2 *lane_tab, int d, int klo_bits, int b,
                                        uint32_t shift_monty, uint32_t wN_monty, uint32_t ninv_monty) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t per = (1u << d) + (1u << klo_bits);
    if (idx >= (per << b)) return;
    const uint32_t j = idx / per, r = idx % per;
    const uint32_t sigma = bb::mmul(shift_monty, mpow(wN_monty, j));
    uint32_t val;
    if (r < (1u << d)) {
        val = mpow(mpow(sigma, 1u << klo_bits), r);
        const uint32_t w = bb::from_monty(val);
        pre_tab[((size_t)j << d) + r] = make_uint2(w, bb::shoup_prime(w));
    } else {
        const uint32_t Kc = r - (1u << d);
        val = bb::mmul(mpow(sigma, Kc), ninv_monty);
        const uint32_t w = bb::from_monty(val);
        lane_tab[((size_t)j << klo_bits) + Kc] = make_uint2(w, bb::shoup_prime(w));
    }
}

// out row r = in row brev(r)  (p3-matrix bit_reverse_rows().to_row_major_matrix()); one warp-ish per row
__global__ void bitrev_rows_kernel(const uint32_t *in, uint32_t *out, int log_h, uint32_t width) {
    const size_t total = ((size_t)1 << log_h) * width;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(i / width), c = (uint32_t)(i % width);
        out[i] = in[(size_t)brev_bits(r, log_h) * width + c];
    }
}
// row k *= shift^k   (coefficient-side coset shift of the plain coset_dft_batch; not on the hot path)
__global__ void scale_rows_pow_kernel(uint32_t *data, int log_h, uint32_t width, uint32_t shift_monty) {
    const size_t total = ((size_t)1 << log_h) * width;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(i / width);
        data[i] = bb::mmul(data[i], mpow(shift_monty, r));
    }
}
// n == 1: every output row equals the single input row
__global__ void broadcast_row_kernel(const uint32_t *in, uint32_t *out, size_t rows, uint32_t width) {
    const size_t total = rows * width;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        out[i] = in[i % width];
}
// synthetic trace: element (r, c) of a rows x total_width matrix = SplitMix64((seed << 40) + r * total_width + c)
// mod p; this kernel fills the column window [col0, col0 + width) of it, so every rank of a column-sharded run holds
// its slice of the SAME matrix.  Same function as oracle.splitmix_matrix.
__global__ void fill_splitmix_kernel(uint32_t *out, size_t rows, uint32_t width, uint64_t seed, uint32_t col0,
                                     uint32_t total_width, int monty) {
    const size_t total = rows * width;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / width;
        const uint32_t c = (uint32_t)(i % width);
        uint64_t z = (seed << 40) + (uint64_t)r * total_width + col0 + c + 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        const uint32_t v = (uint32_t)(z % bb::P);
        out[i] = monty ? bb::to_monty(v) : v;
    }
}
__global__ void monty_convert_kernel(uint32_t *data, size_t n, int to) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        data[i] = to ? bb::to_monty(data[i]) : bb::from_monty(data[i]);
}

}  // namespace ntt
