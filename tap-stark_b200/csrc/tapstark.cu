// tapstark.cu -- context, planning and the extern "C" surface of libtapstark_b200.so (see include/tapstark.h).
//
// Device work is enqueued on one stream per context; nothing here computes field/hash results on the host
// except the Fiat-Shamir challenger and the verifier-side Merkle check, which are host code in the
// reference too.
#include "../../include/tapstark.h"

#include <algorithm>
#include <array>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <atomic>
#include <tuple>
#include <functional>
#include <vector>

#include "fold.cuh"
#include "fri_tail.cuh"
#include "hash.cuh"
#include "host_side.h"
#include "ntt.cuh"
#include "ntt_fast.cuh"
#include "ntt_pm.cuh"
#include "ntt_v4.cuh"
#include "open.cuh"
#include "quotient.cuh"
#include "sha256.cuh"

// ------------------------------------------------------------------------------------------------ structs
struct ts_matrix {
    ts_ctx *ctx;
    uint32_t *d;
    size_t rows, width;
    bool owned;
};

struct ts_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int num_sms = 148;
    uint2 *tw_small = nullptr;
    uint2 *tw_small_inv = nullptr;
    uint2 *round_tab[2] = {nullptr, nullptr};  // nttf::FastTables::rt
    uint2 *tw_big = nullptr;
    int big_log = 0;
    uint32_t *fold_tlo = nullptr;
    std::map<std::tuple<int, int, int, uint32_t>, std::pair<uint2 *, uint2 *>> coset_tabs;
    // ntt_v4.cuh tables: inverse inter-digit twiddles [lo][p] per (m, top digit); per (m, b, low digit, shift) the forward
    // top-digit twiddles with the coset scalars folded in [j][Kc][p] and the coset powers [j][g][c]
    std::map<std::pair<int, int>, uint2 *> v4_post1;
    struct V4Mid { uint2 *post3 = nullptr, *pre4 = nullptr; uint64_t stamp = 0; };
    std::map<std::tuple<int, int, int, uint32_t>, V4Mid> v4_mid;
    uint64_t v4_clock = 0;
    uint32_t *scratch = nullptr;
    size_t scratch_words = 0;
    // stream-ordered caching allocator: freed blocks are reused by later work on the same stream without a
    // device synchronisation (cudaFree would serialise every step of the pipeline)
    // H2D pipeline of the host-buffer entry points: column chunks are copied on copy_stream while the previous
    // chunk's LDE runs on `stream`
    cudaStream_t copy_stream = nullptr;
    // ts_copy_async / ts_copy2d_async: copy-engine transfers (peer buffers over NVLink, strided host windows) on their own
    // stream, ordered after the work queued on `stream` at the call; ts_copy_join makes `stream` wait for them
    static constexpr int XFER_LANES = 8;  // independent queues (e.g. lane 0 host staging, lanes 1..7 one per peer)
    cudaStream_t xfer_stream[XFER_LANES] = {};
    cudaEvent_t ev_xfer_in[XFER_LANES] = {}, ev_xfer_out[XFER_LANES] = {};
    // second compute stream: the leaf hash of column chunk k runs here beside the LDE of chunk k+1 (lde_hash_overlapped)
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_side = nullptr;
    // set around calls whose kernels share the GPU with other work (the sharded prover's LDE runs beside the NCCL
    // all-to-all of the previous column chunk): persistent one-CTA-per-SM launches are avoided there
    bool shares_gpu = false;
    std::vector<uint32_t> tail_final;  // final FRI layer of the last fri_tail call (Montgomery), host side
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    bool ev_free_used[2] = {false, false};
    uint32_t *stage[2] = {nullptr, nullptr};
    size_t stage_words = 0;
    // pageable host sources (stage_pageable_window): worker threads gather row windows into page-locked bounce slots
    // and queue 1-D copies on their own streams; two slots per worker, reused across calls
    static constexpr int BOUNCE_WORKERS = 16;
    static constexpr size_t BOUNCE_SLOT_BYTES = (size_t)4 << 20;
    uint8_t *bounce = nullptr;
    cudaStream_t bounce_stream[BOUNCE_WORKERS] = {};
    cudaEvent_t bounce_ev[BOUNCE_WORKERS][2] = {}, bounce_done[BOUNCE_WORKERS] = {};
    bool bounce_ev_used[BOUNCE_WORKERS][2] = {};
    std::multimap<size_t, void *> pool_free;
    std::map<void *, size_t> pool_sizes;
    // stats
    bool profiling = false;
    double ms[TS_K_COUNT] = {0};
    uint64_t launches[TS_K_COUNT] = {0};
    uint64_t total_launches = 0;
    struct Pending { cudaEvent_t a, b; int kind; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> ev_pool;
};

struct ts_tree {
    ts_ctx *ctx;
    std::vector<ts_matrix *> mats;
    bool own_mats;
    int layout;
    std::vector<size_t> order;
    size_t hmax;
    unsigned lmax;
    uint32_t *digests;               // all layers, 8 words per node
    std::vector<size_t> layer_off;   // in nodes
};

#define TS_FAIL(ctx, code, msg)      \
    do {                             \
        (ctx)->err = (msg);          \
        return (code);               \
    } while (0)
#define TS_CUDA(ctx, call)                                                                     \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                   \
            return TS_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)
#define TS_TRY(expr)                  \
    do {                              \
        int rc_ = (expr);             \
        if (rc_ != TS_OK) return rc_; \
    } while (0)

namespace {
struct KScope {  // counts a kernel launch and, when profiling, brackets it with events on the ctx stream
    ts_ctx *c;
    int kind;
    cudaEvent_t a = nullptr, b = nullptr;
    KScope(ts_ctx *c_, int kind_) : c(c_), kind(kind_) {
        c->launches[kind]++;
        c->total_launches++;
        if (c->profiling) {
            a = get();
            b = get();
            cudaEventRecord(a, c->stream);
        }
    }
    cudaEvent_t get() {
        if (!c->ev_pool.empty()) {
            cudaEvent_t e = c->ev_pool.back();
            c->ev_pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    ~KScope() {
        if (c->profiling) {
            cudaEventRecord(b, c->stream);
            c->pending.push_back({a, b, kind});
        }
    }
};

int check_launch(ts_ctx *c, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        c->err = std::string(what) + ": " + cudaGetErrorString(e);
        return TS_ERR_CUDA;
    }
    return TS_OK;
}

void pool_trim(ts_ctx *c) {
    if (c->pool_free.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (auto &kv : c->pool_free) {
        cudaFree(kv.second);
        c->pool_sizes.erase(kv.second);
    }
    c->pool_free.clear();
}
cudaError_t pool_alloc(ts_ctx *c, void **p, size_t bytes) {
#ifdef TS_EMULATE  // exact sizes: under the AddressSanitizer build of the emulator every byte past the end is a red zone
    const size_t sz = std::max<size_t>(bytes, 1);
#else
    const size_t sz = (std::max<size_t>(bytes, 1) + 511) & ~(size_t)511;
#endif
    // best fit within 2x (and at most 64 MiB of slack): a prover with varying shapes reuses its cached blocks instead of holding
    // one per distinct size; the block keeps its real size in pool_sizes
    auto it = c->pool_free.lower_bound(sz);
    if (it != c->pool_free.end() && (it->first == sz || (it->first <= 2 * sz && it->first - sz <= ((size_t)64 << 20)))) {
        *p = it->second;
        c->pool_free.erase(it);
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(p, sz);
    if (e != cudaSuccess) {  // give cached blocks back to the driver and retry once
        cudaGetLastError();
        pool_trim(c);
        e = cudaMalloc(p, sz);
    }
    if (e == cudaSuccess) c->pool_sizes[*p] = sz;
    return e;
}
void pool_release(ts_ctx *c, void *p) {
    if (!p) return;
    auto it = c->pool_sizes.find(p);
    if (it == c->pool_sizes.end()) {
        cudaFree(p);
        return;
    }
    c->pool_free.insert({it->second, p});
}

int log2_strict(size_t x) {
    int l = 0;
    while (((size_t)1 << l) < x) l++;
    return (((size_t)1 << l) == x) ? l : -1;
}
int log2_ceil(size_t x) {
    int l = 0;
    while (((size_t)1 << l) < x) l++;
    return l;
}
uint32_t h_to_monty(uint32_t x) { return (uint32_t)((((uint64_t)x) << 32) % bb::P); }
uint32_t h_from_monty(uint32_t x) { return bb::cmul(x, bb::cinv(bb::MONTY_ONE)); }

int ensure_scratch(ts_ctx *c, size_t words) {
    if (c->scratch_words >= words) return TS_OK;
    if (c->scratch) cudaFree(c->scratch);
    c->scratch = nullptr;
    c->scratch_words = 0;
    TS_CUDA(c, cudaMalloc((void **)&c->scratch, words * 4));
    c->scratch_words = words;
    return TS_OK;
}

int gen_twiddles(ts_ctx *c, uint2 *out, int log, bool inverse = false) {
    ntt::RootPows rp;
    uint32_t r = bb::two_adic_generator(log);
    if (inverse) r = bb::cinv(r);
    for (int k = 0; k < 28; k++) {
        rp.v[k] = h_to_monty(r);
        r = bb::cmul(r, r);
    }
    const uint32_t n = 1u << log;
    KScope ks(c, TS_K_MISC);
    auto kfn = ntt::gen_twiddles_kernel;
    TS_LAUNCH(kfn, (n + 255) / 256, 256, 0, c->stream, out, log, rp);
    return check_launch(c, "gen_twiddles_kernel");
}

int ensure_big(ts_ctx *c, int log) {
    if (c->tw_big && c->big_log >= log) return TS_OK;
    if (c->tw_big) {
        TS_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->tw_big);
        c->tw_big = nullptr;
    }
    TS_CUDA(c, cudaMalloc((void **)&c->tw_big, sizeof(uint2) << log));
    c->big_log = log;
    return gen_twiddles(c, c->tw_big, log);
}

// lanes of a tile: K = Bt x Ct, all powers of two
struct Lanes {
    int logBt, logCt, pad;
};
Lanes choose_lanes(int d, size_t w, int batch_bits, size_t max_tile_words) {
    int logKmax = 0;
    while (logKmax < 6 && (((size_t)2 << logKmax) << d) <= max_tile_words) logKmax++;
    Lanes l;
    l.logCt = std::min(log2_ceil(w), logKmax);
    l.logBt = std::min(logKmax - l.logCt, batch_bits);
    const int K = 1 << (l.logBt + l.logCt);
    l.pad = K == 1 ? 0 : (K >= 32 ? 1 : 32 / K);
    return l;
}

constexpr size_t kTileWords = 16384;

// one in-place (or src->dst) DIF digit pass over rows of `bits_total` index bits
nttf::FastTables fast_tables(ts_ctx *c) {
    nttf::FastTables t;
    t.tw_small = c->tw_small;
    t.tw_small_inv = c->tw_small_inv;
    t.tw_big = c->tw_big;
    t.big_log = c->big_log;
    t.small_log = ntt::SMALL_LOG;
    t.rt[0] = c->round_tab[0];
    t.rt[1] = c->round_tab[1];
    return t;
}
bool fast_shape(int d, size_t w) {
    static const bool off = getenv("TS_NO_FAST") != nullptr;
    return !off && d >= 9 && d <= 11 && w >= 8 && (w & 3) == 0;
}
bool use_pm() {
    static const bool off = getenv("TS_NO_PM") != nullptr;
    return !off;
}
bool use_persistent() {  // measured slower than 2 independent CTAs per SM (profiles/r01/README.md); opt-in
    static const bool on = getenv("TS_PERSIST") != nullptr;
    return on;
}
template <int D>
int launch_pass_fast(ts_ctx *c, bool inverse, const nttf::FastPassParams &p, size_t blocks) {
    const size_t smem = (size_t)16384 * 4;
    KScope ks(c, TS_K_NTT_PASS);
    if (use_pm() && use_persistent()) {
        nttp::PersistPassParams pp;
        pp.p = p;
        pp.n_tiles = (uint32_t)blocks;
        const unsigned grid = (unsigned)std::min<size_t>(blocks, (size_t)c->num_sms);
        if (inverse) {
            auto kfn = nttp::ntt_pass_pm2_kernel<D, true>;
            TS_LAUNCH(kfn, grid, nttp::PM2_NT, 2 * smem, c->stream, pp);
        } else {
            auto kfn = nttp::ntt_pass_pm2_kernel<D, false>;
            TS_LAUNCH(kfn, grid, nttp::PM2_NT, 2 * smem, c->stream, pp);
        }
        return check_launch(c, "ntt_pass_pm2_kernel");
    }
    if (use_pm()) {
        if (inverse) {
            auto kfn = nttp::ntt_pass_pm_kernel<D, true>;
            TS_LAUNCH(kfn, (unsigned)blocks, nttp::PM_PASS_NT, smem, c->stream, p);
        } else {
            auto kfn = nttp::ntt_pass_pm_kernel<D, false>;
            TS_LAUNCH(kfn, (unsigned)blocks, nttp::PM_PASS_NT, smem, c->stream, p);
        }
        return check_launch(c, "ntt_pass_pm_kernel");
    }
    if (inverse) {
        auto kfn = nttf::ntt_pass_fast_kernel<D, true>;
        TS_LAUNCH(kfn, (unsigned)blocks, nttf::PASS_NT, smem, c->stream, p);
    } else {
        auto kfn = nttf::ntt_pass_fast_kernel<D, false>;
        TS_LAUNCH(kfn, (unsigned)blocks, nttf::PASS_NT, smem, c->stream, p);
    }
    return check_launch(c, "ntt_pass_fast_kernel");
}

// w columns are transformed; src_pitch/dst_pitch (0 = w) are the row strides when src/dst are column windows of
// wider matrices (fast path only)
struct PeerDst {  // row owners of a column-sharded LDE (ts_coset_lde_batch_scatter)
    uint32_t *ptr[8];
    int n;
};
int launch_pass(ts_ctx *c, bool inverse, const uint32_t *src, uint32_t *dst, size_t w, int d, int lo_bits,
                int hi_bits, bool has_scale, uint2 scale, size_t src_pitch = 0, size_t dst_pitch = 0,
                size_t src_slice = 0, size_t dst_slice = 0, const PeerDst *peers = nullptr) {
    if (!src_pitch) src_pitch = w;
    if (!dst_pitch) dst_pitch = w;
    if ((src_slice || dst_slice) && !(use_pm() && !has_scale && fast_shape(d, w)))
        TS_FAIL(c, TS_ERR_ARG, "blocked layouts need the position-major NTT path");
    if (!has_scale && fast_shape(d, w)) {
        nttf::FastPassParams fp;
        fp.src = src;
        fp.dst = dst;
        fp.src_pitch = (uint32_t)src_pitch;
        fp.dst_pitch = (uint32_t)dst_pitch;
        fp.src_slice = src_slice;
        fp.dst_slice = dst_slice;
        fp.ncols = (uint32_t)w;
        fp.lo_bits = lo_bits;
        fp.hi_bits = hi_bits;
        const size_t K = (size_t)1 << (14 - d);
        fp.n_col_slices = (uint32_t)((w + K - 1) / K);
        fp.t = fast_tables(c);
        fp.tw_shift = lo_bits > 0 ? c->big_log - (lo_bits + d) : 0;
        if (peers) {
            if (!use_pm() || use_persistent() || lo_bits != 0 || inverse)
                TS_FAIL(c, TS_ERR_ARG, "lde scatter: needs the position-major last forward pass");
            for (int i = 0; i < peers->n; i++) fp.peer[i] = peers->ptr[i];
            fp.peer_log_rows = hi_bits + d - log2_strict((size_t)peers->n);  // rows of the output / peers
        }
        const size_t blocks = ((size_t)1 << (hi_bits + lo_bits)) * fp.n_col_slices;
        switch (d) {
            case 9: return launch_pass_fast<9>(c, inverse, fp, blocks);
            case 10: return launch_pass_fast<10>(c, inverse, fp, blocks);
            default: return launch_pass_fast<11>(c, inverse, fp, blocks);
        }
    }
    if (src_pitch != w || dst_pitch != w) TS_FAIL(c, TS_ERR_ARG, "column windows need the fast NTT path");
    ntt::PassParams p;
    p.src = src;
    p.dst = dst;
    p.width = (uint32_t)w;
    p.d = d;
    p.lo_bits = lo_bits;
    p.hi_bits = hi_bits;
    p.batch_lo = lo_bits > 0;
    const Lanes ln = choose_lanes(d, w, p.batch_lo ? lo_bits : hi_bits, kTileWords);
    p.logBt = ln.logBt;
    p.logCt = ln.logCt;
    p.pad = ln.pad;
    p.n_col_slices = (uint32_t)((w + ((size_t)1 << ln.logCt) - 1) >> ln.logCt);
    p.post_tw = lo_bits > 0;
    p.big_log = c->big_log;
    p.tw_shift = p.post_tw ? c->big_log - (lo_bits + d) : 0;
    p.tw_small = c->tw_small;
    p.tw_big = c->tw_big;
    p.has_scale = has_scale;
    p.scale = scale;
    size_t tiles = p.batch_lo ? (((size_t)1 << hi_bits) << (lo_bits - ln.logBt))
                              : ((((size_t)1 << hi_bits) + ((size_t)1 << ln.logBt) - 1) >> ln.logBt);
    const size_t blocks = tiles * p.n_col_slices;
    const size_t smem = ((size_t)((1 << d) + ln.pad) << (ln.logBt + ln.logCt)) * 4;
    KScope ks(c, TS_K_NTT_PASS);
    if (inverse) {
        auto kfn = ntt::ntt_pass_kernel<true>;
        TS_LAUNCH(kfn, (unsigned)blocks, 256, smem, c->stream, p);
    } else {
        auto kfn = ntt::ntt_pass_kernel<false>;
        TS_LAUNCH(kfn, (unsigned)blocks, 256, smem, c->stream, p);
    }
    return check_launch(c, "ntt_pass_kernel");
}

std::vector<int> split_digits(int m) {
    std::vector<int> d;
    if (m <= 0) return d;
    if (const char *env = getenv("TS_DIGITS")) {  // test hook: force a digit split, e.g. "11,9"
        int sum = 0;
        for (const char *q = env; *q;) {
            const int v = (int)strtol(q, (char **)&q, 10);
            if (v < 1 || v > ntt::MAX_DIGIT) break;
            d.push_back(v);
            sum += v;
            if (*q == ',') q++;
        }
        if (sum == m) return d;
        d.clear();
    }
    const int D = (m + ntt::MAX_DIGIT - 1) / ntt::MAX_DIGIT;
    for (int i = 0; i < D; i++) d.push_back(m / D + (i < m % D ? 1 : 0));
    return d;
}

// all digits of a full transform, top-down, in place on `data` (first pass may read `src`).
// Output is in bit-reversed row order.
int full_transform(ts_ctx *c, bool inverse, const uint32_t *src, uint32_t *data, size_t w, int m) {
    const std::vector<int> dg = split_digits(m);
    if (dg.size() > 1) TS_TRY(ensure_big(c, m));
    int used = 0;
    for (size_t i = 0; i < dg.size(); i++) {
        const int lo_bits = m - used - dg[i];
        const bool last = i + 1 == dg.size();
        uint2 scale = make_uint2(0, 0);
        if (inverse && last) {
            const uint32_t ninv = bb::cinv((uint32_t)(((uint64_t)1 << m) % bb::P));
            scale = make_uint2(ninv, bb::cshoup_prime(ninv));
        }
        TS_TRY(launch_pass(c, inverse, i == 0 ? src : data, data, w, dg[i], lo_bits, used, inverse && last, scale));
        used += dg[i];
    }
    return TS_OK;
}

int get_coset_tables(ts_ctx *c, int m, int b, int d, uint32_t shift_monty, uint2 **pre, uint2 **lane) {
    auto key = std::make_tuple(m, b, d, shift_monty);
    auto it = c->coset_tabs.find(key);
    if (it != c->coset_tabs.end()) {
        *pre = it->second.first;
        *lane = it->second.second;
        return TS_OK;
    }
    const int klo = m - d;
    uint2 *p = nullptr, *l = nullptr;
    TS_CUDA(c, cudaMalloc((void **)&p, (sizeof(uint2) << d) << b));
    TS_CUDA(c, cudaMalloc((void **)&l, (sizeof(uint2) << klo) << b));
    const uint32_t wN = h_to_monty(bb::two_adic_generator(m + b));
    const uint32_t ninv = h_to_monty(bb::cinv((uint32_t)(((uint64_t)1 << m) % bb::P)));
    const uint32_t total = ((1u << d) + (1u << klo)) << b;
    {
        KScope ks(c, TS_K_MISC);
        auto kfn = ntt::gen_coset_tables_kernel;
        TS_LAUNCH(kfn, (total + 255) / 256, 256, 0, c->stream, p, l, d, klo, b, shift_monty, wN, ninv);
    }
    TS_TRY(check_launch(c, "gen_coset_tables_kernel"));
    c->coset_tabs[key] = {p, l};
    *pre = p;
    *lane = l;
    return TS_OK;
}

// dst (n << b) x w  <-  committed-order coset LDE of src (n x w)
bool all_digits_fast(int m, size_t w) {
    const std::vector<int> dg = split_digits(m);
    if (dg.empty()) return false;
    for (int d : dg)
        if (!fast_shape(d, w)) return false;
    return true;
}

bool v4_shape(int m, size_t w);
int lde_committed_v4(ts_ctx *c, const uint32_t *src, size_t n, size_t w, unsigned b, uint32_t shift_monty, uint32_t *dst,
                     size_t src_pitch, size_t dst_pitch);
// windows: src is n x w with row stride src_pitch, dst is (n<<b) x w with row stride dst_pitch (0 = w)
int lde_committed(ts_ctx *c, const uint32_t *src, size_t n, size_t w, unsigned b, uint32_t shift_monty,
                  uint32_t *dst, size_t src_pitch = 0, size_t dst_pitch = 0, const PeerDst *peers = nullptr) {
    if (!src_pitch) src_pitch = w;
    if (!dst_pitch) dst_pitch = w;
    const int m = log2_strict(n);
    if (m < 0 || w == 0 || m + (int)b > 27) TS_FAIL(c, TS_ERR_ARG, "lde: rows must be a power of two, rows<<added_bits <= 2^27");
    if ((src_pitch != w || dst_pitch != w) && !all_digits_fast(m, w))
        TS_FAIL(c, TS_ERR_ARG, "column windows need the fast NTT path");
    if (peers && m < 2) TS_FAIL(c, TS_ERR_ARG, "lde scatter: shape not on the blocked position-major path");
    if (m == 0) {
        KScope ks(c, TS_K_MISC);
        auto kfn = ntt::broadcast_row_kernel;
        TS_LAUNCH(kfn, 64, 256, 0, c->stream, src, dst, (size_t)1 << b, (uint32_t)w);
        return check_launch(c, "broadcast_row_kernel");
    }
    if (!peers && v4_shape(m, w)) return lde_committed_v4(c, src, n, w, b, shift_monty, dst, src_pitch, dst_pitch);
    const std::vector<int> dg = split_digits(m);
    const size_t D = dg.size();
    const int dK = dg[D - 1];
    const int klo = m - dK;
    // Blocked intermediates ([col/8][row][8], see ntt_pm.cuh word_off): only the first read and the last write
    // of a multi-digit LDE use the caller's row-major matrices.
    const bool blocked = D > 1 && use_pm() && all_digits_fast(m, w) && getenv("TS_NO_BLOCKED") == nullptr;
    if (peers && !blocked) TS_FAIL(c, TS_ERR_ARG, "lde scatter: shape not on the blocked position-major path");
    const size_t w8 = (w + 7) & ~(size_t)7, N = n << b;
    const size_t s_slice = blocked ? n * 8 : 0, i_slice = blocked ? N * 8 : 0;
    uint32_t *inter = nullptr;  // blocked N x w intermediate between lde_mid and the last forward pass
    if (D > 1) {
        TS_TRY(ensure_big(c, m));
        TS_TRY(ensure_scratch(c, n * (blocked ? w8 : w)));
    }
    if (blocked) TS_CUDA(c, pool_alloc(c, (void **)&inter, N * w8 * 4));
    uint2 *pre, *lane;
    int rc = get_coset_tables(c, m, (int)b, dK, shift_monty, &pre, &lane);
    // inverse passes over the top digits
    int used = 0;
    for (size_t i = 0; i + 1 < D && rc == TS_OK; i++) {
        const int lo_bits = m - used - dg[i];
        rc = launch_pass(c, true, i == 0 ? src : c->scratch, c->scratch, w, dg[i], lo_bits, used, false, make_uint2(0, 0),
                         i == 0 ? src_pitch : w, w, i == 0 ? 0 : s_slice, s_slice);
        used += dg[i];
    }
    // middle kernel
    uint32_t *mid_dst = blocked ? inter : dst;
    if (rc != TS_OK) {
    } else if (fast_shape(dK, w)) {
        nttf::FastMidParams fp;
        fp.src = D > 1 ? c->scratch : src;
        fp.dst = mid_dst;
        fp.src_pitch = (uint32_t)(D > 1 ? w : src_pitch);
        fp.dst_pitch = (uint32_t)dst_pitch;
        fp.src_slice = D > 1 ? s_slice : 0;
        fp.dst_slice = i_slice;
        fp.ncols = (uint32_t)w;
        fp.klo_bits = klo;
        fp.b = (int)b;
        const size_t K = (size_t)1 << (14 - dK);
        fp.n_col_slices = (uint32_t)((w + K - 1) / K);
        fp.t = fast_tables(c);
        fp.tw_shift = klo > 0 ? c->big_log - m : 0;
        fp.pre_tab = pre;
        fp.lane_tab = lane;
        const size_t blocks = ((size_t)1 << klo) * fp.n_col_slices;
        const size_t smem = (size_t)2 * 16384 * 4 + (((size_t)2 << dK) + 1024 + K) * sizeof(uint2) + ((size_t)8 << dK);  // + post table
        KScope ks(c, TS_K_LDE_MID);
        if (use_pm() && use_persistent()) {
            nttp::PersistMidParams pp;
            pp.p = fp;
            pp.n_tiles = (uint32_t)blocks;
            const unsigned grid = (unsigned)std::min<size_t>(blocks, (size_t)c->num_sms);
            const size_t smem2 = (size_t)3 * 16384 * 4 + (((size_t)1 << dK) + 1024) * sizeof(uint2);
            if (dK == 9) {
                auto kfn = nttp::lde_mid_pm2_kernel<9>;
                TS_LAUNCH(kfn, grid, nttp::PM2_NT, smem2, c->stream, pp);
            } else if (dK == 10) {
                auto kfn = nttp::lde_mid_pm2_kernel<10>;
                TS_LAUNCH(kfn, grid, nttp::PM2_NT, smem2, c->stream, pp);
            } else {
                auto kfn = nttp::lde_mid_pm2_kernel<11>;
                TS_LAUNCH(kfn, grid, nttp::PM2_NT, smem2, c->stream, pp);
            }
        } else if (use_pm()) {
            // one resident CTA per SM (the kernel's 180+ KiB of shared memory allow no more), each walking a contiguous
            // tile range (8 ranges per SM: still ~50 tiles per table fill, and the ranges rebalance if an SM is late).
            // Not when the GPU is shared (ts_coset_lde_batch_into / _scatter, the sharded prover): beside the NCCL
            // all-to-all of the previous column chunk the long-lived CTAs were displaced and ran late (N=2: 11.4 ->
            // 12.3 ms with 8 ranges per SM, N=8: 3.1 -> 4.6 ms with one), so those calls launch one CTA per tile.
            fp.n_tiles = (uint32_t)blocks;
            const unsigned grid = (c->shares_gpu || getenv("TS_MID_NOT_PERSISTENT"))
                                      ? (unsigned)blocks
                                      : (unsigned)std::min<size_t>(blocks, (size_t)c->num_sms * 8);
            if (dK == 9) {
                auto kfn = nttp::lde_mid_pm_kernel<9>;
                TS_LAUNCH(kfn, grid, nttp::PM_MID_NT, smem, c->stream, fp);
            } else if (dK == 10) {
                auto kfn = nttp::lde_mid_pm_kernel<10>;
                TS_LAUNCH(kfn, grid, nttp::PM_MID_NT, smem, c->stream, fp);
            } else {
                auto kfn = nttp::lde_mid_pm_kernel<11>;
                TS_LAUNCH(kfn, grid, nttp::PM_MID_NT, smem, c->stream, fp);
            }
        } else if (dK == 9) {
            auto kfn = nttf::lde_mid_fast_kernel<9>;
            TS_LAUNCH(kfn, (unsigned)blocks, nttf::MID_NT, smem, c->stream, fp);
        } else if (dK == 10) {
            auto kfn = nttf::lde_mid_fast_kernel<10>;
            TS_LAUNCH(kfn, (unsigned)blocks, nttf::MID_NT, smem, c->stream, fp);
        } else {
            auto kfn = nttf::lde_mid_fast_kernel<11>;
            TS_LAUNCH(kfn, (unsigned)blocks, nttf::MID_NT, smem, c->stream, fp);
        }
        rc = check_launch(c, "lde_mid (fast) kernel");
    } else {
        ntt::MidParams p;
        p.src = D > 1 ? c->scratch : src;
        p.dst = dst;
        p.width = (uint32_t)w;
        p.d = dK;
        p.klo_bits = klo;
        p.b = (int)b;
        const Lanes ln = choose_lanes(dK, w, klo, kTileWords);
        p.logBt = ln.logBt;
        p.logCt = ln.logCt;
        p.pad = ln.pad;
        p.n_col_slices = (uint32_t)((w + ((size_t)1 << ln.logCt) - 1) >> ln.logCt);
        p.big_log = c->big_log;
        p.tw_shift = klo > 0 ? c->big_log - m : 0;
        p.tw_small = c->tw_small;
        p.tw_big = c->tw_big;
        p.pre_tab = pre;
        p.lane_tab = lane;
        const size_t blocks = ((size_t)1 << (klo - ln.logBt)) * p.n_col_slices;
        const int K = 1 << (ln.logBt + ln.logCt);
        const size_t smem = (size_t)2 * K * ((1 << dK) + ln.pad) * 4 + (size_t)K * sizeof(uint2);
        KScope ks(c, TS_K_LDE_MID);
        auto kfn = ntt::lde_mid_kernel;
        TS_LAUNCH(kfn, (unsigned)blocks, 512, smem, c->stream, p);
        rc = check_launch(c, "lde_mid_kernel");
    }
    // forward passes over the remaining digits, all cosets at once; the last one writes the caller's matrix
    used = 0;
    for (size_t i = 0; i + 1 < D && rc == TS_OK; i++) {
        const int lo_bits = klo - used - dg[i];
        const int hi_bits = m + (int)b - lo_bits - dg[i];
        const bool last = i + 2 == D;
        if (blocked)
            rc = launch_pass(c, false, inter, last ? dst : inter, w, dg[i], lo_bits, hi_bits, false, make_uint2(0, 0), w,
                             last ? dst_pitch : w, i_slice, last ? 0 : i_slice, last ? peers : nullptr);
        else
            rc = launch_pass(c, false, dst, dst, w, dg[i], lo_bits, hi_bits, false, make_uint2(0, 0), dst_pitch, dst_pitch);
        used += dg[i];
    }
    if (inter) pool_release(c, inter);  // stream-ordered: later allocations on this stream may reuse it
    return rc;
}

// ---- ntt_v4.cuh: two-digit LDE as four launches of one pass kernel ---------------------------------------------------
bool use_v4() {
    static const bool off = getenv("TS_NO_V4") != nullptr;
    return !off;
}
bool v4_shape(int m, size_t w) {
    if (!use_v4() || !use_pm() || getenv("TS_NO_BLOCKED")) return false;
    const std::vector<int> dg = split_digits(m);
    return dg.size() == 2 && all_digits_fast(m, w);
}
template <int D>
int launch_v4(ts_ctx *c, int kind, const ntt4::PassParams &p, size_t blocks) {  // kind: 1..4 = P1..P4 (ntt_v4.cuh)
    const size_t smem = (size_t)16384 * 4;
    KScope ks(c, kind == 3 ? TS_K_LDE_MID : TS_K_NTT_PASS);
    if (kind == 1) {
        auto kfn = ntt4::pass_kernel<D, true, false, false>;
        TS_LAUNCH(kfn, (unsigned)blocks, ntt4::V4_NT, smem, c->stream, p);
    } else if (kind == 2) {
        auto kfn = ntt4::pass_kernel<D, true, false, true>;
        TS_LAUNCH(kfn, (unsigned)blocks, ntt4::V4_NT, smem, c->stream, p);
    } else if (kind == 3) {
        auto kfn = ntt4::pass_kernel<D, false, true, true>;
        TS_LAUNCH(kfn, (unsigned)blocks, ntt4::V4_NT, smem, c->stream, p);
    } else {
        auto kfn = ntt4::pass_kernel<D, false, false, true>;
        TS_LAUNCH(kfn, (unsigned)blocks, ntt4::V4_NT, smem, c->stream, p);
    }
    return check_launch(c, "ntt4::pass_kernel");
}
#ifndef TS_EMULATE
// TMA form of the contiguous-source passes (ntt_v4.cuh: pass_tma_kernel); TS_TMA=1
bool use_tma() { return getenv("TS_TMA") != nullptr; }  // read per call: tests and A/B runs flip it inside one process
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}
template <int D>
int launch_v4_tma(ts_ctx *c, bool inverse, const ntt4::PassParams &p, size_t blocks, size_t rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) TS_FAIL(c, TS_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    // blocked source [col / 8][row][8] as a 4-D tensor {8 columns, 256 rows, rows / 256, column groups}
    CUtensorMap map;
    const cuuint64_t dims[4] = {8, 256, rows / 256, (p.ncols + 7) / 8};
    const cuuint64_t strides[3] = {32, 32 * 256, (cuuint64_t)p.src_slice * 4};
    const cuuint32_t box[4] = {4, 256, (cuuint32_t)((1u << D) / 256), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<uint32_t *>(p.src), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) TS_FAIL(c, TS_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    const size_t smem = (size_t)16384 * 4 + 1024 + 16;
    KScope ks(c, TS_K_NTT_PASS);
    if (inverse) {
        auto kfn = ntt4::pass_tma_kernel<D, true>;
        TS_LAUNCH(kfn, (unsigned)blocks, ntt4::V4_NT, smem, c->stream, p, map);
    } else {
        auto kfn = ntt4::pass_tma_kernel<D, false>;
        TS_LAUNCH(kfn, (unsigned)blocks, ntt4::V4_NT, smem, c->stream, p, map);
    }
    return check_launch(c, "ntt4::pass_tma_kernel");
}
#endif
int launch_v4_d(ts_ctx *c, int d, int kind, ntt4::PassParams &p, size_t tiles) {
    const size_t K = (size_t)1 << (14 - d);
    p.n_col_slices = (uint32_t)((p.ncols + K - 1) / K);
    p.cs_shift = log2_strict(p.n_col_slices);
    p.t = fast_tables(c);
    const size_t blocks = tiles * p.n_col_slices;
#ifndef TS_EMULATE
    if (d == 11 && getenv("TS_SHFL") != nullptr) {  // last two rounds exchanged through warp shuffles (A/B form, D = 11 tiles)
        const size_t smem = (size_t)16384 * 4;
        KScope ks(c, kind == 3 ? TS_K_LDE_MID : TS_K_NTT_PASS);
        if (kind == 1) {
            auto kfn = ntt4::pass_shfl_kernel<11, true, false, false>;
            TS_LAUNCH(kfn, (unsigned)blocks, ntt4::V4_NT, smem, c->stream, p);
        } else if (kind == 2) {
            auto kfn = ntt4::pass_shfl_kernel<11, true, false, true>;
            TS_LAUNCH(kfn, (unsigned)blocks, ntt4::V4_NT, smem, c->stream, p);
        } else if (kind == 3) {
            auto kfn = ntt4::pass_shfl_kernel<11, false, true, true>;
            TS_LAUNCH(kfn, (unsigned)blocks, ntt4::V4_NT, smem, c->stream, p);
        } else {
            auto kfn = ntt4::pass_shfl_kernel<11, false, false, true>;
            TS_LAUNCH(kfn, (unsigned)blocks, ntt4::V4_NT, smem, c->stream, p);
        }
        return check_launch(c, "ntt4::pass_shfl_kernel");
    }
    if (use_tma() && (kind == 2 || kind == 4) && p.src_slice != 0 && p.lo_bits == 0) {
        const size_t rows = tiles << d;  // contiguous source: tiles * 2^d rows
        switch (d) {
            case 9: return launch_v4_tma<9>(c, kind == 2, p, blocks, rows);
            case 10: return launch_v4_tma<10>(c, kind == 2, p, blocks, rows);
            default: return launch_v4_tma<11>(c, kind == 2, p, blocks, rows);
        }
    }
#endif
    switch (d) {
        case 9: return launch_v4<9>(c, kind, p, blocks);
        case 10: return launch_v4<10>(c, kind, p, blocks);
        default: return launch_v4<11>(c, kind, p, blocks);
    }
}
int v4_get_post1(ts_ctx *c, int m, int d0, uint2 **out) {
    auto key = std::make_pair(m, d0);
    auto it = c->v4_post1.find(key);
    if (it != c->v4_post1.end()) {
        *out = it->second;
        return TS_OK;
    }
    uint2 *t = nullptr;
    TS_CUDA(c, cudaMalloc((void **)&t, sizeof(uint2) << m));
    {
        KScope ks(c, TS_K_MISC);
        auto kfn = ntt4::fill_post_table_kernel;
        TS_LAUNCH(kfn, (unsigned)((((size_t)1 << m) + 255) / 256), 256, 0, c->stream, t, c->tw_big, c->big_log, m, d0, m - d0, 1);
    }
    TS_TRY(check_launch(c, "fill_post_table_kernel"));
    c->v4_post1[key] = t;
    *out = t;
    return TS_OK;
}
int v4_get_mid(ts_ctx *c, int m, int b, int dK, uint32_t shift_monty, uint2 **post3, uint2 **pre4) {
    auto key = std::make_tuple(m, b, dK, shift_monty);
    auto it = c->v4_mid.find(key);
    if (it != c->v4_mid.end()) {
        it->second.stamp = ++c->v4_clock;
        *post3 = it->second.post3;
        *pre4 = it->second.pre4;
        return TS_OK;
    }
    // a prover commits a handful of (size, shift) pairs per proof (trace + quotient chunks); beyond 12 cached shapes the
    // least recently used tables go back to the pool (stream-ordered, so work already queued on them is safe)
    while (c->v4_mid.size() >= 12) {
        auto old = c->v4_mid.begin();
        for (auto q = c->v4_mid.begin(); q != c->v4_mid.end(); ++q)
            if (q->second.stamp < old->second.stamp) old = q;
        pool_release(c, old->second.post3);
        pool_release(c, old->second.pre4);
        c->v4_mid.erase(old);
    }
    uint2 *pre, *lane;
    TS_TRY(get_coset_tables(c, m, b, dK, shift_monty, &pre, &lane));
    const int klo = m - dK;
    ts_ctx::V4Mid t;
    TS_CUDA(c, pool_alloc(c, (void **)&t.post3, sizeof(uint2) << (m + b)));
    TS_CUDA(c, pool_alloc(c, (void **)&t.pre4, sizeof(uint2) << (dK + b)));
    {
        KScope ks(c, TS_K_MISC);
        auto kfn = ntt4::fill_mid_post_table_kernel;
        TS_LAUNCH(kfn, (unsigned)((((size_t)1 << (m + b)) + 255) / 256), 256, 0, c->stream, t.post3, c->tw_big, lane, c->big_log, m,
                  dK, klo, b);
    }
    TS_TRY(check_launch(c, "fill_mid_post_table_kernel"));
    {
        KScope ks(c, TS_K_MISC);
        auto kfn = ntt4::fill_pre_table_kernel;
        TS_LAUNCH(kfn, ((1u << (dK + b)) + 255) / 256, 256, 0, c->stream, t.pre4, pre, dK, b);
    }
    TS_TRY(check_launch(c, "fill_pre_table_kernel"));
    t.stamp = ++c->v4_clock;
    c->v4_mid[key] = t;
    *post3 = t.post3;
    *pre4 = t.pre4;
    return TS_OK;
}
// dst (n << b) x w  <-  committed-order coset LDE of src (n x w); windows as in lde_committed
int lde_committed_v4(ts_ctx *c, const uint32_t *src, size_t n, size_t w, unsigned b, uint32_t shift_monty, uint32_t *dst,
                     size_t src_pitch, size_t dst_pitch) {
    const int m = log2_strict(n);
    const std::vector<int> dg = split_digits(m);
    const int d0 = dg[0], dK = dg[1], klo = m - dK;
    const size_t w8 = (w + 7) & ~(size_t)7, N = n << b;
    const size_t s_slice = n * 8, i_slice = N * 8;
    TS_TRY(ensure_big(c, m));
    TS_TRY(ensure_scratch(c, n * w8));
    uint2 *post1, *post3, *pre4;
    TS_TRY(v4_get_post1(c, m, d0, &post1));
    TS_TRY(v4_get_mid(c, m, (int)b, dK, shift_monty, &post3, &pre4));
    uint32_t *inter = nullptr;
    TS_CUDA(c, pool_alloc(c, (void **)&inter, N * w8 * 4));
    int rc;
    {  // P1: inverse, top digit; caller's rows -> blocked scratch
        ntt4::PassParams p;
        p.src = src, p.dst = c->scratch, p.src_pitch = (uint32_t)src_pitch, p.dst_pitch = (uint32_t)w, p.ncols = (uint32_t)w;
        p.src_slice = 0, p.dst_slice = s_slice, p.lo_bits = m - d0, p.hi_bits = 0, p.post = post1;
        rc = launch_v4_d(c, d0, 1, p, (size_t)1 << (m - d0));
    }
    if (rc == TS_OK) {  // P2: inverse, low digit, in place (bit-reversed coefficient order, unscaled)
        ntt4::PassParams p;
        p.src = c->scratch, p.dst = c->scratch, p.src_pitch = p.dst_pitch = (uint32_t)w, p.ncols = (uint32_t)w;
        p.src_slice = p.dst_slice = s_slice, p.lo_bits = 0, p.hi_bits = m - dK;
        rc = launch_v4_d(c, dK, 2, p, (size_t)1 << (m - dK));
    }
    if (rc == TS_OK) {  // P3: forward, top digit, once per coset
        ntt4::PassParams p;
        p.src = c->scratch, p.dst = inter, p.src_pitch = p.dst_pitch = (uint32_t)w, p.ncols = (uint32_t)w;
        p.src_slice = s_slice, p.dst_slice = i_slice, p.klo_bits = klo, p.b = (int)b, p.m = m, p.post = post3, p.pre = pre4;
        rc = launch_v4_d(c, dK, 3, p, (size_t)1 << (klo + b));
    }
    if (rc == TS_OK) {  // P4: forward, low digit; blocked intermediate -> caller's rows in committed order
        ntt4::PassParams p;
        p.src = inter, p.dst = dst, p.src_pitch = (uint32_t)w, p.dst_pitch = (uint32_t)dst_pitch, p.ncols = (uint32_t)w;
        p.src_slice = i_slice, p.dst_slice = 0, p.lo_bits = klo - d0, p.hi_bits = m + (int)b - klo;
        rc = launch_v4_d(c, d0, 4, p, (size_t)1 << (m + b - d0));
    }
    pool_release(c, inter);
    return rc;
}

// Host trace -> committed LDE on the device, H2D overlapped with compute: the matrix is cut into column chunks;
// chunk k+1 is copied (strided 2-D copy out of the row-major host matrix) while chunk k is transformed.  Columns
// are independent polynomials, so the result is identical to the one-shot LDE.
size_t chunk_cols() {
    if (const char *e = getenv("TS_CHUNK_COLS")) return (size_t)strtoul(e, nullptr, 10);  // test hook
    return 64;
}
bool taper_last_chunk() { return getenv("TS_NO_TAPER") == nullptr; }
bool pipeline_eligible(size_t n, size_t w) {
    const int m = log2_strict(n);
    const size_t wc = chunk_cols();
    return m >= 18 && wc >= 8 && w >= 2 * wc && (w & 3) == 0 && all_digits_fast(m, wc) &&
           (w % wc == 0 || all_digits_fast(m, w % wc)) && getenv("TS_NO_PIPELINE") == nullptr;
}
// Leaf hashing of ONE matrix restricted to the 64-byte blocks [blk_begin, blk_end) of every row; the chaining value
// lives in `digests` between calls (hash.cuh: hash_rows_fast_kernel).  Rows of at most one Blake3 chunk.
int hash_rows_window(ts_ctx *c, const uint32_t *mat, size_t width, size_t n_leaves, uint32_t *digests, uint32_t blk_begin,
                     uint32_t blk_end) {
    b3::FastSegs fs;
    for (int i = 0; i < b3::MAX_SEG; i++) fs.ptr[i] = i == 0 ? mat : nullptr;
    fs.n = 1;
    fs.seg_w = (uint32_t)width;
    fs.log_seg_w = 0;
    KScope ks(c, TS_K_HASH_LEAVES);
    auto kfn = b3::hash_rows_fast_kernel;
    const size_t per_block = (size_t)b3::FAST_WARPS * 32;
    TS_LAUNCH(kfn, (unsigned)((n_leaves + per_block - 1) / per_block), b3::FAST_WARPS * 32, (size_t)b3::FAST_WARPS * 512 * 4,
              c->stream, fs, (uint32_t)width, n_leaves, 1, digests, blk_begin, blk_end);
    return check_launch(c, "hash_rows_fast_kernel");
}
// rows the pipeline may hash chunk by chunk: one Blake3 chunk (<= 256 words), 16-byte aligned row segments
bool incremental_hash_eligible(size_t w) {
    return w <= 256 && (w & 3) == 0 && chunk_cols() % 16 == 0 && getenv("TS_NO_FAST") == nullptr &&
           getenv("TS_NO_INC_HASH") == nullptr;
}

#ifndef TS_EMULATE
// true when `host` is ordinary pageable memory (neither cudaHostAlloc'ed nor cudaHostRegister'ed)
bool host_is_pageable(const void *host) {
    if (getenv("TS_NO_BOUNCE")) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}
int bounce_workers() {
    int t = (int)std::min<unsigned>(8, std::max(1u, std::thread::hardware_concurrency()));
    if (const char *e = getenv("TS_BOUNCE_THREADS")) t = atoi(e);
    return std::max(1, std::min(t, (int)ts_ctx::BOUNCE_WORKERS));
}
// Window [col0, col0 + cols) of all n rows of a PAGEABLE row-major n x w host matrix -> dense n x cols device buffer.
// A strided 2-D copy out of pageable memory is staged row by row by the driver (measured 5.9 GB/s), and page-locking the
// caller's buffer for the call costs more than the copy (cudaHostRegister of 4 GiB: ~1 s).  Instead worker threads gather
// blocks of rows into page-locked bounce slots (the strided part, done by the CPU cores in parallel) and queue a dense 1-D
// copy per block on their own stream; the DMA of one block overlaps the gather of the next.  `after` (optional) is an event
// the copies must wait for (the destination's previous reader); on return `c->stream` is ordered after all copies.
int stage_pageable_window(ts_ctx *c, const uint32_t *host, size_t n, size_t w, size_t col0, size_t cols, uint32_t *dev,
                          cudaEvent_t after) {
    const int T = bounce_workers();
    if (!c->bounce) {
        TS_CUDA(c, cudaHostAlloc((void **)&c->bounce, ts_ctx::BOUNCE_WORKERS * 2 * ts_ctx::BOUNCE_SLOT_BYTES, cudaHostAllocDefault));
        for (int t = 0; t < ts_ctx::BOUNCE_WORKERS; t++) {
            TS_CUDA(c, cudaStreamCreateWithFlags(&c->bounce_stream[t], cudaStreamNonBlocking));
            TS_CUDA(c, cudaEventCreateWithFlags(&c->bounce_ev[t][0], cudaEventDisableTiming));
            TS_CUDA(c, cudaEventCreateWithFlags(&c->bounce_ev[t][1], cudaEventDisableTiming));
            TS_CUDA(c, cudaEventCreateWithFlags(&c->bounce_done[t], cudaEventDisableTiming));
        }
    }
    const size_t row_bytes = cols * 4;
    const size_t rows_per_block = std::max<size_t>(1, ts_ctx::BOUNCE_SLOT_BYTES / row_bytes);
    const size_t nblocks = (n + rows_per_block - 1) / rows_per_block;
    std::atomic<int> failed{(int)cudaSuccess};
    auto work = [&](int t) {
        cudaError_t e = cudaSetDevice(c->device);
        if (e == cudaSuccess && after) e = cudaStreamWaitEvent(c->bounce_stream[t], after, 0);
        size_t k = 0;
        for (size_t blk = (size_t)t; blk < nblocks && e == cudaSuccess; blk += (size_t)T, k++) {
            const int slot = (int)(k & 1);
            uint8_t *buf = c->bounce + ((size_t)t * 2 + slot) * ts_ctx::BOUNCE_SLOT_BYTES;
            if (c->bounce_ev_used[t][slot]) e = cudaEventSynchronize(c->bounce_ev[t][slot]);  // the slot's previous copy
            if (e != cudaSuccess) break;
            const size_t r0 = blk * rows_per_block, r1 = std::min(n, r0 + rows_per_block);
            const uint8_t *src = reinterpret_cast<const uint8_t *>(host + r0 * w + col0);
            for (size_t r = r0; r < r1; r++, src += w * 4) memcpy(buf + (r - r0) * row_bytes, src, row_bytes);
            e = cudaMemcpyAsync(dev + r0 * cols, buf, (r1 - r0) * row_bytes, cudaMemcpyHostToDevice, c->bounce_stream[t]);
            if (e == cudaSuccess) e = cudaEventRecord(c->bounce_ev[t][slot], c->bounce_stream[t]);
            c->bounce_ev_used[t][slot] = true;
        }
        if (e == cudaSuccess) e = cudaEventRecord(c->bounce_done[t], c->bounce_stream[t]);
        if (e != cudaSuccess) failed.store((int)e);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < T; t++) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
    if (failed.load() != (int)cudaSuccess) {
        for (int t = 0; t < T; t++) cudaStreamSynchronize(c->bounce_stream[t]);  // nothing of ours left in flight
        c->err = std::string("pageable staging: ") + cudaGetErrorString((cudaError_t)failed.load());
        return TS_ERR_CUDA;
    }
    for (int t = 0; t < T; t++) TS_CUDA(c, cudaStreamWaitEvent(c->stream, c->bounce_done[t], 0));
    return TS_OK;
}
#endif

// leaf_digests != nullptr: also absorb every finished chunk into the row hashes (incremental_hash_eligible(w)), so
// that only the last chunk's blocks remain to be hashed when the copy ends.
int lde_from_host_pipelined(ts_ctx *c, const uint32_t *host, size_t n, size_t w, unsigned b, uint32_t shift_monty,
                            uint32_t *dst, uint32_t *leaf_digests = nullptr) {
    const size_t wc = chunk_cols();
    // chunk schedule: (first column, columns).  Everything after the last copy is exposed latency, so the final
    // full chunk is cut in two (TS_NO_TAPER disables): narrower 2-D copies are slower on PCIe (profiles/r01/h2d_2d_b200.jsonl:
    // 256-byte segments 55.6 GB/s, 128-byte 47.7, 64-byte 25), which bounds how far tapering pays.
    std::vector<std::pair<size_t, size_t>> sched;
    for (size_t c0 = 0; c0 < w; c0 += wc) sched.push_back({c0, std::min(wc, w - c0)});
    if (taper_last_chunk() && sched.back().second == wc && wc % 32 == 0 && all_digits_fast(log2_strict(n), wc / 2)) {
        const size_t c0 = sched.back().first;
        sched.back() = {c0, wc / 2};
        sched.push_back({c0 + wc / 2, wc / 2});
    }
    const size_t nchunks = sched.size();
    if (c->stage_words < n * wc) {
        TS_CUDA(c, cudaStreamSynchronize(c->stream));
        for (int s = 0; s < 2; s++) {
            if (c->stage[s]) cudaFree(c->stage[s]);
            c->stage[s] = nullptr;
            TS_CUDA(c, cudaMalloc((void **)&c->stage[s], n * wc * 4));
            c->ev_free_used[s] = false;
        }
        c->stage_words = n * wc;
    }
#ifndef TS_EMULATE
    if (!c->copy_stream) {
        TS_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int s = 0; s < 2; s++) {
            TS_CUDA(c, cudaEventCreateWithFlags(&c->ev_copy[s], cudaEventDisableTiming));
            TS_CUDA(c, cudaEventCreateWithFlags(&c->ev_free[s], cudaEventDisableTiming));
        }
    }
#endif
#ifndef TS_EMULATE
    const bool pageable = host_is_pageable(host);
#endif
    for (size_t ch = 0; ch < nchunks; ch++) {
        const int s = (int)(ch & 1);
        const size_t col0 = sched[ch].first, cols = sched[ch].second;
#ifndef TS_EMULATE
        if (pageable) {
            TS_TRY(stage_pageable_window(c, host, n, w, col0, cols, c->stage[s], c->ev_free_used[s] ? c->ev_free[s] : nullptr));
        } else {
            if (c->ev_free_used[s]) TS_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_free[s], 0));
            TS_CUDA(c, cudaMemcpy2DAsync(c->stage[s], cols * 4, host + col0, w * 4, cols * 4, n, cudaMemcpyHostToDevice,
                                         c->copy_stream));
            TS_CUDA(c, cudaEventRecord(c->ev_copy[s], c->copy_stream));
            TS_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copy[s], 0));
        }
#else
        TS_CUDA(c, cudaMemcpy2DAsync(c->stage[s], cols * 4, host + col0, w * 4, cols * 4, n, cudaMemcpyHostToDevice,
                                     c->copy_stream));
#endif
        TS_TRY(lde_committed(c, c->stage[s], n, cols, b, shift_monty, dst + col0, cols, w));
#ifndef TS_EMULATE
        TS_CUDA(c, cudaEventRecord(c->ev_free[s], c->stream));
        c->ev_free_used[s] = true;
#endif
        if (leaf_digests)
            TS_TRY(hash_rows_window(c, dst, w, n << b, leaf_digests, (uint32_t)(col0 / 16),
                                    (uint32_t)((col0 + cols + 15) / 16)));
    }
    return TS_OK;
}

// Device-resident trace -> committed LDE + row hashes, column chunk by column chunk, the Blake3 window of chunk k on
// a second stream beside the LDE of chunk k+1.  Neither kernel class fills the SM's issue slots on its own (NTT
// 45-65 %, Blake3 73 %), and an lde_mid CTA (512 threads x 96 registers, 184 KiB) leaves exactly the registers and
// shared memory of one leaf-hash CTA.  MEASURED NEUTRAL (profiles/r01/README.md): the kernels do co-run, but each slows
// down by what the other gains (step 51.6 vs 51.9 ms), so this path is opt-in (TS_OVERLAP_HASH=1) and kept as a
// tested experiment.
bool overlap_hash_eligible(size_t n, size_t w) {
    const size_t wc = chunk_cols();
    return getenv("TS_OVERLAP_HASH") != nullptr && incremental_hash_eligible(w) && log2_strict(n) >= 18 && w >= 2 * wc &&
           w % wc == 0 && all_digits_fast(log2_strict(n), wc);
}
int lde_hash_overlapped(ts_ctx *c, const uint32_t *src, size_t n, size_t w, unsigned b, uint32_t shift_monty, uint32_t *dst,
                        uint32_t *leaf_digests) {
    const size_t wc = chunk_cols();
#ifndef TS_EMULATE
    if (!c->side_stream) {
        TS_CUDA(c, cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
        TS_CUDA(c, cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming));
    }
    // the side stream starts after everything already queued on the main stream (the digests' previous users)
    TS_CUDA(c, cudaEventRecord(c->ev_side, c->stream));
    TS_CUDA(c, cudaStreamWaitEvent(c->side_stream, c->ev_side, 0));
#endif
    int rc = TS_OK;
    for (size_t col0 = 0; col0 < w && rc == TS_OK; col0 += wc) {
        rc = lde_committed(c, src + col0, n, wc, b, shift_monty, dst + col0, w, w);
        if (rc != TS_OK) break;
#ifndef TS_EMULATE
        TS_CUDA(c, cudaEventRecord(c->ev_side, c->stream));
        TS_CUDA(c, cudaStreamWaitEvent(c->side_stream, c->ev_side, 0));
        cudaStream_t main_stream = c->stream;
        c->stream = c->side_stream;  // launch + statistics events on the side stream
        rc = hash_rows_window(c, dst, w, n << b, leaf_digests, (uint32_t)(col0 / 16), (uint32_t)((col0 + wc) / 16));
        c->stream = main_stream;
#else
        rc = hash_rows_window(c, dst, w, n << b, leaf_digests, (uint32_t)(col0 / 16), (uint32_t)((col0 + wc) / 16));
#endif
    }
#ifndef TS_EMULATE
    // the tree (main stream) waits for the last window; also on errors, so nothing is left running on `dst`
    cudaEventRecord(c->ev_side, c->side_stream);
    cudaStreamWaitEvent(c->stream, c->ev_side, 0);
#endif
    return rc;
}

int new_matrix(ts_ctx *c, size_t rows, size_t width, ts_matrix **out) {
    ts_matrix *m = new ts_matrix{c, nullptr, rows, width, true};
    cudaError_t e = pool_alloc(c, (void **)&m->d, std::max<size_t>(rows * width, 1) * 4);
    if (e != cudaSuccess) {
        delete m;
        c->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        return TS_ERR_CUDA;
    }
    *out = m;
    return TS_OK;
}

int bitrev_rows(ts_ctx *c, const uint32_t *in, uint32_t *out, int log_h, size_t w) {
    KScope ks(c, TS_K_MISC);
    const size_t total = ((size_t)1 << log_h) * w;
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)c->num_sms * 16);
    auto kfn = ntt::bitrev_rows_kernel;
    TS_LAUNCH(kfn, blocks, 256, 0, c->stream, in, out, log_h, (uint32_t)w);
    return check_launch(c, "bitrev_rows_kernel");
}

// ---- hashing ------------------------------------------------------------------------------------
int hash_rows(ts_ctx *c, const std::vector<const ts_matrix *> &mats, const std::vector<uint32_t> &shifts,
              size_t n_leaves, uint32_t *digests) {
    if (mats.size() > (size_t)b3::MAX_SEG) TS_FAIL(c, TS_ERR_ARG, "mmcs: more than 32 matrices hashed into one layer");
    if (mats.size() == 1 && shifts[0] == 0 && mats[0]->width <= 16) {
        KScope ks(c, TS_K_HASH_LEAVES);
        auto kfn = b3::hash_leaves_small_kernel;
        TS_LAUNCH(kfn, (unsigned)((n_leaves + 255) / 256), 256, 0, c->stream, (const uint32_t *)mats[0]->d,
                  (uint32_t)mats[0]->width, n_leaves, 1, digests);
        return check_launch(c, "hash_leaves_small_kernel");
    }
    {
        // fast path: one matrix, or several of equal power-of-two width (multi-GPU column blocks), one row per leaf
        bool fast = getenv("TS_NO_FAST") == nullptr && (mats[0]->width & 3) == 0;
        size_t total = 0;
        for (size_t i = 0; i < mats.size() && fast; i++) {
            fast = shifts[i] == 0 && mats[i]->width == mats[0]->width;
            total += mats[i]->width;
        }
        if (fast && mats.size() > 1) fast = (mats[0]->width & (mats[0]->width - 1)) == 0 && mats[0]->width >= 4;
        if (fast && total <= ((size_t)256 << b3::MAX_STACK)) {
            b3::FastSegs fs;
            for (int i = 0; i < b3::MAX_SEG; i++) fs.ptr[i] = i < (int)mats.size() ? mats[i]->d : nullptr;
            fs.n = (int)mats.size();
            fs.seg_w = (uint32_t)mats[0]->width;
            fs.log_seg_w = 0;
            while ((1u << fs.log_seg_w) < fs.seg_w) fs.log_seg_w++;
            KScope ks(c, TS_K_HASH_LEAVES);
            auto kfn = b3::hash_rows_fast_kernel;
            const size_t per_block = (size_t)b3::FAST_WARPS * 32;
            TS_LAUNCH(kfn, (unsigned)((n_leaves + per_block - 1) / per_block), b3::FAST_WARPS * 32,
                      (size_t)b3::FAST_WARPS * 512 * 4, c->stream, fs, (uint32_t)total, n_leaves, 1, digests, 0u,
                      (uint32_t)((total + 15) / 16));
            return check_launch(c, "hash_rows_fast_kernel");
        }
    }
    b3::Segments sg;
    sg.n = (int)mats.size();
    sg.total_words = 0;
    for (int i = 0; i < b3::MAX_SEG; i++) {
        sg.ptr[i] = nullptr;
        sg.width[i] = 0xffffffffu;  // sentinel: the segment walk never runs past the last real one
        sg.shift[i] = 0;
    }
    for (size_t i = 0; i < mats.size(); i++) {
        sg.ptr[i] = mats[i]->d;
        sg.width[i] = (uint32_t)mats[i]->width;
        sg.shift[i] = shifts[i];
        sg.total_words += (uint32_t)mats[i]->width;
    }
    if ((sg.total_words + 255) / 256 > (1u << b3::MAX_STACK)) TS_FAIL(c, TS_ERR_ARG, "mmcs: leaf wider than 64 KiB");
    KScope ks(c, TS_K_HASH_LEAVES);
    auto kfn = b3::hash_leaves_kernel;
    const size_t smem = (size_t)b3::LEAVES_PER_CTA * b3::PITCH * 4;
    TS_LAUNCH(kfn, (unsigned)((n_leaves + b3::LEAVES_PER_CTA - 1) / b3::LEAVES_PER_CTA), b3::LEAVES_PER_CTA, smem,
              c->stream, sg, n_leaves, 1, digests);
    return check_launch(c, "hash_leaves_kernel");
}

// leaves_done: the leaf layer of t->digests was filled by the caller (pipelined commit: hash_rows_window)
// Layers of at least 2^k children take tree_reduce3_kernel (three levels per launch, every lane busy: throughput), smaller ones
// tree_reduce_kernel (eight levels per launch out of shared memory: fewer dependent launches).  Both are latency bound on small
// layers (7 serial compressions per thread = 18.5 us against 8 barriers = 10.5 us), so the switch point decides how many
// launches the small FRI rounds cost.  TS_TREE3_MIN_LOG overrides (A/B in profiles/r02/README.md).
int tree3_min_log() {
    if (const char *e = getenv("TS_TREE3_MIN_LOG")) return std::max(3, atoi(e));
    return 19;  // measured on a B200: FRI commit phase of 2^24 elements 2.22 ms (14) / 2.12 (17) / 2.11 (19)
}

int build_tree(ts_ctx *c, ts_tree *t, bool leaves_done = false) {
    const size_t k = t->mats.size();
    size_t n_first = 0;
    if (t->layout == TS_LAYOUT_PADDED) n_first = k;
    else
        while (n_first < k && t->mats[t->order[n_first]]->rows == t->hmax) n_first++;
    if (!leaves_done) {
        std::vector<const ts_matrix *> ms;
        std::vector<uint32_t> sh;
        for (size_t q = 0; q < n_first; q++) {
            const ts_matrix *m = t->mats[t->order[q]];
            ms.push_back(m);
            sh.push_back(t->lmax - (unsigned)log2_strict(m->rows));
        }
        TS_TRY(hash_rows(c, ms, sh, t->hmax, t->digests));
    }
    size_t next_mat = n_first;
    unsigned l = 1;
    while (l <= t->lmax) {
        const size_t len = t->hmax >> l;
        size_t inj0 = next_mat;
        if (t->layout == TS_LAYOUT_P3_INJECT)
            while (next_mat < k && t->mats[t->order[next_mat]]->rows == len) next_mat++;
        const uint32_t *children = t->digests + t->layer_off[l - 1] * 8;
        if (next_mat > inj0) {
            std::vector<const ts_matrix *> ms;
            std::vector<uint32_t> sh;
            for (size_t q = inj0; q < next_mat; q++) {
                ms.push_back(t->mats[t->order[q]]);
                sh.push_back(0);
            }
            TS_TRY(ensure_scratch(c, len * 8));
            TS_TRY(hash_rows(c, ms, sh, len, c->scratch));
            KScope ks(c, TS_K_TREE);
            auto kfn = b3::compress_inject_kernel;
            TS_LAUNCH(kfn, (unsigned)((len + 127) / 128), 128, 0, c->stream, children, (const uint32_t *)c->scratch,
                      len, t->digests + t->layer_off[l] * 8);
            TS_TRY(check_launch(c, "compress_inject_kernel"));
            l++;
            continue;
        }
        // as many plain levels as possible in one launch (stop before the next injection layer)
        int levels = 0;
        while (levels < b3::TREE_MAX_LEVELS && l + levels <= t->lmax) {
            if (levels > 0 && t->layout == TS_LAYOUT_P3_INJECT && next_mat < k &&
                t->mats[t->order[next_mat]]->rows == (t->hmax >> (l + levels)))
                break;
            levels++;
        }
        const size_t n_ch = t->hmax >> (l - 1);
        if (levels >= 3 && n_ch >= ((size_t)1 << tree3_min_log()) && getenv("TS_NO_TREE3") == nullptr) {
            // big layer: three levels per launch, one thread per 8 children
            KScope ks(c, TS_K_TREE);
            auto kfn3 = b3::tree_reduce3_kernel;
            TS_LAUNCH(kfn3, (unsigned)((n_ch / 8 + 127) / 128), 128, 0, c->stream, children, n_ch,
                      t->digests + t->layer_off[l] * 8, t->digests + t->layer_off[l + 1] * 8, t->digests + t->layer_off[l + 2] * 8);
            TS_TRY(check_launch(c, "tree_reduce3_kernel"));
            l += 3;
            continue;
        }
        b3::TreeLevels lv;
        lv.levels = levels;
        for (int i = 0; i < b3::TREE_MAX_LEVELS; i++)
            lv.out[i] = i < levels ? t->digests + t->layer_off[l + i] * 8 : nullptr;
        const size_t n_children = t->hmax >> (l - 1);
        KScope ks(c, TS_K_TREE);
        auto kfn = b3::tree_reduce_kernel;
        TS_LAUNCH(kfn, (unsigned)((n_children / 2 + b3::TREE_T - 1) / b3::TREE_T), b3::TREE_T,
                  (size_t)b3::TREE_T * b3::NODE_PITCH * 4, c->stream, children, n_children, lv);
        TS_TRY(check_launch(c, "tree_reduce_kernel"));
        l += levels;
    }
    return TS_OK;
}

// fill_leaves (optional): called with the leaf-digest array once it is allocated; it produces the matrix AND its row
// hashes (the host pipeline), after which only the levels above the leaves are built here.
// prealloc_digests (optional, single-matrix trees): a pool buffer of (2 hmax - 1) x 8 words whose leaf layer is already filled
// (fold_hash_kernel); the tree takes it over -- also when the call fails.
int mmcs_commit(ts_ctx *ctx, ts_matrix *const *mats, size_t n_mats, int layout, int take_ownership,
                uint8_t root[32], ts_tree **out, bool sync_root,
                const std::function<int(uint32_t *)> *fill_leaves = nullptr, uint32_t *prealloc_digests = nullptr) {
    if (n_mats == 0) {
        if (prealloc_digests) pool_release(ctx, prealloc_digests);
        TS_FAIL(ctx, TS_ERR_ARG, "mmcs: no matrices");
    }
    ts_tree *t = new ts_tree;
    t->ctx = ctx;
    t->mats.assign(mats, mats + n_mats);
    t->own_mats = take_ownership != 0;
    t->layout = layout;
    t->digests = nullptr;
    t->order.resize(n_mats);
    for (size_t i = 0; i < n_mats; i++) t->order[i] = i;
    std::stable_sort(t->order.begin(), t->order.end(),
                     [&](size_t a, size_t b) { return mats[a]->rows > mats[b]->rows; });
    t->hmax = mats[t->order[0]]->rows;
    for (size_t i = 0; i < n_mats; i++)
        if (log2_strict(mats[i]->rows) < 0) {
            t->own_mats = false;
            ts_tree_free(t);
            if (prealloc_digests) pool_release(ctx, prealloc_digests);
            TS_FAIL(ctx, TS_ERR_ARG, "mmcs: heights must be powers of two");
        }
    t->lmax = (unsigned)log2_strict(t->hmax);
    size_t off = 0;
    for (unsigned l = 0; l <= t->lmax; l++) {
        t->layer_off.push_back(off);
        off += t->hmax >> l;
    }
    cudaError_t e = cudaSuccess;
    if (prealloc_digests) t->digests = prealloc_digests;
    else e = pool_alloc(ctx, (void **)&t->digests, off * 32);
    if (e != cudaSuccess) {
        t->own_mats = false;
        ts_tree_free(t);
        TS_FAIL(ctx, TS_ERR_CUDA, std::string("cudaMalloc digests: ") + cudaGetErrorString(e));
    }
    int rc = fill_leaves ? (*fill_leaves)(t->digests) : TS_OK;
    if (rc == TS_OK) rc = build_tree(ctx, t, fill_leaves != nullptr || prealloc_digests != nullptr);
    if (rc == TS_OK && root) {
        cudaError_t e2 = cudaMemcpyAsync(root, t->digests + t->layer_off[t->lmax] * 8, 32, cudaMemcpyDeviceToHost,
                                         ctx->stream);
        if (e2 == cudaSuccess && sync_root) e2 = cudaStreamSynchronize(ctx->stream);
        if (e2 != cudaSuccess) {
            ctx->err = std::string("root download: ") + cudaGetErrorString(e2);
            rc = TS_ERR_CUDA;
        }
    }
    if (rc != TS_OK) {
        t->own_mats = false;
        ts_tree_free(t);
        return rc;
    }
    *out = t;
    return TS_OK;
}

// host-side canonical EF helpers (transcript values; a handful per FRI round)
void h_ef_mul(const uint32_t a[4], const uint32_t b[4], uint32_t o[4]) {
    uint32_t r[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) r[i + j] = (uint32_t)((r[i + j] + (uint64_t)a[i] * b[j]) % bb::P);
    for (int i = 0; i < 4; i++) o[i] = i < 3 ? (uint32_t)((r[i] + 11ull * r[i + 4]) % bb::P) : r[i];
}

// beta_canon == nullptr: beta/2 is read from half_beta_dev (device, Montgomery) by the kernel
int fold_ext_launch(ts_ctx *c, const uint32_t *in, uint32_t *out, const uint32_t *addend, int log_h,
                    const uint32_t beta_canon[4], size_t first = 0, size_t h_local = 0, const uint32_t *half_beta_dev = nullptr) {
    const uint32_t half = bb::cinv(2);
    ef::E4 hb;
    for (int i = 0; i < 4; i++) hb.c[i] = beta_canon ? h_to_monty(bb::cmul(beta_canon[i], half)) : 0;
    if (!beta_canon && !half_beta_dev) TS_FAIL(c, TS_ERR_ARG, "fold: no beta");
    const uint32_t g_inv = bb::cinv(bb::two_adic_generator(log_h + 1));
    fold::InvRootPows rp;
    uint32_t r = g_inv;
    for (int k = 0; k < 28; k++) {
        rp.v[k] = h_to_monty(r);
        r = bb::cmul(r, r);
    }
    fold::ChunkDeltas dl;
    const int Lh = log_h - 8;
    for (int t = 0; t < 28; t++) {
        uint32_t v = 1;
        if (Lh > 0 && t < Lh) {
            // bitrev(c+1) - bitrev(c) for c ending in t ones: + 2^(Lh-1-t) - sum_{k<t} 2^(Lh-1-k), mod 2h
            int64_t delta = (int64_t)1 << (Lh - 1 - t);
            for (int k = 0; k < t; k++) delta -= (int64_t)1 << (Lh - 1 - k);
            const int64_t order = (int64_t)2 << log_h;
            delta = ((delta % order) + order) % order;
            v = bb::cpow(g_inv, (uint64_t)delta);
        }
        dl.v[t] = h_to_monty(v);
    }
    const size_t h = h_local ? h_local : (size_t)1 << log_h;
    unsigned blocks = 1;
    if (log_h >= 8) blocks = (unsigned)std::min<size_t>(h >> 8, (size_t)c->num_sms * 8);
    KScope ks(c, TS_K_FOLD);
    auto kfn = fold::fold_ext_kernel;
    TS_LAUNCH(kfn, blocks, fold::FOLD_T, 0, c->stream, (const uint4 *)in, (uint4 *)out, (const uint4 *)addend, log_h,
              first, h, hb, rp, dl, (const uint32_t *)c->fold_tlo, beta_canon ? (const uint32_t *)nullptr : half_beta_dev);
    return check_launch(c, "fold_ext_kernel");
}

// fold + the leaf hashes of the next round's tree (fold.cuh: fold_hash_kernel); log_h >= 9.  next_digests: (h/2) x 8 words.
bool fold_hash_eligible(int log_h) { return log_h >= 9 && getenv("TS_NO_FOLD_HASH") == nullptr; }
int fold_hash_launch(ts_ctx *c, const uint32_t *in, uint32_t *out, const uint32_t *addend, int log_h, const uint32_t beta_canon[4],
                     const uint32_t *half_beta_dev, uint32_t *next_digests, size_t first = 0, size_t h_local = 0) {
    const uint32_t half = bb::cinv(2);
    ef::E4 hb;
    for (int i = 0; i < 4; i++) hb.c[i] = beta_canon ? h_to_monty(bb::cmul(beta_canon[i], half)) : 0;
    if (!beta_canon && !half_beta_dev) TS_FAIL(c, TS_ERR_ARG, "fold: no beta");
    const uint32_t g_inv = bb::cinv(bb::two_adic_generator(log_h + 1));
    fold::InvRootPows rp;
    uint32_t r = g_inv;
    for (int k = 0; k < 28; k++) {
        rp.v[k] = h_to_monty(r);
        r = bb::cmul(r, r);
    }
    fold::ChunkDeltas dl;
    const int Lh = log_h - 9;  // chunk = 512 outputs = 256 leaves of the next tree
    for (int t = 0; t < 28; t++) {
        uint32_t v = 1;
        if (Lh > 0 && t < Lh) {
            int64_t delta = (int64_t)1 << (Lh - 1 - t);
            for (int k = 0; k < t; k++) delta -= (int64_t)1 << (Lh - 1 - k);
            const int64_t order = (int64_t)2 << log_h;
            delta = ((delta % order) + order) % order;
            v = bb::cpow(g_inv, (uint64_t)delta);
        }
        dl.v[t] = h_to_monty(v);
    }
    const size_t h = h_local ? h_local : (size_t)1 << log_h;
    if ((first | h) & 511) TS_FAIL(c, TS_ERR_ARG, "fold_hash: range must be a multiple of 512 rows");
    const size_t chunks = h >> 9;
    KScope ks(c, TS_K_FOLD);
    auto kfn = fold::fold_hash_kernel;
    TS_LAUNCH(kfn, (unsigned)std::min<size_t>(chunks, (size_t)c->num_sms * 8), fold::FOLD_T, 0, c->stream, (const uint4 *)in,
              (uint4 *)out, (const uint4 *)addend, log_h, first, h, hb, rp, dl, beta_canon ? (const uint32_t *)nullptr : half_beta_dev,
              next_digests);
    return check_launch(c, "fold_hash_kernel");
}

}  // namespace

// ================================================================================================ C ABI
// Pageable host matrices: the column-chunk pipeline stages them through page-locked bounce slots (stage_pageable_window).
// Page-locking the caller's buffer for the duration of the call instead is opt-in (TS_AUTO_PIN=1): measured on a 2^22 x 256
// trace it costs more than it saves (cudaHostRegister of 4 GiB ~ 1 s; commit 1.69 s against 0.73 s with the driver's own
// row-by-row staging).  A caller that keeps its trace for several calls registers it once itself (ts_host_register).
struct ScopedHostPin {
    const void *p = nullptr;
    ScopedHostPin(const void *host, size_t bytes) {
#ifndef TS_EMULATE
        if (!host || bytes < ((size_t)16 << 20) || !getenv("TS_AUTO_PIN")) return;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        if (at.type != cudaMemoryTypeUnregistered) return;
        if (cudaHostRegister(const_cast<void *>(host), bytes, cudaHostRegisterDefault) == cudaSuccess) p = host;
        else cudaGetLastError();
#else
        (void)host, (void)bytes;
#endif
    }
    ~ScopedHostPin() {
#ifndef TS_EMULATE
        if (p) cudaHostUnregister(const_cast<void *>(p));
#endif
    }
};

extern "C" {

int ts_is_device_build(void) {
#ifdef TS_EMULATE
    return 0;
#else
    return 1;
#endif
}

int ts_ctx_create(int device, void *stream, ts_ctx **out) {
    if (!out) return TS_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) return TS_ERR_CUDA;  // no GPU: fail loudly
    if (cudaSetDevice(device) != cudaSuccess) return TS_ERR_CUDA;
    ts_ctx *c = new ts_ctx;
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
#ifndef TS_EMULATE
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete c;
            return TS_ERR_CUDA;
        }
        c->own_stream = true;
#endif
    }
    // kernels that need more than 48 KiB of dynamic shared memory
    cudaFuncSetAttribute(ntt::ntt_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(ntt::ntt_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(ntt::lde_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(nttf::ntt_pass_fast_kernel<9, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(nttf::ntt_pass_fast_kernel<9, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(nttf::ntt_pass_fast_kernel<10, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(nttf::ntt_pass_fast_kernel<10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(nttf::ntt_pass_fast_kernel<11, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(nttf::ntt_pass_fast_kernel<11, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(nttp::ntt_pass_pm_kernel<9, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(nttp::ntt_pass_pm_kernel<9, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(nttp::ntt_pass_pm_kernel<10, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(nttp::ntt_pass_pm_kernel<10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(nttp::ntt_pass_pm_kernel<11, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(nttp::ntt_pass_pm_kernel<11, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_kernel<9, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_kernel<9, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_kernel<9, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_kernel<9, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_kernel<10, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_kernel<10, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_kernel<10, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_kernel<10, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_kernel<11, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_kernel<11, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_kernel<11, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_kernel<11, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
#ifndef TS_EMULATE
    cudaFuncSetAttribute(ntt4::pass_shfl_kernel<11, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_shfl_kernel<11, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_shfl_kernel<11, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_shfl_kernel<11, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt4::pass_tma_kernel<9, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    cudaFuncSetAttribute(ntt4::pass_tma_kernel<9, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    cudaFuncSetAttribute(ntt4::pass_tma_kernel<10, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    cudaFuncSetAttribute(ntt4::pass_tma_kernel<10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    cudaFuncSetAttribute(ntt4::pass_tma_kernel<11, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    cudaFuncSetAttribute(ntt4::pass_tma_kernel<11, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
#endif
    cudaFuncSetAttribute(nttp::ntt_pass_pm2_kernel<9, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaFuncSetAttribute(nttp::ntt_pass_pm2_kernel<9, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaFuncSetAttribute(nttp::ntt_pass_pm2_kernel<10, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaFuncSetAttribute(nttp::ntt_pass_pm2_kernel<10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaFuncSetAttribute(nttp::ntt_pass_pm2_kernel<11, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaFuncSetAttribute(nttp::ntt_pass_pm2_kernel<11, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaFuncSetAttribute(nttp::lde_mid_pm2_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    cudaFuncSetAttribute(nttp::lde_mid_pm2_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    cudaFuncSetAttribute(nttp::lde_mid_pm2_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    cudaFuncSetAttribute(nttp::lde_mid_pm_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
    cudaFuncSetAttribute(nttp::lde_mid_pm_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
    cudaFuncSetAttribute(nttp::lde_mid_pm_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
    cudaFuncSetAttribute(nttf::lde_mid_fast_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
    cudaFuncSetAttribute(nttf::lde_mid_fast_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
    cudaFuncSetAttribute(nttf::lde_mid_fast_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
    // small twiddle table w_4096^e and the fold's 256-entry low table, built once
    bool ok = cudaMalloc((void **)&c->tw_small, sizeof(uint2) << ntt::SMALL_LOG) == cudaSuccess &&
              cudaMalloc((void **)&c->tw_small_inv, sizeof(uint2) << ntt::SMALL_LOG) == cudaSuccess &&
              cudaMalloc((void **)&c->fold_tlo, 256 * 4) == cudaSuccess;
    if (ok) ok = gen_twiddles(c, c->tw_small, ntt::SMALL_LOG) == TS_OK;
    if (ok) ok = gen_twiddles(c, c->tw_small_inv, ntt::SMALL_LOG, true) == TS_OK;
    for (int dir = 0; dir < 2 && ok; dir++) {  // product-indexed inter-round tables (ntt_pm.cuh: rtw)
        ok = cudaMalloc((void **)&c->round_tab[dir], sizeof(uint2) * nttf::RT_ENTRIES) == cudaSuccess;
        const uint2 *src = dir ? c->tw_small_inv : c->tw_small;
        const int spec[5][3] = {{9, 8, nttf::RT_R1_OFF(9)}, {10, 8, nttf::RT_R1_OFF(10)}, {11, 8, nttf::RT_R1_OFF(11)},
                                {8, 5, nttf::RT_R2_OFF}, {5, 2, nttf::RT_R3_OFF}};
        for (int k = 0; k < 5 && ok; k++) {
            auto kfn = nttf::fill_round_table_kernel;
            TS_LAUNCH(kfn, ((1u << spec[k][0]) + 255) / 256, 256, 0, c->stream, c->round_tab[dir] + spec[k][2], src, spec[k][0],
                      spec[k][1], (int)ntt::SMALL_LOG);
            ok = check_launch(c, "fill_round_table_kernel") == TS_OK;
        }
    }
    if (ok) {
        uint32_t tlo[256];
        const uint32_t w512inv = bb::cinv(bb::two_adic_generator(9));
        for (uint32_t t = 0; t < 256; t++) {
            uint32_t r = 0;
            for (int k = 0; k < 8; k++) r |= ((t >> k) & 1) << (7 - k);
            tlo[t] = h_to_monty(bb::cpow(w512inv, r));
        }
        ok = cudaMemcpyAsync(c->fold_tlo, tlo, sizeof tlo, cudaMemcpyHostToDevice, c->stream) == cudaSuccess &&
             cudaStreamSynchronize(c->stream) == cudaSuccess;
    }
    if (!ok) {
        ts_ctx_destroy(c);
        return TS_ERR_CUDA;
    }
    *out = c;
    return TS_OK;
}

void ts_ctx_destroy(ts_ctx *c) {
    if (!c) return;
    cudaStreamSynchronize(c->stream);
    cudaFree(c->tw_small);
    cudaFree(c->tw_small_inv);
    cudaFree(c->round_tab[0]);
    cudaFree(c->round_tab[1]);
    cudaFree(c->tw_big);
    cudaFree(c->fold_tlo);
    cudaFree(c->scratch);
    cudaFree(c->stage[0]);
    cudaFree(c->stage[1]);
#ifndef TS_EMULATE
    if (c->side_stream) {
        cudaStreamSynchronize(c->side_stream);
        cudaEventDestroy(c->ev_side);
        cudaStreamDestroy(c->side_stream);
    }
    for (int lane = 0; lane < ts_ctx::XFER_LANES; lane++)
        if (c->xfer_stream[lane]) {
            cudaStreamSynchronize(c->xfer_stream[lane]);
            cudaEventDestroy(c->ev_xfer_in[lane]);
            cudaEventDestroy(c->ev_xfer_out[lane]);
            cudaStreamDestroy(c->xfer_stream[lane]);
        }
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        for (int s = 0; s < 2; s++) {
            cudaEventDestroy(c->ev_copy[s]);
            cudaEventDestroy(c->ev_free[s]);
        }
        cudaStreamDestroy(c->copy_stream);
    }
    if (c->bounce) {
        for (int t = 0; t < ts_ctx::BOUNCE_WORKERS; t++) {
            cudaStreamSynchronize(c->bounce_stream[t]);
            cudaEventDestroy(c->bounce_ev[t][0]);
            cudaEventDestroy(c->bounce_ev[t][1]);
            cudaEventDestroy(c->bounce_done[t]);
            cudaStreamDestroy(c->bounce_stream[t]);
        }
        cudaFreeHost(c->bounce);
    }
#endif
    for (auto &kv : c->v4_post1) cudaFree(kv.second);
    for (auto &kv : c->v4_mid) {
        pool_release(c, kv.second.post3);
        pool_release(c, kv.second.pre4);
    }
    pool_trim(c);
    for (auto &kv : c->coset_tabs) {
        cudaFree(kv.second.first);
        cudaFree(kv.second.second);
    }
    for (auto &p : c->pending) {
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
#ifndef TS_EMULATE
    if (c->own_stream) cudaStreamDestroy(c->stream);
#endif
    delete c;
}

const char *ts_last_error(const ts_ctx *c) { return c ? c->err.c_str() : "null context"; }
int ts_ctx_synchronize(ts_ctx *c) {
    TS_CUDA(c, cudaStreamSynchronize(c->stream));
    return TS_OK;
}
int ts_ctx_set_profiling(ts_ctx *c, int on) {
    c->profiling = on != 0;
    return TS_OK;
}
static void drain_pending(ts_ctx *c) {
    if (c->pending.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (auto &p : c->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) c->ms[p.kind] += ms;
        c->ev_pool.push_back(p.a);
        c->ev_pool.push_back(p.b);
    }
    c->pending.clear();
}
int ts_ctx_reset_stats(ts_ctx *c) {
    drain_pending(c);
    for (int i = 0; i < TS_K_COUNT; i++) {
        c->ms[i] = 0;
        c->launches[i] = 0;
    }
    c->total_launches = 0;
    return TS_OK;
}
int ts_ctx_get_stats(ts_ctx *c, int kind, double *ms, uint64_t *launches) {
    if (kind < 0 || kind >= TS_K_COUNT) TS_FAIL(c, TS_ERR_ARG, "bad kernel kind");
    drain_pending(c);
    if (ms) *ms = c->ms[kind];
    if (launches) *launches = c->launches[kind];
    return TS_OK;
}
uint64_t ts_ctx_total_launches(const ts_ctx *c) { return c->total_launches; }
int ts_ctx_trim(ts_ctx *c) {
    pool_trim(c);
    return TS_OK;
}

// ---------------------------------------------------------------- matrices
int ts_matrix_alloc(ts_ctx *c, size_t rows, size_t width, ts_matrix **out) { return new_matrix(c, rows, width, out); }
int ts_matrix_from_host(ts_ctx *c, const uint32_t *host, size_t rows, size_t width, ts_matrix **out) {
    TS_TRY(new_matrix(c, rows, width, out));
    TS_CUDA(c, cudaMemcpyAsync((*out)->d, host, rows * width * 4, cudaMemcpyHostToDevice, c->stream));
    return TS_OK;
}
int ts_matrix_from_device(ts_ctx *c, const uint32_t *dev, size_t rows, size_t width, ts_matrix **out) {
    TS_TRY(new_matrix(c, rows, width, out));
    TS_CUDA(c, cudaMemcpyAsync((*out)->d, dev, rows * width * 4, cudaMemcpyDeviceToDevice, c->stream));
    return TS_OK;
}
int ts_matrix_wrap_device(ts_ctx *c, uint32_t *dev, size_t rows, size_t width, ts_matrix **out) {
    *out = new ts_matrix{c, dev, rows, width, false};
    return TS_OK;
}
int ts_matrix_download(ts_ctx *c, const ts_matrix *m, size_t row0, size_t nrows, uint32_t *host) {
    if (row0 + nrows > m->rows) TS_FAIL(c, TS_ERR_ARG, "download: row range out of bounds");
    TS_CUDA(c, cudaMemcpyAsync(host, m->d + row0 * m->width, nrows * m->width * 4, cudaMemcpyDeviceToHost, c->stream));
    TS_CUDA(c, cudaStreamSynchronize(c->stream));
    return TS_OK;
}
uint32_t *ts_matrix_device_ptr(const ts_matrix *m) { return m->d; }
size_t ts_matrix_rows(const ts_matrix *m) { return m->rows; }
size_t ts_matrix_width(const ts_matrix *m) { return m->width; }
void ts_matrix_free(ts_matrix *m) {
    if (!m) return;
    if (m->owned && m->d) pool_release(m->ctx, m->d);
    delete m;
}
static int monty_convert(ts_ctx *c, ts_matrix *m, int to) {
    KScope ks(c, TS_K_MISC);
    const size_t n = m->rows * m->width;
    auto kfn = ntt::monty_convert_kernel;
    TS_LAUNCH(kfn, (unsigned)std::min<size_t>((n + 255) / 256, (size_t)c->num_sms * 16), 256, 0, c->stream, m->d, n, to);
    return check_launch(c, "monty_convert_kernel");
}
int ts_matrix_to_monty(ts_ctx *c, ts_matrix *m) { return monty_convert(c, m, 1); }
int ts_matrix_from_monty(ts_ctx *c, ts_matrix *m) { return monty_convert(c, m, 0); }
int ts_matrix_bit_reverse_rows(ts_ctx *c, const ts_matrix *m, ts_matrix **out) {
    const int lh = log2_strict(m->rows);
    if (lh < 0) TS_FAIL(c, TS_ERR_ARG, "bit_reverse_rows: height must be a power of two");
    TS_TRY(new_matrix(c, m->rows, m->width, out));
    return bitrev_rows(c, m->d, (*out)->d, lh, m->width);
}

// ---------------------------------------------------------------- DFT family
int ts_coset_lde_batch(ts_ctx *c, const ts_matrix *evals, unsigned added_bits, uint32_t shift_monty, int natural_order,
                       ts_matrix **out) {
    ts_matrix *o = nullptr;
    TS_TRY(new_matrix(c, evals->rows << added_bits, evals->width, &o));
    int rc = lde_committed(c, evals->d, evals->rows, evals->width, added_bits, shift_monty, o->d);
    if (rc == TS_OK && natural_order) {
        ts_matrix *nat = nullptr;
        rc = ts_matrix_bit_reverse_rows(c, o, &nat);
        ts_matrix_free(o);
        o = nat;
    }
    if (rc != TS_OK) {
        ts_matrix_free(o);
        return rc;
    }
    *out = o;
    return TS_OK;
}
int ts_lde_batch(ts_ctx *c, const ts_matrix *evals, unsigned added_bits, ts_matrix **out) {
    return ts_coset_lde_batch(c, evals, added_bits, bb::MONTY_ONE, 1, out);
}
static int plain_transform(ts_ctx *c, bool inverse, const ts_matrix *in, bool with_shift, uint32_t shift_monty,
                           ts_matrix **out) {
    const int m = log2_strict(in->rows);
    if (m < 0 || m > 27) TS_FAIL(c, TS_ERR_ARG, "dft: height must be a power of two <= 2^27");
    ts_matrix *work = nullptr;
    TS_TRY(ts_matrix_from_device(c, in->d, in->rows, in->width, &work));
    int rc = TS_OK;
    if (with_shift) {
        KScope ks(c, TS_K_MISC);
        const size_t total = in->rows * in->width;
        auto kfn = ntt::scale_rows_pow_kernel;
        TS_LAUNCH(kfn, (unsigned)std::min<size_t>((total + 255) / 256, (size_t)c->num_sms * 16), 256, 0, c->stream,
                  work->d, m, (uint32_t)in->width, shift_monty);
        rc = check_launch(c, "scale_rows_pow_kernel");
    }
    if (rc == TS_OK) rc = full_transform(c, inverse, work->d, work->d, in->width, m);
    ts_matrix *nat = nullptr;
    if (rc == TS_OK) rc = ts_matrix_bit_reverse_rows(c, work, &nat);
    ts_matrix_free(work);
    if (rc != TS_OK) return rc;
    *out = nat;
    return TS_OK;
}
int ts_dft_batch(ts_ctx *c, const ts_matrix *coeffs, ts_matrix **out) {
    return plain_transform(c, false, coeffs, false, 0, out);
}
int ts_idft_batch(ts_ctx *c, const ts_matrix *evals, ts_matrix **out) {
    return plain_transform(c, true, evals, false, 0, out);
}
int ts_coset_dft_batch(ts_ctx *c, const ts_matrix *coeffs, uint32_t shift_monty, ts_matrix **out) {
    return plain_transform(c, false, coeffs, true, shift_monty, out);
}
int ts_dft_batch_host(ts_ctx *c, int kind, const uint32_t *mat_host, size_t rows, size_t width, uint32_t shift_monty,
                      uint32_t *out_host) {
    if (!c || !mat_host || !out_host || kind < TS_DFT || kind > TS_COSET_DFT) TS_FAIL(c, TS_ERR_ARG, "dft_batch_host: bad argument");
    ts_matrix *in = nullptr, *out = nullptr;
    TS_TRY(ts_matrix_from_host(c, mat_host, rows, width, &in));
    int rc = plain_transform(c, kind == TS_IDFT, in, kind == TS_COSET_DFT, shift_monty, &out);
    ts_matrix_free(in);
    if (rc != TS_OK) return rc;
    rc = ts_matrix_download(c, out, 0, rows, out_host);
    ts_matrix_free(out);
    return rc;
}
int ts_host_register(ts_ctx *c, const void *host, size_t bytes) {
#ifndef TS_EMULATE
    TS_CUDA(c, cudaHostRegister(const_cast<void *>(host), bytes, cudaHostRegisterDefault));
#else
    (void)host, (void)bytes;
#endif
    return TS_OK;
}
int ts_host_unregister(ts_ctx *c, const void *host) {
#ifndef TS_EMULATE
    TS_CUDA(c, cudaHostUnregister(const_cast<void *>(host)));
#else
    (void)host;
#endif
    return TS_OK;
}
int ts_coset_lde_batch_host(ts_ctx *c, const uint32_t *evals_host, size_t rows, size_t width, unsigned added_bits,
                            uint32_t shift_monty, int natural_order, uint32_t *out_host) {
    ts_matrix *in = nullptr, *o = nullptr;
    ScopedHostPin pin_in(evals_host, rows * width * 4), pin_out(out_host, (rows << added_bits) * width * 4);
    int rc;
    if (!natural_order && log2_strict(rows) >= 0 && pipeline_eligible(rows, width)) {
        rc = new_matrix(c, rows << added_bits, width, &o);
        if (rc == TS_OK) rc = lde_from_host_pipelined(c, evals_host, rows, width, added_bits, shift_monty, o->d);
    } else {
        TS_TRY(ts_matrix_from_host(c, evals_host, rows, width, &in));
        rc = ts_coset_lde_batch(c, in, added_bits, shift_monty, natural_order, &o);
    }
    if (rc == TS_OK) rc = ts_matrix_download(c, o, 0, o->rows, out_host);
    ts_matrix_free(in);
    ts_matrix_free(o);
    return rc;
}

// ---------------------------------------------------------------- MMCS
int ts_mmcs_commit(ts_ctx *ctx, ts_matrix *const *mats, size_t n_mats, int layout, int take_ownership,
                   uint8_t root[32], ts_tree **out) {
    return mmcs_commit(ctx, mats, n_mats, layout, take_ownership, root, out, true);
}
size_t ts_tree_num_matrices(const ts_tree *t) { return t->mats.size(); }
ts_matrix *ts_tree_matrix(const ts_tree *t, size_t i) { return i < t->mats.size() ? t->mats[i] : nullptr; }
size_t ts_tree_depth(const ts_tree *t) { return t->lmax; }
size_t ts_tree_max_height(const ts_tree *t) { return t->hmax; }
int ts_tree_root_copy(ts_ctx *c, const ts_tree *t, uint8_t *dst_device) {
    TS_CUDA(c, cudaMemcpyAsync(dst_device, t->digests + t->layer_off[t->lmax] * 8, 32, cudaMemcpyDeviceToDevice, c->stream));
    return TS_OK;
}
namespace {
int open_many(ts_ctx *c, const ts_tree *t, const std::vector<size_t> &idx, std::vector<uint32_t> &rows, std::vector<uint8_t> &paths,
              size_t *row_words_out);  // defined with ts_pcs_open below
}
int ts_mmcs_open_batch(ts_ctx *c, const ts_tree *t, size_t index, uint32_t *rows_out, uint8_t *path_out) {
    // one gather launch and one read-back for the rows and the whole path (round 1: a 32-byte copy per tree level and per row)
    std::vector<uint32_t> rows;
    std::vector<uint8_t> path;
    size_t row_words = 0;
    TS_TRY(open_many(c, t, std::vector<size_t>{index}, rows, path, &row_words));
    if (row_words) memcpy(rows_out, rows.data(), row_words * 4);
    if (t->lmax) memcpy(path_out, path.data(), 32 * (size_t)t->lmax);
    return TS_OK;
}
int ts_tree_layer(ts_ctx *c, const ts_tree *t, size_t layer, uint8_t *out, size_t *n_nodes) {
    if (layer > t->lmax) TS_FAIL(c, TS_ERR_ARG, "tree_layer: no such layer");
    const size_t n = t->hmax >> layer;
    if (n_nodes) *n_nodes = n;
    if (out) {
        TS_CUDA(c, cudaMemcpyAsync(out, t->digests + t->layer_off[layer] * 8, n * 32, cudaMemcpyDeviceToHost, c->stream));
        TS_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return TS_OK;
}
void ts_tree_free(ts_tree *t) {
    if (!t) return;
    if (t->digests) pool_release(t->ctx, t->digests);
    for (ts_matrix *m : t->mats)
        if (t->own_mats) ts_matrix_free(m);
    delete t;
}
int ts_mmcs_verify_batch(const size_t *heights, const size_t *widths, size_t k, int layout, size_t index,
                         const uint32_t *rows_monty, const uint8_t *path, size_t depth, const uint8_t root[32]) {
    if (k == 0) return TS_ERR_ARG;
    std::vector<size_t> order(k), offs(k);
    size_t o = 0;
    for (size_t i = 0; i < k; i++) {
        order[i] = i;
        offs[i] = o;
        o += widths[i];
    }
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return heights[a] > heights[b]; });
    const size_t hmax = heights[order[0]];
    if (((size_t)1 << depth) != hmax) return TS_ERR_ARG;
    auto hash_group = [&](size_t q0, size_t q1, uint8_t out[32]) {
        std::vector<uint8_t> buf;
        for (size_t q = q0; q < q1; q++) {
            const size_t m = order[q];
            for (size_t cidx = 0; cidx < widths[m]; cidx++) {
                const uint32_t v = h_from_monty(rows_monty[offs[m] + cidx]);
                for (int b = 0; b < 4; b++) buf.push_back((uint8_t)(v >> (8 * b)));
            }
        }
        hostb3::hash(buf.data(), buf.size(), out);
    };
    size_t n_first = 0;
    if (layout == TS_LAYOUT_PADDED) n_first = k;
    else
        while (n_first < k && heights[order[n_first]] == hmax) n_first++;
    uint8_t cur[32], pair[64];
    hash_group(0, n_first, cur);
    size_t next = n_first, len = hmax;
    for (size_t l = 0; l < depth; l++) {
        if (index & 1) {
            memcpy(pair, path + 32 * l, 32);
            memcpy(pair + 32, cur, 32);
        } else {
            memcpy(pair, cur, 32);
            memcpy(pair + 32, path + 32 * l, 32);
        }
        hostb3::hash(pair, 64, cur);
        index >>= 1;
        len >>= 1;
        const size_t inj0 = next;
        if (layout == TS_LAYOUT_P3_INJECT)
            while (next < k && heights[order[next]] == len) next++;
        if (next > inj0) {
            memcpy(pair, cur, 32);
            hash_group(inj0, next, pair + 32);
            hostb3::hash(pair, 64, cur);
        }
    }
    return memcmp(cur, root, 32) == 0 ? TS_OK : TS_ERR_ARG;
}

// ---------------------------------------------------------------- challenger
int ts_challenger_new(ts_challenger **out) {
    *out = new ts_challenger;
    return TS_OK;
}
int ts_challenger_clone(const ts_challenger *c, ts_challenger **out) {
    *out = new ts_challenger(*c);
    return TS_OK;
}
void ts_challenger_free(ts_challenger *c) { delete c; }
void ts_challenger_observe(ts_challenger *c, const uint8_t word[4]) { c->observe(word); }
void ts_challenger_observe_digest(ts_challenger *c, const uint8_t digest[32]) {
    for (int i = 0; i < 8; i++) c->observe(digest + 4 * i);
}
uint32_t ts_challenger_sample_base(ts_challenger *c) { return c->sample_base(); }
void ts_challenger_sample_ext(ts_challenger *c, uint32_t out[4]) {
    for (int i = 0; i < 4; i++) out[i] = c->sample_base();
}
size_t ts_challenger_sample_bits(ts_challenger *c, unsigned bits, int ext) { return c->sample_bits(bits, ext != 0); }
int ts_challenger_check_witness(ts_challenger *c, unsigned bits, uint32_t witness, int ext) {
    return c->check_witness(bits, witness, ext != 0) ? 1 : 0;
}
int ts_challenger_grind(ts_challenger *c, unsigned bits, int ext, uint32_t *witness) {
    for (uint32_t w = 0; w < 4096; w++) {
        ts_challenger t(*c);
        if (t.check_witness(bits, w, ext != 0)) {
            c->check_witness(bits, w, ext != 0);
            *witness = w;
            return TS_OK;
        }
    }
    return TS_ERR_NO_WITNESS;
}

// ---------------------------------------------------------------- fold
int ts_fri_fold_ext(ts_ctx *c, const uint32_t *in_dev, size_t h, const uint32_t beta_monty[4], uint32_t *out_dev) {
    const int lh = log2_strict(h);
    if (lh < 0 || lh > 26) TS_FAIL(c, TS_ERR_ARG, "fold: h must be a power of two <= 2^26");
    uint32_t beta[4];
    for (int i = 0; i < 4; i++) beta[i] = h_from_monty(beta_monty[i]);
    return fold_ext_launch(c, in_dev, out_dev, nullptr, lh, beta);
}
int ts_fri_fold_base(ts_ctx *c, const uint32_t *in_dev, size_t h, uint32_t beta_monty, uint32_t *out_dev) {
    const int lh = log2_strict(h);
    if (lh < 0 || lh > 26) TS_FAIL(c, TS_ERR_ARG, "fold: h must be a power of two <= 2^26");
    const uint32_t hb = h_to_monty(bb::cmul(h_from_monty(beta_monty), bb::cinv(2)));
    fold::InvRootPows rp;
    uint32_t r = bb::cinv(bb::two_adic_generator(lh + 1));
    for (int k = 0; k < 28; k++) {
        rp.v[k] = h_to_monty(r);
        r = bb::cmul(r, r);
    }
    KScope ks(c, TS_K_FOLD);
    auto kfn = fold::fold_base_kernel;
    TS_LAUNCH(kfn, (unsigned)std::min<size_t>((h + 255) / 256, (size_t)c->num_sms * 8), fold::FOLD_T, 0, c->stream,
              (const uint2 *)in_dev, out_dev, lh, hb, rp);
    return check_launch(c, "fold_base_kernel");
}
int ts_fri_fold_ext_host(ts_ctx *c, const uint32_t *in_host, size_t h, const uint32_t beta_monty[4],
                         uint32_t *out_host) {
    ts_matrix *in = nullptr, *o = nullptr;
    TS_TRY(ts_matrix_from_host(c, in_host, 2 * h, 4, &in));
    int rc = new_matrix(c, h, 4, &o);
    if (rc == TS_OK) rc = ts_fri_fold_ext(c, in->d, h, beta_monty, o->d);
    if (rc == TS_OK) rc = ts_matrix_download(c, o, 0, h, out_host);
    ts_matrix_free(in);
    ts_matrix_free(o);
    return rc;
}

// ---------------------------------------------------------------- commit phase (fri/src/prover.rs:93-141)
// The remaining commit-phase rounds of a small layer in one launch (fri_tail.cuh).  `cur` (len x 4 words) is the current
// layer; when it is owned it becomes the leaf matrix of the first tail tree (or is released), otherwise it is only read.
// Appends to commits / trees from index `round` on and advances it; the final layer lands in c->tail_final.
static int fri_tail(ts_ctx *c, ts_challenger *chal, uint32_t *cur, bool cur_owned, size_t len, unsigned log_blowup,
                    uint8_t *commits, ts_tree **trees, size_t &round, const uint32_t *prev_h_dev = nullptr,
                    const std::function<void()> *before_replay = nullptr) {
    const int log_len = log2_strict(len);
    const int rounds = log_len - (int)log_blowup;
    ftail::Params p;
    p.layer0 = cur;
    p.log_len0 = log_len;
    p.rounds = rounds;
    memcpy(p.prev_h, &chal->state[8][0], 32);
    p.prev_h_dev = prev_h_dev;
    for (int b = 0; b < ftail::MAX_LOG_LEN + 2; b++) p.inv_gen[b] = h_to_monty(bb::cinv(bb::two_adic_generator(b)));
    std::vector<ts_tree *> tt((size_t)rounds, nullptr);
    uint32_t *roots_dev = nullptr;
    int rc = TS_OK;
    cudaError_t e = pool_alloc(c, (void **)&roots_dev, (size_t)rounds * 32);
    if (e != cudaSuccess) rc = TS_ERR_CUDA;
    uint32_t *layer_in = cur;
    bool layer_owned = cur_owned;
    for (int r = 0; r < rounds && rc == TS_OK; r++) {
        const size_t h = len >> (r + 1);
        ts_tree *t = new ts_tree;
        t->ctx = c;
        t->mats.push_back(new ts_matrix{c, layer_in, h, 8, layer_owned});  // leaves of round r (prover.rs:112)
        t->own_mats = true;
        t->layout = TS_LAYOUT_P3_INJECT;
        t->order.assign(1, 0);
        t->hmax = h;
        t->lmax = (unsigned)log2_strict(h);
        t->digests = nullptr;
        size_t off = 0;
        for (unsigned l = 0; l <= t->lmax; l++) {
            t->layer_off.push_back(off);
            off += h >> l;
        }
        tt[(size_t)r] = t;
        uint32_t *next = nullptr;
        if (pool_alloc(c, (void **)&t->digests, off * 32) != cudaSuccess || pool_alloc(c, (void **)&next, h * 16) != cudaSuccess) {
            if (next) pool_release(c, next);
            rc = TS_ERR_CUDA;
            layer_in = nullptr;
            break;
        }
        p.digests[r] = t->digests;
        p.layers[r] = next;
        layer_in = next;  // leaves of the next round, owned by that round's tree
        layer_owned = true;
    }
    p.roots_out = roots_dev;
    std::vector<uint32_t> host((size_t)rounds * 8 + ((size_t)4 << log_blowup));
    if (rc == TS_OK) {
        {
            KScope ks(c, TS_K_FOLD);
            auto kfn = ftail::fri_tail_kernel;
            TS_LAUNCH(kfn, 1, ftail::NT, 12 * sizeof(uint32_t), c->stream, p);
            rc = check_launch(c, "fri_tail_kernel");
        }
        if (rc == TS_OK) {
            e = cudaMemcpyAsync(host.data(), roots_dev, (size_t)rounds * 32, cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(host.data() + (size_t)rounds * 8, layer_in, (size_t)16 << log_blowup, cudaMemcpyDeviceToHost,
                                    c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) {
                c->err = std::string("fri tail read-back: ") + cudaGetErrorString(e);
                rc = TS_ERR_CUDA;
            }
        }
    } else {
        c->err = "fri tail: allocation failed";
    }
    if (layer_in) pool_release(c, layer_in);  // the final layer is owned by no tree
    if (roots_dev) pool_release(c, roots_dev);
    if (rc == TS_OK && before_replay) (*before_replay)();  // the chained rounds before the tail are replayed first
    if (rc == TS_OK) {
        for (int r = 0; r < rounds; r++) {  // replay the sponge on the host: it stays the source of truth
            const uint8_t *root = reinterpret_cast<const uint8_t *>(host.data() + (size_t)r * 8);
            memcpy(commits + 32 * round, root, 32);
            ts_challenger_observe_digest(chal, root);
            uint32_t beta[4];
            ts_challenger_sample_ext(chal, beta);
            if (trees) trees[round] = tt[(size_t)r];
            else ts_tree_free(tt[(size_t)r]);
            tt[(size_t)r] = nullptr;
            round++;
        }
        c->tail_final.assign(host.begin() + (size_t)rounds * 8, host.end());
    }
    // error paths: trees that were set up free their layers; if not even the first one exists, the caller's owned layer
    // has no owner yet
    if (rc != TS_OK && cur_owned && tt[0] == nullptr) pool_release(c, cur);
    for (ts_tree *t : tt)
        if (t) ts_tree_free(t);
    return rc;
}

int ts_fri_commit_phase(ts_ctx *c, ts_matrix *const *inputs, size_t n_inputs, unsigned log_blowup, ts_challenger *chal,
                        uint8_t *commits, ts_tree **trees, uint32_t final_poly[4], size_t *rounds_out) {
    if (n_inputs == 0) TS_FAIL(c, TS_ERR_ARG, "commit phase: no inputs");
    for (size_t i = 0; i < n_inputs; i++) {
        if (inputs[i]->width != 4 || log2_strict(inputs[i]->rows) < 0)
            TS_FAIL(c, TS_ERR_ARG, "commit phase: inputs must be EF vectors (width 4) of power-of-two length");
        if (i > 0 && inputs[i]->rows >= inputs[i - 1]->rows)
            TS_FAIL(c, TS_ERR_ARG, "commit phase: inputs must be sorted by strictly descending length");
    }
    size_t len = inputs[0]->rows, next_in = 1, round = 0;
    uint32_t *next_digests = nullptr;  // digest buffer of the next round's tree, leaf layer filled by fold_hash_kernel
    // current layer.  When prover data is kept, round 0 commits to a clone of the first input
    // (`folded.clone()`, prover.rs:112) so every returned tree owns its layer; otherwise it is borrowed.
    uint32_t *cur = inputs[0]->d;
    bool cur_owned = false;
    const size_t blowup = (size_t)1 << log_blowup;
    int rc = TS_OK;
    if (trees && len > blowup) {
        uint32_t *cl = nullptr;
        TS_CUDA(c, pool_alloc(c, (void **)&cl, len * 16));
        cudaError_t e = cudaMemcpyAsync(cl, cur, len * 16, cudaMemcpyDeviceToDevice, c->stream);
        if (e != cudaSuccess) {
            pool_release(c, cl);
            TS_FAIL(c, TS_ERR_CUDA, std::string("clone first input: ") + cudaGetErrorString(e));
        }
        cur = cl;
        cur_owned = true;
    }
    // Chained rounds: while whole digests are all the challenger observes, its state after a round is (root, previous
    // squeeze), so the sponge step runs on the device (fri_tail.cuh: sponge_step_kernel) and leaves beta/2 where the fold
    // kernel reads it: no round waits for a 32-byte read-back and a host hash.  The roots come back ONCE, with the final
    // layer, and the host challenger replays them (it stays the source of truth for grinding and the query phase).
    const bool chain = chal->in_buf.empty() && getenv("TS_NO_FRI_CHAIN") == nullptr;
    const size_t max_rounds = (size_t)std::max(log2_strict(len) - (int)log_blowup, 0);
    uint32_t *d_chain = nullptr;  // [0,8): h  [8,12): beta/2  [12, 12 + 8*max_rounds): roots
    std::vector<uint32_t> chain_roots;
    size_t chain_pending = 0, chain_first = 0;
    if (chain && max_rounds > 0) {
        TS_CUDA(c, pool_alloc(c, (void **)&d_chain, (12 + 8 * max_rounds) * 4));
        uint32_t h0[8];
        memcpy(h0, &chal->state[8][0], 32);
        cudaError_t e = cudaMemcpyAsync(d_chain, h0, 32, cudaMemcpyHostToDevice, c->stream);  // pageable source: staged before return
        if (e != cudaSuccess) {
            pool_release(c, d_chain);
            TS_FAIL(c, TS_ERR_CUDA, std::string("chain init: ") + cudaGetErrorString(e));
        }
        chain_roots.resize(8 * max_rounds);
    }
    auto replay_chain = [&]() {  // after a synchronisation: the host challenger catches up with the chained rounds
        for (size_t k = 0; k < chain_pending; k++) {
            const uint8_t *root = reinterpret_cast<const uint8_t *>(chain_roots.data() + 8 * k);
            memcpy(commits + 32 * (chain_first + k), root, 32);
            ts_challenger_observe_digest(chal, root);  // prover.rs:114
            uint32_t beta[4];
            ts_challenger_sample_ext(chal, beta);      // prover.rs:116
        }
        chain_pending = 0;
    };
    while (len > blowup) {
        if (len <= ((size_t)1 << ftail::MAX_LOG_LEN) && next_in >= n_inputs && chal->in_buf.empty() &&
            getenv("TS_NO_FRI_TAIL") == nullptr) {
            // every remaining round in one launch, sponge included (fri_tail.cuh)
            const std::function<void()> hook = replay_chain;
            if (chain_pending) {
                cudaError_t e = cudaMemcpyAsync(chain_roots.data(), d_chain + 12 + 8 * chain_first, chain_pending * 32,
                                                cudaMemcpyDeviceToHost, c->stream);
                if (e != cudaSuccess) {
                    c->err = std::string("chain roots download: ") + cudaGetErrorString(e);
                    rc = TS_ERR_CUDA;
                    break;
                }
            }
            rc = fri_tail(c, chal, cur, cur_owned, len, log_blowup, commits, trees, round, d_chain, &hook);
            cur = nullptr;  // consumed: owned by the first tail tree or released
            cur_owned = false;
            if (rc == TS_OK) {
                // the tail left the final layer in c->tail_final (host)
                len = blowup;
                for (int i = 0; i < 4; i++) final_poly[i] = h_from_monty(c->tail_final[i]);
                for (size_t i = 1; i < blowup; i++)
                    if (memcmp(&c->tail_final[4 * i], &c->tail_final[0], 16) != 0) {  // prover.rs:130-134
                        c->err = "commit phase: final layer is not constant (input not low-degree)";
                        rc = TS_ERR_NOT_CONSTANT;
                    }
            }
            if (d_chain) pool_release(c, d_chain);
            if (rounds_out) *rounds_out = round;
            return rc;
        }
        const size_t h = len / 2;
        // leaves = RowMajorMatrix::new(folded.clone(), 2): h rows of 2 EF = 8 u32 (prover.rs:112)
        ts_matrix *leaves = new ts_matrix{c, cur, h, 8, cur_owned};
        ts_tree *tree = nullptr;
        uint8_t root[32];
        uint32_t *filled = next_digests;  // leaf layer already hashed by the previous round's fold_hash_kernel
        next_digests = nullptr;
        rc = mmcs_commit(c, &leaves, 1, TS_LAYOUT_P3_INJECT, 1, d_chain ? nullptr : root, &tree, true, nullptr, filled);  // prover.rs:113
        if (rc != TS_OK) {
            ts_matrix_free(leaves);
            cur = nullptr;
            break;
        }
        // the tree now owns `leaves` (and the layer buffer when it was ours)
        uint32_t beta[4];
        if (d_chain) {
            if (chain_pending == 0) chain_first = round;
            KScope ks(c, TS_K_TREE);
            auto kfn = ftail::sponge_step_kernel;
            TS_LAUNCH(kfn, 1, 32, 0, c->stream, (const uint32_t *)(tree->digests + tree->layer_off[tree->lmax] * 8), 1, d_chain,
                      d_chain + 12 + 8 * round, d_chain + 8);
            rc = check_launch(c, "sponge_step_kernel");
            chain_pending++;
        } else {
            memcpy(commits + 32 * round, root, 32);
            ts_challenger_observe_digest(chal, root);  // prover.rs:114
            ts_challenger_sample_ext(chal, beta);      // prover.rs:116
        }
        uint32_t *nf = nullptr;
        cudaError_t e = rc == TS_OK ? pool_alloc(c, (void **)&nf, h * 16) : cudaSuccess;
        if (rc != TS_OK) {
        } else if (e != cudaSuccess) {
            c->err = std::string("cudaMalloc folded: ") + cudaGetErrorString(e);
            rc = TS_ERR_CUDA;
        } else {
            const uint32_t *addend = nullptr;
            if (next_in < n_inputs && inputs[next_in]->rows == h) addend = inputs[next_in++]->d;  // prover.rs:124-126
            // the next iteration commits to nf as h/2 rows unless it is the last one or goes to the tail kernel
            const bool next_is_tail = h <= ((size_t)1 << ftail::MAX_LOG_LEN) && next_in >= n_inputs && getenv("TS_NO_FRI_TAIL") == nullptr;
            if (h > blowup && !next_is_tail && fold_hash_eligible(log2_strict(h)) &&
                pool_alloc(c, (void **)&next_digests, (h - 1) * 32) == cudaSuccess) {
                rc = fold_hash_launch(c, cur, nf, addend, log2_strict(h), d_chain ? nullptr : beta, d_chain ? d_chain + 8 : nullptr,
                                      next_digests);  // prover.rs:119 + the row hashes of :112-113
            } else {
                next_digests = nullptr;
                rc = fold_ext_launch(c, cur, nf, addend, log2_strict(h), d_chain ? nullptr : beta, 0, 0, d_chain ? d_chain + 8 : nullptr);  // prover.rs:119
            }
        }
        cur = nullptr;
        cur_owned = false;
        if (rc != TS_OK || !trees) ts_tree_free(tree);  // the layer returns to the stream-ordered pool
        else trees[round] = tree;
        if (rc != TS_OK) {
            if (nf) pool_release(c, nf);
            break;
        }
        cur = nf;
        cur_owned = true;
        len = h;
        round++;
    }
    if (rc == TS_OK) {
        std::vector<uint32_t> tail(len * 4);
        cudaError_t e = cudaMemcpyAsync(tail.data(), cur, len * 16, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess && chain_pending)
            e = cudaMemcpyAsync(chain_roots.data(), d_chain + 12 + 8 * chain_first, chain_pending * 32, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            c->err = std::string("final layer download: ") + cudaGetErrorString(e);
            rc = TS_ERR_CUDA;
        } else {
            replay_chain();
            for (int i = 0; i < 4; i++) final_poly[i] = h_from_monty(tail[i]);
            for (size_t i = 1; i < len; i++)
                if (memcmp(&tail[4 * i], &tail[0], 16) != 0) {  // prover.rs:130-134
                    c->err = "commit phase: final layer is not constant (input not low-degree)";
                    rc = TS_ERR_NOT_CONSTANT;
                }
        }
    }
    if (d_chain) pool_release(c, d_chain);
    if (next_digests) pool_release(c, next_digests);
    if (cur_owned && cur) pool_release(c, cur);
    if (rounds_out) *rounds_out = round;
    return rc;
}

// ---------------------------------------------------------------- PCS
int ts_pcs_commit(ts_ctx *c, ts_matrix *const *evals, const uint32_t *domain_shifts_monty, size_t n,
                  unsigned log_blowup, int layout, uint8_t root[32], ts_tree **out) {
    std::vector<ts_matrix *> ldes;
    int rc = TS_OK;
    for (size_t i = 0; i < n && rc == TS_OK; i++) {
        // shift = Val::generator() / domain.shift   (two_adic_pcs.rs:235)
        const uint32_t dshift = h_from_monty(domain_shifts_monty[i]);
        if (dshift == 0) {
            c->err = "pcs commit: zero domain shift";
            rc = TS_ERR_ARG;
            break;
        }
        const uint32_t shift = h_to_monty(bb::cmul(31, bb::cinv(dshift)));
        ts_matrix *lde = nullptr;
        if (n == 1 && log2_strict(evals[i]->rows) >= 0 && overlap_hash_eligible(evals[i]->rows, evals[i]->width)) {
            rc = new_matrix(c, evals[i]->rows << log_blowup, evals[i]->width, &lde);
            if (rc != TS_OK) break;
            const ts_matrix *ev = evals[i];
            const std::function<int(uint32_t *)> fill = [&](uint32_t *leaf_digests) {
                return lde_hash_overlapped(c, ev->d, ev->rows, ev->width, log_blowup, shift, lde->d, leaf_digests);
            };
            rc = mmcs_commit(c, &lde, 1, layout, 1, root, out, true, &fill);
            if (rc != TS_OK) ts_matrix_free(lde);
            return rc;
        }
        rc = ts_coset_lde_batch(c, evals[i], log_blowup, shift, 0, &lde);
        if (rc == TS_OK) ldes.push_back(lde);
    }
    if (rc == TS_OK) rc = mmcs_commit(c, ldes.data(), ldes.size(), layout, 1, root, out, true);
    if (rc != TS_OK)
        for (ts_matrix *m : ldes) ts_matrix_free(m);
    return rc;
}
int ts_pcs_commit_host(ts_ctx *c, const uint32_t *const *evals_host, const size_t *rows, const size_t *widths,
                       const uint32_t *domain_shifts_monty, size_t n, unsigned log_blowup, int layout,
                       uint8_t root[32], ts_tree **out) {
    std::vector<ts_matrix *> ldes;
    std::vector<std::unique_ptr<ScopedHostPin>> pins;
    for (size_t i = 0; i < n; i++) pins.emplace_back(new ScopedHostPin(evals_host[i], rows[i] * widths[i] * 4));
    int rc = TS_OK;
    for (size_t i = 0; i < n && rc == TS_OK; i++) {
        const uint32_t dshift = h_from_monty(domain_shifts_monty[i]);
        if (dshift == 0 || log2_strict(rows[i]) < 0) {
            c->err = "pcs commit: bad domain";
            rc = TS_ERR_ARG;
            break;
        }
        const uint32_t shift = h_to_monty(bb::cmul(31, bb::cinv(dshift)));  // two_adic_pcs.rs:235
        ts_matrix *lde = nullptr;
        if (n == 1 && pipeline_eligible(rows[i], widths[i]) && incremental_hash_eligible(widths[i])) {
            // one matrix: rows are hashed chunk by chunk behind the copy, the tree is finished when the copy ends
            rc = new_matrix(c, rows[i] << log_blowup, widths[i], &lde);
            if (rc != TS_OK) break;
            const std::function<int(uint32_t *)> fill = [&](uint32_t *leaf_digests) {
                return lde_from_host_pipelined(c, evals_host[i], rows[i], widths[i], log_blowup, shift, lde->d, leaf_digests);
            };
            rc = mmcs_commit(c, &lde, 1, layout, 1, root, out, true, &fill);
            if (rc != TS_OK) ts_matrix_free(lde);
            return rc;
        }
        if (pipeline_eligible(rows[i], widths[i])) {
            rc = new_matrix(c, rows[i] << log_blowup, widths[i], &lde);
            if (rc == TS_OK) rc = lde_from_host_pipelined(c, evals_host[i], rows[i], widths[i], log_blowup, shift, lde->d);
        } else {
            ts_matrix *in = nullptr;
            rc = ts_matrix_from_host(c, evals_host[i], rows[i], widths[i], &in);
            if (rc == TS_OK) rc = ts_coset_lde_batch(c, in, log_blowup, shift, 0, &lde);
            ts_matrix_free(in);
        }
        if (rc == TS_OK) ldes.push_back(lde);
        else ts_matrix_free(lde);
    }
    if (rc == TS_OK) rc = mmcs_commit(c, ldes.data(), ldes.size(), layout, 1, root, out, true);
    if (rc != TS_OK)
        for (ts_matrix *m : ldes) ts_matrix_free(m);
    return rc;
}
int ts_pcs_get_evaluations_on_domain(ts_ctx *c, const ts_tree *t, size_t idx, size_t domain_size, uint32_t *out_host) {
    if (idx >= t->mats.size()) TS_FAIL(c, TS_ERR_ARG, "get_evaluations_on_domain: no such matrix");
    const ts_matrix *lde = t->mats[idx];
    const int ld = log2_strict(domain_size);
    if (ld < 0 || domain_size > lde->rows) TS_FAIL(c, TS_ERR_ARG, "get_evaluations_on_domain: bad domain size");
    ts_matrix *tmp = nullptr;
    TS_TRY(new_matrix(c, domain_size, lde->width, &tmp));
    int rc = bitrev_rows(c, lde->d, tmp->d, ld, lde->width);  // split_rows(size).0.bit_reverse_rows()
    if (rc == TS_OK) rc = ts_matrix_download(c, tmp, 0, domain_size, out_host);
    ts_matrix_free(tmp);
    return rc;
}
// apow_dev: device array of alpha^(k) for the columns of m (Montgomery EF, one uint4 per column), readable for
// roundup16(width) entries (columns past the width are multiplied by zero-filled data)
static int dot_ext_powers_launch(ts_ctx *c, const ts_matrix *m, const uint32_t *apow_dev, ts_matrix *o, int accumulate) {
    KScope ks(c, TS_K_MISC);
    if ((m->width == 4 || m->width == 8) && getenv("TS_NO_FAST") == nullptr) {
        auto kfn = fold::dot_rows_small_kernel;
        TS_LAUNCH(kfn, (unsigned)std::min<size_t>((m->rows + 255) / 256, (size_t)c->num_sms * 16), 256, 0, c->stream,
                  (const uint4 *)m->d, m->rows, (uint32_t)(m->width / 4), (const uint4 *)apow_dev, (uint4 *)o->d, accumulate);
        return check_launch(c, "dot_rows_small_kernel");
    }
    if ((m->width & 3) == 0 && getenv("TS_NO_FAST") == nullptr) {
        auto kfn = fold::dot_rows_fast_kernel;
        const size_t per_block = (size_t)fold::DOT_FAST_WARPS * 32;
        fold::DotSegs sg;
        for (int i = 0; i < fold::DOT_MAX_SEG; i++) sg.ptr[i] = i == 0 ? m->d : nullptr;
        sg.n = 1;
        sg.seg_w = (uint32_t)m->width;
        sg.log_seg_w = 0;
        TS_LAUNCH(kfn, (unsigned)((m->rows + per_block - 1) / per_block), fold::DOT_FAST_WARPS * 32,
                  fold::DOT_SMEM_BYTES, c->stream, sg, m->rows, (uint32_t)m->width,
                  (const uint4 *)apow_dev, (uint4 *)o->d, accumulate);
        return check_launch(c, "dot_rows_fast_kernel");
    }
    auto kfn = fold::dot_ext_powers_kernel;
    TS_LAUNCH(kfn, (unsigned)((m->rows + fold::DOT_ROWS - 1) / fold::DOT_ROWS), 256,
              (size_t)fold::DOT_ROWS * (fold::DOT_COLS + 1) * 4, c->stream, (const uint32_t *)m->d, m->rows,
              (uint32_t)m->width, (const uint4 *)apow_dev, (uint4 *)o->d, accumulate);
    return check_launch(c, "dot_ext_powers_kernel");
}
// alpha^0 .. alpha^(count-1) followed by 16 zero entries, as a (count+16) x 4 device matrix
int ts_alpha_powers(ts_ctx *c, const uint32_t alpha_monty[4], size_t count, ts_matrix **out) {
    std::vector<uint32_t> apow((count + 16) * 4, 0u);
    uint32_t a[4], cur[4] = {1, 0, 0, 0};
    for (int i = 0; i < 4; i++) a[i] = h_from_monty(alpha_monty[i]);
    for (size_t k = 0; k < count; k++) {
        for (int i = 0; i < 4; i++) apow[4 * k + i] = h_to_monty(cur[i]);
        uint32_t nx[4];
        h_ef_mul(cur, a, nx);
        memcpy(cur, nx, 16);
    }
    ts_matrix *ap = nullptr;
    TS_TRY(ts_matrix_from_host(c, apow.data(), count + 16, 4, &ap));
    cudaError_t e = cudaStreamSynchronize(c->stream);  // apow is a stack-lifetime pageable buffer
    if (e != cudaSuccess) {
        ts_matrix_free(ap);
        TS_FAIL(c, TS_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = ap;
    return TS_OK;
}
int ts_dot_ext_powers(ts_ctx *c, const ts_matrix *m, const uint32_t alpha_monty[4], ts_matrix **out) {
    ts_matrix *o = nullptr, *ap = nullptr;
    TS_TRY(ts_alpha_powers(c, alpha_monty, m->width, &ap));
    int rc = new_matrix(c, m->rows, 4, &o);
    if (rc == TS_OK) rc = dot_ext_powers_launch(c, m, ap->d, o, 0);
    ts_matrix_free(ap);  // returns to the stream-ordered pool: safe while the kernel is still queued
    if (rc != TS_OK) {
        ts_matrix_free(o);
        return rc;
    }
    *out = o;
    return TS_OK;
}
int ts_dot_ext_powers_acc(ts_ctx *c, const ts_matrix *m, const ts_matrix *alpha_powers, size_t first_power,
                          ts_matrix *acc, int accumulate) {
    if (acc->rows != m->rows || acc->width != 4) TS_FAIL(c, TS_ERR_ARG, "dot_ext_powers_acc: acc must be rows x 4");
    if (alpha_powers->width != 4 || first_power + ((m->width + 15) & ~(size_t)15) > alpha_powers->rows)
        TS_FAIL(c, TS_ERR_ARG, "dot_ext_powers_acc: alpha_powers too short (use ts_alpha_powers(total_width))");
    return dot_ext_powers_launch(c, m, alpha_powers->d + 4 * first_power, acc, accumulate);
}
int ts_dot_ext_powers_blocks(ts_ctx *c, ts_matrix *const *blocks, size_t n_blocks, const ts_matrix *alpha_powers,
                             ts_matrix *acc) {
    if (n_blocks == 0) TS_FAIL(c, TS_ERR_ARG, "dot_ext_powers_blocks: no blocks");
    size_t total = 0;
    bool one_launch = n_blocks <= (size_t)fold::DOT_MAX_SEG && getenv("TS_NO_FAST") == nullptr;
    for (size_t i = 0; i < n_blocks; i++) {
        if (blocks[i]->rows != acc->rows) TS_FAIL(c, TS_ERR_ARG, "dot_ext_powers_blocks: row counts differ");
        one_launch = one_launch && blocks[i]->width == blocks[0]->width;
        total += blocks[i]->width;
    }
    const size_t w0 = blocks[0]->width;
    one_launch = one_launch && w0 >= 4 && (w0 & (w0 - 1)) == 0;
    if (acc->width != 4) TS_FAIL(c, TS_ERR_ARG, "dot_ext_powers_blocks: acc must be rows x 4");
    if (alpha_powers->width != 4 || ((total + 15) & ~(size_t)15) > alpha_powers->rows)
        TS_FAIL(c, TS_ERR_ARG, "dot_ext_powers_blocks: alpha_powers too short (use ts_alpha_powers(total_width))");
    if (!one_launch) {  // ragged blocks: one accumulating pass per block
        size_t first = 0;
        for (size_t i = 0; i < n_blocks; i++) {
            TS_TRY(ts_dot_ext_powers_acc(c, blocks[i], alpha_powers, first, acc, i > 0));
            first += blocks[i]->width;
        }
        return TS_OK;
    }
    KScope ks(c, TS_K_MISC);
    fold::DotSegs sg;
    for (int i = 0; i < fold::DOT_MAX_SEG; i++) sg.ptr[i] = i < (int)n_blocks ? blocks[i]->d : nullptr;
    sg.n = (int)n_blocks;
    sg.seg_w = (uint32_t)w0;
    sg.log_seg_w = log2_strict(w0);
    auto kfn = fold::dot_rows_fast_kernel;
    const size_t per_block = (size_t)fold::DOT_FAST_WARPS * 32;
    TS_LAUNCH(kfn, (unsigned)((acc->rows + per_block - 1) / per_block), fold::DOT_FAST_WARPS * 32,
              fold::DOT_SMEM_BYTES, c->stream, sg, acc->rows, (uint32_t)total, (const uint4 *)alpha_powers->d,
              (uint4 *)acc->d, 0);
    return check_launch(c, "dot_rows_fast_kernel");
}

// ---------------------------------------------------------------- reduced openings (TwoAdicFriPcs::open, f1)
static opn::RootPows root_pows(int log_h) {
    opn::RootPows rp;
    uint32_t r = bb::two_adic_generator(log_h);
    for (int k = 0; k < 28; k++) {
        rp.v[k] = h_to_monty(r);
        r = bb::cmul(r, r);
    }
    return rp;
}
static ef::E4 to_e4(const uint32_t v[4]) {
    ef::E4 e;
    for (int i = 0; i < 4; i++) e.c[i] = v[i];
    return e;
}
// ---------------------------------------------------------------- quotient values (uni_stark::prove, f3)
int ts_quotient_values(ts_ctx *c, const ts_matrix *trace_lde, unsigned log_n, unsigned log_quotient_degree,
                       const uint32_t *program, size_t n_instr, const uint32_t *consts_monty, size_t n_consts,
                       const uint32_t *public_values_monty, size_t n_public, const uint32_t alpha_monty[4],
                       ts_matrix **chunks_out) {
    const unsigned log_m = log_n + log_quotient_degree;
    const int log_N = log2_strict(trace_lde->rows);
    if (log_N < 0 || log_m > (unsigned)log_N || log_m > 27)
        TS_FAIL(c, TS_ERR_ARG, "quotient_values: the quotient domain must fit inside the committed LDE");
    if (log_quotient_degree > 4 || (1u << log_quotient_degree) > (unsigned)quo::MAX_ZH)
        TS_FAIL(c, TS_ERR_ARG, "quotient_values: quotient degree too large");
    // validate the program on the host: a bad index must be an error, not an out-of-bounds access
    for (size_t pc = 0; pc < n_instr; pc++) {
        const uint32_t op = program[4 * pc], dst = program[4 * pc + 1];
        if (op > quo::OP_ASSERT_ZERO) TS_FAIL(c, TS_ERR_ARG, "quotient_values: unknown opcode");
        if (op != quo::OP_ASSERT_ZERO && dst >= (uint32_t)quo::MAX_REGS) TS_FAIL(c, TS_ERR_ARG, "quotient_values: register index");
        const int n_ops = (op == quo::OP_ASSERT_ZERO || op == quo::OP_NEG) ? 1 : 2;
        for (int k = 0; k < n_ops; k++) {
            const uint32_t code = program[4 * pc + 2 + k], kind = code >> 28, idx = code & 0x0fffffffu;
            const size_t lim = kind == quo::K_REG ? (size_t)quo::MAX_REGS
                               : (kind == quo::K_LOCAL || kind == quo::K_NEXT) ? trace_lde->width
                               : kind == quo::K_PUBLIC ? n_public
                               : kind == quo::K_CONST ? n_consts
                               : kind == quo::K_SEL ? (size_t)3 : (size_t)0;
            if (idx >= lim) TS_FAIL(c, TS_ERR_ARG, "quotient_values: operand out of range");
        }
    }
    const size_t m = (size_t)1 << log_m, n_chunks = (size_t)1 << log_quotient_degree;
    // program, constants and public values in one device buffer
    const size_t words = 4 * n_instr + n_consts + n_public + 4;
    std::vector<uint32_t> host(words, 0);
    if (n_instr) memcpy(host.data(), program, 16 * n_instr);
    if (n_consts) memcpy(host.data() + 4 * n_instr, consts_monty, 4 * n_consts);
    if (n_public) memcpy(host.data() + 4 * n_instr + n_consts, public_values_monty, 4 * n_public);
    uint32_t *dev = nullptr;
    TS_CUDA(c, pool_alloc(c, (void **)&dev, words * 4));
    cudaError_t e = cudaMemcpyAsync(dev, host.data(), words * 4, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);  // `host` is pageable and dies with this frame
    if (e != cudaSuccess) {
        pool_release(c, dev);
        TS_FAIL(c, TS_ERR_CUDA, std::string("quotient_values upload: ") + cudaGetErrorString(e));
    }
    quo::Params p;
    p.lde = trace_lde->d;
    p.width = (uint32_t)trace_lde->width;
    p.log_n = (int)log_n;
    p.log_m = (int)log_m;
    p.log_chunks = (int)log_quotient_degree;
    p.program = dev;
    p.n_instr = (uint32_t)n_instr;
    p.consts = dev + 4 * n_instr;
    p.publics = dev + 4 * n_instr + n_consts;
    for (int i = 0; i < 4; i++) p.alpha.c[i] = alpha_monty[i];
    p.rp = root_pows((int)log_m);
    p.g_monty = TS_GENERATOR_MONTY;
    p.wn_inv = h_to_monty(bb::cinv(bb::two_adic_generator((int)log_n)));
    {
        const uint32_t s_pow_n = bb::cpow(31, (uint64_t)1 << log_n);              // shift^n
        const uint32_t wq = bb::two_adic_generator((int)log_quotient_degree);      // w_m^n
        uint32_t x = 1;
        for (size_t k = 0; k < n_chunks; k++) {
            const uint32_t z = (uint32_t)(((uint64_t)bb::cmul(s_pow_n, x) + bb::P - 1) % bb::P);
            p.zh[k] = h_to_monty(z);
            p.zh_inv[k] = h_to_monty(bb::cinv(z));
            x = bb::cmul(x, wq);
        }
    }
    int rc = TS_OK;
    std::vector<ts_matrix *> outs;
    for (size_t k = 0; k < n_chunks && rc == TS_OK; k++) {
        ts_matrix *o = nullptr;
        rc = new_matrix(c, m >> log_quotient_degree, 4, &o);
        if (rc == TS_OK) {
            outs.push_back(o);
            p.out[k] = o->d;
        }
    }
    if (rc == TS_OK) {
        KScope ks(c, TS_K_MISC);
        auto kfn = quo::quotient_values_kernel;
        // TS_QV_CTAS_PER_SM=k: k resident CTAs per SM walking the rows grid-stride (fewer rows in flight per L1); default one
        // CTA per 128 rows
        size_t grid = (m + 127) / 128;
        if (const char *e = getenv("TS_QV_CTAS_PER_SM"))
            if (atoi(e) > 0) grid = std::min(grid, (size_t)atoi(e) * (size_t)c->num_sms);
        TS_LAUNCH(kfn, (unsigned)grid, 128, 0, c->stream, p);
        rc = check_launch(c, "quotient_values_kernel");
    }
    pool_release(c, dev);  // stream-ordered
    if (rc != TS_OK) {
        for (ts_matrix *o : outs) ts_matrix_free(o);
        return rc;
    }
    for (size_t k = 0; k < n_chunks; k++) chunks_out[k] = outs[k];
    return TS_OK;
}

int ts_inv_denoms(ts_ctx *c, unsigned log_h, const uint32_t z_monty[4], ts_matrix **out) {
    if (log_h > 27) TS_FAIL(c, TS_ERR_ARG, "inv_denoms: log_h > 27");
    ts_matrix *o = nullptr;
    TS_TRY(new_matrix(c, (size_t)1 << log_h, 4, &o));
    {
        KScope ks(c, TS_K_MISC);
        const size_t h = (size_t)1 << log_h;
        auto kfn = opn::inv_denoms_kernel;
        TS_LAUNCH(kfn, (unsigned)std::min<size_t>((h + 255) / 256, (size_t)c->num_sms * 16), 256, 0, c->stream, (uint4 *)o->d,
                  (int)log_h, h_to_monty(31), root_pows((int)log_h), to_e4(z_monty));
    }
    int rc = check_launch(c, "inv_denoms_kernel");
    if (rc != TS_OK) {
        ts_matrix_free(o);
        return rc;
    }
    *out = o;
    return TS_OK;
}
int ts_interpolate_low_coset(ts_ctx *c, const ts_matrix *lde, size_t n, const uint32_t z_monty[4],
                             const ts_matrix *inv_denoms, uint32_t *ys_out_monty) {
    const int log_n = log2_strict(n), log_h = log2_strict(lde->rows);
    if (log_n < 0 || log_h < 0 || n > lde->rows || inv_denoms->rows < n || inv_denoms->width != 4)
        TS_FAIL(c, TS_ERR_ARG, "interpolate_low_coset: bad sizes");
    const size_t w = lde->width;
    // 16-byte row loads for wide matrices (every committed trace); narrow ones (quotient chunks of 4 columns) are bound by the
    // per-block weight computation and run 2.6x faster in the scalar kernel with its 4+ CTAs per SM (105 against 277 us, measured)
    const bool quad = (w & 3) == 0 && w >= 32 && getenv("TS_NO_BARY4") == nullptr;
    const size_t n_blocks = quad ? (n + opn::BARY4_RB - 1) / opn::BARY4_RB : (n + opn::BARY_RB - 1) / opn::BARY_RB;
    size_t n_ctas = std::min<size_t>(n_blocks, (size_t)c->num_sms * (quad ? 2 : 4));
    if (const char *e = getenv("TS_BARY_CTAS")) n_ctas = std::max<size_t>(1, std::min<size_t>(n_blocks, strtoul(e, nullptr, 10)));  // test hook
    int tpr_log = 0;  // threads per row: the smallest power of two >= width (column quads when `quad`), at most 256
    while (tpr_log < 8 && ((size_t)1 << tpr_log) < (quad ? w / 4 : w)) tpr_log++;
    uint4 *partial = nullptr, *ys = nullptr;
    TS_CUDA(c, pool_alloc(c, (void **)&partial, n_ctas * w * 16));
    cudaError_t e = pool_alloc(c, (void **)&ys, w * 16);
    if (e != cudaSuccess) {
        pool_release(c, partial);
        TS_FAIL(c, TS_ERR_CUDA, cudaGetErrorString(e));
    }
    int rc;
    {
        // the low coset of the committed LDE: rows r < n hold p(g * w_n^bitrev_n(r))
        KScope ks(c, TS_K_MISC);
        if (quad) {
            auto kfn = opn::bary_partial4_kernel;
            TS_LAUNCH(kfn, (unsigned)n_ctas, 256, (size_t)(opn::BARY4_RB * 4 + 256 * 16) * 4, c->stream, (const uint4 *)lde->d, n,
                      (uint32_t)(w / 4), log_n, h_to_monty(31), root_pows(log_n), (const uint4 *)inv_denoms->d, partial, tpr_log);
            rc = check_launch(c, "bary_partial4_kernel");
        } else {
            auto kfn = opn::bary_partial_kernel;
            TS_LAUNCH(kfn, (unsigned)n_ctas, 256, (size_t)(opn::BARY_RB + 256) * 16, c->stream, (const uint32_t *)lde->d, n, (uint32_t)w,
                      log_n, h_to_monty(31), root_pows(log_n), (const uint4 *)inv_denoms->d, partial, tpr_log);
            rc = check_launch(c, "bary_partial_kernel");
        }
    }
    if (rc == TS_OK) {
        KScope ks(c, TS_K_MISC);
        auto kfn = opn::bary_final_kernel;
        TS_LAUNCH(kfn, (unsigned)((w + 31) / 32), 256, (size_t)256 * 16, c->stream, (const uint4 *)partial, n_ctas, (uint32_t)w, ys);
        rc = check_launch(c, "bary_final_kernel");
    }
    std::vector<uint32_t> s(w * 4);
    if (rc == TS_OK) {
        e = cudaMemcpyAsync(s.data(), ys, w * 16, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            c->err = cudaGetErrorString(e);
            rc = TS_ERR_CUDA;
        }
    }
    pool_release(c, partial);
    pool_release(c, ys);
    if (rc != TS_OK) return rc;
    // ys = -((z/g)^n - 1)/n * S   (host: a handful of extension-field operations on `width` values)
    uint32_t zc[4], t[4], f[4];
    const uint32_t ginv = bb::cinv(31);
    for (int i = 0; i < 4; i++) zc[i] = bb::cmul(h_from_monty(z_monty[i]), ginv);
    memcpy(t, zc, 16);
    for (int k = 0; k < log_n; k++) {
        uint32_t sq[4];
        h_ef_mul(t, t, sq);
        memcpy(t, sq, 16);
    }
    t[0] = (t[0] + bb::P - 1) % bb::P;  // (z/g)^n - 1
    const uint32_t fac = (bb::P - bb::cinv((uint32_t)(n % bb::P))) % bb::P;  // -1/n
    for (int i = 0; i < 4; i++) f[i] = bb::cmul(t[i], fac);
    for (size_t col = 0; col < w; col++) {
        uint32_t sc[4], y[4];
        for (int i = 0; i < 4; i++) sc[i] = h_from_monty(s[4 * col + i]);
        h_ef_mul(sc, f, y);
        for (int i = 0; i < 4; i++) ys_out_monty[4 * col + i] = h_to_monty(y[i]);
    }
    return TS_OK;
}
int ts_reduce_opening_acc(ts_ctx *c, const ts_matrix *dot, const ts_matrix *inv_denoms, const uint32_t alpha_pow_offset_monty[4],
                          const uint32_t reduced_ys_monty[4], ts_matrix *ro) {
    if (dot->width != 4 || ro->width != 4 || inv_denoms->width != 4 || dot->rows != ro->rows || inv_denoms->rows < ro->rows)
        TS_FAIL(c, TS_ERR_ARG, "reduce_opening_acc: EF vectors of equal length expected");
    KScope ks(c, TS_K_MISC);
    const size_t h = ro->rows;
    auto kfn = opn::reduce_rows_kernel;
    TS_LAUNCH(kfn, (unsigned)std::min<size_t>((h + 255) / 256, (size_t)c->num_sms * 16), 256, 0, c->stream,
              (const uint4 *)dot->d, (const uint4 *)inv_denoms->d, to_e4(alpha_pow_offset_monty), to_e4(reduced_ys_monty), h,
              (uint4 *)ro->d);
    return check_launch(c, "reduce_rows_kernel");
}
int ts_fill_splitmix(ts_ctx *c, uint32_t *dev, size_t rows, size_t width, uint64_t seed, size_t col0, size_t total_width,
                     int monty) {
    if (!c || !dev || width == 0 || col0 + width > total_width || total_width > 0xffffffffu) TS_FAIL(c, TS_ERR_ARG, "fill_splitmix: bad window");
    KScope ks(c, TS_K_MISC);
    auto kfn = ntt::fill_splitmix_kernel;
    const size_t total = rows * width;
    TS_LAUNCH(kfn, (unsigned)std::min<size_t>((total + 255) / 256, (size_t)c->num_sms * 16), 256, 0, c->stream, dev, rows,
              (uint32_t)width, seed, (uint32_t)col0, (uint32_t)total_width, monty);
    return check_launch(c, "fill_splitmix_kernel");
}
int ts_matrix_zero(ts_ctx *c, ts_matrix *m) {
    TS_CUDA(c, cudaMemsetAsync(m->d, 0, m->rows * m->width * 4, c->stream));
    return TS_OK;
}

// ---------------------------------------------------------------- sharded (multi-GPU) building blocks
int ts_coset_lde_batch_into(ts_ctx *c, const ts_matrix *evals, unsigned added_bits, uint32_t shift_monty,
                            ts_matrix *out) {
    if (out->rows != (evals->rows << added_bits) || out->width != evals->width)
        TS_FAIL(c, TS_ERR_ARG, "coset_lde_batch_into: out must be (rows<<added_bits) x width");
    c->shares_gpu = true;
    const int rc = lde_committed(c, evals->d, evals->rows, evals->width, added_bits, shift_monty, out->d);
    c->shares_gpu = false;
    return rc;
}
int ts_coset_lde_batch_scatter(ts_ctx *c, const ts_matrix *evals, unsigned added_bits, uint32_t shift_monty,
                               uint32_t *const *owner_ptrs, size_t n_owners, size_t dst_pitch) {
    if (n_owners < 1 || n_owners > 8 || (n_owners & (n_owners - 1))) TS_FAIL(c, TS_ERR_ARG, "lde scatter: 1, 2, 4 or 8 owners");
    if (dst_pitch < evals->width || (dst_pitch & 3)) TS_FAIL(c, TS_ERR_ARG, "lde scatter: bad pitch");
    PeerDst pd;
    pd.n = (int)n_owners;
    for (size_t i = 0; i < 8; i++) pd.ptr[i] = i < n_owners ? owner_ptrs[i] : nullptr;
    c->shares_gpu = true;
    const int rc = lde_committed(c, evals->d, evals->rows, evals->width, added_bits, shift_monty, nullptr, 0, dst_pitch, &pd);
    c->shares_gpu = false;
    return rc;
}
// CUDA IPC plumbing for the peer-mapped receive buffers (one process per GPU)
int ts_device_malloc(ts_ctx *c, size_t bytes, void **out) {
    TS_CUDA(c, cudaMalloc(out, bytes));
    TS_CUDA(c, cudaMemsetAsync(*out, 0, bytes, c->stream));  // zero-filled: mailboxes start with no flag set
    return TS_OK;
}
int ts_device_free(ts_ctx *c, void *p) {
    TS_CUDA(c, cudaStreamSynchronize(c->stream));
    TS_CUDA(c, cudaFree(p));
    return TS_OK;
}
int ts_ipc_get_handle(ts_ctx *c, void *dev_ptr, uint8_t handle[64]) {
#ifndef TS_EMULATE
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    cudaIpcMemHandle_t h;
    TS_CUDA(c, cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle, &h, 64);
    return TS_OK;
#else
    (void)dev_ptr;
    (void)handle;
    TS_FAIL(c, TS_ERR_ARG, "ipc: not available in the emulated build");
#endif
}
int ts_ipc_open(ts_ctx *c, const uint8_t handle[64], void **out) {
#ifndef TS_EMULATE
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    TS_CUDA(c, cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return TS_OK;
#else
    (void)handle;
    (void)out;
    TS_FAIL(c, TS_ERR_ARG, "ipc: not available in the emulated build");
#endif
}
int ts_ipc_close(ts_ctx *c, void *p) {
#ifndef TS_EMULATE
    TS_CUDA(c, cudaStreamSynchronize(c->stream));
    TS_CUDA(c, cudaIpcCloseMemHandle(p));
    return TS_OK;
#else
    (void)p;
    TS_FAIL(c, TS_ERR_ARG, "ipc: not available in the emulated build");
#endif
}
// ---- incremental commit of ONE matrix whose rows arrive as equal-width column blocks (row-sharded prover) ---------------
// The row hash is sequential in the row, so blocks are absorbed in column order; between windows the chaining value of
// every row is parked in the leaf-digest array (hash.cuh: hash_rows_fast_kernel, rows of at most one Blake3 chunk).
int ts_mmcs_commit_begin(ts_ctx *c, ts_matrix *const *blocks, size_t n_blocks, ts_tree **out) {
    if (!c || !blocks || !out || n_blocks == 0 || n_blocks > (size_t)b3::MAX_SEG) TS_FAIL(c, TS_ERR_ARG, "commit_begin: 1..32 column blocks");
    const size_t rows = blocks[0]->rows, bw = blocks[0]->width;
    if (log2_strict(rows) < 0 || bw < 4 || (bw & (bw - 1)) || bw * n_blocks > 256)
        TS_FAIL(c, TS_ERR_ARG, "commit_begin: power-of-two rows, equal power-of-two block widths >= 4, at most 256 columns in total");
    for (size_t i = 1; i < n_blocks; i++)
        if (blocks[i]->rows != rows || blocks[i]->width != bw) TS_FAIL(c, TS_ERR_ARG, "commit_begin: blocks differ in shape");
    ts_tree *t = new ts_tree;
    t->ctx = c;
    t->mats.assign(blocks, blocks + n_blocks);
    t->own_mats = false;
    t->layout = TS_LAYOUT_P3_INJECT;
    t->digests = nullptr;
    t->order.resize(n_blocks);
    for (size_t i = 0; i < n_blocks; i++) t->order[i] = i;
    t->hmax = rows;
    t->lmax = (unsigned)log2_strict(rows);
    size_t off = 0;
    for (unsigned l = 0; l <= t->lmax; l++) {
        t->layer_off.push_back(off);
        off += rows >> l;
    }
    if (pool_alloc(c, (void **)&t->digests, off * 32) != cudaSuccess) {
        ts_tree_free(t);
        TS_FAIL(c, TS_ERR_CUDA, "commit_begin: digest allocation failed");
    }
    *out = t;
    return TS_OK;
}
int ts_mmcs_commit_window(ts_ctx *c, ts_tree *t, size_t block_begin, size_t block_end) {
    const size_t nb = t->mats.size(), bw = t->mats[0]->width;
    if (block_begin >= block_end || block_end > nb || (block_begin * bw) % 16 || ((block_end * bw) % 16 && block_end != nb))
        TS_FAIL(c, TS_ERR_ARG, "commit_window: windows must cover whole 64-byte blocks of the row");
    b3::FastSegs fs;
    for (int i = 0; i < b3::MAX_SEG; i++) fs.ptr[i] = i < (int)nb ? t->mats[i]->d : nullptr;
    fs.n = (int)nb;
    fs.seg_w = (uint32_t)bw;
    fs.log_seg_w = 0;
    while ((1u << fs.log_seg_w) < fs.seg_w) fs.log_seg_w++;
    const size_t total = bw * nb;
    KScope ks(c, TS_K_HASH_LEAVES);
    auto kfn = b3::hash_rows_fast_kernel;
    const size_t per_block = (size_t)b3::FAST_WARPS * 32;
    TS_LAUNCH(kfn, (unsigned)((t->hmax + per_block - 1) / per_block), b3::FAST_WARPS * 32, (size_t)b3::FAST_WARPS * 512 * 4, c->stream, fs,
              (uint32_t)total, t->hmax, 1, t->digests, (uint32_t)(block_begin * bw / 16), (uint32_t)((block_end * bw + 15) / 16));
    return check_launch(c, "hash_rows_fast_kernel");
}
int ts_mmcs_commit_finish(ts_ctx *c, ts_tree *t, uint8_t *root_or_null) {
    TS_TRY(build_tree(c, t, true));
    if (root_or_null) {
        TS_CUDA(c, cudaMemcpyAsync(root_or_null, t->digests + t->layer_off[t->lmax] * 8, 32, cudaMemcpyDeviceToHost, c->stream));
        TS_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return TS_OK;
}

// ---- chained commit-phase rounds for a row-sharded layer (one process per GPU; collectives stay with the caller) -------
static int xfer_ready(ts_ctx *c, int lane, bool after_main) {
    if (lane < 0 || lane >= ts_ctx::XFER_LANES) TS_FAIL(c, TS_ERR_ARG, "copy: lane must be 0..7");
#ifndef TS_EMULATE
    if (!c->xfer_stream[lane]) {
        TS_CUDA(c, cudaStreamCreateWithFlags(&c->xfer_stream[lane], cudaStreamNonBlocking));
        TS_CUDA(c, cudaEventCreateWithFlags(&c->ev_xfer_in[lane], cudaEventDisableTiming));
        TS_CUDA(c, cudaEventCreateWithFlags(&c->ev_xfer_out[lane], cudaEventDisableTiming));
    }
    if (after_main) {
        TS_CUDA(c, cudaEventRecord(c->ev_xfer_in[lane], c->stream));
        TS_CUDA(c, cudaStreamWaitEvent(c->xfer_stream[lane], c->ev_xfer_in[lane], 0));
    }
#else
    (void)after_main;
#endif
    return TS_OK;
}
int ts_copy_async(ts_ctx *c, int lane, void *dst, const void *src, size_t bytes, int after_main) {
    if (!c || !dst || !src) TS_FAIL(c, TS_ERR_ARG, "copy_async: null pointer");
    TS_TRY(xfer_ready(c, lane, after_main != 0));
#ifndef TS_EMULATE
    TS_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, c->xfer_stream[lane]));
#else
    memmove(dst, src, bytes);
#endif
    return TS_OK;
}
int ts_copy2d_async(ts_ctx *c, int lane, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width_bytes,
                    size_t rows, int after_main) {
    if (!c || !dst || !src || width_bytes > dst_pitch || width_bytes > src_pitch) TS_FAIL(c, TS_ERR_ARG, "copy2d_async: bad argument");
    TS_TRY(xfer_ready(c, lane, after_main != 0));
#ifndef TS_EMULATE
    TS_CUDA(c, cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows, cudaMemcpyDefault, c->xfer_stream[lane]));
#else
    for (size_t r = 0; r < rows; r++) memmove((char *)dst + r * dst_pitch, (const char *)src + r * src_pitch, width_bytes);
#endif
    return TS_OK;
}
int ts_copy_join(ts_ctx *c, int lane) {
    if (lane < 0 || lane >= ts_ctx::XFER_LANES) TS_FAIL(c, TS_ERR_ARG, "copy: lane must be 0..7");
#ifndef TS_EMULATE
    if (!c->xfer_stream[lane]) return TS_OK;
    TS_CUDA(c, cudaEventRecord(c->ev_xfer_out[lane], c->xfer_stream[lane]));
    TS_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_xfer_out[lane], 0));
#endif
    return TS_OK;
}
int ts_fri_chain_begin(ts_ctx *c, const ts_challenger *chal, size_t max_rounds, uint32_t **chain_dev) {
    if (!c || !chal || !chain_dev) TS_FAIL(c, TS_ERR_ARG, "fri_chain_begin: null argument");
    if (!chal->in_buf.empty()) TS_FAIL(c, TS_ERR_ARG, "fri_chain_begin: the challenger has partially observed input");
    uint32_t *d = nullptr;
    TS_CUDA(c, pool_alloc(c, (void **)&d, (12 + 8 * std::max<size_t>(max_rounds, 1)) * 4));
    uint32_t h0[8];
    memcpy(h0, &chal->state[8][0], 32);
    cudaError_t e = cudaMemcpyAsync(d, h0, 32, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) {
        pool_release(c, d);
        TS_FAIL(c, TS_ERR_CUDA, std::string("fri_chain_begin: ") + cudaGetErrorString(e));
    }
    *chain_dev = d;
    return TS_OK;
}
int ts_fri_chain_step(ts_ctx *c, uint32_t *chain_dev, const uint8_t *sub_roots_dev, size_t n_sub, size_t round) {
    if (!chain_dev || !sub_roots_dev || n_sub == 0 || n_sub > 32 || (n_sub & (n_sub - 1))) TS_FAIL(c, TS_ERR_ARG, "fri_chain_step: 1..32 sub-roots, a power of two");
    KScope ks(c, TS_K_TREE);
    auto kfn = ftail::sponge_step_kernel;
    TS_LAUNCH(kfn, 1, 32, 0, c->stream, reinterpret_cast<const uint32_t *>(sub_roots_dev), (int)n_sub, chain_dev, chain_dev + 12 + 8 * round,
              chain_dev + 8);
    return check_launch(c, "sponge_step_kernel");
}
int ts_fri_fold_ext_shard_chain(ts_ctx *c, const uint32_t *in_dev, size_t h_global, size_t first, size_t h_local,
                                const uint32_t *chain_dev, const uint32_t *addend_dev, uint32_t *out_dev) {
    const int lh = log2_strict(h_global);
    if (!chain_dev || lh < 0 || lh > 26 || first + h_local > h_global) TS_FAIL(c, TS_ERR_ARG, "fold shard: bad range");
    if (lh >= 8 && ((first | h_local) & 255)) TS_FAIL(c, TS_ERR_ARG, "fold shard: range must be a multiple of 256 rows");
    if (lh < 8 && (first != 0 || h_local != h_global)) TS_FAIL(c, TS_ERR_ARG, "fold shard: small layers are not sharded");
    return fold_ext_launch(c, in_dev, out_dev, addend_dev, lh, nullptr, first, h_local, chain_dev + 8);
}
int ts_fri_chain_end(ts_ctx *c, uint32_t *chain_dev, ts_challenger *chal, size_t n_rounds, uint8_t *commits_out) {
    if (!chain_dev) return TS_OK;
    std::vector<uint32_t> roots(8 * std::max<size_t>(n_rounds, 1));
    cudaError_t e = cudaSuccess;
    if (n_rounds) e = cudaMemcpyAsync(roots.data(), chain_dev + 12, n_rounds * 32, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    pool_release(c, chain_dev);
    if (e != cudaSuccess) TS_FAIL(c, TS_ERR_CUDA, std::string("fri_chain_end: ") + cudaGetErrorString(e));
    for (size_t r = 0; r < n_rounds && chal; r++) {  // the host challenger replays the rounds (prover.rs:114-116)
        const uint8_t *root = reinterpret_cast<const uint8_t *>(roots.data() + 8 * r);
        if (commits_out) memcpy(commits_out + 32 * r, root, 32);
        ts_challenger_observe_digest(chal, root);
        uint32_t beta[4];
        ts_challenger_sample_ext(chal, beta);
    }
    return TS_OK;
}
int ts_fri_mailbox_words(void) { return ftail::MAIL_WORDS; }
// All sharded commit-phase rounds of one step in one call: per round leaf hash + subtree, publish the sub-root into every
// rank's mailbox, sponge step once all sub-roots of the round have arrived, fold with the beta it leaves on the device.
int ts_fri_commit_phase_sharded(ts_ctx *c, const uint32_t *cur_dev, size_t len_global, size_t rank, size_t world, size_t n_rounds,
                                uint32_t *const *mailboxes, uint32_t epoch, ts_challenger *chal, uint8_t *commits_out,
                                uint32_t *out_dev) {
    if (!c || !cur_dev || !mailboxes || !chal || !out_dev) TS_FAIL(c, TS_ERR_ARG, "fri sharded: null argument");
    if (world < 2 || world > (size_t)ftail::MAIL_RANKS || (world & (world - 1)) || rank >= world || n_rounds == 0 || n_rounds > (size_t)ftail::MAIL_ROUNDS)
        TS_FAIL(c, TS_ERR_ARG, "fri sharded: 2..8 ranks (a power of two), 1..32 rounds");
    if (log2_strict(len_global) < 0 || (len_global >> n_rounds) / world < 1 || ((len_global >> n_rounds) / world) * world != (len_global >> n_rounds))
        TS_FAIL(c, TS_ERR_ARG, "fri sharded: the layers must stay divisible among the ranks");
    uint32_t *chain = nullptr;
    TS_TRY(ts_fri_chain_begin(c, chal, n_rounds, &chain));
    ftail::Mailboxes mb;
    for (size_t i = 0; i < (size_t)ftail::MAIL_RANKS; i++) mb.box[i] = i < world ? mailboxes[i] : nullptr;
    uint32_t *cur = const_cast<uint32_t *>(cur_dev);
    bool cur_owned = false;
    size_t len_g = len_global, local = len_global / world;
    int rc = TS_OK;
    uint32_t *next_digests = nullptr;
    for (size_t r = 0; r < n_rounds && rc == TS_OK; r++) {
        const size_t h_g = len_g / 2, h_l = local / 2;
        ts_matrix *leaves = new ts_matrix{c, cur, h_l, 8, cur_owned};
        ts_tree *tree = nullptr;
        uint32_t *filled = next_digests;  // sub-tree leaf layer already hashed by the previous round's fold_hash_kernel
        next_digests = nullptr;
        rc = mmcs_commit(c, &leaves, 1, TS_LAYOUT_P3_INJECT, 1, nullptr, &tree, false, nullptr, filled);
        if (rc != TS_OK) {
            ts_matrix_free(leaves);
            cur = nullptr;
            break;
        }
        {
            KScope ks(c, TS_K_TREE);
            auto kfn = ftail::publish_subroot_kernel;
            TS_LAUNCH(kfn, 1, 32, 0, c->stream, (const uint32_t *)(tree->digests + tree->layer_off[tree->lmax] * 8), mb, (int)world, (int)rank, (int)r,
                      epoch);
            rc = check_launch(c, "publish_subroot_kernel");
        }
        if (rc == TS_OK) {
            KScope ks(c, TS_K_TREE);
            auto kfn = ftail::sponge_step_mailbox_kernel;
            TS_LAUNCH(kfn, 1, 32, 0, c->stream, (const uint32_t *)mailboxes[rank], (int)world, (int)r, epoch, chain, chain + 12 + 8 * r, chain + 8);
            rc = check_launch(c, "sponge_step_mailbox_kernel");
        }
        uint32_t *nf = out_dev;
        if (rc == TS_OK && r + 1 < n_rounds && pool_alloc(c, (void **)&nf, h_l * 16) != cudaSuccess) {
            c->err = "fri sharded: layer allocation failed";
            rc = TS_ERR_CUDA;
        }
        if (rc == TS_OK) {
            if (r + 1 < n_rounds && (h_l & 511) == 0 && fold_hash_eligible(log2_strict(h_g)) &&
                pool_alloc(c, (void **)&next_digests, (h_l - 1) * 32) == cudaSuccess) {
                rc = fold_hash_launch(c, cur, nf, nullptr, log2_strict(h_g), nullptr, chain + 8, next_digests, rank * h_l, h_l);
            } else {
                next_digests = nullptr;
                rc = fold_ext_launch(c, cur, nf, nullptr, log2_strict(h_g), nullptr, rank * h_l, h_l, chain + 8);
            }
        }
        ts_tree_free(tree);  // releases the layer it owns (stream-ordered pool)
        cur = nf;
        cur_owned = r + 1 < n_rounds;
        len_g = h_g;
        local = h_l;
    }
    if (next_digests) pool_release(c, next_digests);
    if (rc != TS_OK) {
        if (cur_owned && cur) pool_release(c, cur);
        ts_fri_chain_end(c, chain, nullptr, 0, nullptr);
        return rc;
    }
    return ts_fri_chain_end(c, chain, chal, n_rounds, commits_out);
}
int ts_fri_fold_ext_shard(ts_ctx *c, const uint32_t *in_dev, size_t h_global, size_t first, size_t h_local,
                          const uint32_t beta_monty[4], const uint32_t *addend_dev, uint32_t *out_dev) {
    const int lh = log2_strict(h_global);
    if (lh < 0 || lh > 26 || first + h_local > h_global) TS_FAIL(c, TS_ERR_ARG, "fold shard: bad range");
    if (lh >= 8 && ((first | h_local) & 255)) TS_FAIL(c, TS_ERR_ARG, "fold shard: range must be a multiple of 256 rows");
    if (lh < 8 && (first != 0 || h_local != h_global)) TS_FAIL(c, TS_ERR_ARG, "fold shard: small layers are not sharded");
    uint32_t beta[4];
    for (int i = 0; i < 4; i++) beta[i] = h_from_monty(beta_monty[i]);
    return fold_ext_launch(c, in_dev, out_dev, addend_dev, lh, beta, first, h_local);
}

int ts_fri_fold_hash_shard(ts_ctx *c, const uint32_t *in_dev, size_t h_global, size_t first, size_t h_local,
                           const uint32_t beta_monty[4], const uint32_t *addend_dev, uint32_t *out_dev, uint32_t *next_digests_dev) {
    const int lh = log2_strict(h_global);
    if (lh < 9 || lh > 26 || first + h_local > h_global || !next_digests_dev) TS_FAIL(c, TS_ERR_ARG, "fold_hash shard: bad range");
    uint32_t beta[4];
    for (int i = 0; i < 4; i++) beta[i] = h_from_monty(beta_monty[i]);
    return fold_hash_launch(c, in_dev, out_dev, addend_dev, lh, beta, nullptr, next_digests_dev, first, h_local);
}

// ---------------------------------------------------------------- Pcs::open + bf_prove behind the ABI, proof bytes
namespace {
// rows (concatenated over the tree's matrices) and paths of many indices of one tree: one gather launch, one read-back
int open_many(ts_ctx *c, const ts_tree *t, const std::vector<size_t> &idx, std::vector<uint32_t> &rows, std::vector<uint8_t> &paths,
              size_t *row_words_out) {
    size_t row_words = 0;
    for (const ts_matrix *m : t->mats) row_words += m->width;
    *row_words_out = row_words;
    const size_t q = idx.size(), per = row_words + 8 * (size_t)t->lmax;
    rows.assign(q * row_words, 0);
    paths.assign(q * 32 * (size_t)t->lmax, 0);
    if (q == 0 || per == 0) return TS_OK;
    std::vector<opn::GatherSeg> segs;
    for (size_t k = 0; k < q; k++) {
        if (idx[k] >= t->hmax) TS_FAIL(c, TS_ERR_ARG, "open: index out of range");
        size_t o = k * per;
        for (const ts_matrix *m : t->mats) {
            const size_t row = idx[k] >> (t->lmax - (unsigned)log2_strict(m->rows));
            segs.push_back({m->d + row * m->width, (uint32_t)o, (uint32_t)m->width});
            o += m->width;
        }
        for (unsigned l = 0; l < t->lmax; l++) {
            const size_t node = (idx[k] >> l) ^ 1;
            segs.push_back({t->digests + (t->layer_off[l] + node) * 8, (uint32_t)o, 8u});
            o += 8;
        }
    }
    opn::GatherSeg *dsegs = nullptr;
    uint32_t *dbuf = nullptr;
    TS_CUDA(c, pool_alloc(c, (void **)&dsegs, segs.size() * sizeof(opn::GatherSeg)));
    cudaError_t e = pool_alloc(c, (void **)&dbuf, q * per * 4);
    if (e != cudaSuccess) {
        pool_release(c, dsegs);
        TS_FAIL(c, TS_ERR_CUDA, cudaGetErrorString(e));
    }
    std::vector<uint32_t> host(q * per);
    int rc = TS_OK;
    e = cudaMemcpyAsync(dsegs, segs.data(), segs.size() * sizeof(opn::GatherSeg), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        KScope ks(c, TS_K_MISC);
        auto kfn = opn::gather_segments_kernel;
        TS_LAUNCH(kfn, (unsigned)std::min<size_t>(segs.size(), 4096), 64, 0, c->stream, (const opn::GatherSeg *)dsegs,
                  (uint32_t)segs.size(), dbuf);
        rc = check_launch(c, "gather_segments_kernel");
    }
    if (e == cudaSuccess && rc == TS_OK) e = cudaMemcpyAsync(host.data(), dbuf, q * per * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && rc == TS_OK) e = cudaStreamSynchronize(c->stream);
    pool_release(c, dsegs);
    pool_release(c, dbuf);
    if (e != cudaSuccess) TS_FAIL(c, TS_ERR_CUDA, cudaGetErrorString(e));
    if (rc != TS_OK) return rc;
    for (size_t k = 0; k < q; k++) {
        memcpy(rows.data() + k * row_words, host.data() + k * per, row_words * 4);
        memcpy(paths.data() + k * 32 * (size_t)t->lmax, host.data() + k * per + row_words, 32 * (size_t)t->lmax);
    }
    return TS_OK;
}

// postcard (serde) wire format: unsigned integers as LEB128 varints, Vec<T> = varint length + elements, fixed arrays and
// tuples/structs = elements in order, u8 = one raw byte.  BabyBear = its canonical u32 ([MEM] p3-baby-bear Serialize).
struct Postcard {
    std::vector<uint8_t> b;
    void varint(uint64_t v) {
        while (v >= 0x80) {
            b.push_back((uint8_t)(v | 0x80));
            v >>= 7;
        }
        b.push_back((uint8_t)v);
    }
    void len(size_t n) { varint(n); }
    void bb_monty(uint32_t m) { varint(h_from_monty(m)); }
    void bb_canon(uint32_t v) { varint(v); }
    void digest(const uint8_t *d) { b.insert(b.end(), d, d + 32); }
    void path(const uint8_t *p, size_t depth) {  // Vec<[u8; 32]>
        len(depth);
        b.insert(b.end(), p, p + 32 * depth);
    }
};
void h_ef_pow(const uint32_t a[4], uint64_t e, uint32_t o[4]) {
    uint32_t r[4] = {1, 0, 0, 0}, base[4], t[4];
    memcpy(base, a, 16);
    while (e) {
        if (e & 1) {
            h_ef_mul(r, base, t);
            memcpy(r, t, 16);
        }
        h_ef_mul(base, base, t);
        memcpy(base, t, 16);
        e >>= 1;
    }
    memcpy(o, r, 16);
}
}  // namespace

int ts_pcs_open(ts_ctx *c, const ts_tree *const *rounds, size_t n_rounds, const size_t *n_points, const uint32_t *points_monty,
                unsigned log_blowup, unsigned num_queries, unsigned pow_bits, ts_challenger *chal, uint8_t **out_bytes,
                size_t *out_len) {
    if (!c || !rounds || !n_rounds || !n_points || !chal || !out_bytes || !out_len) TS_FAIL(c, TS_ERR_ARG, "pcs_open: null argument");
    *out_bytes = nullptr;
    *out_len = 0;
    uint32_t alpha[4], alpha_m[4];  // :312  alpha = challenger.sample()
    ts_challenger_sample_ext(chal, alpha);
    for (int i = 0; i < 4; i++) alpha_m[i] = h_to_monty(alpha[i]);
    // (round, matrix) -> first point index
    struct MatRef { const ts_matrix *m; size_t first_point, n_points; };
    std::vector<std::vector<MatRef>> mats(n_rounds);
    size_t pt = 0, mi = 0;
    unsigned log_global_max_height = 0;
    for (size_t r = 0; r < n_rounds; r++) {
        if (!rounds[r]) TS_FAIL(c, TS_ERR_ARG, "pcs_open: null round");
        for (const ts_matrix *m : rounds[r]->mats) {
            const int lh = log2_strict(m->rows);
            if (lh < (int)log_blowup) TS_FAIL(c, TS_ERR_ARG, "pcs_open: matrix shorter than the blowup");
            mats[r].push_back({m, pt, n_points[mi]});
            pt += n_points[mi++];
            log_global_max_height = std::max(log_global_max_height, (unsigned)lh);
        }
    }
    if (pt && !points_monty) TS_FAIL(c, TS_ERR_ARG, "pcs_open: null points");
    // compute_inverse_denominators (:677-720): per unique point, for the largest height it is opened at
    std::map<std::array<uint32_t, 4>, unsigned> max_lh;
    auto key_of = [&](size_t p) { return std::array<uint32_t, 4>{points_monty[4 * p], points_monty[4 * p + 1], points_monty[4 * p + 2], points_monty[4 * p + 3]}; };
    for (auto &rm : mats)
        for (auto &mr : rm)
            for (size_t k = 0; k < mr.n_points; k++) {
                auto key = key_of(mr.first_point + k);
                max_lh[key] = std::max(max_lh.count(key) ? max_lh[key] : 0u, (unsigned)log2_strict(mr.m->rows));
            }
    std::map<std::array<uint32_t, 4>, ts_matrix *> inv;
    std::map<unsigned, ts_matrix *> reduced;
    std::map<unsigned, size_t> num_reduced;
    std::vector<ts_tree *> fri_trees;
    auto cleanup = [&]() {
        for (auto &kv : inv) ts_matrix_free(kv.second);
        for (auto &kv : reduced) ts_matrix_free(kv.second);
        for (ts_tree *t : fri_trees) ts_tree_free(t);
    };
#define PO_TRY(expr)           \
    do {                       \
        int rc__ = (expr);     \
        if (rc__ != TS_OK) {   \
            cleanup();         \
            return rc__;       \
        }                      \
    } while (0)
    for (auto &kv : max_lh) {
        ts_matrix *d = nullptr;
        PO_TRY(ts_inv_denoms(c, kv.second, kv.first.data(), &d));
        inv[kv.first] = d;
    }
    Postcard pc;
    // all_opened_values: Vec (round) < Vec (matrix) < Vec (point) < Vec<Challenge> > > >
    pc.len(n_rounds);
    for (size_t r = 0; r < n_rounds; r++) {
        pc.len(mats[r].size());
        for (auto &mr : mats[r]) {
            const ts_matrix *m = mr.m;
            const unsigned lh = (unsigned)log2_strict(m->rows);
            if (!reduced.count(lh)) {
                ts_matrix *ro = nullptr;
                PO_TRY(new_matrix(c, m->rows, 4, &ro));
                reduced[lh] = ro;
                num_reduced[lh] = 0;
                PO_TRY(ts_matrix_zero(c, ro));
            }
            ts_matrix *dot = nullptr;  // sum_i alpha^i p_i[X], shared by all points of this matrix (:375)
            if (mr.n_points) PO_TRY(ts_dot_ext_powers(c, m, alpha_m, &dot));
            pc.len(mr.n_points);
            std::vector<uint32_t> ys(m->width * 4);
            for (size_t k = 0; k < mr.n_points; k++) {
                const uint32_t *zm = points_monty + 4 * (mr.first_point + k);
                ts_matrix *idn = inv[key_of(mr.first_point + k)];
                int rc = ts_interpolate_low_coset(c, m, m->rows >> log_blowup, zm, idn, ys.data());  // :358-369
                uint32_t apo[4], rys[4] = {0, 0, 0, 0}, ap[4] = {1, 0, 0, 0}, apo_m[4], rys_m[4];
                if (rc == TS_OK) {
                    h_ef_pow(alpha, num_reduced[lh], apo);
                    pc.len(m->width);
                    for (size_t col = 0; col < m->width; col++) {  // dot_product(alpha.powers(), ys)
                        uint32_t y[4], t[4];
                        for (int i = 0; i < 4; i++) {
                            y[i] = h_from_monty(ys[4 * col + i]);
                            pc.bb_canon(y[i]);
                        }
                        h_ef_mul(ap, y, t);
                        for (int i = 0; i < 4; i++) rys[i] = (uint32_t)(((uint64_t)rys[i] + t[i]) % bb::P);
                        h_ef_mul(ap, alpha, t);
                        memcpy(ap, t, 16);
                    }
                    for (int i = 0; i < 4; i++) apo_m[i] = h_to_monty(apo[i]), rys_m[i] = h_to_monty(rys[i]);
                    rc = ts_reduce_opening_acc(c, dot, idn, apo_m, rys_m, reduced[lh]);  // :371-381
                    num_reduced[lh] += m->width;
                }
                if (rc != TS_OK) {
                    ts_matrix_free(dot);
                    cleanup();
                    return rc;
                }
            }
            if (dot) ts_matrix_free(dot);
        }
    }
    // fri_input: reduced openings sorted by height, descending (:389); bf_prove (fri/src/prover.rs:19-67)
    std::vector<ts_matrix *> fri_in;
    for (auto it = reduced.rbegin(); it != reduced.rend(); ++it) fri_in.push_back(it->second);
    const size_t max_rounds = log_global_max_height > log_blowup ? log_global_max_height - log_blowup : 0;
    std::vector<uint8_t> commits(32 * std::max<size_t>(max_rounds, 1));
    fri_trees.assign(std::max<size_t>(max_rounds, 1), nullptr);
    uint32_t final_poly[4];
    size_t n_fri = 0;
    {
        int rc = ts_fri_commit_phase(c, fri_in.data(), fri_in.size(), log_blowup, chal, commits.data(), fri_trees.data(), final_poly, &n_fri);
        PO_TRY(rc);
        fri_trees.resize(n_fri);  // entries beyond the rounds actually run are null
    }
    uint32_t pow_witness = 0;
    PO_TRY(ts_challenger_grind(chal, pow_bits, 1, &pow_witness));
    std::vector<size_t> index(num_queries);
    for (unsigned q = 0; q < num_queries; q++) index[q] = ts_challenger_sample_bits(chal, log_global_max_height, 1);
    // query phase: per tree, every query's rows + path in one gather
    struct Opened { std::vector<uint32_t> rows; std::vector<uint8_t> paths; size_t row_words; };
    std::vector<Opened> in_open(n_rounds), fri_open(n_fri);
    for (size_t r = 0; r < n_rounds; r++) {
        std::vector<size_t> idx(num_queries);
        for (unsigned q = 0; q < num_queries; q++) idx[q] = index[q] >> (log_global_max_height - rounds[r]->lmax);  // :399-413
        PO_TRY(open_many(c, rounds[r], idx, in_open[r].rows, in_open[r].paths, &in_open[r].row_words));
    }
    for (size_t i = 0; i < n_fri; i++) {
        std::vector<size_t> idx(num_queries);
        for (unsigned q = 0; q < num_queries; q++) idx[q] = index[q] >> i >> 1;  // prover.rs:69-90
        PO_TRY(open_many(c, fri_trees[i], idx, fri_open[i].rows, fri_open[i].paths, &fri_open[i].row_words));
    }
    // FriProof { commit_phase_commits, query_proofs, final_poly, pow_witness }  (fri/src/proof.rs:13-33)
    pc.len(n_fri);
    for (size_t i = 0; i < n_fri; i++) pc.digest(commits.data() + 32 * i);
    pc.len(num_queries);
    for (unsigned q = 0; q < num_queries; q++) {
        pc.len(n_rounds);  // input_proof: Vec<BatchOpening { opened_values: Vec<Vec<Val>>, opening_proof }>
        for (size_t r = 0; r < n_rounds; r++) {
            const ts_tree *t = rounds[r];
            pc.len(t->mats.size());
            size_t o = q * in_open[r].row_words;
            for (const ts_matrix *m : t->mats) {
                pc.len(m->width);
                for (size_t k = 0; k < m->width; k++) pc.bb_monty(in_open[r].rows[o + k]);
                o += m->width;
            }
            pc.path(in_open[r].paths.data() + (size_t)q * 32 * t->lmax, t->lmax);
        }
        pc.len(n_fri);  // commit_phase_openings: Vec<(Vec<Vec<F>>, Proof)>
        for (size_t i = 0; i < n_fri; i++) {
            const ts_tree *t = fri_trees[i];
            pc.len(1);
            pc.len(2);
            for (int k = 0; k < 8; k++) pc.bb_monty(fri_open[i].rows[q * fri_open[i].row_words + k]);
            pc.path(fri_open[i].paths.data() + (size_t)q * 32 * t->lmax, t->lmax);
        }
    }
    for (int i = 0; i < 4; i++) pc.bb_canon(final_poly[i]);
    pc.bb_canon(pow_witness);
    cleanup();
#undef PO_TRY
    uint8_t *buf = (uint8_t *)malloc(pc.b.size() ? pc.b.size() : 1);
    if (!buf) TS_FAIL(c, TS_ERR_ARG, "pcs_open: out of host memory");
    memcpy(buf, pc.b.data(), pc.b.size());
    *out_bytes = buf;
    *out_len = pc.b.size();
    return TS_OK;
}
void ts_bytes_free(uint8_t *p) { free(p); }


// ---------------------------------------------------------------- TapTree commitment (f2, first slice)
struct ts_taptree {
    ts_ctx *ctx;
    size_t n_leaves;
    unsigned log_n;
    uint32_t *nodes;     // all levels, leaves first: (2 n - 1) x 8 state words
    uint8_t *swapped;    // n - 1 flags, level 0 first
    uint32_t *leaf_idx;  // n: reverse_idx_dict
};
void ts_taptree_free(ts_taptree *t) {
    if (!t) return;
    pool_release(t->ctx, t->nodes);
    pool_release(t->ctx, t->swapped);
    pool_release(t->ctx, t->leaf_idx);
    delete t;
}
static void be_bytes(const uint32_t w[8], uint8_t out[32]) {
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(w[i] >> 24), out[4 * i + 1] = (uint8_t)(w[i] >> 16), out[4 * i + 2] = (uint8_t)(w[i] >> 8), out[4 * i + 3] = (uint8_t)w[i];
    }
}
int ts_taptree_commit(ts_ctx *c, const ts_matrix *leaf_rows, const uint8_t *segs, const size_t *seg_offsets, const uint32_t *push_word,
                      size_t n_push, uint8_t root[32], ts_taptree **out) {
    if (!c || !leaf_rows || !segs || !seg_offsets || !out || n_push == 0 || n_push > sha::MAX_PUSH) TS_FAIL(c, TS_ERR_ARG, "taptree: bad argument");
    if (n_push > 1 && !push_word) TS_FAIL(c, TS_ERR_ARG, "taptree: push_word missing");
    const size_t n = leaf_rows->rows;
    const int log_n = log2_strict(n);
    if (log_n < 0 || n > ((size_t)1 << 31)) TS_FAIL(c, TS_ERR_ARG, "taptree: the leaf count must be a power of two");  // builder.rs:40
    for (size_t k = 1; k < n_push; k++)
        if (push_word[k - 1] >= leaf_rows->width) TS_FAIL(c, TS_ERR_ARG, "taptree: push_word out of range");
    const size_t total = seg_offsets[n_push + 1];
    if (total > 0xffffffffu) TS_FAIL(c, TS_ERR_ARG, "taptree: template too long");
    ts_taptree *t = new ts_taptree{c, n, (unsigned)log_n, nullptr, nullptr, nullptr};
    uint8_t *d_segs = nullptr;
    uint32_t *d_off = nullptr, *d_pw = nullptr;
    auto fail = [&](int rc, const char *msg) {
        if (msg) c->err = msg;
        pool_release(c, d_segs), pool_release(c, d_off), pool_release(c, d_pw);
        ts_taptree_free(t);
        return rc;
    };
    if (pool_alloc(c, (void **)&t->nodes, (2 * n - 1) * 32) != cudaSuccess || pool_alloc(c, (void **)&t->swapped, std::max<size_t>(n - 1, 1)) != cudaSuccess ||
        pool_alloc(c, (void **)&t->leaf_idx, n * 4) != cudaSuccess || pool_alloc(c, (void **)&d_segs, std::max<size_t>(total, 1)) != cudaSuccess ||
        pool_alloc(c, (void **)&d_off, (n_push + 2) * 4) != cudaSuccess || pool_alloc(c, (void **)&d_pw, std::max<size_t>(n_push - 1, 1) * 4) != cudaSuccess)
        return fail(TS_ERR_CUDA, "taptree: device allocation failed");
    std::vector<uint32_t> off32(n_push + 2);
    for (size_t k = 0; k < n_push + 2; k++) off32[k] = (uint32_t)seg_offsets[k];
    cudaError_t e = cudaMemcpyAsync(d_segs, segs, total, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, off32.data(), off32.size() * 4, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n_push > 1) e = cudaMemcpyAsync(d_pw, push_word, (n_push - 1) * 4, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) return fail(TS_ERR_CUDA, cudaGetErrorString(e));
    int rc;
    {
        sha::LeafParams p;
        p.rows = leaf_rows->d, p.width = (uint32_t)leaf_rows->width, p.n_leaves = n, p.segs = d_segs, p.seg_off = d_off, p.push_word = d_pw;
        p.n_push = (uint32_t)n_push, p.const_len = (uint32_t)total, p.out = t->nodes;
        for (int i = 0; i < 8; i++) p.midstate[i] = sha::TAPLEAF_MID[i];
        KScope ks(c, TS_K_HASH_LEAVES);
        auto kfn = sha::taptree_leaf_kernel;
        TS_LAUNCH(kfn, (unsigned)((n + sha::LEAF_NT - 1) / sha::LEAF_NT), sha::LEAF_NT, (size_t)16 * sha::LEAF_NT * 4, c->stream, p);
        rc = check_launch(c, "taptree_leaf_kernel");
    }
    size_t in_off = 0, flag_off = 0;
    for (size_t w = n; w > 1 && rc == TS_OK; w >>= 1) {
        KScope ks(c, TS_K_TREE);
        auto kfn = sha::taptree_branch_kernel;
        TS_LAUNCH(kfn, (unsigned)((w / 2 + 255) / 256), 256, 0, c->stream, (const uint32_t *)(t->nodes + in_off * 8), w / 2,
                  t->nodes + (in_off + w) * 8, t->swapped + flag_off);
        rc = check_launch(c, "taptree_branch_kernel");
        in_off += w;
        flag_off += w / 2;
    }
    if (rc == TS_OK) {
        KScope ks(c, TS_K_MISC);
        auto kfn = sha::taptree_perm_kernel;
        TS_LAUNCH(kfn, (unsigned)std::min<size_t>((n + 255) / 256, (size_t)c->num_sms * 8), 256, 0, c->stream, (const uint8_t *)t->swapped,
                  (uint32_t)log_n, t->leaf_idx);
        rc = check_launch(c, "taptree_perm_kernel");
    }
    uint32_t rw[8];
    if (rc == TS_OK) {
        e = cudaMemcpyAsync(rw, t->nodes + (2 * n - 2) * 8, 32, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return fail(TS_ERR_CUDA, cudaGetErrorString(e));
    }
    pool_release(c, d_segs), pool_release(c, d_off), pool_release(c, d_pw);
    d_segs = nullptr, d_off = nullptr, d_pw = nullptr;
    if (rc != TS_OK) return fail(rc, nullptr);
    if (root) be_bytes(rw, root);
    *out = t;
    return TS_OK;
}
int ts_padded_leaf_rows(ts_ctx *c, const ts_matrix *const *mats, size_t n_mats, ts_matrix **out) {
    if (!c || !mats || !out || n_mats == 0 || n_mats > (size_t)b3::MAX_SEG) TS_FAIL(c, TS_ERR_ARG, "padded_leaf_rows: 1..32 matrices");
    std::vector<size_t> order(n_mats);
    for (size_t i = 0; i < n_mats; i++) {
        order[i] = i;
        if (!mats[i] || log2_strict(mats[i]->rows) < 0) TS_FAIL(c, TS_ERR_ARG, "padded_leaf_rows: heights must be powers of two");
    }
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return mats[a]->rows > mats[b]->rows; });  // tcs/mod.rs:346
    const size_t hmax = mats[order[0]]->rows;
    b3::Segments sg;
    sg.n = (int)n_mats, sg.total_words = 0;
    for (int i = 0; i < b3::MAX_SEG; i++) sg.ptr[i] = nullptr, sg.width[i] = 0xffffffffu, sg.shift[i] = 0;
    for (size_t i = 0; i < n_mats; i++) {
        const ts_matrix *m = mats[order[i]];
        sg.ptr[i] = m->d, sg.width[i] = (uint32_t)m->width;
        sg.shift[i] = (uint32_t)(log2_strict(hmax) - log2_strict(m->rows));
        sg.total_words += (uint32_t)m->width;
    }
    if (sg.total_words == 0) TS_FAIL(c, TS_ERR_ARG, "padded_leaf_rows: empty rows");
    ts_matrix *o = nullptr;
    TS_TRY(new_matrix(c, hmax, sg.total_words, &o));
    {
        KScope ks(c, TS_K_MISC);
        auto kfn = b3::padded_rows_kernel;
        const size_t total = hmax * sg.total_words;
        TS_LAUNCH(kfn, (unsigned)std::min<size_t>((total + 255) / 256, (size_t)c->num_sms * 16), 256, 0, c->stream, sg, hmax, o->d);
        const int rc = check_launch(c, "padded_rows_kernel");
        if (rc != TS_OK) {
            ts_matrix_free(o);
            return rc;
        }
    }
    *out = o;
    return TS_OK;
}
int ts_taptree_open(ts_ctx *c, const ts_taptree *t, size_t index, uint8_t *path_out, uint32_t *position_out) {
    if (!c || !t || index >= t->n_leaves || (t->log_n && !path_out)) TS_FAIL(c, TS_ERR_ARG, "taptree_open: bad argument");
    // TapBranch hashes are symmetric in their children, so the branch of Merkle leaf m is its siblings in construction order
    // (complete_taptree.rs:34-49: the TaprootMerkleBranch of leaf_indices[m]), leaf level first
    std::vector<opn::GatherSeg> segs;
    size_t off = 0;
    for (unsigned l = 0; l < t->log_n; l++) {
        segs.push_back({t->nodes + (off + ((index >> l) ^ 1)) * 8, (uint32_t)(8 * l), 8u});
        off += t->n_leaves >> l;
    }
    segs.push_back({t->leaf_idx + index, (uint32_t)(8 * t->log_n), 1u});
    opn::GatherSeg *dsegs = nullptr;
    uint32_t *dbuf = nullptr;
    const size_t words = 8 * (size_t)t->log_n + 1;
    TS_CUDA(c, pool_alloc(c, (void **)&dsegs, segs.size() * sizeof(opn::GatherSeg)));
    cudaError_t e = pool_alloc(c, (void **)&dbuf, words * 4);
    if (e != cudaSuccess) {
        pool_release(c, dsegs);
        TS_FAIL(c, TS_ERR_CUDA, cudaGetErrorString(e));
    }
    std::vector<uint32_t> host(words);
    int rc = TS_OK;
    e = cudaMemcpyAsync(dsegs, segs.data(), segs.size() * sizeof(opn::GatherSeg), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        KScope ks(c, TS_K_MISC);
        auto kfn = opn::gather_segments_kernel;
        TS_LAUNCH(kfn, (unsigned)segs.size(), 32, 0, c->stream, (const opn::GatherSeg *)dsegs, (uint32_t)segs.size(), dbuf);
        rc = check_launch(c, "gather_segments_kernel");
    }
    if (e == cudaSuccess && rc == TS_OK) e = cudaMemcpyAsync(host.data(), dbuf, words * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && rc == TS_OK) e = cudaStreamSynchronize(c->stream);
    pool_release(c, dsegs);
    pool_release(c, dbuf);
    if (e != cudaSuccess) TS_FAIL(c, TS_ERR_CUDA, cudaGetErrorString(e));
    if (rc != TS_OK) return rc;
    for (unsigned l = 0; l < t->log_n; l++) be_bytes(&host[8 * l], path_out + 32 * l);
    if (position_out) *position_out = host[8 * t->log_n];
    return TS_OK;
}
int ts_taptree_leaf_indices(ts_ctx *c, const ts_taptree *t, uint32_t *out_host) {
    TS_CUDA(c, cudaMemcpyAsync(out_host, t->leaf_idx, t->n_leaves * 4, cudaMemcpyDeviceToHost, c->stream));
    TS_CUDA(c, cudaStreamSynchronize(c->stream));
    return TS_OK;
}
int ts_taptree_level(ts_ctx *c, const ts_taptree *t, unsigned level, uint8_t *out_host) {
    if (level > t->log_n) TS_FAIL(c, TS_ERR_ARG, "taptree: level out of range");
    size_t off = 0;
    for (unsigned l = 0; l < level; l++) off += t->n_leaves >> l;
    const size_t cnt = t->n_leaves >> level;
    std::vector<uint32_t> w(cnt * 8);
    TS_CUDA(c, cudaMemcpyAsync(w.data(), t->nodes + off * 8, cnt * 32, cudaMemcpyDeviceToHost, c->stream));
    TS_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < cnt; i++) be_bytes(&w[8 * i], out_host + 32 * i);
    return TS_OK;
}

void ts_blake3_host(const uint8_t *in, size_t len, uint8_t out[32]) { hostb3::hash(in, len, out); }

}  // extern "C"
