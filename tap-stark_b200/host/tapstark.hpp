// tapstark.hpp -- C++ host-side mirror of the reference's Plonky3-derived interfaces over the C ABI
// (include/tapstark.h).  The reference is Rust; there is no Rust toolchain in the build image, so the host layer
// a `tapstark-gpu` crate would provide (tap-stark_b200/rust/) is mirrored here in C++ with the reference's names,
// argument meaning and error behaviour (a reference `panic!`/`expect` is a `tapstark::Panic` exception):
//
//   RowMajorMatrix<Val>          p3_matrix::dense::RowMajorMatrix                       (host, Montgomery u32)
//   GpuDft                       p3_dft::TwoAdicSubgroupDft<BabyBear>      fri/src/two_adic_pcs.rs:207,237-240
//   Blake3MerkleMmcs             basic::mmcs::bf_mmcs::BFMmcs<T>           basic/src/mmcs/bf_mmcs.rs:17-68
//   BfChallenger                 basic::challenger::BfChallenger           basic/src/challenger/mod.rs
//   FriConfig                    fri::FriConfig                            fri/src/config.rs:10-22
//   fold_even_odd                fri::fold_even_odd                        fri/src/fold_even_odd.rs:20-52
//   bf_commit_phase              fri::prover::bf_commit_phase              fri/src/prover.rs:93-141
//   TwoAdicFriPcs                impl Pcs for TwoAdicFriPcs: commit :227-258, open :260-419 (one ts_pcs_open call)
//   FriProof / BfQueryProof / BatchOpening   fri/src/proof.rs:13-33, decoded from the postcard bytes of ts_pcs_open
//
// Header only; link with libtapstark_b200.so.  All compute is in the CUDA library.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/tapstark.h"

namespace tapstark {

struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

constexpr uint32_t P = TS_P;
using Val = uint32_t;                      // BabyBear, Montgomery form (p3's in-memory representation)
using Challenge = std::array<uint32_t, 4>;  // BinomialExtensionField<BabyBear, 4>, Montgomery coefficients
using Digest = std::array<uint8_t, 32>;     // [[u8;4];8]

inline uint32_t to_monty(uint32_t x) { return (uint32_t)((((uint64_t)x) << 32) % P); }
inline uint32_t from_monty(uint32_t x) {
    // x * 2^-32 mod p ; 2^-32 mod p = 943718400
    return (uint32_t)(((uint64_t)x * 943718400ull) % P);
}

template <class T>
struct RowMajorMatrix {
    std::vector<T> values;
    size_t width_ = 0;
    RowMajorMatrix() = default;
    RowMajorMatrix(std::vector<T> v, size_t w) : values(std::move(v)), width_(w) {
        if (w == 0 || values.size() % w) throw Panic("RowMajorMatrix: values.len() % width != 0");
    }
    size_t width() const { return width_; }
    size_t height() const { return width_ ? values.size() / width_ : 0; }
    const T *row(size_t r) const { return values.data() + r * width_; }
};

class Context {
public:
    explicit Context(int device = 0, void *stream = nullptr) {
        if (ts_ctx_create(device, stream, &h_) != TS_OK)
            throw Panic("ts_ctx_create failed: no CUDA device (there is no CPU fallback)");
    }
    ~Context() { ts_ctx_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    ts_ctx *raw() const { return h_; }
    void check(int rc, const char *what) const {
        if (rc != TS_OK) throw Panic(std::string(what) + ": " + ts_last_error(h_));
    }

private:
    ts_ctx *h_ = nullptr;
};

// device-resident RowMajorMatrix<Val>
class DeviceMatrix {
public:
    DeviceMatrix(const Context &c, ts_matrix *m, bool owned = true) : c_(&c), m_(m), owned_(owned) {}
    DeviceMatrix(const Context &c, const RowMajorMatrix<Val> &host) : c_(&c) {
        c.check(ts_matrix_from_host(c.raw(), host.values.data(), host.height(), host.width(), &m_), "matrix_from_host");
        c.check(ts_ctx_synchronize(c.raw()), "synchronize");
    }
    DeviceMatrix(DeviceMatrix &&o) noexcept : c_(o.c_), m_(o.m_), owned_(o.owned_) { o.m_ = nullptr; }
    DeviceMatrix(const DeviceMatrix &) = delete;
    ~DeviceMatrix() {
        if (m_ && owned_) ts_matrix_free(m_);
    }
    ts_matrix *raw() const { return m_; }
    ts_matrix *release() {
        ts_matrix *m = m_;
        m_ = nullptr;
        return m;
    }
    size_t height() const { return ts_matrix_rows(m_); }
    size_t width() const { return ts_matrix_width(m_); }
    RowMajorMatrix<Val> to_row_major_matrix() const {
        std::vector<Val> v(height() * width());
        c_->check(ts_matrix_download(c_->raw(), m_, 0, height(), v.data()), "matrix_download");
        return RowMajorMatrix<Val>(std::move(v), width());
    }

private:
    const Context *c_;
    ts_matrix *m_ = nullptr;
    bool owned_ = true;
};

// p3_dft::TwoAdicSubgroupDft<BabyBear>.  Host-matrix methods have the trait's signatures; *_device keep the result
// on the GPU in committed (bit-reversed) order, the form TwoAdicFriPcs::commit wants.
class GpuDft {
public:
    explicit GpuDft(const Context &c) : c_(&c) {}
    RowMajorMatrix<Val> dft_batch(const RowMajorMatrix<Val> &mat) const { return unary(ts_dft_batch, mat); }
    RowMajorMatrix<Val> idft_batch(const RowMajorMatrix<Val> &mat) const { return unary(ts_idft_batch, mat); }
    std::vector<Val> dft(const std::vector<Val> &v) const { return dft_batch(RowMajorMatrix<Val>(v, 1)).values; }
    RowMajorMatrix<Val> coset_dft_batch(const RowMajorMatrix<Val> &mat, Val shift) const {
        DeviceMatrix in(*c_, mat);
        ts_matrix *o = nullptr;
        c_->check(ts_coset_dft_batch(c_->raw(), in.raw(), shift, &o), "coset_dft_batch");
        return DeviceMatrix(*c_, o).to_row_major_matrix();
    }
    // natural row order (trait semantics)
    RowMajorMatrix<Val> coset_lde_batch(const RowMajorMatrix<Val> &mat, size_t added_bits, Val shift) const {
        std::vector<Val> out((mat.height() << added_bits) * mat.width());
        c_->check(ts_coset_lde_batch_host(c_->raw(), mat.values.data(), mat.height(), mat.width(), (unsigned)added_bits, shift,
                                          1, out.data()),
                  "coset_lde_batch");
        return RowMajorMatrix<Val>(std::move(out), mat.width());
    }
    // committed order, device resident: `.coset_lde_batch(..).bit_reverse_rows().to_row_major_matrix()` with no extra pass
    DeviceMatrix coset_lde_batch_committed(const DeviceMatrix &evals, size_t added_bits, Val shift) const {
        ts_matrix *o = nullptr;
        c_->check(ts_coset_lde_batch(c_->raw(), evals.raw(), (unsigned)added_bits, shift, 0, &o), "coset_lde_batch");
        return DeviceMatrix(*c_, o);
    }

private:
    template <class F>
    RowMajorMatrix<Val> unary(F fn, const RowMajorMatrix<Val> &mat) const {
        DeviceMatrix in(*c_, mat);
        ts_matrix *o = nullptr;
        c_->check(fn(c_->raw(), in.raw(), &o), "dft");
        return DeviceMatrix(*c_, o).to_row_major_matrix();
    }
    const Context *c_;
};

// BFMmcs::ProverData
class ProverData {
public:
    ProverData(const Context &c, ts_tree *t) : c_(&c), t_(t) {}
    ProverData(ProverData &&o) noexcept : c_(o.c_), t_(o.t_) { o.t_ = nullptr; }
    ProverData(const ProverData &) = delete;
    ~ProverData() {
        if (t_) ts_tree_free(t_);
    }
    ts_tree *raw() const { return t_; }

private:
    const Context *c_;
    ts_tree *t_;
};

struct BatchOpeningProof {
    std::vector<Digest> siblings;
};

class Blake3MerkleMmcs {
public:
    explicit Blake3MerkleMmcs(const Context &c, int layout = TS_LAYOUT_P3_INJECT) : c_(&c), layout_(layout) {}
    int layout() const { return layout_; }
    // commit takes ownership of the matrices (bf_mmcs.rs:23), like the reference
    std::pair<Digest, ProverData> commit(std::vector<DeviceMatrix> inputs) const {
        std::vector<ts_matrix *> raw;
        for (auto &m : inputs) raw.push_back(m.release());
        Digest root{};
        ts_tree *t = nullptr;
        int rc = ts_mmcs_commit(c_->raw(), raw.data(), raw.size(), layout_, 1, root.data(), &t);
        if (rc != TS_OK)
            for (auto *m : raw) ts_matrix_free(m);
        c_->check(rc, "mmcs commit");
        return {root, ProverData(*c_, t)};
    }
    std::pair<Digest, ProverData> commit_matrix(DeviceMatrix input) const {
        std::vector<DeviceMatrix> v;
        v.push_back(std::move(input));
        return commit(std::move(v));
    }
    // (opened rows per matrix, proof); query_times_index of the reference selects one of its Taptrees: single tree here
    std::pair<std::vector<std::vector<Val>>, BatchOpeningProof> open_batch(size_t query_index, const ProverData &pd) const {
        const size_t k = ts_tree_num_matrices(pd.raw()), depth = ts_tree_depth(pd.raw());
        size_t total = 0;
        std::vector<size_t> widths(k);
        for (size_t i = 0; i < k; i++) total += widths[i] = ts_matrix_width(ts_tree_matrix(pd.raw(), i));
        std::vector<Val> rows(total);
        std::vector<uint8_t> path(32 * (depth ? depth : 1));
        c_->check(ts_mmcs_open_batch(c_->raw(), pd.raw(), query_index, rows.data(), path.data()), "open_batch");
        std::vector<std::vector<Val>> out;
        size_t o = 0;
        for (size_t i = 0; i < k; o += widths[i++]) out.emplace_back(rows.begin() + o, rows.begin() + o + widths[i]);
        BatchOpeningProof pr;
        pr.siblings.resize(depth);
        for (size_t l = 0; l < depth; l++) std::copy(path.begin() + 32 * l, path.begin() + 32 * l + 32, pr.siblings[l].begin());
        return {out, pr};
    }
    // Result<(), Error>: false = root mismatch
    bool verify_batch(const std::vector<size_t> &heights, const std::vector<std::vector<Val>> &opened, size_t query_index,
                      const BatchOpeningProof &proof, const Digest &root) const {
        std::vector<size_t> widths;
        std::vector<Val> flat;
        for (auto &r : opened) {
            widths.push_back(r.size());
            flat.insert(flat.end(), r.begin(), r.end());
        }
        std::vector<uint8_t> path(32 * (proof.siblings.empty() ? 1 : proof.siblings.size()));
        for (size_t l = 0; l < proof.siblings.size(); l++) std::copy(proof.siblings[l].begin(), proof.siblings[l].end(), path.begin() + 32 * l);
        return ts_mmcs_verify_batch(heights.data(), widths.data(), heights.size(), layout_, query_index, flat.data(), path.data(),
                                    proof.siblings.size(), root.data()) == TS_OK;
    }
    std::vector<DeviceMatrix> get_matrices(const ProverData &pd) const {
        std::vector<DeviceMatrix> v;
        for (size_t i = 0; i < ts_tree_num_matrices(pd.raw()); i++) v.emplace_back(*c_, ts_tree_matrix(pd.raw(), i), false);
        return v;
    }
    size_t get_max_height(const ProverData &pd) const { return ts_tree_max_height(pd.raw()); }
    const Context &ctx() const { return *c_; }

private:
    const Context *c_;
    int layout_;
};

// The commitment TapTreeMmcs computes (basic/src/tcs/mod.rs:238-282) over a script template; the template (bit-commitment locking
// scripts around pushed integers) is host work of the caller, as the key generation is in the reference.
struct ScriptTemplate {
    std::vector<std::vector<uint8_t>> segments;  // n_push + 1 constant byte runs
    std::vector<uint32_t> push_word;             // word of the leaf row behind push k + 1 (push 0 is the leaf index)
};
class TapTree {
public:
    // PolyTCS::padding_matrix on the device: the rows the leaves commit to when matrices of several heights share a tree
    static DeviceMatrix padded_rows(const Context &c, const std::vector<const DeviceMatrix *> &mats) {
        std::vector<const ts_matrix *> raw;
        for (auto *m : mats) raw.push_back(m->raw());
        ts_matrix *o = nullptr;
        c.check(ts_padded_leaf_rows(c.raw(), raw.data(), raw.size(), &o), "padded_leaf_rows");
        return DeviceMatrix(c, o);
    }
    TapTree(const Context &c, const DeviceMatrix &leaf_rows, const ScriptTemplate &tpl) : c_(&c) {
        if (tpl.segments.size() != tpl.push_word.size() + 2) throw Panic("taptree: need one segment more than pushes");
        std::vector<uint8_t> blob;
        std::vector<size_t> off{0};
        for (auto &sgm : tpl.segments) {
            blob.insert(blob.end(), sgm.begin(), sgm.end());
            off.push_back(blob.size());
        }
        c.check(ts_taptree_commit(c.raw(), leaf_rows.raw(), blob.data(), off.data(), tpl.push_word.data(), tpl.push_word.size() + 1,
                                  root_.data(), &t_), "taptree_commit");
        n_leaves_ = leaf_rows.height();
    }
    TapTree(TapTree &&o) noexcept : c_(o.c_), t_(o.t_), root_(o.root_), n_leaves_(o.n_leaves_) { o.t_ = nullptr; }
    TapTree(const TapTree &) = delete;
    ~TapTree() {
        if (t_) ts_taptree_free(t_);
    }
    const Digest &root() const { return root_; }
    // CompleteTaptree::leaf_indices (reverse_idx_dict, builder.rs:96-102)
    std::vector<uint32_t> leaf_indices() const {
        std::vector<uint32_t> v(n_leaves_);
        c_->check(ts_taptree_leaf_indices(c_->raw(), t_, v.data()), "taptree_leaf_indices");
        return v;
    }
    // (TaprootMerkleBranch of Merkle leaf `index`, its position among the TapTree's leaves)
    std::pair<std::vector<Digest>, uint32_t> open(size_t index) const {
        size_t depth = 0;
        while (((size_t)1 << depth) < n_leaves_) depth++;
        std::vector<uint8_t> path(32 * (depth ? depth : 1));
        uint32_t pos = 0;
        c_->check(ts_taptree_open(c_->raw(), t_, index, path.data(), &pos), "taptree_open");
        std::vector<Digest> br(depth);
        for (size_t l = 0; l < depth; l++) std::copy(path.begin() + 32 * l, path.begin() + 32 * l + 32, br[l].begin());
        return {br, pos};
    }

private:
    const Context *c_;
    ts_taptree *t_ = nullptr;
    Digest root_{};
    size_t n_leaves_ = 0;
};

// BfChallenger<Challenge, U32, Blake3Permutation, 16>
class BfChallenger {
public:
    BfChallenger() { ts_challenger_new(&h_); }
    BfChallenger(const BfChallenger &o) { ts_challenger_clone(o.h_, &h_); }
    BfChallenger &operator=(const BfChallenger &) = delete;
    ~BfChallenger() { ts_challenger_free(h_); }
    void observe(const std::array<uint8_t, 4> &v) { ts_challenger_observe(h_, v.data()); }
    void observe(const Digest &commitment) { ts_challenger_observe_digest(h_, commitment.data()); }
    Challenge sample() {  // canonical transcript values -> Montgomery field element
        uint32_t c[4];
        ts_challenger_sample_ext(h_, c);
        return {to_monty(c[0]), to_monty(c[1]), to_monty(c[2]), to_monty(c[3])};
    }
    size_t sample_bits(size_t bits) { return ts_challenger_sample_bits(h_, (unsigned)bits, 1); }
    uint32_t grind(size_t bits) {
        uint32_t w = 0;
        if (ts_challenger_grind(h_, (unsigned)bits, 1, &w) != TS_OK) throw Panic("failed to find witness");
        return w;
    }
    bool check_witness(size_t bits, uint32_t witness) { return ts_challenger_check_witness(h_, (unsigned)bits, witness, 1) != 0; }
    ts_challenger *raw() const { return h_; }

private:
    ts_challenger *h_ = nullptr;
};

struct FriConfig {
    size_t log_blowup, num_queries, proof_of_work_bits;
    const Blake3MerkleMmcs *mmcs;
    size_t blowup() const { return (size_t)1 << log_blowup; }
};

// fri::fold_even_odd on host vectors (bit-reversed evaluations in, folded vector out)
inline std::vector<Challenge> fold_even_odd(const Context &c, const std::vector<Challenge> &poly, const Challenge &beta) {
    std::vector<Challenge> out(poly.size() / 2);
    c.check(ts_fri_fold_ext_host(c.raw(), poly.data()->data(), out.size(), beta.data(), out.data()->data()), "fold_even_odd");
    return out;
}

struct CommitPhaseResult {
    std::vector<Digest> commits;
    std::vector<ProverData> data;
    Challenge final_poly;  // canonical
};

// fri::prover::bf_commit_phase; inputs: extension vectors (rows x 4), lengths strictly descending
inline CommitPhaseResult bf_commit_phase(const FriConfig &config, const std::vector<const DeviceMatrix *> &inputs,
                                         BfChallenger &challenger) {
    const Context &c = config.mmcs->ctx();
    std::vector<ts_matrix *> raw;
    for (auto *m : inputs) raw.push_back(m->raw());
    size_t max_rounds = 0;
    while (((size_t)1 << (max_rounds + config.log_blowup)) < inputs[0]->height()) max_rounds++;
    std::vector<uint8_t> commits(32 * (max_rounds ? max_rounds : 1));
    std::vector<ts_tree *> trees(max_rounds ? max_rounds : 1, nullptr);
    CommitPhaseResult res;
    size_t rounds = 0;
    int rc = ts_fri_commit_phase(c.raw(), raw.data(), raw.size(), (unsigned)config.log_blowup, challenger.raw(), commits.data(),
                                 trees.data(), res.final_poly.data(), &rounds);
    for (size_t i = 0; i < rounds; i++) {
        if (trees[i]) res.data.emplace_back(c, trees[i]);
        Digest d;
        std::copy(commits.begin() + 32 * i, commits.begin() + 32 * i + 32, d.begin());
        res.commits.push_back(d);
    }
    if (rc == TS_ERR_NOT_CONSTANT) throw Panic("assertion failed: final layer is not constant");  // prover.rs:130-134
    c.check(rc, "bf_commit_phase");
    return res;
}

// ---- proof objects (fri/src/proof.rs:13-33; canonical u32 values, as serialised) -------------------------------------
struct BatchOpening {
    std::vector<std::vector<Val>> opened_values;  // canonical
    BatchOpeningProof opening_proof;
};
struct CommitPhaseStep {
    std::vector<std::vector<Challenge>> opened_rows;  // one matrix: [[e0, e1]], canonical
    BatchOpeningProof opening_proof;
};
struct BfQueryProof {
    std::vector<BatchOpening> input_proof;
    std::vector<CommitPhaseStep> commit_phase_openings;
};
struct FriProof {
    std::vector<Digest> commit_phase_commits;
    std::vector<BfQueryProof> query_proofs;
    Challenge final_poly;  // canonical
    uint32_t pow_witness;
};
// all_opened_values[round][matrix][point][column], canonical BabyBear^4
using OpenedValues = std::vector<std::vector<std::vector<std::vector<Challenge>>>>;

// postcard reader (include/tapstark.h: ts_pcs_open documents the layout)
class PostcardReader {
public:
    PostcardReader(const uint8_t *p, size_t n) : p_(p), n_(n) {}
    uint64_t varint() {
        uint64_t v = 0;
        for (int s = 0;; s += 7) {
            if (o_ >= n_ || s > 63) throw Panic("postcard: truncated varint");
            const uint8_t x = p_[o_++];
            v |= (uint64_t)(x & 0x7f) << s;
            if (x < 0x80) return v;
        }
    }
    Challenge ef() { return {(uint32_t)varint(), (uint32_t)varint(), (uint32_t)varint(), (uint32_t)varint()}; }
    Digest digest() {
        if (o_ + 32 > n_) throw Panic("postcard: truncated digest");
        Digest d;
        std::copy(p_ + o_, p_ + o_ + 32, d.begin());
        o_ += 32;
        return d;
    }
    BatchOpeningProof path() {
        BatchOpeningProof pr;
        pr.siblings.resize(varint());
        for (auto &d : pr.siblings) d = digest();
        return pr;
    }
    bool done() const { return o_ == n_; }

private:
    const uint8_t *p_;
    size_t n_, o_ = 0;
};
inline std::pair<OpenedValues, FriProof> decode_opening(const std::vector<uint8_t> &bytes) {
    PostcardReader r(bytes.data(), bytes.size());
    OpenedValues ov(r.varint());
    for (auto &rnd : ov) {
        rnd.resize(r.varint());
        for (auto &mat : rnd) {
            mat.resize(r.varint());
            for (auto &pt : mat) {
                pt.resize(r.varint());
                for (auto &y : pt) y = r.ef();
            }
        }
    }
    FriProof pr;
    pr.commit_phase_commits.resize(r.varint());
    for (auto &d : pr.commit_phase_commits) d = r.digest();
    pr.query_proofs.resize(r.varint());
    for (auto &q : pr.query_proofs) {
        q.input_proof.resize(r.varint());
        for (auto &bo : q.input_proof) {
            bo.opened_values.resize(r.varint());
            for (auto &row : bo.opened_values) {
                row.resize(r.varint());
                for (auto &v : row) v = (uint32_t)r.varint();
            }
            bo.opening_proof = r.path();
        }
        q.commit_phase_openings.resize(r.varint());
        for (auto &st : q.commit_phase_openings) {
            st.opened_rows.resize(r.varint());
            for (auto &row : st.opened_rows) {
                row.resize(r.varint());
                for (auto &e : row) e = r.ef();
            }
            st.opening_proof = r.path();
        }
    }
    pr.final_poly = r.ef();
    pr.pow_witness = (uint32_t)r.varint();
    if (!r.done()) throw Panic("postcard: trailing bytes after the opening proof");
    return {std::move(ov), std::move(pr)};
}

struct TwoAdicMultiplicativeCoset {
    size_t log_n;
    Val shift;  // Montgomery
    size_t size() const { return (size_t)1 << log_n; }
};

class TwoAdicFriPcs {
public:
    TwoAdicFriPcs(const GpuDft &dft, const Blake3MerkleMmcs &mmcs, FriConfig fri) : dft_(&dft), mmcs_(&mmcs), fri_(fri) {}
    TwoAdicMultiplicativeCoset natural_domain_for_degree(size_t degree) const {
        size_t l = 0;
        while (((size_t)1 << l) < degree) l++;
        if (((size_t)1 << l) != degree) throw Panic("log2_strict_usize: not a power of two");
        return {l, to_monty(1)};
    }
    // two_adic_pcs.rs:227-245 on host matrices (what uni_stark::prove hands over)
    std::pair<Digest, ProverData> commit(const std::vector<std::pair<TwoAdicMultiplicativeCoset, RowMajorMatrix<Val>>> &evals) const {
        const Context &c = mmcs_->ctx();
        std::vector<const uint32_t *> ptrs;
        std::vector<size_t> rows, widths;
        std::vector<uint32_t> shifts;
        for (auto &e : evals) {
            if (e.first.size() != e.second.height()) throw Panic("assertion failed: domain.size() == evals.height()");  // :234
            ptrs.push_back(e.second.values.data());
            rows.push_back(e.second.height());
            widths.push_back(e.second.width());
            shifts.push_back(e.first.shift);
        }
        Digest root{};
        ts_tree *t = nullptr;
        c.check(ts_pcs_commit_host(c.raw(), ptrs.data(), rows.data(), widths.data(), shifts.data(), evals.size(),
                                   (unsigned)fri_.log_blowup, mmcs_->layout(), root.data(), &t),
                "pcs commit");
        return {root, ProverData(c, t)};
    }
    // two_adic_pcs.rs:247-258
    RowMajorMatrix<Val> get_evaluations_on_domain(const ProverData &pd, size_t idx, TwoAdicMultiplicativeCoset domain) const {
        if (domain.shift != TS_GENERATOR_MONTY) throw Panic("assertion failed: domain.shift == Val::generator()");
        const Context &c = mmcs_->ctx();
        const size_t w = ts_matrix_width(ts_tree_matrix(pd.raw(), idx));
        std::vector<Val> v(domain.size() * w);
        c.check(ts_pcs_get_evaluations_on_domain(c.raw(), pd.raw(), idx, domain.size(), v.data()), "get_evaluations_on_domain");
        return RowMajorMatrix<Val>(std::move(v), w);
    }
    // two_adic_pcs.rs:260-419 + fri/src/prover.rs:19-90.  rounds: (prover data, per matrix its opening points in Montgomery
    // form).  The serialised `(OpenedValues, FriProof)` exactly as the library emits it:
    std::vector<uint8_t> open_bytes(const std::vector<std::pair<const ProverData *, std::vector<std::vector<Challenge>>>> &rounds,
                                    BfChallenger &challenger) const {
        const Context &c = mmcs_->ctx();
        std::vector<const ts_tree *> trees;
        std::vector<size_t> counts;
        std::vector<uint32_t> pts;
        for (auto &r : rounds) {
            if (r.second.size() != ts_tree_num_matrices(r.first->raw())) throw Panic("open: one point list per committed matrix");
            trees.push_back(r.first->raw());
            for (auto &per_mat : r.second) {
                counts.push_back(per_mat.size());
                for (auto &z : per_mat) pts.insert(pts.end(), z.begin(), z.end());
            }
        }
        uint8_t *buf = nullptr;
        size_t n = 0;
        c.check(ts_pcs_open(c.raw(), trees.data(), trees.size(), counts.data(), pts.data(), (unsigned)fri_.log_blowup,
                            (unsigned)fri_.num_queries, (unsigned)fri_.proof_of_work_bits, challenger.raw(), &buf, &n),
                "pcs open");
        std::vector<uint8_t> out(buf, buf + n);
        ts_bytes_free(buf);
        return out;
    }
    std::pair<OpenedValues, FriProof> open(const std::vector<std::pair<const ProverData *, std::vector<std::vector<Challenge>>>> &rounds,
                                           BfChallenger &challenger) const {
        return decode_opening(open_bytes(rounds, challenger));
    }
    const FriConfig &fri() const { return fri_; }
    const Blake3MerkleMmcs &mmcs() const { return *mmcs_; }

private:
    const GpuDft *dft_;
    const Blake3MerkleMmcs *mmcs_;
    FriConfig fri_;
};

}  // namespace tapstark
