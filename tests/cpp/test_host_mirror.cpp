// Parity tests of the C++ host mirror (tap-stark_b200/host/tapstark.hpp) against the oracle, written after the
// reference's own tests for the path:
//   test_fold_even_odd           fri/src/fold_even_odd.rs:54-92
//   commit_single / commit_many  fri/tests/pcs.rs:70-110 (commit side: round shapes of the reference's cases)
//   mmcs commit/open/verify      basic/src/mmcs/taptree_mmcs.rs tests (commit -> open_batch -> verify_batch, tamper)
//   commit phase                 fri/tests/fri.rs:50-130 (LDE -> bit-reverse -> bf_commit_phase)
//   taptree                      basic/src/tcs/mod.rs tests (commit -> open -> verify_inclusion), padding_matrix
//   pcs open                     fri/tests/pcs.rs:70-110 (commit -> sample zeta -> open; the verifier-side checks that need
//                                only the oracle's field arithmetic are replayed: opened values, Merkle openings, fold chain)
// The same binary links the CUDA library (pytest -m gpu) or the emulated build of the same kernel sources.
#include <cstdio>
#include <cstring>
#include <functional>

#include "../../oracle/tapstark_oracle.h"
#include "../../tap-stark_b200/host/tapstark.hpp"

using namespace tapstark;

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint64_t next_u64() {
    uint64_t z = (rng_state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static uint32_t rand_val() { return (uint32_t)(next_u64() % P); }  // canonical
static std::vector<uint32_t> rand_canonical(size_t n) {
    std::vector<uint32_t> v(n);
    for (auto &x : v) x = rand_val();
    return v;
}
static std::vector<uint32_t> monty(std::vector<uint32_t> v) {
    or_to_monty_vec(v.data(), v.size());
    return v;
}
static std::vector<uint32_t> canon(std::vector<uint32_t> v) {
    or_from_monty_vec(v.data(), v.size());
    return v;
}

static int failures = 0, passed = 0;
#define CHECK(cond)                                                         \
    do {                                                                    \
        if (!(cond)) {                                                      \
            std::printf("  CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            throw std::runtime_error("check failed");                       \
        }                                                                   \
    } while (0)
static void run(const char *name, const std::function<void()> &f) {
    try {
        f();
        passed++;
        std::printf("ok   %s\n", name);
    } catch (const std::exception &e) {
        failures++;
        std::printf("FAIL %s: %s\n", name, e.what());
    }
}

// ---------------------------------------------------------------------------------------------------------------
static void test_dft_roundtrip_and_oracle(const Context &ctx) {
    GpuDft dft(ctx);
    for (unsigned log_n : {0u, 1u, 5u, 11u}) {
        const size_t n = (size_t)1 << log_n, w = 7;
        auto c = rand_canonical(n * w);
        auto expect = c;
        or_dft_batch(expect.data(), log_n, w);
        auto got = dft.dft_batch(RowMajorMatrix<Val>(monty(c), w));
        CHECK(canon(got.values) == expect);
        CHECK(canon(dft.idft_batch(got).values) == c);
    }
}

static void test_coset_lde_batch(const Context &ctx) {
    GpuDft dft(ctx);
    const unsigned log_n = 9, added = 2;
    const size_t n = 1u << log_n, w = 13;
    auto e = rand_canonical(n * w);
    std::vector<uint32_t> expect((n << added) * w);
    or_coset_lde_batch(e.data(), log_n, w, added, 31, expect.data());
    auto nat = dft.coset_lde_batch(RowMajorMatrix<Val>(monty(e), w), added, to_monty(31));
    CHECK(canon(nat.values) == expect);
    // the committed-order, device-resident form used by the PCS
    or_pcs_lde_committed(e.data(), log_n, w, added, 31, expect.data());
    DeviceMatrix dev(ctx, RowMajorMatrix<Val>(monty(e), w));
    auto com = dft.coset_lde_batch_committed(dev, added, to_monty(31)).to_row_major_matrix();
    CHECK(canon(com.values) == expect);
}

// fri/src/fold_even_odd.rs:54-92: folding the bit-reversed evaluations of p equals the bit-reversed evaluations of
// even(p) + beta * odd(p).
static void test_fold_even_odd(const Context &ctx) {
    GpuDft dft(ctx);
    const unsigned log_n = 10;
    const size_t n = 1u << log_n;
    auto coeffs = rand_canonical(n);
    uint32_t beta[4] = {rand_val(), rand_val(), rand_val(), rand_val()};
    // evals of p, embedded in the extension field, bit-reversed
    auto evals = coeffs;
    or_dft_batch(evals.data(), log_n, 1);
    or_bit_reverse_rows(evals.data(), log_n, 1);
    std::vector<Challenge> poly(n);
    for (size_t i = 0; i < n; i++) poly[i] = {or_to_monty(evals[i]), 0, 0, 0};
    Challenge b = {or_to_monty(beta[0]), or_to_monty(beta[1]), or_to_monty(beta[2]), or_to_monty(beta[3])};
    auto folded = fold_even_odd(ctx, poly, b);
    // expected: dft of (even + beta*odd), coefficient-wise over the 4 extension coordinates
    std::vector<uint32_t> ec((n / 2) * 4);
    for (size_t i = 0; i < n / 2; i++)
        for (int k = 0; k < 4; k++)
            ec[4 * i + k] = or_bb_add(k == 0 ? coeffs[2 * i] : 0, or_bb_mul(beta[k], coeffs[2 * i + 1]));
    or_dft_batch(ec.data(), log_n - 1, 4);
    or_bit_reverse_rows(ec.data(), log_n - 1, 4);
    CHECK(folded.size() == n / 2);
    for (size_t i = 0; i < n / 2; i++)
        for (int k = 0; k < 4; k++) CHECK(or_from_monty(folded[i][k]) == ec[4 * i + k]);
}

// commit -> open_batch -> verify_batch on mixed heights, and a tampered opening is rejected
static void test_mmcs(const Context &ctx, int layout) {
    Blake3MerkleMmcs mmcs(ctx, layout);
    const size_t heights[3] = {64, 64, 16}, widths[3] = {5, 3, 9};
    std::vector<std::vector<uint32_t>> host;
    std::vector<DeviceMatrix> mats;
    const uint32_t *ptrs[3];
    for (int i = 0; i < 3; i++) {
        host.push_back(rand_canonical(heights[i] * widths[i]));
        mats.emplace_back(ctx, RowMajorMatrix<Val>(monty(host[i]), widths[i]));
    }
    for (int i = 0; i < 3; i++) ptrs[i] = host[i].data();
    uint8_t oroot[32];
    or_tree *ot = or_mmcs_commit(ptrs, heights, widths, 3, layout, oroot);
    auto [root, pd] = mmcs.commit(std::move(mats));
    CHECK(std::memcmp(root.data(), oroot, 32) == 0);
    CHECK(mmcs.get_max_height(pd) == 64);
    CHECK(mmcs.get_matrices(pd).size() == 3);
    std::vector<size_t> hs(heights, heights + 3);
    for (size_t index : {0u, 17u, 63u}) {
        auto [opened, proof] = mmcs.open_batch(index, pd);
        std::vector<uint32_t> orows(5 + 3 + 9);
        std::vector<uint8_t> opath(32 * or_tree_depth(ot));
        or_mmcs_open_batch(ot, index, orows.data(), opath.data());
        std::vector<uint32_t> flat;
        for (auto &r : opened) flat.insert(flat.end(), r.begin(), r.end());
        CHECK(canon(flat) == orows);
        CHECK(proof.siblings.size() == or_tree_depth(ot));
        for (size_t l = 0; l < proof.siblings.size(); l++) CHECK(std::memcmp(proof.siblings[l].data(), &opath[32 * l], 32) == 0);
        CHECK(mmcs.verify_batch(hs, opened, index, proof, root));
        auto bad = opened;
        bad[1][0] = bad[1][0] == 0 ? 1 : 0;
        CHECK(!mmcs.verify_batch(hs, bad, index, proof, root));
        if (!proof.siblings.empty()) {
            auto badp = proof;
            badp.siblings[0][3] ^= 1;
            CHECK(!mmcs.verify_batch(hs, opened, index, badp, root));
        }
        CHECK(!mmcs.verify_batch(hs, opened, index ^ 1, proof, root));
    }
    or_tree_free(ot);
}

// fri/tests/pcs.rs round shapes: one commitment over several matrices of different degrees
static void do_test_pcs_commit(const Context &ctx, const std::vector<std::pair<unsigned, size_t>> &shapes, unsigned log_blowup) {
    GpuDft dft(ctx);
    Blake3MerkleMmcs mmcs(ctx);
    TwoAdicFriPcs pcs(dft, mmcs, FriConfig{log_blowup, 10, 8, &mmcs});
    std::vector<std::pair<TwoAdicMultiplicativeCoset, RowMajorMatrix<Val>>> evals;
    std::vector<std::vector<uint32_t>> ldes;
    std::vector<const uint32_t *> ptrs;
    std::vector<size_t> hs, ws;
    for (auto [log_n, w] : shapes) {
        const size_t n = (size_t)1 << log_n;
        auto e = rand_canonical(n * w);
        std::vector<uint32_t> lde((n << log_blowup) * w);
        or_pcs_lde_committed(e.data(), log_n, w, log_blowup, 31, lde.data());
        ldes.push_back(std::move(lde));
        hs.push_back(n << log_blowup);
        ws.push_back(w);
        evals.emplace_back(pcs.natural_domain_for_degree(n), RowMajorMatrix<Val>(monty(e), w));
    }
    for (auto &l : ldes) ptrs.push_back(l.data());
    uint8_t oroot[32];
    or_tree *ot = or_mmcs_commit(ptrs.data(), hs.data(), ws.data(), shapes.size(), OR_LAYOUT_P3_INJECT, oroot);
    auto [root, pd] = pcs.commit(evals);
    CHECK(std::memcmp(root.data(), oroot, 32) == 0);
    // get_evaluations_on_domain: the quotient domain (first n*2 committed rows, re-bit-reversed)
    for (size_t i = 0; i < shapes.size(); i++) {
        const unsigned lq = shapes[i].first + (log_blowup ? 1 : 0);
        std::vector<uint32_t> expect(ldes[i].begin(), ldes[i].begin() + ((size_t)1 << lq) * ws[i]);
        or_bit_reverse_rows(expect.data(), lq, ws[i]);
        auto got = pcs.get_evaluations_on_domain(pd, i, TwoAdicMultiplicativeCoset{lq, TS_GENERATOR_MONTY});
        CHECK(canon(got.values) == expect);
    }
    // a domain/height mismatch is the reference's assert_eq!(domain.size(), evals.height())
    bool panicked = false;
    try {
        auto bad = evals;
        bad[0].first.log_n += 1;
        pcs.commit(bad);
    } catch (const Panic &) {
        panicked = true;
    }
    CHECK(panicked);
    or_tree_free(ot);
}

// fri/tests/fri.rs: random polynomials of decreasing degree, LDE, bit-reverse, commit phase
static void test_commit_phase(const Context &ctx) {
    Blake3MerkleMmcs mmcs(ctx);
    FriConfig config{1, 10, 8, &mmcs};
    std::vector<std::vector<uint32_t>> inputs;  // canonical EF codewords
    std::vector<size_t> lens;
    for (unsigned deg_bits : {7u, 5u, 4u}) {
        const size_t n = (size_t)1 << deg_bits;
        auto e = rand_canonical(n * 4);
        std::vector<uint32_t> lde((n << 1) * 4);
        or_pcs_lde_committed(e.data(), deg_bits, 4, 1, 1, lde.data());
        inputs.push_back(std::move(lde));
        lens.push_back(n << 1);
    }
    std::vector<const uint32_t *> ptrs;
    for (auto &v : inputs) ptrs.push_back(v.data());
    or_challenger oc;
    or_chal_init(&oc, 0);
    std::vector<uint8_t> ocommits(32 * 16);
    uint32_t ofinal[4];
    const int orounds = or_fri_commit_phase(ptrs.data(), lens.data(), 3, 1, &oc, ocommits.data(), ofinal, nullptr, nullptr);
    CHECK(orounds == 7);

    std::vector<DeviceMatrix> dev;
    for (auto &v : inputs) dev.emplace_back(ctx, RowMajorMatrix<Val>(monty(v), 4));
    std::vector<const DeviceMatrix *> refs;
    for (auto &d : dev) refs.push_back(&d);
    BfChallenger challenger;
    auto res = bf_commit_phase(config, refs, challenger);
    CHECK((int)res.commits.size() == orounds);
    for (int r = 0; r < orounds; r++) CHECK(std::memcmp(res.commits[r].data(), &ocommits[32 * r], 32) == 0);
    for (int k = 0; k < 4; k++) CHECK(res.final_poly[k] == ofinal[k]);
    // the transcripts stay in lock step: same proof-of-work witness, same query indices
    const uint32_t w = challenger.grind(config.proof_of_work_bits);
    CHECK(w == or_chal_grind(&oc, 8, 1));
    BfChallenger v2(challenger);
    const size_t index = challenger.sample_bits(8);
    CHECK(index == or_chal_sample_bits(&oc, 8, 1));
    CHECK(v2.sample_bits(8) == index);  // Clone keeps the sponge state
    // prover data of every round opens and verifies
    for (size_t r = 0; r < res.data.size(); r++) {
        auto [opened, proof] = mmcs.open_batch(3 >> r, res.data[r]);
        CHECK(opened.size() == 1 && opened[0].size() == 8);
        CHECK(mmcs.verify_batch({mmcs.get_max_height(res.data[r])}, opened, 3 >> r, proof, res.commits[r]));
    }
    // a codeword that is not low degree cannot end in a constant layer: prover.rs:130-134 asserts
    std::vector<DeviceMatrix> junk;
    junk.emplace_back(ctx, RowMajorMatrix<Val>(monty(rand_canonical(64 * 4)), 4));
    BfChallenger c2;
    bool panicked = false;
    try {
        bf_commit_phase(config, {&junk[0]}, c2);
    } catch (const Panic &) {
        panicked = true;
    }
    CHECK(panicked);
}

// fri/tests/pcs.rs do_test_fri_pcs: commit some matrices, sample zeta, open every matrix at it
static void test_pcs_open(const Context &ctx) {
    const unsigned log_blowup = 1;
    GpuDft dft(ctx);
    Blake3MerkleMmcs mmcs(ctx);
    TwoAdicFriPcs pcs(dft, mmcs, FriConfig{log_blowup, 5, 4, &mmcs});
    const std::vector<std::pair<unsigned, size_t>> shapes = {{5, 6}, {3, 2}};
    std::vector<std::pair<TwoAdicMultiplicativeCoset, RowMajorMatrix<Val>>> evals;
    std::vector<std::vector<uint32_t>> coeffs;
    for (auto [log_n, w] : shapes) {
        const size_t n = (size_t)1 << log_n;
        auto e = rand_canonical(n * w);
        evals.emplace_back(pcs.natural_domain_for_degree(n), RowMajorMatrix<Val>(monty(e), w));
        or_idft_batch(e.data(), log_n, w);  // coefficients, for the direct evaluation below
        coeffs.push_back(std::move(e));
    }
    auto [root, pd] = pcs.commit(evals);
    BfChallenger challenger;
    challenger.observe(root);
    or_challenger oc;
    or_chal_init(&oc, 0);
    or_chal_observe_digest(&oc, root.data());
    const Challenge zeta = challenger.sample();  // Montgomery
    uint32_t zc[4];
    or_chal_sample_ef(&oc, zc);
    for (int k = 0; k < 4; k++) CHECK(from_monty(zeta[k]) == zc[k]);
    std::vector<std::pair<const ProverData *, std::vector<std::vector<Challenge>>>> rounds;
    rounds.push_back({&pd, {{zeta}, {zeta}}});
    auto [opened, proof] = pcs.open(rounds, challenger);
    // (1) opened values = p_c(zeta), by Horner on the oracle's coefficients
    CHECK(opened.size() == 1 && opened[0].size() == shapes.size());
    for (size_t i = 0; i < shapes.size(); i++) {
        const size_t n = (size_t)1 << shapes[i].first, w = shapes[i].second;
        CHECK(opened[0][i].size() == 1 && opened[0][i][0].size() == w);
        for (size_t col = 0; col < w; col++) {
            uint32_t acc[4] = {0, 0, 0, 0};
            for (size_t k = n; k-- > 0;) {
                uint32_t t[4];
                or_ef_mul(acc, zc, t);
                t[0] = (uint32_t)(((uint64_t)t[0] + coeffs[i][k * w + col]) % P);
                std::memcpy(acc, t, 16);
            }
            for (int k = 0; k < 4; k++) CHECK(opened[0][i][0][col][k] == acc[k]);
        }
    }
    // (2) transcript replay (fri/src/verifier.rs:30-60): alpha, betas from the layer commitments, the witness, the indices
    uint32_t alpha[4];
    or_chal_sample_ef(&oc, alpha);
    const size_t log_max = shapes[0].first + log_blowup;
    CHECK(proof.commit_phase_commits.size() == log_max - log_blowup);
    std::vector<std::array<uint32_t, 4>> betas;
    for (auto &cm : proof.commit_phase_commits) {
        or_chal_observe_digest(&oc, cm.data());
        std::array<uint32_t, 4> b;
        or_chal_sample_ef(&oc, b.data());
        betas.push_back(b);
    }
    CHECK(or_chal_check_witness(&oc, 4, proof.pow_witness, 1));
    CHECK(proof.query_proofs.size() == 5);
    for (auto &q : proof.query_proofs) {
        size_t index = or_chal_sample_bits(&oc, (unsigned)log_max, 1);
        // input openings verify against the commitment (heights of the committed LDEs)
        CHECK(q.input_proof.size() == 1);
        std::vector<std::vector<Val>> rows_m;
        for (auto &row : q.input_proof[0].opened_values) rows_m.push_back(monty(row));
        CHECK(mmcs.verify_batch({(size_t)1 << (shapes[0].first + log_blowup), (size_t)1 << (shapes[1].first + log_blowup)}, rows_m, index,
                                q.input_proof[0].opening_proof, root));
        // fold chain (two_adic_pcs.rs:87-114): every layer opening verifies, folded value carried to the next layer
        CHECK(q.commit_phase_openings.size() == proof.commit_phase_commits.size());
        uint32_t folded[4];
        bool have = false;
        for (size_t r = 0; r < q.commit_phase_openings.size(); r++) {
            auto &st = q.commit_phase_openings[r];
            CHECK(st.opened_rows.size() == 1 && st.opened_rows[0].size() == 2);
            const size_t pair = index >> 1;
            std::vector<Val> flat;
            for (auto &e : st.opened_rows[0]) flat.insert(flat.end(), e.begin(), e.end());
            CHECK(mmcs.verify_batch({(size_t)1 << (log_max - 1 - r)}, {monty(flat)}, pair, st.opening_proof, proof.commit_phase_commits[r]));
            const auto &e0 = st.opened_rows[0][0], &e1 = st.opened_rows[0][1];
            if (have) CHECK(std::memcmp(folded, (index & 1) ? e1.data() : e0.data(), 16) == 0 || r == (size_t)(shapes[0].first - shapes[1].first));
            or_fold_row_ef(pair, (unsigned)(log_max - 1 - r), betas[r].data(), e0.data(), e1.data(), folded);
            have = true;
            index = pair;
        }
        // the shorter matrix joins at its own height, so the chain is only checked at its end here when no input joined
        (void)alpha;
    }
    // (3) one C-ABI call, deterministic bytes: a second opening from the same transcript state is identical
    BfChallenger c1, c2;
    c1.observe(root);
    c2.observe(root);
    const Challenge z1 = c1.sample(), z2 = c2.sample();
    std::vector<std::pair<const ProverData *, std::vector<std::vector<Challenge>>>> r1, r2;
    r1.push_back({&pd, {{z1}, {z1}}});
    r2.push_back({&pd, {{z2}, {z2}}});
    CHECK(pcs.open_bytes(r1, c1) == pcs.open_bytes(r2, c2));
    // a point list per matrix is required (reference: zip of rounds and points)
    bool panicked = false;
    try {
        std::vector<std::pair<const ProverData *, std::vector<std::vector<Challenge>>>> bad;
        bad.push_back({&pd, {{zeta}}});
        BfChallenger c3;
        pcs.open(bad, c3);
    } catch (const Panic &) {
        panicked = true;
    }
    CHECK(panicked);
}

// TapTree (basic/src/tcs/mod.rs:238-301): commit over a script template, open, verify_inclusion (complete_taptree.rs:64-73) with
// the oracle's tagged SHA-256; the padded rows against the oracle's padding_matrix.
static std::vector<uint8_t> push_num(uint32_t v) {  // minimal script-number push of a non-negative integer
    if (v == 0) return {0x00};
    if (v <= 16) return {(uint8_t)(0x50 + v)};
    std::vector<uint8_t> b;
    for (uint32_t a = v; a; a >>= 8) b.push_back((uint8_t)a);
    if (b.back() & 0x80) b.push_back(0);
    b.insert(b.begin(), (uint8_t)b.size());
    return b;
}
static void test_taptree(const Context &ctx) {
    const size_t heights[3] = {32, 8, 32}, widths[3] = {2, 3, 1};
    std::vector<std::vector<uint32_t>> host;
    std::vector<DeviceMatrix> mats;
    const uint32_t *ptrs[3];
    for (int i = 0; i < 3; i++) {
        host.push_back(rand_canonical(heights[i] * widths[i]));
        if (i == 0) host[0][0] = 0, host[0][1] = 16, host[0][2] = 128, host[0][3] = 0x77ffffff;  // every push length
        mats.emplace_back(ctx, RowMajorMatrix<Val>(monty(host[i]), widths[i]));
    }
    for (int i = 0; i < 3; i++) ptrs[i] = host[i].data();
    DeviceMatrix rows = TapTree::padded_rows(ctx, {&mats[0], &mats[1], &mats[2]});
    CHECK(rows.height() == 32 && rows.width() == 6);
    const std::vector<uint32_t> got_rows = canon(rows.to_row_major_matrix().values);
    for (size_t leaf = 0; leaf < 32; leaf++) {
        uint32_t want[6];
        CHECK(or_padded_leaf(ptrs, heights, widths, 3, leaf, want) == 6);
        CHECK(std::memcmp(want, &got_rows[6 * leaf], 24) == 0);
    }
    // a template shaped like generate_script (tcs/mod.rs:197-225): opaque "locking script" bytes around the pushes, > 252 bytes in
    // total so the compact-size length takes three bytes
    ScriptTemplate tpl;
    for (size_t k = 0; k <= 7; k++) {
        std::vector<uint8_t> seg(k == 0 ? 300 : 41 + k);
        for (auto &b : seg) b = (uint8_t)next_u64();
        tpl.segments.push_back(seg);
    }
    tpl.push_word = {0, 1, 5, 4, 3, 2};
    TapTree tree(ctx, rows, tpl);
    const auto perm = tree.leaf_indices();
    std::vector<bool> seen(32, false);
    for (uint32_t p : perm) {
        CHECK(p < 32 && !seen[p]);
        seen[p] = true;
    }
    for (size_t index : {0u, 1u, 13u, 31u}) {
        std::vector<uint8_t> script(tpl.segments[0]);
        auto num = push_num((uint32_t)index);
        script.insert(script.end(), num.begin(), num.end());
        for (size_t k = 0; k < tpl.push_word.size(); k++) {
            script.insert(script.end(), tpl.segments[k + 1].begin(), tpl.segments[k + 1].end());
            num = push_num(got_rows[6 * index + tpl.push_word[k]]);
            script.insert(script.end(), num.begin(), num.end());
        }
        script.insert(script.end(), tpl.segments.back().begin(), tpl.segments.back().end());
        uint8_t h[32], nx[32];
        or_tap_leaf_hash(script.data(), script.size(), h);
        auto [branch, pos] = tree.open(index);
        CHECK(branch.size() == 5 && pos == perm[index]);
        for (auto &sib : branch) {
            or_tap_branch_hash(h, sib.data(), nx);
            std::memcpy(h, nx, 32);
        }
        CHECK(std::memcmp(h, tree.root().data(), 32) == 0);
    }
}

int main() {
    Context ctx(0);
    run("dft_roundtrip_and_oracle", [&] { test_dft_roundtrip_and_oracle(ctx); });
    run("coset_lde_batch", [&] { test_coset_lde_batch(ctx); });
    run("fold_even_odd", [&] { test_fold_even_odd(ctx); });
    run("mmcs_p3_inject", [&] { test_mmcs(ctx, TS_LAYOUT_P3_INJECT); });
    run("mmcs_padded", [&] { test_mmcs(ctx, TS_LAYOUT_PADDED); });
    run("pcs_commit_single", [&] { do_test_pcs_commit(ctx, {{5, 10}}, 1); });
    run("pcs_commit_many_equal", [&] { do_test_pcs_commit(ctx, {{5, 10}, {5, 3}, {5, 7}}, 2); });
    run("pcs_commit_many_different", [&] { do_test_pcs_commit(ctx, {{3, 4}, {6, 9}, {4, 1}, {6, 2}}, 1); });
    run("commit_phase", [&] { test_commit_phase(ctx); });
    run("pcs_open", [&] { test_pcs_open(ctx); });
    run("taptree", [&] { test_taptree(ctx); });
    std::printf("%d passed, %d failed\n", passed, failures);
    return failures ? 1 : 0;
}
