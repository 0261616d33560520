"""Parity on a real B200: the shipped CUDA library, through the C ABI, bit-exact against the CPU oracle on
seeded inputs at oracle-friendly sizes, against the committed golden vectors, and - at BASELINE.json's
sizes - through size-independent properties (column-subset LDE, open->verify, transcript replay)."""
import numpy as np
import pytest

import parity_cases as pc

pytestmark = pytest.mark.gpu
P = pc.P


@pytest.fixture(scope="module")
def ts():
    from __graft_entry__ import build_device, load_pkg

    build_device()
    pkg = load_pkg()
    pkg.load_library()  # device build only; raises if missing
    assert pkg.lib().ts_is_device_build() == 1
    return pkg


@pytest.fixture(scope="module")
def ctx(ts):
    c = ts.Context(0)
    yield c
    c.close()


# ---------------------------------------------------------------- LDE
@pytest.mark.parametrize("log_n,width,b", [(0, 3, 2), (1, 1, 1), (2, 2, 2), (3, 5, 1), (4, 8, 2), (5, 3, 3), (7, 2, 2),
                                           (9, 9, 1), (10, 2, 2), (11, 8, 2), (11, 64, 1)])
def test_lde_single_digit(ts, ctx, orc, log_n, width, b):
    pc.check_lde(ts, ctx, orc, log_n, width, b)


@pytest.mark.parametrize("log_n,width,b", [(12, 3, 1), (13, 8, 2), (14, 1, 2), (12, 17, 2), (16, 64, 2), (18, 16, 2),
                                           (20, 4, 2), (22, 2, 1), (21, 3, 2)])
def test_lde_two_digits(ts, ctx, orc, log_n, width, b):
    pc.check_lde(ts, ctx, orc, log_n, width, b)


@pytest.mark.parametrize("log_n,width,b", [(23, 1, 1), (24, 2, 1), (23, 4, 2)])
def test_lde_three_digits(ts, ctx, orc, log_n, width, b):
    pc.check_lde(ts, ctx, orc, log_n, width, b)


def test_lde_other_shifts(ts, ctx, orc):
    pc.check_lde(ts, ctx, orc, 6, 4, 2, shift=1)
    pc.check_lde(ts, ctx, orc, 12, 2, 1, shift=pow(31, 5, P))
    pc.check_lde(ts, ctx, orc, 15, 5, 2, shift=31 * pow(7, P - 2, P) % P)


def test_lde_golden(ts, ctx, golden):
    dft = ts.GpuDft(ctx)
    for case in golden["lde"]:
        ev = np.array(case["evals"], dtype=np.uint32)
        got = dft.coset_lde_batch(ts.DeviceMatrix.from_canonical(ctx, ev), case["added_bits"], case["shift"],
                                  committed_order=True)
        assert got.to_canonical().tolist() == case["committed"]


def test_lde_natural_and_host(ts, ctx, orc):
    pc.check_lde_natural_and_host(ts, ctx, orc, 6, 3, 2)
    pc.check_lde_natural_and_host(ts, ctx, orc, 14, 5, 2)


@pytest.mark.parametrize("log_n,width", [(0, 2), (3, 3), (8, 2), (12, 2), (17, 3)])
def test_dft_family(ts, ctx, orc, log_n, width):
    pc.check_dft_family(ts, ctx, orc, log_n, width)


@pytest.mark.parametrize("log_n,width,b", [(18, 8, 2), (19, 24, 1), (20, 16, 2), (21, 12, 1), (22, 8, 1)])
def test_lde_tma_staging(ts, ctx, orc, monkeypatch, log_n, width, b):
    """TS_TMA=1: the contiguous-source passes stage their tile with cp.async.bulk.tensor (ntt_v4.cuh pass_tma_kernel,
    128-byte hardware swizzle = the tile's own sigma): every LDE word equals the oracle's, for digits 9, 10 and 11 and a
    ragged last column slice."""
    monkeypatch.setenv("TS_TMA", "1")
    pc.check_lde(ts, ctx, orc, log_n, width, b)


@pytest.mark.parametrize("log_n,width,b", [(22, 8, 1), (21, 12, 2)])
def test_lde_shuffle_exchange(ts, ctx, orc, monkeypatch, log_n, width, b):
    """TS_SHFL=1: the exchange between the last two rounds of a D = 11 tile goes through __shfl_xor (ntt_v4.cuh
    r34_shfl_st) instead of shared memory; same LDE words."""
    monkeypatch.setenv("TS_SHFL", "1")
    pc.check_lde(ts, ctx, orc, log_n, width, b)


def test_lde_config2_shape_column_subset(ts, ctx, orc):
    """BASELINE config 2 (2^20 x 64, log_blowup 2): the LDE is column-independent, so the oracle LDE of a column
    subset must equal the same columns of the full device result; plus the low-coset property
    (fri/src/two_adic_pcs.rs:247-258)."""
    log_n, w, b = 20, 64, 2
    ev = orc.splitmix_matrix(0, 1 << log_n, w)
    dft = ts.GpuDft(ctx)
    got = dft.coset_lde_batch(ts.DeviceMatrix.from_canonical(ctx, ev), b, 31, committed_order=True).to_canonical()
    cols = [0, 17, 63]
    want = orc.pcs_lde_committed(np.ascontiguousarray(ev[:, cols]), b)
    assert np.array_equal(got[:, cols], want)
    one = dft.coset_lde_batch(ts.DeviceMatrix.from_canonical(ctx, ev), b, 1, committed_order=True)
    low = one.to_canonical(0, 1 << log_n)
    assert np.array_equal(orc.bit_reverse_rows(low), ev)  # shift 1: the low coset is the input itself


# ---------------------------------------------------------------- MMCS
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("shapes", [[(1, 5)], [(2, 1)], [(64, 8)], [(256, 16)], [(128, 40)], [(8, 300)], [(4, 600)],
                                    [(2, 5000)], [(32, 3), (32, 7)], [(8, 3), (32, 2), (8, 5), (32, 1), (2, 4)],
                                    [(16, 8), (4, 8)], [(1 << 14, 256)], [(1 << 12, 200), (1 << 12, 16)],
                                    [(1 << 15, 8)], [(1 << 10, 64), (1 << 7, 9), (1 << 10, 1)]])
def test_mmcs(ts, ctx, orc, shapes, layout):
    pc.check_mmcs(ts, ctx, orc, shapes, layout, indices=(0, 1, 5, 31, 12345))


def test_mmcs_golden(ts, ctx, golden):
    mm = ts.Blake3MerkleMmcs(ctx)
    for case in golden["merkle_single"]:
        root, data = mm.commit([ts.DeviceMatrix.from_canonical(ctx, np.array(case["rows"], dtype=np.uint32))])
        assert root.hex() == case["root"]
        assert data.layer(0)[0].tobytes().hex() == case["leaf0"]


def test_blake3_reference_kats_on_device(ts, ctx, golden):
    """scripts/src/hashes/blake3.rs:537-587: 16 (resp. 15) LE u32 ones -> the reference's digests; here as the
    leaf hash of a one-row matrix."""
    mm = ts.Blake3MerkleMmcs(ctx)
    for k, w in zip(golden["blake3"]["kats"], (16, 15)):
        root, _ = mm.commit([ts.DeviceMatrix.from_canonical(ctx, np.ones((1, w), dtype=np.uint32))])
        assert root.hex() == k["hash"], k["src"]


def test_mmcs_large_open_verify(ts, ctx, orc):
    """2^20 leaves x 64 columns: root against the oracle, random openings verify on the host."""
    m = orc.splitmix_matrix(3, 1 << 20, 64)
    mm = ts.Blake3MerkleMmcs(ctx)
    root, data = mm.commit([ts.DeviceMatrix.from_canonical(ctx, m)])
    assert root == orc.mmcs_commit([m]).root
    rng = np.random.default_rng(0)
    for idx in rng.integers(0, 1 << 20, 8):
        rows, path = mm.open_batch(int(idx), data)
        assert np.array_equal(rows[0], m[idx])
        mm.verify_batch([1 << 20], rows, int(idx), path, root)


# ---------------------------------------------------------------- fold / challenger / commit phase
@pytest.mark.parametrize("log_h", [0, 1, 3, 7, 8, 9, 11, 16, 20])
def test_fold_ext(ts, ctx, orc, log_h):
    pc.check_fold_ext(ts, ctx, orc, log_h)


def test_fold_golden(ts, ctx, golden):
    for case in golden["fold_ef"]:
        got = ts.fold_even_odd(ctx, np.array(case["vals"], dtype=np.uint32), case["beta"])
        assert got.tolist() == case["out"]


def test_fold_base_reference_property(ts, ctx, orc):
    pc.check_fold_base_reference_property(ts, ctx, orc, log_n=10)  # fri/src/fold_even_odd.rs:65-95
    pc.check_fold_base_reference_property(ts, ctx, orc, log_n=16)


def test_challenger(ts, orc, golden):
    pc.check_challenger(ts, golden)
    pc.check_challenger_grind(ts, orc)


def test_commit_phase(ts, ctx, orc):
    pc.check_commit_phase(ts, ctx, orc, [6], 2)
    pc.check_commit_phase(ts, ctx, orc, [7, 5, 4], 1)
    pc.check_commit_phase(ts, ctx, orc, [16, 12], 2)
    pc.check_commit_phase(ts, ctx, orc, [10], 4)
    pc.check_commit_phase_rejects_high_degree(ts, ctx)


@pytest.mark.parametrize("rows,bw,nb,windows", [(64, 8, 4, [(0, 2), (2, 4)]), (256, 16, 8, [(0, 1), (1, 5), (5, 8)]), (32, 4, 8, [(0, 4), (4, 8)])])
def test_mmcs_incremental_commit(ts, ctx, orc, rows, bw, nb, windows):
    """ts_mmcs_commit_begin / _window / _finish: the row hash absorbed block window by block window (the row-sharded
    prover hashes a chunk's columns while the next chunk is still in flight) gives the root of the one-shot commit."""
    import ctypes as C

    m = pc.rand_mat(77, rows, bw * nb)
    blocks = [ts.DeviceMatrix.from_canonical(ctx, np.ascontiguousarray(m[:, i * bw : (i + 1) * bw])) for i in range(nb)]
    want = orc.mmcs_commit([m]).root
    root, data = ts.Blake3MerkleMmcs(ctx).commit(blocks)
    assert root == want
    arr = (C.c_void_p * nb)(*[b._h for b in blocks])
    th = C.c_void_p()
    ctx.check(ctx._L.ts_mmcs_commit_begin(ctx._h, arr, nb, C.byref(th)), "commit_begin")
    for b0, b1 in windows:
        ctx.check(ctx._L.ts_mmcs_commit_window(ctx._h, th, b0, b1), "commit_window")
    got = (C.c_uint8 * 32)()
    ctx.check(ctx._L.ts_mmcs_commit_finish(ctx._h, th, got), "commit_finish")
    assert bytes(got) == want
    assert ctx._L.ts_mmcs_commit_window(ctx._h, th, 1, 1) != 0  # empty window
    ctx._L.ts_tree_free(th)


@pytest.mark.parametrize("env", [{"TS_NO_FRI_TAIL": "1"}, {"TS_NO_FRI_CHAIN": "1"}, {"TS_NO_FRI_CHAIN": "1", "TS_NO_FRI_TAIL": "1"}, {},
                                 {"TS_NO_FRI_TAIL": "1", "TS_NO_FOLD_HASH": "1"}, {"TS_TREE3_MIN_LOG": "14"}])
def test_commit_phase_round_forms(ts, ctx, orc, monkeypatch, env):
    """The three forms of a commit-phase round produce one transcript: chained on the device (sponge_step_kernel, beta read
    by the fold kernel from device memory), the single-CTA tail, and the round-1 form with a host sponge per round."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    pc.check_commit_phase(ts, ctx, orc, [12], 2, seed=31)
    pc.check_commit_phase(ts, ctx, orc, [8, 6, 3], 1, seed=32)
    pc.check_commit_phase(ts, ctx, orc, [11, 10, 9], 1, seed=33)  # later inputs added inside fold_hash_kernel (prover.rs:124-126)
    pc.check_commit_phase(ts, ctx, orc, [16], 2, seed=34)         # fused fold + leaf hash above the tail, all tree kernels


@pytest.mark.parametrize("log_h,shards,with_addend", [(9, 1, False), (10, 2, True), (12, 4, False), (20, 8, True), (22, 2, False)])
def test_fold_hash_shard(ts, ctx, orc, log_h, shards, with_addend):
    pc.check_fold_hash_shard(ts, ctx, orc, log_h, shards, with_addend)


def test_commit_phase_golden(ts, ctx, orc, golden):
    g = golden["commit_phase"]
    cw = orc.pcs_lde_committed(np.array(g["evals"], dtype=np.uint32), g["log_blowup"])
    cfg = ts.FriConfig(g["log_blowup"], 4, 8, ts.Blake3MerkleMmcs(ctx))
    res = ts.bf_commit_phase(cfg, [ts.DeviceMatrix.from_canonical(ctx, cw)], ts.BfChallenger())
    assert [c.hex() for c in res.commits] == g["commits"]
    assert res.final_poly.tolist() == g["final_poly"]


def test_commit_phase_2_20(ts, ctx, orc):
    """FRI sweep point (config 5): codeword 2^20 of BabyBear^4, blowup 4 -> 18 rounds, bit-exact transcript."""
    pc.check_commit_phase(ts, ctx, orc, [18], 2, seed=40)


# ---------------------------------------------------------------- PCS
def test_pcs_commit(ts, ctx, orc):
    pc.check_pcs_commit(ts, ctx, orc, [(5, 3)], 2)
    pc.check_pcs_commit(ts, ctx, orc, [(6, 2), (4, 5), (6, 1)], 1)
    pc.check_pcs_commit(ts, ctx, orc, [(6, 2), (4, 5)], 1, layout=1)
    pc.check_pcs_commit(ts, ctx, orc, [(10, 2)], 2)  # config 1 shape: Fibonacci trace 2^10 x 2, log_blowup 2
    pc.check_pcs_commit(ts, ctx, orc, [(14, 40), (14, 4)], 2)


@pytest.mark.parametrize("width", [192, 160, 256, 320])
def test_pcs_commit_host_pipelined(ts, ctx, orc, width):
    """ts_pcs_commit_host on a trace wide enough for the H2D/LDE pipeline (64-column chunks, last one partial).
    Rows of at most one Blake3 chunk (<= 256 columns) are also hashed chunk by chunk behind the copy."""
    ev = orc.splitmix_matrix(9, 1 << 18, width)
    mm = ts.Blake3MerkleMmcs(ctx)
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mm, ts.FriConfig(1, 2, 8, mm))
    dom = pcs.natural_domain_for_degree(1 << 18)
    ev_m = ts.to_monty(ev)
    root_h, data_h = pcs.commit_host([(dom, ev_m)])  # pageable source: gathered through the page-locked bounce slots
    root_d, _ = pcs.commit([(dom, ts.DeviceMatrix.from_canonical(ctx, ev))])
    assert root_h == root_d
    L = ts.lib()
    ctx.check(L.ts_host_register(ctx._h, ev_m.ctypes.data, ev_m.nbytes), "host_register")
    try:
        root_p, _ = pcs.commit_host([(dom, ev_m)])  # page-locked source: strided 2-D copies straight from the caller's rows
    finally:
        ctx.check(L.ts_host_unregister(ctx._h, ev_m.ctypes.data), "host_unregister")
    assert root_p == root_d
    cols = [0, 63, 64, width - 1]
    lde = mm.get_matrices(data_h)[0].to_canonical()
    assert np.array_equal(lde[:, cols], orc.pcs_lde_committed(np.ascontiguousarray(ev[:, cols]), 1))
    assert root_h == orc.mmcs_commit([orc.pcs_lde_committed(ev, 1)]).root


def test_dot_ext_powers(ts, ctx, orc):
    pc.check_dot_ext_powers(ts, ctx, orc, 100, 70)
    pc.check_dot_ext_powers(ts, ctx, orc, 64, 3)
    pc.check_dot_ext_powers(ts, ctx, orc, 1000, 4)   # one thread per row (quotient chunks)
    pc.check_dot_ext_powers(ts, ctx, orc, 77, 8)
    pc.check_dot_ext_powers(ts, ctx, orc, 1 << 12, 256)


def test_full_pipeline_config1(ts, ctx, orc):
    """Config 1 shape end to end on the device path: trace 2^10 x 2 -> commit -> alpha-reduction -> FRI commit
    phase (10 rounds to 4 equal values), every transcript value equal to the oracle's."""
    log_n, w, b = 10, 2, 2
    trace = orc.splitmix_matrix(1, 1 << log_n, w)
    mm = ts.Blake3MerkleMmcs(ctx)
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mm, ts.FriConfig(b, 16, 8, mm))
    ch, rch = ts.BfChallenger(), orc.BfChallenger()
    root, data = pcs.commit([(pcs.natural_domain_for_degree(1 << log_n), ts.DeviceMatrix.from_canonical(ctx, trace))])
    lde_ref = orc.pcs_lde_committed(trace, b)
    assert root == orc.mmcs_commit([lde_ref]).root
    ch.observe(root)
    rch.observe_digest(root)
    alpha = ch.sample()
    assert np.array_equal(alpha, rch.sample_ef())
    fri_in = pcs.dot_ext_powers(mm.get_matrices(data)[0], alpha)
    ref_in = fri_in.to_canonical()
    res = ts.bf_commit_phase(pcs.fri, [fri_in], ch)
    ref = orc.fri_commit_phase([ref_in], b, rch)
    assert len(res.commits) == 10 and res.commits == ref["commits"]
    assert np.array_equal(res.final_poly, ref["final_poly"])
    w1, w2 = ch.grind(8), rch.grind(8)
    assert w1 == w2
    assert [ch.sample_bits(12) for _ in range(16)] == [rch.sample_bits(12) for _ in range(16)]


# ---------------------------------------------------------------- BASELINE.json configs 3-5 (shapes named there)
def test_config4_shape_reduced_height(ts, ctx, orc):
    """Config 4 shape at an oracle-friendly height: ~200-column trace + degree-4 extension quotient as 4 chunks
    of 4 base columns (uni-stark/src/prover.rs:78-83: chunk i has domain shift g * w^i), one commit each."""
    log_n, b = 14, 2
    mm = ts.Blake3MerkleMmcs(ctx)
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mm, ts.FriConfig(b, 28, 8, mm))
    trace = orc.splitmix_matrix(4, 1 << log_n, 200)
    root, data = pcs.commit([(pcs.natural_domain_for_degree(1 << log_n), ts.DeviceMatrix.from_canonical(ctx, trace))])
    assert root == orc.mmcs_commit([orc.pcs_lde_committed(trace, b)]).root
    # quotient chunks: domain shift 31 * w_{4n}^i, i = 0..3  ->  LDE shift 31 / (31 w^i) = w^-i
    w4n = orc.two_adic_generator(log_n + 2)
    chunks = [orc.splitmix_matrix(10 + i, 1 << log_n, 4) for i in range(4)]
    doms = [ts.TwoAdicMultiplicativeCoset(log_n, 31 * pow(w4n, i, P) % P) for i in range(4)]
    rootq, dataq = pcs.commit([(d, ts.DeviceMatrix.from_canonical(ctx, c)) for d, c in zip(doms, chunks)])
    ldes = [orc.pcs_lde_committed(c, b, pow(pow(w4n, i, P), P - 2, P)) for i, c in enumerate(chunks)]
    assert rootq == orc.mmcs_commit(ldes).root
    rows, path = mm.open_batch(12345, dataq)
    mm.verify_batch([1 << (log_n + b)] * 4, rows, 12345, path, rootq)


def test_config4_full_height_properties(ts, ctx, orc):
    """2^21 x 200, log_blowup 2 (RISC0-recursion scale): column-subset LDE parity, openings verify on the host."""
    log_n, w, b = 21, 200, 2
    import torch

    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    t = torch.randint(0, P, (1 << log_n, w), dtype=torch.int32, device="cuda", generator=g)
    torch.cuda.synchronize()
    mm = ts.Blake3MerkleMmcs(ctx)
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mm, ts.FriConfig(b, 28, 8, mm))
    ev = ts.DeviceMatrix.wrap_device(ctx, t.data_ptr(), 1 << log_n, w, keepalive=t)
    root, data = pcs.commit([(pcs.natural_domain_for_degree(1 << log_n), ev)])
    lde = mm.get_matrices(data)[0]
    cols = [0, 7, 8, 199]
    host_cols = ts.from_monty(t[:, cols].cpu().numpy().view(np.uint32))
    want = orc.pcs_lde_committed(np.ascontiguousarray(host_cols), b)
    for r0 in (0, (1 << 23) - 4096, 3 << 21):
        got = lde.to_canonical(r0, 4096)
        assert np.array_equal(got[:, cols], want[r0 : r0 + 4096])
    for idx in (0, 1, (1 << 23) - 1, 4242424):
        rows, path = mm.open_batch(idx, data)
        assert np.array_equal(rows[0][cols], want[idx])
        mm.verify_batch([1 << 23], rows, idx, path, root)


def _replay_query_path(ts, orc, cfg, res, log_len, idx, rch=None):
    """Host replay of one FRI query path (fold_row, two_adic_pcs.rs:87-114) over the opened layer rows: betas from the
    commitments, every opening verified against its layer commitment, the last folded value = final_poly.  rch: an
    oracle challenger in the state the commit phase started from."""
    rch = rch if rch is not None else orc.BfChallenger()
    folded = None
    for r, (commit, pd) in enumerate(zip(res.commits, res.data)):
        rch.observe_digest(commit)
        beta = rch.sample_ef()
        pair = idx >> 1
        rows, path = cfg.mmcs.open_batch(pair, pd)
        cfg.mmcs.verify_batch([1 << (log_len - 1 - r)], rows, pair, path, commit)
        e0, e1 = rows[0][:4], rows[0][4:]
        if folded is not None:
            assert np.array_equal(folded, e1 if idx & 1 else e0)
        folded = orc.fold_row_ef(pair, log_len - 1 - r, beta, e0, e1)
        idx = pair
    assert np.array_equal(folded, res.final_poly)


def test_config3_full_size(ts, ctx, orc):
    """BASELINE config 3 = the bench workload at its own size (2^22 x 256 SplitMix trace, log_blowup 2), one GPU:
      * the device trace generator against the oracle's definition (column subset);
      * LDE (two_adic_pcs.rs:237-240): a column subset against the oracle LDE at three row windows;
      * leaf digests of sampled rows = Blake3 of the row's canonical LE bytes (orc.blake3), >= 8 openings verified by the
        ORACLE's verify_batch against the device root;
      * root = the oracle's 2-to-1 tree over the device's 2^24-leaf layer;
      * alpha-reduction on sampled rows against big-int arithmetic, then the FRI commit phase (prover.rs:93-141) replayed
        along two query paths with the oracle's fold_row, ending in the device's final polynomial."""
    log_n, w, b, seed = 22, 256, 2, 0
    n, N = 1 << log_n, 1 << (log_n + b)
    mm = ts.Blake3MerkleMmcs(ctx)
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mm, ts.FriConfig(b, 16, 8, mm))
    ev = ts.DeviceMatrix.splitmix(ctx, seed, n, w)
    cols = [0, 9, 130, 255]
    host_cols = orc.splitmix_columns(seed, n, w, cols)
    for r0 in (0, n - 1024, 1234567):
        assert np.array_equal(ev.to_canonical(r0, 1024)[:, cols], host_cols[r0 : r0 + 1024])
    root, data = pcs.commit([(pcs.natural_domain_for_degree(n), ev)])
    lde = mm.get_matrices(data)[0]
    want = orc.pcs_lde_committed(host_cols, b)
    for r0 in (0, N - 4096, 5 << 21):
        assert np.array_equal(lde.to_canonical(r0, 4096)[:, cols], want[r0 : r0 + 4096])
    leaves = data.layer(0)
    assert leaves.shape == (N, 32)
    assert orc.merkle_root_from_leaves(leaves) == root
    rng = np.random.default_rng(3)
    sample = [0, 1, N - 1] + [int(x) for x in rng.integers(0, N, 7)]
    for idx in sample:
        rows, path = mm.open_batch(idx, data)
        assert np.array_equal(rows[0][cols], want[idx])
        assert orc.blake3(rows[0].astype("<u4").tobytes()) == leaves[idx].tobytes()
        assert orc.mmcs_verify_batch([N], rows, idx, path, root)
    # alpha-reduction + FRI commit phase, as bench.py runs them
    ch = ts.BfChallenger()
    ch.observe(root)
    alpha = ch.sample()
    och = orc.BfChallenger()
    och.observe_digest(root)
    assert np.array_equal(np.asarray(och.sample_ef(), dtype=np.uint32), np.asarray(alpha, dtype=np.uint32))
    fri_in = pcs.dot_ext_powers(lde, alpha)
    for idx in sample[:4]:
        row = lde.to_canonical(idx, 1)[0]
        acc, apow = np.zeros(4, dtype=np.uint32), np.array([1, 0, 0, 0], dtype=np.uint32)
        for c in range(w):
            acc = (acc.astype(np.uint64) + apow.astype(np.uint64) * int(row[c])) % P
            apow = orc.ef_mul(apow, np.asarray(alpha, dtype=np.uint32))
        assert np.array_equal(fri_in.to_canonical(idx, 1)[0], acc.astype(np.uint32))
    res = ts.bf_commit_phase(pcs.fri, [fri_in], ch)
    assert len(res.commits) == log_n
    for idx in (987654321 % N, 3):
        _replay_query_path(ts, orc, pcs.fri, res, log_n + b, idx, rch=och.clone())


@pytest.mark.parametrize("log_len,log_blowup", [(18, 1), (20, 3), (22, 4), (24, 2)])
def test_config5_fri_sweep(ts, ctx, orc, log_len, log_blowup):
    """FRI commit-phase sweep over BabyBear^4 codewords (config 5): the codeword is the LDE of a random
    low-degree extension polynomial (made on the device), folded down to the final polynomial.  Up to 2^20 the
    whole transcript is compared with the oracle; above, the library's own constancy check (prover.rs:130-134)
    and a host replay of fold_row along one query path (two_adic_pcs.rs:87-114) must agree."""
    import torch

    log_n = log_len - log_blowup
    g = torch.Generator(device="cuda")
    g.manual_seed(log_len)
    coeff_evals = torch.randint(0, P, (1 << log_n, 4), dtype=torch.int32, device="cuda", generator=g)
    torch.cuda.synchronize()
    ev = ts.DeviceMatrix.wrap_device(ctx, coeff_evals.data_ptr(), 1 << log_n, 4, keepalive=coeff_evals)
    cw = ts.GpuDft(ctx).coset_lde_batch(ev, log_blowup, 31, committed_order=True)
    cfg = ts.FriConfig(log_blowup, 4, 8, ts.Blake3MerkleMmcs(ctx))
    ch = ts.BfChallenger()
    res = ts.bf_commit_phase(cfg, [cw], ch)
    assert len(res.commits) == log_n
    if log_len <= 20:
        ref = orc.fri_commit_phase([cw.to_canonical()], log_blowup, orc.BfChallenger())
        assert ref["ok"] and res.commits == ref["commits"] and np.array_equal(res.final_poly, ref["final_poly"])
        return
    # replay: betas from the commitments, then fold_row along the path of one index using opened layer rows
    _replay_query_path(ts, orc, cfg, res, log_len, 987654321 % (1 << log_len))


@pytest.mark.parametrize("log_n,width,ctas", [(3, 1, None), (9, 2, None), (10, 4, 3), (9, 60, 2), (9, 200, 3), (8, 300, 1), (7, 1100, 1), (7, 301, 2), (11, 8, 3),
                                              (18, 4, None), (17, 200, None)])
def test_interpolate_low_coset(ts, ctx, orc, monkeypatch, log_n, width, ctas):
    """interpolate_coset on the device against sum_k c_k z^k; the last two shapes have more row blocks than CTAs."""
    if ctas is not None:
        monkeypatch.setenv("TS_BARY_CTAS", str(ctas))
    pc.check_interpolate_low_coset(ts, ctx, orc, log_n, width, 1, seed=60 + width)


def test_pcs_open_verify(ts, ctx, orc):
    """TwoAdicFriPcs::open end to end on the device, accepted by the restated reference verifier (parity rung L5)."""
    pc.check_pcs_open_verify(ts, ctx, orc, [[(5, 3, 2)], [(5, 4, 1), (5, 4, 1)]], 2)
    pc.check_pcs_open_verify(ts, ctx, orc, [[(4, 2, 1), (6, 3, 1), (3, 2, 1)]], 1, seed=90)
    pc.check_pcs_open_verify(ts, ctx, orc, [[(8, 12, 2)], [(8, 4, 1)]], 2, num_queries=16, pow_bits=8, seed=110)


def test_config4_open_at_scale(ts, ctx, orc):
    """BASELINE config 4 through Pcs::open in ONE C-ABI call (ts_pcs_open): a 2^21 x 200 trace opened at zeta and
    zeta*g plus four 2^21 x 4 quotient chunks (their own cosets, uni-stark/src/prover.rs:78-83) opened at zeta, 28 queries,
    8 proof-of-work bits.  The bytes decode to a proof the restated reference verifier (two_adic_pcs.rs:421-530,
    fri/src/verifier.rs) accepts; a flipped opened value is rejected.  A wrong opened value anywhere would also make the
    reduced opening high-degree, i.e. the commit phase's final layer non-constant (prover.rs:130-134)."""
    import importlib

    from oracle import verifier as V

    log_n, w, b, nq, pow_bits = 21, 200, 2, 28, 8
    n = 1 << log_n
    mm = ts.Blake3MerkleMmcs(ctx)
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mm, ts.FriConfig(b, nq, pow_bits, mm))
    ch, och = ts.BfChallenger(), orc.BfChallenger()
    trace = ts.DeviceMatrix.splitmix(ctx, 4, n, w)
    root_t, data_t = pcs.commit([(pcs.natural_domain_for_degree(n), trace)])
    ch.observe(root_t)
    och.observe_digest(root_t)
    gen22 = orc.two_adic_generator(log_n + 2)
    chunks = [ts.DeviceMatrix.splitmix(ctx, 10 + k, n, 4) for k in range(4)]
    doms = [ts.TwoAdicMultiplicativeCoset(log_n, 31 * pow(gen22, k, P) % P) for k in range(4)]  # split_domains
    root_q, data_q = pcs.commit(list(zip(doms, chunks)))
    ch.observe(root_q)
    och.observe_digest(root_q)
    zeta = [int(x) for x in ch.sample()]
    assert zeta == [int(x) for x in och.sample_ef()]
    g = orc.two_adic_generator(log_n)
    zeta_next = [c * g % P for c in zeta]
    rounds = [(data_t, [[zeta, zeta_next]]), (data_q, [[zeta]] * 4)]
    och_v = och.clone()
    blob = pcs.open_bytes(rounds, ch)
    opened, proof = importlib.import_module("tapstark_b200.proofio").decode_opening(blob)
    assert len(proof.query_proofs) == nq and len(proof.commit_phase_commits) == log_n
    assert opened[0][0][0].shape == (w, 4) and opened[1][3][0].shape == (4, 4)
    v_rounds = [(root_t, [(log_n, [(zeta, opened[0][0][0].tolist()), (zeta_next, opened[0][0][1].tolist())])]),
                (root_q, [(log_n, [(zeta, opened[1][k][0].tolist())]) for k in range(4)])]
    assert V.pcs_verify(b, nq, pow_bits, v_rounds, proof, och_v.clone())
    v_rounds[0][1][0][1][0][1][7][2] = (v_rounds[0][1][0][1][0][1][7][2] + 1) % P
    with pytest.raises(V.VerifyError):
        V.pcs_verify(b, nq, pow_bits, v_rounds, proof, och_v.clone())


def test_stark_fibonacci_config1(ts, ctx, orc):
    """BASELINE config 1: Fibonacci AIR, 2^10 x 2 trace, log_blowup 2 -- uni_stark::prove on the device path
    (trace commit -> quotient values kernel -> quotient commit -> open at zeta, zeta*g -> FRI), verified by the
    restated uni_stark::verify; quotient chunks bit-exact against the oracle's folder evaluation."""
    import airs

    trace = airs.fibonacci_trace(0, 1, 1 << 10)
    pc.check_stark_prove_verify(ts, ctx, orc, airs.FibonacciAir(), trace, [0, 1, int(trace[-1, 1])], 2, num_queries=8, pow_bits=8)


@pytest.mark.parametrize("degree,log_n", [(3, 6), (5, 7)])
def test_stark_mul_air(ts, ctx, orc, degree, log_n):
    """Constraint degree 3 / 5: quotient degree 2 / 4, chunks committed on their own cosets (split_domains)."""
    import airs

    air = airs.MulAir(degree=degree, reps=3)
    pc.check_stark_prove_verify(ts, ctx, orc, air, airs.mul_trace(air, 1 << log_n, 11), [], 3 if degree == 5 else 2)


def test_stark_program_errors(ts, ctx):
    import ctypes as C

    m = ts.DeviceMatrix.from_canonical(ctx, np.zeros((32, 2), dtype=np.uint32))
    out = (C.c_void_p * 1)()
    alpha = np.zeros(4, dtype=np.uint32)
    one = np.zeros(1, dtype=np.uint32)
    for prog in ([[9, 0, 0, 0]], [[0, 64, 0, 0]], [[0, 0, (1 << 28) | 2, 0]], [[4, 0, (3 << 28) | 0, 0]]):
        p = np.array(prog, dtype=np.uint32)
        rc = ctx._L.ts_quotient_values(ctx._h, m._h, 3, 0, p.ctypes.data_as(C.c_void_p), 1, one.ctypes.data_as(C.c_void_p), 0,
                                       one.ctypes.data_as(C.c_void_p), 0, alpha.ctypes.data_as(C.c_void_p), out)
        assert rc == 2  # TS_ERR_ARG


def test_stark_quotient_golden(ts, ctx, golden):
    pc.check_stark_golden(ts, ctx, golden["stark_fibonacci"])


def test_stark_counter_air(ts, ctx, orc):
    """Constants, negation, assert_one, when(<expression>): the NEG / CONST operands of the constraint-program kernel."""
    import airs

    n = 1 << 8
    pc.check_stark_prove_verify(ts, ctx, orc, airs.CounterAir(), airs.counter_trace(n), [n - 1], 2)
