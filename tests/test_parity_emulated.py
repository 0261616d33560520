"""Kernel-source parity WITHOUT a GPU: tap-stark_b200/csrc compiled by g++ against the SIMT emulator in
tests/emul (test-only; see tests/emul/cuda_emul.h) and compared bit-exactly with the oracle.  These are the
same cases the GPU module runs on the real library, at emulator-friendly sizes."""
import numpy as np
import pytest

import parity_cases as pc


@pytest.fixture(scope="module")
def ts():
    from emul.build_emul import build

    from __graft_entry__ import load_pkg

    pkg = load_pkg()
    pkg.load_library(build(), allow_emulated=True)
    yield pkg
    pkg._lib = None


@pytest.fixture(scope="module")
def ctx(ts):
    c = ts.Context(0)
    yield c
    c.close()


def test_product_loader_rejects_emulated_library(ts):
    from emul.build_emul import OUT

    keep = ts._lib
    with pytest.raises(ts.TapStarkError):
        ts.load_library(OUT)  # allow_emulated defaults to False: the product path never computes on a CPU
    ts._lib = keep


@pytest.mark.parametrize("log_n,width,b", [(0, 3, 2), (1, 1, 1), (2, 2, 2), (3, 5, 1), (4, 8, 2), (5, 3, 3),
                                           (7, 2, 2), (9, 9, 1), (10, 2, 2), (11, 8, 2)])
def test_lde_single_digit(ts, ctx, orc, log_n, width, b):
    pc.check_lde(ts, ctx, orc, log_n, width, b)


@pytest.mark.parametrize("log_n,width,b", [(12, 3, 1), (13, 8, 2), (14, 1, 2), (12, 17, 2)])
def test_lde_two_digits(ts, ctx, orc, log_n, width, b):
    pc.check_lde(ts, ctx, orc, log_n, width, b)


@pytest.mark.parametrize("log_n,width,b,digits", [(18, 8, 1, None), (18, 12, 2, None), (19, 8, 1, None),
                                                  (20, 8, 1, "11,9"), (20, 8, 1, "9,11")])
def test_lde_fast_path(ts, ctx, orc, log_n, width, b, digits, monkeypatch):
    """ntt_fast.cuh kernels (digits 9..11, width % 4 == 0); TS_DIGITS forces the split so that D = 11 is
    reachable at emulator-friendly sizes."""
    if digits:
        monkeypatch.setenv("TS_DIGITS", digits)
    pc.check_lde(ts, ctx, orc, log_n, width, b)


@pytest.mark.parametrize("chunk,width", [(8, 20), (16, 40)])
def test_pcs_commit_host_pipelined(ts, ctx, orc, monkeypatch, chunk, width):
    """ts_pcs_commit_host cuts a wide trace into column chunks (H2D of chunk k+1 overlaps the LDE of chunk k on the
    GPU); each chunk is an LDE window into the full-width output.  TS_CHUNK_COLS shrinks the chunk for the emulator.
    Chunks of 8, 8 and 4 columns: rows hashed after the LDE; chunks of 16, 16 and 8: rows hashed incrementally, one
    64-byte Blake3 block per chunk, the chaining value parked in the leaf-digest array between launches."""
    import numpy as np

    monkeypatch.setenv("TS_CHUNK_COLS", str(chunk))
    ev = pc.rand_mat(21, 1 << 18, width)
    mm = ts.Blake3MerkleMmcs(ctx)
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mm, ts.FriConfig(1, 2, 8, mm))
    ctx.reset_stats()
    root, data = pcs.commit_host([(pcs.natural_domain_for_degree(1 << 18), ts.to_monty(ev))])
    assert ctx.stats()["hash_leaves"]["launches"] == (3 if chunk == 16 else 1)
    lde = orc.pcs_lde_committed(ev, 1)
    assert np.array_equal(mm.get_matrices(data)[0].to_canonical(), lde)
    assert root == orc.mmcs_commit([lde]).root


def test_pcs_commit_overlapped_hash(ts, ctx, orc, monkeypatch):
    """Device-resident commit of one wide matrix: LDE by column chunks with the Blake3 window of chunk k issued on a
    second stream beside the LDE of chunk k+1 (lde_hash_overlapped); same root and LDE as the oracle."""
    import numpy as np

    monkeypatch.setenv("TS_CHUNK_COLS", "16")
    monkeypatch.setenv("TS_OVERLAP_HASH", "1")  # opt-in: measured neutral on the B200 (profiles/r01/README.md)
    ev = pc.rand_mat(33, 1 << 18, 48)
    mm = ts.Blake3MerkleMmcs(ctx)
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mm, ts.FriConfig(1, 2, 8, mm))
    ctx.reset_stats()
    root, data = pcs.commit([(pcs.natural_domain_for_degree(1 << 18), ts.DeviceMatrix.from_canonical(ctx, ev))])
    assert ctx.stats()["hash_leaves"]["launches"] == 3
    lde = orc.pcs_lde_committed(ev, 1)
    assert np.array_equal(mm.get_matrices(data)[0].to_canonical(), lde)
    assert root == orc.mmcs_commit([lde]).root


def test_lde_other_shift(ts, ctx, orc):
    pc.check_lde(ts, ctx, orc, 6, 4, 2, shift=1)
    pc.check_lde(ts, ctx, orc, 12, 2, 1, shift=pow(31, 5, pc.P))


def test_lde_natural_and_host(ts, ctx, orc):
    pc.check_lde_natural_and_host(ts, ctx, orc, 6, 3, 2)


@pytest.mark.parametrize("log_n,width", [(0, 2), (3, 3), (8, 2), (12, 2)])
def test_dft_family(ts, ctx, orc, log_n, width):
    pc.check_dft_family(ts, ctx, orc, log_n, width)


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("shapes", [[(1, 5)], [(2, 1)], [(64, 8)], [(256, 16)], [(128, 40)], [(8, 300)],
                                    [(4, 600)], [(32, 3), (32, 7)], [(8, 3), (32, 2), (8, 5), (32, 1), (2, 4)],
                                    [(16, 8), (4, 8)]])
def test_mmcs(ts, ctx, orc, shapes, layout):
    pc.check_mmcs(ts, ctx, orc, shapes, layout, indices=(0, 1, 5, 31))


def test_mmcs_fri_leaf_width(ts, ctx, orc):
    pc.check_mmcs(ts, ctx, orc, [(512, 8)], 0, indices=(0, 300))


@pytest.mark.parametrize("log_h", [0, 1, 3, 7, 8, 9, 11])
def test_fold_ext(ts, ctx, orc, log_h):
    pc.check_fold_ext(ts, ctx, orc, log_h)


def test_fold_golden(ts, ctx, golden):
    import numpy as np

    for case in golden["fold_ef"]:
        got = ts.fold_even_odd(ctx, np.array(case["vals"], dtype=np.uint32), case["beta"])
        assert got.tolist() == case["out"]


def test_fold_base_reference_property(ts, ctx, orc):
    pc.check_fold_base_reference_property(ts, ctx, orc, log_n=8)


def test_challenger(ts, orc, golden):
    pc.check_challenger(ts, golden)
    pc.check_challenger_grind(ts, orc)


def test_commit_phase(ts, ctx, orc):
    pc.check_commit_phase(ts, ctx, orc, [6], 2)
    pc.check_commit_phase(ts, ctx, orc, [7, 5, 4], 1)
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
                                 {"TS_NO_FRI_TAIL": "1", "TS_NO_FOLD_HASH": "1"}])
def test_commit_phase_round_forms(ts, ctx, orc, monkeypatch, env):
    """The three forms of a commit-phase round produce one transcript: chained on the device (sponge_step_kernel, beta read
    by the fold kernel from device memory), the single-CTA tail, and the round-1 form with a host sponge per round."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    pc.check_commit_phase(ts, ctx, orc, [12], 2, seed=31)
    pc.check_commit_phase(ts, ctx, orc, [8, 6, 3], 1, seed=32)
    pc.check_commit_phase(ts, ctx, orc, [11, 10, 9], 1, seed=33)  # later inputs added inside fold_hash_kernel (prover.rs:124-126)


@pytest.mark.parametrize("log_h,shards,with_addend", [(9, 1, False), (10, 2, True), (12, 4, False)])
def test_fold_hash_shard(ts, ctx, orc, log_h, shards, with_addend):
    pc.check_fold_hash_shard(ts, ctx, orc, log_h, shards, with_addend)


def test_commit_phase_golden(ts, ctx, orc, golden):
    import numpy as np

    g = golden["commit_phase"]
    cw = orc.pcs_lde_committed(np.array(g["evals"], dtype=np.uint32), g["log_blowup"])
    cfg = ts.FriConfig(g["log_blowup"], 4, 8, ts.Blake3MerkleMmcs(ctx))
    res = ts.bf_commit_phase(cfg, [ts.DeviceMatrix.from_canonical(ctx, cw)], ts.BfChallenger())
    assert [c.hex() for c in res.commits] == g["commits"]
    assert res.final_poly.tolist() == g["final_poly"]


def test_pcs_commit(ts, ctx, orc):
    pc.check_pcs_commit(ts, ctx, orc, [(5, 3)], 2)
    pc.check_pcs_commit(ts, ctx, orc, [(6, 2), (4, 5), (6, 1)], 1)
    pc.check_pcs_commit(ts, ctx, orc, [(6, 2), (4, 5)], 1, layout=1)


def test_dot_ext_powers(ts, ctx, orc):
    pc.check_dot_ext_powers(ts, ctx, orc, 100, 70)
    pc.check_dot_ext_powers(ts, ctx, orc, 64, 3)
    pc.check_dot_ext_powers(ts, ctx, orc, 1000, 4)   # one thread per row (quotient chunks)
    pc.check_dot_ext_powers(ts, ctx, orc, 77, 8)
    pc.check_dot_ext_powers(ts, ctx, orc, 300, 72)   # width % 4 == 0: fast kernel, partial last block
    pc.check_dot_ext_powers(ts, ctx, orc, 33, 256)


@pytest.mark.parametrize("log_n,width,ctas", [(3, 1, None), (9, 2, None), (10, 4, 3), (9, 60, 2), (9, 200, 3), (8, 300, 1), (7, 1100, 1), (7, 301, 2), (11, 8, 3)])
def test_interpolate_low_coset(ts, ctx, orc, monkeypatch, log_n, width, ctas):
    """Row-lane mapping for narrow matrices, one lane for wide ones, column groups past 256, grid-stride row blocks."""
    if ctas is not None:
        monkeypatch.setenv("TS_BARY_CTAS", str(ctas))
    pc.check_interpolate_low_coset(ts, ctx, orc, log_n, width, 1, seed=60 + width)


def test_pcs_open_verify(ts, ctx, orc):
    # uni-stark shape (uni-stark/src/prover.rs:94-104): trace opened at zeta and zeta*w, quotient chunks at zeta
    pc.check_pcs_open_verify(ts, ctx, orc, [[(5, 3, 2)], [(5, 4, 1), (5, 4, 1)]], 2)
    # fri/tests/pcs.rs "many_different": mixed heights in one commit
    pc.check_pcs_open_verify(ts, ctx, orc, [[(4, 2, 1), (6, 3, 1), (3, 2, 1)]], 1, seed=90)


def test_stark_fibonacci_prove_verify(ts, ctx, orc):
    """Config 1's AIR (uni-stark/tests/fib_air.rs: 2^3 rows, log_blowup 2): device quotient values == oracle folder,
    proof accepted by the restated uni_stark::verify, wrong public value rejected."""
    import airs

    trace = airs.fibonacci_trace(0, 1, 1 << 3)
    pc.check_stark_prove_verify(ts, ctx, orc, airs.FibonacciAir(), trace, [0, 1, int(trace[-1, 1])], 2)


def test_stark_mul_air_quotient_chunks(ts, ctx, orc):
    """Degree-3 constraints: quotient degree 2, two chunks on the cosets g*H and g*w_2n*H (split_evals / split_domains)."""
    import airs

    air = airs.MulAir(degree=3, reps=2)
    pc.check_stark_prove_verify(ts, ctx, orc, air, airs.mul_trace(air, 1 << 4, 5), [], 2)


def test_stark_program_errors(ts, ctx):
    """A malformed constraint program is an argument error at the ABI, not an out-of-bounds access."""
    import ctypes as C
    import numpy as np

    m = ts.DeviceMatrix.from_canonical(ctx, np.zeros((32, 2), dtype=np.uint32))
    out = (C.c_void_p * 1)()
    alpha = np.zeros(4, dtype=np.uint32)
    one = np.zeros(1, dtype=np.uint32)
    for prog in ([[9, 0, 0, 0]], [[0, 64, 0, 0]], [[0, 0, (1 << 28) | 2, 0]], [[4, 0, (3 << 28) | 0, 0]]):
        p = np.array(prog, dtype=np.uint32)
        rc = ctx._L.ts_quotient_values(ctx._h, m._h, 3, 0, p.ctypes.data_as(C.c_void_p), 1, one.ctypes.data_as(C.c_void_p), 0,
                                       one.ctypes.data_as(C.c_void_p), 0, alpha.ctypes.data_as(C.c_void_p), out)
        assert rc == 2  # TS_ERR_ARG


@pytest.mark.parametrize("widths", [[8, 8, 8], [8, 4], [12]])
def test_dot_ext_powers_blocks(ts, ctx, orc, widths):
    """ts_dot_ext_powers_blocks: a row held as several column blocks (equal power-of-two widths: one pass over all of
    them; ragged: one accumulating pass per block) against the oracle's dot over the concatenated matrix."""
    import ctypes as C

    import numpy as np

    rows = 96
    mats = [pc.rand_mat(50 + i, rows, w) for i, w in enumerate(widths)]
    alpha = np.array([5, 6, 7, 8], dtype=np.uint32)
    dev = [ts.DeviceMatrix.from_canonical(ctx, m) for m in mats]
    ap = C.c_void_p()
    am = ts.to_monty(alpha)
    ctx.check(ctx._L.ts_alpha_powers(ctx._h, am.ctypes.data_as(C.c_void_p), sum(widths), C.byref(ap)), "alpha_powers")
    acc = ts.DeviceMatrix.from_canonical(ctx, np.zeros((rows, 4), dtype=np.uint32))
    arr = (C.c_void_p * len(dev))(*[d._h for d in dev])
    ctx.check(ctx._L.ts_dot_ext_powers_blocks(ctx._h, arr, len(dev), ap, acc._h), "dot_ext_powers_blocks")
    ctx._L.ts_matrix_free(ap)
    ref = orc.dot_ext_powers(np.ascontiguousarray(np.concatenate(mats, axis=1)), alpha)
    assert np.array_equal(acc.to_canonical(), ref)


def test_stark_edge_sizes(ts, ctx, orc):
    """Smallest traces (2 and 4 rows) still prove and verify; a quotient domain larger than the committed LDE
    (log_quotient_degree > log_blowup) is an argument error, as get_evaluations_on_domain would refuse it."""
    import importlib

    import airs

    for n in (2, 4):
        trace = airs.fibonacci_trace(3, 5, n)
        pc.check_stark_prove_verify(ts, ctx, orc, airs.FibonacciAir(), trace, [3, 5, int(trace[-1, 1])], 2, tamper=False)
    st = importlib.import_module(ts.__name__ + ".stark")
    air = airs.MulAir(degree=5, reps=1)  # quotient degree 4
    mmcs = ts.Blake3MerkleMmcs(ctx)
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mmcs, ts.FriConfig(1, 2, 4, mmcs))
    with pytest.raises(ts.TapStarkError):
        st.prove(pcs, air, ts.BfChallenger(), airs.mul_trace(air, 16, 3), [])


def test_stark_wide_air(ts, ctx, orc):
    """A C4-shaped AIR (66 triples = 198 columns, not a multiple of 4; ~400-instruction constraint program)."""
    import airs

    air = airs.MulAir(degree=3, reps=66)
    pc.check_stark_prove_verify(ts, ctx, orc, air, airs.mul_trace(air, 1 << 4, 9), [], 2, num_queries=2, tamper=False)


def test_stark_quotient_golden(ts, ctx, golden):
    pc.check_stark_golden(ts, ctx, golden["stark_fibonacci"])


@pytest.mark.parametrize("layout", [0, 1])
def test_mmcs_big_layers(ts, ctx, orc, monkeypatch, layout):
    """Big layers go through tree_reduce3_kernel (one thread per 8 children, three levels per launch; from 2^19 children by
    default, from 2^14 here), followed by an injection layer (2^12 rows) and the shared-memory kernel for the top."""
    monkeypatch.setenv("TS_TREE3_MIN_LOG", "14")
    pc.check_mmcs(ts, ctx, orc, [(1 << 16, 3), (1 << 12, 5)], layout, indices=(0, 4097, (1 << 16) - 1))


def test_stark_counter_air(ts, ctx, orc):
    """Constants, negation, assert_one and when(<expression>) through the symbolic builder, the program compiler and the
    kernel's NEG / CONST operands (constraint degree 3 -> two quotient chunks)."""
    import airs

    n = 1 << 5
    pc.check_stark_prove_verify(ts, ctx, orc, airs.CounterAir(), airs.counter_trace(n), [n - 1], 2)
