"""Pins the CPU oracle (oracle/tapstark_oracle.c) against the reference's own known answers and the
committed golden vectors (tests/golden/golden.json, made by tests/golden/make_golden.py from
definitions + the `blake3` package).  No GPU."""
import numpy as np
import pytest

P = 0x78000001


def test_field_constants(orc):
    # basic/src/field/mod.rs:45,70-85 (test_subgroup): w4^4 = 1, w4^2 = -1; SURVEY App. A constants
    L = orc.lib()
    w4 = orc.two_adic_generator(2)
    assert L.or_bb_pow(w4, 4) == 1 and L.or_bb_pow(w4, 2) == P - 1
    assert orc.two_adic_generator(27) == 0x1A427A41 == pow(31, 15, P)
    assert orc.two_adic_generator(1) == P - 1
    assert L.or_to_monty(1) == 0x0FFFFFFE
    assert L.or_from_monty(L.or_to_monty(123456789)) == 123456789
    assert L.or_bb_inv(2) == 1006632961


def test_ef_mul_inv(orc):
    rng = np.random.default_rng(1)
    for _ in range(20):
        a = rng.integers(0, P, 4, dtype=np.uint32)
        b = rng.integers(0, P, 4, dtype=np.uint32)
        from oracle import pyref
        assert list(orc.ef_mul(a, b)) == pyref.ef_mul([int(x) for x in a], [int(x) for x in b])
        inv = np.zeros(4, dtype=np.uint32)
        orc.lib().or_ef_inv(orc._u32p(a), orc._u32p(inv))
        assert list(orc.ef_mul(a, inv)) == [1, 0, 0, 0]


def test_blake3_reference_kats(orc, golden):
    for k in golden["blake3"]["kats"]:
        assert orc.blake3(bytes.fromhex(k["input_hex"])).hex() == k["hash"], k["src"]


def test_blake3_pattern_vectors(orc, golden):
    for v in golden["blake3"]["pattern_mod251"]:
        data = bytes(i % 251 for i in range(v["len"]))
        assert orc.blake3(data).hex() == v["hash"], v["len"]


def test_challenger_reference_golden(orc, golden):
    # script_expr/src/challenger_expr.rs:278-296
    ch = orc.BfChallenger(ext=False)
    ch.observe(b"\x01\x01\x01\x01")
    assert ch.sample_bb() == golden["challenger"]["ref_golden"]["first"]
    ch.observe(b"\x01\x01\x01\x01")
    assert ch.sample_bb() == 1103171332


def test_challenger_script(orc, golden):
    ch = orc.BfChallenger()
    outs = iter(golden["challenger"]["outputs"])
    for op in golden["challenger"]["script"]:
        if op[0] == "observe_digest":
            ch.observe_digest(bytes.fromhex(op[1]))
        elif op[0] == "observe":
            ch.observe(bytes.fromhex(op[1]))
        elif op[0] == "sample_ef":
            assert list(ch.sample_ef()) == next(outs)
        elif op[0] == "sample_bb":
            assert ch.sample_bb() == next(outs)


def test_challenger_grind_and_fake_perm(orc):
    ch = orc.BfChallenger()
    ch.observe_digest(bytes(range(32)))
    v = ch.clone()
    w = ch.grind(8)
    assert w < 4096
    assert v.check_witness(8, w)
    for smaller in range(w):
        assert not orc.BfChallenger.clone(_replay(orc)).check_witness(8, smaller)
    # fri/tests/fri.rs:37-48 TestPermutation: state reversed
    f = orc.BfChallenger(fake_perm=True, ext=False)
    for i in range(8):
        f.observe(i + 1)
    # state = [1..8, 0*8] reversed -> out = state[8..16] = [8,7,...,1]; pop() yields 1 first
    assert [f.sample_bb() for _ in range(8)] == [1, 2, 3, 4, 5, 6, 7, 8]


def _replay(orc):
    ch = orc.BfChallenger()
    ch.observe_digest(bytes(range(32)))
    return ch


def test_lde_golden(orc, golden):
    for case in golden["lde"]:
        ev = np.array(case["evals"], dtype=np.uint32)
        got = orc.pcs_lde_committed(ev, case["added_bits"], case["shift"])
        assert got.tolist() == case["committed"], (case["log_n"], case["width"])


def test_dft_matches_definition(orc):
    rng = np.random.default_rng(7)
    for log_n in range(0, 8):
        m = rng.integers(0, P, (1 << log_n, 3), dtype=np.uint32)
        assert np.array_equal(orc.dft_batch(m), orc.naive_dft(m))
        assert np.array_equal(orc.idft_batch(orc.dft_batch(m)), m)


def test_lde_low_coset_is_input_domain(orc):
    # fri/src/two_adic_pcs.rs:247-258: first n committed rows, re-bit-reversed, are p on g*H_n; with shift
    # 1 they are the input itself.
    rng = np.random.default_rng(3)
    m = rng.integers(0, P, (64, 5), dtype=np.uint32)
    lde = orc.pcs_lde_committed(m, 2, 1)
    assert np.array_equal(orc.bit_reverse_rows(lde[:64]), m)


def test_fold_reference_property(orc):
    # fri/src/fold_even_odd.rs:65-95 with n = 2^10
    rng = np.random.default_rng(11)
    n = 1 << 10
    coeffs = rng.integers(0, P, (n, 1), dtype=np.uint32)
    evals = orc.dft_batch(coeffs)
    even, odd = orc.dft_batch(coeffs[0::2]), orc.dft_batch(coeffs[1::2])
    beta = int(rng.integers(0, P))
    expected = (even.astype(np.uint64) + beta * odd.astype(np.uint64)) % P
    folded = orc.fold_matrix_bb(orc.bit_reverse_rows(evals).reshape(-1), beta)
    folded = orc.bit_reverse_rows(folded.reshape(-1, 1))
    assert np.array_equal(folded.astype(np.uint64), expected)


def test_fold_ef_golden(orc, golden):
    for case in golden["fold_ef"]:
        got = orc.fold_matrix_ef(np.array(case["vals"], dtype=np.uint32), case["beta"])
        assert got.tolist() == case["out"]


def test_fold_row_matches_fold_matrix(orc):
    # verifier's fold_row (two_adic_pcs.rs:87-114) agrees with the prover's fold_matrix
    rng = np.random.default_rng(5)
    log_h = 4
    vals = rng.integers(0, P, (2 << log_h, 4), dtype=np.uint32)
    beta = rng.integers(0, P, 4, dtype=np.uint32)
    out = orc.fold_matrix_ef(vals, beta)
    for i in range(1 << log_h):
        assert np.array_equal(orc.fold_row_ef(i, log_h, beta, vals[2 * i], vals[2 * i + 1]), out[i])


def test_merkle_single_golden(orc, golden):
    for case in golden["merkle_single"]:
        rows = np.array(case["rows"], dtype=np.uint32)
        t = orc.mmcs_commit([rows])
        assert t.root.hex() == case["root"]
        assert t.layer(0)[0].tobytes().hex() == case["leaf0"]
        tp = orc.mmcs_commit([rows], orc.LAYOUT_PADDED)
        assert tp.root == t.root  # layouts coincide for a single matrix


def test_padded_layout_reference_vectors(orc, golden):
    # basic/src/tcs/mod.rs:520-602: mat_1 (4x2), mat_2 (4x4), mat_3 (8x1) of test_taptree_mmcs and the
    # per-leaf value lists written in its comment
    mat1 = np.array([[0, 1], [2, 1], [2, 2], [1, 0]], dtype=np.uint32)
    mat2 = np.array([[0, 1, 2, 1], [2, 2, 1, 0], [0, 1, 2, 1], [2, 2, 1, 0]], dtype=np.uint32)
    mat3 = np.array([[0], [1], [2], [1], [2], [2], [1], [0]], dtype=np.uint32)
    want = golden["padded_layout"]["leaves"]
    for i in range(8):
        assert orc.padded_leaf([mat1, mat2, mat3], i).tolist() == want[i]


@pytest.mark.parametrize("layout", [0, 1])
def test_mmcs_mixed_heights_roundtrip(orc, layout):
    # mirrors basic/src/mmcs/taptree_mmcs.rs:133-231 (commit -> open -> verify, mixed heights)
    rng = np.random.default_rng(9)
    mats = [rng.integers(0, P, s, dtype=np.uint32) for s in [(8, 3), (32, 2), (8, 5), (32, 1), (2, 4)]]
    t = orc.mmcs_commit(mats, layout)
    for idx in [0, 1, 7, 18, 31]:
        rows, path = t.open_batch(idx)
        for m, r in zip(mats, rows):
            assert np.array_equal(r, m[idx >> (5 - (m.shape[0].bit_length() - 1))])
        assert t.verify_batch(idx, rows, path)
        bad = [r.copy() for r in rows]
        bad[2][0] ^= 1
        assert not t.verify_batch(idx, bad, path)
        assert not t.verify_batch(idx ^ 1, rows, path)


def test_commit_phase_golden(orc, golden):
    g = golden["commit_phase"]
    cw = orc.pcs_lde_committed(np.array(g["evals"], dtype=np.uint32), g["log_blowup"])
    res = orc.fri_commit_phase([cw], g["log_blowup"], orc.BfChallenger(), want_layers=True)
    assert res["ok"]
    assert [c.hex() for c in res["commits"]] == g["commits"]
    assert res["betas"].tolist() == g["betas"]
    assert res["final_poly"].tolist() == g["final_poly"]


def test_commit_phase_mixed_inputs_verify(orc):
    # fri/tests/fri.rs:52-147 shape: inputs of several heights; restated verify_query accepts
    rng = np.random.default_rng(13)
    b = 1
    inputs = []
    for log_n in [6, 4, 3]:
        ev = rng.integers(0, P, (1 << log_n, 4), dtype=np.uint32)
        inputs.append(orc.pcs_lde_committed(ev, b))
    ch = orc.BfChallenger()
    res = orc.fri_commit_phase(inputs, b, ch, want_layers=True)
    assert res["ok"] and res["rounds"] == 6
    log_max = 7
    for q in [0, 5, 77, 127]:
        idx, folded = q, np.zeros(4, dtype=np.uint32)
        ro = {v.shape[0].bit_length() - 1: v for v in inputs}
        for r in range(res["rounds"]):
            lfh = log_max - 1 - r
            if lfh + 1 in ro:
                folded = (folded.astype(np.uint64) + ro[lfh + 1][idx >> 0 if lfh + 1 == log_max else idx]) % P
            layer = res["layers"][r]
            assert np.array_equal(folded.astype(np.uint32), layer[idx]) or r == 0
            pair = idx >> 1
            folded = orc.fold_row_ef(pair, lfh, res["betas"][r], layer[2 * pair], layer[2 * pair + 1])
            idx = pair
        assert np.array_equal(folded, res["final_poly"])


def test_stark_fibonacci_quotient_golden(golden, orc):
    """Row f3: the oracle's folder-based quotient values equal the big-int definition frozen in golden.json
    (uni-stark/src/prover.rs:122-194 on the reference's Fibonacci AIR)."""
    import numpy as np

    import airs
    from oracle import stark as OS

    g = golden["stark_fibonacci"]
    trace = np.array(g["trace"], dtype=np.uint32)
    assert np.array_equal(trace, airs.fibonacci_trace(0, 1, 1 << g["log_n"]))
    lde = orc.pcs_lde_committed(trace, g["log_blowup"])
    assert orc.mmcs_commit([lde]).root.hex() == g["trace_root"]
    ch = orc.BfChallenger()
    ch.observe_digest(bytes.fromhex(g["trace_root"]))
    assert [int(x) for x in ch.sample_ef()] == g["alpha"]
    chunks = OS.quotient_values(airs.FibonacciAir(), g["public_values"], lde, g["log_n"], 0, g["alpha"])
    assert len(chunks) == 1 and chunks[0].tolist() == g["quotient"]
