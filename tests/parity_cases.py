"""Parity cases shared by the emulated (no GPU) and the real-GPU test modules.  Every case drives the C ABI
through the host-side mirror (tap-stark_b200/__init__.py) and compares bit-exactly with the CPU oracle."""
import numpy as np

P = 0x78000001


def rand_mat(seed, rows, width):
    rng = np.random.default_rng(seed)
    return rng.integers(0, P, (rows, width), dtype=np.uint32)


# ---- LDE (fri/src/two_adic_pcs.rs:235-240) -----------------------------------------------------
def check_lde(ts, ctx, orc, log_n, width, added_bits, shift=31, seed=0):
    ev = rand_mat(seed, 1 << log_n, width)
    dft = ts.GpuDft(ctx)
    got = dft.coset_lde_batch(ts.DeviceMatrix.from_canonical(ctx, ev), added_bits, shift, committed_order=True)
    want = orc.pcs_lde_committed(ev, added_bits, shift)
    assert got.rows == want.shape[0] and got.width == width
    assert np.array_equal(got.to_canonical(), want), f"LDE mismatch log_n={log_n} w={width} b={added_bits}"


def check_lde_natural_and_host(ts, ctx, orc, log_n, width, added_bits, seed=1):
    ev = rand_mat(seed, 1 << log_n, width)
    dft = ts.GpuDft(ctx)
    nat = dft.coset_lde_batch(ts.DeviceMatrix.from_canonical(ctx, ev), added_bits, 31)
    want = orc.coset_lde_batch(ev, added_bits, 31)
    assert np.array_equal(nat.to_canonical(), want)
    assert np.array_equal(dft.coset_lde_batch_host(ev, added_bits, 31), want)


def check_dft_family(ts, ctx, orc, log_n, width, seed=2):
    m = rand_mat(seed, 1 << log_n, width)
    dft = ts.GpuDft(ctx)
    dm = ts.DeviceMatrix.from_canonical(ctx, m)
    assert np.array_equal(dft.dft_batch(dm).to_canonical(), orc.dft_batch(m))
    assert np.array_equal(dft.idft_batch(dm).to_canonical(), orc.idft_batch(m))
    assert np.array_equal(dft.coset_dft_batch(dm, 31).to_canonical(), orc.coset_dft_batch(m, 31))
    assert np.array_equal(dft.lde_batch(dm, 1).to_canonical(), orc.coset_lde_batch(m, 1, 1))


# ---- MMCS ------------------------------------------------------------------------------------------
def check_mmcs(ts, ctx, orc, shapes, layout, seed=3, indices=(0, 1)):
    mats = [rand_mat(seed + i, h, w) for i, (h, w) in enumerate(shapes)]
    mmcs = ts.Blake3MerkleMmcs(ctx, layout)
    dms = [ts.DeviceMatrix.from_canonical(ctx, m) for m in mats]
    root, data = mmcs.commit(dms)
    ref = orc.mmcs_commit(mats, layout)
    assert root == ref.root, f"root mismatch shapes={shapes} layout={layout}"
    assert np.array_equal(data.layer(0), ref.layer(0))
    hmax = max(h for h, _ in shapes)
    assert mmcs.get_max_height(data) == hmax
    for idx in indices:
        idx = idx % hmax
        rows, path = mmcs.open_batch(idx, data)
        rrows, rpath = ref.open_batch(idx)
        for a, b in zip(rows, rrows):
            assert np.array_equal(a, b)
        assert np.array_equal(path, rpath)
        mmcs.verify_batch([h for h, _ in shapes], rows, idx, path, root)
        assert ref.verify_batch(idx, rows, path)
        bad = [r.copy() for r in rows]
        bad[0][0] = (int(bad[0][0]) + 1) % P
        try:
            mmcs.verify_batch([h for h, _ in shapes], bad, idx, path, root)
            raise AssertionError("tampered opening verified")
        except ts.TapStarkError:
            pass


# ---- fold (fri/src/two_adic_pcs.rs:116-147, fri/src/fold_even_odd.rs) ------------------------------
def check_fold_ext(ts, ctx, orc, log_h, seed=4):
    vals = rand_mat(seed, 2 << log_h, 4)
    beta = rand_mat(seed + 1, 1, 4)[0]
    got = ts.fold_even_odd(ctx, vals, beta)
    assert np.array_equal(got, orc.fold_matrix_ef(vals, beta)), f"fold mismatch log_h={log_h}"


def check_fold_hash_shard(ts, ctx, orc, log_h, shards, with_addend, seed=44):
    """ts_fri_fold_hash_shard: every shard's folded rows equal the oracle's fold (+ the next input), and the digests it emits are
    Blake3 of the next round's rows (two folded extension elements as canonical LE words, fri/src/prover.rs:112-113)."""
    import ctypes as C

    h = 1 << log_h
    vals = rand_mat(seed, 2 * h, 4)
    beta = rand_mat(seed + 1, 1, 4)[0]
    addend = rand_mat(seed + 2, h, 4) if with_addend else None
    want = orc.fold_matrix_ef(vals, beta)
    if with_addend:
        want = ((want.astype(np.uint64) + addend) % P).astype(np.uint32)
    L = ts.lib()
    src = ts.DeviceMatrix.from_canonical(ctx, vals.reshape(h, 8))
    add_d = ts.DeviceMatrix.from_canonical(ctx, addend) if with_addend else None
    bm = ts.to_monty(beta)
    h_l = h // shards
    for r in range(shards):
        out = ts.DeviceMatrix.from_canonical(ctx, np.zeros((h_l, 4), dtype=np.uint32))
        dig = ts.DeviceMatrix.from_canonical(ctx, np.zeros((h_l // 2, 8), dtype=np.uint32))
        ctx.check(L.ts_fri_fold_hash_shard(ctx._h, C.c_void_p(src.device_ptr + r * h_l * 32), h, r * h_l, h_l,
                                           bm.ctypes.data_as(C.c_void_p),
                                           C.c_void_p(add_d.device_ptr + r * h_l * 16) if with_addend else None,
                                           C.c_void_p(out.device_ptr), C.c_void_p(dig.device_ptr)), "fold_hash_shard")
        got = out.to_canonical()
        assert np.array_equal(got, want[r * h_l : (r + 1) * h_l]), f"fold_hash shard {r}/{shards} log_h={log_h}"
        d = dig.to_monty_host()  # raw digest words
        rows = want[r * h_l : (r + 1) * h_l].reshape(-1, 8)
        for j in sorted({0, 1, h_l // 4, h_l // 2 - 1}):
            assert d[j].astype("<u4").tobytes() == orc.blake3(rows[j].astype("<u4").tobytes()), f"digest {j} of shard {r}"
    if log_h >= 9:
        assert L.ts_fri_fold_hash_shard(ctx._h, C.c_void_p(src.device_ptr), h, 256, 256, bm.ctypes.data_as(C.c_void_p), None,
                                        C.c_void_p(src.device_ptr), C.c_void_p(src.device_ptr)) != 0  # not a multiple of 512


def check_fold_base_reference_property(ts, ctx, orc, log_n=10, seed=5):
    """fri/src/fold_even_odd.rs:65-95 verbatim, on the device path."""
    n = 1 << log_n
    coeffs = rand_mat(seed, n, 1)
    dft = ts.GpuDft(ctx)
    dev = lambda a: ts.DeviceMatrix.from_canonical(ctx, a)
    evals = dft.dft_batch(dev(coeffs)).to_canonical()
    even = dft.dft_batch(dev(coeffs[0::2])).to_canonical()
    odd = dft.dft_batch(dev(coeffs[1::2])).to_canonical()
    beta = int(rand_mat(seed + 1, 1, 1)[0, 0])
    expected = (even.astype(np.uint64) + beta * odd.astype(np.uint64)) % P
    folded = ts.fold_even_odd(ctx, orc.bit_reverse_rows(evals).reshape(-1), beta)
    folded = orc.bit_reverse_rows(folded.reshape(-1, 1))
    assert np.array_equal(folded.astype(np.uint64), expected)


# ---- challenger (basic/src/challenger/mod.rs) -------------------------------------------------------
def check_challenger(ts, golden):
    ch = ts.BfChallenger(ext=False)
    ch.observe(b"\x01\x01\x01\x01")
    assert ch.sample() == golden["challenger"]["ref_golden"]["first"]
    ch.observe(b"\x01\x01\x01\x01")
    assert ch.sample() == 1103171332  # script_expr/src/challenger_expr.rs:278-296
    ch = ts.BfChallenger()
    outs = iter(golden["challenger"]["outputs"])
    for op in golden["challenger"]["script"]:
        if op[0] == "observe_digest":
            ch.observe(bytes.fromhex(op[1]))
        elif op[0] == "observe":
            ch.observe(bytes.fromhex(op[1]))
        elif op[0] == "sample_ef":
            assert list(ch.sample_ext()) == next(outs)
        else:
            assert ch.sample_base() == next(outs)


def check_challenger_grind(ts, orc):
    a, b = ts.BfChallenger(), orc.BfChallenger()
    d = bytes(range(32))
    a.observe(d)
    b.observe_digest(d)
    v = a.clone()
    w = a.grind(8)
    assert w == b.grind(8)
    assert v.check_witness(8, w)
    assert a.sample_bits(10) == b.sample_bits(10)


# ---- commit phase (fri/src/prover.rs:93-141) -----------------------------------------------------------
def check_commit_phase(ts, ctx, orc, log_ns, log_blowup, seed=6):
    inputs = [orc.pcs_lde_committed(rand_mat(seed + i, 1 << ln, 4), log_blowup) for i, ln in enumerate(log_ns)]
    cfg = ts.FriConfig(log_blowup, 4, 8, ts.Blake3MerkleMmcs(ctx))
    res = ts.bf_commit_phase(cfg, [ts.DeviceMatrix.from_canonical(ctx, v) for v in inputs], ts.BfChallenger())
    ref = orc.fri_commit_phase(inputs, log_blowup, orc.BfChallenger(), want_layers=True)
    assert ref["ok"]
    assert res.commits == ref["commits"]
    assert np.array_equal(res.final_poly, ref["final_poly"])
    assert len(res.data) == ref["rounds"]
    # prover data of round r opens the committed layer r (bf_answer_query, prover.rs:69-90)
    mm = cfg.mmcs
    for r, pd in enumerate(res.data[:3]):
        idx = 1 % (ref["layers"][r].shape[0] // 2)
        rows, path = mm.open_batch(idx, pd)
        assert np.array_equal(rows[0], ref["layers"][r].reshape(-1, 8)[idx])
        mm.verify_batch([ref["layers"][r].shape[0] // 2], rows, idx, path, res.commits[r])


def check_commit_phase_rejects_high_degree(ts, ctx):
    bad = rand_mat(9, 64, 4)  # random codeword: not low degree (prover.rs:130-134 assert)
    cfg = ts.FriConfig(1, 4, 8, ts.Blake3MerkleMmcs(ctx))
    try:
        ts.bf_commit_phase(cfg, [ts.DeviceMatrix.from_canonical(ctx, bad)], ts.BfChallenger())
    except ts.TapStarkError as e:
        assert "not constant" in str(e)
    else:
        raise AssertionError("high-degree input accepted")


# ---- PCS commit (fri/src/two_adic_pcs.rs:227-258) ---------------------------------------------------------
def check_pcs_commit(ts, ctx, orc, shapes, log_blowup, layout=0, seed=7):
    mmcs = ts.Blake3MerkleMmcs(ctx, layout)
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mmcs, ts.FriConfig(log_blowup, 2, 8, mmcs))
    evs = [rand_mat(seed + i, 1 << ln, w) for i, (ln, w) in enumerate(shapes)]
    doms = [pcs.natural_domain_for_degree(1 << ln) for ln, _ in shapes]
    root, data = pcs.commit([(d, ts.DeviceMatrix.from_canonical(ctx, e)) for d, e in zip(doms, evs)])
    ldes = [orc.pcs_lde_committed(e, log_blowup) for e in evs]
    assert root == orc.mmcs_commit(ldes, layout).root
    root2, _ = pcs.commit_host([(d, ts.to_monty(e)) for d, e in zip(doms, evs)])
    assert root2 == root
    # get_evaluations_on_domain: evaluations on g*H_n, natural order (two_adic_pcs.rs:247-258)
    ln, _ = shapes[0]
    got = pcs.get_evaluations_on_domain(data, 0, ts.TwoAdicMultiplicativeCoset(ln, 31))
    assert np.array_equal(got, orc.coset_dft_batch(orc.idft_batch(evs[0]), 31))
    # quotient-chunk style domain (uni-stark/src/prover.rs:78-83): domain.shift = 31 => LDE shift 1
    q = rand_mat(seed + 50, 1 << ln, 4)
    rootq, _ = pcs.commit([(ts.TwoAdicMultiplicativeCoset(ln, 31), ts.DeviceMatrix.from_canonical(ctx, q))])
    assert rootq == orc.mmcs_commit([orc.pcs_lde_committed(q, log_blowup, 1)], layout).root


def check_dot_ext_powers(ts, ctx, orc, rows, width, seed=8):
    m = rand_mat(seed, rows, width)
    alpha = rand_mat(seed + 1, 1, 4)[0]
    mmcs = ts.Blake3MerkleMmcs(ctx)
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mmcs, ts.FriConfig(1, 2, 8, mmcs))
    got = pcs.dot_ext_powers(ts.DeviceMatrix.from_canonical(ctx, m), alpha).to_canonical()
    acc = np.zeros((rows, 4), dtype=np.uint64)
    cur = np.array([1, 0, 0, 0], dtype=np.uint32)
    for c in range(width):
        acc = (acc + m[:, c : c + 1].astype(np.uint64) * cur.astype(np.uint64)) % P
        cur = orc.ef_mul(cur, alpha)
    assert np.array_equal(got, acc.astype(np.uint32))


# ---- interpolate_coset (fri/src/two_adic_pcs.rs:358-369) --------------------------------------------------------
def check_interpolate_low_coset(ts, ctx, orc, log_n, width, log_blowup, seed=60):
    """p_c(z) from the low coset of a committed LDE == sum_k coeff[k][c] z^k (coefficients from the oracle's inverse DFT)."""
    import ctypes as C

    n = 1 << log_n
    ev = rand_mat(seed, n, width)
    z = rand_mat(seed + 1, 1, 4)[0]
    coeffs = orc.idft_batch(ev).astype(np.uint64)          # p on H_n, natural order -> coefficients
    zpow = np.zeros((n, 4), dtype=np.uint64)                # z^k by doubling
    zpow[0, 0] = 1
    m, zm_ = 1, z.astype(np.uint32)
    while m < n:
        # zpow[m:2m] = zpow[0:m] * z^m, vectorised extension-field product
        a, b = zpow[:m], zm_.astype(np.uint64)
        out = np.zeros((m, 4), dtype=np.uint64)
        for i in range(4):
            for j in range(4):
                t = (a[:, i] * b[j]) % P
                if i + j >= 4:
                    t = (t * 11) % P
                out[:, (i + j) % 4] = (out[:, (i + j) % 4] + t) % P
        zpow[m : 2 * m] = out
        zm_ = orc.ef_mul(zm_, zm_)
        m *= 2
    want = np.zeros((width, 4), dtype=np.uint32)
    for k in range(4):
        want[:, k] = ((coeffs * zpow[:, k : k + 1]) % P).sum(axis=0) % P
    dm = ts.DeviceMatrix.from_canonical(ctx, ev)
    lde = ts.GpuDft(ctx).coset_lde_batch(dm, log_blowup, 31, committed_order=True)  # committed order, shift g (two_adic_pcs.rs:235-243)
    L = ts.lib()
    zmont = ts.to_monty(z)
    h = C.c_void_p()
    ctx.check(L.ts_inv_denoms(ctx._h, log_n + log_blowup, zmont.ctypes.data_as(C.c_void_p), C.byref(h)), "inv_denoms")
    inv = ts.DeviceMatrix(ctx, h)
    ys = np.empty((width, 4), dtype=np.uint32)
    ctx.check(L.ts_interpolate_low_coset(ctx._h, lde._h, n, zmont.ctypes.data_as(C.c_void_p), inv._h,
                                         ys.ctypes.data_as(C.c_void_p)), "interpolate_low_coset")
    assert np.array_equal(ts.from_monty(ys), want)


# ---- Pcs::open + verify (fri/src/two_adic_pcs.rs:260-530, fri/src/prover.rs, fri/src/verifier.rs) ------------------
def check_pcs_open_verify(ts, ctx, orc, round_shapes, log_blowup, num_queries=6, pow_bits=4, seed=70):
    """round_shapes: [[(log_n, width, n_points)] per commit round].  The device `open` must (1) return the
    oracle's opened values, (2) feed FRI the oracle's reduced openings (same layer commitments / final poly) and
    (3) produce a proof the restated reference verifier accepts; tampering must be rejected."""
    from oracle import verifier as V

    mmcs = ts.Blake3MerkleMmcs(ctx)
    fri = ts.FriConfig(log_blowup, num_queries, pow_bits, mmcs)
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mmcs, fri)
    ch, och = ts.BfChallenger(), orc.BfChallenger()
    rounds_dev, rounds_ora, commits = [], [], []
    rng = np.random.default_rng(seed)
    for ri, shapes in enumerate(round_shapes):
        evs = [rand_mat(seed + 10 * ri + i, 1 << ln, w) for i, (ln, w, _) in enumerate(shapes)]
        doms = [pcs.natural_domain_for_degree(1 << ln) for ln, _, _ in shapes]
        root, data = pcs.commit([(d, ts.DeviceMatrix.from_canonical(ctx, e)) for d, e in zip(doms, evs)])
        ldes = [orc.pcs_lde_committed(e, log_blowup) for e in evs]
        assert root == orc.mmcs_commit(ldes).root
        ch.observe(root)
        och.observe_digest(root)
        commits.append(root)
        rounds_dev.append((data, shapes))
        rounds_ora.append(ldes)
    zeta = [int(x) for x in ch.sample()]
    assert zeta == [int(x) for x in och.sample_ef()]
    pts_dev, pts_ora = [], []
    for (data, shapes), ldes in zip(rounds_dev, rounds_ora):
        per_mat = []
        for (ln, w, npts) in shapes:
            wgen = orc.two_adic_generator(ln)
            per_mat.append([[c * pow(wgen, k, P) % P for c in zeta] for k in range(npts)])  # zeta, zeta*w, ...
        pts_dev.append((data, per_mat))
        pts_ora.append(list(zip(ldes, per_mat)))
    och_v = och.clone()
    ch_abi = ch.clone()
    opened, proof = pcs.open(pts_dev, ch)
    # the same opening through the single C-ABI call (ts_pcs_open): its bytes are the oracle's postcard encoding of the
    # host-orchestrated result, they decode to a proof the restated verifier accepts, and the transcript continues equally
    from oracle import serialize as S
    import importlib

    blob = pcs.open_bytes(pts_dev, ch_abi)
    assert blob == S.encode_opening(opened, proof), "ts_pcs_open bytes differ from the oracle's encoding"
    opened_abi, proof_abi = importlib.import_module("tapstark_b200.proofio").decode_opening(blob)
    assert [int(x) for x in ch_abi.sample()] == [int(x) for x in ch.clone().sample()]
    # (1)+(2): oracle's open
    alpha = [int(x) for x in och.sample_ef()]
    o_opened, o_reduced = V.pcs_open_reduced(pts_ora, log_blowup, alpha)
    for r_dev, r_ora in zip(opened, o_opened):
        for m_dev, m_ora in zip(r_dev, r_ora):
            for y_dev, y_ora in zip(m_dev, m_ora):
                assert y_dev.tolist() == y_ora
    fri_in = [o_reduced[lh] for lh in sorted(o_reduced, reverse=True)]
    ref = orc.fri_commit_phase(fri_in, log_blowup, och)
    assert ref["ok"] and proof.commit_phase_commits == ref["commits"]
    assert proof.final_poly.tolist() == ref["final_poly"].tolist()
    assert proof.pow_witness == och.grind(pow_bits)
    # (3): restated reference verifier
    v_rounds = []
    for commit, (data, per_mat), (_, shapes), o_r in zip(commits, pts_dev, rounds_dev, opened):
        v_rounds.append((commit, [(ln, list(zip(pts, ys))) for (ln, _, _), pts, ys in zip(shapes, per_mat, o_r)]))
    assert V.pcs_verify(log_blowup, num_queries, pow_bits, v_rounds, proof, och_v.clone())
    v_rounds_abi = []
    for commit, (data, per_mat), (_, shapes), o_r in zip(commits, pts_dev, rounds_dev, opened_abi):
        v_rounds_abi.append((commit, [(ln, list(zip(pts, [y.tolist() for y in ys]))) for (ln, _, _), pts, ys in zip(shapes, per_mat, o_r)]))
    assert V.pcs_verify(log_blowup, num_queries, pow_bits, v_rounds_abi, proof_abi, och_v.clone())
    bad = proof.query_proofs[0].commit_phase_openings[0][0][0]
    bad[0] = (int(bad[0]) + 1) % P
    try:
        V.pcs_verify(log_blowup, num_queries, pow_bits, v_rounds, proof, och_v.clone())
    except V.VerifyError:
        pass
    else:
        raise AssertionError("tampered proof accepted")
    bad[0] = (int(bad[0]) - 1) % P
    opened[0][0][0][0][0] = (int(opened[0][0][0][0][0]) + 1) % P  # wrong claimed p(zeta)
    try:
        V.pcs_verify(log_blowup, num_queries, pow_bits, v_rounds, proof, och_v.clone())
    except V.VerifyError:
        pass
    else:
        raise AssertionError("wrong opened value accepted")


# ---- f3: quotient values and the uni-stark prove -> verify round trip ---------------------------------------------
def check_stark_prove_verify(ts, ctx, orc, air, trace, public_values, log_blowup, num_queries=4, pow_bits=4, tamper=True):
    """uni_stark::prove on the device path (tap-stark_b200/stark.py) against the oracle: (1) the quotient chunks the
    constraint-program kernel produces equal the oracle's row-by-row folder evaluation, (2) the proof is accepted by
    the restated uni_stark::verify, (3) a wrong public value / a tampered opened value is rejected."""
    from importlib import import_module

    from oracle import stark as OS

    st = import_module(ts.__name__ + ".stark")
    mmcs = ts.Blake3MerkleMmcs(ctx)
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mmcs, ts.FriConfig(log_blowup, num_queries, pow_bits, mmcs))
    log_n = trace.shape[0].bit_length() - 1
    log_qd = st.get_log_quotient_degree(air, len(public_values))
    # (1) quotient chunks
    dom = pcs.natural_domain_for_degree(trace.shape[0])
    root, data = pcs.commit([(dom, ts.DeviceMatrix.from_canonical(ctx, trace))])
    ch = ts.BfChallenger()
    ch.observe(root)
    alpha = [int(x) for x in ch.sample()]
    chunks = st.quotient_values(pcs, data, air, public_values, log_n, log_qd, alpha)
    lde = orc.pcs_lde_committed(trace, log_blowup)
    ref = OS.quotient_values(air, public_values, lde, log_n, log_qd, alpha)
    assert len(chunks) == len(ref) == 1 << log_qd
    for c_dev, c_ref in zip(chunks, ref):
        assert np.array_equal(c_dev.to_canonical(), c_ref)
    # (2) full proof
    proof = st.prove(pcs, air, ts.BfChallenger(), trace, public_values)
    assert proof.commitments.trace == root and proof.degree_bits == log_n
    assert OS.verify(log_blowup, num_queries, pow_bits, air, orc.BfChallenger(), proof, public_values, log_qd)
    # serialized proof bytes (uni-stark/src/proof.rs:19-37, postcard): the product's bytes -- its opening_proof part is
    # what ts_pcs_open emitted, verbatim -- equal the oracle's independent encoding of the decoded proof objects
    from oracle import serialize as S

    assert proof.to_bytes() == S.encode_stark_proof(proof)
    if not tamper:
        return proof
    # (3) rejection
    bad_pis = list(public_values)
    if bad_pis:
        bad_pis[-1] = (int(bad_pis[-1]) + 1) % P
        try:
            OS.verify(log_blowup, num_queries, pow_bits, air, orc.BfChallenger(), proof, bad_pis, log_qd)
        except OS.VerificationError as e:
            assert "OodEvaluationMismatch" in str(e)
        else:
            raise AssertionError("wrong public value accepted")
    v = proof.opened_values.trace_local
    v[0][0] = (int(v[0][0]) + 1) % P
    try:
        OS.verify(log_blowup, num_queries, pow_bits, air, orc.BfChallenger(), proof, public_values, log_qd)
    except OS.VerificationError:
        pass
    else:
        raise AssertionError("tampered opened value accepted")
    v[0][0] = (int(v[0][0]) - 1) % P
    return proof


def check_stark_golden(ts, ctx, g):
    """Device quotient values of the Fibonacci AIR against the golden fixture (big-int definition)."""
    from importlib import import_module

    import airs

    st = import_module(ts.__name__ + ".stark")
    mmcs = ts.Blake3MerkleMmcs(ctx)
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mmcs, ts.FriConfig(g["log_blowup"], 2, 4, mmcs))
    trace = np.array(g["trace"], dtype=np.uint32)
    root, data = pcs.commit([(pcs.natural_domain_for_degree(trace.shape[0]), ts.DeviceMatrix.from_canonical(ctx, trace))])
    assert root.hex() == g["trace_root"]
    chunks = st.quotient_values(pcs, data, airs.FibonacciAir(), g["public_values"], g["log_n"], 0, g["alpha"])
    assert len(chunks) == 1 and chunks[0].to_canonical().tolist() == g["quotient"]
