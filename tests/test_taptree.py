"""TapTree commitment (first slice): the GPU tagged-SHA-256 leaf hashes over the templated leaf scripts, the
sorted-pair TapBranch tree and CompleteTaptree's leaf index table against oracle/taptree.py (hashlib), on 2^0 .. 2^12 leaves.
The locking-script bytes come from an external crate in the reference: PARITY UNPINNED for the template itself
(oracle/taptree.py header); what is pinned here is hashing, ordering and the permutation on identical scripts."""
import hashlib

import numpy as np
import pytest

import parity_cases as pc


def test_oracle_script_primitives():
    from oracle import taptree as T

    # script-number pushes ([MEM] bitcoin script minimal encoding) and BIP-341 tagged hashes against hashlib directly
    assert T.push_int(0) == b"\x00" and T.push_int(16) == b"\x60" and T.push_int(17) == b"\x01\x11"
    assert T.push_int(127) == b"\x01\x7f" and T.push_int(128) == b"\x02\x80\x00" and T.push_int(0x78000000) == b"\x04\x00\x00\x00\x78"
    assert T.push_int(0xFFFFFFFF) == b"\x05\xff\xff\xff\xff\x00"
    t = hashlib.sha256(b"TapLeaf").digest()
    assert T.tap_leaf_hash(b"\x51") == hashlib.sha256(t + t + b"\xc0\x01\x51").digest()
    # winternitz.rs:265-281: the public key is the (DIGITS + 1)-fold hash160 of secret || digit index
    h = bytes.fromhex("1234") + b"\x03"
    for _ in range(16):
        h = hashlib.new("ripemd160", hashlib.sha256(h).digest()).digest()
    assert T.generate_public_key("1234", 3) == h
    # checksig_verify: N digit blocks of 15 DUP/HASH160 pairs and one 20-byte key push each
    lock = T.locking_script_u32("1234")
    assert lock.count(bytes([0x76, 0xA9]) * 15) >= T.N and lock.count(b"\x14") >= T.N
    # build_tree on 4 leaves: a permutation of the index set; sorted pairs make the root blind to swaps INSIDE a pair
    # (combine_with_order sorts), not to moving a leaf into the other pair
    hs = [hashlib.sha256(bytes([i])).digest() for i in range(4)]
    root, perm = T.build_tree(hs)
    assert sorted(perm) == [0, 1, 2, 3]
    assert root == T.build_tree([hs[1], hs[0], hs[3], hs[2]])[0] and root != T.build_tree([hs[0], hs[2], hs[1], hs[3]])[0]


def test_c_oracle_sha256_pinned_to_hashlib(orc):
    """oracle/tapstark_oracle.c carries its own SHA-256 for the C++ host-mirror test: pinned to hashlib and to oracle/taptree.py."""
    from oracle import taptree as T

    L = orc.lib()
    rng = np.random.default_rng(5)
    for n in (0, 1, 55, 56, 63, 64, 65, 119, 120, 300, 1000):
        data = rng.integers(0, 256, n, dtype=np.uint8)
        out = np.zeros(32, dtype=np.uint8)
        L.or_sha256(orc._u8p(data if n else np.zeros(1, dtype=np.uint8)), n, orc._u8p(out))
        assert out.tobytes() == hashlib.sha256(data.tobytes()).digest()
        if n:
            L.or_tap_leaf_hash(orc._u8p(data), n, orc._u8p(out))
            assert out.tobytes() == T.tap_leaf_hash(data.tobytes())
    a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
    out = np.zeros(32, dtype=np.uint8)
    for x, y in ((a, b), (b, a)):
        L.or_tap_branch_hash(orc._u8p(x), orc._u8p(y), orc._u8p(out))
        assert out.tobytes() == T.tap_branch_hash(a.tobytes(), b.tobytes())[0]


def _check(ts, ctx, T, log_n, width, limbs, seed):
    n = 1 << log_n
    n_eval = width // limbs
    segs, order = T.template("00" * 20, [f"{e + 1:040x}" for e in range(n_eval)], limbs)
    rows = pc.rand_mat(seed, n, width) if width else np.zeros((n, 1), dtype=np.uint32)
    # a few small and boundary values so that every push length occurs
    if width:
        rows[0, 0], rows[n - 1, width - 1] = 0, 0x77FFFFFF
        if n > 2:
            rows[1, 0], rows[2, width - 1] = 16, 0x80
    want_root, want_perm, want_leaves = T.commit(segs, order, rows)
    dev = ts.DeviceMatrix.from_canonical(ctx, rows)
    got = ts.TapTreeCommit(ctx, dev, segs, order)
    assert got.level(0).tobytes() == b"".join(want_leaves), "TapLeaf hashes differ"
    assert got.root == want_root, "TapTree root differs"
    assert got.leaf_indices().tolist() == want_perm, "leaf index table differs"
    assert got.level(log_n).tobytes() == want_root
    got.free()


@pytest.mark.parametrize("log_n,width,limbs", [(0, 1, 1), (3, 2, 1), (5, 4, 4), (7, 3, 1)])
def test_taptree_emulated(log_n, width, limbs):
    from __graft_entry__ import load_pkg
    from emul import build_emul
    from oracle import taptree as T

    ts = load_pkg()
    ts.load_library(build_emul.build(), allow_emulated=True)
    ctx = ts.Context(0)
    _check(ts, ctx, T, log_n, width, limbs, seed=200 + log_n)


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


@pytest.mark.gpu
@pytest.mark.parametrize("log_n,width,limbs", [(3, 2, 1), (6, 8, 4), (9, 5, 1), (12, 3, 1), (12, 8, 4)])
def test_taptree_gpu(ts, ctx, log_n, width, limbs):
    from oracle import taptree as T

    _check(ts, ctx, T, log_n, width, limbs, seed=300 + log_n)


# ---- TapTreeMmcs host mirror (basic/src/mmcs/taptree_mmcs.rs) -------------------------------------------------------------
def _secrets(seed: bytes, count: int):
    """the product's stand-in secret generator, restated: SHA-256(seed || counter)[:20], counter from 1"""
    return [hashlib.sha256(seed + (k + 1).to_bytes(8, "little")).digest()[:20] for k in range(count)]


def test_product_template_matches_oracle():
    """tap-stark_b200/taptree.py builds its own script template (the product never imports oracle/): same bytes as the oracle's."""
    from importlib import import_module

    from __graft_entry__ import load_pkg
    from oracle import taptree as T

    tt = import_module(load_pkg().__name__ + ".taptree")
    for v in (0, 1, 16, 17, 127, 128, 255, 256, 0x7FFF, 0x8000, 0x77FFFFFF, 0x78000000, 0xFFFFFFFF, -1, -5):
        assert tt.script_num_push(v) == T.push_int(v)
    for limbs, n_eval in ((1, 3), (4, 2)):
        sec = _secrets(b"t", 1 + n_eval)
        use = tt.UseBComm(tt.BitCommitment.from_secret(sec[0]), [tt.BitCommitment.from_secret(s) for s in sec[1:]], limbs)
        segs, order = use.template()
        osegs, oorder = T.template(sec[0].hex(), [s.hex() for s in sec[1:]], limbs)
        assert segs == osegs and order == oorder
        row = [5, 0, 0x77FFFFFF, 200, 16, 17, 128, 3][: n_eval * limbs]
        script = use.leaf_script(9, row)
        assert script == T.leaf_script(osegs, oorder, 9, row)
        assert tt.tap_leaf_hash(script) == T.tap_leaf_hash(script)


def _check_mmcs(ts, ctx, orc, shapes, limbs, num_queries, seed):
    """shapes: [(log_height, width)] in commit order.  Roots, openings, proofs against oracle/taptree.py over the oracle's padded rows."""
    from importlib import import_module

    from oracle import taptree as T

    tt = import_module(ts.__name__ + ".taptree")
    mats = [pc.rand_mat(seed + i, 1 << lh, w) for i, (lh, w) in enumerate(shapes)]
    hmax = max(m.shape[0] for m in mats)
    rows = np.stack([orc.padded_leaf(mats, leaf) for leaf in range(hmax)])  # PolyTCS::padding_matrix, restated in the C oracle
    n_eval = rows.shape[1] // limbs
    mmcs = tt.TapTreeMmcs(ctx, num_queries, limbs, tt.SecretGen(b"seed%d" % seed))
    dev = [ts.DeviceMatrix.from_canonical(ctx, m) for m in mats]
    assert np.array_equal(mmcs.padded_rows(dev).to_canonical(), rows)
    roots, data = mmcs.commit(dev)
    assert len(roots) == num_queries and mmcs.get_matrices(data) == dev
    sec = _secrets(b"seed%d" % seed, num_queries * (1 + n_eval))
    for q in range(num_queries):
        s = sec[q * (1 + n_eval): (q + 1) * (1 + n_eval)]
        segs, order = T.template(s[0].hex(), [x.hex() for x in s[1:]], limbs)
        want_root, want_perm, want_leaves = T.commit(segs, order, rows)
        assert roots[q] == want_root, "root of query tree %d differs" % q
        for index in sorted({0, 1 % hmax, hmax // 2, hmax - 1}):
            opened, proof = mmcs.open_batch(q, index, data)
            for m, o in zip(mats, opened):  # taptree_mmcs.rs:54-64
                assert np.array_equal(o, m[index >> ((hmax // m.shape[0]).bit_length() - 1)])
            assert proof.leaf_script == T.leaf_script(segs, order, index, rows[index])
            h = want_leaves[index]
            for sib in proof.merkle_branch:  # verify_inclusion with the oracle's sorted-pair hash
                h, _ = T.tap_branch_hash(h, sib)
            assert h == want_root and len(proof.merkle_branch) == hmax.bit_length() - 1
            assert data[q].tree.open(index)[1] == want_perm[index]
            heights = [m.shape[0] for m in mats]
            mmcs.verify_batch(q, opened, proof, roots, heights)
            bad = [o.copy() for o in opened]
            bad[-1][0] ^= 1
            with pytest.raises(ts.TapStarkError):
                mmcs.verify_batch(q, bad, proof, roots, heights)
            if num_queries > 1:
                with pytest.raises(ts.TapStarkError):  # a proof of one query tree does not verify under another root
                    mmcs.verify_batch((q + 1) % num_queries, opened, proof, roots, heights)
            if proof.merkle_branch:
                forged = tt.CommitedProof(proof.leaf_script, proof.merkle_branch[::-1] if len(proof.merkle_branch) > 1
                                          else [bytes(32)], proof.use_bcs, proof.query_index)
                with pytest.raises(ts.TapStarkError):
                    mmcs.verify_batch(q, opened, forged, roots, heights)


MMCS_CASES = [
    ([(2, 2)], 1, 2),                     # one FRI-layer-like matrix: DEFAULT_MATRIX_WIDTH = 2 (taptree_mmcs.rs:19)
    ([(2, 2), (2, 4), (3, 1)], 1, 2),     # test_taptree_mmcs's shapes (taptree_mmcs.rs:139-215): heights 4, 4, 8
    ([(4, 8)], 4, 1),                     # extension-field rows: 2 evaluations of 4 limbs
    ([(0, 3)], 1, 1),                     # a single leaf
    ([(5, 1), (3, 2), (5, 2), (1, 1)], 1, 3),
]


@pytest.mark.parametrize("shapes,limbs,num_queries", MMCS_CASES)
def test_taptree_mmcs_emulated(orc, shapes, limbs, num_queries):
    from __graft_entry__ import load_pkg
    from emul import build_emul

    ts = load_pkg()
    ts.load_library(build_emul.build(), allow_emulated=True)
    ctx = ts.Context(0)
    _check_mmcs(ts, ctx, orc, shapes, limbs, num_queries, seed=400 + len(shapes))


@pytest.mark.gpu
@pytest.mark.parametrize("shapes,limbs,num_queries", MMCS_CASES + [([(10, 2), (8, 4)], 1, 2), ([(11, 8)], 4, 2)])
def test_taptree_mmcs_gpu(ts, ctx, orc, shapes, limbs, num_queries):
    _check_mmcs(ts, ctx, orc, shapes, limbs, num_queries, seed=500 + len(shapes))
