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


@pytest.mark.gpu
@pytest.mark.parametrize("log_n,width,limbs", [(3, 2, 1), (6, 8, 4), (9, 5, 1), (12, 3, 1), (12, 8, 4)])
def test_taptree_gpu(ts, ctx, log_n, width, limbs):
    from oracle import taptree as T

    _check(ts, ctx, T, log_n, width, limbs, seed=300 + log_n)
