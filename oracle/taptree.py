"""TEST INFRASTRUCTURE (oracle/): CPU restatement of the commitment the reference really computes.

    leaf script        basic/src/tcs/mod.rs:197-225 (CommitedLeaf::generate_script): index bit-commitment locking script,
                       push(index) OP_EQUALVERIFY, then per evaluation its locking script and its u32 limbs pushed in
                       reverse order each followed by OP_EQUALVERIFY, then OP_1
    locking script     `bc.locking_script_with_type(CompressType::U32)` lives in the EXTERNAL crate `bitcomm` (git dependency,
                       not on this box).  Restated from the in-tree twin: scripts/src/bit_comm/bit_comm_u32.rs:75-80
                       (recover_message_at_stack = checksig_verify ‖ u32_compress), scripts/src/bit_comm/winternitz.rs:170-263
                       (checksig_verify), scripts/src/u32/u32_std.rs:122-175 (u32_compress), scripts/src/pseudo.rs:106-113
                       (OP_256MUL).  PARITY UNPINNED: the external crate's bytes, its secret generator and the `script!`
                       macro's push encoding cannot be checked here; secrets below are fixed test values.
    public keys        scripts/src/bit_comm/winternitz.rs:265-281 (hash160 chain of length DIGITS + 1 over secret ‖ digit index)
    leaf / branch hash BIP-341 tagged SHA-256 as rust-bitcoin computes them for `NodeInfo::new_leaf_with_ver(script, TapScript)`
                       (basic/src/tcs/builder.rs:26) and `NodeInfo::combine_with_order` (:64): TapLeaf = H_TapLeaf(0xc0 ‖
                       compact_size(len) ‖ script), TapBranch = H_TapBranch(min(a, b) ‖ max(a, b)) byte-lexicographically [MEM]
    tree + permutation basic/src/tcs/builder.rs:38-93 (build_tree): pairwise combination level by level; when the right node
                       sorts first the two leaf ranges swap; the result is reverse_idx_dict of that permutation.

Used by tests/ only, as the checker of the GPU TapTree kernels (csrc/sha256.cuh).
"""
from __future__ import annotations

import hashlib
from typing import List, Sequence, Tuple

LOG_D = 4
DIGITS = (1 << LOG_D) - 1
N0, N1 = 8, 2
N = N0 + N1

OP = {
    "OP_0": 0x00, "OP_1NEGATE": 0x4F, "OP_1": 0x51, "OP_IF": 0x63, "OP_ELSE": 0x67, "OP_ENDIF": 0x68, "OP_TOALTSTACK": 0x6B,
    "OP_FROMALTSTACK": 0x6C, "OP_2DROP": 0x6D, "OP_DUP": 0x76, "OP_PICK": 0x79, "OP_ROLL": 0x7A, "OP_ROT": 0x7B, "OP_SWAP": 0x7C,
    "OP_TUCK": 0x7D, "OP_EQUALVERIFY": 0x88, "OP_NEGATE": 0x8F, "OP_ADD": 0x93, "OP_SUB": 0x94, "OP_MIN": 0xA3,
    "OP_GREATERTHAN": 0xA0, "OP_HASH160": 0xA9,
}


def push_int(v: int) -> bytes:
    """bitcoin script-number push, minimal encoding (what `{ n }` in the script! macro emits [MEM])."""
    if v == 0:
        return bytes([OP["OP_0"]])
    if v == -1:
        return bytes([OP["OP_1NEGATE"]])
    if 1 <= v <= 16:
        return bytes([OP["OP_1"] + v - 1])
    neg, a = v < 0, abs(v)
    out = bytearray()
    while a:
        out.append(a & 0xFF)
        a >>= 8
    if out[-1] & 0x80:
        out.append(0x80 if neg else 0x00)
    elif neg:
        out[-1] |= 0x80
    return bytes([len(out)]) + bytes(out)


def push_bytes(b: bytes) -> bytes:
    assert 1 <= len(b) <= 75
    return bytes([len(b)]) + b


def ops(*names) -> bytes:
    return bytes(OP[n] for n in names)


def hash160(b: bytes) -> bytes:
    return hashlib.new("ripemd160", hashlib.sha256(b).digest()).digest()


def generate_public_key(secret_hex: str, digit_index: int) -> bytes:
    """winternitz.rs:265-281"""
    h = hash160(bytes.fromhex(secret_hex) + bytes([digit_index]))
    for _ in range(DIGITS):
        h = hash160(h)
    return h


def checksig_verify(pub_key: Sequence[bytes]) -> bytes:
    """winternitz.rs:184-263"""
    s = bytearray()
    for digit_index in range(N):
        s += push_int(DIGITS) + ops("OP_MIN", "OP_DUP", "OP_TOALTSTACK", "OP_TOALTSTACK")
        s += ops("OP_DUP", "OP_HASH160") * DIGITS
        s += ops("OP_FROMALTSTACK", "OP_PICK") + push_bytes(pub_key[N - 1 - digit_index]) + ops("OP_EQUALVERIFY")
        s += ops("OP_2DROP") * ((DIGITS + 1) // 2)
    s += ops("OP_FROMALTSTACK", "OP_DUP", "OP_NEGATE")
    s += ops("OP_FROMALTSTACK", "OP_TUCK", "OP_SUB") * (N0 - 1)
    s += push_int(DIGITS * N0) + ops("OP_ADD")
    s += ops("OP_FROMALTSTACK")
    for _ in range(N1 - 1):
        s += ops("OP_DUP", "OP_ADD") * LOG_D + ops("OP_FROMALTSTACK", "OP_ADD")
    s += ops("OP_EQUALVERIFY")
    for i in range(N0 // 2):
        s += ops("OP_SWAP") + ops("OP_DUP", "OP_ADD") * LOG_D + ops("OP_ADD")
        if i != N0 // 2 - 1:
            s += ops("OP_TOALTSTACK")
    s += ops("OP_FROMALTSTACK") * (N0 // 2 - 1)
    return bytes(s)


def u32_compress() -> bytes:
    """u32_std.rs:122-175 with OP_256MUL = 8 x (OP_DUP OP_ADD) (pseudo.rs:106-113)"""
    mul256 = ops("OP_DUP", "OP_ADD") * 8
    s = ops("OP_SWAP", "OP_ROT") + push_int(3) + ops("OP_ROLL")
    s += ops("OP_DUP") + push_int(127) + ops("OP_GREATERTHAN", "OP_IF") + push_int(128) + ops("OP_SUB") + push_int(1)
    s += ops("OP_ELSE") + push_int(0) + ops("OP_ENDIF", "OP_TOALTSTACK")
    s += (mul256 + ops("OP_ADD")) * 3
    s += ops("OP_FROMALTSTACK", "OP_IF", "OP_NEGATE", "OP_ENDIF")
    return s


def locking_script_u32(secret_hex: str) -> bytes:
    """the in-tree twin of bitcomm's locking_script_with_type(CompressType::U32): bit_comm_u32.rs:75-80"""
    pk = [generate_public_key(secret_hex, i) for i in range(N)]
    return checksig_verify(pk) + u32_compress()


def template(index_secret: str, eval_secrets: Sequence[str], limbs_per_eval: int = 1) -> Tuple[List[bytes], List[int]]:
    """The leaf script as constant segments around pushed integers (tcs/mod.rs:197-225):
         script(i) = seg[0] push(x_0) seg[1] push(x_1) ... seg[m]      x_0 = leaf index, then per evaluation its limbs REVERSED.
    Returns (segments, push_order): push_order[k] = which word of the leaf row feeds push k + 1."""
    segs = [locking_script_u32(index_secret)]
    order: List[int] = []
    eqv = ops("OP_EQUALVERIFY")
    for e, sec in enumerate(eval_secrets):
        lock = locking_script_u32(sec)
        for j in range(limbs_per_eval - 1, -1, -1):
            segs.append(eqv + (lock if j == limbs_per_eval - 1 else b""))
            order.append(e * limbs_per_eval + j)
    segs.append(eqv + ops("OP_1"))
    return segs, order


def leaf_script(segs: Sequence[bytes], order: Sequence[int], index: int, row_words: Sequence[int]) -> bytes:
    s = bytearray(segs[0]) + push_int(index)
    for k, w in enumerate(order):
        s += segs[k + 1] + push_int(int(row_words[w]))
    s += segs[len(order) + 1]
    return bytes(s)


def compact_size(n: int) -> bytes:
    if n < 0xFD:
        return bytes([n])
    if n <= 0xFFFF:
        return b"\xfd" + n.to_bytes(2, "little")
    if n <= 0xFFFFFFFF:
        return b"\xfe" + n.to_bytes(4, "little")
    return b"\xff" + n.to_bytes(8, "little")


def tagged_hash(tag: str, data: bytes) -> bytes:
    t = hashlib.sha256(tag.encode()).digest()
    return hashlib.sha256(t + t + data).digest()


def tap_leaf_hash(script: bytes) -> bytes:
    return tagged_hash("TapLeaf", b"\xc0" + compact_size(len(script)) + script)


def tap_branch_hash(a: bytes, b: bytes) -> Tuple[bytes, bool]:
    """(hash, left_first) of NodeInfo::combine_with_order"""
    left_first = a <= b
    lo, hi = (a, b) if left_first else (b, a)
    return tagged_hash("TapBranch", lo + hi), left_first


def build_tree(leaf_hashes: Sequence[bytes]) -> Tuple[bytes, List[int]]:
    """builder.rs:38-93 on leaf hashes: (root, leaf_indices) with leaf_indices[m] = taptree position of merkle leaf m."""
    n = len(leaf_hashes)
    assert n and n & (n - 1) == 0
    nodes = [(h, 1) for h in leaf_hashes]
    t_to_m = list(range(n))
    while len(nodes) > 1:
        nxt, a_start = [], 0
        for k in range(0, len(nodes), 2):
            (ha, sa), (hb, sb) = nodes[k], nodes[k + 1]
            h, left_first = tap_branch_hash(ha, hb)
            nxt.append((h, sa + sb))
            if not left_first:
                t_to_m[a_start: a_start + sa + sb] = t_to_m[a_start + sa: a_start + sa + sb] + t_to_m[a_start: a_start + sa]
            a_start += sa + sb
        nodes = nxt
    rev = [0] * n
    for t, m in enumerate(t_to_m):
        rev[m] = t
    return nodes[0][0], rev


def commit(segs: Sequence[bytes], order: Sequence[int], rows) -> Tuple[bytes, List[int], List[bytes]]:
    """rows: [n_leaves][words] canonical u32.  Returns (root, leaf_indices, leaf_hashes)."""
    hashes = [tap_leaf_hash(leaf_script(segs, order, i, r)) for i, r in enumerate(rows)]
    root, perm = build_tree(hashes)
    return root, perm, hashes
