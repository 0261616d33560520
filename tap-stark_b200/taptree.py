"""Host mirror of `TapTreeMmcs` (basic/src/mmcs/taptree_mmcs.rs:23-122) and of the taptree commitment scheme under it
(`TCS`, basic/src/tcs/mod.rs:227-330) over the TapTree kernels of the C ABI.

What runs where, as in the reference:
    host    the bit-commitment assignment of one tree (Winternitz public keys: hash160 chains over a secret -- key generation,
            not per-leaf work; tcs/mod.rs:252-262) and the script TEMPLATE built from them; the single leaf script and the
            inclusion check of an opening
    device  everything proportional to the number of leaves: the padded leaf rows (padding_matrix, tcs/mod.rs:341-383), one
            TapLeaf hash per leaf, the sorted-pair TapBranch tree, the leaf permutation, the Merkle branch of an opening

    TapTreeMmcs.commit        taptree_mmcs.rs:101-114 -> tcs/mod.rs:284-292: `num_queries` trees over the SAME matrices, each with
                              freshly assigned bit-commitments
    TapTreeMmcs.open_batch    taptree_mmcs.rs:46-74   -> tcs/mod.rs:294-301, :141-146 (query_proof)
    TapTreeMmcs.verify_batch  taptree_mmcs.rs:76-99   -> tcs/mod.rs:423-434 + :148-152 (verify_proof): inclusion under the root AND
                              the leaf script accepting the opened values.  The reference EXECUTES the script with the Winternitz
                              witness of the values; here the script is rebuilt from the proof's bit-commitments, the query index
                              and the opened values and compared byte for byte with the committed leaf script -- the same
                              statement (every OP_EQUALVERIFY of tcs/mod.rs:203-223 compares a recovered value with a pushed
                              constant) without a Bitcoin script interpreter.

PARITY UNPINNED for the locking-script bytes and the secret generator: they live in the external `bitcomm` crate (git dependency,
absent from the reference tree); the bytes below follow the in-tree twin scripts/src/bit_comm/winternitz.rs:170-281,
bit_comm_u32.rs:75-80 and u32/u32_std.rs:122-175.  The secret generator is a stand-in: a counter hashed with a seed.
"""
from __future__ import annotations

import ctypes as C
import hashlib
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import Context, DeviceMatrix, TapStarkError, TapTreeCommit

# winternitz.rs:20-33: 4-bit digits, a u32 message = 8 digits + 2 checksum digits
LOG_D = 4
DIGITS = (1 << LOG_D) - 1
N0, N1 = 8, 2
N = N0 + N1

_OP = dict(OP_0=0x00, OP_1NEGATE=0x4F, OP_1=0x51, OP_IF=0x63, OP_ELSE=0x67, OP_ENDIF=0x68, OP_TOALTSTACK=0x6B, OP_FROMALTSTACK=0x6C,
           OP_2DROP=0x6D, OP_DUP=0x76, OP_PICK=0x79, OP_ROLL=0x7A, OP_ROT=0x7B, OP_SWAP=0x7C, OP_TUCK=0x7D, OP_EQUALVERIFY=0x88,
           OP_NEGATE=0x8F, OP_ADD=0x93, OP_SUB=0x94, OP_MIN=0xA3, OP_GREATERTHAN=0xA0, OP_HASH160=0xA9)


def _ops(*names: str) -> bytes:
    return bytes(_OP[n] for n in names)


def script_num_push(v: int) -> bytes:
    """Minimal script-number push, what `{ n }` inside `script!` emits (and what taptree_leaf_kernel streams per pushed word)."""
    if v == 0:
        return b"\x00"
    if v == -1:
        return b"\x4f"
    if 1 <= v <= 16:
        return bytes([0x50 + v])
    neg, a, out = v < 0, abs(v), bytearray()
    while a:
        out.append(a & 0xFF)
        a >>= 8
    if out[-1] & 0x80:
        out.append(0x80 if neg else 0x00)
    elif neg:
        out[-1] |= 0x80
    return bytes([len(out)]) + bytes(out)


def hash160(b: bytes) -> bytes:
    return hashlib.new("ripemd160", hashlib.sha256(b).digest()).digest()


@dataclass(frozen=True)
class BitCommitment:
    """One u32 Winternitz bit-commitment: its N = 10 public keys (winternitz.rs:265-281).  The secret stays with the prover."""

    public_keys: Tuple[bytes, ...]
    secret: Optional[bytes] = field(default=None, compare=False, repr=False)

    @staticmethod
    def from_secret(secret: bytes) -> "BitCommitment":
        keys = []
        for digit_index in range(N):
            h = hash160(secret + bytes([digit_index]))
            for _ in range(DIGITS):
                h = hash160(h)
            keys.append(h)
        return BitCommitment(tuple(keys), secret)

    def locking_script(self) -> bytes:
        """checksig_verify (winternitz.rs:184-263) followed by u32_compress (u32_std.rs:122-175): recovers the committed u32."""
        s = bytearray()
        for digit_index in range(N):
            s += script_num_push(DIGITS) + _ops("OP_MIN", "OP_DUP", "OP_TOALTSTACK", "OP_TOALTSTACK")
            s += _ops("OP_DUP", "OP_HASH160") * DIGITS
            key = self.public_keys[N - 1 - digit_index]
            s += _ops("OP_FROMALTSTACK", "OP_PICK") + bytes([len(key)]) + key + _ops("OP_EQUALVERIFY")
            s += _ops("OP_2DROP") * ((DIGITS + 1) // 2)
        s += _ops("OP_FROMALTSTACK", "OP_DUP", "OP_NEGATE") + _ops("OP_FROMALTSTACK", "OP_TUCK", "OP_SUB") * (N0 - 1)
        s += script_num_push(DIGITS * N0) + _ops("OP_ADD", "OP_FROMALTSTACK")
        for _ in range(N1 - 1):
            s += _ops("OP_DUP", "OP_ADD") * LOG_D + _ops("OP_FROMALTSTACK", "OP_ADD")
        s += _ops("OP_EQUALVERIFY")
        for i in range(N0 // 2):
            s += _ops("OP_SWAP") + _ops("OP_DUP", "OP_ADD") * LOG_D + _ops("OP_ADD")
            if i != N0 // 2 - 1:
                s += _ops("OP_TOALTSTACK")
        s += _ops("OP_FROMALTSTACK") * (N0 // 2 - 1)
        mul256 = _ops("OP_DUP", "OP_ADD") * 8  # pseudo.rs:106-113
        s += _ops("OP_SWAP", "OP_ROT") + script_num_push(3) + _ops("OP_ROLL", "OP_DUP") + script_num_push(127)
        s += _ops("OP_GREATERTHAN", "OP_IF") + script_num_push(128) + _ops("OP_SUB") + script_num_push(1) + _ops("OP_ELSE")
        s += script_num_push(0) + _ops("OP_ENDIF", "OP_TOALTSTACK") + (mul256 + _ops("OP_ADD")) * 3
        s += _ops("OP_FROMALTSTACK", "OP_IF", "OP_NEGATE", "OP_ENDIF")
        return bytes(s)


class SecretGen:
    """Stand-in for bitcomm's SecretGenIns (external): secret k = SHA-256(seed || k)[:20]."""

    def __init__(self, seed: bytes = b"tapstark-b200"):
        self.seed, self.counter = seed, 0

    def next(self) -> bytes:
        self.counter += 1
        return hashlib.sha256(self.seed + self.counter.to_bytes(8, "little")).digest()[:20]


@dataclass
class UseBComm:
    """tcs/mod.rs:164-169: the bit-commitments of one tree -- one for the leaf index, one per u32 limb group of an evaluation."""

    index_bc: BitCommitment
    evaluations_bc: List[BitCommitment]
    limbs: int  # u32 limbs per evaluation: F::U32_SIZE, 1 (BabyBear) or 4 (BabyBear^4), tcs/mod.rs:239-245

    def template(self) -> Tuple[List[bytes], List[int]]:
        """generate_script (tcs/mod.rs:197-225) as constant segments around pushed integers:
        script(i) = seg[0] P(i) seg[1] P(x_1) ... seg[m]; push k + 1 takes word push_word[k] of the leaf row (limbs reversed, :214)."""
        eqv = _ops("OP_EQUALVERIFY")
        segs, order = [self.index_bc.locking_script()], []
        for e, bc in enumerate(self.evaluations_bc):
            lock = bc.locking_script()
            for j in range(self.limbs - 1, -1, -1):
                segs.append(eqv + (lock if j == self.limbs - 1 else b""))
                order.append(e * self.limbs + j)
        segs.append(eqv + _ops("OP_1"))
        return segs, order

    def leaf_script(self, index: int, row_words: Sequence[int]) -> bytes:
        segs, order = self.template()
        s = bytearray(segs[0]) + script_num_push(index)
        for k, w in enumerate(order):
            s += segs[k + 1] + script_num_push(int(row_words[w]))
        return bytes(s + segs[-1])


def _tagged(tag: bytes, data: bytes) -> bytes:
    t = hashlib.sha256(tag).digest()
    return hashlib.sha256(t + t + data).digest()


def _compact_size(n: int) -> bytes:
    if n < 0xFD:
        return bytes([n])
    if n <= 0xFFFF:
        return b"\xfd" + n.to_bytes(2, "little")
    return b"\xfe" + n.to_bytes(4, "little")


def tap_leaf_hash(script: bytes) -> bytes:
    """LeafNode::node_hash of `NodeInfo::new_leaf_with_ver(script, TapScript)` (builder.rs:26)."""
    return _tagged(b"TapLeaf", b"\xc0" + _compact_size(len(script)) + script)


def verify_inclusion(root: bytes, leaf_hash: bytes, branch: Sequence[bytes]) -> bool:
    """complete_taptree.rs:64-73: TapNodeHash::from_node_hashes sorts each pair."""
    h = leaf_hash
    for sib in branch:
        lo, hi = (h, sib) if h <= sib else (sib, h)
        h = _tagged(b"TapBranch", lo + hi)
    return h == root


@dataclass
class CommitedProof:
    """tcs/mod.rs:106-112: the opened TapLeaf (script + TaprootMerkleBranch), the tree's bit-commitments, the query index."""

    leaf_script: bytes
    merkle_branch: List[bytes]
    use_bcs: UseBComm
    query_index: int


@dataclass
class CommitedData:
    """tcs/mod.rs:85-91.  `leaves` are the committed matrices (device resident), `rows` the padded leaf rows."""

    leaves: List[DeviceMatrix]
    rows: DeviceMatrix
    tree: TapTreeCommit
    use_bcs: UseBComm

    def get_max_height(self) -> int:
        return max(m.rows for m in self.leaves)


class TapTreeMmcs:
    """BFMmcs over taptrees (taptree_mmcs.rs:41-122).  Commitment = one 32-byte TapNodeHash per query; matrices hold BabyBear
    (`limbs` = 1) or BabyBear^4 stored as 4 consecutive words (`limbs` = 4)."""

    def __init__(self, ctx: Context, num_queries: int, limbs: int = 1, secret_gen: Optional[SecretGen] = None):
        if limbs not in (1, 4):
            raise TapStarkError("only support 1 or 4")  # tcs/mod.rs:243
        self.ctx, self.num_queries, self.limbs = ctx, num_queries, limbs
        self.secret_gen = secret_gen or SecretGen()

    # ---- PolyTCS::commit_polys (tcs/mod.rs:238-282) on the device
    def _commit_polys(self, leaves: List[DeviceMatrix], rows: DeviceMatrix) -> CommitedData:
        if rows.width % self.limbs:
            raise TapStarkError("taptree: row width is not a multiple of the limb count")
        n_eval = rows.width // self.limbs
        use_bcs = UseBComm(BitCommitment.from_secret(self.secret_gen.next()),
                           [BitCommitment.from_secret(self.secret_gen.next()) for _ in range(n_eval)], self.limbs)
        segs, order = use_bcs.template()
        return CommitedData(leaves, rows, TapTreeCommit(self.ctx, rows, segs, order), use_bcs)

    def padded_rows(self, inputs: Sequence[DeviceMatrix]) -> DeviceMatrix:
        L = self.ctx._L
        arr = (C.c_void_p * len(inputs))(*[m._h for m in inputs])
        h = C.c_void_p()
        self.ctx.check(L.ts_padded_leaf_rows(self.ctx._h, arr, len(inputs), C.byref(h)), "padded_leaf_rows")
        return DeviceMatrix(self.ctx, h)

    def commit(self, inputs: Sequence[DeviceMatrix]) -> Tuple[List[bytes], List[CommitedData]]:
        inputs = list(inputs)
        rows = self.padded_rows(inputs)  # identical for every query: built once (the reference clones the matrices, :290)
        data = [self._commit_polys(inputs, rows) for _ in range(self.num_queries)]
        return [d.tree.root for d in data], data

    def get_matrices(self, prover_data: List[CommitedData]) -> List[DeviceMatrix]:
        return prover_data[0].leaves

    def open_batch(self, query_times_index: int, query_index: int, prover_data: List[CommitedData]):
        if len(prover_data) != self.num_queries:
            raise TapStarkError("open_batch: prover data of another query count")  # tcs/mod.rs:413
        d = prover_data[query_times_index]
        max_height = d.get_max_height()
        if not 0 <= query_index < max_height:
            raise TapStarkError("open_batch: index out of range")
        log_max = max_height.bit_length() - 1
        openings = []
        for m in prover_data[0].leaves:  # taptree_mmcs.rs:54-64
            reduced = query_index >> (log_max - (m.rows.bit_length() - 1))
            openings.append(m.to_canonical(reduced, 1)[0])
        row = d.rows.to_canonical(query_index, 1)[0]  # padding_matrix(...)[index], tcs/mod.rs:299
        if not np.array_equal(self._flatten_like_rows(prover_data[0].leaves, openings), row):
            raise TapStarkError("open_batch: padded row and matrix rows disagree")  # the reference's assert_eq!, :71
        branch, _ = d.tree.open(query_index)
        proof = CommitedProof(d.use_bcs.leaf_script(query_index, row), branch, d.use_bcs, query_index)
        return openings, proof

    @staticmethod
    def _flatten_like_rows(leaves: Sequence[DeviceMatrix], openings: Sequence[np.ndarray]) -> np.ndarray:
        """opened rows in padding_matrix order (tallest matrix first, stable)"""
        order = sorted(range(len(leaves)), key=lambda i: -leaves[i].rows)
        return np.concatenate([openings[i] for i in order])

    def verify_batch(self, query_times_index: int, opened_values: Sequence[np.ndarray], proof: CommitedProof, roots: Sequence[bytes],
                     heights: Optional[Sequence[int]] = None) -> None:
        """Raises TapStarkError("InvalidOpenedValue") like BfError::InvalidOpenedValue (taptree_mmcs.rs:93-98).  `heights` (of the
        committed matrices, in commit order) gives the padded order of `opened_values`; omitted = they are already in it."""
        vals = [np.asarray(v, dtype=np.uint32).reshape(-1) for v in opened_values]
        if heights is not None:
            vals = [vals[i] for i in sorted(range(len(vals)), key=lambda i: -heights[i])]
        flat = np.concatenate(vals) if vals else np.zeros(0, dtype=np.uint32)
        ok = len(flat) == len(proof.use_bcs.evaluations_bc) * proof.use_bcs.limbs
        ok = ok and proof.use_bcs.leaf_script(proof.query_index, flat) == proof.leaf_script
        ok = ok and verify_inclusion(roots[query_times_index], tap_leaf_hash(proof.leaf_script), proof.merkle_branch)
        if not ok:
            raise TapStarkError("InvalidOpenedValue")
