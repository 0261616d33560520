"""Wire format of the opening proof returned by ts_pcs_open (include/tapstark.h): postcard (serde) of the pair
`(OpenedValues, FriProof)` -- fri/src/proof.rs:13-33, fri/src/two_adic_pcs.rs:325-386,408-411 -- and of
`uni_stark::Proof` (uni-stark/src/proof.rs:19-37).

postcard: unsigned integers and lengths are LEB128 varints; Vec<T> = length + elements; structs, tuples and fixed arrays are
their fields in order; u8 is one raw byte.  BabyBear is its canonical u32 ([MEM] p3-baby-bear Serialize: parity unpinned, the
reference holds no serialized vector), BabyBear^4 four of them (low coefficient first), a digest 32 raw bytes.  The Merkle
sibling path stands where the reference has its Taproot `CommitedProof`.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


class Reader:
    def __init__(self, data: bytes):
        self.b, self.o = memoryview(data), 0

    def varint(self) -> int:
        v, s = 0, 0
        while True:
            x = self.b[self.o]
            self.o += 1
            v |= (x & 0x7F) << s
            if x < 0x80:
                return v
            s += 7

    def ef(self) -> np.ndarray:
        return np.array([self.varint() for _ in range(4)], dtype=np.uint32)

    def digest(self) -> bytes:
        d = bytes(self.b[self.o:self.o + 32])
        self.o += 32
        return d

    def path(self) -> np.ndarray:
        n = self.varint()
        p = np.frombuffer(bytes(self.b[self.o:self.o + 32 * n]), dtype=np.uint8).reshape(n, 32).copy()
        self.o += 32 * n
        return p

    def done(self) -> bool:
        return self.o == len(self.b)


def decode_opening(data: bytes, want_split: bool = False):
    """bytes of ts_pcs_open -> (opened_values[round][matrix][point] -> (width, 4) array, FriProof).
    want_split: also return the byte offset at which the FriProof starts (uni_stark::Proof embeds it as opening_proof)."""
    from . import BatchOpening, BfQueryProof, FriProof

    r = Reader(data)
    opened = []
    for _ in range(r.varint()):
        rnd = []
        for _ in range(r.varint()):
            mat = []
            for _ in range(r.varint()):
                w = r.varint()
                mat.append(np.stack([r.ef() for _ in range(w)]) if w else np.zeros((0, 4), dtype=np.uint32))
            rnd.append(mat)
        opened.append(rnd)
    split = r.o
    proof = decode_fri_proof(r)
    if not r.done():
        raise ValueError("trailing bytes after the opening proof")
    return (opened, proof, split) if want_split else (opened, proof)


def decode_fri_proof(r: Reader):
    from . import BatchOpening, BfQueryProof, FriProof

    commits = [r.digest() for _ in range(r.varint())]
    queries = []
    for _ in range(r.varint()):
        input_proof = []
        for _ in range(r.varint()):
            vals = []
            for _ in range(r.varint()):
                vals.append(np.array([r.varint() for _ in range(r.varint())], dtype=np.uint32))
            input_proof.append(BatchOpening(vals, r.path()))
        steps = []
        for _ in range(r.varint()):
            rows = []
            for _ in range(r.varint()):
                n = r.varint()
                rows.append(np.concatenate([r.ef() for _ in range(n)]) if n else np.zeros(0, dtype=np.uint32))
            steps.append((rows, r.path()))
        queries.append(BfQueryProof(input_proof, steps))
    final_poly = r.ef()
    pow_witness = r.varint()
    return FriProof(commits, queries, final_poly, pow_witness)


def varint(v: int) -> bytes:
    out = bytearray()
    v = int(v)
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _ef_vec(a) -> bytes:
    a = np.asarray(a, dtype=np.uint32).reshape(-1, 4)
    return varint(a.shape[0]) + b"".join(varint(int(x)) for x in a.reshape(-1))


def encode_stark_proof(commit_trace: bytes, commit_quotient: bytes, trace_local, trace_next, quotient_chunks,
                       fri_proof_bytes: bytes, degree_bits: int) -> bytes:
    """postcard of uni_stark::Proof (uni-stark/src/proof.rs:19-37): commitments { trace, quotient_chunks }, opened_values
    { trace_local, trace_next, quotient_chunks }, opening_proof (the FriProof bytes of ts_pcs_open, verbatim), degree_bits."""
    out = bytes(commit_trace) + bytes(commit_quotient)
    out += _ef_vec(trace_local) + _ef_vec(trace_next)
    out += varint(len(quotient_chunks)) + b"".join(_ef_vec(c) for c in quotient_chunks)
    return out + fri_proof_bytes + varint(degree_bits)
