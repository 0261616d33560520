"""TEST INFRASTRUCTURE (oracle/): an independent postcard encoder of the proof objects, used to compare the bytes that
ts_pcs_open produces with a restatement of the reference's serde derive order.

    FriProof / BfQueryProof        fri/src/proof.rs:13-33 (field order = declaration order, as #[derive(Serialize)] emits)
    BatchOpening                   [MEM] p3-fri { opened_values, opening_proof }, used at fri/src/two_adic_pcs.rs:408-411
    OpenedValues of Pcs::open      fri/src/two_adic_pcs.rs:325-386: Vec<Vec<Vec<Vec<Challenge>>>>
    uni_stark::Proof               uni-stark/src/proof.rs:19-37

postcard: varint (LEB128) unsigned integers and lengths; Vec = length + items; struct/tuple/array = items; u8 raw.
BabyBear = canonical u32 varint [MEM]; parity unpinned (the reference holds no serialized bytes; its one serialisation call
is the commented-out postcard::to_allocvec at uni-stark/tests/mul_air.rs:133).
"""
from __future__ import annotations


def varint(v: int) -> bytes:
    out = bytearray()
    v = int(v)
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _vec(items, enc) -> bytes:
    items = list(items)
    return varint(len(items)) + b"".join(enc(x) for x in items)


def _ef(e) -> bytes:
    e = [int(x) for x in e]
    assert len(e) == 4
    return b"".join(varint(x) for x in e)


def _ef_vec(flat) -> bytes:
    """Vec<Challenge> from a (k, 4) array or a flat array of 4k canonical u32"""
    import numpy as np

    a = np.asarray(flat, dtype=np.uint32).reshape(-1, 4)
    return _vec(a, _ef)


def _path(p) -> bytes:
    import numpy as np

    p = np.asarray(p, dtype=np.uint8).reshape(-1, 32)
    return varint(p.shape[0]) + p.tobytes()


def encode_opened_values(opened) -> bytes:
    return _vec(opened, lambda rnd: _vec(rnd, lambda mat: _vec(mat, _ef_vec)))


def encode_fri_proof(proof) -> bytes:
    out = _vec(proof.commit_phase_commits, lambda c: bytes(c))
    def query(q):
        ip = _vec(q.input_proof, lambda bo: _vec(bo.opened_values, lambda row: _vec(row, varint)) + _path(bo.opening_proof))
        steps = _vec(q.commit_phase_openings, lambda st: _vec(st[0], _ef_vec) + _path(st[1]))
        return ip + steps
    out += _vec(proof.query_proofs, query)
    out += _ef(proof.final_poly)
    out += varint(proof.pow_witness)
    return out


def encode_opening(opened, proof) -> bytes:
    return encode_opened_values(opened) + encode_fri_proof(proof)


def encode_stark_proof(proof) -> bytes:
    """uni-stark/src/proof.rs:19-37, field by field in declaration order."""
    out = bytes(proof.commitments.trace) + bytes(proof.commitments.quotient_chunks)
    ov = proof.opened_values
    out += _ef_vec(ov.trace_local) + _ef_vec(ov.trace_next) + _vec(ov.quotient_chunks, _ef_vec)
    out += encode_fri_proof(proof.opening_proof)
    return out + varint(proof.degree_bits)
