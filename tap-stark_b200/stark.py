"""Host-side mirror of uni-stark's prover (uni-stark/src/prover.rs:26-118) over the device path: trace commit ->
quotient values on the device (ts_quotient_values) -> quotient commit -> open at zeta / zeta*g -> FRI proof.

An AIR is any object with `width()` and `eval(builder)` written against the p3 `AirBuilder` surface the reference's
AIRs use (uni-stark/tests/fib_air.rs:29-58): `main()` -> (local, next) row variables, `public_values()`,
`when_first_row()`, `when_transition()`, `when_last_row()`, `when(cond)`, `assert_zero`, `assert_eq`.
`SymbolicAirBuilder` records the constraints as expressions (uni-stark/src/symbolic_builder.rs), from which come
the constraint degree (symbolic_expression.rs:41-61) and the constraint program the CUDA kernel interprets
(csrc/quotient.cuh).  The verifier lives in oracle/stark.py (test infrastructure): it evaluates the same AIR
directly in the extension field, never through this module's compiler.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

from . import (P, GENERATOR, DeviceMatrix, TapStarkError, TwoAdicFriPcs, TwoAdicMultiplicativeCoset, BfChallenger, FriProof,
               to_monty, _ptr)

K_REG, K_LOCAL, K_NEXT, K_PUBLIC, K_CONST, K_SEL = range(6)
OP_ADD, OP_SUB, OP_MUL, OP_NEG, OP_ASSERT_ZERO = range(5)
MAX_REGS = 64
_ROOT27 = 0x1A427A41


def two_adic_generator(bits: int) -> int:
    return pow(_ROOT27, 1 << (27 - bits), P)


class Expr:
    """SymbolicExpression (uni-stark/src/symbolic_expression.rs:11-37) with its degree multiple (:41-61)."""

    __slots__ = ("kind", "index", "x", "y", "degree")

    def __init__(self, kind, index=0, x=None, y=None, degree=0):
        self.kind, self.index, self.x, self.y, self.degree = kind, index, x, y, degree

    @staticmethod
    def of(v) -> "Expr":
        if isinstance(v, Expr):
            return v
        return Expr("const", int(v) % P)

    def __add__(self, o):
        o = Expr.of(o)
        return Expr("add", x=self, y=o, degree=max(self.degree, o.degree))

    __radd__ = __add__

    def __sub__(self, o):
        o = Expr.of(o)
        return Expr("sub", x=self, y=o, degree=max(self.degree, o.degree))

    def __rsub__(self, o):
        return Expr.of(o) - self

    def __mul__(self, o):
        o = Expr.of(o)
        return Expr("mul", x=self, y=o, degree=self.degree + o.degree)

    __rmul__ = __mul__

    def __neg__(self):
        return Expr("neg", x=self, degree=self.degree)


class _Filtered:
    """p3_air::FilteredAirBuilder: assert_zero(x) -> inner.assert_zero(condition * x)."""

    def __init__(self, inner, condition):
        self.inner, self.condition = inner, condition

    def assert_zero(self, x):
        self.inner.assert_zero(self.condition * Expr.of(x))

    def assert_eq(self, x, y):
        self.assert_zero(Expr.of(x) - Expr.of(y))

    def assert_one(self, x):
        self.assert_zero(Expr.of(x) - 1)

    def when(self, condition):
        return _Filtered(self, condition)


class SymbolicAirBuilder:
    """uni-stark/src/symbolic_builder.rs:57-150: collects the constraints of `air.eval` in order."""

    def __init__(self, width: int, num_public_values: int):
        self.local = [Expr("local", c, degree=1) for c in range(width)]
        self.next = [Expr("next", c, degree=1) for c in range(width)]
        self.publics = [Expr("public", k, degree=0) for k in range(num_public_values)]
        self.constraints: List[Expr] = []

    def main(self):
        return self.local, self.next

    def public_values(self):
        return self.publics

    def is_first_row(self):
        return Expr("sel", 0, degree=1)

    def is_last_row(self):
        return Expr("sel", 1, degree=1)

    def is_transition(self):
        return Expr("sel", 2, degree=0)

    def when(self, condition):
        return _Filtered(self, condition)

    def when_first_row(self):
        return self.when(self.is_first_row())

    def when_last_row(self):
        return self.when(self.is_last_row())

    def when_transition(self):
        return self.when(self.is_transition())

    def assert_zero(self, x):
        self.constraints.append(Expr.of(x))

    def assert_eq(self, x, y):
        self.assert_zero(Expr.of(x) - Expr.of(y))

    def assert_one(self, x):
        self.assert_zero(Expr.of(x) - 1)


def get_symbolic_constraints(air, num_public_values: int) -> List[Expr]:
    b = SymbolicAirBuilder(air.width(), num_public_values)
    air.eval(b)
    return b.constraints


def get_log_quotient_degree(air, num_public_values: int) -> int:
    """uni-stark/src/symbolic_builder.rs:15-32: pad the constraint degree to >= 2; log2_ceil(degree - 1)."""
    d = max([c.degree for c in get_symbolic_constraints(air, num_public_values)] + [2])
    return (d - 2).bit_length()  # log2_ceil(d - 1)


def compile_program(constraints: Sequence[Expr]):
    """Expression trees -> the three-address program of csrc/quotient.cuh.  Returns (program (n, 4) u32, constants)."""
    prog, consts, const_ix, free = [], [], {}, list(range(MAX_REGS - 1, -1, -1))

    def leaf(e):
        if e.kind == "local":
            return K_LOCAL << 28 | e.index
        if e.kind == "next":
            return K_NEXT << 28 | e.index
        if e.kind == "public":
            return K_PUBLIC << 28 | e.index
        if e.kind == "sel":
            return K_SEL << 28 | e.index
        if e.kind == "const":
            if e.index not in const_ix:
                const_ix[e.index] = len(consts)
                consts.append(e.index)
            return K_CONST << 28 | const_ix[e.index]
        return None

    def release(code):
        if code >> 28 == K_REG:
            free.append(code & 0x0FFFFFFF)

    def emit(e) -> int:
        code = leaf(e)
        if code is not None:
            return code
        a = emit(e.x)
        b = emit(e.y) if e.y is not None else 0
        release(a)
        if e.y is not None:
            release(b)
        if not free:
            raise TapStarkError(f"constraint program needs more than {MAX_REGS} registers")
        dst = free.pop()
        prog.append(({"add": OP_ADD, "sub": OP_SUB, "mul": OP_MUL, "neg": OP_NEG}[e.kind], dst, a, b))
        return K_REG << 28 | dst

    for c in constraints:
        code = emit(c)
        prog.append((OP_ASSERT_ZERO, 0, code, 0))
        release(code)
    return np.array(prog, dtype=np.uint32).reshape(-1, 4), np.array(consts, dtype=np.uint32)


@dataclass
class Commitments:  # uni-stark/src/proof.rs:27-31
    trace: bytes
    quotient_chunks: bytes


@dataclass
class OpenedValues:  # uni-stark/src/proof.rs:33-37
    trace_local: np.ndarray      # (width, 4) canonical
    trace_next: np.ndarray
    quotient_chunks: List[np.ndarray]  # per chunk (4, 4)


@dataclass
class Proof:  # uni-stark/src/proof.rs:19-25
    commitments: Commitments
    opened_values: OpenedValues
    opening_proof: FriProof
    degree_bits: int
    opening_proof_bytes: bytes = b""  # the FriProof exactly as ts_pcs_open serialised it

    def to_bytes(self) -> bytes:
        """postcard of the whole proof (the reference's one serialisation call is postcard::to_allocvec(&proof),
        commented out at uni-stark/tests/mul_air.rs:133)."""
        from . import proofio

        return proofio.encode_stark_proof(self.commitments.trace, self.commitments.quotient_chunks, self.opened_values.trace_local,
                                          self.opened_values.trace_next, self.opened_values.quotient_chunks,
                                          self.opening_proof_bytes, self.degree_bits)


def quotient_values(pcs: TwoAdicFriPcs, trace_data, air, public_values: Sequence[int], log_n: int, log_quotient_degree: int,
                    alpha) -> List[DeviceMatrix]:
    """prover.rs:69-80 on the device: the quotient chunks (each n x 4, natural order on its coset)."""
    ctx, L = pcs.ctx, pcs.ctx._L
    prog, consts = compile_program(get_symbolic_constraints(air, len(public_values)))
    prog = np.ascontiguousarray(prog)
    cm = to_monty(consts) if consts.size else np.zeros(1, dtype=np.uint32)
    pm = to_monty(np.array([int(v) % P for v in public_values], dtype=np.uint32)) if len(public_values) else np.zeros(1, dtype=np.uint32)
    am = to_monty(np.asarray(alpha, dtype=np.uint32))
    lde = pcs.mmcs.get_matrices(trace_data)[0]
    qd = 1 << log_quotient_degree
    outs = (C.c_void_p * qd)()
    ctx.check(L.ts_quotient_values(ctx._h, lde._h, log_n, log_quotient_degree, _ptr(prog), prog.shape[0], _ptr(cm), consts.size,
                                   _ptr(pm), len(public_values), _ptr(am), outs), "quotient_values")
    return [DeviceMatrix(ctx, C.c_void_p(outs[k])) for k in range(qd)]


def prove(pcs: TwoAdicFriPcs, air, challenger: BfChallenger, trace: np.ndarray, public_values: Sequence[int]) -> Proof:
    """uni_stark::prove (uni-stark/src/prover.rs:26-118).  trace: canonical (n, width) array."""
    degree = trace.shape[0]
    log_degree = degree.bit_length() - 1
    if 1 << log_degree != degree or trace.shape[1] != air.width():
        raise TapStarkError("prove: the trace must be 2^k x air.width()")
    log_qd = get_log_quotient_degree(air, len(public_values))
    qd = 1 << log_qd
    trace_domain = pcs.natural_domain_for_degree(degree)
    trace_commit, trace_data = pcs.commit([(trace_domain, DeviceMatrix.from_canonical(pcs.ctx, trace))])  # :51-52
    challenger.observe(trace_commit)  # :59
    alpha = [int(x) for x in challenger.sample()]  # :62
    chunks = quotient_values(pcs, trace_data, air, public_values, log_degree, log_qd, alpha)  # :64-80
    w_m = two_adic_generator(log_degree + log_qd)
    qc_domains = [TwoAdicMultiplicativeCoset(log_degree, GENERATOR * pow(w_m, k, P) % P) for k in range(qd)]  # split_domains
    quotient_commit, quotient_data = pcs.commit(list(zip(qc_domains, chunks)))  # :83-84
    challenger.observe(quotient_commit)  # :85
    zeta = [int(x) for x in challenger.sample()]  # :92
    g_n = two_adic_generator(log_degree)
    zeta_next = [c * g_n % P for c in zeta]  # trace_domain.next_point(zeta)
    # :95-105 -- Pcs::open through ONE C-ABI call (ts_pcs_open); the bytes decode into the proof objects
    from . import proofio

    blob = pcs.open_bytes([(trace_data, [[zeta, zeta_next]]), (quotient_data, [[zeta]] * qd)], challenger)
    opened, opening_proof, split = proofio.decode_opening(blob, want_split=True)
    ov = OpenedValues(opened[0][0][0], opened[0][0][1], [opened[1][k][0] for k in range(qd)])
    return Proof(Commitments(trace_commit, quotient_commit), ov, opening_proof, log_degree, blob[split:])
