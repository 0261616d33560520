"""Oracle restatement of the uni-stark layer around the PCS: quotient_values (uni-stark/src/prover.rs:122-194) and
uni_stark::verify (uni-stark/src/verifier.rs:20-163).  TEST INFRASTRUCTURE ONLY (see oracle/tapstark_oracle.h).

The AIR's `eval` is run DIRECTLY against folders that compute with field elements (uni-stark/src/folder.rs): base
field per quotient-domain row for the prover reference, extension field at zeta for the verifier.  Nothing here
goes through the product's expression compiler or constraint program.

Parity note: the selector normalisation (is_first_row = Z_H(x)/(x-1), is_last_row = Z_H(x)/(x - w_n^-1),
is_transition = x - w_n^-1, inv_zeroifier = 1/Z_H(x)) is [MEM] p3-commit @72b2fc16 `selectors_on_coset` /
`selectors_at_point`: the reference has no known-answer vector for quotient values, so this layer is "parity
unpinned" against the reference binary and pinned by prove -> verify agreement, like the reference's own test
(uni-stark/tests/fib_air.rs:117-149).
"""
from __future__ import annotations

import numpy as np

from . import oracle as orc
from . import pyref
from . import verifier as V

P = pyref.P


class Val:
    """A field element for the folders: base (int) or extension (4-list); mixed arithmetic promotes."""

    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v if isinstance(v, list) else int(v) % P

    @staticmethod
    def of(x):
        return x if isinstance(x, Val) else Val(x)

    def ext(self):
        return self.v if isinstance(self.v, list) else [self.v, 0, 0, 0]

    def _bin(self, o, fb, fe):
        o = Val.of(o)
        if isinstance(self.v, int) and isinstance(o.v, int):
            return Val(fb(self.v, o.v))
        return Val(fe(self.ext(), o.ext()))

    def __add__(self, o):
        return self._bin(o, lambda a, b: (a + b) % P, pyref.ef_add)

    __radd__ = __add__

    def __sub__(self, o):
        return self._bin(o, lambda a, b: (a - b) % P, pyref.ef_sub)

    def __rsub__(self, o):
        return Val.of(o) - self

    def __mul__(self, o):
        return self._bin(o, lambda a, b: a * b % P, pyref.ef_mul)

    __rmul__ = __mul__

    def __neg__(self):
        return Val(0) - self


class _Filtered:
    def __init__(self, inner, condition):
        self.inner, self.condition = inner, condition

    def assert_zero(self, x):
        self.inner.assert_zero(self.condition * Val.of(x))

    def assert_eq(self, x, y):
        self.assert_zero(Val.of(x) - Val.of(y))

    def assert_one(self, x):
        self.assert_zero(Val.of(x) - 1)

    def when(self, condition):
        return _Filtered(self, condition)


class ConstraintFolder:
    """ProverConstraintFolder / VerifierConstraintFolder (uni-stark/src/folder.rs): accumulator = accumulator * alpha + x."""

    def __init__(self, local, nxt, public_values, is_first_row, is_last_row, is_transition, alpha):
        self.local, self.next = [Val.of(x) for x in local], [Val.of(x) for x in nxt]
        self.publics = [Val(int(v)) for v in public_values]
        self.sels = (Val.of(is_first_row), Val.of(is_last_row), Val.of(is_transition))
        self.alpha = Val([int(x) for x in alpha])
        self.accumulator = Val([0, 0, 0, 0])

    def main(self):
        return self.local, self.next

    def public_values(self):
        return self.publics

    def is_first_row(self):
        return self.sels[0]

    def is_last_row(self):
        return self.sels[1]

    def is_transition(self):
        return self.sels[2]

    def when(self, condition):
        return _Filtered(self, condition)

    def when_first_row(self):
        return self.when(self.sels[0])

    def when_last_row(self):
        return self.when(self.sels[1])

    def when_transition(self):
        return self.when(self.sels[2])

    def assert_zero(self, x):
        self.accumulator = self.accumulator * self.alpha + Val.of(x)  # folder.rs:60-64 / :101-105

    def assert_eq(self, x, y):
        self.assert_zero(Val.of(x) - Val.of(y))

    def assert_one(self, x):
        self.assert_zero(Val.of(x) - 1)


def quotient_values(air, public_values, trace_lde_committed: np.ndarray, log_n: int, log_quotient_degree: int, alpha):
    """prover.rs:122-194 row by row.  trace_lde_committed: the committed (bit-reversed) LDE, canonical; its first m rows
    re-bit-reversed are the trace on the quotient domain g*H_m (two_adic_pcs.rs:247-258).
    Returns the quotient chunks [(n, 4) canonical arrays] after flatten_to_base + split_evals (prover.rs:79-80)."""
    log_m = log_n + log_quotient_degree
    m, n = 1 << log_m, 1 << log_n
    tq = orc.bit_reverse_rows(np.ascontiguousarray(trace_lde_committed[:m]))
    w_m, w_n_inv = orc.two_adic_generator(log_m), pow(orc.two_adic_generator(log_n), P - 2, P)
    next_step = 1 << log_quotient_degree
    out = np.zeros((m, 4), dtype=np.uint32)
    x = 31
    for i in range(m):
        zh = (pow(x, n, P) - 1) % P
        f = ConstraintFolder([int(v) for v in tq[i]], [int(v) for v in tq[(i + next_step) % m]], public_values,
                             zh * pow((x - 1) % P, P - 2, P) % P, zh * pow((x - w_n_inv) % P, P - 2, P) % P, (x - w_n_inv) % P, alpha)
        air.eval(f)
        q = f.accumulator * Val(pow(zh, P - 2, P))  # prover.rs:177
        out[i] = q.ext()
        x = x * w_m % P
    qd = 1 << log_quotient_degree
    return [np.ascontiguousarray(out[k::qd]) for k in range(qd)]  # split_evals: vertically strided


class VerificationError(Exception):
    pass


def _ef_inv(a):
    return V.ef_inv(a)


def _zp_at_point(log_n: int, shift: int, z):
    """TwoAdicMultiplicativeCoset::zp_at_point [MEM]: (z / shift)^n - 1."""
    t = pyref.ef_scale([int(c) for c in z], pow(shift, P - 2, P))
    r = V.ef_pow(t, 1 << log_n)
    return pyref.ef_sub(r, [1, 0, 0, 0])


def verify(log_blowup: int, num_queries: int, pow_bits: int, air, challenger, proof, public_values, log_quotient_degree: int):
    """uni-stark/src/verifier.rs:20-163.  `challenger`: a fresh oracle BfChallenger.  `log_quotient_degree` is what
    get_log_quotient_degree gives for the AIR (symbolic_builder.rs:15-32); the caller states it so that this file does
    not depend on the product's symbolic builder."""
    degree_bits = proof.degree_bits
    qd = 1 << log_quotient_degree
    width = air.width()
    ov = proof.opened_values
    if (len(ov.trace_local) != width or len(ov.trace_next) != width or len(ov.quotient_chunks) != qd
            or any(len(qc) != 4 for qc in ov.quotient_chunks)):
        raise VerificationError("InvalidProofShape")  # :49-59
    challenger.observe_digest(proof.commitments.trace)  # :69
    alpha = [int(x) for x in challenger.sample_ef()]
    challenger.observe_digest(proof.commitments.quotient_chunks)  # :72
    zeta = [int(x) for x in challenger.sample_ef()]
    g_n = orc.two_adic_generator(degree_bits)
    zeta_next = pyref.ef_scale(zeta, g_n)  # :75
    w_m = orc.two_adic_generator(degree_bits + log_quotient_degree)
    chunk_shifts = [31 * pow(w_m, k, P) % P for k in range(qd)]  # split_domains of the disjoint domain
    rounds = [
        (proof.commitments.trace, [(degree_bits, [(zeta, [list(map(int, v)) for v in ov.trace_local]),
                                                  (zeta_next, [list(map(int, v)) for v in ov.trace_next])])]),
        (proof.commitments.quotient_chunks, [(degree_bits, [(zeta, [list(map(int, v)) for v in qc])]) for qc in ov.quotient_chunks]),
    ]
    try:
        V.pcs_verify(log_blowup, num_queries, pow_bits, rounds, proof.opening_proof, challenger)  # :77-100
    except V.VerifyError as e:
        raise VerificationError(f"InvalidOpeningArgument({e})")
    # :102-133  quotient(zeta) from the chunks
    zps = []
    for i in range(qd):
        acc = [1, 0, 0, 0]
        for j in range(qd):
            if j == i:
                continue
            num = _zp_at_point(degree_bits, chunk_shifts[j], zeta)
            den = _zp_at_point(degree_bits, chunk_shifts[j], [chunk_shifts[i], 0, 0, 0])  # first_point of domain i
            acc = pyref.ef_mul(acc, pyref.ef_mul(num, _ef_inv(den)))
        zps.append(acc)
    quotient = [0, 0, 0, 0]
    for i, qc in enumerate(ov.quotient_chunks):
        for e_i in range(4):
            mono = [0, 0, 0, 0]
            mono[e_i] = 1
            term = pyref.ef_mul(pyref.ef_mul(zps[i], mono), [int(x) for x in qc[e_i]])
            quotient = pyref.ef_add(quotient, term)
    # :135-152  constraints at zeta
    zh = _zp_at_point(degree_bits, 1, zeta)
    w_n_inv = pow(g_n, P - 2, P)
    d_first, d_last = pyref.ef_sub(zeta, [1, 0, 0, 0]), pyref.ef_sub(zeta, [w_n_inv, 0, 0, 0])
    f = ConstraintFolder([Val(list(map(int, v))) for v in ov.trace_local], [Val(list(map(int, v))) for v in ov.trace_next],
                         public_values, Val(pyref.ef_mul(zh, _ef_inv(d_first))), Val(pyref.ef_mul(zh, _ef_inv(d_last))), Val(d_last), alpha)
    air.eval(f)
    if pyref.ef_mul(f.accumulator.ext(), _ef_inv(zh)) != quotient:  # :154-158
        raise VerificationError("OodEvaluationMismatch")
    return True
