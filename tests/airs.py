"""Example AIRs written against the AirBuilder surface, as in the reference's tests: the same `eval` runs on the
product's SymbolicAirBuilder (-> constraint program -> CUDA) and on the oracle's folders (direct field arithmetic)."""
import numpy as np

P = 0x78000001


class FibonacciAir:
    """uni-stark/tests/fib_air.rs:21-58: public values (a, b, x); first row = (a, b); a' = b, b' = a + b; last b = x."""

    def width(self):
        return 2

    def eval(self, builder):
        local, nxt = builder.main()
        pis = builder.public_values()
        a, b, x = pis[0], pis[1], pis[2]
        when_first_row = builder.when_first_row()
        when_first_row.assert_eq(local[0], a)
        when_first_row.assert_eq(local[1], b)
        when_transition = builder.when_transition()
        when_transition.assert_eq(local[1], nxt[0])
        when_transition.assert_eq(local[0] + local[1], nxt[1])
        builder.when_last_row().assert_eq(local[1], x)


def fibonacci_trace(a: int, b: int, n: int) -> np.ndarray:
    """generate_trace_rows (fib_air.rs:60-80)"""
    t = np.zeros((n, 2), dtype=np.uint32)
    t[0] = (a % P, b % P)
    for i in range(1, n):
        t[i, 0] = t[i - 1, 1]
        t[i, 1] = (int(t[i - 1, 0]) + int(t[i - 1, 1])) % P
    return t


class MulAir:
    """After the shape of uni-stark/tests/mul_air.rs (commented out upstream): `reps` triples (a, b, c) per row with
    a^(degree-1) * b = c; on transition rows the first a of the next row is the first a of this row + 1; the first
    row's first a is 0 -- so constraint degree `degree`, quotient degree 2^ceil(log2(degree-1))."""

    def __init__(self, degree: int = 3, reps: int = 2):
        self.degree, self.reps = degree, reps

    def width(self):
        return 3 * self.reps

    def eval(self, builder):
        local, nxt = builder.main()
        for r in range(self.reps):
            a, b, c = local[3 * r], local[3 * r + 1], local[3 * r + 2]
            t = a
            for _ in range(self.degree - 2):
                t = t * a
            builder.assert_zero(t * b - c)
        builder.when_first_row().assert_zero(local[0])
        builder.when_transition().assert_eq(local[0] + 1, nxt[0])


def mul_trace(air: MulAir, n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.zeros((n, air.width()), dtype=np.uint32)
    for i in range(n):
        for r in range(air.reps):
            a = i if r == 0 else int(rng.integers(0, P))
            b = int(rng.integers(0, P))
            t[i, 3 * r], t[i, 3 * r + 1] = a, b
            t[i, 3 * r + 2] = pow(a, air.degree - 1, P) * b % P
    return t


class CounterAir:
    """Exercises the rest of the AirBuilder surface: constants, negation, assert_one, when(<expression>).
    Columns (i, flag, sq, neg): i counts 0,1,2,...; flag = 1 on every row; sq = i*i; neg = -i.  Constraints:
    flag == 1; when(flag): sq - i*i == 0; neg + i == 0 written as -(i) - neg == 0; first row i == 0;
    transition i' == i + 1; last row i == public[0]."""

    def width(self):
        return 4

    def eval(self, builder):
        local, nxt = builder.main()
        i, flag, sq, neg = local
        builder.assert_one(flag)
        builder.when(flag).assert_zero(sq - i * i)
        builder.assert_zero(-i - neg)
        builder.when_first_row().assert_zero(i)
        builder.when_transition().assert_eq(nxt[0], i + 1)
        builder.when_last_row().assert_eq(i, builder.public_values()[0])


def counter_trace(n: int) -> np.ndarray:
    t = np.zeros((n, 4), dtype=np.uint32)
    for i in range(n):
        t[i] = (i, 1, i * i % P, (P - i) % P)
    return t
