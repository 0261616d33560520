"""Independent pure-Python restatement (big ints, definitions only) used to PIN the C oracle.

TEST INFRASTRUCTURE ONLY.  Deliberately shares no code with oracle/tapstark_oracle.c: every function
here is the mathematical definition (O(n^2) sums, Python ints) or the SURVEY Appendix-A model of the
reference (challenger).  tests/golden/make_golden.py uses it, together with the third-party `blake3`
PyPI package (the official Rust implementation behind a Python binding), to write the committed
golden vectors.
"""
from __future__ import annotations

P = 0x78000001
GENERATOR = 31
W = 11  # x^4 = 11


def two_adic_generator(bits: int) -> int:
    return pow(0x1A427A41, 1 << (27 - bits), P)


def bitrev(x: int, bits: int) -> int:
    return int(format(x, f"0{bits}b")[::-1], 2) if bits else 0


def idft_def(col):
    n = len(col)
    w_inv = pow(two_adic_generator(n.bit_length() - 1), P - 2, P)
    n_inv = pow(n, P - 2, P)
    return [sum(col[x] * pow(w_inv, k * x, P) for x in range(n)) * n_inv % P for k in range(n)]


def eval_poly(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % P
    return acc


def lde_committed_def(rows, added_bits: int, shift: int = GENERATOR):
    """Committed LDE by definition: row r, column c = p_c(shift * w_N^bitrev(r))."""
    n, w = len(rows), len(rows[0])
    log_n = n.bit_length() - 1
    log_N = log_n + added_bits
    N = 1 << log_N
    wN = two_adic_generator(log_N)
    coeffs = [idft_def([rows[r][c] for r in range(n)]) for c in range(w)]
    out = []
    for r in range(N):
        x = shift * pow(wN, bitrev(r, log_N), P) % P
        out.append([eval_poly(coeffs[c], x) for c in range(w)])
    return out


# ---- extension field --------------------------------------------------------------------------
def ef_add(a, b):
    return [(x + y) % P for x, y in zip(a, b)]


def ef_sub(a, b):
    return [(x - y) % P for x, y in zip(a, b)]


def ef_mul(a, b):
    r = [0] * 7
    for i in range(4):
        for j in range(4):
            r[i + j] = (r[i + j] + a[i] * b[j]) % P
    return [(r[i] + W * r[i + 4]) % P if i < 3 else r[i] for i in range(4)]


def ef_scale(a, s):
    return [x * s % P for x in a]


def fold_def_ef(vals, beta):
    """out[i] = 1/2 (lo+hi) + beta/2 * g_inv^bitrev(i) * (lo-hi)   ([CHK])."""
    h = len(vals) // 2
    log_h = h.bit_length() - 1
    g_inv = pow(two_adic_generator(log_h + 1), P - 2, P)
    half = pow(2, P - 2, P)
    out = []
    for i in range(h):
        lo, hi = vals[2 * i], vals[2 * i + 1]
        s = ef_scale(ef_add(lo, hi), half)
        d = ef_scale(ef_sub(lo, hi), half * pow(g_inv, bitrev(i, log_h), P) % P)
        out.append(ef_add(s, ef_mul(beta, d)))
    return out


# ---- challenger (SURVEY App. A model of basic/src/challenger/mod.rs) ----------------------------
class PyChallenger:
    def __init__(self, hash_fn):
        self.hash_fn = hash_fn  # bytes -> 32 bytes (Blake3)
        self.state = [b"\0\0\0\0"] * 16
        self.inp = []
        self.out = []

    def _duplex(self):
        for i, v in enumerate(self.inp):
            self.state[i] = v
        self.inp = []
        h = self.hash_fn(b"".join(self.state))
        self.state = [b"\0\0\0\0"] * 8 + [h[4 * i : 4 * i + 4] for i in range(8)]
        self.out = list(self.state[8:16])

    def observe(self, word: bytes):
        self.out = []
        self.inp.append(bytes(word))
        if len(self.inp) == 8:
            self._duplex()

    def observe_digest(self, d: bytes):
        for i in range(8):
            self.observe(d[4 * i : 4 * i + 4])

    def sample_bb(self) -> int:
        if self.inp or not self.out:
            self._duplex()
        return int.from_bytes(self.out.pop(), "little") % P

    def sample_ef(self):
        return [self.sample_bb() for _ in range(4)]


# ---- Merkle (P3 inject layout, equal heights only) ---------------------------------------------
def merkle_root_single(hash_fn, rows):
    """Blake3 row hash (canonical LE u32 bytes) + blake3(left||right) binary tree, one matrix."""
    layer = [hash_fn(b"".join(int(v).to_bytes(4, "little") for v in row)) for row in rows]
    layers = [layer]
    while len(layer) > 1:
        layer = [hash_fn(layer[2 * i] + layer[2 * i + 1]) for i in range(len(layer) // 2)]
        layers.append(layer)
    return layer[0], layers
