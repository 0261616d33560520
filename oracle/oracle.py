"""ctypes front-end of the CPU ORACLE (oracle/tapstark_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (tap-stark_b200/) never imports this module.

All values are CANONICAL BabyBear u32 (not Montgomery) unless a function says otherwise.
Extension-field (BabyBear^4) arrays carry a trailing axis of 4 coefficients, low degree first
(reference: basic/src/field/mod.rs:53-64).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

P = 0x78000001
GENERATOR = 31  # [MEM] p3-baby-bear Val::generator(); fri/src/two_adic_pcs.rs:235,254
LAYOUT_P3_INJECT = 0
LAYOUT_PADDED = 1

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libtapstark_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile the C oracle with gcc (recipe: oracle/Makefile)."""
    src = _HERE / "tapstark_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < max(
        src.stat().st_mtime, (_HERE / "tapstark_oracle.h").stat().st_mtime
    ):
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


class Challenger(C.Structure):
    _fields_ = [
        ("state", C.c_uint8 * 64),
        ("in_buf", C.c_uint8 * 64),
        ("n_in", C.c_int),
        ("out_buf", C.c_uint8 * 64),
        ("n_out", C.c_int),
        ("fake_perm", C.c_int),
    ]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(str(_LIB_PATH))
    u32p = C.POINTER(C.c_uint32)
    u8p = C.POINTER(C.c_uint8)
    szp = C.POINTER(C.c_size_t)
    sigs = {
        "or_bb_mul": (C.c_uint32, [C.c_uint32, C.c_uint32]),
        "or_bb_pow": (C.c_uint32, [C.c_uint32, C.c_uint64]),
        "or_bb_inv": (C.c_uint32, [C.c_uint32]),
        "or_two_adic_generator": (C.c_uint32, [C.c_uint]),
        "or_to_monty": (C.c_uint32, [C.c_uint32]),
        "or_from_monty": (C.c_uint32, [C.c_uint32]),
        "or_to_monty_vec": (None, [u32p, C.c_size_t]),
        "or_from_monty_vec": (None, [u32p, C.c_size_t]),
        "or_ef_mul": (None, [u32p, u32p, u32p]),
        "or_ef_inv": (None, [u32p, u32p]),
        "or_naive_dft": (None, [u32p, u32p, C.c_uint, C.c_size_t]),
        "or_dft_batch": (None, [u32p, C.c_uint, C.c_size_t]),
        "or_idft_batch": (None, [u32p, C.c_uint, C.c_size_t]),
        "or_coset_dft_batch": (None, [u32p, C.c_uint, C.c_size_t, C.c_uint32]),
        "or_coset_lde_batch": (None, [u32p, C.c_uint, C.c_size_t, C.c_uint, C.c_uint32, u32p]),
        "or_bit_reverse_rows": (None, [u32p, C.c_uint, C.c_size_t]),
        "or_pcs_lde_committed": (None, [u32p, C.c_uint, C.c_size_t, C.c_uint, C.c_uint32, u32p]),
        "or_blake3": (None, [u8p, C.c_size_t, u8p]),
        "or_mmcs_commit": (C.c_void_p, [C.POINTER(u32p), szp, szp, C.c_size_t, C.c_int, u8p]),
        "or_tree_depth": (C.c_size_t, [C.c_void_p]),
        "or_tree_num_layers": (C.c_size_t, [C.c_void_p]),
        "or_tree_layer": (u8p, [C.c_void_p, C.c_size_t, szp]),
        "or_mmcs_open_batch": (None, [C.c_void_p, C.c_size_t, u32p, u8p]),
        "or_mmcs_verify_batch": (C.c_int, [szp, szp, C.c_size_t, C.c_int, C.c_size_t, u32p, u8p, C.c_size_t, u8p]),
        "or_tree_free": (None, [C.c_void_p]),
        "or_padded_leaf": (C.c_size_t, [C.POINTER(u32p), szp, szp, C.c_size_t, C.c_size_t, u32p]),
        "or_sha256": (None, [u8p, C.c_size_t, u8p]),
        "or_tap_leaf_hash": (None, [u8p, C.c_size_t, u8p]),
        "or_tap_branch_hash": (None, [u8p, u8p, u8p]),
        "or_fold_matrix_bb": (None, [u32p, C.c_uint, C.c_uint32, u32p]),
        "or_fold_matrix_ef": (None, [u32p, C.c_uint, u32p, u32p]),
        "or_fold_row_ef": (None, [C.c_size_t, C.c_uint, u32p, u32p, u32p, u32p]),
        "or_dot_ext_powers": (None, [u32p, C.c_size_t, C.c_size_t, u32p, u32p]),
        "or_chal_init": (None, [C.POINTER(Challenger), C.c_int]),
        "or_chal_observe": (None, [C.POINTER(Challenger), u8p]),
        "or_chal_observe_digest": (None, [C.POINTER(Challenger), u8p]),
        "or_chal_sample_bb": (C.c_uint32, [C.POINTER(Challenger)]),
        "or_chal_sample_ef": (None, [C.POINTER(Challenger), u32p]),
        "or_chal_sample_bits": (C.c_size_t, [C.POINTER(Challenger), C.c_uint, C.c_int]),
        "or_chal_check_witness": (C.c_int, [C.POINTER(Challenger), C.c_uint, C.c_uint32, C.c_int]),
        "or_chal_grind": (C.c_uint32, [C.POINTER(Challenger), C.c_uint, C.c_int]),
        "or_fri_commit_phase": (C.c_int, [C.POINTER(u32p), szp, C.c_size_t, C.c_uint, C.POINTER(Challenger), u8p, u32p, C.POINTER(u32p), u32p]),
        "or_num_threads": (C.c_int, []),
        "or_set_num_threads": (None, [C.c_int]),
        "or_splitmix_fill": (None, [u32p, C.c_size_t, C.c_uint64]),
        "or_merkle_root_from_leaves": (None, [u8p, C.c_size_t, u8p]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _u32p(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _u8p(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def log2_strict(x: int) -> int:
    l = x.bit_length() - 1
    assert 1 << l == x, f"{x} is not a power of two"
    return l


# ----------------------------------------------------------------------------- field helpers
def to_monty(a: np.ndarray) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.uint32).copy()
    lib().or_to_monty_vec(_u32p(out.reshape(-1)), out.size)
    return out


def from_monty(a: np.ndarray) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.uint32).copy()
    lib().or_from_monty_vec(_u32p(out.reshape(-1)), out.size)
    return out


def two_adic_generator(bits: int) -> int:
    return lib().or_two_adic_generator(bits)


def ef_mul(a, b) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().or_ef_mul(_u32p(a), _u32p(b), _u32p(o))
    return o


# ----------------------------------------------------------------------------- synthetic inputs
def splitmix_matrix(seed: int, rows: int, width: int) -> np.ndarray:
    """Counter-based synthetic trace: element i = SplitMix64(seed*2^40 + i) mod p, canonical."""
    if rows * width >= 1 << 24:  # big traces: the C loop (bit-identical; tests/test_oracle_pins.py compares the two)
        out = np.empty((rows, width), dtype=np.uint32)
        lib().or_splitmix_fill(_u32p(out), rows * width, seed)
        return out
    idx = np.arange(rows * width, dtype=np.uint64) + (np.uint64(seed) << np.uint64(40))
    with np.errstate(over="ignore"):
        z = idx + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z % np.uint64(P)).astype(np.uint32).reshape(rows, width)


# ----------------------------------------------------------------------------- DFT family
def naive_dft(mat: np.ndarray) -> np.ndarray:
    mat = np.ascontiguousarray(mat, dtype=np.uint32)
    out = np.empty_like(mat)
    lib().or_naive_dft(_u32p(mat), _u32p(out), log2_strict(mat.shape[0]), mat.shape[1])
    return out


def dft_batch(mat: np.ndarray) -> np.ndarray:
    out = np.ascontiguousarray(mat, dtype=np.uint32).copy()
    lib().or_dft_batch(_u32p(out), log2_strict(out.shape[0]), out.shape[1])
    return out


def idft_batch(mat: np.ndarray) -> np.ndarray:
    out = np.ascontiguousarray(mat, dtype=np.uint32).copy()
    lib().or_idft_batch(_u32p(out), log2_strict(out.shape[0]), out.shape[1])
    return out


def coset_dft_batch(mat: np.ndarray, shift: int) -> np.ndarray:
    out = np.ascontiguousarray(mat, dtype=np.uint32).copy()
    lib().or_coset_dft_batch(_u32p(out), log2_strict(out.shape[0]), out.shape[1], shift)
    return out


def coset_lde_batch(mat: np.ndarray, added_bits: int, shift: int) -> np.ndarray:
    """[MEM] TwoAdicSubgroupDft::coset_lde_batch, natural row order."""
    mat = np.ascontiguousarray(mat, dtype=np.uint32)
    n, w = mat.shape
    out = np.empty((n << added_bits, w), dtype=np.uint32)
    lib().or_coset_lde_batch(_u32p(mat), log2_strict(n), w, added_bits, shift, _u32p(out))
    return out


def bit_reverse_rows(mat: np.ndarray) -> np.ndarray:
    out = np.ascontiguousarray(mat, dtype=np.uint32).copy()
    lib().or_bit_reverse_rows(_u32p(out), log2_strict(out.shape[0]), out.shape[1])
    return out


def pcs_lde_committed(mat: np.ndarray, added_bits: int, shift: int = GENERATOR) -> np.ndarray:
    """fri/src/two_adic_pcs.rs:235-240: the matrix the MMCS commits to (bit-reversed rows)."""
    mat = np.ascontiguousarray(mat, dtype=np.uint32)
    n, w = mat.shape
    out = np.empty((n << added_bits, w), dtype=np.uint32)
    lib().or_pcs_lde_committed(_u32p(mat), log2_strict(n), w, added_bits, shift, _u32p(out))
    return out


# ----------------------------------------------------------------------------- Blake3 / MMCS
def blake3(data: bytes) -> bytes:
    buf = np.frombuffer(data, dtype=np.uint8).copy() if len(data) else np.zeros(1, dtype=np.uint8)
    out = np.zeros(32, dtype=np.uint8)
    lib().or_blake3(_u8p(buf), len(data), _u8p(out))
    return out.tobytes()


class Tree:
    """Prover data of the oracle MMCS.  Keeps the committed matrices alive."""

    def __init__(self, mats, layout=LAYOUT_P3_INJECT):
        self.mats = [np.ascontiguousarray(m, dtype=np.uint32).reshape(m.shape[0], -1) for m in mats]
        self.layout = layout
        k = len(self.mats)
        self._ptrs = (C.POINTER(C.c_uint32) * k)(*[_u32p(m) for m in self.mats])
        self._h = (C.c_size_t * k)(*[m.shape[0] for m in self.mats])
        self._w = (C.c_size_t * k)(*[m.shape[1] for m in self.mats])
        root = np.zeros(32, dtype=np.uint8)
        self._t = lib().or_mmcs_commit(self._ptrs, self._h, self._w, k, layout, _u8p(root))
        self.root = root.tobytes()
        self.depth = lib().or_tree_depth(self._t)

    def layer(self, i: int) -> np.ndarray:
        n = C.c_size_t()
        p = lib().or_tree_layer(self._t, i, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value, 32)).copy()

    def open_batch(self, index: int):
        total = sum(m.shape[1] for m in self.mats)
        rows = np.zeros(total, dtype=np.uint32)
        path = np.zeros((max(self.depth, 1), 32), dtype=np.uint8)
        lib().or_mmcs_open_batch(self._t, index, _u32p(rows), _u8p(path.reshape(-1)))
        out, o = [], 0
        for m in self.mats:
            out.append(rows[o : o + m.shape[1]].copy())
            o += m.shape[1]
        return out, path[: self.depth].copy()

    def verify_batch(self, index: int, rows, path, root: bytes | None = None) -> bool:
        flat = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.uint32) for r in rows]))
        path = np.ascontiguousarray(path, dtype=np.uint8).reshape(-1)
        if path.size == 0:
            path = np.zeros(32, dtype=np.uint8)
        r = np.frombuffer(root if root is not None else self.root, dtype=np.uint8).copy()
        return bool(
            lib().or_mmcs_verify_batch(self._h, self._w, len(self.mats), self.layout, index,
                                       _u32p(flat), _u8p(path), self.depth, _u8p(r))
        )

    def __del__(self):
        try:
            lib().or_tree_free(self._t)
        except Exception:
            pass


def mmcs_commit(mats, layout=LAYOUT_P3_INJECT) -> Tree:
    return Tree(mats, layout)


def mmcs_verify_batch(heights, rows, index: int, path, root: bytes, layout=LAYOUT_P3_INJECT) -> bool:
    """basic/src/mmcs/bf_mmcs.rs:54-62 verify_batch without a Tree object: heights of the committed matrices, the
    opened rows (canonical), the authentication path and the root."""
    k = len(heights)
    hs = (C.c_size_t * k)(*heights)
    ws = (C.c_size_t * k)(*[len(r) for r in rows])
    flat = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.uint32) for r in rows]))
    path = np.ascontiguousarray(path, dtype=np.uint8)
    depth = path.shape[0] if path.size else 0
    pbuf = path.reshape(-1) if path.size else np.zeros(32, dtype=np.uint8)
    r = np.frombuffer(root, dtype=np.uint8).copy()
    return bool(lib().or_mmcs_verify_batch(hs, ws, k, layout, index, _u32p(flat), _u8p(pbuf), depth, _u8p(r)))


def merkle_root_from_leaves(leaves: np.ndarray) -> bytes:
    """Root of the 2-to-1 Blake3 tree over a power-of-two number of 32-byte leaf digests ([n, 32] uint8)."""
    leaves = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, 32)
    log2_strict(leaves.shape[0])
    root = np.zeros(32, dtype=np.uint8)
    lib().or_merkle_root_from_leaves(_u8p(leaves), leaves.shape[0], _u8p(root))
    return root.tobytes()


def splitmix_columns(seed: int, rows: int, total_width: int, cols) -> np.ndarray:
    """Columns `cols` of splitmix_matrix(seed, rows, total_width) without materialising the matrix."""
    idx = (np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(total_width) + np.asarray(cols, dtype=np.uint64)[None, :]
           + (np.uint64(seed) << np.uint64(40)))
    with np.errstate(over="ignore"):
        z = idx + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z % np.uint64(P)).astype(np.uint32)


def padded_leaf(mats, leaf: int) -> np.ndarray:
    mats = [np.ascontiguousarray(m, dtype=np.uint32) for m in mats]
    k = len(mats)
    ptrs = (C.POINTER(C.c_uint32) * k)(*[_u32p(m) for m in mats])
    h = (C.c_size_t * k)(*[m.shape[0] for m in mats])
    w = (C.c_size_t * k)(*[m.shape[1] for m in mats])
    out = np.zeros(sum(m.shape[1] for m in mats), dtype=np.uint32)
    n = lib().or_padded_leaf(ptrs, h, w, k, leaf, _u32p(out))
    return out[:n]


# ----------------------------------------------------------------------------- fold
def fold_matrix_bb(vals: np.ndarray, beta: int) -> np.ndarray:
    vals = np.ascontiguousarray(vals, dtype=np.uint32).reshape(-1)
    h = vals.size // 2
    out = np.empty(h, dtype=np.uint32)
    lib().or_fold_matrix_bb(_u32p(vals), log2_strict(h), beta, _u32p(out))
    return out


def fold_matrix_ef(vals: np.ndarray, beta) -> np.ndarray:
    """vals: (2h, 4) EF codeword viewed as h x 2; returns (h, 4).  fri/src/two_adic_pcs.rs:116-147."""
    vals = np.ascontiguousarray(vals, dtype=np.uint32).reshape(-1, 4)
    h = vals.shape[0] // 2
    beta = np.ascontiguousarray(beta, dtype=np.uint32)
    out = np.empty((h, 4), dtype=np.uint32)
    lib().or_fold_matrix_ef(_u32p(vals.reshape(-1)), log2_strict(h), _u32p(beta), _u32p(out.reshape(-1)))
    return out


def fold_row_ef(index: int, log_height: int, beta, e0, e1) -> np.ndarray:
    out = np.zeros(4, dtype=np.uint32)
    a = [np.ascontiguousarray(x, dtype=np.uint32) for x in (beta, e0, e1)]
    lib().or_fold_row_ef(index, log_height, _u32p(a[0]), _u32p(a[1]), _u32p(a[2]), _u32p(out))
    return out


def dot_ext_powers(m: np.ndarray, alpha) -> np.ndarray:
    """fri/src/two_adic_pcs.rs:375: (rows, 4) EF vector sum_c alpha^c * m[:, c]."""
    m = np.ascontiguousarray(m, dtype=np.uint32)
    a = np.ascontiguousarray(alpha, dtype=np.uint32)
    out = np.empty((m.shape[0], 4), dtype=np.uint32)
    lib().or_dot_ext_powers(_u32p(m.reshape(-1)), m.shape[0], m.shape[1], _u32p(a), _u32p(out.reshape(-1)))
    return out


# ----------------------------------------------------------------------------- challenger
class BfChallenger:
    """basic/src/challenger/mod.rs BfChallenger<F, U32, Blake3Permutation, 16>."""

    def __init__(self, fake_perm: bool = False, ext: bool = True):
        self.c = Challenger()
        self.ext = ext
        lib().or_chal_init(C.byref(self.c), int(fake_perm))

    def clone(self) -> "BfChallenger":
        o = BfChallenger.__new__(BfChallenger)
        o.c = Challenger.from_buffer_copy(self.c)
        o.ext = self.ext
        return o

    def observe(self, word) -> None:
        if isinstance(word, int):
            word = word.to_bytes(4, "little")
        buf = (C.c_uint8 * 4)(*word)
        lib().or_chal_observe(C.byref(self.c), buf)

    def observe_digest(self, digest: bytes) -> None:
        assert len(digest) == 32
        buf = (C.c_uint8 * 32)(*digest)
        lib().or_chal_observe_digest(C.byref(self.c), buf)

    def sample_bb(self) -> int:
        return lib().or_chal_sample_bb(C.byref(self.c))

    def sample_ef(self) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint32)
        lib().or_chal_sample_ef(C.byref(self.c), _u32p(out))
        return out

    def sample_bits(self, bits: int) -> int:
        return lib().or_chal_sample_bits(C.byref(self.c), bits, int(self.ext))

    def check_witness(self, bits: int, witness: int) -> bool:
        return bool(lib().or_chal_check_witness(C.byref(self.c), bits, witness, int(self.ext)))

    def grind(self, bits: int) -> int:
        return lib().or_chal_grind(C.byref(self.c), bits, int(self.ext))


# ----------------------------------------------------------------------------- commit phase
def fri_commit_phase(inputs, log_blowup: int, chal: BfChallenger, want_layers: bool = False):
    """fri/src/prover.rs:93-141.  inputs: list of (len_i, 4) EF arrays, lengths descending.
    Returns dict(commits=[bytes], final_poly, betas, layers?)."""
    inputs = [np.ascontiguousarray(v, dtype=np.uint32).reshape(-1, 4) for v in inputs]
    k = len(inputs)
    ptrs = (C.POINTER(C.c_uint32) * k)(*[_u32p(v.reshape(-1)) for v in inputs])
    lens = (C.c_size_t * k)(*[v.shape[0] for v in inputs])
    max_rounds = max(log2_strict(inputs[0].shape[0]) - log_blowup, 0)
    commits = np.zeros(max(max_rounds, 1) * 32, dtype=np.uint8)
    final = np.zeros(4, dtype=np.uint32)
    betas = np.zeros(max(max_rounds, 1) * 4, dtype=np.uint32)
    layer_ptrs = (C.POINTER(C.c_uint32) * max(max_rounds, 1))()
    r = lib().or_fri_commit_phase(ptrs, lens, k, log_blowup, C.byref(chal.c), _u8p(commits), _u32p(final),
                                  layer_ptrs if want_layers else None, _u32p(betas))
    res = {
        "ok": r >= 0,
        "rounds": max_rounds,
        "commits": [commits[32 * i : 32 * i + 32].tobytes() for i in range(max_rounds)],
        "final_poly": final,
        "betas": betas.reshape(-1, 4)[:max_rounds].copy(),
    }
    if want_layers:
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        layers = []
        for i in range(max_rounds):
            n = inputs[0].shape[0] >> i
            layers.append(np.ctypeslib.as_array(layer_ptrs[i], shape=(n, 4)).copy())
            libc.free(C.cast(layer_ptrs[i], C.c_void_p))
        res["layers"] = layers
    return res


def num_threads() -> int:
    return lib().or_num_threads()


def use_all_cores() -> int:
    """Sets the OpenMP team to every core this process may run on (torch.distributed.run exports
    OMP_NUM_THREADS=1, which would otherwise turn the multi-core CPU arm of bench.py into a 1-core run)."""
    import os

    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().or_set_num_threads(n)
    return num_threads()
