"""tapstark_b200 -- host-side mirror of the reference's Plonky3-derived interfaces over the C ABI of
libtapstark_b200.so (include/tapstark.h).

The reference is a Rust workspace; no Rust toolchain exists in this image, so the host side above the C ABI
is mirrored here with the reference's names and argument meaning, and panics become `TapStarkError`:

    GpuDft                 p3_dft::TwoAdicSubgroupDft<BabyBear>            fri/src/two_adic_pcs.rs:207,237-240
    Blake3MerkleMmcs       basic::mmcs::bf_mmcs::BFMmcs<T>                 basic/src/mmcs/bf_mmcs.rs:17-68
    BfChallenger           basic::challenger::BfChallenger                 basic/src/challenger/mod.rs
    FriConfig              fri::FriConfig                                  fri/src/config.rs:10-22
    TwoAdicFriPcs          basic::bf_pcs::Pcs for TwoAdicFriPcs            fri/src/two_adic_pcs.rs:197-258
    fold_even_odd          fri::fold_even_odd                              fri/src/fold_even_odd.rs:20-52
    bf_commit_phase        fri::prover::bf_commit_phase                    fri/src/prover.rs:93-141

All compute happens in the CUDA library.  There is no CPU fallback: if the library (or a GPU) is missing,
constructing a Context raises.  This package never imports anything from oracle/.

Values crossing this API are CANONICAL numpy uint32 unless a name says `monty`; conversion to the ABI's
Montgomery form is a cheap host-side numpy step for test-sized inputs and a device kernel otherwise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

P = 0x78000001
GENERATOR = 31
MONTY_R = (1 << 32) % P
MONTY_RINV = pow(MONTY_R, P - 2, P)
LAYOUT_P3_INJECT = 0
LAYOUT_PADDED = 1
K_NTT_PASS, K_LDE_MID, K_HASH_LEAVES, K_TREE, K_FOLD, K_MISC = range(6)
KERNEL_KINDS = ["ntt_pass", "lde_mid", "hash_leaves", "tree", "fold", "misc"]

_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / "libtapstark_b200.so"


class TapStarkError(RuntimeError):
    """Raised where the reference would panic (assert!/expect) or a CUDA call fails."""


_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_szp = C.POINTER(C.c_size_t)
_vp = C.c_void_p
_vpp = C.POINTER(C.c_void_p)

_SIGNATURES = {
    "ts_is_device_build": (C.c_int, []),
    "ts_ctx_create": (C.c_int, [C.c_int, _vp, _vpp]),
    "ts_ctx_destroy": (None, [_vp]),
    "ts_last_error": (C.c_char_p, [_vp]),
    "ts_ctx_synchronize": (C.c_int, [_vp]),
    "ts_ctx_set_profiling": (C.c_int, [_vp, C.c_int]),
    "ts_ctx_reset_stats": (C.c_int, [_vp]),
    "ts_ctx_get_stats": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "ts_ctx_total_launches": (C.c_uint64, [_vp]),
    "ts_ctx_trim": (C.c_int, [_vp]),
    "ts_matrix_alloc": (C.c_int, [_vp, C.c_size_t, C.c_size_t, _vpp]),
    "ts_matrix_from_host": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vpp]),
    "ts_matrix_from_device": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vpp]),
    "ts_matrix_wrap_device": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vpp]),
    "ts_matrix_download": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp]),
    "ts_matrix_device_ptr": (_vp, [_vp]),
    "ts_matrix_rows": (C.c_size_t, [_vp]),
    "ts_matrix_width": (C.c_size_t, [_vp]),
    "ts_matrix_free": (None, [_vp]),
    "ts_matrix_to_monty": (C.c_int, [_vp, _vp]),
    "ts_matrix_from_monty": (C.c_int, [_vp, _vp]),
    "ts_matrix_bit_reverse_rows": (C.c_int, [_vp, _vp, _vpp]),
    "ts_coset_lde_batch": (C.c_int, [_vp, _vp, C.c_uint, C.c_uint32, C.c_int, _vpp]),
    "ts_dft_batch": (C.c_int, [_vp, _vp, _vpp]),
    "ts_idft_batch": (C.c_int, [_vp, _vp, _vpp]),
    "ts_coset_dft_batch": (C.c_int, [_vp, _vp, C.c_uint32, _vpp]),
    "ts_lde_batch": (C.c_int, [_vp, _vp, C.c_uint, _vpp]),
    "ts_coset_lde_batch_host": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_uint, C.c_uint32, C.c_int, _vp]),
    "ts_mmcs_commit": (C.c_int, [_vp, _vpp, C.c_size_t, C.c_int, C.c_int, _u8p, _vpp]),
    "ts_tree_num_matrices": (C.c_size_t, [_vp]),
    "ts_tree_matrix": (_vp, [_vp, C.c_size_t]),
    "ts_tree_depth": (C.c_size_t, [_vp]),
    "ts_tree_max_height": (C.c_size_t, [_vp]),
    "ts_mmcs_open_batch": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp]),
    "ts_mmcs_verify_batch": (C.c_int, [_szp, _szp, C.c_size_t, C.c_int, C.c_size_t, _vp, _vp, C.c_size_t, _vp]),
    "ts_tree_layer": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _szp]),
    "ts_tree_free": (None, [_vp]),
    "ts_challenger_new": (C.c_int, [_vpp]),
    "ts_challenger_clone": (C.c_int, [_vp, _vpp]),
    "ts_challenger_free": (None, [_vp]),
    "ts_challenger_observe": (None, [_vp, _vp]),
    "ts_challenger_observe_digest": (None, [_vp, _vp]),
    "ts_challenger_sample_base": (C.c_uint32, [_vp]),
    "ts_challenger_sample_ext": (None, [_vp, _vp]),
    "ts_challenger_sample_bits": (C.c_size_t, [_vp, C.c_uint, C.c_int]),
    "ts_challenger_check_witness": (C.c_int, [_vp, C.c_uint, C.c_uint32, C.c_int]),
    "ts_challenger_grind": (C.c_int, [_vp, C.c_uint, C.c_int, _u32p]),
    "ts_fri_fold_base": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint32, _vp]),
    "ts_fri_fold_ext": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp]),
    "ts_fri_fold_ext_host": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp]),
    "ts_fri_commit_phase": (C.c_int, [_vp, _vpp, C.c_size_t, C.c_uint, _vp, _vp, _vpp, _vp, _szp]),
    "ts_pcs_commit": (C.c_int, [_vp, _vpp, _vp, C.c_size_t, C.c_uint, C.c_int, _u8p, _vpp]),
    "ts_pcs_commit_host": (C.c_int, [_vp, _vpp, _szp, _szp, _vp, C.c_size_t, C.c_uint, C.c_int, _u8p, _vpp]),
    "ts_pcs_get_evaluations_on_domain": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp]),
    "ts_dot_ext_powers": (C.c_int, [_vp, _vp, _vp, _vpp]),
    "ts_inv_denoms": (C.c_int, [_vp, C.c_uint, _vp, _vpp]),
    "ts_interpolate_low_coset": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp, _vp]),
    "ts_reduce_opening_acc": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "ts_matrix_zero": (C.c_int, [_vp, _vp]),
    "ts_fill_splitmix": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_uint64, C.c_size_t, C.c_size_t, C.c_int]),
    "ts_coset_lde_batch_into": (C.c_int, [_vp, _vp, C.c_uint, C.c_uint32, _vp]),
    "ts_alpha_powers": (C.c_int, [_vp, _vp, C.c_size_t, _vpp]),
    "ts_dot_ext_powers_acc": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _vp, C.c_int]),
    "ts_dot_ext_powers_blocks": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp]),
    "ts_tree_root_copy": (C.c_int, [_vp, _vp, _vp]),
    "ts_coset_lde_batch_scatter": (C.c_int, [_vp, _vp, C.c_uint, C.c_uint32, _vp, C.c_size_t, C.c_size_t]),
    "ts_device_malloc": (C.c_int, [_vp, C.c_size_t, _vpp]),
    "ts_device_free": (C.c_int, [_vp, _vp]),
    "ts_ipc_get_handle": (C.c_int, [_vp, _vp, _vp]),
    "ts_ipc_open": (C.c_int, [_vp, _vp, _vpp]),
    "ts_ipc_close": (C.c_int, [_vp, _vp]),
    "ts_quotient_values": (C.c_int, [_vp, _vp, C.c_uint, C.c_uint, _vp, C.c_size_t, _vp, C.c_size_t, _vp, C.c_size_t, _vp, _vp]),
    "ts_fri_fold_ext_shard": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp, _vp, _vp]),
    "ts_fri_fold_hash_shard": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp, _vp, _vp, _vp]),
    "ts_blake3_host": (None, [_vp, C.c_size_t, _vp]),
    "ts_pcs_open": (C.c_int, [_vp, _vpp, C.c_size_t, _szp, _vp, C.c_uint, C.c_uint, C.c_uint, _vp, _vpp, _szp]),
    "ts_bytes_free": (None, [_vp]),
    "ts_fri_mailbox_words": (C.c_int, []),
    "ts_fri_commit_phase_sharded": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, _vpp, C.c_uint32, _vp, _vp, _vp]),
    "ts_mmcs_commit_begin": (C.c_int, [_vp, _vpp, C.c_size_t, _vpp]),
    "ts_mmcs_commit_window": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t]),
    "ts_mmcs_commit_finish": (C.c_int, [_vp, _vp, _vp]),
    "ts_taptree_commit": (C.c_int, [_vp, _vp, _vp, _szp, _vp, C.c_size_t, _u8p, _vpp]),
    "ts_taptree_leaf_indices": (C.c_int, [_vp, _vp, _vp]),
    "ts_taptree_level": (C.c_int, [_vp, _vp, C.c_uint, _vp]),
    "ts_taptree_free": (None, [_vp]),
    "ts_padded_leaf_rows": (C.c_int, [_vp, _vpp, C.c_size_t, _vpp]),
    "ts_taptree_open": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp]),
    "ts_copy_async": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_size_t, C.c_int]),
    "ts_copy2d_async": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t, _vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]),
    "ts_copy_join": (C.c_int, [_vp, C.c_int]),
    "ts_fri_chain_begin": (C.c_int, [_vp, _vp, C.c_size_t, _vpp]),
    "ts_fri_chain_step": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_size_t]),
    "ts_fri_fold_ext_shard_chain": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp, _vp, _vp]),
    "ts_fri_chain_end": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _vp]),
    "ts_dft_batch_host": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t, C.c_size_t, C.c_uint32, _vp]),
    "ts_host_register": (C.c_int, [_vp, _vp, C.c_size_t]),
    "ts_host_unregister": (C.c_int, [_vp, _vp]),
}
ABI_SYMBOLS = sorted(_SIGNATURES)

_lib = None


def load_library(path: Optional[Path] = None, allow_emulated: bool = False):
    """Loads the CUDA library.  `allow_emulated` exists only for tests/emul (kernel-source checks in the
    GPU-less build container); the product path never sets it and a non-device build is rejected."""
    global _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise TapStarkError(
            f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback."
        )
    L = C.CDLL(str(p))
    for name, (res, args) in _SIGNATURES.items():
        f = getattr(L, name)  # AttributeError if the ABI is incomplete
        f.restype = res
        f.argtypes = args
    if not L.ts_is_device_build() and not allow_emulated:
        raise TapStarkError(f"{p} is not a device build; refusing to run the hot path on a CPU")
    _lib = L
    return L


def lib():
    return _lib if _lib is not None else load_library()


# ------------------------------------------------------------------------------------------- field helpers
def to_monty(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.uint64)
    return ((a << np.uint64(32)) % np.uint64(P)).astype(np.uint32)


def from_monty(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.uint64)
    return ((a * np.uint64(MONTY_RINV)) % np.uint64(P)).astype(np.uint32)


def two_adic_generator(bits: int) -> int:
    return pow(0x1A427A41, 1 << (27 - bits), P)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------------------- context
class Context:
    """Owns a CUDA stream, twiddle caches and scratch.  One host thread at a time (like the reference's
    single-threaded callers)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.ts_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise TapStarkError(f"ts_ctx_create failed (rc={rc}): no usable CUDA device; there is no CPU fallback")
        self._h = h
        self.stream = stream  # the caller's cudaStream_t handle, or None for the context's private stream

    def check(self, rc: int, what: str = ""):
        if rc != 0:
            msg = self._L.ts_last_error(self._h)
            raise TapStarkError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")

    def synchronize(self):
        self.check(self._L.ts_ctx_synchronize(self._h), "synchronize")

    def set_profiling(self, on: bool):
        self._L.ts_ctx_set_profiling(self._h, int(on))

    def reset_stats(self):
        self._L.ts_ctx_reset_stats(self._h)

    def stats(self) -> dict:
        out = {}
        for k, name in enumerate(KERNEL_KINDS):
            ms, n = C.c_double(), C.c_uint64()
            self._L.ts_ctx_get_stats(self._h, k, C.byref(ms), C.byref(n))
            out[name] = {"ms": ms.value, "launches": n.value}
        return out

    def total_launches(self) -> int:
        return self._L.ts_ctx_total_launches(self._h)

    def trim(self):
        self._L.ts_ctx_trim(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._L.ts_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceMatrix:
    """Device-resident p3_matrix::RowMajorMatrix<BabyBear> (Montgomery form).  EF matrices: width = 4*w."""

    def __init__(self, ctx: Context, handle, owned: bool = True, keepalive=None):
        self.ctx, self._h, self._owned, self._keep = ctx, handle, owned, keepalive

    # constructors ------------------------------------------------------------------------------
    @classmethod
    def from_canonical(cls, ctx: Context, a) -> "DeviceMatrix":
        a = np.asarray(a, dtype=np.uint32)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        if a.ndim == 3:  # (rows, ext width, 4)
            a = a.reshape(a.shape[0], -1)
        return cls.from_monty(ctx, to_monty(a))

    @classmethod
    def from_monty(cls, ctx: Context, a: np.ndarray) -> "DeviceMatrix":
        a = np.ascontiguousarray(a, dtype=np.uint32)
        h = C.c_void_p()
        ctx.check(ctx._L.ts_matrix_from_host(ctx._h, _ptr(a), a.shape[0], a.shape[1], C.byref(h)), "matrix_from_host")
        ctx.synchronize()  # `a` may be a temporary
        return cls(ctx, h)

    @classmethod
    def wrap_device(cls, ctx: Context, dev_ptr: int, rows: int, width: int, keepalive=None) -> "DeviceMatrix":
        h = C.c_void_p()
        ctx.check(ctx._L.ts_matrix_wrap_device(ctx._h, C.c_void_p(dev_ptr), rows, width, C.byref(h)), "wrap")
        return cls(ctx, h, keepalive=keepalive)

    @classmethod
    def splitmix(cls, ctx: Context, seed: int, rows: int, width: int, col0: int = 0, total_width: Optional[int] = None,
                 monty: bool = True) -> "DeviceMatrix":
        """Columns [col0, col0+width) of the synthetic rows x total_width trace of SURVEY 8(d), generated on the device
        (ts_fill_splitmix): the same matrix as oracle.splitmix_matrix(seed, rows, total_width), in Montgomery form."""
        h = C.c_void_p()
        ctx.check(ctx._L.ts_matrix_alloc(ctx._h, rows, width, C.byref(h)), "matrix_alloc")
        m = cls(ctx, h)
        ctx.check(ctx._L.ts_fill_splitmix(ctx._h, C.c_void_p(m.device_ptr), rows, width, seed, col0,
                                          width if total_width is None else total_width, 1 if monty else 0), "fill_splitmix")
        return m

    # accessors ---------------------------------------------------------------------------------
    @property
    def rows(self) -> int:
        return self.ctx._L.ts_matrix_rows(self._h)

    @property
    def width(self) -> int:
        return self.ctx._L.ts_matrix_width(self._h)

    def height(self) -> int:
        return self.rows

    @property
    def device_ptr(self) -> int:
        return self.ctx._L.ts_matrix_device_ptr(self._h)

    def to_monty_host(self, row0: int = 0, nrows: Optional[int] = None) -> np.ndarray:
        nrows = self.rows - row0 if nrows is None else nrows
        out = np.empty((nrows, self.width), dtype=np.uint32)
        self.ctx.check(self.ctx._L.ts_matrix_download(self.ctx._h, self._h, row0, nrows, _ptr(out)), "download")
        return out

    def to_canonical(self, row0: int = 0, nrows: Optional[int] = None) -> np.ndarray:
        return from_monty(self.to_monty_host(row0, nrows))

    def bit_reverse_rows(self) -> "DeviceMatrix":
        h = C.c_void_p()
        self.ctx.check(self.ctx._L.ts_matrix_bit_reverse_rows(self.ctx._h, self._h, C.byref(h)), "bit_reverse_rows")
        return DeviceMatrix(self.ctx, h)

    def free(self):
        if self._h and self._owned:
            self.ctx._L.ts_matrix_free(self._h)
        self._h = None

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------- Dft
class GpuDft:
    """p3_dft::TwoAdicSubgroupDft<BabyBear> (the `Dft` type parameter of TwoAdicFriPcs).  Constructible as a
    value like `Radix2DitParallel` in uni-stark/tests/fib_air.rs:113,122."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def _unary(self, fn, mat: DeviceMatrix, *args) -> DeviceMatrix:
        h = C.c_void_p()
        self.ctx.check(fn(self.ctx._h, mat._h, *args, C.byref(h)), fn.__name__)
        return DeviceMatrix(self.ctx, h)

    def dft_batch(self, mat: DeviceMatrix) -> DeviceMatrix:
        return self._unary(self.ctx._L.ts_dft_batch, mat)

    def idft_batch(self, mat: DeviceMatrix) -> DeviceMatrix:
        return self._unary(self.ctx._L.ts_idft_batch, mat)

    def coset_dft_batch(self, mat: DeviceMatrix, shift: int) -> DeviceMatrix:
        return self._unary(self.ctx._L.ts_coset_dft_batch, mat, int(to_monty(shift)))

    def lde_batch(self, mat: DeviceMatrix, added_bits: int) -> DeviceMatrix:
        return self._unary(self.ctx._L.ts_lde_batch, mat, added_bits)

    def coset_lde_batch(self, mat: DeviceMatrix, added_bits: int, shift: int, committed_order: bool = False) -> DeviceMatrix:
        """Trait semantics (natural row order) by default; `committed_order=True` returns what
        `.bit_reverse_rows().to_row_major_matrix()` would give, with no extra pass (the PCS hot call)."""
        return self._unary(self.ctx._L.ts_coset_lde_batch, mat, added_bits, int(to_monty(shift)),
                           0 if committed_order else 1)

    def coset_lde_batch_host(self, evals: np.ndarray, added_bits: int, shift: int, committed_order: bool = False) -> np.ndarray:
        """Host-buffer form: canonical numpy in, canonical numpy out (H2D + LDE + D2H in one ABI call)."""
        ev = to_monty(evals)
        out = np.empty((ev.shape[0] << added_bits, ev.shape[1]), dtype=np.uint32)
        self.ctx.check(self.ctx._L.ts_coset_lde_batch_host(self.ctx._h, _ptr(ev), ev.shape[0], ev.shape[1], added_bits,
                                                           int(to_monty(shift)), 0 if committed_order else 1, _ptr(out)),
                       "coset_lde_batch_host")
        return from_monty(out)


# ------------------------------------------------------------------------------------------- Mmcs
class ProverData:
    """BFMmcs::ProverData: the digest layers plus the committed matrices (device resident)."""

    def __init__(self, ctx: Context, handle, mats: Sequence[DeviceMatrix], owns_mats: bool):
        self.ctx, self._h, self._mats, self._owns = ctx, handle, list(mats), owns_mats
        if owns_mats:  # the tree frees them
            for m in self._mats:
                m._owned = False

    @property
    def depth(self) -> int:
        return self.ctx._L.ts_tree_depth(self._h)

    def layer(self, i: int) -> np.ndarray:
        n = C.c_size_t()
        self.ctx.check(self.ctx._L.ts_tree_layer(self.ctx._h, self._h, i, None, C.byref(n)), "tree_layer")
        out = np.empty((n.value, 32), dtype=np.uint8)
        self.ctx.check(self.ctx._L.ts_tree_layer(self.ctx._h, self._h, i, _ptr(out), C.byref(n)), "tree_layer")
        return out

    def root_to_device(self, dst_ptr: int) -> None:
        """Copies the 32-byte root to device memory at dst_ptr on the context's stream (no host synchronisation)."""
        self.ctx.check(self.ctx._L.ts_tree_root_copy(self.ctx._h, self._h, C.c_void_p(dst_ptr)), "tree_root_copy")

    def free(self):
        if self._h:
            self.ctx._L.ts_tree_free(self._h)
            self._h = None

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


class Blake3MerkleMmcs:
    """BFMmcs<T> with Commitment = [[u8;4];8] (32 bytes).  See include/tapstark.h for the two leaf layouts."""

    def __init__(self, ctx: Context, layout: int = LAYOUT_P3_INJECT):
        self.ctx, self.layout = ctx, layout

    def commit(self, inputs: Sequence[DeviceMatrix], take_ownership: bool = False, host_root: bool = True):
        """host_root=False: the root stays on the device (no synchronisation); returns (None, ProverData) and the
        caller fetches it with ProverData.root_to_device (multi-GPU sub-roots go straight into the all-gather)."""
        k = len(inputs)
        arr = (C.c_void_p * k)(*[m._h for m in inputs])
        root = (C.c_uint8 * 32)()
        h = C.c_void_p()
        self.ctx.check(self.ctx._L.ts_mmcs_commit(self.ctx._h, arr, k, self.layout, int(take_ownership),
                                                  root if host_root else None, C.byref(h)), "mmcs_commit")
        return (bytes(root) if host_root else None), ProverData(self.ctx, h, inputs, take_ownership)

    def commit_matrix(self, m: DeviceMatrix) -> Tuple[bytes, ProverData]:
        return self.commit([m])

    def open_batch(self, query_index: int, prover_data: ProverData) -> Tuple[List[np.ndarray], np.ndarray]:
        L = self.ctx._L
        k = L.ts_tree_num_matrices(prover_data._h)
        widths = [L.ts_matrix_width(L.ts_tree_matrix(prover_data._h, i)) for i in range(k)]
        rows = np.empty(sum(widths), dtype=np.uint32)
        depth = prover_data.depth
        path = np.empty((max(depth, 1), 32), dtype=np.uint8)
        self.ctx.check(L.ts_mmcs_open_batch(self.ctx._h, prover_data._h, query_index, _ptr(rows), _ptr(path)), "open_batch")
        rows = from_monty(rows)
        out, o = [], 0
        for w in widths:
            out.append(rows[o : o + w].copy())
            o += w
        return out, path[:depth].copy()

    def verify_batch(self, heights: Sequence[int], opened_values: Sequence[np.ndarray], query_index: int,
                     proof: np.ndarray, root: bytes) -> None:
        """Raises TapStarkError on mismatch (reference: Result<(), Error>)."""
        k = len(heights)
        hs = (C.c_size_t * k)(*heights)
        ws = (C.c_size_t * k)(*[len(r) for r in opened_values])
        flat = np.ascontiguousarray(to_monty(np.concatenate([np.asarray(r, dtype=np.uint32) for r in opened_values])))
        proof = np.ascontiguousarray(proof, dtype=np.uint8)
        depth = proof.shape[0] if proof.size else 0
        pbuf = proof.reshape(-1) if proof.size else np.zeros(32, dtype=np.uint8)
        rbuf = np.frombuffer(root, dtype=np.uint8).copy()
        rc = self.ctx._L.ts_mmcs_verify_batch(hs, ws, k, self.layout, query_index, _ptr(flat), _ptr(pbuf), depth, _ptr(rbuf))
        if rc != 0:
            raise TapStarkError("verify_batch: root mismatch")

    def get_matrices(self, prover_data: ProverData) -> List[DeviceMatrix]:
        return prover_data._mats

    def get_max_height(self, prover_data: ProverData) -> int:
        return self.ctx._L.ts_tree_max_height(prover_data._h)


# ------------------------------------------------------------------------------------------- challenger
class BfChallenger:
    """BfChallenger<F, U32, Blake3Permutation, 16>; `ext` selects F = BabyBear^4 (the STARK config) or BabyBear."""

    def __init__(self, ext: bool = True, _handle=None):
        self._L = lib()
        self.ext = ext
        if _handle is None:
            _handle = C.c_void_p()
            self._L.ts_challenger_new(C.byref(_handle))
        self._h = _handle

    def clone(self) -> "BfChallenger":
        h = C.c_void_p()
        self._L.ts_challenger_clone(self._h, C.byref(h))
        return BfChallenger(self.ext, h)

    def observe(self, value) -> None:
        """value: 4 bytes, an int (LE u32 word), or a 32-byte commitment (observed as 8 words)."""
        if isinstance(value, int):
            value = value.to_bytes(4, "little")
        value = bytes(value)
        if len(value) == 32:
            self._L.ts_challenger_observe_digest(self._h, C.c_char_p(value))
        elif len(value) == 4:
            self._L.ts_challenger_observe(self._h, C.c_char_p(value))
        else:
            raise TapStarkError("observe: expected 4 or 32 bytes")

    def sample_base(self) -> int:
        return self._L.ts_challenger_sample_base(self._h)

    def sample_ext(self) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint32)
        self._L.ts_challenger_sample_ext(self._h, _ptr(out))
        return out

    def sample(self):
        return self.sample_ext() if self.ext else self.sample_base()

    def sample_bits(self, bits: int) -> int:
        return self._L.ts_challenger_sample_bits(self._h, bits, int(self.ext))

    def check_witness(self, bits: int, witness: int) -> bool:
        return bool(self._L.ts_challenger_check_witness(self._h, bits, witness, int(self.ext)))

    def grind(self, bits: int) -> int:
        w = C.c_uint32()
        if self._L.ts_challenger_grind(self._h, bits, int(self.ext), C.byref(w)) != 0:
            raise TapStarkError("failed to find witness")  # basic/src/challenger/mod.rs:101
        return w.value

    def __del__(self):
        try:
            self._L.ts_challenger_free(self._h)
        except Exception:
            pass


class TapTreeCommit:
    """basic::tcs::TCS::commit_polys on the device (first slice of SURVEY f2): TapLeaf hashes of the templated leaf scripts,
    the sorted-pair TapBranch tree and CompleteTaptree's leaf index table.  `segments` / `push_word` describe the leaf script
    template (include/tapstark.h: ts_taptree_commit); `rows` is a DeviceMatrix with one row of field words per leaf."""

    def __init__(self, ctx: Context, rows: DeviceMatrix, segments: Sequence[bytes], push_word: Sequence[int]):
        self.ctx = ctx
        n_push = len(push_word) + 1
        if len(segments) != n_push + 1:
            raise TapStarkError("taptree: need one segment more than pushes")
        blob = b"".join(segments)
        offs = [0]
        for s_ in segments:
            offs.append(offs[-1] + len(s_))
        seg_off = (C.c_size_t * len(offs))(*offs)
        pw = np.asarray(list(push_word) or [0], dtype=np.uint32)
        root = (C.c_uint8 * 32)()
        h = C.c_void_p()
        ctx.check(ctx._L.ts_taptree_commit(ctx._h, rows._h, C.c_char_p(blob), seg_off, _ptr(pw), n_push, root, C.byref(h)), "taptree_commit")
        self._h, self.root, self.n_leaves = h, bytes(root), rows.rows

    def leaf_indices(self) -> np.ndarray:
        out = np.empty(self.n_leaves, dtype=np.uint32)
        self.ctx.check(self.ctx._L.ts_taptree_leaf_indices(self.ctx._h, self._h, _ptr(out)), "taptree_leaf_indices")
        return out

    def level(self, level: int) -> np.ndarray:
        out = np.empty((self.n_leaves >> level, 32), dtype=np.uint8)
        self.ctx.check(self.ctx._L.ts_taptree_level(self.ctx._h, self._h, level, _ptr(out)), "taptree_level")
        return out

    def open(self, index: int) -> Tuple[List[bytes], int]:
        """(TaprootMerkleBranch of Merkle leaf `index`, leaf level first; its position among the TapTree's leaves)."""
        depth = self.n_leaves.bit_length() - 1
        path = np.zeros((max(depth, 1), 32), dtype=np.uint8)
        pos = C.c_uint32(0)
        self.ctx.check(self.ctx._L.ts_taptree_open(self.ctx._h, self._h, index, _ptr(path), C.byref(pos)), "taptree_open")
        return [bytes(path[l]) for l in range(depth)], int(pos.value)

    def free(self):
        if self._h:
            self.ctx._L.ts_taptree_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------- FRI
@dataclass
class FriConfig:
    """fri/src/config.rs:10-22"""

    log_blowup: int
    num_queries: int
    proof_of_work_bits: int
    mmcs: Blake3MerkleMmcs

    def blowup(self) -> int:
        return 1 << self.log_blowup


def fold_even_odd(ctx: Context, poly, beta) -> np.ndarray:
    """fri/src/fold_even_odd.rs:20-52 on host buffers.  poly: (2h,) base or (2h,4) ext canonical, bit-reversed
    evaluations; beta matching.  Returns the folded vector (h,) / (h,4)."""
    poly = np.asarray(poly, dtype=np.uint32)
    if poly.ndim == 1:
        h = poly.shape[0] // 2
        src = DeviceMatrix.from_canonical(ctx, poly.reshape(h, 2))
        dst = DeviceMatrix(ctx, _alloc(ctx, h, 1))
        ctx.check(ctx._L.ts_fri_fold_base(ctx._h, C.c_void_p(src.device_ptr), h, int(to_monty(int(beta))),
                                          C.c_void_p(dst.device_ptr)), "fri_fold_base")
        return dst.to_canonical().reshape(-1)
    h = poly.shape[0] // 2
    pin = to_monty(poly.reshape(-1, 4))
    out = np.empty((h, 4), dtype=np.uint32)
    b = to_monty(np.asarray(beta, dtype=np.uint32))
    ctx.check(ctx._L.ts_fri_fold_ext_host(ctx._h, _ptr(pin), h, _ptr(b), _ptr(out)), "fri_fold_ext_host")
    return from_monty(out)


def fold_matrix(ctx: Context, beta, m: DeviceMatrix) -> DeviceMatrix:
    """TwoAdicFriGenericConfig::fold_matrix (fri/src/two_adic_pcs.rs:116-147) device to device; m is h x 8."""
    h = m.rows * m.width // 8
    dst = DeviceMatrix(ctx, _alloc(ctx, h, 4))
    b = to_monty(np.asarray(beta, dtype=np.uint32))
    ctx.check(ctx._L.ts_fri_fold_ext(ctx._h, C.c_void_p(m.device_ptr), h, _ptr(b), C.c_void_p(dst.device_ptr)), "fri_fold_ext")
    return dst


def _alloc(ctx: Context, rows: int, width: int):
    h = C.c_void_p()
    ctx.check(ctx._L.ts_matrix_alloc(ctx._h, rows, width, C.byref(h)), "matrix_alloc")
    return h


@dataclass
class CommitPhaseResult:
    """fri/src/prover.rs:143-147"""

    commits: List[bytes]
    data: List[ProverData]
    final_poly: np.ndarray


def bf_commit_phase(config: FriConfig, inputs: Sequence[DeviceMatrix], challenger: BfChallenger,
                    keep_data: bool = True) -> CommitPhaseResult:
    """fri/src/prover.rs:93-141.  inputs: EF vectors (rows x 4), lengths strictly descending."""
    ctx = config.mmcs.ctx
    k = len(inputs)
    arr = (C.c_void_p * k)(*[m._h for m in inputs])
    max_rounds = max(int(inputs[0].rows).bit_length() - 1 - config.log_blowup, 0)
    commits = np.zeros((max(max_rounds, 1), 32), dtype=np.uint8)
    trees = (C.c_void_p * max(max_rounds, 1))()
    final = np.zeros(4, dtype=np.uint32)
    rounds = C.c_size_t()
    rc = ctx._L.ts_fri_commit_phase(ctx._h, arr, k, config.log_blowup, challenger._h, _ptr(commits),
                                    trees if keep_data else None, _ptr(final), C.byref(rounds))
    data = []
    if keep_data:
        for i in range(rounds.value):
            if trees[i]:
                data.append(ProverData(ctx, C.c_void_p(trees[i]), [], True))
    ctx.check(rc, "fri_commit_phase")
    return CommitPhaseResult([commits[i].tobytes() for i in range(rounds.value)], data, final)


# ------------------------------------------------------------------------------------------- PCS
@dataclass(frozen=True)
class TwoAdicMultiplicativeCoset:
    """[MEM] p3_commit::TwoAdicMultiplicativeCoset { log_n, shift }"""

    log_n: int
    shift: int = 1

    def size(self) -> int:
        return 1 << self.log_n


class TwoAdicFriPcs:
    """fri/src/two_adic_pcs.rs:37-52, Pcs impl :197-258 (commit side)."""

    def __init__(self, dft: GpuDft, mmcs: Blake3MerkleMmcs, fri: FriConfig):
        self.dft, self.mmcs, self.fri = dft, mmcs, fri
        self.ctx = dft.ctx

    def natural_domain_for_degree(self, degree: int) -> TwoAdicMultiplicativeCoset:
        log_n = degree.bit_length() - 1
        if 1 << log_n != degree:
            raise TapStarkError("log2_strict_usize: not a power of two")
        return TwoAdicMultiplicativeCoset(log_n, 1)

    def commit(self, evaluations: Sequence[Tuple[TwoAdicMultiplicativeCoset, DeviceMatrix]]) -> Tuple[bytes, ProverData]:
        k = len(evaluations)
        for dom, ev in evaluations:
            if dom.size() != ev.rows:
                raise TapStarkError("assertion failed: domain.size() == evals.height()")  # two_adic_pcs.rs:234
        arr = (C.c_void_p * k)(*[ev._h for _, ev in evaluations])
        shifts = to_monty(np.array([dom.shift for dom, _ in evaluations], dtype=np.uint32))
        root = (C.c_uint8 * 32)()
        h = C.c_void_p()
        L = self.ctx._L
        self.ctx.check(L.ts_pcs_commit(self.ctx._h, arr, _ptr(shifts), k, self.fri.log_blowup, self.mmcs.layout, root,
                                       C.byref(h)), "pcs_commit")
        mats = [DeviceMatrix(self.ctx, C.c_void_p(L.ts_tree_matrix(h, i)), owned=False) for i in range(k)]
        return bytes(root), ProverData(self.ctx, h, mats, False)

    def commit_host(self, evaluations: Sequence[Tuple[TwoAdicMultiplicativeCoset, np.ndarray]]) -> Tuple[bytes, ProverData]:
        """Host-buffer form of `commit`: Montgomery-form numpy matrices in (as a Rust Vec<BabyBear> would be),
        H2D inside the call, 32-byte commitment out."""
        k = len(evaluations)
        mats = [np.ascontiguousarray(ev, dtype=np.uint32) for _, ev in evaluations]
        ptrs = (C.c_void_p * k)(*[m.ctypes.data for m in mats])
        rows = (C.c_size_t * k)(*[m.shape[0] for m in mats])
        widths = (C.c_size_t * k)(*[m.shape[1] for m in mats])
        shifts = to_monty(np.array([dom.shift for dom, _ in evaluations], dtype=np.uint32))
        root = (C.c_uint8 * 32)()
        h = C.c_void_p()
        L = self.ctx._L
        self.ctx.check(L.ts_pcs_commit_host(self.ctx._h, ptrs, rows, widths, _ptr(shifts), k, self.fri.log_blowup,
                                            self.mmcs.layout, root, C.byref(h)), "pcs_commit_host")
        dm = [DeviceMatrix(self.ctx, C.c_void_p(L.ts_tree_matrix(h, i)), owned=False) for i in range(k)]
        return bytes(root), ProverData(self.ctx, h, dm, False)

    def get_evaluations_on_domain(self, prover_data: ProverData, idx: int, domain: TwoAdicMultiplicativeCoset) -> np.ndarray:
        if domain.shift != GENERATOR:
            raise TapStarkError("assertion failed: domain.shift == Val::generator()")  # two_adic_pcs.rs:254
        w = prover_data._mats[idx].width
        out = np.empty((domain.size(), w), dtype=np.uint32)
        self.ctx.check(self.ctx._L.ts_pcs_get_evaluations_on_domain(self.ctx._h, prover_data._h, idx, domain.size(),
                                                                    _ptr(out)), "get_evaluations_on_domain")
        return from_monty(out)

    def dot_ext_powers(self, m: DeviceMatrix, alpha) -> DeviceMatrix:
        a = to_monty(np.asarray(alpha, dtype=np.uint32))
        h = C.c_void_p()
        self.ctx.check(self.ctx._L.ts_dot_ext_powers(self.ctx._h, m._h, _ptr(a), C.byref(h)), "dot_ext_powers")
        return DeviceMatrix(self.ctx, h)


# ------------------------------------------------------------------------------------------- open / prove (f1, f4)
def _ef_mul(a, b):
    """BabyBear^4 product on canonical Python ints (transcript-side arithmetic: a few hundred per proof)."""
    r = [0] * 7
    for i in range(4):
        for j in range(4):
            r[i + j] = (r[i + j] + int(a[i]) * int(b[j])) % P
    return [(r[i] + 11 * r[i + 4]) % P if i < 3 else r[i] for i in range(4)]


def _ef_pow(a, e: int):
    r, b = [1, 0, 0, 0], [int(x) for x in a]
    while e:
        if e & 1:
            r = _ef_mul(r, b)
        b = _ef_mul(b, b)
        e >>= 1
    return r


@dataclass
class BatchOpening:
    """[MEM] p3_fri::BatchOpening as used at fri/src/two_adic_pcs.rs:408-411"""

    opened_values: List[np.ndarray]
    opening_proof: np.ndarray


@dataclass
class BfQueryProof:
    """fri/src/proof.rs:23-33"""

    input_proof: List[BatchOpening]
    commit_phase_openings: List[Tuple[List[np.ndarray], np.ndarray]]


@dataclass
class FriProof:
    """fri/src/proof.rs:13-21"""

    commit_phase_commits: List[bytes]
    query_proofs: List[BfQueryProof]
    final_poly: np.ndarray
    pow_witness: int


def bf_prove(config: FriConfig, inputs: Sequence[DeviceMatrix], challenger: BfChallenger, open_input) -> FriProof:
    """fri/src/prover.rs:19-67: commit phase, proof-of-work grinding, query phase.
    open_input(query_index) -> input proof.  (The reference's query_times_index selects one of its num_queries
    Taptrees; a Merkle commitment has a single tree.)"""
    log_max_height = int(inputs[0].rows).bit_length() - 1
    cp = bf_commit_phase(config, inputs, challenger)
    pow_witness = challenger.grind(config.proof_of_work_bits)
    queries = []
    for _ in range(config.num_queries):
        index = challenger.sample_bits(log_max_height)
        openings = [config.mmcs.open_batch(index >> i >> 1, data) for i, data in enumerate(cp.data)]  # prover.rs:69-90
        queries.append(BfQueryProof(open_input(index), openings))
    return FriProof(cp.commits, queries, cp.final_poly, pow_witness)


def _pcs_open(self, rounds, challenger: BfChallenger):
    """TwoAdicFriPcs::open (fri/src/two_adic_pcs.rs:260-419) on the device-resident LDEs.
    rounds: [(ProverData, [[point, ...] per matrix])], points = canonical BabyBear^4 (4 ints).
    Returns (opened_values[round][matrix][point] -> (width, 4) canonical array, FriProof)."""
    ctx, L, b = self.ctx, self.ctx._L, self.fri.log_blowup
    alpha = [int(x) for x in challenger.sample()]  # :312
    mats_and_points = [(self.mmcs.get_matrices(data), points) for data, points in rounds]
    heights = [m.rows for mats, _ in mats_and_points for m in mats]
    log_global_max_height = max(heights).bit_length() - 1
    # inverse denominators per unique point, for the largest height opened at it (:677-720)
    max_lh = {}
    for mats, points in mats_and_points:
        for m, pts in zip(mats, points):
            for z in pts:
                key = tuple(int(x) for x in z)
                max_lh[key] = max(max_lh.get(key, 0), m.rows.bit_length() - 1)
    inv = {}
    for key, lh in max_lh.items():
        zm = to_monty(np.array(key, dtype=np.uint32))
        h = C.c_void_p()
        ctx.check(L.ts_inv_denoms(ctx._h, lh, _ptr(zm), C.byref(h)), "inv_denoms")
        inv[key] = DeviceMatrix(ctx, h)
    reduced, num_reduced, opened = {}, {}, []
    for mats, points in mats_and_points:
        opened_round = []
        for m, pts in zip(mats, points):
            lh = m.rows.bit_length() - 1
            if lh not in reduced:
                ro = DeviceMatrix(ctx, _alloc(ctx, m.rows, 4))
                ctx.check(L.ts_matrix_zero(ctx._h, ro._h), "matrix_zero")
                reduced[lh], num_reduced[lh] = ro, 0
            dot = self.dot_ext_powers(m, alpha) if pts else None  # sum_i alpha^i p_i[X], shared by all points (:375)
            opened_mat = []
            for z in pts:
                key = tuple(int(x) for x in z)
                zm = to_monty(np.array(key, dtype=np.uint32))
                ys_m = np.empty((m.width, 4), dtype=np.uint32)
                ctx.check(L.ts_interpolate_low_coset(ctx._h, m._h, m.rows >> b, _ptr(zm), inv[key]._h, _ptr(ys_m)),
                          "interpolate_low_coset")  # :358-369
                ys = from_monty(ys_m)
                apo = _ef_pow(alpha, num_reduced[lh])
                rys, ap = [0, 0, 0, 0], [1, 0, 0, 0]
                for y in ys:  # dot_product(alpha.powers(), ys)
                    t = _ef_mul(ap, y)
                    rys = [(rys[k] + t[k]) % P for k in range(4)]
                    ap = _ef_mul(ap, alpha)
                apo_m, rys_m = to_monty(np.array(apo, dtype=np.uint32)), to_monty(np.array(rys, dtype=np.uint32))
                ctx.check(L.ts_reduce_opening_acc(ctx._h, dot._h, inv[key]._h, _ptr(apo_m), _ptr(rys_m), reduced[lh]._h),
                          "reduce_opening_acc")  # :371-381
                num_reduced[lh] += m.width
                opened_mat.append(ys)
            opened_round.append(opened_mat)
        opened.append(opened_round)
    fri_input = [reduced[lh] for lh in sorted(reduced, reverse=True)]  # :389

    def open_input(index):
        out = []
        for data, _ in rounds:
            log_max = self.mmcs.get_max_height(data).bit_length() - 1
            vals, proof = self.mmcs.open_batch(index >> (log_global_max_height - log_max), data)  # :399-413
            out.append(BatchOpening(vals, proof))
        return out

    proof = bf_prove(self.fri, fri_input, challenger, open_input)
    return opened, proof


TwoAdicFriPcs.open = _pcs_open


def _pcs_open_bytes(self, rounds, challenger: BfChallenger) -> bytes:
    """Pcs::open through ONE C-ABI call (ts_pcs_open): same arguments as `open`, returns the postcard bytes of
    `(OpenedValues, FriProof)` -- decode with proofio.decode_opening."""
    ctx, L = self.ctx, self.ctx._L
    k = len(rounds)
    trees = (C.c_void_p * k)(*[data._h for data, _ in rounds])
    counts, pts = [], []
    for data, points in rounds:
        mats = self.mmcs.get_matrices(data)
        if len(points) != len(mats):
            raise TapStarkError("open: one point list per committed matrix")
        for p_ in points:
            counts.append(len(p_))
            pts += [[int(x) for x in z] for z in p_]
    n_points = (C.c_size_t * max(len(counts), 1))(*counts)
    pm = to_monty(np.array(pts, dtype=np.uint32).reshape(-1)) if pts else np.zeros(4, dtype=np.uint32)
    out, n = C.c_void_p(), C.c_size_t()
    ctx.check(L.ts_pcs_open(ctx._h, trees, k, n_points, _ptr(pm), self.fri.log_blowup, self.fri.num_queries,
                            self.fri.proof_of_work_bits, challenger._h, C.byref(out), C.byref(n)), "pcs_open")
    try:
        return C.string_at(out, n.value)
    finally:
        L.ts_bytes_free(out)


TwoAdicFriPcs.open_bytes = _pcs_open_bytes
