"""No-GPU checks of the boundary: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/tapstark.h declares; the product loader has no CPU path."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def built():
    from __graft_entry__ import build_device

    return build_device()


def declared_symbols():
    text = (ROOT / "include" / "tapstark.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ts_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_the_mirror_binds():
    from __graft_entry__ import load_pkg

    assert declared_symbols() == load_pkg().ABI_SYMBOLS


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(str(built))
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/tapstark.h but not exported"
    lib.ts_is_device_build.restype = ctypes.c_int
    assert lib.ts_is_device_build() == 1


def test_sass_is_sm100a(built):
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", str(built)], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from __graft_entry__ import load_pkg

    ts = load_pkg()
    ts.load_library(built)
    with pytest.raises(ts.TapStarkError):
        ts.Context(0)
    ts._lib = None


def test_product_never_imports_oracle():
    for p in (ROOT / "tap-stark_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".h", ".cpp", ".rs"}:
            txt = p.read_text()
            assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), p
            assert "tapstark_oracle" not in txt and "liboracle" not in txt, p


def _c_params(decl: str):
    """number of parameters of a C prototype string `name(...)`"""
    inner = decl[decl.index("(") + 1: decl.rindex(")")].strip()
    return 0 if inner in ("", "void") else inner.count(",") + 1


def test_rust_shim_binds_declared_entry_points_with_matching_arity():
    """The Rust shim cannot be compiled here (no rustc): at least every `extern "C"` item of its `sys` module must be an
    entry point include/tapstark.h declares, with the same number of parameters, and the trait impls the reference needs
    (bf_pcs.rs:19-88, bf_mmcs.rs:17-68, TwoAdicSubgroupDft) must be present."""
    hdr = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "tapstark.h").read_text(), flags=re.S)
    c_decl = {m.group(1): _c_params(m.group(0)) for m in re.finditer(r"\b(ts_[a-z0-9_]+)\s*\([^;{]*\)\s*;", hdr, flags=re.S)}
    rs = (ROOT / "tap-stark_b200" / "rust" / "src" / "lib.rs").read_text()
    sys_mod = rs[rs.index("pub mod sys"): rs.index("pub const LAYOUT_P3_INJECT")]
    rust_decl = {m.group(1): _c_params(m.group(0)) for m in re.finditer(r"pub fn (ts_[a-z0-9_]+)\s*\([^;]*\)[^;]*;", sys_mod, flags=re.S)}
    assert len(rust_decl) >= 30
    for name, arity in rust_decl.items():
        assert name in c_decl, f"rust/src/lib.rs binds {name}, which include/tapstark.h does not declare"
        assert arity == c_decl[name], f"{name}: {arity} parameters in the Rust shim, {c_decl[name]} in the header"
    for needle in ("impl TwoAdicSubgroupDft<BabyBear> for GpuDft", "impl<T: DeviceElem> BFMmcs<T> for GpuBlake3Mmcs",
                   "impl Pcs<Challenge, GpuChallenger> for GpuTwoAdicFriPcs", "fn dft_batch", "fn idft_batch", "fn coset_dft_batch",
                   "fn coset_lde_batch", "fn open_batch", "fn verify_batch", "fn get_matrices", "fn get_evaluations_on_domain",
                   "fn open(", "fn verify(", "sys::ts_pcs_open", "sys::ts_pcs_commit_host", "sys::ts_dft_batch_host"):
        assert needle in rs, needle
    assert "ts_coset_lde_batch_host(c.0, as_u32(&mat.values), h, w, 0" not in rs  # round-1 bug: dft_batch as an LDE of 0 bits
