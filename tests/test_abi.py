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
