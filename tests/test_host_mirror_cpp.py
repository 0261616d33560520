"""The C++ host mirror (tap-stark_b200/host/tapstark.hpp) through tests/cpp/test_host_mirror.cpp, a program written
after the reference's own Rust tests for the path.  Built with g++ against the C ABI: once over the emulated build of
the kernel sources (no GPU), once over the CUDA library (-m gpu)."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "cpp" / "test_host_mirror.cpp"
OUT = ROOT / "tests" / "cpp" / "_build"


def _oracle_lib() -> Path:
    lib = ROOT / "oracle" / "_build" / "libtapstark_oracle.so"
    if not lib.exists():
        subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle")])
    return lib


def _build_and_run(lib: Path, name: str):
    OUT.mkdir(exist_ok=True)
    exe = OUT / name
    orc = _oracle_lib()
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-o", str(exe), str(SRC), str(lib), str(orc),
           f"-Wl,-rpath,{lib.parent}", f"-Wl,-rpath,{orc.parent}"]
    subprocess.check_call(cmd)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout
    return r.stdout


def test_host_mirror_emulated():
    from emul import build_emul
    out = _build_and_run(build_emul.build(), "host_mirror_emul")
    assert "11 passed" in out


@pytest.mark.gpu
def test_host_mirror_gpu():
    lib = ROOT / "tap-stark_b200" / "libtapstark_b200.so"
    assert lib.exists(), "CUDA library not built: run __graft_entry__.build()"
    out = _build_and_run(lib, "host_mirror_gpu")
    assert "11 passed" in out
