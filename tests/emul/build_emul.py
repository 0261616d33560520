"""Builds tests/emul/_build/libtapstark_emul.so: the kernel SOURCES of tap-stark_b200/csrc compiled by g++
against the fiber-based SIMT emulator (cuda_emul.h).  TEST-ONLY: lets the GPU-less build container check
kernel logic against the oracle.  The product package rejects this library unless a test passes
allow_emulated=True."""
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
OUT = HERE / "_build" / "libtapstark_emul.so"


def build(force: bool = False) -> Path:
    srcs = list((ROOT / "tap-stark_b200" / "csrc").glob("*")) + [HERE / "cuda_emul.h", HERE / "cuda_emul.cpp",
                                                                  ROOT / "include" / "tapstark.h"]
    if not force and OUT.exists() and all(s.stat().st_mtime <= OUT.stat().st_mtime for s in srcs):
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-DTS_EMULATE", "-w", f"-I{HERE}",
           "-o", str(OUT), "-x", "c++", str(ROOT / "tap-stark_b200" / "csrc" / "tapstark.cu"),
           "-x", "c++", str(HERE / "cuda_emul.cpp")]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
