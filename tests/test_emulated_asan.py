"""Memory check of the kernel logic without a GPU: the kernel sources compiled with AddressSanitizer against the SIMT
emulator (tests/emul/asan_cases.py).  compute-sanitizer is closed on the GPU pool; this is the substitute."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
EMUL = ROOT / "tests" / "emul"


def test_kernels_under_address_sanitizer():
    libasan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not libasan or not Path(libasan).exists():
        pytest.skip("libasan not available")
    out = EMUL / "_build" / "libtapstark_emul_asan.so"
    out.parent.mkdir(exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-fsanitize=address", "-fno-omit-frame-pointer",
                           "-DTS_EMULATE", "-w", f"-I{EMUL}", "-o", str(out), "-x", "c++",
                           str(ROOT / "tap-stark_b200" / "csrc" / "tapstark.cu"), "-x", "c++", str(EMUL / "cuda_emul.cpp")])
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    r = subprocess.run([sys.executable, str(EMUL / "asan_cases.py"), str(out)], env=env, capture_output=True, text=True,
                       timeout=1500)
    assert r.returncode == 0 and "asan cases ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    assert "ERROR: AddressSanitizer" not in r.stderr
