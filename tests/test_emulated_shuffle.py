"""Race check of the kernel logic without a GPU: a subset of the emulated parity tests re-run with TS_EMUL_SHUFFLE, which
resumes the threads of a block in a different pseudo-random order in every barrier interval -- a missing
__syncthreads / __syncwarp changes the result here instead of racing on the GPU."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_kernels_with_shuffled_thread_order():
    env = dict(os.environ, TS_EMUL_SHUFFLE="11")
    sel = "(lde_fast_path and 18-12-2) or (pipelined and 16) or stark_fibonacci or stark_mul or dot_ext_powers_blocks or mmcs"
    r = subprocess.run([sys.executable, "-m", "pytest", str(ROOT / "tests" / "test_parity_emulated.py"), "-q", "-x", "-k", sel],
                       env=env, capture_output=True, text=True, timeout=1500, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
