"""Run by tests/test_emulated_asan.py in a subprocess with libasan preloaded: the kernel sources, compiled by g++ with
-fsanitize=address against the SIMT emulator, execute the main paths once -- any out-of-bounds global or shared
access of a kernel shows up as an AddressSanitizer report (device allocations are heap blocks in the emulator).
compute-sanitizer is not available on the GPU pool, so this is the memory checker of the kernel logic.  TEST-ONLY."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import airs  # noqa: E402
import parity_cases as pc  # noqa: E402
from __graft_entry__ import load_pkg  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main(lib):
    ts = load_pkg()
    ts.load_library(lib, allow_emulated=True)
    ctx = ts.Context(0)
    pc.check_lde(ts, ctx, orc, 12, 3, 1)            # generic digit kernels, odd width
    pc.check_lde(ts, ctx, orc, 18, 12, 1)           # position-major passes + persistent lde_mid, ragged column slice
    pc.check_dot_ext_powers(ts, ctx, orc, 100, 70)  # ragged rows / columns of the warp-transposing kernels
    pc.check_dot_ext_powers(ts, ctx, orc, 64, 3)
    pc.check_mmcs(ts, ctx, orc, [(64, 5), (64, 3), (16, 9)], 0)
    pc.check_mmcs(ts, ctx, orc, [(32, 300)], 0)     # multi-chunk rows
    trace = airs.fibonacci_trace(0, 1, 1 << 5)
    pc.check_stark_prove_verify(ts, ctx, orc, airs.FibonacciAir(), trace, [0, 1, int(trace[-1, 1])], 2)
    air = airs.MulAir(degree=3, reps=2)
    pc.check_stark_prove_verify(ts, ctx, orc, air, airs.mul_trace(air, 1 << 4, 5), [], 2)
    pc.check_pcs_open_verify(ts, ctx, orc, [[(5, 6, 2)], [(5, 4, 1), (3, 8, 1)]], 1)
    print("asan cases ok")


if __name__ == "__main__":
    main(sys.argv[1])
