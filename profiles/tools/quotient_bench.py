#!/usr/bin/env python3
"""Times ts_quotient_values (csrc/quotient.cuh) on device-resident committed LDEs: CUDA events through the context's
kernel statistics, 3 warm-up + 5 timed launches, inputs larger than L2 for the wide case.  One JSON line per case.
usage (GPU box, repo root): python profiles/tools/quotient_bench.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import airs  # noqa: E402
from __graft_entry__ import load_pkg  # noqa: E402

ts = load_pkg()
from importlib import import_module  # noqa: E402

st = import_module("tapstark_b200.stark")
ctx = ts.Context(0)
mmcs = ts.Blake3MerkleMmcs(ctx)


def run(name, air, log_n, log_blowup, n_pub):
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mmcs, ts.FriConfig(log_blowup, 16, 8, mmcs))
    n, w = 1 << log_n, air.width()
    t = torch.randint(0, ts.P, (n, w), dtype=torch.int32, device="cuda")  # timing only: the constraints need not hold
    ev = ts.DeviceMatrix.wrap_device(ctx, t.data_ptr(), n, w, keepalive=t)
    root, data = pcs.commit([(pcs.natural_domain_for_degree(n), ev)])
    log_qd = st.get_log_quotient_degree(air, n_pub)
    prog, consts = st.compile_program(st.get_symbolic_constraints(air, n_pub))
    for i in range(8):
        if i == 3:
            ctx.synchronize(); ctx.set_profiling(True); ctx.reset_stats()
        chunks = st.quotient_values(pcs, data, air, list(range(n_pub)), log_n, log_qd, [3, 1, 4, 1])
        ctx.synchronize()
        for c in chunks:
            c.free()
    ms = ctx.stats()["misc"]["ms"] / 5
    ctx.set_profiling(False)
    m = n << log_qd
    alg = m * (2 * w * 4 + 16)
    print(json.dumps({"case": name, "log_n": log_n, "width": w, "log_quotient_degree": log_qd, "instructions": int(prog.shape[0]),
                      "ms": round(ms, 4), "quotient_points_per_s": m / (ms * 1e-3), "alg_bytes": alg,
                      "achieved_GBs": alg / (ms * 1e-3) / 1e9}))
    data.free()


run("fibonacci 2^22 x 2, blowup 4", airs.FibonacciAir(), 22, 2, 3)
run("mul_air degree 3, 20 triples (width 60), 2^20, blowup 4", airs.MulAir(3, 20), 20, 2, 0)
run("mul_air degree 5, 20 triples (width 60), 2^19, blowup 8", airs.MulAir(5, 20), 19, 3, 0)
run("C4 scale: mul_air degree 3, 66 triples (width 198), 2^21, blowup 4", airs.MulAir(3, 66), 21, 2, 0)
