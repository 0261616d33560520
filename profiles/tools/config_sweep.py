#!/usr/bin/env python3
"""Per-stage numbers for the other BASELINE.json configurations: C2 (2^20 x 64), the C4 trace shape
(2^21 x 200) -- both through bench.py, whose `stages` carry the per-kernel-class CUDA-event times -- and the C5 FRI
commit-phase sweep (codewords of 2^18 .. 2^26 BabyBear^4 elements).  One JSON line per case.
usage (GPU box, repo root): python profiles/tools/config_sweep.py"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def bench(name, log_rows, width, log_blowup):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--log-rows", str(log_rows), "--width", str(width),
                        "--log-blowup", str(log_blowup), "--steps", "5", "--warmup", "3", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        print(json.dumps({"case": name, "error": r.stderr[-400:]}))
        return
    d = json.loads(r.stdout.strip().splitlines()[-1])
    print(json.dumps({"case": name, "ms_per_step": d["ms_per_step"], "elems_per_s": d["value"],
                      "e2e_ms": d["e2e"]["ms_per_step"] if d.get("e2e") else None,
                      "stages_ms": {k: round(v["ms_per_step"], 3) for k, v in d["stages"].items()}, "clocks": d["clocks"]}))


def fri_sweep():
    import torch

    from __graft_entry__ import load_pkg

    ts = load_pkg()
    ctx = ts.Context(0)
    mmcs = ts.Blake3MerkleMmcs(ctx)
    for log_len, b in ((18, 1), (20, 2), (22, 2), (24, 2), (26, 2), (24, 4)):
        cfg = ts.FriConfig(b, 16, 8, mmcs)
        n = 1 << (log_len - b)
        t = torch.randint(0, ts.P, (n, 4), dtype=torch.int32, device="cuda")
        ev = ts.DeviceMatrix.wrap_device(ctx, t.data_ptr(), n, 4, keepalive=t)
        lde = ts.GpuDft(ctx).coset_lde_batch(ev, b, 1, committed_order=True)  # a low-degree codeword (prover.rs:130-134)
        best = None
        for rep in range(4):
            ch = ts.BfChallenger()
            ctx.synchronize()
            t0 = time.perf_counter()
            res = ts.bf_commit_phase(cfg, [lde], ch, keep_data=False)
            ctx.synchronize()
            dt = time.perf_counter() - t0
            if rep and (best is None or dt < best):
                best = dt
        print(json.dumps({"case": "C5 fri commit phase", "log_len": log_len, "log_blowup": b, "rounds": len(res.commits),
                          "ms": round(best * 1e3, 3), "ext_elems_per_s": (1 << log_len) / best}))
        lde.free()


def fold_rates():
    """fold_ext_kernel alone (fri/src/two_adic_pcs.rs:116-147): streaming read-2 / write-1, 48 bytes per output extension
    element, against the measured HBM peak; layers of 2^22 .. 2^26 elements (C5's largest is 1 GiB in, 512 MiB out)."""
    import ctypes as C

    import numpy as np
    import torch

    from __graft_entry__ import load_pkg

    ts = load_pkg()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = ts.Context(0, stream.cuda_stream)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    beta = ts.to_monty(np.array([3, 1, 4, 1], dtype=np.uint32))
    for log_len in (22, 24, 26):
        n = 1 << log_len
        src = torch.randint(0, ts.P, (n, 4), dtype=torch.int32, device="cuda")
        dst = torch.empty((n // 2, 4), dtype=torch.int32, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for rep in range(6):
            e0.record(stream)
            ctx.check(ctx._L.ts_fri_fold_ext(ctx._h, C.c_void_p(src.data_ptr()), n // 2, beta.ctypes.data_as(C.c_void_p), C.c_void_p(dst.data_ptr())), "fold")
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            if rep >= 2 and (best is None or ms < best):
                best = ms
        gbs = (n // 2) * 48 / (best * 1e-3) / 1e9
        print(json.dumps({"case": "fold_ext_kernel alone", "log_len_in": log_len, "ms": round(best, 4), "achieved_GBs": round(gbs, 1),
                          "hbm_peak_GBs": peak, "frac": round(gbs / peak, 3)}))
        del src, dst
    ctx.close()


def open_c4():
    """Pcs::open at the C4 shape through ts_pcs_open: per-kernel-class device times (the context's CUDA-event statistics)
    and the HBM rate of the dominant passes (alpha-reduction reads the LDE once per matrix, the barycentric sums read the
    low coset once per point)."""
    from __graft_entry__ import load_pkg

    ts = load_pkg()
    ctx = ts.Context(0)
    log_n, w, b, nq = 21, 200, 2, 28
    n = 1 << log_n
    mm = ts.Blake3MerkleMmcs(ctx)
    pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mm, ts.FriConfig(b, nq, 8, mm))
    trace = ts.DeviceMatrix.splitmix(ctx, 4, n, w)
    root_t, data_t = pcs.commit([(pcs.natural_domain_for_degree(n), trace)])
    chunks = [ts.DeviceMatrix.splitmix(ctx, 10 + k, n, 4) for k in range(4)]
    doms = [ts.TwoAdicMultiplicativeCoset(log_n, 31) for _ in range(4)]
    root_q, data_q = pcs.commit(list(zip(doms, chunks)))
    best, stats = None, None
    for rep in range(3):
        ch = ts.BfChallenger()
        ch.observe(root_t)
        ch.observe(root_q)
        zeta = [int(x) for x in ch.sample()]
        zn = [c * 3 % ts.P for c in zeta]
        ctx.set_profiling(True)
        ctx.reset_stats()
        ctx.synchronize()
        t0 = time.perf_counter()
        blob = pcs.open_bytes([(data_t, [[zeta, zn]]), (data_q, [[zeta]] * 4)], ch)
        dt = time.perf_counter() - t0
        st = ctx.stats()
        ctx.set_profiling(False)
        if rep and (best is None or dt < best):
            best, stats = dt, st
    N = n << b
    lde_bytes = N * w * 4 + 4 * N * 4 * 4
    print(json.dumps({"case": "C4 Pcs::open (ts_pcs_open): 2^21x200 trace at 2 points + 4 chunks of 2^21x4 at 1 point, 28 queries",
                      "ms": round(best * 1e3, 2), "proof_bytes": len(blob),
                      "stages_ms": {k: round(v["ms"], 3) for k, v in stats.items() if v["launches"]},
                      "launches": {k: v["launches"] for k, v in stats.items() if v["launches"]},
                      "lde_GB_read_once": round(lde_bytes / 1e9, 2)}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "open":
        open_c4()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fri":
        fri_sweep()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fold":
        fold_rates()
        sys.exit(0)
    bench("C2: 2^20 x 64, log_blowup 2", 20, 64, 2)
    bench("C4 trace shape: 2^21 x 200, log_blowup 2", 21, 200, 2)
    fri_sweep()
    fold_rates()
    open_c4()
