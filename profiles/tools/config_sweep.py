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


if __name__ == "__main__":
    bench("C2: 2^20 x 64, log_blowup 2", 20, 64, 2)
    bench("C4 trace shape: 2^21 x 200, log_blowup 2", 21, 200, 2)
    fri_sweep()
