import os, sys, json, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from __graft_entry__ import load_pkg
ts = load_pkg()
ctx = ts.Context(0)
mmcs = ts.Blake3MerkleMmcs(ctx)
for logn in (12, 16, 20, 24):
    b = 2
    cfg = ts.FriConfig(b, 16, 8, mmcs)
    # low-degree codeword: LDE of random EF poly of degree < 2^(logn-b)
    n = 1 << (logn - b)
    t = torch.randint(0, ts.P, (n, 4), dtype=torch.int32, device="cuda")
    ev = ts.DeviceMatrix.wrap_device(ctx, t.data_ptr(), n, 4, keepalive=t)
    lde = ts.GpuDft(ctx).coset_lde_batch(ev, b, 1, committed_order=True)
    for rep in range(3):
        ch = ts.BfChallenger()
        ctx.synchronize(); t0 = time.perf_counter()
        res = ts.bf_commit_phase(cfg, [lde], ch, keep_data=False)
        ctx.synchronize(); t1 = time.perf_counter()
    print(json.dumps({"log_len": logn, "rounds": len(res.commits), "ms": round((t1 - t0) * 1e3, 3), "us_per_round": round((t1 - t0) * 1e6 / len(res.commits), 1)}))
