import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
from __graft_entry__ import load_pkg
ts = load_pkg()
if os.environ.get("TAPSTARK_LIB"): ts.load_library(os.environ["TAPSTARK_LIB"])
ctx = ts.Context(0)
n, w = 1 << 22, 256
t = torch.randint(0, 0x78000001, (n, w), dtype=torch.int32, device="cuda")
m = ts.DeviceMatrix.wrap_device(ctx, t.data_ptr(), n, w, keepalive=t)
dft = ts.GpuDft(ctx)
for i in range(3):
    if i == 1:
        ctx.set_profiling(True); ctx.reset_stats()
    o = dft.coset_lde_batch(m, 2, 31, committed_order=True); ctx.synchronize(); o.free()
st = ctx.stats()
print(json.dumps({k: round(v["ms"] / 2, 3) for k, v in st.items() if v["launches"]}))
