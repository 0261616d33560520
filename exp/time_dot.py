import os, sys, json
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from __graft_entry__ import load_pkg
ts = load_pkg()
if os.environ.get("TAPSTARK_LIB"): ts.load_library(os.environ["TAPSTARK_LIB"])
ctx = ts.Context(0)
n, w = 1 << 24, 256
t = torch.randint(0, 0x78000001, (n, w), dtype=torch.int32, device="cuda")
m = ts.DeviceMatrix.wrap_device(ctx, t.data_ptr(), n, w, keepalive=t)
mm = ts.Blake3MerkleMmcs(ctx)
pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mm, ts.FriConfig(2, 16, 8, mm))
alpha = np.array([5, 6, 7, 8], dtype=np.uint32)
for i in range(4):
    if i == 1:
        ctx.set_profiling(True); ctx.reset_stats()
    o = pcs.dot_ext_powers(m, alpha); ctx.synchronize(); o.free()
st = ctx.stats()
print(json.dumps({k: round(v["ms"] / 3, 3) for k, v in st.items() if v["launches"]}))
