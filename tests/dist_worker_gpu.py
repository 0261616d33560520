"""Worker of tests/test_sharded_gpu.py: one rank of the sharded prover on its own GPU (NCCL) with the CUDA library.
Runs the resident and the host-panel (end-to-end) forms of the step; TEST-ONLY."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    out_path, log_n, width, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    from __graft_entry__ import load_pkg
    from oracle import oracle as orc

    ts = load_pkg()
    from tapstark_b200.parallel import ShardedProver

    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = ts.Context(local, stream=stream.cuda_stream)
        trace = orc.splitmix_matrix(5, 1 << log_n, width)  # canonical, full trace (every rank derives its shard)
        prover = ShardedProver(ts, ctx, rank, world, b, torch.device("cuda", local))
        shard = ts.to_monty(np.ascontiguousarray(trace[:, prover.owned_columns(width)]))  # chunk-major column ownership
        shard_t = torch.from_numpy(shard.view(np.int32).copy()).cuda()
        prover.REPLICATE_BELOW = int(os.environ.get("TS_REPLICATE_BELOW", str(1 << 12)))
        out = {}
        full_host = torch.from_numpy(ts.to_monty(trace).view(np.int32).copy()).pin_memory()
        for name in ("resident", "resident_again", "host_panels", "host_full"):
            if name == "host_panels":
                res = prover.commit_and_fri(None, host_panels=prover.host_panels(shard_t))
            elif name == "host_full":
                res = prover.commit_and_fri(None, host_full=full_host)
            else:
                res = prover.commit_and_fri(shard_t)
            out[name] = {"root": res["root"].hex(), "commits": [c.hex() for c in res["commits"]], "final_poly": res["final_poly"]}
        out["reshard"] = prover.reshard_mode()
        out["peer_buffers"] = any(v is not None for v in getattr(prover, "_p2p_cache", {}).values())
        prover.close()
        torch.cuda.synchronize()
    Path(f"{out_path}.{rank}").write_text(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
