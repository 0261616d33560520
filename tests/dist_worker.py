"""Worker of tests/test_sharded_gloo.py: one rank of the sharded prover on CPU tensors (gloo) against the
emulated kernel library.  TEST-ONLY."""
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
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from emul.build_emul import OUT

    from __graft_entry__ import load_pkg
    from oracle import oracle as orc

    ts = load_pkg()
    ts.load_library(OUT, allow_emulated=True)
    from tapstark_b200.parallel import ShardedProver

    ctx = ts.Context(0)
    trace = orc.splitmix_matrix(5, 1 << log_n, width)  # canonical, full trace (every rank derives its shard)
    prover = ShardedProver(ts, ctx, rank, world, b, torch.device("cpu"))
    shard = ts.to_monty(np.ascontiguousarray(trace[:, prover.owned_columns(width)]))  # chunk-major column ownership
    shard_t = torch.from_numpy(shard.view(np.int32).copy())
    prover.REPLICATE_BELOW = int(os.environ.get("TS_REPLICATE_BELOW", "64"))  # exercise the row-sharded FRI rounds
    res = prover.commit_and_fri(shard_t)
    res["root"] = res["root"].hex()
    res["commits"] = [c.hex() for c in res["commits"]]
    Path(f"{out_path}.{rank}").write_text(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
