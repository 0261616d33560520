"""The N>1 path on real GPUs (needs >= 2 devices; skipped otherwise): NCCL ranks of tap-stark_b200/parallel.py with the
three forms of the re-shard (copy-engine peer copies, NCCL all-to-all, fused LDE + peer stores); all must reproduce the
single-process oracle transcript bit for bit."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

from test_sharded_gloo import free_port, oracle_transcript

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def n_gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def run_world(tmp_path, world, log_n, width, b, extra_env):
    out = tmp_path / "res"
    env = dict(os.environ, **extra_env)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(ROOT / "tests" / "dist_worker_gpu.py"), str(out), str(log_n), str(width), str(b)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return [json.loads(Path(f"{out}.{k}").read_text()) for k in range(world)]


@pytest.mark.parametrize("mode", ["ce", "nccl", "fused"])
def test_sharded_gpu_matches_oracle(tmp_path, orc, mode):
    """ce: copy-engine peer copies into IPC-mapped receive buffers (default); nccl: all-to-all; fused: the last butterfly
    pass stores into the peers' buffers.  Resident shard, per-chunk host panels and the row-major host trace (strided
    windows) all give the single-process oracle transcript."""
    if n_gpus() < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n_gpus() >= 4 else 2
    log_n, width, b = 18, 64, 2  # 16 or 32 columns per rank -> 2 or 4 chunks of 8
    results = run_world(tmp_path, world, log_n, width, b, {"TS_RESHARD": mode})
    root, commits, final = oracle_transcript(orc, log_n, width, b)
    for res in results:
        assert res["reshard"] == mode
        assert res["peer_buffers"] == (mode != "nccl")
        for name in ("resident", "resident_again", "host_panels", "host_full"):
            assert res[name]["root"] == root, name
            assert res[name]["commits"] == commits, name
            assert res[name]["final_poly"] == final, name
