"""The N>1 path on CPU: world_size-2 (and 4) gloo runs of tap-stark_b200/parallel.py against the emulated
kernel library; every rank must reproduce the single-process oracle transcript bit for bit (same root, same
FRI layer commitments, same final polynomial)."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_world(tmp_path, world, log_n, width, b):
    from emul.build_emul import build

    build()
    out = tmp_path / "res"
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "dist_worker.py"), str(out), str(log_n),
                                       str(width), str(b)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    logs = []
    for p in procs:
        o, _ = p.communicate(timeout=600)
        logs.append(o.decode())
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    return [json.loads(Path(f"{out}.{r}").read_text()) for r in range(world)]


def oracle_transcript(orc, log_n, width, b):
    trace = orc.splitmix_matrix(5, 1 << log_n, width)
    lde = orc.pcs_lde_committed(trace, b)
    tree = orc.mmcs_commit([lde])
    ch = orc.BfChallenger()
    ch.observe_digest(tree.root)
    alpha = ch.sample_ef()
    res = orc.fri_commit_phase([orc.dot_ext_powers(lde, alpha)], b, ch)
    assert res["ok"]
    return tree.root.hex(), [c.hex() for c in res["commits"]], res["final_poly"].tolist()


@pytest.mark.parametrize("world,log_n,width,b", [(2, 10, 8, 2), (4, 11, 8, 1), (2, 6, 4, 2), (2, 9, 64, 1)])
def test_sharded_matches_oracle(tmp_path, orc, world, log_n, width, b):
    results = run_world(tmp_path, world, log_n, width, b)
    root, commits, final = oracle_transcript(orc, log_n, width, b)
    for res in results:
        assert res["root"] == root
        assert res["commits"] == commits
        assert res["final_poly"] == final
