"""bench.py's reference arm runs without a GPU: its JSON line must carry the contract's keys, keep the metric/unit/config of the
GPU arm and use every core even when torch.distributed.run has exported OMP_NUM_THREADS=1 (round 1's N > 1 denominators were 16x
too small because it did not)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(extra_env, *args):
    env = dict(os.environ, **extra_env)
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-log-rows", "12", "--width", "16", *args], env=env, capture_output=True, text=True, timeout=600, check=True)
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_reference_arm_line():
    d = _run({"OMP_NUM_THREADS": "1"})
    assert d["impl"] == "reference" and d["metric"] == "committed_babybear_lde_elems_per_s" and d["unit"] == "elems/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["dtype"] == "u32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "sample" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["unit"] == "elems/s" and cb["value"] == d["value"] and cb["sample"]
    # OMP_NUM_THREADS=1 from the launcher must not reach the OpenMP team of the oracle
    usable = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    assert cb["cores"] == usable
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without output."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                          "--cpu-log-rows", "12", "--width", "16"], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip() == ""
