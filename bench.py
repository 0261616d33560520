#!/usr/bin/env python3
"""bench.py -- committed BabyBear LDE elements / second on the BASELINE.json config
(2^22 x 256 trace, log_blowup 2: coset LDE -> Blake3 Merkle commit -> alpha-reduction -> FRI commit phase).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the oracle port of the reference's
                                                              # algorithms on all host cores, bounded sample

One step = one pass of the hot path over one synthetic trace.  `value` is measured with the trace resident
in HBM; `e2e` goes through the host-buffer C-ABI entry point (pinned host trace -> H2D inside the call ->
commitment + FRI commitments + final polynomial read back).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

P = 0x78000001
METRIC = "committed_babybear_lde_elems_per_s"
UNIT = "elems/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-rows", type=int, default=22)
    ap.add_argument("--width", type=int, default=256)
    ap.add_argument("--log-blowup", type=int, default=2)
    ap.add_argument("--cpu-log-rows", type=int, default=18, help="rows of the bounded CPU sample")
    ap.add_argument("--cpu-same-config", action="store_true",
                    help="reference arm: additionally time ONE step of the full --log-rows configuration (about 30 s of CPU work)")
    ap.add_argument("--seed", type=int, default=0, help="SplitMix seed of the synthetic trace (SURVEY 8d)")
    ap.add_argument("--no-self-check", action="store_true",
                    help="skip the in-process comparison of root / final_poly with the CPU oracle on the 2^cpu-log-rows sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return f"trace 2^{a.log_rows}x{a.width} BabyBear, log_blowup {a.log_blowup}: coset LDE + Blake3 Merkle commit + FRI commit phase"


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_pipeline(orc, trace, b):
    """The same step on the CPU oracle (oracle/ is test infrastructure; here it is the thing being timed as
    the reference arm / cpu_baseline, never part of the product path)."""
    lde = orc.pcs_lde_committed(trace, b)
    tree = orc.mmcs_commit([lde])
    ch = orc.BfChallenger()
    ch.observe_digest(tree.root)
    alpha = ch.sample_ef()
    fri_in = orc.dot_ext_powers(lde, alpha)
    res = orc.fri_commit_phase([fri_in], b, ch)
    assert res["ok"]
    return tree.root, res


def cpu_sample(a, steps, warmup, log_rows=None):
    from oracle import oracle as orc

    cores = orc.use_all_cores()  # torch.distributed.run exports OMP_NUM_THREADS=1: set the team explicitly
    log_rows = a.cpu_log_rows if log_rows is None else log_rows
    rows, w, b = 1 << log_rows, a.width, a.log_blowup
    trace = orc.splitmix_matrix(a.seed, rows, w)
    for _ in range(warmup):
        cpu_pipeline(orc, trace, b)
    t0 = time.perf_counter()
    for _ in range(steps):
        root, res = cpu_pipeline(orc, trace, b)
    dt = (time.perf_counter() - t0) / steps
    elems = (rows << b) * w
    return {
        "value": elems / dt,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"oracle C port (OpenMP, {cores} threads) of the same step on a 2^{log_rows}x{w} SplitMix trace (seed {a.seed}), "
                  f"log_blowup {b} ({elems} output elems, {dt:.2f} s/step); the reference itself is Rust and cannot be built here",
        "ms_per_step": dt * 1e3,
        "root": root.hex(),
        "final_poly": [int(x) for x in res["final_poly"]],
    }


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_sample(a, max(a.steps, 1), min(a.warmup, 1))
    line = {
        "impl": "reference",
        "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": cb["sample"]},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if a.cpu_same_config and a.cpu_log_rows != a.log_rows:
        full = cpu_sample(a, 1, 0, log_rows=a.log_rows)
        line["same_config"] = {"value": full["value"], "unit": UNIT, "ms_per_step": full["ms_per_step"], "steps": 1,
                               "cores": full["cores"], "root": full["root"], "final_poly": full["final_poly"],
                               "note": "ONE untimed-warm-up-free step of the full configuration; `value` above is the K-step sample"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ roofline
def algorithmic_bytes(n, w, b, digits):
    """Compulsory HBM bytes per step and kernel class (each kernel reads its inputs once, writes its outputs
    once; DESIGN.md 'Rooflines').  n rows, w columns, N = n << b."""
    N = n << b
    D = digits
    fri_rounds = max((N.bit_length() - 1) - b, 0)
    out = {
        # two-digit LDE = P1, P2 (n rows each), P4 (N rows) as passes and P3 as "lde_mid" (ntt_v4.cuh)
        "ntt_pass": (2 * 2 * n * w * 4 + 2 * N * w * 4) if D == 2 else ((D - 1) * 2 * n * w * 4 + (D - 1) * 2 * N * w * 4),
        "lde_mid": (n + N) * w * 4,
        "hash_leaves": N * w * 4 + N * 32 + sum((N >> (r + 1)) * 64 for r in range(fri_rounds)),
        "tree": 96 * (N - 1) + sum(96 * ((N >> (r + 1)) - 1) for r in range(fri_rounds)),
        "fold": sum(48 * (N >> (r + 1)) for r in range(fri_rounds)),
        "misc": N * w * 4 + N * 16,
    }
    out["lde_stage_5B_per_elem"] = (n + N) * w * 4
    return out


def split_digits(m, max_digit=11):
    D = (m + max_digit - 1) // max_digit
    return max(D, 1)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(a):
    import torch
    import torch.distributed as dist

    from __graft_entry__ import build_device, load_pkg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner to stdout when the first communicator is created; rank 0's stdout must
        # carry ONE JSON line, so fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    if rank == 0:
        build_device()
    if world > 1:
        dist.barrier()
    ts = load_pkg()
    ts.load_library(os.environ.get("TAPSTARK_LIB"))  # TAPSTARK_LIB: A/B builds of the same sources (experiments)
    # a dedicated (non-default) torch stream is handed to the library, so torch.cuda.Event timing, torch's own
    # kernels (input generation, NCCL) and the library's kernels are all ordered on ONE stream
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = ts.Context(local_rank, stream.cuda_stream)

    if world > 1:
        from tapstark_b200 import parallel as par  # column-sharded LDE / all-to-all / row-sharded hash + fold

        def make_runner(log_rows):
            return par.ShardedRunner(ts, ctx, log_rows, a.width, a.log_blowup, rank, world, seed=a.seed)
    else:
        def make_runner(log_rows):
            return SingleGpuRunner(ts, ctx, log_rows, a.width, a.log_blowup, seed=a.seed)

    # ---- self-check: the same pipeline on the bounded 2^cpu_log_rows sample of the same trace generator, against the
    # CPU oracle, in this process (at N > 1 the sharded path is what is checked; only rank 0 runs the oracle)
    self_check = None
    if not a.no_self_check:
        small = make_runner(a.cpu_log_rows)
        got = small.step_resident()
        small.close()
        if rank == 0:
            from oracle import oracle as orc

            orc.use_all_cores()
            ref_root, ref = cpu_pipeline(orc, orc.splitmix_matrix(a.seed, 1 << a.cpu_log_rows, a.width), a.log_blowup)
            self_check = {
                "sample": f"2^{a.cpu_log_rows}x{a.width}, seed {a.seed}, {world} GPU(s) vs oracle",
                "root_match": got["root"] == ref_root,
                "final_poly_match": [int(x) for x in got["final_poly"]] == [int(x) for x in ref["final_poly"]],
                "fri_commits_match": (got.get("commits") == ref["commits"]) if got.get("commits") is not None else None,
            }
            if not (self_check["root_match"] and self_check["final_poly_match"]):
                raise SystemExit(f"bench.py: self-check FAILED against the oracle: {self_check}")
    runner = make_runner(a.log_rows)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input throughput -------------------------------------------------------------------
    for _ in range(a.warmup):
        runner.step_resident()
    barrier()
    ctx.set_profiling(True)
    ctx.reset_stats()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    last = None
    for _ in range(a.steps):
        last = runner.step_resident()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1) / a.steps
    clocks = sampler.stop() if sampler else None
    phases = runner.prover.last_phases() if hasattr(runner, "prover") else None
    stats = ctx.stats()
    launches = ctx.total_launches()
    ctx.set_profiling(False)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    elems = (1 << (a.log_rows + a.log_blowup)) * a.width
    value = elems / (ms * 1e-3)

    # ---- end to end through the host-buffer ABI ------------------------------------------------------------
    e2e = None
    if not a.no_e2e:
        runner.prepare_host()
        for _ in range(3):
            runner.step_e2e()
        barrier()
        k = max(2, min(a.steps, 5))
        ev0.record(stream)
        last_e = None
        for _ in range(k):
            last_e = runner.step_e2e()
        ev1.record(stream)
        barrier()
        ms_e = ev0.elapsed_time(ev1) / k
        if world > 1:
            t = torch.tensor([ms_e], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e = float(t.item())
        e2e = {"value": elems / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": runner.h2d_bytes * world,
               "d2h_bytes_per_step": runner.d2h_bytes, "ms_per_step": ms_e,
               # the host-buffer call must produce the resident step's transcript (same trace)
               "same_result_as_resident": bool(last and last_e and last_e["root"] == last["root"]
                                               and list(last_e["final_poly"]) == list(last["final_poly"]))}
        runner.release_host()
        if world == 1 and hasattr(runner, "prepare_host"):
            # the same call on a PAGEABLE host trace (a Rust Vec that was not registered with ts_host_register): the library
            # gathers the column windows into page-locked bounce slots with host threads (tapstark.cu: stage_pageable_window)
            try:
                runner.prepare_host(pinned=False)
                runner.step_e2e()
                barrier()
                ev0.record(stream)
                for _ in range(2):
                    runner.step_e2e()
                ev1.record(stream)
                barrier()
                ms_p = ev0.elapsed_time(ev1) / 2
                e2e["pageable_source"] = {"value": elems / (ms_p * 1e-3), "unit": UNIT, "ms_per_step": ms_p, "steps": 2}
            except TypeError:
                pass
            runner.release_host()

    if rank == 0:
        peaks = {}
        pf = ROOT / "MEASURED_PEAKS.json"
        if pf.exists():
            peaks = json.loads(pf.read_text())
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if pf.exists() else "fallback 6650 GB/s (B200_PROFILING.md)"
        n = 1 << a.log_rows
        wl = a.width // world if world > 1 else a.width
        alg = algorithmic_bytes(n, wl if world > 1 else a.width, a.log_blowup, split_digits(a.log_rows))
        per_kind = {}
        for kind, s in stats.items():
            if s["launches"] == 0:
                continue
            t_ms = s["ms"] / a.steps
            per_kind[kind] = {"ms_per_step": t_ms, "launches_per_step": s["launches"] / a.steps,
                              "alg_GB_per_step": alg.get(kind, 0) / 1e9,
                              "achieved_GBs": (alg.get(kind, 0) / 1e9) / (t_ms * 1e-3) if t_ms > 0 else None}
        dom = max(per_kind, key=lambda k: per_kind[k]["ms_per_step"]) if per_kind else None
        traffic = None
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists() and dom:
            traffic = json.loads(tf.read_text()).get(dom)
        roof = None
        if dom:
            d = per_kind[dom]
            per_launch_bytes = alg[dom] / d["launches_per_step"]
            per_launch_s = d["ms_per_step"] * 1e-3 / d["launches_per_step"]
            roof = {"kernel": dom, "bound": "hbm", "achieved": per_launch_bytes / per_launch_s / 1e9, "peak": peak,
                    "unit": "GB/s", "frac": per_launch_bytes / per_launch_s / 1e9 / peak, "traffic": traffic,
                    "alg_bytes_per_launch": per_launch_bytes, "avg_launch_ms": per_launch_s * 1e3, "peak_source": peak_src,
                    "note": "rank 0's kernels; algorithmic bytes = compulsory I/O of that kernel class (DESIGN.md 4): every digit pass "
                            "reads and writes its whole matrix once. The NTT kernels are INT32-issue bound, not HBM bound (both integer "
                            "pipes ~57 % busy, issue slots 61 %, profiles/r02/a_ncu_v4.md); `lde_stage` below is the whole LDE against the "
                            "5 B per output element of SURVEY 8(d); Blake3 leaves run the ALU pipe at 74 % (profiles/r01/v11_ncu_full.md)"}
        lde_ms = sum(per_kind.get(k, {}).get("ms_per_step", 0) for k in ("ntt_pass", "lde_mid"))
        # INT32 view of the LDE (the bound that actually applies): butterflies per second of this rank's
        # shard against the butterfly issue rate measured on a B200 by profiles/tools/int_pipes.cu
        int32 = None
        try:
            rate = None
            for ln in (ROOT / "profiles" / "r01" / "int_pipes_b200_v2.jsonl").read_text().splitlines():
                rec = json.loads(ln)
                if rec.get("test") == "dif_butterfly":
                    rate = rec["lane_ops_per_clk_per_sm_at_max_clock"]
            if rate and lde_ms and clocks and clocks.get("sm_mhz"):
                n_, w_ = 1 << a.log_rows, a.width // world
                bf = (0.5 * a.log_rows * n_ + 0.5 * a.log_rows * (n_ << a.log_blowup)) * w_  # inverse + 2^b forward size-n
                sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
                peak_bf = rate * sms * clocks["sm_mhz"] * 1e6
                int32 = {"butterflies_per_step": bf, "achieved_per_s": bf / (lde_ms * 1e-3), "peak_per_s": peak_bf,
                         "frac": bf / (lde_ms * 1e-3) / peak_bf,
                         "peak_source": "profiles/r01/int_pipes_b200_v2.jsonl dif_butterfly x SMs x sm_mhz (twiddle-only "
                                        "multiplies of the rounds are not counted as butterflies)"}
        except Exception:  # the extra view must never break the bench line
            int32 = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(a), "log_rows": a.log_rows, "width": a.width, "log_blowup": a.log_blowup,
                       "l2_policy": "inputs_larger_than_L2 (4 GiB trace, 16 GiB LDE; nothing is reused across steps)",
                       "parallelism": runner.parallelism, "fri_rounds": last["rounds"] if last else None},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof,
            "stages": per_kind, "phases_last_step_ms": phases,
            "lde_stage": {"alg_GB": alg["lde_stage_5B_per_elem"] / 1e9, "ms": lde_ms,
                          "achieved_GBs": alg["lde_stage_5B_per_elem"] / 1e9 / (lde_ms * 1e-3) if lde_ms else None,
                          "frac_of_hbm_peak": alg["lde_stage_5B_per_elem"] / 1e9 / (lde_ms * 1e-3) / peak if lde_ms else None,
                          "int32": int32},
            "result": {"root": last["root"].hex() if last else None, "final_poly": last["final_poly"] if last else None,
                       "trace": f"SplitMix64 seed {a.seed} (oracle.splitmix_matrix): identical at every N"},
            "self_check": self_check,
        }
        if not a.no_cpu_baseline and world == 1:
            cb = cpu_sample(a, 1, 0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    runner.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


class SingleGpuRunner:
    parallelism = "1 GPU"

    def __init__(self, ts, ctx, log_rows, width, log_blowup, seed=0):
        import torch

        self.ts, self.ctx, self.torch = ts, ctx, torch
        self.log_rows, self.width, self.b = log_rows, width, log_blowup
        n = 1 << log_rows
        # element (r, c) = SplitMix64((seed << 40) + r*width + c) mod p, generated on the device in
        # Montgomery form -- the matrix oracle.splitmix_matrix(seed, n, width) holds in canonical form
        self.trace_t = torch.empty((n, width), dtype=torch.int32, device="cuda")
        ctx.check(ctx._L.ts_fill_splitmix(ctx._h, self.trace_t.data_ptr(), n, width, seed, 0, width, 1), "fill_splitmix")
        self.trace = ts.DeviceMatrix.wrap_device(ctx, self.trace_t.data_ptr(), n, width, keepalive=self.trace_t)
        mm = ts.Blake3MerkleMmcs(ctx)
        self.pcs = ts.TwoAdicFriPcs(ts.GpuDft(ctx), mm, ts.FriConfig(log_blowup, 16, 8, mm))
        self.dom = self.pcs.natural_domain_for_degree(n)
        self.host = None
        self.h2d_bytes = n * width * 4
        self.d2h_bytes = 32 + 32 * log_rows + 16

    def _finish(self, root, data):
        ts = self.ts
        ch = ts.BfChallenger()
        ch.observe(root)
        alpha = ch.sample()
        lde = self.pcs.mmcs.get_matrices(data)[0]
        fri_in = self.pcs.dot_ext_powers(lde, alpha)
        res = ts.bf_commit_phase(self.pcs.fri, [fri_in], ch, keep_data=False)
        fri_in.free()
        data.free()
        return {"root": root, "rounds": len(res.commits), "final_poly": res.final_poly.tolist(), "commits": list(res.commits)}

    def step_resident(self):
        root, data = self.pcs.commit([(self.dom, self.trace)])
        return self._finish(root, data)

    def prepare_host(self, pinned=True):
        t = self.torch.empty((1 << self.log_rows, self.width), dtype=self.torch.int32, pin_memory=pinned)
        t.copy_(self.trace_t)
        self.torch.cuda.synchronize()
        self.host_t = t
        self.host = t.numpy().view(np.uint32)

    def step_e2e(self):
        root, data = self.pcs.commit_host([(self.dom, self.host)])
        return self._finish(root, data)

    def release_host(self):
        self.host = None
        self.host_t = None

    def close(self):
        self.trace = None
        self.trace_t = None
        self.ctx.trim()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
