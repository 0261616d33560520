#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into a small CSV + markdown table that can be committed.
usage: summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01/name"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp_instr"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    table = []
    for r in rows[2:]:
        rec = {"kernel": r[idx["Kernel Name"]].split("(")[0]}
        for m, name in METRICS:
            if m in idx:
                rec[name] = f"{r[idx[m]]} {units[idx[m]]}".strip()
        stalls = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), r[i])
                  for i, h in enumerate(hdr)
                  if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
        stalls = sorted(stalls, key=lambda x: -float(x[1]) if x[1] not in ("", "nan", "-nan") else 0)[:5]
        rec["top_stalls_per_issue"] = ", ".join(f"{a}={float(b):.2f}" for a, b in stalls if b not in ("", "nan", "-nan"))
        table.append(rec)
    keys = ["kernel"] + [n for _, n in METRICS] + ["top_stalls_per_issue"]
    with open(out + ".csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        for rec in table:
            w.writerow(rec)
    with open(out + ".md", "w") as f:
        f.write(f"ncu --set full --clock-control none summary of `{rep}`\n\n")
        for rec in table:
            f.write(f"### {rec['kernel']}\n")
            for k in keys[1:]:
                f.write(f"- {k}: {rec.get(k, '')}\n")
            f.write("\n")
    print(f"wrote {out}.csv / .md ({len(table)} kernels)")


if __name__ == "__main__":
    main()
