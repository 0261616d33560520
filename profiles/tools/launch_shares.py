"""Kernel-class shares of one bench step from an `ncu --metrics gpu__time_duration.sum` launch list (CSV), next to the CUDA-event
stage times of a bench line.  usage: launch_shares.py LAUNCHES.csv [BENCH.json]
The ncu times are cold-cache and serialised: the SHARES are what must agree with the bench's own stage times."""
import csv
import json
import re
import sys

CLASSES = [
    ("lde_mid", r"ntt4::pass_\w*kernel<\d+, 0, 1,|lde_mid"), ("ntt_pass", r"ntt4::pass_\w*kernel<|ntt_pass"),  # <D, INV, MID, CSRC>
    ("hash_leaves", r"hash_rows|hash_leaves"), ("tree", r"tree_reduce|compress_inject|sponge_step|fri_tail|publish_subroot"),
    ("fold", r"fold_ext_kernel|fold_hash_kernel|fold_base"), ("misc", r"dot_rows|dot_ext|fill_|bitrev|gather|inv_denoms|bary|reduce_rows"),
]


def main():
    lines = open(sys.argv[1]).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    seq = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1e6) for r in rows]  # ms
    # the last step: from the last P1 launch (first pass of an LDE chunk whose predecessor is not a pass of the same LDE)
    dots = [i for i, (n, _) in enumerate(seq) if "dot_rows" in n]
    end = len(seq)
    first = dots[-2] + 1 if len(dots) >= 2 else 0
    # a step = [LDE chunks ... hash ... tree ... dot ... FRI]; it ends with the FRI tail after the LAST dot
    tails = [i for i, (n, _) in enumerate(seq) if "fri_tail" in n]
    prev_tail = max([t for t in tails if t < dots[-1]], default=-1)
    first, end = prev_tail + 1, tails[-1] + 1
    agg = {}
    for n, ms in seq[first:end]:
        cls = next((c for c, pat in CLASSES if re.search(pat, n)), "other")
        a = agg.setdefault(cls, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(v[1] for v in agg.values())
    bench = None
    if len(sys.argv) > 2:
        bench = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(f"{'class':12s} {'launches':>8s} {'ncu ms':>9s} {'share':>7s}" + ("   bench ms   share" if bench else ""))
    for cls, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = f"{cls:12s} {n:8d} {ms:9.3f} {100 * ms / total:6.1f}%"
        if bench and cls in bench["stages"]:
            b = bench["stages"][cls]["ms_per_step"]
            line += f"   {b:8.3f}  {100 * b / bench['ms_per_step']:5.1f}%"
        print(line)
    print(f"{'total':12s} {sum(v[0] for v in agg.values()):8d} {total:9.3f}" + (f"            {bench['ms_per_step']:8.3f}" if bench else ""))


if __name__ == "__main__":
    main()
