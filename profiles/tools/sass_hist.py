#!/usr/bin/env python3
"""Per-kernel SASS instruction histograms (static), split at BAR.SYNC (= per round of the NTT kernels).

    python profiles/tools/sass_hist.py tap-stark_b200/libtapstark_b200.so [substring ...] [--dump DIR]

--dump DIR also writes the raw `cuobjdump -sass` text of every selected kernel to DIR/<demangled-ish name>.sass.
"""
import collections
import re
import subprocess
import sys
from pathlib import Path


def kernels(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    for blk in txt.split("Function : ")[1:]:
        name, body = blk.split("\n", 1)
        yield name.strip(), body


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def ops_of(body):
    out = []
    for l in body.split("\n"):
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
            s = re.sub(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?", "", l)
            out.append(s.split()[0].rstrip(";"))
    return out


def main():
    args = sys.argv[1:]
    dump = None
    if "--dump" in args:
        i = args.index("--dump")
        dump = Path(args[i + 1])
        dump.mkdir(parents=True, exist_ok=True)
        del args[i:i + 2]
    lib, pats = args[0], args[1:]
    for name, body in kernels(lib):
        dn = demangle(name)
        if pats and not any(p in dn for p in pats):
            continue
        ops = ops_of(body)
        segs = [[]]
        for o in ops:
            segs[-1].append(o)
            if o.startswith("BAR"):
                segs.append([])
        print(f"== {dn}: {len(ops)} instructions, segments between barriers: {[len(s) for s in segs]}")
        for k, s in enumerate(segs):
            h = collections.Counter(x.split(".")[0] if x.split(".")[0] not in ("IMAD", "LDG", "STG", "LDS", "STS") else x for x in s)
            print(f"   seg {k}: " + ", ".join(f"{a} {b}" for a, b in h.most_common(14)))
        if dump:
            safe = re.sub(r"[^A-Za-z0-9_<>,]+", "_", dn)[:100]
            (dump / f"{safe}.sass").write_text("Function : " + name + "\n" + body)


if __name__ == "__main__":
    main()
