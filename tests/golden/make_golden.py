#!/usr/bin/env python3
"""Generates tests/golden/golden.json -- run in the BUILD container only.

Sources of truth, none of which is oracle/tapstark_oracle.c:
  * the reference's own known-answer values (cited per entry);
  * oracle/pyref.py: mathematical definitions in Python big ints;
  * the `blake3` PyPI package (official implementation) for every hash.
The reference itself is Rust and cannot be built here (no rustc/cargo, un-vendored git deps), so no
fixture comes from running it.
"""
import json
import random
import sys
from pathlib import Path

import blake3 as _b3

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pyref  # noqa: E402

P = pyref.P


def H(b: bytes) -> bytes:
    return _b3.blake3(b).digest()


def main():
    rng = random.Random(20261018)
    g = {}

    # --- Blake3: reference KATs (scripts/src/hashes/blake3.rs:537-571) + official test-vector pattern
    kats = [
        {"src": "scripts/src/hashes/blake3.rs:537-556", "input_hex": (b"\x01\0\0\0" * 16).hex(),
         "hash": "86ca95aefdee3d969af9bcc78b48a5c1115be5d66cafc2fc106bbd982d820e70"},
        {"src": "scripts/src/hashes/blake3.rs:558-571", "input_hex": (b"\x01\0\0\0" * 15).hex(),
         "hash": "11b4167bd0184b9fc8b3474a4c29d08e801cbc1596b63a5ab380ce0fc83a15cd"},
    ]
    for k in kats:
        assert H(bytes.fromhex(k["input_hex"])).hex() == k["hash"], "reference KAT disagrees with blake3 package"
    pattern = []
    for n in [0, 1, 2, 3, 4, 31, 32, 33, 63, 64, 65, 127, 128, 129, 1023, 1024, 1025, 2048, 2049, 3072, 3073,
              4096, 4097, 5120, 5121, 6144, 6145, 7168, 7169, 8192, 8193, 16384, 31744, 102400]:
        data = bytes(i % 251 for i in range(n))
        pattern.append({"len": n, "hash": H(data).hex()})
    g["blake3"] = {"kats": kats, "pattern_mod251": pattern}

    # --- Challenger: reference golden (script_expr/src/challenger_expr.rs:278-296)
    ch = pyref.PyChallenger(H)
    ch.observe(b"\x01\x01\x01\x01")
    s0 = ch.sample_bb()
    ch.observe(b"\x01\x01\x01\x01")
    s1 = ch.sample_bb()
    assert s1 == 1103171332, s1
    # longer transcript: digests + ext samples, exercising buffer wrap and pop order
    ch2 = pyref.PyChallenger(H)
    script, outs = [], []
    for step in range(6):
        d = bytes(rng.randrange(256) for _ in range(32))
        ch2.observe_digest(d)
        script.append(["observe_digest", d.hex()])
        e = ch2.sample_ef()
        script.append(["sample_ef"])
        outs.append(e)
        if step % 2:
            wv = rng.randrange(1 << 32).to_bytes(4, "little")
            ch2.observe(wv)
            script.append(["observe", wv.hex()])
            b = ch2.sample_bb()
            script.append(["sample_bb"])
            outs.append(b)
    for _ in range(11):  # drain past one duplex without observing
        outs.append(ch2.sample_bb())
        script.append(["sample_bb"])
    g["challenger"] = {"ref_golden": {"src": "script_expr/src/challenger_expr.rs:278-296",
                                      "first": s0, "second": s1},
                       "script": script, "outputs": outs}

    # --- LDE by definition; parity unpinned w.r.t. the reference binary
    ldes = []
    for (log_n, w, b, shift) in [(0, 2, 2, 31), (1, 3, 1, 31), (3, 3, 2, 31), (4, 2, 1, 31), (5, 1, 2, 31),
                                 (3, 2, 2, pow(31, P - 2, P) * 31 % P), (4, 3, 3, 31 * pow(7, P - 2, P) % P)]:
        n = 1 << log_n
        rows = [[rng.randrange(P) for _ in range(w)] for _ in range(n)]
        ldes.append({"log_n": log_n, "width": w, "added_bits": b, "shift": shift, "evals": rows,
                     "committed": pyref.lde_committed_def(rows, b, shift)})
    g["lde"] = ldes

    # --- fold: definition + the reference's property (fri/src/fold_even_odd.rs:65-95) is tested separately
    folds = []
    for log_h in [0, 1, 2, 5]:
        h = 1 << log_h
        vals = [[rng.randrange(P) for _ in range(4)] for _ in range(2 * h)]
        beta = [rng.randrange(P) for _ in range(4)]
        folds.append({"log_h": log_h, "vals": vals, "beta": beta, "out": pyref.fold_def_ef(vals, beta)})
    g["fold_ef"] = folds

    # --- Merkle (single matrix): rows of canonical u32 -> root
    merk = []
    for (log_h, w) in [(0, 5), (1, 1), (3, 8), (4, 16), (2, 300)]:
        rows = [[rng.randrange(P) for _ in range(w)] for _ in range(1 << log_h)]
        root, layers = pyref.merkle_root_single(H, rows)
        merk.append({"rows": rows, "root": root.hex(), "leaf0": layers[0][0].hex()})
    g["merkle_single"] = merk

    # --- padded leaf layout: the reference's comment vectors (basic/src/tcs/mod.rs:540-602)
    g["padded_layout"] = {
        "src": "basic/src/tcs/mod.rs:594-602",
        "leaves": [[0, 0, 1, 0, 1, 2, 1], [1, 0, 1, 0, 1, 2, 1], [2, 2, 1, 2, 2, 1, 0], [1, 2, 1, 2, 2, 1, 0],
                   [2, 2, 2, 0, 1, 2, 1], [2, 2, 2, 0, 1, 2, 1], [1, 1, 0, 2, 2, 1, 0], [0, 1, 0, 2, 2, 1, 0]],
    }

    # --- a full commit phase by definition: EF codeword = LDE of a random low-degree EF polynomial
    log_n, b = 4, 2
    n = 1 << log_n
    rows = [[rng.randrange(P) for _ in range(4)] for _ in range(n)]  # 4 base columns = 1 EF column
    cw = pyref.lde_committed_def(rows, b, 31)
    chal = pyref.PyChallenger(H)
    folded, commits, betas = cw, [], []
    while len(folded) > (1 << b):
        leaves = [folded[2 * i] + folded[2 * i + 1] for i in range(len(folded) // 2)]
        root, _ = pyref.merkle_root_single(H, leaves)
        commits.append(root.hex())
        chal.observe_digest(root)
        beta = chal.sample_ef()
        betas.append(beta)
        folded = pyref.fold_def_ef(folded, beta)
    assert all(x == folded[0] for x in folded), "final layer not constant"
    g["commit_phase"] = {"log_n": log_n, "log_blowup": b, "evals": rows, "commits": commits, "betas": betas,
                         "final_poly": folded[0]}

    # --- row f3: quotient values of the reference's Fibonacci AIR (uni-stark/tests/fib_air.rs: 2^3 rows, log_blowup 2)
    # by definition (uni-stark/src/prover.rs:122-194; selectors [MEM] p3-commit selectors_on_coset), big ints only
    log_n, b = 3, 2
    n = 1 << log_n
    tr = [[0, 1]]
    for _ in range(1, n):
        tr.append([tr[-1][1], (tr[-1][0] + tr[-1][1]) % P])
    pis = [0, 1, tr[-1][1]]
    lde = pyref.lde_committed_def(tr, b, 31)
    root, _ = pyref.merkle_root_single(H, lde)
    chal = pyref.PyChallenger(H)
    chal.observe_digest(root)
    alpha = chal.sample_ef()
    tq = [lde[pyref.bitrev(i, log_n)] for i in range(n)]  # first n committed rows, re-bit-reversed: the trace on g*H_n
    w_n = pyref.two_adic_generator(log_n)
    w_n_inv = pow(w_n, P - 2, P)
    quotient = []
    for i in range(n):
        x = 31 * pow(w_n, i, P) % P
        zh = (pow(x, n, P) - 1) % P
        first = zh * pow((x - 1) % P, P - 2, P) % P
        last = zh * pow((x - w_n_inv) % P, P - 2, P) % P
        trans = (x - w_n_inv) % P
        l, nx = tq[i], tq[(i + 1) % n]
        cons = [first * (l[0] - pis[0]), first * (l[1] - pis[1]), trans * (l[1] - nx[0]), trans * (l[0] + l[1] - nx[1]),
                last * (l[1] - pis[2])]
        acc = [0, 0, 0, 0]
        for c in cons:  # accumulator = accumulator * alpha + constraint (uni-stark/src/folder.rs:60-64)
            acc = pyref.ef_mul(acc, alpha)
            acc[0] = (acc[0] + c) % P
        quotient.append(pyref.ef_scale(acc, pow(zh, P - 2, P)))
    g["stark_fibonacci"] = {"src": "uni-stark/tests/fib_air.rs:117-149, uni-stark/src/prover.rs:122-194", "log_n": log_n,
                            "log_blowup": b, "trace": tr, "public_values": pis, "trace_root": root.hex(), "alpha": alpha,
                            "quotient": quotient}

    out = Path(__file__).with_name("golden.json")
    out.write_text(json.dumps(g, separators=(",", ":")))
    print(f"wrote {out} ({out.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
