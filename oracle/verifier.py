"""Oracle restatement of the opening side of the PCS: TwoAdicFriPcs::open's reduced openings
(fri/src/two_adic_pcs.rs:260-419) and TwoAdicFriPcs::verify (:421-530) with the FRI verifier
(fri/src/verifier.rs:20-165).  TEST INFRASTRUCTURE ONLY (see oracle/tapstark_oracle.h).

Everything is canonical integers; extension elements are 4-lists / (.., 4) arrays.  Deliberately different
from the device path: opened values come from coefficient-form evaluation (inverse DFT + Horner), not from the
barycentric formula; reduced openings follow the reference's loop literally.
"""
from __future__ import annotations

import numpy as np

from . import oracle as orc
from . import pyref

P = pyref.P


def ef_pow(a, e):
    r, b = [1, 0, 0, 0], [int(x) for x in a]
    while e:
        if e & 1:
            r = pyref.ef_mul(r, b)
        b = pyref.ef_mul(b, b)
        e >>= 1
    return r


def ef_inv(a):
    o = np.zeros(4, dtype=np.uint32)
    orc.lib().or_ef_inv(orc._u32p(np.ascontiguousarray(a, dtype=np.uint32)), orc._u32p(o))
    return [int(x) for x in o]


def eval_matrix_at(lde_committed: np.ndarray, log_blowup: int, z) -> list:
    """p_c(z) for every column: the low coset (first n committed rows, bit-reversed) holds p on g*H_n
    (two_adic_pcs.rs:358-369).  Coefficients by inverse DFT, then Horner in the extension field."""
    n = lde_committed.shape[0] >> log_blowup
    low = orc.bit_reverse_rows(np.ascontiguousarray(lde_committed[:n]))  # natural order on g*H_n
    coeffs = orc.idft_batch(low).astype(object)  # c_k * g^k
    ginv = pow(31, P - 2, P)
    gk = 1
    for k in range(n):
        coeffs[k] = (coeffs[k] * gk) % P
        gk = gk * ginv % P
    ys = []
    for c in range(lde_committed.shape[1]):
        acc = [0, 0, 0, 0]
        for k in range(n - 1, -1, -1):
            acc = pyref.ef_mul(acc, z)
            acc[0] = (acc[0] + int(coeffs[k][c])) % P
        ys.append(acc)
    return ys


def subgroup_bitrev(log_h: int) -> list:
    g = orc.two_adic_generator(log_h)
    return [31 * pow(g, pyref.bitrev(X, log_h), P) % P for X in range(1 << log_h)]


def pcs_open_reduced(rounds, log_blowup: int, alpha):
    """rounds: [[(lde_committed ndarray, [points])]] -> (opened values, {log_height: (h,4) reduced opening})."""
    reduced, num_reduced, opened = {}, {}, []
    for mats in rounds:
        opened_round = []
        for lde, points in mats:
            h = lde.shape[0]
            lh = h.bit_length() - 1
            ro = reduced.setdefault(lh, [[0, 0, 0, 0] for _ in range(h)])
            num_reduced.setdefault(lh, 0)
            xs = subgroup_bitrev(lh)
            opened_mat = []
            for z in points:
                ys = eval_matrix_at(lde, log_blowup, z)
                apo = ef_pow(alpha, num_reduced[lh])
                rys, ap = [0, 0, 0, 0], [1, 0, 0, 0]
                for y in ys:
                    rys = pyref.ef_add(rys, pyref.ef_mul(ap, y))
                    ap = pyref.ef_mul(ap, alpha)
                for X in range(h):
                    row, ap = [0, 0, 0, 0], [1, 0, 0, 0]
                    for v in lde[X]:
                        row = pyref.ef_add(row, pyref.ef_scale(ap, int(v)))
                        ap = pyref.ef_mul(ap, alpha)
                    inv = ef_inv(pyref.ef_sub([xs[X], 0, 0, 0], z))
                    t = pyref.ef_mul(pyref.ef_mul(apo, pyref.ef_sub(row, rys)), inv)
                    ro[X] = pyref.ef_add(ro[X], t)
                num_reduced[lh] += lde.shape[1]
                opened_mat.append(ys)
            opened_round.append(opened_mat)
        opened.append(opened_round)
    return opened, {lh: np.array(v, dtype=np.uint32) for lh, v in reduced.items()}


class VerifyError(Exception):
    pass


def verify_query(log_blowup, commits, index, query, betas, reduced_openings, log_max_height):
    """fri/src/verifier.rs:97-165"""
    folded = [0, 0, 0, 0]
    ro = list(reduced_openings)
    for r, (commit, (opened, path), beta) in enumerate(zip(commits, query.commit_phase_openings, betas)):
        log_folded_height = log_max_height - 1 - r
        point_index, index_pair = index & 1, index >> 1
        if ro and ro[0][0] == log_folded_height + 1:
            folded = pyref.ef_add(folded, ro.pop(0)[1])
        assert len(opened) == 1
        row = [int(x) for x in opened[0]]
        committed = row[4 * point_index : 4 * point_index + 4]
        if log_folded_height < log_max_height - 1 and committed != folded:
            raise VerifyError("folded evaluation does not match the committed layer")
        if r == 0 and committed != folded:
            raise VerifyError("reduced opening does not match the first committed layer")
        ok = orc.lib().or_mmcs_verify_batch(
            (orc.C.c_size_t * 1)(1 << log_folded_height), (orc.C.c_size_t * 1)(8), 1, orc.LAYOUT_P3_INJECT, index_pair,
            orc._u32p(np.ascontiguousarray(opened[0], dtype=np.uint32)),
            orc._u8p(np.ascontiguousarray(path, dtype=np.uint8).reshape(-1) if len(path) else np.zeros(32, dtype=np.uint8)),
            log_folded_height, orc._u8p(np.frombuffer(commit, dtype=np.uint8).copy()))
        if not ok:
            raise VerifyError("commit-phase MMCS opening rejected")
        index = index_pair
        folded = [int(x) for x in orc.fold_row_ef(index, log_folded_height, np.array(beta, dtype=np.uint32),
                                                    np.array(row[:4], dtype=np.uint32), np.array(row[4:], dtype=np.uint32))]
    return folded


def pcs_verify(log_blowup, num_queries, pow_bits, rounds, proof, challenger, layout=orc.LAYOUT_P3_INJECT):
    """fri/src/two_adic_pcs.rs:421-530.  rounds: [(commit bytes, [(log_domain_size, [(z, ys)])])]."""
    alpha = [int(x) for x in challenger.sample_ef()]
    log_global_max_height = len(proof.commit_phase_commits) + log_blowup
    betas = []
    for comm in proof.commit_phase_commits:  # verifier.rs:33-40
        challenger.observe_digest(comm)
        betas.append([int(x) for x in challenger.sample_ef()])
    if len(proof.query_proofs) != num_queries:
        raise VerifyError("InvalidProofShape")
    if not challenger.check_witness(pow_bits, proof.pow_witness):
        raise VerifyError("InvalidPowWitness")
    indices = [challenger.sample_bits(log_global_max_height) for _ in range(num_queries)]
    for index, query in zip(indices, proof.query_proofs):
        reduced = {}
        for batch, (commit, mats) in zip(query.input_proof, rounds):
            heights = [(1 << ld) << log_blowup for ld, _ in mats]
            widths = [len(v) for v in batch.opened_values]
            log_batch_max = max(heights).bit_length() - 1
            reduced_index = index >> (log_global_max_height - log_batch_max)
            flat = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.uint32) for v in batch.opened_values]))
            path = np.ascontiguousarray(batch.opening_proof, dtype=np.uint8).reshape(-1)
            ok = orc.lib().or_mmcs_verify_batch(
                (orc.C.c_size_t * len(heights))(*heights), (orc.C.c_size_t * len(widths))(*widths), len(heights), layout,
                reduced_index, orc._u32p(flat), orc._u8p(path if path.size else np.zeros(32, dtype=np.uint8)), log_batch_max,
                orc._u8p(np.frombuffer(commit, dtype=np.uint8).copy()))
            if not ok:
                raise VerifyError("input MMCS opening rejected")
            for mat_opening, (log_domain, points_and_values) in zip(batch.opened_values, mats):
                log_height = log_domain + log_blowup
                rev = pyref.bitrev(index >> (log_global_max_height - log_height), log_height)
                x = 31 * pow(orc.two_adic_generator(log_height), rev, P) % P
                alpha_pow, ro = reduced.get(log_height, ([1, 0, 0, 0], [0, 0, 0, 0]))
                for z, ps_at_z in points_and_values:
                    acc = [0, 0, 0, 0]
                    for p_at_x, p_at_z in zip(mat_opening, ps_at_z):
                        diff = pyref.ef_sub([int(p_at_x), 0, 0, 0], [int(v) for v in p_at_z])
                        acc = pyref.ef_add(acc, pyref.ef_mul(alpha_pow, diff))
                        alpha_pow = pyref.ef_mul(alpha_pow, alpha)
                    ro = pyref.ef_add(ro, pyref.ef_mul(acc, ef_inv(pyref.ef_sub([x, 0, 0, 0], [int(v) for v in z]))))
                reduced[log_height] = (alpha_pow, ro)
        ro_list = [(lh, reduced[lh][1]) for lh in sorted(reduced, reverse=True)]
        folded = verify_query(log_blowup, proof.commit_phase_commits, index, query, betas, ro_list, log_global_max_height)
        if folded != [int(x) for x in proof.final_poly]:
            raise VerifyError("FinalPolyMismatch")
    return True
