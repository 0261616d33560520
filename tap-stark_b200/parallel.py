"""One-process-per-GPU sharding of the commitment path, over torch.distributed.

    LDE            column-sharded: rank r transforms columns [r*w/G, (r+1)*w/G) of every row; no communication
                   (columns are independent polynomials, fri/src/two_adic_pcs.rs:237-240 is column-wise).
    re-shard       ONE all-to-all (NCCL over NVLink): rank s sends rows [r*N/G, (r+1)*N/G) of its LDE columns to
                   rank r.  Committed order is bit-reversed, so a contiguous row range is one Merkle subtree.
    hash           rank r hashes its N/G full rows -- the G received column blocks are fed to the leaf kernel as
                   G segments, no interleave pass -- and builds its subtree; the G sub-roots (32 B each) are
                   all-gathered and the top log2(G) levels hashed on every host: the root equals the 1-GPU root.
    alpha-reduce   row-local (ts_dot_ext_powers_acc per column block).
    FRI            row-sharded: pairs (2i, 2i+1) are adjacent, so folding is local; per round one 32-byte
                   all-gather; when a shard drops below 256 pairs the layer is all-gathered and the remaining
                   rounds run replicated (identical transcript on every rank).

The same code runs on CPU tensors with the gloo backend against the emulated library (tests only).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np


class ShardedProver:
    def __init__(self, ts, ctx, rank: int, world: int, log_blowup: int, device):
        import torch
        import torch.distributed as dist

        if world & (world - 1):
            raise ts.TapStarkError("world size must be a power of two")
        self.ts, self.ctx, self.rank, self.world, self.b = ts, ctx, rank, world, log_blowup
        self.torch, self.dist, self.device = torch, dist, device
        self.mmcs = ts.Blake3MerkleMmcs(ctx)
        self.cfg = ts.FriConfig(log_blowup, 16, 8, self.mmcs)
        self.use_batched_p2p = dist.get_backend() != "nccl"

    # ---- plumbing -------------------------------------------------------------------------------------
    def _wrap(self, t, rows, width):
        return self.ts.DeviceMatrix.wrap_device(self.ctx, t.data_ptr(), rows, width, keepalive=t)

    def _sync_lib(self):
        # library work is enqueued on the ctx stream (= torch's current stream on CUDA); the emulated build
        # is synchronous
        pass

    def exchange(self, send: List, recv: List):
        """all-to-all of equally sized chunks: send[s] -> rank s, recv[s] <- rank s."""
        dist = self.dist
        if not self.use_batched_p2p:
            dist.all_to_all(recv, send)
            return
        recv[self.rank].copy_(send[self.rank])
        ops = []
        for s in range(self.world):
            if s == self.rank:
                continue
            ops.append(dist.P2POp(dist.isend, send[s], s))
            ops.append(dist.P2POp(dist.irecv, recv[s], s))
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    # ---- peer-mapped receive buffers: the LDE's last pass stores every row at its owner (NVLink), no all-to-all ------
    def _p2p_setup(self, N: int, wc: int, n_chunks: int):
        """Allocates this rank's receive buffers ([G, N/G, wc] per chunk), exchanges their CUDA IPC handles and maps every
        peer's buffers.  Returns per-chunk (own base pointer, owner pointer table) or None when the fused path is
        unavailable or not requested (CPU/gloo, TS_P2P != 1, IPC or peer access refused): the caller then uses the NCCL
        all-to-all overlapped with the next chunk's LDE, which is the faster form on NVSwitch B200s.  Collective."""
        import os

        key = (N, wc, n_chunks)
        cache = self.__dict__.setdefault("_p2p_cache", {})
        if key in cache:
            return cache[key]
        torch, dist, ctx, L, G, r = self.torch, self.dist, self.ctx, self.ctx._L, self.world, self.rank
        # opt-in: measured SLOWER than the chunk-overlapped NCCL all-to-all on 2/4/8 B200s (profiles/r01/README.md)
        ok = self.device.type == "cuda" and dist.get_backend() == "nccl" and os.environ.get("TS_P2P", "0") == "1" and G <= 8
        Nl = N // G
        bases, handles = [], []
        if ok:
            for _ in range(n_chunks):
                p_ = C.c_void_p()
                h_ = (C.c_uint8 * 64)()
                if L.ts_device_malloc(ctx._h, G * Nl * wc * 4, C.byref(p_)) != 0 or L.ts_ipc_get_handle(ctx._h, p_, h_) != 0:
                    ok = False
                    break
                bases.append(p_.value)
                handles.append(bytes(h_))
        gathered = [None] * G
        dist.all_gather_object(gathered, handles if ok else None)
        ok = ok and all(g is not None for g in gathered)
        plan = None
        if ok:
            plan = []
            for c in range(n_chunks):
                owners = (C.c_void_p * G)()
                for d in range(G):
                    if d == r:
                        base_d = bases[c]
                    else:
                        q_ = C.c_void_p()
                        if L.ts_ipc_open(ctx._h, C.c_char_p(gathered[d][c]), C.byref(q_)) != 0:
                            ok = False
                            break
                        base_d = q_.value
                    owners[d] = base_d + r * Nl * wc * 4  # my block inside rank d's [G, Nl, wc] buffer
                if not ok:
                    break
                plan.append((bases[c], owners))
        flag = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # everyone or no one; also: all mappings exist before the first store
        cache[key] = plan if int(flag.item()) == 1 else None
        return cache[key]

    def combine_roots(self, data) -> bytes:
        """all-gather the G sub-roots (device to device: the sub-root never visits the host on its own) and hash the
        top log2(G) levels (parent = Blake3(left || right)) on every host: one device->host read per call."""
        torch, dist = self.torch, self.dist
        G = self.world
        buf = torch.empty((G + 1) * 32, dtype=torch.uint8, device=self.device)
        mine = buf[G * 32 :]
        data.root_to_device(mine.data_ptr())
        dist.all_gather_into_tensor(buf[: G * 32], mine)
        flat = buf[: G * 32].cpu().numpy()
        layer = [flat[32 * i : 32 * i + 32].tobytes() for i in range(G)]
        L = self.ctx._L
        while len(layer) > 1:
            nxt = []
            for i in range(0, len(layer), 2):
                pair = layer[i] + layer[i + 1]
                out = (C.c_uint8 * 32)()
                L.ts_blake3_host(C.c_char_p(pair), 64, out)
                nxt.append(bytes(out))
            layer = nxt
        return layer[0]

    # ---- the step -------------------------------------------------------------------------------------
    REPLICATE_BELOW = 1 << 20  # FRI layers of at most this many elements are gathered and folded on every rank

    def _chunks(self, wl: int) -> int:
        """column chunks per rank: the LDE of chunk c+1 overlaps the all-to-all of chunk c (NCCL runs on its own
        stream).  Chunk widths stay powers of two >= 8 so the received blocks feed the fast leaf kernel."""
        import os

        cmax = int(os.environ.get("TS_SHARD_CHUNKS", "4"))  # measurement hook
        c = 1
        while c < cmax and wl % (2 * c) == 0 and wl // (2 * c) >= 8 and (wl // (2 * c)) & (wl // (2 * c) - 1) == 0:
            c *= 2
        return c if self.world > 1 else 1

    def host_panels(self, trace_t):
        """Pinned host copy of this rank's shard as column panels (one contiguous [n, wc] tensor per chunk): the
        layout a column-sharded host prover hands over, so that panel c+1 is copied while panel c is transformed."""
        torch = self.torch
        n, wl = trace_t.shape
        C_ = self._chunks(wl)
        wc = wl // C_
        panels = []
        for c in range(C_):
            t = torch.empty((n, wc), dtype=torch.int32, pin_memory=True)
            t.copy_(trace_t[:, c * wc : (c + 1) * wc])
            panels.append(t)
        return panels

    def commit_and_fri(self, trace_t, host_panels=None):
        """trace_t: torch int32 [n, w/G] -- this rank's columns of the trace (Montgomery-form bits), resident; or
        host_panels (see host_panels()): the same data in pinned host memory, copied H2D inside the step.
        Returns dict(root, commits, final_poly, rounds); identical on every rank."""
        ts, ctx, torch, dist, G, r, b = self.ts, self.ctx, self.torch, self.dist, self.world, self.rank, self.b
        L = ctx._L
        if host_panels is not None:
            n, wl = host_panels[0].shape[0], sum(p_.shape[1] for p_ in host_panels)
        else:
            n, wl = trace_t.shape
        N = n << b
        Nl = N // G
        C_ = self._chunks(wl)
        wc = wl // C_
        gen = int(ts.to_monty(ts.GENERATOR))
        recv, works, keep = [], [], []
        staged = []
        marks = [self._mark()]
        if host_panels is not None:
            # all panels are queued on a side stream up front; the main stream waits panel by panel
            if not hasattr(self, "_copy_stream"):
                self._copy_stream = torch.cuda.Stream()
            main = torch.cuda.current_stream()
            self._copy_stream.wait_stream(main)
            with torch.cuda.stream(self._copy_stream):
                for c in range(C_):
                    d_ = torch.empty((n, wc), dtype=torch.int32, device=self.device)
                    d_.copy_(host_panels[c], non_blocking=True)
                    e_ = torch.cuda.Event()
                    e_.record(self._copy_stream)
                    d_.record_stream(main)
                    staged.append((d_, e_))
        plan = self._p2p_setup(N, wc, C_) if (G > 1 and n >= 1 << 18 and wc % 4 == 0) else None
        for c in range(C_):
            if host_panels is not None:
                src_t, ev_ = staged[c]
                torch.cuda.current_stream().wait_event(ev_)
            else:
                src_t = trace_t if C_ == 1 else trace_t[:, c * wc : (c + 1) * wc].contiguous()
            ev = self._wrap(src_t, n, wc)
            if plan is not None:
                # fused LDE + re-shard: the last butterfly pass writes each row range into its owner's buffer
                ctx.check(L.ts_coset_lde_batch_scatter(ctx._h, ev._h, b, gen, plan[c][1], G, wc), "coset_lde_batch_scatter")
                keep.append(src_t)
                continue
            lde_t = torch.empty((N, wc), dtype=torch.int32, device=self.device)
            lde = self._wrap(lde_t, N, wc)
            ctx.check(L.ts_coset_lde_batch_into(ctx._h, ev._h, b, gen, lde._h), "coset_lde_batch_into")
            recv_t = torch.empty((G, Nl, wc), dtype=torch.int32, device=self.device)
            send_list, recv_list = list(lde_t.view(G, Nl, wc).unbind(0)), list(recv_t.unbind(0))
            if self.use_batched_p2p:
                self.exchange(send_list, recv_list)
            else:
                works.append(dist.all_to_all(recv_list, send_list, async_op=True))  # overlaps the next chunk's LDE
            recv.append(recv_t)
            keep.append((lde_t, src_t))
        for w_ in works:
            w_.wait()
        if plan is not None:
            # every rank's stores into my buffers are complete once all ranks have passed this point in stream order
            if not hasattr(self, "_p2p_flag"):
                self._p2p_flag = torch.zeros(1, device=self.device, dtype=torch.int32)
            dist.all_reduce(self._p2p_flag)
        del keep
        marks.append(self._mark())  # LDE + re-shard done
        # global column order: rank-major, then chunk
        blocks = []
        for s_ in range(G):
            for c in range(C_):
                if plan is not None:
                    blocks.append(ts.DeviceMatrix.wrap_device(ctx, plan[c][0] + s_ * Nl * wc * 4, Nl, wc))
                else:
                    blocks.append(self._wrap(recv[c][s_], Nl, wc))
        _, data = self.mmcs.commit(blocks, host_root=False)
        root = self.combine_roots(data)
        marks.append(self._mark())  # leaf hashes, sub-tree, sub-root all-gather
        ch = ts.BfChallenger()
        ch.observe(root)
        alpha = ch.sample()
        am = ts.to_monty(alpha)
        ap = C.c_void_p()
        ctx.check(L.ts_alpha_powers(ctx._h, am.ctypes.data_as(C.c_void_p), wl * G, C.byref(ap)), "alpha_powers")
        fri_t = torch.empty((Nl, 4), dtype=torch.int32, device=self.device)
        fri = self._wrap(fri_t, Nl, 4)
        # blocks are in global column order (first_col ascending and contiguous): one pass over all of them
        arr = (C.c_void_p * len(blocks))(*[blk._h for blk in blocks])
        ctx.check(L.ts_dot_ext_powers_blocks(ctx._h, arr, len(blocks), ap, fri._h), "dot_ext_powers_blocks")
        L.ts_matrix_free(ap)
        data.free()
        marks.append(self._mark())  # alpha reduction
        commits, final = self._fri_commit_phase(fri_t, N, ch)
        marks.append(self._mark())  # FRI commit phase
        self._marks = marks
        return {"root": root, "commits": commits, "final_poly": final, "rounds": len(commits)}

    PHASES = ("lde_and_reshard", "hash_tree_roots", "alpha_reduction", "fri_commit_phase")

    def _mark(self):
        if not self.torch.cuda.is_available():
            return None
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def last_phases(self):
        """Device time of the last step's phases in ms (CUDA events on the step's stream); None on CPU."""
        m = getattr(self, "_marks", None)
        if not m or m[0] is None:
            return None
        m[-1].synchronize()
        return {k: m[i].elapsed_time(m[i + 1]) for i, k in enumerate(self.PHASES)}

    def _fri_commit_phase(self, cur_t, len_g: int, ch):
        """fri/src/prover.rs:93-141 on a row-sharded codeword (cur_t: this rank's [len_g/G, 4] slice)."""
        ts, ctx, torch, dist, G, r = self.ts, self.ctx, self.torch, self.dist, self.world, self.rank
        L = ctx._L
        commits: List[bytes] = []
        local = cur_t.shape[0]
        blowup = 1 << self.b
        while len_g > blowup:
            h_g, h_l = len_g // 2, local // 2
            if len_g > self.REPLICATE_BELOW and h_l >= 256 and h_l * G == h_g:
                leaves = self._wrap(cur_t, h_l, 8)
                _, data = self.mmcs.commit([leaves], host_root=False)
                root = self.combine_roots(data)
                data.free()
                commits.append(root)
                ch.observe(root)
                beta = ts.to_monty(ch.sample())
                out_t = torch.empty((h_l, 4), dtype=torch.int32, device=self.device)
                ctx.check(L.ts_fri_fold_ext_shard(ctx._h, C.c_void_p(cur_t.data_ptr()), h_g, r * h_l, h_l,
                                                  beta.ctypes.data_as(C.c_void_p), None, C.c_void_p(out_t.data_ptr())),
                          "fri_fold_ext_shard")
                cur_t, len_g, local = out_t, h_g, h_l
                continue
            # small layer: gather it everywhere and finish replicated (no further communication)
            full_t = torch.empty((len_g, 4), dtype=torch.int32, device=self.device)
            dist.all_gather_into_tensor(full_t.view(-1), cur_t.contiguous().view(-1))
            res = ts.bf_commit_phase(self.cfg, [self._wrap(full_t, len_g, 4)], ch, keep_data=False)
            commits += res.commits
            return commits, res.final_poly.tolist()
        # len_g <= blowup without ever gathering (only when the input was already tiny)
        parts = [torch.empty_like(cur_t) for _ in range(G)]
        dist.all_gather(parts, cur_t)
        full = torch.cat(parts, dim=0).cpu().numpy().view(np.uint32)
        vals = ts.from_monty(full)
        if not all((vals[i] == vals[0]).all() for i in range(vals.shape[0])):
            raise ts.TapStarkError("commit phase: final layer is not constant")
        return commits, vals[0].tolist()


class ShardedRunner:
    """bench.py driver for N > 1 (strong scaling: the 2^log_rows x width trace is split by columns)."""

    def __init__(self, ts, ctx, log_rows, width, log_blowup, rank, world, seed=0):
        import torch

        self.ts, self.ctx, self.torch = ts, ctx, torch
        self.rank, self.world = rank, world
        if width % world:
            raise ts.TapStarkError("width must be divisible by the number of GPUs")
        self.n, self.wl, self.b = 1 << log_rows, width // world, log_blowup
        dev = torch.device("cuda", torch.cuda.current_device())
        # this rank's columns [rank*wl, (rank+1)*wl) of the ONE synthetic trace every N works on (element
        # (r, c) = SplitMix64((seed << 40) + r*width + c) mod p), so root and final polynomial are the 1-GPU ones
        self.trace_t = torch.empty((self.n, self.wl), dtype=torch.int32, device="cuda")
        ctx.check(ctx._L.ts_fill_splitmix(ctx._h, self.trace_t.data_ptr(), self.n, self.wl, seed, rank * self.wl, width, 1),
                  "fill_splitmix")
        self.prover = ShardedProver(ts, ctx, rank, world, log_blowup, dev)
        self.parallelism = (f"{world} GPUs: LDE column-sharded ({self.wl} cols/GPU), NCCL all-to-all re-shard by rows, "
                            f"row-sharded Blake3 subtrees + FRI folding, 32-byte sub-root all-gathers")
        self.h2d_bytes = self.n * self.wl * 4
        self.d2h_bytes = 32 + 32 * log_rows + 16
        self.host_t = None

    def step_resident(self):
        return self.prover.commit_and_fri(self.trace_t)

    def prepare_host(self):
        self.host_t = self.prover.host_panels(self.trace_t)
        self.torch.cuda.synchronize()

    def step_e2e(self):
        # this rank's column shard: pinned host panels -> HBM inside the step, overlapped with the LDE
        return self.prover.commit_and_fri(None, host_panels=self.host_t)

    def release_host(self):
        self.host_t = None

    def close(self):
        self.trace_t = None
        self.ctx.trim()
