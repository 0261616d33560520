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
        import os

        self.peer_lanes = max(1, min(7, int(os.environ.get("TS_PEER_LANES", "3"))))  # copy queues of the copy-engine re-shard
        if "TS_REPLICATE_BELOW" in os.environ:  # measurement hook: FRI layers of at most this many elements are gathered
            self.REPLICATE_BELOW = self.MAILBOX_REPLICATE_BELOW = int(os.environ["TS_REPLICATE_BELOW"])
        if device.type == "cuda":
            # torch tensors, the NCCL collectives and the caching allocator's frees are ordered on torch's current stream; the
            # library's kernels run on the context's stream.  They must be ONE stream ("Multi-GPU"): a
            # context with a private stream would let the all-to-all read an LDE that is still being written.
            cur = torch.cuda.current_stream(device).cuda_stream
            if ctx.stream is None or int(ctx.stream) != int(cur):
                raise ts.TapStarkError(
                    "ShardedProver: create the Context on torch's current stream (ts.Context(device, "
                    "torch.cuda.current_stream().cuda_stream)) and keep that stream current while proving")

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
        # peer-mapped receive buffers serve two forms of the re-shard (self.reshard_mode()): "ce" = copy-engine peer copies
        # beside the next chunk's LDE (default), "fused" = the last butterfly pass stores straight into them (TS_P2P=1,
        # measured slower in round 1); "nccl" needs none
        ok = self.device.type == "cuda" and dist.get_backend() == "nccl" and self.reshard_mode() in ("ce", "fused") and G <= 8
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
                        self.__dict__.setdefault("_p2p_mapped", []).append(base_d)
                    owners[d] = base_d + r * Nl * wc * 4  # my block inside rank d's [G, Nl, wc] buffer
                if not ok:
                    break
                plan.append((bases[c], owners))
        flag = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # everyone or no one; also: all mappings exist before the first store
        cache[key] = plan if int(flag.item()) == 1 else None
        return cache[key]

    def _mailbox_setup(self):
        """Peer-mapped mailboxes of the sharded FRI rounds (ts_fri_commit_phase_sharded): allocated and exchanged once.
        Returns the pointer table (rank order) or None when unavailable (CPU/gloo, IPC refused, TS_NO_MAILBOX): the caller
        then exchanges sub-roots with a 32-byte all-gather per round.  Collective on first use."""
        import os

        if hasattr(self, "_mail"):
            return self._mail
        torch, dist, ctx, L, G, r = self.torch, self.dist, self.ctx, self.ctx._L, self.world, self.rank
        ok = self.device.type == "cuda" and dist.get_backend() == "nccl" and G <= 8 and os.environ.get("TS_NO_MAILBOX") is None
        own, handle = C.c_void_p(), (C.c_uint8 * 64)()
        if ok:
            ok = L.ts_device_malloc(ctx._h, L.ts_fri_mailbox_words() * 4, C.byref(own)) == 0 and L.ts_ipc_get_handle(ctx._h, own, handle) == 0
        gathered = [None] * G
        dist.all_gather_object(gathered, bytes(handle) if ok else None)
        ok = ok and all(g is not None for g in gathered)
        table = (C.c_void_p * G)()
        if ok:
            for d in range(G):
                if d == r:
                    table[d] = own.value
                    continue
                q_ = C.c_void_p()
                if L.ts_ipc_open(ctx._h, C.c_char_p(gathered[d]), C.byref(q_)) != 0:
                    ok = False
                    break
                table[d] = q_.value
                self.__dict__.setdefault("_p2p_mapped", []).append(q_.value)
        flag = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # everyone or no one; also orders every mailbox's zero-fill before its first use
        self._mail = table if int(flag.item()) == 1 else None
        self._mail_own = own.value if ok else None
        return self._mail

    def reshard_mode(self) -> str:
        """How LDE rows travel to their owners.  TS_RESHARD = ce | nccl | fused (TS_P2P=1 is the round-1 spelling of fused)."""
        import os

        if os.environ.get("TS_P2P", "0") == "1":
            return "fused"
        return os.environ.get("TS_RESHARD", "ce")

    def close(self):
        """Releases the peer-mapped receive buffers of the fused re-shard (TS_P2P=1): closes the IPC mappings of the other
        ranks' buffers and frees this rank's own.  Collective-free; call after the last step."""
        L, ctx = self.ctx._L, self.ctx
        for plan in self.__dict__.pop("_p2p_cache", {}).values():
            if not plan:
                continue
            for own_base, _owners in plan:
                L.ts_device_free(ctx._h, C.c_void_p(own_base))
        for p_ in self.__dict__.pop("_p2p_mapped", []):
            L.ts_ipc_close(ctx._h, C.c_void_p(p_))
        if self.__dict__.pop("_mail_own", None):
            pass  # the mailbox (36 KiB) lives as long as the process: peers may still have it mapped
        self.__dict__.pop("_mail", None)

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
    MAILBOX_REPLICATE_BELOW = 1 << 20  # the same threshold when sub-roots travel through peer mailboxes (measured per N)

    def _chunks(self, wl: int) -> int:
        """column chunks per rank: the LDE of chunk c+1 overlaps the all-to-all of chunk c (NCCL runs on its own
        stream).  Chunk widths stay powers of two >= 8 so the received blocks feed the fast leaf kernel."""
        import os

        cmax = int(os.environ.get("TS_SHARD_CHUNKS", "4"))  # measurement hook
        c = 1
        # the row hash takes at most 32 equal-width column blocks (b3::MAX_SEG, fold::DOT_MAX_SEG): G * chunks <= 32
        while (c < cmax and 2 * c * self.world <= 32 and wl % (2 * c) == 0 and wl // (2 * c) >= 8
               and (wl // (2 * c)) & (wl // (2 * c) - 1) == 0):
            c *= 2
        return c if self.world > 1 else 1

    def owned_columns(self, width_total: int):
        """Global column indices of this rank's shard, in local order.  Ownership is CHUNK-MAJOR: the w columns are cut
        into C chunks of G*wc columns and rank r owns columns [r*wc, (r+1)*wc) of every chunk.  After the re-shard of chunk c
        a rank therefore holds the CONTIGUOUS global columns [c*G*wc, (c+1)*G*wc) of its rows, i.e. whole 64-byte blocks of
        the row hash, which is absorbed while the next chunk is still being transformed and exchanged."""
        G, r = self.world, self.rank
        wl = width_total // G
        C_ = self._chunks(wl)
        wc = wl // C_
        return [c * G * wc + r * wc + j for c in range(C_) for j in range(wc)]

    def host_panels(self, trace_t):
        """Pinned host copy of this rank's shard as column panels (one contiguous [n, wc] tensor per chunk): the
        layout a column-sharded host prover hands over, so that panel c+1 is copied while panel c is transformed."""
        torch = self.torch
        if isinstance(trace_t, (list, tuple)):
            trace_t = torch.cat(list(trace_t), dim=1)
        n, wl = trace_t.shape
        C_ = self._chunks(wl)
        wc = wl // C_
        panels = []
        for c in range(C_):
            t = torch.empty((n, wc), dtype=torch.int32, pin_memory=True)
            t.copy_(trace_t[:, c * wc : (c + 1) * wc])
            panels.append(t)
        return panels

    def commit_and_fri(self, trace_t, host_panels=None, host_full=None):
        """trace_t: torch int32 [n, w/G] -- this rank's columns of the trace (Montgomery-form bits), resident; or
        host_panels (see host_panels()): the same data in pinned host memory as per-chunk panels; or host_full: the WHOLE
        row-major n x w trace in pinned host memory (what a RowMajorMatrix is), of which this rank copies its column
        windows with strided 2-D copies inside the step.
        Returns dict(root, commits, final_poly, rounds); identical on every rank."""
        ts, ctx, torch, dist, G, r, b = self.ts, self.ctx, self.torch, self.dist, self.world, self.rank, self.b
        L = ctx._L
        if host_full is not None:
            n, wl = host_full.shape[0], host_full.shape[1] // G
        elif host_panels is not None:
            n, wl = host_panels[0].shape[0], sum(p_.shape[1] for p_ in host_panels)
        elif isinstance(trace_t, (list, tuple)):
            n, wl = trace_t[0].shape[0], sum(t_.shape[1] for t_ in trace_t)
        else:
            n, wl = trace_t.shape
        N = n << b
        Nl = N // G
        C_ = self._chunks(wl)
        wc = wl // C_
        gen = int(ts.to_monty(ts.GENERATOR))
        import os

        recv, works, keep = [], [], []
        staged = []
        marks = [self._mark()]
        slab = False
        if host_full is not None:
            # transfer lane 0: window c+1 is copied while window c is transformed (ts_copy_join below makes the LDE of a window
            # wait for exactly the copies issued before it)
            W = host_full.shape[1]
            # Two ways to read a rank's share out of the row-major pinned trace:
            #   window: this rank's wc columns of ALL rows per chunk -- (wc * 4)-byte segments, 32 bytes at 8 GPUs, which PCIe
            #           copies at a fraction of the link rate (measured: 49 ms per step from 4 GPUs on, whatever N)
            #   slab  : ALL G * wc columns of the chunk for this rank's n / G ROWS -- 256-byte segments at every N -- followed by
            #           an all-to-all over NVLink that hands every owner its columns (the transpose of the re-shard after the LDE)
            slab = G > 1 and n % G == 0 and os.environ.get("TS_HOST_INPUT", "slab") == "slab"
            nl_in, Wc = n // G, G * wc
            for c in range(C_):
                staged.append(torch.empty((nl_in, Wc) if slab else (n, wc), dtype=torch.int32, device=self.device))

            def issue_h2d(c):
                if slab:
                    src = host_full.data_ptr() + (r * nl_in * W + c * Wc) * 4
                    ctx.check(L.ts_copy2d_async(ctx._h, 0, C.c_void_p(staged[c].data_ptr()), Wc * 4, C.c_void_p(src), W * 4, Wc * 4, nl_in,
                                                1 if c == 0 else 0), "copy2d_async")
                    return
                src = host_full.data_ptr() + (c * G * wc + r * wc) * 4  # chunk-major ownership (owned_columns)
                ctx.check(L.ts_copy2d_async(ctx._h, 0, C.c_void_p(staged[c].data_ptr()), wc * 4, C.c_void_p(src), W * 4, wc * 4, n,
                                            1 if c == 0 else 0), "copy2d_async")

            def columns_to_owners(c):
                """slab form: rows [r * n/G, ...) x the chunk's G * wc columns -> this rank's wc columns of all n rows"""
                send = staged[c].view(nl_in, G, wc).permute(1, 0, 2).contiguous()  # [owner][row][column]
                out = torch.empty((n, wc), dtype=torch.int32, device=self.device)  # rows in source-rank order = global row order
                send_list, recv_list = list(send.unbind(0)), list(out.view(G, nl_in, wc).unbind(0))
                if self.use_batched_p2p:
                    self.exchange(send_list, recv_list)
                else:
                    dist.all_to_all(recv_list, send_list)
                keep.append(send)
                return out

            issue_h2d(0)
            ctx.check(L.ts_copy_join(ctx._h, 0), "copy_join")
        elif host_panels is not None:
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
        mode = self.reshard_mode()
        plan = self._p2p_setup(N, wc, C_) if (G > 1 and n >= 1 << 18 and wc % 4 == 0) else None
        ce = plan is not None and mode != "fused"
        # copy-engine re-shard: every chunk's LDE stays allocated (its own row range is hashed in place) and the row hash is
        # INCREMENTAL: chunk c-1's columns are absorbed while chunk c is transformed and exchanged
        incremental = ce and (G * wc) % 16 == 0 and G * C_ <= 32 and G * C_ * wc <= 256 and os.environ.get("TS_NO_INC_HASH") is None
        blk = Nl * wc * 4
        lde_ts = [torch.empty((N, wc), dtype=torch.int32, device=self.device) for _ in range(C_)] if ce else []

        def chunk_blocks(c):  # global column order inside a chunk: rank-major
            out = []
            for s_ in range(G):
                if ce and s_ == r:
                    out.append(ts.DeviceMatrix.wrap_device(ctx, lde_ts[c].data_ptr() + r * blk, Nl, wc, keepalive=lde_ts[c]))
                elif plan is not None:
                    out.append(ts.DeviceMatrix.wrap_device(ctx, plan[c][0] + s_ * blk, Nl, wc))
                else:
                    out.append(self._wrap(recv[c][s_], Nl, wc))
            return out

        tree = None
        if incremental:
            blocks = [b_ for c in range(C_) for b_ in chunk_blocks(c)]
            arr = (C.c_void_p * len(blocks))(*[b_._h for b_ in blocks])
            th = C.c_void_p()
            ctx.check(L.ts_mmcs_commit_begin(ctx._h, arr, len(blocks), C.byref(th)), "mmcs_commit_begin")
            tree = ts.ProverData(ctx, th, blocks, False)
        if not hasattr(self, "_p2p_flag") and plan is not None:
            self._p2p_flag = torch.zeros(1, device=self.device, dtype=torch.int32)
        n_lanes = min(self.peer_lanes, 2)
        lag = max(1, min(2, int(os.environ.get("TS_HASH_LAG", "2"))))  # a chunk is hashed `lag` LDEs after its own

        def lanes_of(c):  # three rotating sets of copy queues: "chunk c has landed" is awaited without waiting for c+1, c+2
            return [1 + (c % 3) * 2 + k for k in range(n_lanes)]

        def chunk_landed(c):
            """this rank's copies of chunk c are done and, after the collective, so are everyone's into this rank"""
            for lane in lanes_of(c):
                ctx.check(L.ts_copy_join(ctx._h, lane), "copy_join")
            dist.all_reduce(self._p2p_flag)

        for c in range(C_):
            if host_full is not None:
                if c + 1 < C_:
                    issue_h2d(c + 1)
                src_t = columns_to_owners(c) if slab else staged[c]
            elif host_panels is not None:
                src_t, ev_ = staged[c]
                torch.cuda.current_stream().wait_event(ev_)
            elif isinstance(trace_t, (list, tuple)):
                src_t = trace_t[c]
            else:
                src_t = trace_t if C_ == 1 else trace_t[:, c * wc : (c + 1) * wc].contiguous()
            ev = self._wrap(src_t, n, wc)
            if plan is not None and mode == "fused":
                # fused LDE + re-shard: the last butterfly pass writes each row range into its owner's buffer
                ctx.check(L.ts_coset_lde_batch_scatter(ctx._h, ev._h, b, gen, plan[c][1], G, wc), "coset_lde_batch_scatter")
                keep.append(src_t)
                if host_full is not None and c + 1 < C_:
                    ctx.check(L.ts_copy_join(ctx._h, 0), "copy_join")
                continue
            lde_t = lde_ts[c] if ce else torch.empty((N, wc), dtype=torch.int32, device=self.device)
            lde = self._wrap(lde_t, N, wc)
            ctx.check(L.ts_coset_lde_batch_into(ctx._h, ev._h, b, gen, lde._h), "coset_lde_batch_into")
            if host_full is not None and c + 1 < C_:
                ctx.check(L.ts_copy_join(ctx._h, 0), "copy_join")  # the next window must have landed before its LDE starts
            if ce:
                # this rank's rows [d*Nl, (d+1)*Nl) of the chunk go to rank d's receive buffer over NVLink (copy engines) while
                # the SMs run the next chunk's LDE; peers are visited starting after this rank so that no receiver is
                # everybody's first target; this rank's own rows stay where they are
                lanes = lanes_of(c)
                pieces = max(1, -(-len(lanes) // (G - 1)))  # keep every queue busy whatever the peer count
                issued = set()
                q = 0
                for k in range(G - 1):
                    d = (r + 1 + k) % G
                    for pc_ in range(pieces):
                        lane = lanes[q % len(lanes)]
                        q += 1
                        lo = blk * pc_ // pieces // 16 * 16
                        hi = blk * (pc_ + 1) // pieces // 16 * 16 if pc_ + 1 < pieces else blk
                        ctx.check(L.ts_copy_async(ctx._h, lane, C.c_void_p(plan[c][1][d] + lo), C.c_void_p(lde_t.data_ptr() + d * blk + lo),
                                                  hi - lo, 0 if lane in issued else 1), "copy_async")
                        issued.add(lane)
                keep.append(src_t)
                if incremental and c >= lag:
                    # the exchange of chunk c-lag has had `lag` LDEs (and the windows between them) to finish
                    chunk_landed(c - lag)
                    ctx.check(L.ts_mmcs_commit_window(ctx._h, tree._h, (c - lag) * G, (c - lag + 1) * G), "mmcs_commit_window")
                continue
            recv_t = torch.empty((G, Nl, wc), dtype=torch.int32, device=self.device)
            send_list, recv_list = list(lde_t.view(G, Nl, wc).unbind(0)), list(recv_t.unbind(0))
            if self.use_batched_p2p:
                self.exchange(send_list, recv_list)
            else:
                works.append(dist.all_to_all(recv_list, send_list, async_op=True))  # overlaps the next chunk's LDE
            recv.append(recv_t)
            keep.append((lde_t, src_t))
        marks_detail = [self._mark()]  # all LDE launches queued behind this point
        for w_ in works:
            w_.wait()
        if incremental:
            for c in range(max(C_ - lag, 0), C_):
                chunk_landed(c)
                if c == C_ - 1:
                    marks.append(self._mark())  # LDE + re-shard done (all but the last hash windows already ran inside it)
                ctx.check(L.ts_mmcs_commit_window(ctx._h, tree._h, c * G, (c + 1) * G), "mmcs_commit_window")
            ctx.check(L.ts_mmcs_commit_finish(ctx._h, tree._h, None), "mmcs_commit_finish")
            data = tree
        else:
            if ce:
                for lane in range(1, 7):
                    ctx.check(L.ts_copy_join(ctx._h, lane), "copy_join")
            if plan is not None:
                # every rank's stores into my buffers are complete once all ranks have passed this point in stream order
                dist.all_reduce(self._p2p_flag)
            if not ce:
                keep = None
            marks.append(self._mark())  # LDE + re-shard done
            # global column order: chunk-major, then rank (owned_columns)
            blocks = [b_ for c in range(C_) for b_ in chunk_blocks(c)]
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
        blocks = lde_ts = keep = tree = None
        marks.append(self._mark())  # alpha reduction
        commits, final = self._fri_commit_phase(fri_t, N, ch)
        marks.append(self._mark())  # FRI commit phase
        self._marks = marks
        self._marks_detail = marks_detail
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
        out = {k: m[i].elapsed_time(m[i + 1]) for i, k in enumerate(self.PHASES)}
        d = getattr(self, "_marks_detail", None)
        if d and d[0] is not None:
            out["lde_kernels_done_at"] = m[0].elapsed_time(d[0])  # the rest of lde_and_reshard is the exposed exchange + barrier
        return out

    def _fri_commit_phase(self, cur_t, len_g: int, ch):
        """fri/src/prover.rs:93-141 on a row-sharded codeword (cur_t: this rank's [len_g/G, 4] slice)."""
        import os

        ts, ctx, torch, dist, G, r = self.ts, self.ctx, self.torch, self.dist, self.world, self.rank
        L = ctx._L
        commits: List[bytes] = []
        local = cur_t.shape[0]
        blowup = 1 << self.b
        # sharded rounds are CHAINED on the device (ts_fri_chain_*): sub-root all-gather -> root + sponge + beta in one small
        # kernel -> fold reading beta from device memory; the host sees the roots once, after the last sharded round
        sharded = 0
        lg, lc = len_g, local
        while lg > blowup and lg > self.REPLICATE_BELOW and (lc // 2) >= 256 and (lc // 2) * G == lg // 2:
            sharded, lg, lc = sharded + 1, lg // 2, lc // 2
        mail = self._mailbox_setup() if (sharded or G > 1) else None
        if mail is not None and "TS_REPLICATE_BELOW" not in os.environ and self.REPLICATE_BELOW == type(self).REPLICATE_BELOW:
            # with the mailbox exchange a sharded round costs a few small launches and no host work: keep sharding down to
            # MAILBOX_REPLICATE_BELOW elements (the fold wants at least 256 local rows)
            sharded, lg, lc = 0, len_g, local
            while lg > blowup and lg > self.MAILBOX_REPLICATE_BELOW and (lc // 2) >= 256 and (lc // 2) * G == lg // 2:
                sharded, lg, lc = sharded + 1, lg // 2, lc // 2
            if sharded == 0:
                mail = None
        if mail is not None:
            # every sharded round in ONE library call; the sub-roots travel through peer-mapped mailboxes inside the kernels
            self._epoch = getattr(self, "_epoch", 0) + 1
            out_t = torch.empty((local >> sharded, 4), dtype=torch.int32, device=self.device)
            cbuf = np.zeros((sharded, 32), dtype=np.uint8)
            ctx.check(L.ts_fri_commit_phase_sharded(ctx._h, C.c_void_p(cur_t.data_ptr()), len_g, r, G, sharded, mail, self._epoch, ch._h,
                                                    cbuf.ctypes.data_as(C.c_void_p), C.c_void_p(out_t.data_ptr())), "fri_commit_phase_sharded")
            commits += [cbuf[i].tobytes() for i in range(sharded)]
            cur_t, len_g, local, sharded = out_t, len_g >> sharded, local >> sharded, 0
        chain = C.c_void_p()
        if sharded:
            ctx.check(L.ts_fri_chain_begin(ctx._h, ch._h, sharded, C.byref(chain)), "fri_chain_begin")
            buf = torch.empty((G + 1) * 32, dtype=torch.uint8, device=self.device)
            mine = buf[G * 32 :]
        for rnd in range(sharded):
            h_g, h_l = len_g // 2, local // 2
            leaves = self._wrap(cur_t, h_l, 8)
            _, data = self.mmcs.commit([leaves], host_root=False)
            data.root_to_device(mine.data_ptr())
            data.free()
            dist.all_gather_into_tensor(buf[: G * 32], mine)
            ctx.check(L.ts_fri_chain_step(ctx._h, chain, C.c_void_p(buf.data_ptr()), G, rnd), "fri_chain_step")
            out_t = torch.empty((h_l, 4), dtype=torch.int32, device=self.device)
            ctx.check(L.ts_fri_fold_ext_shard_chain(ctx._h, C.c_void_p(cur_t.data_ptr()), h_g, r * h_l, h_l, chain, None,
                                                    C.c_void_p(out_t.data_ptr())), "fri_fold_ext_shard_chain")
            cur_t, len_g, local = out_t, h_g, h_l
        if len_g > blowup:
            # small layer: gather it everywhere and finish replicated (no further communication)
            full_t = torch.empty((len_g, 4), dtype=torch.int32, device=self.device)
            dist.all_gather_into_tensor(full_t.view(-1), cur_t.contiguous().view(-1))
            if sharded:
                cbuf = np.zeros((sharded, 32), dtype=np.uint8)
                ctx.check(L.ts_fri_chain_end(ctx._h, chain, ch._h, sharded, cbuf.ctypes.data_as(C.c_void_p)), "fri_chain_end")
                commits += [cbuf[i].tobytes() for i in range(sharded)]
            res = ts.bf_commit_phase(self.cfg, [self._wrap(full_t, len_g, 4)], ch, keep_data=False)
            commits += res.commits
            return commits, res.final_poly.tolist()
        if sharded:
            cbuf = np.zeros((sharded, 32), dtype=np.uint8)
            ctx.check(L.ts_fri_chain_end(ctx._h, chain, ch._h, sharded, cbuf.ctypes.data_as(C.c_void_p)), "fri_chain_end")
            commits += [cbuf[i].tobytes() for i in range(sharded)]
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
        self.n, self.wl, self.b, self.seed = 1 << log_rows, width // world, log_blowup, seed
        dev = torch.device("cuda", torch.cuda.current_device())
        # this rank's columns [rank*wl, (rank+1)*wl) of the ONE synthetic trace every N works on (element
        # (r, c) = SplitMix64((seed << 40) + r*width + c) mod p), so root and final polynomial are the 1-GPU ones
        self.prover = ShardedProver(ts, ctx, rank, world, log_blowup, dev)
        C_ = self.prover._chunks(self.wl)
        wc = self.wl // C_
        self.trace_t = []  # one contiguous [n, wc] tensor per column chunk (ShardedProver.owned_columns: chunk-major ownership)
        for c in range(C_):
            t = torch.empty((self.n, wc), dtype=torch.int32, device="cuda")
            ctx.check(ctx._L.ts_fill_splitmix(ctx._h, t.data_ptr(), self.n, wc, seed, c * world * wc + rank * wc, width, 1), "fill_splitmix")
            self.trace_t.append(t)
        self.parallelism = (f"{world} GPUs: LDE column-sharded ({self.wl} cols/GPU), NCCL all-to-all re-shard by rows, "
                            f"row-sharded Blake3 subtrees + FRI folding, 32-byte sub-root all-gathers")
        self.h2d_bytes = self.n * self.wl * 4
        self.d2h_bytes = 32 + 32 * log_rows + 16
        self.host_t = None

    def step_resident(self):
        return self.prover.commit_and_fri(self.trace_t)

    def prepare_host(self):
        """The boundary hands over ONE row-major n x width matrix (a RowMajorMatrix): every rank keeps a pinned host copy of
        the whole trace and copies only its own column windows (strided 2-D copies) inside the timed step."""
        torch, ts = self.torch, self.ts
        n, W = self.n, self.wl * self.world
        full = torch.empty((n, W), dtype=torch.int32, pin_memory=True)
        tmp = torch.empty((n, self.wl), dtype=torch.int32, device="cuda")
        for s_ in range(self.world):  # the same generator every rank uses for its resident shard, 1/G of the columns at a time
            self.ctx.check(self.ctx._L.ts_fill_splitmix(self.ctx._h, tmp.data_ptr(), n, self.wl, self.seed, s_ * self.wl, W, 1),
                           "fill_splitmix")
            torch.cuda.current_stream().synchronize()
            full[:, s_ * self.wl : (s_ + 1) * self.wl].copy_(tmp)
        self.host_t = full
        torch.cuda.synchronize()

    def step_e2e(self):
        # pinned row-major host trace -> this rank's column windows -> HBM inside the step, overlapped with the LDE
        return self.prover.commit_and_fri(None, host_full=self.host_t)

    def release_host(self):
        self.host_t = None

    def close(self):
        self.trace_t = None
        self.prover.close()
        self.ctx.trim()
