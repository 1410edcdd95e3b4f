"""One scene across the GPUs of a node: host side of the slab (brick) protocol (device side: csrc/mgpu.cu).

One process per GPU.  `SlabDomain.substep()` is pool::update (source/pool.cpp:67-106) cut where data of other ranks is
needed:

    integrate owned                          (velocity_handling)
    ROUTE    particles that left the brick change owner: full 80-byte state, new order
             [arrivals from lower ranks | stayers | arrivals from higher ranks] == the order the single-GPU sort sees
    HALO     owned particles inside another rank's brick grown by the halo width go there as ghosts (32 bytes each)
    search   over owned + ghosts; send lists and ghost slots follow the sort permutation
    spread_kernel_width, then KW exchange    (owners' new widths overwrite the ghosts')
    per iteration: prologue (commit, box collision, pack) -> P4 exchange -> density/lambda sweep -> LAMBDA exchange
                   -> apply sweep
    final commit

Two host synchronisations per substep (the message sizes of ROUTE and HALO); the four per-iteration exchanges reuse the
send lists, their sizes are known, and they are stream-ordered.  All accumulators on the path are integers, hence the
N-rank result equals the 1-rank result bit for bit.

The class is backend-agnostic on purpose: `CudaRankBackend` drives the C-ABI (`apbf_sim_mg_*`), and the CPU tests plug in a
numpy/oracle backend over gloo to check this host logic without a GPU (tests/test_multi_rank_gloo.py).  torch.distributed is
plumbing only.
"""
import ctypes as C
import os

import numpy as np

HALO, KW, P4, LAMBDA, POS = 0, 1, 2, 3, 4
WORDS = {HALO: 12, KW: 1, P4: 4, LAMBDA: 1, POS: 4}
STATE_WORDS = 20


class TorchComm:
    """send/recv of int32 word buffers between ranks through torch.distributed (nccl: device tensors, gloo: CPU tensors)"""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.device = device if device is not None else torch.device("cpu")
        self.bytes_sent = 0
        self.messages = 0
        # one all_to_all per exchange instead of a send/recv pair per neighbour: 6 % faster at 4 GPUs, no gain with a single
        # neighbour (APBF_MG_FLAT=0/1 overrides)
        flat = os.environ.get("APBF_MG_FLAT")
        self.use_all_to_all = dist.is_initialized() and dist.get_backend() == "nccl" and (self.world > 2 if flat is None else flat == "1")

    def all_gather_counts(self, counts):
        """counts: list[world] of ints, or a device tensor of them (no host round trip before the collective)
        -> matrix m[src][dst] on the host; the one read-back synchronises"""
        torch = self.torch
        if not torch.is_tensor(counts):
            counts = torch.tensor([int(c) for c in counts], dtype=torch.int32, device=self.device)
        if self.world == 1:
            return [counts.tolist()]
        if counts.is_cuda:
            out = torch.empty(self.world * counts.numel(), dtype=counts.dtype, device=counts.device)
            self.dist.all_gather_into_tensor(out, counts.contiguous())
            return out.view(self.world, -1).tolist()
        out = [torch.zeros_like(counts) for _ in range(self.world)]
        self.dist.all_gather(out, counts)
        return [o.tolist() for o in out]

    def all_to_all(self, buf, send_counts, recv_counts, words):
        """one collective instead of a send/recv pair per neighbour: buf [sum(send_counts), words] ordered by destination ->
        [sum(recv_counts), words] ordered by source"""
        torch, dist = self.torch, self.dist
        out = torch.empty((int(sum(recv_counts)), words), dtype=torch.int32, device=self.device)
        dist.all_to_all_single(out, buf, [int(c) for c in recv_counts], [int(c) for c in send_counts])
        self.bytes_sent += buf.numel() * 4
        self.messages += sum(1 for c in send_counts if c)
        return out

    def exchange(self, send, recv_counts, words):
        """send: {rank: int32 tensor [count, words]}, recv_counts: {rank: count} -> {rank: int32 tensor [count, words]}"""
        torch, dist = self.torch, self.dist
        recv = {r: torch.empty((int(c), words), dtype=torch.int32, device=self.device) for r, c in recv_counts.items() if c}
        ops = []
        for r in sorted(set(send) | set(recv)):
            if r in send and send[r].numel():
                ops.append(dist.P2POp(dist.isend, send[r].contiguous(), r))
                self.bytes_sent += send[r].numel() * 4
                self.messages += 1
            if r in recv:
                ops.append(dist.P2POp(dist.irecv, recv[r], r))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return recv


class SlabDomain:
    def __init__(self, backend, comm, adaptive, solver_iterations, integrate=True, update_transfers=False, width_from_boundary_distance=False):
        self.b, self.comm = backend, comm
        self.rank, self.world = comm.rank, comm.world
        self.adaptive, self.iters, self.integrate = bool(adaptive), int(solver_iterations), bool(integrate)
        # the reference's default adaptive mode (pool.cpp:77-80 and :99-102): kernel width from the boundary distance before the
        # search, update_transfers after the solver (its flood fill crosses the bricks: the halo record carries the ghosts' old
        # boundary distances, and their committed positions follow the solver)
        self.update_transfers, self.width_from_bd = bool(update_transfers), bool(width_from_boundary_distance)
        self.n_own = backend.n_owned()
        self.gid_base = 0
        self.stats = dict(migrated=0, ghosts=0)
        self.timing = {} if os.environ.get("APBF_MG_TIMING") and hasattr(backend, "torch") else None
        # device backends pack every destination in one go and exchange with a single all_to_all (NCCL); the CPU test
        # backend keeps the pairwise send/recv form (gloo)
        self.flat = hasattr(backend, "pack_all") and getattr(comm, "use_all_to_all", False)
        if self.flat:
            backend.flat_only = True
        if self.world > 1 and hasattr(backend, "init_comm") and getattr(comm, "device", None) is not None and comm.device.type == "cuda" \
                and os.environ.get("APBF_MG_C_LOOP"):
            # APBF_MG_C_LOOP=1: solver loop and its exchanges inside the library (own NCCL communicator, grouped send/recv on the
            # context's stream, one call per substep).  Measured: no gain at 2 ranks, 4 % slower than one torch all_to_all per
            # exchange at 4 ranks -- the exchanges are bound by NCCL's latency, not by the host calls around them.  Off by default.
            backend.init_comm(comm)

    def n_owned(self):
        return self.n_own

    def reset(self, n_owned, gid_base=0):
        """the lists were uploaded afresh: this rank owns the first n_owned entries again (global ids must be re-derived by
        the caller if the bricks are not the initial ones)"""
        self.n_own, self.gid_base = int(n_owned), int(gid_base)
        self.b.set_counts(self.n_own, self.n_own, self.gid_base)

    # ---- one exchange of `what` for the current send lists / ghost slots ------------------------------------------------------
    def _refresh_ghosts(self, what):
        if self.world == 1:
            return
        if self.flat:
            # all destinations packed by one kernel, one all_to_all, all ghosts unpacked by one kernel
            n_ghost = sum(self.ghost_counts.values())
            if self.comm.world > 1:
                recv = self.comm.all_to_all(self.b.pack_all(what), [0 if r == self.rank else self.send_counts[r] for r in range(self.world)],
                                            [self.ghost_counts.get(r, 0) for r in range(self.world)], WORDS[what])
                self.b.unpack(what, 0, n_ghost, recv)
            return
        send = {r: self.b.pack(what, r) for r in range(self.world) if r != self.rank and self.send_counts[r]}
        recv = self.comm.exchange(send, self.ghost_counts, WORDS[what])
        off = 0
        for r in range(self.world):
            c = self.ghost_counts.get(r, 0)
            if c:
                self.b.unpack(what, off, c, recv[r])
            off += c

    def _mark(self, name):
        """APBF_MG_TIMING=1: synchronised wall time per section of the substep (diagnostics, serialises the pipeline)"""
        if self.timing is None:
            return
        import time
        self.b.torch.cuda.synchronize()
        now = time.perf_counter()
        self.timing[name] = self.timing.get(name, 0.0) + (now - self._t_last)
        self._t_last = now

    def substep(self):
        b, me, W = self.b, self.rank, self.world
        if self.timing is not None:
            import time
            b.torch.cuda.synchronize()
            self._t_last = time.perf_counter()
        b.set_counts(self.n_own, self.n_own, self.gid_base)          # ghosts of the last substep are dropped
        if self.integrate:
            b.integrate()
        if self.width_from_bd:
            b.kernel_width_from_boundary_distance()
        self._mark("integrate")
        if W > 1:
            # ---- ROUTE ---------------------------------------------------------------------------------------------------
            raw = b.route()                                           # (device counts: no read-back before the collective)
            self._mark("route")
            m = self.comm.all_gather_counts(raw)                      # host sync 1
            counts = [int(c) for c in m[me]]
            self._mark("route_counts")
            first = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
            send = {r: b.pack_state(int(first[r]), int(counts[r])) for r in range(W) if r != me and counts[r]}
            recv = self.comm.exchange(send, {r: m[r][me] for r in range(W) if r != me}, STATE_WORDS)
            b.assemble(int(first[me]), int(counts[me]), [(r, recv[r]) for r in sorted(recv)], me)
            owned = [sum(m[src][dst] for src in range(W)) for dst in range(W)]
            self.stats["migrated"] = int(sum(counts) - counts[me])
            self.n_own, self.gid_base = int(owned[me]), int(sum(owned[:me]))
            b.set_counts(self.n_own, self.n_own, self.gid_base)
            self._mark("route_exchange")
            # ---- HALO ----------------------------------------------------------------------------------------------------
            raw = b.halo_lists()
            self._mark("halo_lists")
            mh = self.comm.all_gather_counts(raw)                     # host sync 2
            self._mark("halo_counts")
            self.send_counts = [int(c) for c in mh[me]]
            if hasattr(b, "set_send_counts"):
                b.set_send_counts(self.send_counts)
            self.ghost_counts = {r: int(mh[r][me]) for r in range(W) if r != me and mh[r][me]}
            n_ghost = sum(self.ghost_counts.values())
            b.set_counts(self.n_own, self.n_own + n_ghost, self.gid_base)
            b.begin_ghosts(n_ghost)
            self._refresh_ghosts(HALO)
            self.stats["ghosts"] = n_ghost
            self._mark("halo_exchange")
        b.search()
        self._mark("search")
        if W > 1:
            b.remap_after_search(self.send_counts, sum(self.ghost_counts.values()))
        if self.adaptive:
            b.spread()
        if W > 1 and getattr(b, "has_comm", False):
            # the library's own communicator: widths to the ghosts, constants, the whole solver loop with its exchanges and the
            # final commit in one call on the context's stream
            b.solve(self.send_counts, self.ghost_counts, self.adaptive, self.iters)
            self._finish()
            self._mark("solve")
            return
        if self.adaptive:
            self._refresh_ghosts(KW)
        self._mark("remap_kw")
        b.prepare()
        for it in range(self.iters):
            b.iter_begin(it)
            self._mark("iter_begin")
            self._refresh_ghosts(P4)
            self._mark("x_p4")
            b.density_lambda()
            self._mark("density_lambda")
            self._refresh_ghosts(LAMBDA)
            self._mark("x_lambda")
            b.apply_delta()
            self._mark("apply_delta")
        b.final_commit()
        self._finish()
        self._mark("final")

    def _finish(self):
        if self.update_transfers:
            self._refresh_ghosts(POS)                       # distances are taken after the solver
            self.b.update_transfers()
        self.b.set_counts(self.n_own, self.n_own, self.gid_base)


class CudaRankBackend:
    """The device steps of the protocol through the C-ABI (apbf_sim_mg_*); staging buffers are torch device tensors."""

    def __init__(self, sim, n_owned, world, rank, halo_range, ghost_capacity):
        import torch
        from . import _check
        self.torch, self._check = torch, _check
        self.sim, self.lib, self.ctx = sim, sim.lib, sim.ctx
        self.world, self.rank = world, rank
        self._n_owned = int(n_owned)
        dev = torch.device("cuda", sim.ctx.device)
        self.dev = dev
        self.cap_per_dest = int(ghost_capacity)
        self.counts_dev = torch.zeros(8, dtype=torch.int32, device=dev)
        self.halo_counts_dev = torch.zeros(8, dtype=torch.int32, device=dev)
        self.send_ids = torch.zeros((world, self.cap_per_dest), dtype=torch.int32, device=dev)
        self.ghost_ids = torch.zeros(max(self.cap_per_dest * max(world - 1, 1), 1), dtype=torch.int32, device=dev)
        self.send_counts = [0] * world
        self.send_flat = None
        self.flat_only = False   # set by SlabDomain when every exchange goes through pack_all
        self.has_comm = False
        self.ghost_first = 0
        self._ck(self.lib.apbf_sim_mg_enable(sim.handle, rank, world, C.c_float(halo_range)))

    def _ck(self, rc):
        self._check(self.ctx, rc)

    def n_owned(self):
        return self._n_owned

    def set_counts(self, n_owned, n_total, gid_base):
        self._n_owned = n_owned
        self._ck(self.lib.apbf_sim_mg_set_counts(self.sim.handle, n_owned, n_total, gid_base))

    def _phase(self, p, it=0):
        self._ck(self.lib.apbf_sim_mg_phase(self.sim.handle, p, it))

    def integrate(self): self._phase(0)
    def search(self): self._phase(1)
    def spread(self): self._phase(2)
    def prepare(self): self._phase(3)
    def iter_begin(self, it): self._phase(4, it)
    def density_lambda(self): self._phase(5)
    def apply_delta(self): self._phase(6)
    def final_commit(self): self._phase(7)
    def kernel_width_from_boundary_distance(self): self._phase(8)
    def update_transfers(self): self._phase(9)

    def route(self):
        self._ck(self.lib.apbf_sim_mg_route(self.sim.handle, self.counts_dev.data_ptr()))
        return self.counts_dev[: self.world]                         # stays on the device until the counts are all-gathered

    def pack_state(self, first, count):
        out = self.torch.empty((count, STATE_WORDS), dtype=self.torch.int32, device=self.dev)
        self._ck(self.lib.apbf_sim_mg_pack_state(self.sim.handle, first, count, out.data_ptr()))
        return out

    def assemble(self, stay_first, stay_count, arrivals, me):
        off = 0
        placed_stayers = False
        for r, buf in arrivals + [(None, None)]:
            if not placed_stayers and (r is None or r > me):
                self._ck(self.lib.apbf_sim_mg_copy_state(self.sim.handle, stay_first, off, stay_count))
                off += stay_count
                placed_stayers = True
            if r is not None:
                self._ck(self.lib.apbf_sim_mg_unpack_state(self.sim.handle, off, buf.shape[0], buf.data_ptr(), 1))
                off += buf.shape[0]
        self._ck(self.lib.apbf_sim_mg_swap(self.sim.handle))

    def halo_lists(self):
        self._ck(self.lib.apbf_sim_mg_halo_lists(self.sim.handle, self.send_ids.data_ptr(), self.cap_per_dest, self.halo_counts_dev.data_ptr()))
        return self.halo_counts_dev[: self.world]

    def set_send_counts(self, c):
        if max(c) > self.cap_per_dest:
            raise RuntimeError(f"halo send list overflow: {max(c)} > capacity {self.cap_per_dest}")
        self.send_counts = list(c)
        self.send_flat = None

    def begin_ghosts(self, n_ghost):
        if n_ghost > self.ghost_ids.numel():
            raise RuntimeError(f"ghost capacity exceeded: {n_ghost} > {self.ghost_ids.numel()}")
        self.ghost_first = self._n_owned
        self.ghosts_sorted = False

    def pack(self, what, dest):
        n = self.send_counts[dest]
        out = self.torch.empty((n, WORDS[what]), dtype=self.torch.int32, device=self.dev)
        self._ck(self.lib.apbf_sim_mg_pack(self.sim.handle, what, self.send_ids[dest].data_ptr(), n, out.data_ptr()))
        return out

    def _flat_ids(self):
        if self.send_flat is None:   # send lists of all destinations, destination after destination
            parts = [self.send_ids[r, : self.send_counts[r]] for r in range(self.world) if r != self.rank and self.send_counts[r]]
            self.send_flat = self.torch.cat(parts) if parts else self.torch.zeros(0, dtype=self.torch.int32, device=self.dev)
        return self.send_flat

    def pack_all(self, what):
        ids = self._flat_ids()
        n = ids.numel()
        out = self.torch.empty((n, WORDS[what]), dtype=self.torch.int32, device=self.dev)
        if n:
            self._ck(self.lib.apbf_sim_mg_pack(self.sim.handle, what, ids.data_ptr(), n, out.data_ptr()))
        return out

    def init_comm(self, comm):
        """the library's own NCCL communicator: rank 0 makes the id, torch.distributed carries its 128 bytes"""
        torch = self.torch
        buf = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            self._ck(self.lib.apbf_mg_nccl_unique_id(buf.data_ptr()))
        t = buf.to(self.dev)
        comm.dist.broadcast(t, 0)
        buf = t.cpu()
        self._ck(self.lib.apbf_sim_mg_comm_init(self.sim.handle, buf.data_ptr(), self.rank, self.world))
        self.has_comm = True
        self.flat_only = True

    def solve(self, send_counts, ghost_counts, adaptive, iterations):
        ids = self._flat_ids()
        sc = (C.c_uint32 * self.world)(*[0 if r == self.rank else int(send_counts[r]) for r in range(self.world)])
        gc = (C.c_uint32 * self.world)(*[int(ghost_counts.get(r, 0)) for r in range(self.world)])
        self._ck(self.lib.apbf_sim_mg_solve(self.sim.handle, ids.data_ptr() if ids.numel() else None, sc, self.ghost_ids.data_ptr(), gc,
                                            int(bool(adaptive)), int(iterations)))

    def unpack(self, what, ghost_offset, count, buf):
        ids = self.ghost_ids.data_ptr() + 4 * ghost_offset
        self._ck(self.lib.apbf_sim_mg_unpack(self.sim.handle, what, ids, self.ghost_first + ghost_offset, count, buf.data_ptr()))

    def remap_after_search(self, send_counts, n_ghost):
        if self.flat_only:      # one remap over the concatenated send lists
            ids = self._flat_ids()
            if ids.numel():
                self._ck(self.lib.apbf_sim_mg_remap(self.sim.handle, ids.data_ptr(), ids.numel(), 0xFFFFFFFF))
        else:
            self.send_flat = None
            for r in range(self.world):
                if r != self.rank and send_counts[r]:
                    self._ck(self.lib.apbf_sim_mg_remap(self.sim.handle, self.send_ids[r].data_ptr(), send_counts[r], 0xFFFFFFFF))
        self._ck(self.lib.apbf_sim_mg_remap(self.sim.handle, self.ghost_ids.data_ptr(), n_ghost, self.ghost_first))


def brick_of_rank(sim, rank):
    lo, hi, halo = (C.c_uint32 * 3)(), (C.c_uint32 * 3)(), (C.c_uint32 * 3)()
    sim.lib.apbf_sim_mg_brick(sim.handle, rank, lo, hi, halo)
    return list(lo), list(hi), list(halo)


def owner_rank_of_positions(position, min_pos, max_pos, res_log2, dims, world):
    """Host-side initial distribution: rank of every particle = top log2(world) bits of its cell key (same float arithmetic
    as calculate_position_hash.comp:23-36)."""
    f32 = np.float32
    lw = int(np.log2(world))
    if lw == 0:
        return np.zeros(len(position), np.int64)
    mn, mx = np.asarray(min_pos, f32), np.asarray(max_pos, f32)
    p = position[:, :3].astype(f32) * f32(1.0 / 262144.0)
    cell = ((p - mn) / (mx - mn) * f32(1 << res_log2))
    cell = np.clip(cell, 0, None).astype(np.uint32) & np.uint32((1 << res_log2) - 1)
    rank = np.zeros(len(position), np.int64)
    for b in range(lw):
        level, axis = b // dims, dims - 1 - (b % dims)
        bit = (cell[:, axis] >> np.uint32(res_log2 - 1 - level)) & 1
        rank = (rank << 1) | bit.astype(np.int64)
    return rank


class LibraryDomain:
    """One scene in bricks with the whole substep -- route, halo, search, solve and every exchange between the ranks -- driven by the
    library (apbf_sim_mg_substep, csrc/mgpu.cu): one C call per substep, everything on the context's stream, no host read-back.
    Python's part is set-up plumbing: the 128-byte NCCL id and the per-pair message capacities travel through torch.distributed."""

    def __init__(self, sim, n_owned, world, rank, halo_range, ghost_capacity, adaptive, solver_iterations, headroom=2.0, transport=None):
        import os
        import torch
        import torch.distributed as dist
        from . import _check
        transport = transport or os.environ.get("APBF_MG_TRANSPORT", "p2p")   # "p2p" (NVLink stores + flags) | "nccl" (send / recv)
        self.torch, self._check = torch, _check
        self.sim, self.lib, self.ctx = sim, sim.lib, sim.ctx
        self.world, self.rank = world, rank
        dev = torch.device("cuda", sim.ctx.device)
        self._ck(self.lib.apbf_sim_mg_enable(sim.handle, rank, world, C.c_float(halo_range)))
        if world > 1:
            buf = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                self._ck(self.lib.apbf_mg_nccl_unique_id(buf.data_ptr()))
            t = buf.to(dev)
            dist.broadcast(t, 0)
            buf = t.cpu()
            self._ck(self.lib.apbf_sim_mg_comm_init(sim.handle, buf.data_ptr(), rank, world))
        # how many ghosts each pair of ranks exchanges right now -> message capacities, the same number on both sides of a pair
        counts = (C.c_uint32 * 8)()
        self._ck(self.lib.apbf_sim_mg_halo_counts(sim.handle, int(n_owned), counts))
        mine = torch.tensor([int(counts[r]) for r in range(8)], dtype=torch.int64, device=dev)
        if world > 1:
            allc = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allc, mine)
            m = [a.tolist() for a in allc]
        else:
            m = [mine.tolist()]
        caps = (C.c_uint32 * 8)()
        for r in range(world):
            if r != rank:
                caps[r] = int(max(4096, headroom * max(m[rank][r], m[r][rank]) + 1024))
        self.halo_caps = [int(caps[r]) for r in range(8)]
        # (message sizes are part of the protocol: every rank must use the same migration capacity)
        rc_t = torch.tensor([max(4096, int(n_owned) // 64)], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(rc_t, op=dist.ReduceOp.MAX)
        self.route_cap = int(rc_t.item())
        self._ck(self.lib.apbf_sim_mg_loop_init(sim.handle, int(n_owned), self.route_cap, caps))
        self.ghost_capacity = int(ghost_capacity)
        self.transport = "nccl"
        self.transport_note = ""
        if world > 1 and transport == "p2p":
            # every rank describes its receive buffers (CUDA IPC handle + offsets); the table of all ranks goes back into the library
            blob_bytes = 512  # APBF_MG_P2P_BLOB_BYTES
            blob = torch.zeros(blob_bytes, dtype=torch.uint8)
            self._ck(self.lib.apbf_sim_mg_p2p_export(sim.handle, blob.data_ptr()))
            allb = [torch.zeros(blob_bytes, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(allb, blob.to(dev))
            table = torch.cat([b.cpu() for b in allb]).contiguous()
            rc = self.lib.apbf_sim_mg_p2p_import(sim.handle, table.data_ptr())
            if rc != 0:   # (kept for the bench line: why this scene fell back to NCCL)
                self.transport_note = self.lib.apbf_ctx_last_error(self.ctx.handle).decode(errors="replace")
            ok = torch.tensor([1 if rc == 0 else 0, -(1 if rc == 0 else 0)], dtype=torch.int64, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # [min, -max]: all ranks or none, the two transports do not mix
            if int(ok[0].item()) == 1:
                self.transport = "p2p"
            elif int(ok[1].item()) != 0:
                raise RuntimeError("peer-to-peer transport: some ranks could map their peers' buffers and some could not")

    def _ck(self, rc):
        self._check(self.ctx, rc)

    def substep(self, n=1):
        self._ck(self.lib.apbf_sim_mg_substep(self.sim.handle, int(n)))

    def _stats(self):
        w = (C.c_uint32 * 8)()
        self._ck(self.lib.apbf_sim_mg_loop_stats(self.sim.handle, w))
        return [int(x) for x in w]

    def n_owned(self):
        return self._stats()[0]

    def reset(self, n_owned, gid_base=0):
        self._ck(self.lib.apbf_sim_mg_loop_reset(self.sim.handle, int(n_owned), int(gid_base)))

    @property
    def stats(self):
        w = self._stats()
        if w[4]:
            raise RuntimeError(f"slab loop capacity exceeded (flags {w[4]}: 1 migration buffer, 2 ghost list, 4 particle capacity)")
        return dict(owned=w[0], ghosts=w[1] - w[0], gid_base=w[2], migrated=w[3], exchanges=w[5], route_cap=self.route_cap,
                    halo_caps=self.halo_caps[: self.world], transport=self.transport, **({"transport_note": self.transport_note} if self.transport_note else {}))
