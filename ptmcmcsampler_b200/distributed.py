"""Sharding across GPUs (one process per GPU, ``torch.distributed``).

**Walker sharding** (``run``, ``pooled_adapt``).  Every rank owns ``nwalkers`` complete ladders (``walker_offset = rank * nwalkers`` keys the RNG, so a
walker's draws do not depend on the sharding).  The MH step and the swap need no communication.  The
only coupling is the pooled proposal covariance: at every covariance boundary the ranks exchange
their batch moments {n, mean[d], M2c[d*d]} (d*d+d+1 doubles), merge them with Chan's formula in rank
order, and each applies the merged batch -- the multi-device form of the reference's rank-0
``send(cov)`` broadcast (ref PTMCMCSampler.py:545-560).  The DE history stays shard-local.

**Ladder sharding** (``run_ladder``, ``LadderComm``; BASELINE config 5: 256 rungs as 32 per GPU).  Rank g
owns the contiguous rungs ``[g*T/G, (g+1)*T/G)`` of every walker.  MH steps need no communication.  The
swap sweep of the reference runs hottest pair first on rank 0 after a gather of every rung (ref
:660-697); here it is cut at the shard boundaries and only the boundary rung moves, between nearest
neighbours: each shard sends its top rung up (no dependency), sweeps its own pairs once the carry of
the hotter shard has arrived, passes its own carry down, and resolves its lowest position against the
colder shard's top rung.  Both sides of a boundary evaluate the same acceptance from the counter-based
stream, so the result is bit-identical to the unsharded sweep.  The adaptive state lives on the shard
holding T=1: its eigen-factor (d*d+d doubles) is broadcast at every covariance update and its AM ring
before every DE-history update, the multi-device form of the reference's rank-0 ``send(cov)`` /
``send(_DEbuffer)`` (ref :545-571).

The drivers are written against the small engine surface both ``_cabi.Engine`` (device pointers, NCCL)
and the test oracle (host pointers, gloo) expose, so the exchange logic is testable without a GPU.
"""
import ctypes
import os

import numpy as np


def merge_batches(batches):
    """Chan merge of [{n, mean[d], M2c[d*d]}, ...] in list order (deterministic)."""
    batches = [np.asarray(b, dtype=np.float64) for b in batches]
    d = int(round((-1 + np.sqrt(1 + 4 * (len(batches[0]) - 1))) / 2))
    n, mean, m2 = 0.0, np.zeros(d), np.zeros((d, d))
    for b in batches:
        nb, mb, m2b = b[0], b[1:1 + d], b[1 + d:].reshape(d, d)
        if nb == 0:
            continue
        delta = mb - mean
        tot = n + nb
        m2 = m2 + m2b + np.outer(delta, delta) * (n * nb / tot)
        mean = mean + delta * (nb / tot)
        n = tot
    return np.concatenate([[n], mean, m2.ravel()])


class _CudaAlias(object):
    """Exposes a raw device allocation through ``__cuda_array_interface__`` (zero-copy torch view)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def _device_view(engine, ptr, n):
    import torch

    return torch.as_tensor(_CudaAlias(ptr, n), device=torch.device("cuda", engine.device))


def pooled_adapt(engine, group=None):
    """If a covariance update is due, pool the batch moments over the ranks of ``group`` and apply.
    Collective: every rank must call it at the same iteration.  Returns True if an update ran.

    With a CUDA engine over NCCL nothing leaves the device and the host does not wait: the batch moments are
    all-gathered from the engine's own buffer on the engine's stream and merged by a kernel
    (``ptmcmc_adapt_begin_dev`` / ``ptmcmc_adapt_finish_dev``).  Host engines (the test oracle over gloo) exchange the
    same batches through numpy and merge them with :func:`merge_batches`, the same recurrence in the same order."""
    import torch
    import torch.distributed as dist

    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if multi and hasattr(engine, "adapt_begin_dev") and dist.get_backend(group) == "nccl":
        due, ptr, n = engine.adapt_begin_dev()
        if not due:
            return False
        world = dist.get_world_size(group)
        st = getattr(engine, "_adapt_comm", None)
        if st is None:
            dev = torch.device("cuda", engine.device)
            stream = torch.cuda.ExternalStream(engine.stream, device=dev)
            with torch.cuda.stream(stream):
                parts = torch.empty(world * n, dtype=torch.float64, device=dev)
            st = engine._adapt_comm = (stream, _device_view(engine, ptr, n), parts)
        stream, mine, parts = st
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(parts, mine, group=group)
        engine.adapt_finish_dev(parts.data_ptr(), world, world * engine.cov_update * engine.W)
        return True
    batch = engine.adapt_begin()
    if batch is None:
        return False
    if multi:
        backend = dist.get_backend(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        mine = torch.from_numpy(batch).to(dev)
        parts = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, mine, group=group)
        batch = merge_batches([p.cpu().numpy() for p in parts])
    engine.adapt_finish(batch)
    return True


def run(engine, niter, group=None):
    """``engine.run(niter)`` with pooled covariance updates at every boundary."""
    cu = engine.cov_update
    done = 0
    while done < niter:
        it = engine.iteration
        pooled_adapt(engine, group)          # boundary reached by a previous call
        step = min(niter - done, cu - it % cu)
        engine.run(step)
        done += step
    return done


# ------------------------------------------------------------------ ladder sharding ------------
def ladder_slice(ntemps_global, world, rank):
    """Contiguous rungs ``[lo, hi)`` of shard ``rank``; the ladder must divide evenly."""
    if ntemps_global % world != 0:
        raise ValueError("%d temperatures do not divide over %d shards" % (ntemps_global, world))
    n = ntemps_global // world
    return rank * n, (rank + 1) * n


def ladder_shard_kwargs(ladder, world, rank):
    """Engine keyword arguments (ntemps, ladder, temp_offset, ...) of shard ``rank``."""
    ladder = np.asarray(ladder, dtype=np.float64)
    lo, hi = ladder_slice(len(ladder), world, rank)
    return dict(ntemps=hi - lo, ladder=ladder[lo:hi].copy(), temp_offset=lo, ntemps_global=len(ladder),
                ladder_above=float(ladder[hi]) if hi < len(ladder) else 0.0,
                ladder_below=float(ladder[lo - 1]) if lo > 0 else 0.0)


_BULK_GROUPS = {}
_P2P_LINKS = {}  # (ranks, device, message size) -> every shard's mailbox handle, once all of them were mapped


class LadderComm(object):
    """Neighbour exchange and cold-shard broadcasts of one ladder shard over ``torch.distributed``.

    ``device`` is the torch device of the message buffers: a CUDA device for ``_cabi.Engine`` (NCCL) or
    ``"cpu"`` for a host engine (gloo).  With CUDA, ``stream`` (the engine's stream) is made current so
    that collectives and engine kernels are ordered on the device without host synchronisation."""

    def __init__(self, engine, group=None, device=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.coldest, self.hottest = self.rank == 0, self.rank == self.world - 1
        if device is None:
            device = torch.device("cuda", engine.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        self.device = torch.device(device)
        self.stream = None
        if self.device.type == "cuda":
            self.stream = torch.cuda.ExternalStream(engine.stream, device=self.device)
        n = engine.swap_msg_doubles
        with self._on_stream():
            self.up_out, self.below_in, self.carry_in, self.carry_out = (
                torch.zeros(n, dtype=torch.float64, device=self.device) for _ in range(4))
        # bulk transfers (the AM ring, 1.3 GB per covUpdate at C5) get their own communicator: operations of one process
        # group are serialised on one NCCL stream, and a 130 MB broadcast queued ahead of a swap message would stall the
        # whole ladder for its duration
        self.bulk_group = group
        if self.world > 1 and self.device.type == "cuda" and not os.environ.get("PTMCMC_NO_BULK_GROUP"):
            ranks = tuple(dist.get_process_group_ranks(group) if group is not None else range(dist.get_world_size()))
            if ranks not in _BULK_GROUPS:  # one extra communicator per set of ranks, for the life of the process
                _BULK_GROUPS[ranks] = dist.new_group(ranks=list(ranks))
            self.bulk_group = _BULK_GROUPS[ranks]
        # swap messages through peer memory (the kernels store into the neighbour's mailbox over NVLink and spin on a flag
        # in their own memory) instead of a send / receive pair per hop of the 8-stage carry chain; every shard must
        # succeed in mapping its neighbours, else all keep the NCCL messages
        self.p2p = False
        if self.world > 1 and self.device.type == "cuda" and not os.environ.get("PTMCMC_NO_P2P"):
            self.p2p = self._connect_p2p(engine)
        self.maint_done = -1
        self.am_sent = -1      # last iteration whose AM-ring slot has been broadcast from the cold shard
        self.am_works = []     # broadcasts in flight
        self._ring = None
        self._factor = None    # device views of the engine's eigen-factor

    def _connect_p2p(self, engine):
        """Map the neighbours' mailboxes.  The library keeps mailboxes and mappings for the life of the process, so a later
        engine of the same geometry gets the same mailbox back, with the sequence number every shard left it at: then
        nothing is exchanged at all.  Otherwise one collective carries every shard's handle and sequence number (and, the
        first time, a second one the verdict that every rank could map its neighbours)."""
        mine = None
        try:
            mine = (engine.p2p_open()[0], engine.p2p_seq)
        except Exception:  # no CUDA IPC here (e.g. a restricted container)
            pass
        ranks = tuple(self.dist.get_process_group_ranks(self.group) if self.group is not None else range(self.world))
        key = (ranks, engine.device, engine.swap_msg_doubles)
        known = _P2P_LINKS.get(key)
        if mine is not None and known is not None and known[self.rank] == mine[0]:
            engine.p2p_connect(None if self.hottest else known[self.rank + 1], None if self.coldest else known[self.rank - 1], mine[1])
            return True
        parts = [None] * self.world
        self.dist.all_gather_object(parts, mine, group=self.group)
        if any(q is None for q in parts):
            return False
        ok = 1
        try:
            engine.p2p_connect(None if self.hottest else parts[self.rank + 1][0], None if self.coldest else parts[self.rank - 1][0],
                               max(q[1] for q in parts))
        except Exception:
            ok = 0
        flags = [None] * self.world
        self.dist.all_gather_object(flags, ok, group=self.group)
        if all(flags):
            _P2P_LINKS[key] = [q[0] for q in parts]
            return True
        return False

    def check(self, engine):
        """Synchronise and raise -- on every rank together -- if a neighbour's swap message never arrived on any of them
        (peer-memory exchange only)."""
        if not self.p2p:
            return
        failed, msg = 0, ""
        try:
            engine.p2p_error()
        except Exception as exc:  # collective below first: the other ranks must not be left waiting
            failed, msg = 1, str(exc)
        flag = self.torch.tensor([failed], dtype=self.torch.int32, device=self.device)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MAX, group=self.group)
        if int(flag.item()):
            raise RuntimeError("ladder sharding: " + (msg or "a swap message did not arrive on another rank"))

    def _on_stream(self):
        import contextlib

        return self.torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def _peer(self, r):
        return self.dist.get_global_rank(self.group, r) if self.group is not None else r

    def alias(self, ptr, n):
        """torch view of ``n`` doubles at address ``ptr`` in the engine's memory space."""
        if self.device.type == "cuda":
            return self.torch.as_tensor(_CudaAlias(ptr, n), device=self.device)
        buf = (ctypes.c_double * n).from_address(ptr)
        return self.torch.from_numpy(np.ctypeslib.as_array(buf))

    def exchange_top(self):
        """Start: my top rung -> hotter neighbour, colder neighbour's top rung -> me."""
        ops = []
        if not self.hottest:
            ops.append(self.dist.P2POp(self.dist.isend, self.up_out, self._peer(self.rank + 1), self.group))
        if not self.coldest:
            ops.append(self.dist.P2POp(self.dist.irecv, self.below_in, self._peer(self.rank - 1), self.group))
        return self.dist.batch_isend_irecv(ops) if ops else []

    def recv_carry(self):
        # (batched like the top-rung exchange: unbatched send / recv would be serialised with every other operation of
        # the process group)
        for r in self.dist.batch_isend_irecv([self.dist.P2POp(self.dist.irecv, self.carry_in, self._peer(self.rank + 1),
                                                              self.group)]):
            r.wait()

    def send_carry(self):
        return self.dist.batch_isend_irecv([self.dist.P2POp(self.dist.isend, self.carry_out, self._peer(self.rank - 1),
                                                            self.group)])

    def bcast(self, tensor):
        self.dist.broadcast(tensor, self._peer(0), group=self.group)


def ladder_swap(engine, comm):
    """One sharded swap sweep (the engine stopped at a swap iteration)."""
    if comm.p2p:
        for phase in (0, 1, 2):  # only enqueues: the flags in the mailboxes order the shards on the devices
            engine.swap_p2p(phase)
        return
    with comm._on_stream():
        if not comm.hottest:
            engine.swap_pack_top(comm.up_out.data_ptr())
        reqs = comm.exchange_top()
        if not comm.hottest:
            comm.recv_carry()
        engine.swap_sweep(None if comm.hottest else comm.carry_in.data_ptr(),
                          None if comm.coldest else comm.carry_out.data_ptr())
        if not comm.coldest:
            reqs = list(reqs) + list(comm.send_carry())
        for r in reqs:
            r.wait()
        engine.swap_finish(None if comm.coldest else comm.below_in.data_ptr())


def ladder_am_progress(engine, comm, flush=False):
    """Broadcast the AM-ring slots the cold shard has filled since the last call (asynchronously, so the
    transfer overlaps the following MH segments instead of costing 1.3 GB at once before the DE update)."""
    if comm.world == 1:
        return
    it, cu = engine.iteration, engine.cov_update
    lo = max(comm.am_sent + 1, it - cu + 1, 0)
    if it < lo or (not flush and it - lo + 1 < max(1, cu // 10)):
        return
    if comm._ring is None:
        ptr, n = engine.am_ring()
        comm._ring = (comm.alias(ptr, n), n // cu)
    ring, per = comm._ring
    with comm._on_stream():
        while lo <= it:
            s0 = lo % cu
            run = min(it - lo + 1, cu - s0)
            comm.am_works.append(comm.dist.broadcast(ring[s0 * per:(s0 + run) * per], comm._peer(0), group=comm.bulk_group,
                                                     async_op=True))
            lo += run
    comm.am_sent = it


def ladder_maintenance(engine, comm):
    """Covariance / DE maintenance due at the start of the next iteration, with the cold shard's
    AM ring broadcast before a DE update and its eigen-factor after a covariance update."""
    it = engine.iteration
    if it == 0 or comm.maint_done == it:
        return
    due_cov, due_de = it % engine.cov_update == 0, it % engine.burn == 0
    if not (due_cov or due_de):
        return
    torch = comm.torch
    with comm._on_stream():
        if due_de and comm.world > 1:
            ladder_am_progress(engine, comm, flush=True)
            for w in comm.am_works:
                w.wait()
            comm.am_works = []
        engine.maintain()
        if due_cov and comm.world > 1:
            if hasattr(engine, "factor_dev") and comm.device.type == "cuda":
                # in place, device to device: no host copy, the host does not wait
                if comm._factor is None:
                    (pu, nu), (ps, ns) = engine.factor_dev()
                    comm._factor = (comm.alias(pu, nu), comm.alias(ps, ns))
                for t in comm._factor:
                    comm.bcast(t)
                if not comm.coldest:
                    engine.factor_refresh()
            else:
                if comm.coldest:
                    U, S = engine.factor()
                    t = torch.from_numpy(np.concatenate([U, S])).to(comm.device)
                else:
                    t = torch.empty(engine.usize + engine.ssize, dtype=torch.float64, device=comm.device)
                comm.bcast(t)
                if not comm.coldest:
                    f = t.cpu().numpy()
                    engine.set_factor(f[:engine.usize], f[engine.usize:])
    comm.maint_done = it


def _next_stop(it, niter_left, tskip, cov_update, burn):
    """Iterations to the next point a ladder shard must stop at: a swap, or a covariance / DE-history boundary (the
    factor broadcast and the AM-ring hand-over happen there, whether or not the boundary is a multiple of Tskip)."""
    return min(niter_left, tskip - it % tskip, cov_update - it % cov_update, burn - it % burn)


def run_ladder(engine, niter, comm, tskip):
    """``niter`` iterations of one ladder shard; collective over the shards of ``comm``."""
    done = 0
    while done < niter:
        it = engine.iteration
        ladder_maintenance(engine, comm)
        step = _next_stop(it, niter - done, tskip, engine.cov_update, engine.burn)
        engine.run(step)
        done += step
        if engine.swap_pending:
            ladder_swap(engine, comm)
        ladder_am_progress(engine, comm)
    return done


class HostMem(object):
    """Message memory for host engines (ctypes)."""

    def alloc(self, n):
        buf = (ctypes.c_double * n)()
        return ctypes.addressof(buf), buf

    def copy(self, dst, src, n):
        ctypes.memmove(dst, src, 8 * n)

    def sync(self, engine):
        pass


class CudaMem(object):
    """Message memory for ``_cabi.Engine`` shards living on one CUDA device (torch allocations)."""

    def __init__(self, device=0):
        import torch

        self.torch, self.device = torch, torch.device("cuda", device)

    def alloc(self, n):
        t = self.torch.zeros(n, dtype=self.torch.float64, device=self.device)
        self.torch.cuda.synchronize(self.device)
        return t.data_ptr(), t

    def copy(self, dst, src, n):
        a = self.torch.as_tensor(_CudaAlias(dst, n), device=self.device)
        a.copy_(self.torch.as_tensor(_CudaAlias(src, n), device=self.device))
        self.torch.cuda.synchronize(self.device)

    def sync(self, engine):
        engine.sync()


def connect_local_p2p(engines):
    """Shards of one process on one device: the peer-memory exchange with plain device addresses."""
    boxes = [e.p2p_open(want_handle=False)[1] for e in engines]
    seq0 = max(e.p2p_seq for e in engines)
    for g, e in enumerate(engines):
        e.p2p_connect(boxes[g + 1] if g + 1 < len(engines) else None, boxes[g - 1] if g > 0 else None, seq0)


def run_ladder_local(engines, niter, tskip, mem, p2p=False):
    """All shards of a ladder driven by ONE process (several shards on one device; also the reference
    implementation of the protocol for the tests): the same steps as ``run_ladder`` with the messages
    handed over directly.  ``mem`` is a ``HostMem`` / ``CudaMem``; ``p2p``: the shards were connected with
    ``connect_local_p2p`` and exchange the messages themselves (phase by phase: they share the device)."""
    G = len(engines)
    n = engines[0].swap_msg_doubles
    up = [mem.alloc(n) for _ in range(G)]
    carry = [mem.alloc(n) for _ in range(G)]
    done = 0
    while done < niter:
        it = engines[0].iteration
        due_cov = it > 0 and it % engines[0].cov_update == 0
        due_de = it > 0 and it % engines[0].burn == 0
        if due_de and G > 1:
            ptr0, na = engines[0].am_ring()
            mem.sync(engines[0])
            for e in engines[1:]:
                mem.sync(e)
                mem.copy(e.am_ring()[0], ptr0, na)
        if due_cov or due_de:
            for e in engines:
                e.maintain()
        if due_cov and G > 1:
            U, S = engines[0].factor()
            for e in engines[1:]:
                e.set_factor(U, S)
        step = _next_stop(it, niter - done, tskip, engines[0].cov_update, engines[0].burn)
        for e in engines:
            e.run(step)
        done += step
        if engines[0].swap_pending and p2p:
            for phase in (0, 1, 2):
                for g in (reversed(range(G)) if phase == 1 else range(G)):
                    engines[g].swap_p2p(phase)
            for e in engines:
                e.p2p_error()
        elif engines[0].swap_pending:
            for g in range(G - 1):
                engines[g].swap_pack_top(up[g][0])
                mem.sync(engines[g])
            for g in reversed(range(G)):
                engines[g].swap_sweep(carry[g + 1][0] if g < G - 1 else None, carry[g][0] if g > 0 else None)
                mem.sync(engines[g])
            for g in range(G):
                engines[g].swap_finish(up[g - 1][0] if g > 0 else None)
            for e in engines:
                mem.sync(e)
    return done
