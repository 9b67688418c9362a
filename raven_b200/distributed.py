"""Multi-GPU ``raven::FindOverlapsAndCreatePiles`` (RavenLib/src/construct.cc:14-121):
one process per GPU, ``torch.distributed`` (NCCL over NVLink) for the exchanges,
the rvn_dist_* entry points of the C ABI for every compute step.

Sharding (DESIGN.md "Multi-GPU"):

* sketching      reads split into contiguous ranges of equal bases;
* the index      keys owned by ``value mod world`` - an all-to-all of 16-byte
                 minimizer records builds each rank's slice, an all-reduce of the
                 run-length histogram gives ONE global occurrence threshold;
* seed hits      found where the key lives, sent (all-to-all, 20 B per hit) to
                 the rank that owns the query read: read ``r`` belongs to rank
                 ``r mod world`` (with ``avoid_symmetric`` a read only meets higher
                 ids - the interleave gives every rank the same mix);
* chaining       per owned read, the single-GPU kernels;
* piles + lists  of the owned reads: every overlap also goes to the owner of its
                 rhs read (all-to-all, 32 B each); results stay sharded, bit-
                 identical to the single-GPU path (``assemble`` rebuilds the whole).

The exchange logic is written against a small "steps" interface so that the
same code runs over gloo on CPU tensors in the tests (tests/test_dist_cpu.py
plugs a numpy + oracle implementation in); ``CudaSteps`` is the product one.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from ._lib import OVLP, U16P, U32P, U64P


# ---------------------------------------------------------------- partitions
def sketch_bounds(lens, parts):
    """Contiguous read ranges of (nearly) equal bases: parts + 1 boundaries."""
    lens = np.asarray(lens, dtype=np.uint64)
    cum = np.concatenate([[0], np.cumsum(lens, dtype=np.uint64)])
    total = int(cum[-1])
    b = [0]
    for p in range(1, parts):
        b.append(int(np.searchsorted(cum, total * p // parts, side="left")))
    b.append(len(lens))
    return [min(max(x, b[i - 1] if i else 0), len(lens)) for i, x in enumerate(b)]


def index_batches(lens, index_batch_bases):
    """[j, i1) read ranges of the reference's index batches (construct.cc:36-41):
    a batch closes with the read that brings its bases to the threshold."""
    ib = int(index_batch_bases) or (1 << 32)
    n = len(lens)
    cum = np.concatenate([[0], np.cumsum(np.asarray(lens, dtype=np.uint64), dtype=np.uint64)])
    out, j = [], 0
    while j < n:
        i1 = int(np.searchsorted(cum, int(cum[j]) + ib, side="left"))
        i1 = min(max(i1, j + 1), n)
        out.append((j, i1))
        j = i1
    return out


# ---------------------------------------------------------------- collectives
class TorchComm:
    """The three exchanges of the schedule over ``torch.distributed`` (NCCL on
    the GPUs; gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def _exchange_counts(self, send_counts, device):
        s = torch.tensor(send_counts, dtype=torch.int64, device=device)
        r = torch.empty(self.world, dtype=torch.int64, device=device)
        dist.all_to_all_single(r, s, group=self.group)
        return [int(x) for x in r.tolist()]

    def all_to_all_v(self, tensors, send_counts):
        """Rows [sum(send_counts)] of every tensor, split by destination rank ->
        (rows received from every rank, concatenated in source-rank order;
        rows per source rank)."""
        tensors = [_as_torch(t) for t in tensors]
        device = tensors[0].device
        recv_counts = self._exchange_counts(send_counts, device)
        n_recv = sum(recv_counts)
        out = []
        for t in tensors:
            r = torch.empty((n_recv,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
            dist.all_to_all_single(r, t.contiguous(), recv_counts, list(send_counts),
                                   group=self.group)
            out.append(r)
        return out, recv_counts

    def all_gather_v(self, t):
        """Concatenation of every rank's rows, in rank order."""
        device = t.device
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=device)
        counts = [torch.empty(1, dtype=torch.int64, device=device)
                  for _ in range(self.world)]
        dist.all_gather(counts, n, group=self.group)
        counts = [int(c.item()) for c in counts]
        m = max(counts) if counts else 0
        if m == 0:
            return t[:0].clone()
        pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
        pad[: t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(parts, pad, group=self.group)
        return torch.cat([p[:c] for p, c in zip(parts, counts)])

    def all_reduce_sum(self, t):
        dist.all_reduce(t, group=self.group)
        return t

    def all_reduce_max(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def begin_step(self):
        pass

    def gather_objects(self, obj):
        """[obj of rank 0, obj of rank 1, ...] on every rank (host data)."""
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out


class P2PComm(TorchComm):
    """The all-to-alls as direct DMA writes into the destination's receive arena
    over NVLink (rvn_dist_put: CUDA IPC peer memory, copy engines) instead of
    NCCL send/recv.  The arena is bump-allocated over one step: every rank sees
    the whole count matrix, so every rank knows every arena's layout.  An
    exchange that does not fit goes through NCCL and the arena is re-sized at
    the next step boundary."""

    ALIGN = 256

    def __init__(self, engine, device, group=None, arena_bytes=1 << 28):
        super().__init__(group)
        self.e, self.lib, self.h = engine, engine.lib, engine.h
        self.device = torch.device(device)
        self.cap = 0
        self.want = int(arena_bytes)
        self.bump = [0] * self.world
        self.arena = 0
        self.stats = {"p2p": 0, "nccl": 0}
        self._resize()

    def _resize(self):
        err = None

        def call(rc):  # keep taking part in the collectives even if a call fails
            nonlocal err
            if rc != 0 and err is None:
                err = self.lib.rvn_last_error(self.h).decode()

        dist.barrier(group=self.group)  # nobody reads or writes an arena now
        call(self.lib.rvn_dist_arena_close_peers(self.h))
        dist.barrier(group=self.group)  # every mapping is closed: arenas may be freed
        handle = (C.c_uint8 * 64)()
        call(self.lib.rvn_dist_arena_export(self.h, self.want, handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        blob = (C.c_uint8 * (64 * self.world)).from_buffer_copy(b"".join(handles))
        if err is None:
            call(self.lib.rvn_dist_arena_import(self.h, self.world, self.rank, blob))
        ptr, cap = C.c_void_p(), C.c_uint64(0)
        call(self.lib.rvn_dist_arena(self.h, C.byref(ptr), C.byref(cap)))
        self.arena, self.cap = ptr.value or 0, cap.value
        bad = torch.tensor([0 if err is None else 1], device=self.device)
        dist.all_reduce(bad, group=self.group)
        if bad.item():
            raise RuntimeError("peer-memory exchange unavailable: " + (err or "on another rank"))

    def begin_step(self):
        need = max(self.bump)  # bytes the last step wanted in the fullest arena
        if need > self.cap:
            self.want = int(need * 1.25) + (1 << 20)
            self._resize()
        else:
            dist.barrier(group=self.group)  # the previous step is consumed everywhere
        self.bump = [0] * self.world

    def all_to_all_v(self, arrays, send_counts):
        world, me = self.world, self.rank
        s = torch.tensor(send_counts, dtype=torch.int64, device=self.device)
        m = torch.empty(world * world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(m, s, group=self.group)
        m = m.cpu().numpy().reshape(world, world)  # m[src][dst] rows
        recv_counts = [int(x) for x in m[:, me]]
        rows_to = m.sum(axis=0)
        rb = [a.row_bytes() if isinstance(a, DevArray) else
              a.element_size() * int(np.prod(a.shape[1:])) for a in arrays]
        # layout of this exchange in every destination arena
        base = [[0] * world for _ in arrays]
        fits = True
        for d in range(world):
            at = self.bump[d]
            for t, b in enumerate(rb):
                at = (at + self.ALIGN - 1) // self.ALIGN * self.ALIGN
                base[t][d] = at
                at += int(rows_to[d]) * b
            self.bump[d] = at
            fits = fits and at <= self.cap
        if not fits:  # same verdict on every rank
            self.stats["nccl"] += 1
            return super().all_to_all_v(arrays, send_counts)
        self.stats["p2p"] += 1
        before = m[:me, :].sum(axis=0)  # rows of lower ranks in every destination
        send_off = np.concatenate([[0], np.cumsum(send_counts)])
        for t, a in enumerate(arrays):
            src = a.data_ptr()
            for k in range(world):
                d = (me + k) % world
                nbytes = int(send_counts[d]) * rb[t]
                if nbytes:
                    self.e._check(self.lib.rvn_dist_put(
                        self.h, d, base[t][d] + int(before[d]) * rb[t],
                        C.c_void_p(src + int(send_off[d]) * rb[t]), nbytes))
        self.e._check(self.lib.rvn_dist_put_flush(self.h))
        dist.barrier(group=self.group)  # every rank's writes have landed
        n_recv = int(rows_to[me])
        out = []
        for t, a in enumerate(arrays):
            shape = (n_recv,) + tuple(a.shape[1:])
            out.append(DevArray(self.arena + base[t][me], shape, a.dtype, self.device))
        return out, recv_counts


# ---------------------------------------------------------------- the schedule
def find_overlaps_and_create_piles(steps, lens, frequency=0.001, max_overlaps=32,
                                   use_minhash=False, index_batch_bases=0,
                                   query_batch_bases=0, comm=None, fetch=True):
    """The reference's stage 1 over ``world`` ranks.  ``steps`` does the compute
    on this rank (CudaSteps); returns this rank's share - the overlap lists and
    piles of the reads ``rank, rank + world, ...`` (``assemble`` gives the whole)."""
    comm = comm or TorchComm()
    rank, world = comm.rank, comm.world
    lens = np.asarray(lens, dtype=np.uint32)
    n = len(lens)
    if not 0 <= frequency <= 1:
        raise ValueError("[ram::MinimizerEngine::Filter] error: invalid frequency")
    sb = sketch_bounds(lens, world)
    # RVN_DIST_TRACE=1: wall time per step of the schedule (adds a device sync
    # after each; for analysis, never for a reported number)
    trace = {} if os.environ.get("RVN_DIST_TRACE") else None
    t_last = [time.perf_counter()]

    def tick(name):
        if trace is None:
            return
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        now = time.perf_counter()
        trace[name] = trace.get(name, 0.0) + 1e3 * (now - t_last[0])
        t_last[0] = now

    comm.begin_step()
    steps.stage1_begin(world, rank)
    tick("begin")
    occurrences = []
    for (j, i1) in index_batches(lens, index_batch_bases):
        # -- 1. sketch own reads; records to the owners of their keys
        # (index: micromizers only with use_minhash, construct.cc:42-43;
        #  queries: always micromizers, construct.cc:62)
        lo, hi = max(sb[rank], j), min(sb[rank + 1], i1)
        lo, hi = (lo, hi) if lo < hi else (0, 0)
        val, org, cnt = steps.sketch_split(lo, hi, world, use_minhash)
        tick("sketch_split(index)")
        (ival, iorg), _ = comm.all_to_all_v([val, org], cnt)
        tick("a2a index records")
        qlo, qhi = sb[rank], min(sb[rank + 1], i1)
        qlo, qhi = (qlo, qhi) if qlo < qhi else (0, 0)
        if use_minhash and j == 0:  # (same decision on every rank)
            qval, qorg = ival, iorg  # one exchange serves index and queries
        else:
            val, org, cnt = steps.sketch_split(qlo, qhi, world, True)
            tick("sketch_split(query)")
            (qval, qorg), _ = comm.all_to_all_v([val, org], cnt)
            tick("a2a query records")

        # -- 2. index slice + ONE global occurrence threshold. Queries are micromizers
        # only: index records above the largest micromizer value of ANY read that can
        # query this batch are counted but not probe-able (the tiers of index.cu); the
        # bound is the maximum over the ranks (every rank knows its own query reads)
        limit = None
        if not use_minhash:
            local = min(steps.max_threshold(qlo, qhi), (1 << 62))
            t = torch.tensor([local], dtype=torch.int64, device=getattr(steps, "device", "cpu"))
            limit = int(comm.all_reduce_max(t).item())
            tick("index limit (all-reduce)")
        steps.build_index(ival, iorg, int(lens[j:i1].astype(np.uint64).sum()), limit)
        tick("build_index")
        hist, n_keys = steps.histogram()
        tot = torch.cat([hist, torch.tensor([n_keys], dtype=torch.int64,
                                            device=hist.device)])
        tot = comm.all_reduce_sum(tot).cpu().numpy().astype(np.uint64)
        occurrences.append(steps.set_occurrence(tot[:-1], int(tot[-1]), frequency))
        tick("filter (histogram + all-reduce)")

        # -- 3. seed hits where the key lives -> owner of the query read
        # (the received queries are sorted by read: ascending sketch ranges)
        grp, pos, lhs, cnt = steps.hits_split(qval, qorg, world, i1)
        tick("hits_split")
        (grp, pos, lhs), runs = comm.all_to_all_v([grp, pos, lhs], cnt)
        tick("a2a hits")

        # -- 4. chain the owned reads (one run of hits per source rank)
        steps.chain(grp, pos, lhs, runs, world, rank, i1)
        tick("chain")

        # -- 5. every overlap also to the owner of its rhs read; piles + lists of
        # the owned reads with the reference's flush schedule
        ovl, cnt = steps.overlaps_split(world, rank)
        (ovl,), runs = comm.all_to_all_v([ovl], cnt)
        tick("a2a overlaps")
        steps.stage1_add(ovl, runs, i1, max_overlaps, query_batch_bases)
        tick("piles + lists")
    steps.stage1_end()
    tick("end (D2H)")
    res = steps.stage1_results() if fetch else {}
    res.update(occurrences=occurrences, rank=rank, world=world, n_reads=n)
    if trace is not None:
        for k, v in getattr(steps, "call_ms", {}).items():
            trace["call:" + k] = v
        res["trace_ms"] = trace
    return res


def assemble(res, comm=None):
    """The complete stage-1 result (the layout of Engine.find_overlaps_and_
    create_piles) from the per-rank shares, on every rank."""
    comm = comm or TorchComm()
    keys = ("overlaps", "ovl_off", "pile", "pile_off", "num_mapped")
    shares = comm.gather_objects({k: res[k] for k in keys})
    n, world = res["n_reads"], res["world"]
    ocnt = np.zeros(n, dtype=np.uint64)
    pcnt = np.zeros(n, dtype=np.uint64)
    for r, sh in enumerate(shares):
        ocnt[r::world] = np.diff(sh["ovl_off"])
        pcnt[r::world] = np.diff(sh["pile_off"])
    ovl_off = np.concatenate([[0], np.cumsum(ocnt)]).astype(np.uint64)
    pile_off = np.concatenate([[0], np.cumsum(pcnt)]).astype(np.uint64)
    overlaps = np.zeros((int(ovl_off[-1]), 8), dtype=np.uint32)
    pile = np.zeros(int(pile_off[-1]), dtype=np.uint16)
    for r, sh in enumerate(shares):
        ids = np.arange(r, n, world)
        for src_off, dst_off, src, dst in ((sh["ovl_off"], ovl_off, sh["overlaps"], overlaps),
                                           (sh["pile_off"], pile_off, sh["pile"], pile)):
            cnt = np.diff(src_off).astype(np.int64)
            if cnt.sum() == 0:
                continue
            # destination index of every element of the share
            start = np.repeat(dst_off[ids].astype(np.int64), cnt)
            within = np.arange(int(cnt.sum())) - np.repeat(src_off[:-1].astype(np.int64), cnt)
            dst[start + within] = src
    return dict(overlaps=overlaps, ovl_off=ovl_off, pile=pile, pile_off=pile_off,
                num_mapped=int(sum(int(sh["num_mapped"]) for sh in shares)),
                occurrences=res["occurrences"])


# ---------------------------------------------------------------- CUDA steps
class _DevMem:
    """``__cuda_array_interface__`` view of context-owned device memory."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {
            "shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3,
            "strides": None}


_TYPESTR = {torch.int64: "<i8", torch.int32: "<i4"}


def _view(ptr, shape, typestr, dtype, device):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return torch.empty(shape, dtype=dtype, device=device)
    return torch.as_tensor(_DevMem(ptr, tuple(shape), typestr), device=device)


class DevArray:
    """A typed window of device memory (context- or arena-owned): just enough of
    the tensor interface for the schedule; ``torch()`` gives a real tensor view
    when an exchange goes through torch.distributed."""

    def __init__(self, ptr, shape, dtype, device):
        self.ptr, self.shape, self.dtype, self.device = ptr or 0, tuple(shape), dtype, device

    def data_ptr(self):
        return self.ptr

    def numel(self):
        return int(np.prod(self.shape))

    def row_bytes(self):
        return int(np.prod(self.shape[1:])) * torch.empty(0, dtype=self.dtype).element_size()

    def torch(self):
        return _view(self.ptr, self.shape, _TYPESTR[self.dtype], self.dtype, self.device)


def _as_torch(t):
    return t.torch() if isinstance(t, DevArray) else t


class CudaSteps:
    """The rvn_dist_* entry points over an ``Engine`` whose reads are uploaded.
    Tensors returned are views of context memory: consumed (sent) before the
    next call into the same step."""

    def __init__(self, engine, device):
        self.e = engine
        self.lib = engine.lib
        self.h = engine.h
        self.device = torch.device(device)
        self.call_ms = {}  # RVN_DIST_TRACE: wall time inside the C calls
        if os.environ.get("RVN_DIST_TRACE"):
            for name in ("sketch_split", "build_index", "hits_split", "chain",
                         "overlaps_split", "stage1_add", "stage1_end"):
                setattr(self, name, self._timed(name, getattr(self, name)))

    def _timed(self, name, fn):
        def wrapped(*a):
            t = time.perf_counter()
            out = fn(*a)
            self.call_ms[name] = self.call_ms.get(name, 0.0) + 1e3 * (time.perf_counter() - t)
            return out
        return wrapped

    def _p(self, t):  # torch tensor or DevArray
        return C.c_void_p(t.data_ptr() if t.numel() else 0)

    def sketch_split(self, first, last, parts, minhash):
        v, o = C.c_void_p(), C.c_void_p()
        cnt = (C.c_uint64 * parts)()
        self.e._check(self.lib.rvn_dist_sketch_split(
            self.h, first, last, int(minhash), parts, C.byref(v), C.byref(o), cnt))
        cnt = [int(x) for x in cnt]
        n = sum(cnt)
        return (DevArray(v.value, (n,), torch.int64, self.device),
                DevArray(o.value, (n,), torch.int64, self.device), cnt)

    def build_index(self, val, org, bases, limit=None):
        self._keep_index = (val, org)
        lim = 0xFFFFFFFFFFFFFFFF if limit is None else int(limit)
        self.e._check(self.lib.rvn_dist_index_limited(self.h, self._p(val), self._p(org),
                                                      val.numel(), bases, lim))

    def max_threshold(self, first, last):
        v = C.c_uint64(0)
        self.e._check(self.lib.rvn_dist_max_threshold(self.h, first, last, C.byref(v)))
        return int(v.value)

    def histogram(self):
        d, nb, nk = C.c_void_p(), C.c_uint32(0), C.c_uint64(0)
        self.e._check(self.lib.rvn_dist_histogram(self.h, C.byref(d), C.byref(nb),
                                                  C.byref(nk)))
        return _view(d.value, (nb.value,), "<i8", torch.int64, self.device).clone(), nk.value

    def set_occurrence(self, hist, n_keys, frequency):
        h = np.ascontiguousarray(hist, dtype=np.uint64)
        occ = C.c_uint32(0)
        self.e._check(self.lib.rvn_dist_set_occurrence(
            self.h, h.ctypes.data_as(U64P), n_keys, float(frequency), C.byref(occ)))
        return occ.value

    def hits_split(self, qval, qorg, parts, n_query):
        g, p, l = C.c_void_p(), C.c_void_p(), C.c_void_p()
        cnt = (C.c_uint64 * parts)()
        self.e._check(self.lib.rvn_dist_hits_split(
            self.h, self._p(qval), self._p(qorg), qval.numel(), 1, 1, parts, n_query,
            C.byref(g), C.byref(p), C.byref(l), cnt))
        cnt = [int(x) for x in cnt]
        n = sum(cnt)
        return (DevArray(g.value, (n,), torch.int64, self.device),
                DevArray(p.value, (n,), torch.int64, self.device),
                DevArray(l.value, (n,), torch.int32, self.device), cnt)

    @staticmethod
    def _runs(counts):
        return (C.c_uint64 * (len(counts) + 1))(0, *np.cumsum(counts).tolist())

    def chain(self, grp, pos, lhs, run_counts, parts, rank, n_query):
        o, n = C.c_void_p(), C.c_uint64(0)
        self.e._check(self.lib.rvn_dist_chain(
            self.h, self._p(grp), self._p(pos), self._p(lhs), grp.numel(), len(run_counts),
            self._runs(run_counts), parts, rank, n_query, C.byref(o), C.byref(n)))
        return n.value

    def overlaps_split(self, parts, rank):
        o = C.c_void_p()
        cnt = (C.c_uint64 * parts)()
        self.e._check(self.lib.rvn_dist_overlaps_split(self.h, parts, rank, C.byref(o), cnt))
        cnt = [int(x) for x in cnt]
        return DevArray(o.value, (sum(cnt), 8), torch.int32, self.device), cnt

    def stage1_begin(self, parts, rank):
        self.e._check(self.lib.rvn_dist_stage1_begin(self.h, parts, rank))

    def stage1_add(self, ovl, run_counts, n_query, max_overlaps, query_batch_bases):
        self.e._check(self.lib.rvn_dist_stage1_add(
            self.h, self._p(ovl), ovl.shape[0], len(run_counts), self._runs(run_counts),
            n_query, max_overlaps, query_batch_bases))

    def stage1_end(self):
        self.e._check(self.lib.rvn_dist_stage1_end(self.h))

    def stage1_results(self):
        from .engine import _arr
        o, off, p, poff = OVLP(), U64P(), U16P(), U64P()
        n_own, nm = C.c_uint32(0), C.c_uint64(0)
        self.e._check(self.lib.rvn_dist_stage1_results(
            self.h, C.byref(o), C.byref(off), C.byref(p), C.byref(poff), C.byref(n_own),
            C.byref(nm)))
        ovl_off = _arr(off, n_own.value + 1, np.uint64)
        pile_off = _arr(poff, n_own.value + 1, np.uint64)
        return dict(overlaps=_arr(o, int(ovl_off[-1]) * 8, np.uint32).reshape(-1, 8),
                    ovl_off=ovl_off, pile=_arr(p, int(pile_off[-1]), np.uint16),
                    pile_off=pile_off, num_mapped=nm.value)


class DistEngine:
    """One rank of the multi-GPU stage 1: an ``Engine`` on this rank's GPU whose
    work stream is also torch's current stream while the schedule runs, so the
    NCCL exchanges and the kernels order on one stream."""

    def __init__(self, device, comm=None, exchange="p2p", **params):
        """comm: an exchange object (tests), or None for torch.distributed's
        default group with exchange = "p2p" (peer-memory DMA, falls back to NCCL
        if CUDA IPC is unavailable) or "nccl"."""
        from .engine import Engine
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.engine = Engine(self.device.index or 0, stream=self.stream.cuda_stream)
        if params:
            self.engine.configure(**params)
        self.lens = None
        self.exchange = "custom"
        if comm is None:
            comm, self.exchange = TorchComm(), "nccl"
            if exchange == "p2p" and comm.world > 1:
                try:  # (fails on every rank together: no IPC / no peer access)
                    with torch.cuda.device(self.device):
                        comm, self.exchange = P2PComm(self.engine, self.device), "p2p"
                except RuntimeError as e:
                    self.p2p_error = str(e)
        self.comm = comm

    def upload(self, rs):
        """Lengths of all reads, bases of this rank's sketch range only."""
        self.lens = np.asarray(rs.lens, dtype=np.uint32)
        sb = sketch_bounds(self.lens, self.comm.world)
        self.engine.upload(rs, resident=(sb[self.comm.rank], sb[self.comm.rank + 1]))

    def find_overlaps_and_create_piles(self, freq=0.001, max_overlaps=32,
                                       use_minhash=False, index_batch_bases=0,
                                       query_batch_bases=0, fetch=True):
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            return find_overlaps_and_create_piles(
                CudaSteps(self.engine, self.device), self.lens, freq, max_overlaps,
                use_minhash, index_batch_bases, query_batch_bases, self.comm, fetch)
