"""Multi-GPU ``raven::FindOverlapsAndCreatePiles`` (RavenLib/src/construct.cc:14-121):
one process per GPU, ``torch.distributed`` (NCCL over NVLink) for the exchanges,
the rvn_dist_* entry points of the C ABI for every compute step.

Sharding (DESIGN.md "Multi-GPU"):

* sketching      reads split into contiguous ranges of equal bases;
* the index      keys owned by ``value mod world`` - an all-to-all of 16-byte
                 minimizer records builds each rank's slice, an all-reduce of the
                 run-length histogram gives ONE global occurrence threshold;
* seed hits      found where the key lives, sent (all-to-all, 20 B per hit) to
                 the rank that owns the query read;
* chaining       per owned read - ranges balanced for the triangular work of
                 ``avoid_symmetric`` (read i only meets reads above it);
* piles + lists  every rank gets all overlaps (all-gather, 32 B each) and runs
                 the cheap tail itself: the result is complete on every rank and
                 bit-identical to the single-GPU path.

The exchange logic is written against a small "steps" interface so that the
same code runs over gloo on CPU tensors in the tests (tests/test_dist_cpu.py
plugs a numpy + oracle implementation in); ``CudaSteps`` is the product one.
"""
from __future__ import annotations

import ctypes as C
import math

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


def chain_bounds(lens, parts):
    """Contiguous read ranges of equal CHAIN work.  With avoid_symmetric a read
    only keeps hits against higher ids, so the work of the read at base
    fraction x falls like (1 - x): the boundary of part p sits at
    x = 1 - sqrt(1 - p / parts)."""
    lens = np.asarray(lens, dtype=np.uint64)
    cum = np.concatenate([[0], np.cumsum(lens, dtype=np.uint64)])
    total = int(cum[-1])
    b = [0]
    for p in range(1, parts):
        x = 1.0 - math.sqrt(1.0 - p / parts)
        b.append(int(np.searchsorted(cum, int(total * x), side="left")))
    b.append(len(lens))
    for i in range(1, len(b)):
        b[i] = min(max(b[i], b[i - 1]), len(lens))
    return b


def index_batches(lens, index_batch_bases):
    """[j, i1) read ranges of the reference's index batches (construct.cc:36-41)."""
    ib = index_batch_bases or (1 << 32)
    out, bases, j = [], 0, 0
    n = len(lens)
    for i in range(n):
        bases += int(lens[i])
        if i != n - 1 and bases < ib:
            continue
        bases = 0
        out.append((j, i + 1))
        j = i + 1
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
        rows received from every rank, concatenated in source-rank order."""
        device = tensors[0].device
        recv_counts = self._exchange_counts(send_counts, device)
        n_recv = sum(recv_counts)
        out = []
        for t in tensors:
            r = torch.empty((n_recv,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
            dist.all_to_all_single(r, t.contiguous(), recv_counts, list(send_counts),
                                   group=self.group)
            out.append(r)
        return out

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


# ---------------------------------------------------------------- the schedule
def find_overlaps_and_create_piles(steps, lens, frequency=0.001, max_overlaps=32,
                                   index_batch_bases=0, query_batch_bases=0,
                                   comm=None):
    """The reference's stage 1 over ``world`` ranks.  ``steps`` does the compute
    on this rank (CudaSteps); returns what ``steps.stage1_results()`` returns -
    the same on every rank."""
    comm = comm or TorchComm()
    rank, world = comm.rank, comm.world
    lens = np.asarray(lens, dtype=np.uint32)
    n = len(lens)
    if not 0 <= frequency <= 1:
        raise ValueError("[ram::MinimizerEngine::Filter] error: invalid frequency")
    sb = sketch_bounds(lens, world)
    cb = chain_bounds(lens, world)
    steps.stage1_begin()
    occurrences = []
    for (j, i1) in index_batches(lens, index_batch_bases):
        # -- 1. sketch own reads; records to the owners of their keys
        lo, hi = max(sb[rank], j), min(sb[rank + 1], i1)
        lo, hi = (lo, hi) if lo < hi else (0, 0)
        val, org, cnt = steps.sketch_split(lo, hi, world)
        ival, iorg = comm.all_to_all_v([val, org], cnt)
        qlo, qhi = sb[rank], min(sb[rank + 1], i1)
        qlo, qhi = (qlo, qhi) if qlo < qhi else (0, 0)
        if j == 0:  # (same decision on every rank) first batch: queries == index reads
            qval, qorg = ival, iorg  # one exchange serves index and queries
        else:
            val, org, cnt = steps.sketch_split(qlo, qhi, world)
            qval, qorg = comm.all_to_all_v([val, org], cnt)

        # -- 2. index slice + ONE global occurrence threshold
        steps.build_index(ival, iorg, int(lens[j:i1].astype(np.uint64).sum()))
        hist, n_keys = steps.histogram()
        tot = torch.cat([hist, torch.tensor([n_keys], dtype=torch.int64,
                                            device=hist.device)])
        tot = comm.all_reduce_sum(tot).cpu().numpy().astype(np.uint64)
        occurrences.append(steps.set_occurrence(tot[:-1], int(tot[-1]), frequency))

        # -- 3. seed hits where the key lives -> owner of the query read
        qb_ = [min(b, i1) for b in cb]
        grp, pos, lhs, cnt = steps.hits_split(qval, qorg, qb_)
        grp, pos, lhs = comm.all_to_all_v([grp, pos, lhs], cnt)

        # -- 4. chain the owned reads; everybody gets every overlap
        ovl, per_read = steps.chain(grp, pos, lhs, qb_[rank], qb_[rank + 1])
        all_ovl = comm.all_gather_v(ovl)
        all_cnt = comm.all_gather_v(per_read)
        off = np.zeros(i1 + 1, dtype=np.uint64)
        np.cumsum(all_cnt.cpu().numpy().astype(np.uint64), out=off[1:])
        assert int(off[-1]) == all_ovl.shape[0]

        # -- 5. piles + per-read lists (replicated; the reference's flush schedule)
        steps.stage1_add(all_ovl, off, i1, max_overlaps, query_batch_bases)
    steps.stage1_end()
    res = steps.stage1_results(n)
    res["occurrences"] = occurrences
    return res


# ---------------------------------------------------------------- CUDA steps
class _DevMem:
    """``__cuda_array_interface__`` view of context-owned device memory."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {
            "shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3,
            "strides": None}


def _view(ptr, shape, typestr, dtype, device):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return torch.empty(shape, dtype=dtype, device=device)
    return torch.as_tensor(_DevMem(ptr, tuple(shape), typestr), device=device)


class CudaSteps:
    """The rvn_dist_* entry points over an ``Engine`` whose reads are uploaded.
    Tensors returned are views of context memory: consumed (sent) before the
    next call into the same step."""

    def __init__(self, engine, device):
        self.e = engine
        self.lib = engine.lib
        self.h = engine.h
        self.device = torch.device(device)

    def _p(self, t):
        return C.c_void_p(t.data_ptr() if t.numel() else 0)

    def sketch_split(self, first, last, parts):
        v, o = C.c_void_p(), C.c_void_p()
        cnt = (C.c_uint64 * parts)()
        self.e._check(self.lib.rvn_dist_sketch_split(
            self.h, first, last, 1, parts, C.byref(v), C.byref(o), cnt))
        cnt = [int(x) for x in cnt]
        n = sum(cnt)
        return (_view(v.value, (n,), "<i8", torch.int64, self.device),
                _view(o.value, (n,), "<i8", torch.int64, self.device), cnt)

    def build_index(self, val, org, bases):
        self._keep_index = (val, org)
        self.e._check(self.lib.rvn_dist_index(self.h, self._p(val), self._p(org),
                                              val.numel(), bases))

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

    def hits_split(self, qval, qorg, bounds):
        parts = len(bounds) - 1
        b = (C.c_uint32 * (parts + 1))(*bounds)
        g, p, l = C.c_void_p(), C.c_void_p(), C.c_void_p()
        cnt = (C.c_uint64 * parts)()
        self.e._check(self.lib.rvn_dist_hits_split(
            self.h, self._p(qval), self._p(qorg), qval.numel(), 1, 1, parts, b,
            C.byref(g), C.byref(p), C.byref(l), cnt))
        cnt = [int(x) for x in cnt]
        n = sum(cnt)
        return (_view(g.value, (n,), "<i8", torch.int64, self.device),
                _view(p.value, (n,), "<i8", torch.int64, self.device),
                _view(l.value, (n,), "<i4", torch.int32, self.device), cnt)

    def chain(self, grp, pos, lhs, first, last):
        o, c, n = C.c_void_p(), C.c_void_p(), C.c_uint64(0)
        self.e._check(self.lib.rvn_dist_chain(
            self.h, self._p(grp), self._p(pos), self._p(lhs), grp.numel(), first, last,
            C.byref(o), C.byref(c), C.byref(n)))
        return (_view(o.value, (n.value, 8), "<i4", torch.int32, self.device),
                _view(c.value, (last - first,), "<i4", torch.int32, self.device))

    def stage1_begin(self):
        self.e._check(self.lib.rvn_dist_stage1_begin(self.h))

    def stage1_add(self, ovl, off, n_query, max_overlaps, query_batch_bases):
        off = np.ascontiguousarray(off, dtype=np.uint64)
        self.e._check(self.lib.rvn_dist_stage1_add(
            self.h, self._p(ovl), off.ctypes.data_as(U64P), n_query, max_overlaps,
            query_batch_bases))

    def stage1_end(self):
        self.e._check(self.lib.rvn_dist_stage1_end(self.h))

    def stage1_results(self, n):
        from .engine import _arr
        o, off, p, poff, nm = OVLP(), U64P(), U16P(), U64P(), C.c_uint64(0)
        self.e._check(self.lib.rvn_stage1_results(self.h, C.byref(o), C.byref(off),
                                                  C.byref(p), C.byref(poff),
                                                  C.byref(nm)))
        ovl_off = _arr(off, n + 1, np.uint64)
        pile_off = _arr(poff, n + 1, np.uint64)
        return dict(overlaps=_arr(o, int(ovl_off[-1]) * 8, np.uint32).reshape(-1, 8),
                    ovl_off=ovl_off, pile=_arr(p, int(pile_off[-1]), np.uint16),
                    pile_off=pile_off, num_mapped=nm.value)


class DistEngine:
    """One rank of the multi-GPU stage 1: an ``Engine`` on this rank's GPU whose
    work stream is also torch's current stream while the schedule runs, so the
    NCCL exchanges and the kernels order on one stream."""

    def __init__(self, device, comm=None, **params):
        from .engine import Engine
        self.device = torch.device(device)
        self.comm = comm
        self.stream = torch.cuda.Stream(self.device)
        self.engine = Engine(self.device.index or 0, stream=self.stream.cuda_stream)
        if params:
            self.engine.configure(**params)
        self.lens = None

    def upload(self, rs):
        self.engine.upload(rs)
        self.lens = np.asarray(rs.lens, dtype=np.uint32)

    def find_overlaps_and_create_piles(self, freq=0.001, max_overlaps=32,
                                       index_batch_bases=0, query_batch_bases=0):
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            return find_overlaps_and_create_piles(
                CudaSteps(self.engine, self.device), self.lens, freq, max_overlaps,
                index_batch_bases, query_batch_bases, self.comm)
