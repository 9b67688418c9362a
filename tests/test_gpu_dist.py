"""GPU parity of the multi-GPU building blocks (rvn_dist_* through the C ABI).

The driver's GPU box has ONE GPU, so the partitioned schedule runs here as
`world` virtual ranks = threads, each with its own context on cuda:0, exchanging
through an in-process comm that does what the NCCL collectives do.  The real
NCCL path over 2+ GPUs is tests/dist_nccl_check.py (run with torchrun)."""
import threading

import numpy as np
import pytest
import torch

from raven_b200 import distributed, synth

pytestmark = pytest.mark.gpu


class ThreadComm:
    """all-to-all / all-gather / all-reduce between threads of one process."""

    class Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world

    def __init__(self, shared, rank):
        self.sh, self.rank, self.world = shared, rank, shared.world

    def _swap(self, item):
        torch.cuda.synchronize()
        self.sh.slots[self.rank] = item
        self.sh.barrier.wait()
        got = list(self.sh.slots)
        self.sh.barrier.wait()
        return got

    def all_to_all_v(self, tensors, send_counts):
        off = np.concatenate([[0], np.cumsum(send_counts)])
        tensors = [distributed._as_torch(t) for t in tensors]
        mine = [[t[off[d]:off[d + 1]].clone() for d in range(self.world)] for t in tensors]
        got = self._swap(mine)
        out = [torch.cat([got[src][i][self.rank] for src in range(self.world)])
               for i in range(len(tensors))]
        counts = [int(got[src][0][self.rank].shape[0]) for src in range(self.world)]
        torch.cuda.synchronize()
        return out, counts

    def all_gather_v(self, t):
        got = self._swap(t.clone())
        out = torch.cat(got)
        torch.cuda.synchronize()
        return out

    def all_reduce_sum(self, t):
        got = self._swap(t.clone())
        return torch.stack(got).sum(0)

    def all_reduce_max(self, t):
        got = self._swap(t.clone())
        return torch.stack(got).max(0).values

    def gather_objects(self, obj):
        return self._swap(obj)

    def begin_step(self):
        pass


def run_virtual(rs, world, freq, kmax, ib, qb, params=None, minhash=True, tiers=False):
    shared = ThreadComm.Shared(world)
    res, err = [None] * world, [None] * world

    def work(rank):
        try:
            de = distributed.DistEngine("cuda:0", ThreadComm(shared, rank), **(params or {}))
            if tiers:  # the tiered index also on inputs below its default size
                de.engine.set_option("tier_min_records", 0)
            de.upload(rs)
            share = de.find_overlaps_and_create_piles(freq, kmax, minhash, ib, qb)
            assert share["ovl_off"].size - 1 == len(range(rank, rs.n, world))
            res[rank] = distributed.assemble(share, de.comm)
            de.engine.close()
        except BaseException as e:  # noqa: BLE001
            err[rank] = e
            shared.barrier.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    for e in err:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in err:
        if e is not None:
            raise e
    return res


def check(res, want):
    for got in res:
        assert np.array_equal(got["ovl_off"], want["ovl_off"])
        assert np.array_equal(got["overlaps"], want["overlaps"])
        assert np.array_equal(got["pile"], want["pile"])
        assert int(got["num_mapped"]) == int(np.asarray(want["num_mapped"]).ravel()[0])


@pytest.mark.parametrize("minhash", [False, True])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_virtual_ranks_equal_oracle(oracle, world, minhash):
    rs = synth.make_reads(60_000, 150, 6000, seed=11)
    want = oracle.stage1(oracle.engine(15, 5), oracle.reads(rs), 0.001, 16, minhash)
    assert want["overlaps"].shape[0] > 200
    res = run_virtual(rs, world, 0.001, 16, 0, 0, minhash=minhash)
    check(res, want)
    assert list(res[0]["occurrences"]) == list(want["occurrences"])


@pytest.mark.parametrize("minhash", [False, True])
@pytest.mark.parametrize("world", [2, 4])
def test_virtual_ranks_multi_batch_schedule(oracle, world, minhash):
    rs = synth.make_reads(40_000, 160, 4000, seed=12)
    ib, qb = 200_000, 70_000
    want = oracle.stage1(oracle.engine(15, 5), oracle.reads(rs), 0.01, 8, minhash, ib, qb)
    assert len(want["occurrences"]) >= 3
    res = run_virtual(rs, world, 0.01, 8, ib, qb, minhash=minhash)
    check(res, want)
    assert list(res[0]["occurrences"]) == list(want["occurrences"])


@pytest.mark.parametrize("world", [2, 3])
def test_virtual_ranks_tiered_index(oracle, world):
    """The index slices of the partitioned run with the tiers of stage 1: the bound is
    the maximum over the ranks of their largest micromizer value (one all-reduce per
    index batch). Single batch and the multi-batch schedule, low frequency cut."""
    rs = synth.make_reads(60_000, 150, 6000, seed=11)
    want = oracle.stage1(oracle.engine(15, 5), oracle.reads(rs), 0.001, 16, False)
    res = run_virtual(rs, world, 0.001, 16, 0, 0, minhash=False, tiers=True)
    check(res, want)
    assert list(res[0]["occurrences"]) == list(want["occurrences"])
    rs = synth.make_reads(40_000, 160, 4000, seed=12)
    ib, qb = 200_000, 70_000
    want = oracle.stage1(oracle.engine(15, 5), oracle.reads(rs), 0.02, 8, False, ib, qb)
    res = run_virtual(rs, world, 0.02, 8, ib, qb, minhash=False, tiers=True)
    check(res, want)
    assert list(res[0]["occurrences"]) == list(want["occurrences"])


def test_virtual_ranks_equal_single_gpu_lambda(gpu_engine, lambda_reads):
    gpu_engine.configure(k=15, w=5)
    gpu_engine.upload(lambda_reads)
    want = gpu_engine.find_overlaps_and_create_piles(0.001, 32, True)
    res = run_virtual(lambda_reads, 2, 0.001, 32, 0, 0)
    check(res, want)


def test_more_ranks_than_reads(oracle):
    rs = synth.make_reads(20_000, 3, 6000, seed=2)
    want = oracle.stage1(oracle.engine(15, 5), oracle.reads(rs), 0.001, 32, True)
    check(run_virtual(rs, 4, 0.001, 32, 0, 0), want)


def test_nccl_two_gpus():
    """The same schedule over real NCCL when the box has more than one GPU."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU: the NCCL path is exercised by bench.py --gpus N")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29533",
         os.path.join(here, "dist_nccl_check.py")], capture_output=True, text=True,
        timeout=600)
    assert "DIST_NCCL_PARITY PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
