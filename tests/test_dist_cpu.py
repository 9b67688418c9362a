"""N>1 host logic on CPU: the exchange schedule of raven_b200/distributed.py
over gloo (world_size 2 and 3) with numpy + oracle steps, against the
single-process oracle port of raven::FindOverlapsAndCreatePiles."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cfg, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_lib
        from dist_numpy_steps import NumpySteps
        from raven_b200 import distributed, synth
        rs = synth.make_reads(**cfg["reads"])
        steps = NumpySteps(oracle_lib.Oracle(), rs)
        res = distributed.find_overlaps_and_create_piles(
            steps, rs.lens, cfg["freq"], cfg["kmax"], cfg["minhash"], cfg["ib"], cfg["qb"])
        share = int(res["ovl_off"].size) - 1
        assert share == len(range(rank, rs.n, world))
        res = distributed.assemble(res)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"),
                 **{k: np.asarray(v) for k, v in res.items()})
    finally:
        dist.destroy_process_group()


CASES = [
    dict(reads=dict(genome_len=40_000, n_reads=90, mean_len=5000, seed=3), freq=0.001,
         kmax=16, ib=0, qb=0, minhash=True),
    dict(reads=dict(genome_len=40_000, n_reads=90, mean_len=5000, seed=4), freq=0.001,
         kmax=16, ib=0, qb=0, minhash=False),
    # several index batches and query flushes; truncation active
    dict(reads=dict(genome_len=30_000, n_reads=100, mean_len=4000, seed=5), freq=0.01,
         kmax=8, ib=150_000, qb=60_000, minhash=True),
    dict(reads=dict(genome_len=30_000, n_reads=100, mean_len=4000, seed=6), freq=0.01,
         kmax=8, ib=150_000, qb=60_000, minhash=False),
]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_schedule_over_gloo_equals_oracle(oracle, tmp_path, world, case):
    from raven_b200 import synth
    cfg = CASES[case]
    mp.spawn(_worker, args=(world, _free_port(), cfg, str(tmp_path)), nprocs=world,
             join=True)
    rs = synth.make_reads(**cfg["reads"])
    want = oracle.stage1(oracle.engine(15, 5), oracle.reads(rs), cfg["freq"], cfg["kmax"],
                         cfg["minhash"], cfg["ib"] or 1 << 32, cfg["qb"] or 1 << 30)
    assert want["overlaps"].shape[0] > 50
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        assert np.array_equal(got["occurrences"], want["occurrences"])
        assert np.array_equal(got["ovl_off"], want["ovl_off"])
        assert np.array_equal(got["overlaps"], want["overlaps"])
        assert np.array_equal(got["pile"], want["pile"])
        assert int(got["num_mapped"]) == int(want["num_mapped"][0])


def test_partitions():
    from raven_b200 import distributed as d
    rng = np.random.default_rng(1)
    lens = rng.integers(100, 20_000, 1000)
    for parts in (1, 2, 3, 8):
        sb = d.sketch_bounds(lens, parts)
        assert sb[0] == 0 and sb[-1] == len(lens) and len(sb) == parts + 1
        assert all(x <= y for x, y in zip(sb, sb[1:]))
        share = [lens[a:b].sum() / lens.sum() for a, b in zip(sb, sb[1:])]
        assert max(share) - min(share) < 0.02
    def loop(lens, ib):  # construct.cc:36-41 as written
        ib = ib or (1 << 32)
        out, bases, j, n = [], 0, 0, len(lens)
        for i in range(n):
            bases += int(lens[i])
            if i != n - 1 and bases < ib:
                continue
            bases = 0
            out.append((j, i + 1))
            j = i + 1
        return out
    for _ in range(200):
        lens = rng.integers(0, 30, int(rng.integers(0, 40)))
        ib = int(rng.integers(0, 60))
        assert d.index_batches(lens, ib) == loop(lens, ib)
    assert d.index_batches([5, 5, 5, 5, 5], 10) == [(0, 2), (2, 4), (4, 5)]
    assert d.index_batches([5, 5], 0) == [(0, 2)]
    assert d.sketch_bounds([], 2) == [0, 0, 0]
