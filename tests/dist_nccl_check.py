"""Real multi-GPU parity check (NCCL): run as
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
      --master-addr 127.0.0.1 --master-port 29533 tests/dist_nccl_check.py
Every rank runs the partitioned stage 1 and compares the (replicated) result
with the single-GPU path computed on its own GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raven_b200 import distributed, engine, synth  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = [
        (dict(genome_len=300_000, n_reads=1500, mean_len=8000, seed=21), 0.001, 32, 0, 0),
        (dict(genome_len=100_000, n_reads=600, mean_len=5000, seed=22), 0.01, 8,
         1_200_000, 400_000),
    ]
    for case_no, (reads, freq, kmax, ib, qb) in enumerate(cases * 2):
        minhash = case_no >= len(cases)
        rs = synth.make_reads(**reads)
        single = engine.Engine(device=local)
        single.upload(rs)
        want = single.find_overlaps_and_create_piles(freq, kmax, minhash, ib, qb)
        single.close()
        de = distributed.DistEngine(f"cuda:{local}")
        de.upload(rs)
        same = True
        for attempt in range(3):  # the arena is sized after the first pass
            got = distributed.assemble(
                de.find_overlaps_and_create_piles(freq, kmax, minhash, ib, qb))
            same = same and all(np.array_equal(got[k], want[k])
                                for k in ("ovl_off", "overlaps", "pile"))
            same = same and int(got["num_mapped"]) == int(want["num_mapped"])
        print(f"rank {rank} [{de.exchange} {getattr(de.comm, 'stats', '')}]: "
              f"{int(want['ovl_off'][-1])} kept overlaps, "
              f"{int(want['num_mapped'])} mapped, identical={same}", flush=True)
        ok = ok and same
        de.engine.close()
    flag = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_NCCL_PARITY", "PASS" if flag.item() == 0 else "FAIL", flush=True)
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
