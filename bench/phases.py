"""Phase timings of one stage-1 pass (development aid; bench.py is the contract)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth  # noqa: E402
from raven_b200 import engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=200_000)
ap.add_argument("--genome", type=int, default=50_000_000)
ap.add_argument("--mean", type=int, default=10_000)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--minhash", type=int, default=0)
a = ap.parse_args()

t = time.time()
rs = synth.make_reads(20260924, a.genome, a.reads, a.mean)
print(f"generated {rs.n} reads, {rs.bases/1e9:.3f} Gbp in {time.time()-t:.1f}s "
      f"({os.cpu_count()} cpus)", flush=True)
eng = engine.Engine(0)
eng.configure(15, 5)
t = time.time()
eng.upload(rs)
print(f"upload {time.time()-t:.3f}s", flush=True)
for s in range(a.steps):
    eng.set_option("reset_stats", 1)
    t = time.time()
    eng.find_overlaps_and_create_piles(0.001, 32, bool(a.minhash), fetch=False)
    dt = time.time() - t
    st = eng.stats()
    tm = eng.timings()
    print(json.dumps(dict(step=s, wall_s=round(dt, 4), stats=st,
                          ms={k: round(v, 3) for k, v in tm.items()})), flush=True)
