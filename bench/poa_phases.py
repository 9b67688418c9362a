"""POA windows/s of rvn_poa_batch on synthetic racon windows (development aid)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from raven_b200 import engine, synth  # noqa: E402


def tile_windows(w, reps):
    """Repeat a window batch `reps` times (flat layout)."""
    nseq = w["win_first"][-1]
    nb = w["seq_off"][-1]
    wf = np.concatenate([w["win_first"][:-1] + r * nseq for r in range(reps)] + [[reps * nseq]])
    so = np.concatenate([w["seq_off"][:-1] + r * nb for r in range(reps)] + [[reps * nb]])
    return dict(win_first=wf.astype(np.uint32), seq_off=so.astype(np.uint64),
                bases=np.tile(w["bases"], reps),
                quals=None if w["quals"] is None else np.tile(w["quals"], reps),
                seq_begin=np.tile(w["seq_begin"], reps), seq_end=np.tile(w["seq_end"], reps))


ap = argparse.ArgumentParser()
ap.add_argument("--distinct", type=int, default=512)
ap.add_argument("--reps", type=int, default=16)
ap.add_argument("--layers", type=int, default=30)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--cpu", type=int, default=0, help="also time the oracle on this many windows")
a = ap.parse_args()
t = time.time()
w0 = synth.make_windows(n_windows=a.distinct, backbone_len=500, layers=a.layers, seed=11)
w = tile_windows(w0, a.reps)
nw = a.distinct * a.reps
print(f"{nw} windows, {w['bases'].size/1e6:.1f} Mbases, generated in {time.time()-t:.1f}s", flush=True)
eng = engine.Engine(0)
for s in range(a.steps):
    t = time.time()
    r = eng.poa_batch(w, want_coverage=False)
    dt = time.time() - t
    print(json.dumps(dict(step=s, wall_s=round(dt, 4), windows_per_s=round(nw / dt, 1),
                          gcups=round(r["cells"] / dt / 1e9, 2),
                          ms={k: round(v, 2) for k, v in eng.timings().items()})), flush=True)
if a.cpu:
    import oracle_lib
    O = oracle_lib.Oracle()
    sub = tile_windows(w0, 1)
    sub_n = min(a.cpu, a.distinct)
    t = time.time()
    r = O.poa_batch(w0, threads=os.cpu_count())
    dt = time.time() - t
    print(json.dumps(dict(cpu_windows=a.distinct, threads=os.cpu_count(), wall_s=round(dt, 3),
                          windows_per_s=round(a.distinct / dt, 1),
                          gcups=round(float(r["cells"].sum()) / dt / 1e9, 2))), flush=True)
