import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from bench import synth
from raven_b200 import engine
rs = synth.make_reads(7, 50_000_000, 200_000, 10_000)
words = torch.from_numpy(rs.words.view(np.int64)).pin_memory()
woff = torch.from_numpy(rs.word_off.view(np.int64)).pin_memory()
lens = torch.from_numpy(rs.lens.view(np.int32)).pin_memory()
class P: pass
p = P(); p.words = words.numpy().view(np.uint64); p.word_off = woff.numpy().view(np.uint64); p.lens = lens.numpy().view(np.uint32); p.n = rs.n
eng = engine.Engine(device=0); eng.configure(15, 5)
def sync(): torch.cuda.synchronize()
for mode in (0, 1, 0, 1):
    eng.set_option("async_upload", mode)
    eng.upload(p); eng.find_overlaps_and_create_piles(0.001, 32, False, fetch=False); sync()
    ts = []
    for _ in range(4):
        sync(); t0 = time.perf_counter(); eng.upload(p); t1 = time.perf_counter(); sync(); t2 = time.perf_counter()
        eng.find_overlaps_and_create_piles(0.001, 32, False, fetch=False); sync(); t3 = time.perf_counter()
        sync(); t4 = time.perf_counter(); eng.upload(p); eng.find_overlaps_and_create_piles(0.001, 32, False, fetch=False); sync(); t5 = time.perf_counter()
        ts.append((1e3*(t1-t0), 1e3*(t2-t0), 1e3*(t3-t2), 1e3*(t5-t4)))
    print("async", mode, "upload call / upload done / stage1 alone / upload+stage1 (ms):", np.round(np.median(np.array(ts), axis=0), 2))
    print("   phases", {k: round(v, 2) for k, v in eng.timings().items()})
