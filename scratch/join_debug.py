import sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from raven_b200 import engine, seqio
g = np.load('tests/golden/lambda_reads.npz', allow_pickle=True)
rs = seqio.ReadSet(g['words'], g['word_off'], g['lens'])
n = rs.n
eng = engine.Engine(device=0)
eng.configure(15, 5)
eng.upload(rs)
eng.set_option("keep_hits", 1)
res = {}
for join in (0, 1):
    eng.set_option("self_join", join)
    eng.minimize(0, n, False)
    occ = eng.filter(0.001)
    eng.map(0, n, True, True, True)          # first call builds the micromizers
    r = eng.map(0, n, True, True, True)
    h = eng.map_hits(n)
    off = h["hit_off"].astype(np.int64)
    rid = np.repeat(np.arange(n), np.diff(off))
    res[join] = set(zip(rid.tolist(), h["group"].tolist(), h["positions"].tolist()))
    print("join", join, "occ", occ, "hits", len(h["group"]), "ovl", len(r["overlaps"]))
a, b = res[0], res[1]
extra, missing = sorted(b - a), sorted(a - b)
print("extra", len(extra), "missing", len(missing))
full = eng.sketch(0, n, False); mic = eng.sketch(0, n, True)
val = full["value"]; org = full["origin"]
cnt = {}
for v in val.tolist(): cnt[v] = cnt.get(v, 0) + 1
micset = set(zip((mic["origin"] >> np.uint64(32)).tolist(), ((mic["origin"] & np.uint64(0xFFFFFFFF)) >> np.uint64(1)).tolist()))
lookup = {((int(o) >> 32), (int(o) & 0xFFFFFFFF) >> 1): int(v) for v, o in zip(val.tolist(), org.tolist())}
for (r, grp, pos) in extra[:12]:
    lp, rp = pos >> 32, pos & 0xFFFFFFFF
    rhs = grp >> 33
    v = lookup.get((r, lp))
    print("extra: lhs", r, lp, "rhs", rhs, rp, "value", v, "run", cnt.get(v), "lhs is micro", (r, lp) in micset)
for (r, grp, pos) in missing[:12]:
    lp, rp = pos >> 32, pos & 0xFFFFFFFF
    v = lookup.get((r, lp))
    print("missing: lhs", r, lp, "rhs", grp >> 33, rp, "value", v, "run", cnt.get(v), "lhs is micro", (r, lp) in micset)
