"""Regenerate tests/golden/lambda_golden.npz + lambda_golden.json.

Golden intermediates of the overlap path on the reference's own fixture
(RavenTest/data/ERA476754.fastq.gz re-encoded as lambda_reads.npz), produced by
the CPU oracle and, for the in-tree half (batch schedule, gather, AddLayers,
truncation), cross-checked here against the reference's own sources compiled
in place (oracle/_ref).  Upstream holds no per-stage golden vectors for this
path; these pin OUR oracle so that it cannot drift silently.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
from raven_b200 import seqio  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


rs = seqio.ReadSet.load(os.path.join(HERE, "lambda_reads.npz"))
O = oracle_lib.Oracle()
reads = O.reads(rs)
meta = {"k": 15, "w": 5, "freq": 0.001, "reads": rs.n, "bases": rs.bases}
arrays = {}

eng = O.engine(15, 5, threads=4)
for mh in (False, True):
    sk = O.sketch(eng, reads, 0, rs.n, mh)
    tag = "micro" if mh else "full"
    meta[f"sketch_{tag}_n"] = int(sk["value"].size)
    meta[f"sketch_{tag}_value_sha256"] = sha(sk["value"])
    meta[f"sketch_{tag}_origin_sha256"] = sha(sk["origin"])
    meta[f"sketch_{tag}_offsets_sha256"] = sha(sk["offsets"])

O.minimize(eng, reads, 0, rs.n, False)
keys = O.keys(eng)
meta["index_keys"] = int(keys["totals"][0])
meta["index_records"] = int(keys["totals"][1])
meta["index_singletons"] = int((keys["counts"] == 1).sum())
meta["index_max_count"] = int(keys["counts"].max())
meta["occurrence"] = int(O.filter(eng, 0.001))
for mh in (True, False):
    m = O.map(eng, reads, 0, rs.n, True, True, mh, True)
    tag = "micro" if mh else "full"
    meta[f"map_{tag}_hits"] = int(m["match_group"].size)
    meta[f"map_{tag}_overlaps"] = int(m["overlaps"].shape[0])
    meta[f"map_{tag}_overlaps_sha256"] = sha(m["overlaps"])
    meta[f"map_{tag}_filtered"] = int(m["filtered"].size)
    meta[f"map_{tag}_filtered_sha256"] = sha(m["filtered"])
    arrays[f"map_{tag}_overlaps"] = m["overlaps"]
    arrays[f"map_{tag}_ovl_off"] = m["ovl_off"]

R = oracle_lib.Reference() if oracle_lib.Reference.available() else None
for mh in (False, True):
    st = O.stage1(O.engine(15, 5, threads=4), reads, 0.001, 32, mh)
    tag = "minhash" if mh else "plain"
    if R is not None:
        ref = R.stage1(R.reads(rs), 15, 5, 0.001, 32, mh, 4)
        for k in ("overlaps", "ovl_off", "pile", "pile_off", "occurrences"):
            assert np.array_equal(st[k], ref[k]), (tag, k)
        meta[f"stage1_{tag}_checked_against_compiled_reference"] = True
    for k in ("overlaps", "ovl_off", "pile", "pile_off"):
        arrays[f"stage1_{tag}_{k}"] = st[k]
        meta[f"stage1_{tag}_{k}_sha256"] = sha(st[k])
    meta[f"stage1_{tag}_num_mapped"] = int(st["num_mapped"][0])
    meta[f"stage1_{tag}_occurrence"] = int(st["occurrences"][0])

np.savez_compressed(os.path.join(HERE, "lambda_golden.npz"), **arrays)
with open(os.path.join(HERE, "lambda_golden.json"), "w") as f:
    json.dump(meta, f, indent=1, sort_keys=True)
print(json.dumps(meta, indent=1, sort_keys=True))
