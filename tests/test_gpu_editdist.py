"""GPU: batched global edit distance (raven_b200/csrc/editdist.cu) - the
edlibAlign(lhs, rhs, edlibDefaultAlignConfig()).editDistance of the identity filter
(RavenLib/src/construct.cc:176-199,393-416) - against the oracle's textbook
dynamic programme: exact, every band tier, both strands."""
import numpy as np
import pytest

from raven_b200 import seqio, synth

pytestmark = pytest.mark.gpu


def _sub(rs, r, b, n, strand):
    s = rs.codes(r)[b:b + n]
    if not strand:
        s = (3 - s[::-1]).astype(np.uint8)
    return np.frombuffer(b"ACGT", np.uint8)[s].tobytes()


def test_edit_distance_batch_matches_oracle(gpu_engine, oracle):
    rng = np.random.default_rng(5)
    base = rng.integers(0, 4, 9000, dtype=np.uint8)
    seqs = [base]
    # error rates that land in every tier (k = 64, 192, 448, 960) and beyond
    for err in (0.0, 0.002, 0.01, 0.03, 0.06, 0.12, 0.2, 0.35):
        seqs.append(np.asarray(synth.mutate(base, rng, err / 3, err / 3, err / 3), dtype=np.uint8))
    seqs.append((3 - base[::-1]).astype(np.uint8))          # reverse complement
    seqs.append(rng.integers(0, 4, 700, dtype=np.uint8))    # unrelated
    seqs.append(np.zeros(300, np.uint8))
    seqs.append(rng.integers(0, 4, 1, dtype=np.uint8))
    rs = seqio.pack_codes(seqs)
    gpu_engine.configure(15, 5)
    gpu_engine.upload(rs)
    pairs = []
    n0 = len(base)
    for j in range(1, 9):
        pairs.append((0, 0, n0, j, 0, int(rs.lens[j]), 1))
        pairs.append((0, 100, 3000, j, 90, 3100, 1))             # unaligned starts
        pairs.append((j, 17, 4000, 0, 33, 3900, 1))
    pairs.append((0, 0, n0, 9, 0, n0, 0))                         # rhs reverse complemented
    pairs.append((1, 500, 2500, 9, n0 - 3000, 2500, 0))
    pairs.append((0, 0, 63, 1, 0, 64, 1))
    pairs.append((0, 0, 64, 1, 0, 65, 1))
    pairs.append((0, 5, 129, 10, 0, 700, 1))                      # unrelated, unequal
    pairs.append((11, 0, 300, 11, 0, 300, 1))                     # homopolymer with itself
    pairs.append((11, 0, 300, 0, 0, 200, 0))
    pairs.append((12, 0, 1, 12, 0, 1, 0))                         # one base vs its complement
    pairs.append((12, 0, 0, 0, 0, 10, 1))                         # empty lhs
    pairs.append((0, 7, 10, 12, 0, 0, 1))                         # empty rhs
    for _ in range(40):
        a, b = rng.integers(0, 10, 2)
        la, lb = int(rs.lens[a]), int(rs.lens[b])
        na, nb = int(rng.integers(1, min(la, 2500))), int(rng.integers(1, min(lb, 2500)))
        pairs.append((int(a), int(rng.integers(0, la - na + 1)), na, int(b),
                      int(rng.integers(0, lb - nb + 1)), nb, int(rng.integers(0, 2))))
    P = np.array(pairs, dtype=np.int64)
    got = gpu_engine.edit_distance_batch(P[:, 0], P[:, 1], P[:, 2], P[:, 3], P[:, 4], P[:, 5],
                                         P[:, 6])
    want = np.array([oracle.edit_distance(_sub(rs, p[0], p[1], p[2], 1),
                                          _sub(rs, p[3], p[4], p[5], p[6])) for p in pairs],
                    dtype=np.int32)
    assert np.array_equal(got, want)
    # bounded: -1 exactly where the distance exceeds the limit
    for lim in (0, 10, 64, 65, 200, 1000, 5000):
        limit = np.full(len(pairs), lim, dtype=np.int32)
        g = gpu_engine.edit_distance_batch(P[:, 0], P[:, 1], P[:, 2], P[:, 3], P[:, 4], P[:, 5],
                                           P[:, 6], limit)
        assert np.array_equal(g, np.where(want <= lim, want, -1)), lim
    with pytest.raises(ValueError):
        gpu_engine.edit_distance_batch([0], [0], [10**6], [1], [0], [10], [1])
