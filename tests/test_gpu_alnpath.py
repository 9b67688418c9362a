"""GPU: batched alignment paths + window breaking points (raven_b200/csrc/alnpath.cu,
rvn_align_breaking_points) - the per-read edlibAlign(query, target, NW, PATH) of
racon::Polisher::Polish (RavenLib/src/polish.cc:43-51) and the cut of each path at
the target's windows - against the product's host edlib (raven_b200/host/edlib.cc,
itself checked against the oracle's dynamic programme in tests/test_oracle.py):
the same distances and exactly the same first / last aligned pairs per window,
for leaves, one split and several Hirschberg levels, both strands."""
import numpy as np
import pytest

from raven_b200 import seqio, synth

pytestmark = pytest.mark.gpu

LETTERS = np.frombuffer(b"ACGT", np.uint8)


def _sub(rs, r, b, n, strand):
    s = rs.codes(r)[b:b + n]
    if not strand:
        s = (3 - s[::-1]).astype(np.uint8)
    return LETTERS[s].tobytes()


def _expected(path, t_begin, t_len, window):
    """Walk an edlib path (0 match, 1 insert = query only, 2 delete = target only,
    3 mismatch): per touched window first (t, q) and one past the last (t, q)."""
    if t_len == 0:
        return np.zeros((0, 4), np.uint32)
    w0 = t_begin // window
    out = np.full(((t_begin + t_len - 1) // window - w0 + 1, 4), 0xFFFFFFFF, np.uint32)
    ops = np.frombuffer(path, np.uint8)
    diag = (ops == 0) | (ops == 3)
    tq = np.cumsum((ops != 1).astype(np.int64)) - 1 + t_begin   # target position of each op
    qq = np.cumsum((ops != 2).astype(np.int64)) - 1             # query position of each op
    tp, qp = tq[diag], qq[diag]
    if tp.size == 0:
        return out
    win = tp // window - w0
    first = np.concatenate([[True], win[1:] != win[:-1]])
    last = np.concatenate([win[1:] != win[:-1], [True]])
    out[win[first], 0] = tp[first]
    out[win[first], 1] = qp[first]
    out[win[last], 2] = tp[last] + 1
    out[win[last], 3] = qp[last] + 1
    return out


def test_breaking_points_equal_host_edlib(gpu_engine):
    import test_oracle
    align = test_oracle._host_edlib()
    rng = np.random.default_rng(23)
    target = rng.integers(0, 4, 40_000, dtype=np.uint8)
    seqs = [target]
    pairs = []   # (q_read, q_begin, q_len, strand, t_begin, t_len)

    def add_read(t0, n, err, strand, trim=(0, 0)):
        piece = target[t0:t0 + n]
        read = np.asarray(synth.mutate(piece, rng, err / 3, err / 3, err / 3), dtype=np.uint8)
        if not strand:
            read = (3 - read[::-1]).astype(np.uint8)
        seqs.append(read)
        b, e = trim
        pairs.append((len(seqs) - 1, b, len(read) - b - e, strand, t0, n))

    # leaves at the root (traceback data below 1 MiB), one split, several levels
    for n, err in ((40, 0.1), (63, 0.0), (64, 0.2), (300, 0.1), (700, 0.15), (1500, 0.1),
                   (2500, 0.12), (4000, 0.05), (6000, 0.1), (9000, 0.12), (12000, 0.1),
                   (20000, 0.11), (15000, 0.01), (8000, 0.3)):
        for strand in (1, 0):
            t0 = int(rng.integers(0, len(target) - n))
            add_read(t0, n, err, strand, trim=(int(rng.integers(0, 30)), int(rng.integers(0, 30))))
    # window-aligned ends, a target range inside one window, unequal lengths
    add_read(1000, 1000, 0.1, 1)
    add_read(1499, 2, 0.0, 1)
    add_read(1510, 200, 0.1, 0)
    add_read(0, 5000, 0.1, 1, trim=(400, 0))
    seqs.append(rng.integers(0, 4, 3000, dtype=np.uint8))   # unrelated read
    pairs.append((len(seqs) - 1, 0, 3000, 1, 7000, 2500))
    pairs.append((1, 0, 0, 1, 100, 50))                      # empty query
    pairs.append((1, 0, 10, 1, 100, 0))                      # empty target
    rs = seqio.pack_codes(seqs)
    gpu_engine.configure(15, 5)
    gpu_engine.upload(rs)
    P = np.array(pairs, dtype=np.int64)
    host = [align(_sub(rs, qr, qb, ql, st), _sub(rs, 0, tb, tl, 1), 2)
            for (qr, qb, ql, st, tb, tl) in pairs]
    for window in (500, 137):
        dist, off, bp = gpu_engine.align_breaking_points(
            P[:, 0], P[:, 1], P[:, 2], P[:, 3], np.zeros(len(P)), P[:, 4], P[:, 5], window)
        for i, (qr, qb, ql, st, tb, tl) in enumerate(pairs):
            d, path = host[i]
            assert dist[i] == d, (i, pairs[i])
            want = _expected(path, tb, tl, window)
            got = bp[int(off[i]):int(off[i + 1])]
            assert np.array_equal(got, want), (i, pairs[i], window)


def test_breaking_points_reject_bad_input(gpu_engine):
    rng = np.random.default_rng(3)
    rs = seqio.pack_codes([rng.integers(0, 4, 500, dtype=np.uint8) for _ in range(2)])
    gpu_engine.configure(15, 5)
    gpu_engine.upload(rs)
    with pytest.raises(ValueError):   # substring beyond the read
        gpu_engine.align_breaking_points([1], [0], [501], [1], [0], [0], [500])
    with pytest.raises(ValueError):   # unknown read
        gpu_engine.align_breaking_points([2], [0], [10], [1], [0], [0], [10])
    d, off, bp = gpu_engine.align_breaking_points([], [], [], [], [], [], [])
    assert d.size == 0 and off.tolist() == [0] and bp.shape == (0, 4)
