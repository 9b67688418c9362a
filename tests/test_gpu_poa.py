"""GPU parity: racon window consensus (POA) through the C ABI vs the CPU oracle
(oracle/spoa_graph.cpp + oracle/racon_window.cpp) - consensus letters,
coverages and status bit-exact."""
import numpy as np
import pytest

from raven_b200 import synth

pytestmark = pytest.mark.gpu


def check(gpu_engine, oracle, w, **kw):
    got = gpu_engine.poa_batch(w, **kw)
    want = oracle.poa_batch(w, threads=8, **kw)
    assert np.array_equal(got["cons_off"], want["cons_off"])
    assert np.array_equal(got["consensus"], want["consensus"])
    assert np.array_equal(got["status"], want["status"])
    assert np.array_equal(got["coverage"], want["coverage"])
    assert got["cells"] == int(want["cells"].sum())
    return got


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_poa_ont_windows(gpu_engine, oracle, seed):
    w = synth.make_windows(n_windows=24, backbone_len=500, layers=30, seed=seed)
    got = check(gpu_engine, oracle, w)
    assert (got["status"] & 1).all()
    # the consensus is much closer to the hidden truth than the draft backbone
    err_c = err_b = 0
    for i in range(24):
        c = got["consensus"][int(got["cons_off"][i]):int(got["cons_off"][i + 1])].tobytes()
        s0 = int(w["win_first"][i])
        b = w["bases"][int(w["seq_off"][s0]):int(w["seq_off"][s0 + 1])].tobytes()
        err_c += oracle.edit_distance(c, w["truths"][i])
        err_b += oracle.edit_distance(b, w["truths"][i])
    assert err_c * 10 < err_b


def test_poa_variants(gpu_engine, oracle):
    # no qualities (weight 1), all layers partial, other scores, no trimming, NGS type
    w = synth.make_windows(n_windows=12, backbone_len=300, layers=12, seed=5,
                           with_quality=False, partial=1.0)
    check(gpu_engine, oracle, w)
    check(gpu_engine, oracle, w, m=5, n=-4, g=-8)
    check(gpu_engine, oracle, w, trim=False)
    check(gpu_engine, oracle, w, tgs=False)
    # HiFi-like: few errors, deep
    w = synth.make_windows(n_windows=6, backbone_len=500, layers=60, seed=6, sub=0.002,
                           ins=0.0015, dele=0.0015)
    check(gpu_engine, oracle, w)
    # ragged: windows with 0, 1, 2, ... layers (fewer than 3 sequences = backbone)
    w = synth.make_windows(n_windows=16, backbone_len=200, layers=6, seed=7, min_layers=0)
    got = check(gpu_engine, oracle, w)
    nl = np.diff(w["win_first"].astype(np.int64))
    assert ((got["status"] & 1) == (nl >= 3)).all()


def _concat_windows(a, b):
    """Two flat window batches as one."""
    ns, nb = int(a["win_first"][-1]), int(a["seq_off"][-1])
    return dict(
        win_first=np.concatenate([a["win_first"], b["win_first"][1:] + ns]).astype(np.uint32),
        seq_off=np.concatenate([a["seq_off"], b["seq_off"][1:] + np.uint64(nb)]).astype(np.uint64),
        bases=np.concatenate([a["bases"], b["bases"]]),
        quals=None if a["quals"] is None else np.concatenate([a["quals"], b["quals"]]),
        seq_begin=np.concatenate([a["seq_begin"], b["seq_begin"]]),
        seq_end=np.concatenate([a["seq_end"], b["seq_end"]]))


def test_poa_long_layers_and_mixed_batches(gpu_engine, oracle):
    """Layers beyond the 575 bases of the register-row kernel take the generic
    kernel - chosen per WINDOW, so a mixed batch runs both (real ONT windows have
    such layers: 500 target bases, 3 % insertions and more in the reads)."""
    long_w = synth.make_windows(n_windows=5, backbone_len=760, layers=14, seed=11)
    nl = np.diff(long_w["seq_off"].astype(np.int64))
    assert nl.max() > 600
    check(gpu_engine, oracle, long_w)
    short_w = synth.make_windows(n_windows=7, backbone_len=500, layers=20, seed=12)
    both = _concat_windows(short_w, long_w)
    got = check(gpu_engine, oracle, both)
    alone = gpu_engine.poa_batch(short_w)
    n = int(alone["cons_off"][-1])
    assert np.array_equal(got["consensus"][:n], alone["consensus"])
    # one window with a single long, insertion-rich layer among short ones
    w = synth.make_windows(n_windows=3, backbone_len=520, layers=10, seed=13, ins=0.12, dele=0.01)
    assert np.diff(w["seq_off"].astype(np.int64)).max() > 576
    check(gpu_engine, oracle, _concat_windows(w, short_w))


def test_poa_rejects_bad_input(gpu_engine):
    w = synth.make_windows(n_windows=2, backbone_len=100, layers=4, seed=8)
    with pytest.raises(ValueError):
        gpu_engine.poa_batch(w, g=0)          # racon: gap must be negative
    bad = dict(w)
    bad["seq_end"] = w["seq_end"].copy()
    bad["seq_end"][1] = 5000                  # beyond the backbone
    with pytest.raises(ValueError):
        gpu_engine.poa_batch(bad)
    empty = dict(win_first=np.zeros(1, np.uint32), seq_off=np.zeros(1, np.uint64),
                 bases=np.zeros(0, np.uint8), quals=None, seq_begin=np.zeros(0, np.uint32),
                 seq_end=np.zeros(0, np.uint32))
    got = gpu_engine.poa_batch(empty)
    assert got["cons_off"].tolist() == [0]


def test_polisher_facade_equals_oracle_polisher(oracle):
    """racon::Polisher::Polish (polish.cc:43-51) through the product facade (GPU map +
    host edlib path + GPU POA) == the oracle polisher (CPU): same polished targets,
    same name tags, on ONT-like reads over draft contigs, with and without qualities."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from bench import synth as bsynth
    from raven_b200 import polish
    reads = bsynth.make_reads(77, 60_000, 240, 5000, min_len=1500)
    draft = bsynth.make_contigs(77, 60_000, contig_len=25_000)
    assert draft.n == 3
    for with_q in (False, True):
        if with_q:
            rng = np.random.default_rng(1)
            nb = (reads.lens.astype(np.int64) + 63) // 64
            reads.block_quality = rng.integers(6, 20, int(nb.sum())).astype(np.uint8)
            reads.bq_off = np.concatenate([[0], np.cumsum(nb)]).astype(np.uint64)
        q = 10.0 if with_q else 0.0
        got, st = polish.polish(draft, reads, q=q, threads=4)
        names, seqs, ost = oracle.polish(draft, reads, q=q, threads=8)
        assert got.names == names
        assert [got.ascii(i) for i in range(got.n)] == seqs
        assert st["windows"] == int(ost[0]) and st["polished_windows"] == int(ost[1])
        assert st["polished_windows"] > 100
    # the alignment batches of the facade (here forced small) change nothing
    os.environ["RVN_POLISH_BATCH_BASES"] = "150000"
    try:
        again, _ = polish.polish(draft, reads, q=10.0, threads=4)
    finally:
        del os.environ["RVN_POLISH_BATCH_BASES"]
    assert [again.ascii(i) for i in range(again.n)] == seqs
