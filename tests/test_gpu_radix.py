"""GPU: the engine's own stable LSD radix sort (raven_b200/csrc/radix.cu) - the
sort behind the minimizer index (ram::MinimizerEngine::Minimize, call site
RavenLib/src/construct.cc:42-43) - against numpy's stable sort, bit exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref(keys, vals, begin_bit, end_bit, descending):
    width = end_bit - begin_bit
    if width <= 0:
        return keys, vals
    mask = np.uint64((1 << width) - 1)
    d = (keys.astype(np.uint64) >> np.uint64(begin_bit)) & mask
    if descending:
        d = mask - d
    order = np.argsort(d, kind="stable")
    return keys[order], None if vals is None else vals[order]


CASES = [
    # (key dtype, val dtype, n, key distribution bits, begin, end, descending)
    (np.uint32, np.uint64, 0, 30, 0, 30, False),
    (np.uint32, np.uint64, 1, 30, 0, 30, False),
    (np.uint32, np.uint64, 31, 30, 0, 30, False),
    (np.uint32, np.uint64, 8191, 30, 0, 30, False),
    (np.uint32, np.uint64, 8192, 30, 0, 30, False),
    (np.uint32, np.uint64, 8193, 30, 0, 30, False),
    (np.uint32, np.uint64, 100_003, 30, 0, 30, False),
    (np.uint32, np.uint64, 3_000_017, 30, 0, 30, False),
    (np.uint32, np.uint64, 3_000_017, 6, 0, 30, False),      # few distinct keys: long runs
    (np.uint32, np.uint64, 1_000_000, 30, 0, 7, False),      # one short pass
    (np.uint32, np.uint64, 1_000_000, 32, 0, 32, False),
    (np.uint64, np.uint64, 2_000_003, 38, 0, 38, False),     # k = 19
    (np.uint64, np.uint64, 4097, 62, 0, 62, False),          # k = 31
    (np.uint64, np.uint32, 1_500_001, 30, 14, 30, False),    # probe order: top 16 bits
    (np.uint64, np.uint32, 700_001, 40, 0, 37, False),       # overlap emission keys
    (np.uint32, np.uint32, 900_001, 12, 0, 12, True),        # pair sizes, descending
    (np.uint32, np.uint32, 50_001, 17, 0, 17, False),        # rhs ids of the gather
    (np.uint32, None, 2_500_000, 30, 0, 30, False),          # keys only
]


@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_radix_sort_matches_stable_numpy(gpu_engine, case):
    kt, vt, n, kbits, b, e, desc = case
    rng = np.random.default_rng(1000 + n % 97 + kbits)
    keys = rng.integers(0, 1 << kbits, n, dtype=np.uint64).astype(kt)
    if n > 10:
        keys[: n // 3] = keys[n // 2]            # a heavy run that spans tiles
    vals = None if vt is None else rng.integers(0, 1 << 31, n, dtype=np.uint64).astype(vt)
    gk, gv = gpu_engine.debug_sort_pairs(keys, vals, b, e, desc)
    wk, wv = _ref(keys, vals, b, e, desc)
    assert np.array_equal(gk, wk)
    if vt is not None:
        assert np.array_equal(gv, wv)


def test_radix_sort_sorted_and_reversed_input(gpu_engine):
    n = 1_200_000
    keys = (np.arange(n, dtype=np.uint64) * 7919 % (1 << 30)).astype(np.uint32)
    vals = np.arange(n, dtype=np.uint64)
    for k in (np.sort(keys), np.sort(keys)[::-1].copy(), np.zeros(n, np.uint32)):
        gk, gv = gpu_engine.debug_sort_pairs(k, vals, 0, 30)
        wk, wv = _ref(k, vals, 0, 30, False)
        assert np.array_equal(gk, wk) and np.array_equal(gv, wv)
