"""The GPU truncation step replays libstdc++'s std::sort move for move
(raven_b200/csrc/introsort.cuh). Here the same header, compiled for the host,
is compared with the real std::sort (oracle) on element ORDER, ties included."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
U64P = C.POINTER(C.c_uint64)


@pytest.fixture(scope="module")
def ours():
    so = os.path.join(HERE, "_introsort_host.so")
    src = os.path.join(HERE, "introsort_host.cpp")
    hdr = os.path.join(HERE, "..", "raven_b200", "csrc", "introsort.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src),
                                                            os.path.getmtime(hdr)):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                        src, "-o", so], check=True)
    lib = C.CDLL(so)
    lib.rvn_test_stdsort.argtypes = [U64P, C.c_uint64]
    return lib


def run_both(ours, oracle, keys):
    keys = np.asarray(keys, dtype=np.uint64)
    data = (keys << np.uint64(32)) | np.arange(keys.size, dtype=np.uint64)
    a, b = data.copy(), data.copy()
    if a.size:
        ours.rvn_test_stdsort(a.ctypes.data_as(U64P), a.size)
        oracle.lib.orc_std_sort_hi32_desc.argtypes = [U64P, C.c_uint64]
        oracle.lib.orc_std_sort_hi32_desc(b.ctypes.data_as(U64P), b.size)
    assert np.array_equal(a, b)
    assert (np.diff((a >> np.uint64(32)).astype(np.int64)) <= 0).all()


def median_of_3_killer(n):
    """Musser's adversary for median-of-3 quicksort (drives the depth limit)."""
    n -= n % 2
    k = n // 2
    a = [0] * n
    for i in range(k):
        if i % 2 == 0:
            a[i] = i + 1
        else:
            a[i] = k + i + (1 - (k % 2 == 0) * 0)
        a[k + i] = 2 * (i + 1)
    return a


def test_against_std_sort(ours, oracle):
    rng = np.random.default_rng(5)
    for n in list(range(0, 40)) + [63, 64, 65, 100, 257, 1000, 4097, 20000]:
        run_both(ours, oracle, rng.integers(0, 2**31, n))           # distinct-ish
        run_both(ours, oracle, rng.integers(0, 7, n))               # tie-heavy
        run_both(ours, oracle, rng.integers(3000, 3040, n))         # overlap lengths
        run_both(ours, oracle, np.zeros(n))                         # all equal
        run_both(ours, oracle, np.arange(n))                        # ascending
        run_both(ours, oracle, np.arange(n)[::-1])                  # descending
        run_both(ours, oracle, np.minimum(np.arange(n), np.arange(n)[::-1]))  # organ pipe
    for n in (50, 200, 1000, 5000, 30000):
        run_both(ours, oracle, median_of_3_killer(n))
        run_both(ours, oracle, 2**20 - np.array(median_of_3_killer(n)))
