"""GPU: the reference's own construct.cc (compiled in place, tests/cpp/Makefile)
running on the B200 engine through the ram::MinimizerEngine facade, and our
batched FindOverlapsAndCreatePiles replacement with the reference signature -
both against the CPU oracle, bit exact."""
import os
import struct
import subprocess

import numpy as np
import pytest

from raven_b200 import synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "cpp", "_build", "dropin_test")


def write_vec(f, a):
    f.write(struct.pack("<Q", a.size))
    f.write(np.ascontiguousarray(a).tobytes())


def read_vec(f, dt):
    (n,) = struct.unpack("<Q", f.read(8))
    return np.frombuffer(f.read(n * np.dtype(dt).itemsize), dtype=dt).copy()


def run_dropin(tmp_path, rs, k, w, freq, kmax, minhash):
    if not os.path.exists(BIN):
        pytest.skip("tests/cpp/_build/dropin_test not built (needs /root/reference at build time)")
    inp, out = str(tmp_path / "reads.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        write_vec(f, rs.words.astype(np.uint64))
        write_vec(f, rs.word_off.astype(np.uint64))
        write_vec(f, rs.lens.astype(np.uint32))
    subprocess.run([BIN, inp, out, str(k), str(w), str(freq), str(kmax), str(int(minhash))],
                   check=True, stderr=subprocess.DEVNULL, timeout=600)
    res = []
    with open(out, "rb") as f:
        for _ in range(2):
            res.append(dict(overlaps=read_vec(f, np.uint32).reshape(-1, 8),
                            ovl_off=read_vec(f, np.uint64), pile=read_vec(f, np.uint16),
                            pile_off=read_vec(f, np.uint64)))
    return res


@pytest.mark.parametrize("minhash", [False, True])
def test_reference_construct_on_b200_facade(tmp_path, oracle, lambda_reads, minhash):
    a, b = run_dropin(tmp_path, lambda_reads, 15, 5, 0.001, 32, minhash)
    want = oracle.stage1(oracle.engine(15, 5, threads=4), oracle.reads(lambda_reads),
                         0.001, 32, minhash)
    for got, name in ((a, "reference construct.cc over the facade"),
                      (b, "batched replacement")):
        for key in ("overlaps", "ovl_off", "pile", "pile_off"):
            assert np.array_equal(got[key], want[key]), (name, key)


def test_dropin_synthetic_small_kmax(tmp_path, oracle):
    rs = synth.make_reads(40_000, 150, 3000, seed=13)
    a, b = run_dropin(tmp_path, rs, 15, 5, 0.001, 6, False)
    want = oracle.stage1(oracle.engine(15, 5, threads=4), oracle.reads(rs), 0.001, 6, False)
    for got in (a, b):
        for key in ("overlaps", "ovl_off", "pile", "pile_off"):
            assert np.array_equal(got[key], want[key]), key
