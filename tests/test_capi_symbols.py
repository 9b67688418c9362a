"""CPU: the C-ABI library loads and exports every symbol include/raven_b200.h
declares; with no GPU every compute entry point must refuse loudly."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build_cuda()
    from raven_b200 import _lib
    return _lib.load()


def test_header_and_bindings_agree(lib):
    from raven_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "raven_b200.h")).read()
    declared = set(re.findall(r"\b(rvn_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"rvn_ctx", "rvn_overlap", "rvn_stats"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name


def test_version(lib):
    assert lib.rvn_version() >= 100


def test_no_silent_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.rvn_ctx_create(0, None, C.byref(h)) == -2  # RVN_ERR_CUDA
    assert not h.value
    from raven_b200 import engine
    with pytest.raises(RuntimeError):
        engine.Engine(device=0)


def test_product_does_not_touch_the_oracle():
    """Nothing under raven_b200/ may import, link or load oracle/."""
    pkg = os.path.join(ROOT, "raven_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "liboracle" not in txt, f
                assert not re.search(r'#include\s+"[^"]*oracle', txt), f
