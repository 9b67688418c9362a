"""ctypes wrapper of bench/synth_reads.cpp (seeded synthetic reads at bench
scale, written straight into the biosoup wire format)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsynth.so")


def build():
    src = os.path.join(HERE, "synth_reads.cpp")
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return
    subprocess.run(["/usr/bin/g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread",
                    "-fvisibility=hidden", "-o", LIB, src], check=True)


def make_reads(seed, genome_len, n_reads, mean_len, sub=0.03, ins=0.03, dele=0.04,
               min_len=2000, threads=0):
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from raven_b200 import seqio
    build()
    lib = C.CDLL(LIB)
    lib.synth_reads.restype = C.c_void_p
    lib.synth_reads.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_double,
                                C.c_double, C.c_double, C.c_uint32, C.c_uint32]
    for f, t in (("synth_n_words", C.c_uint64), ("synth_words", C.c_void_p),
                 ("synth_word_off", C.c_void_p), ("synth_lens", C.c_void_p)):
        getattr(lib, f).restype = t
        getattr(lib, f).argtypes = [C.c_void_p]
    lib.synth_free.argtypes = [C.c_void_p]
    h = lib.synth_reads(seed, genome_len, n_reads, mean_len, sub, ins, dele, min_len, threads)
    nw = lib.synth_n_words(h)

    def view(ptr, n, dt):
        if n == 0:
            return np.zeros(0, dt)
        buf = (C.c_uint8 * (n * np.dtype(dt).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt).copy()

    rs = seqio.ReadSet(view(lib.synth_words(h), nw, np.uint64),
                       view(lib.synth_word_off(h), n_reads + 1, np.uint64),
                       view(lib.synth_lens(h), n_reads, np.uint32))
    lib.synth_free(h)
    return rs


def make_contigs(seed, genome_len, contig_len=1_000_000, sub=0.004, ins=0.003, dele=0.003,
                 threads=0):
    """Draft contigs of the genome `make_reads(seed, genome_len, ...)` samples from."""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from raven_b200 import seqio
    build()
    lib = C.CDLL(LIB)
    lib.synth_contigs.restype = C.c_void_p
    lib.synth_contigs.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_double,
                                  C.c_double, C.c_uint32]
    lib.synth_n_reads.restype = C.c_uint32
    lib.synth_n_reads.argtypes = [C.c_void_p]
    for f, t in (("synth_n_words", C.c_uint64), ("synth_words", C.c_void_p),
                 ("synth_word_off", C.c_void_p), ("synth_lens", C.c_void_p)):
        getattr(lib, f).restype = t
        getattr(lib, f).argtypes = [C.c_void_p]
    lib.synth_free.argtypes = [C.c_void_p]
    h = lib.synth_contigs(seed, genome_len, contig_len, sub, ins, dele, threads)
    n, nw = lib.synth_n_reads(h), lib.synth_n_words(h)

    def view(ptr, cnt, dt):
        buf = (C.c_uint8 * (cnt * np.dtype(dt).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt).copy()

    rs = seqio.ReadSet(view(lib.synth_words(h), nw, np.uint64),
                       view(lib.synth_word_off(h), n + 1, np.uint64),
                       view(lib.synth_lens(h), n, np.uint32))
    lib.synth_free(h)
    return rs
