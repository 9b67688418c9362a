"""Seeded synthetic long reads (host side, numpy): uniform random genome,
reads drawn from both strands with substitution / insertion / deletion errors
(ONT 10 % = 3/3/4, HiFi 0.5 % = 0.2/0.15/0.15)."""
from __future__ import annotations

import numpy as np

from . import seqio


def mutate(codes, rng, sub, ins, dele):
    """Apply i.i.d. errors to a uint8 code array."""
    n = len(codes)
    if n == 0 or sub + ins + dele == 0:
        return codes
    u = rng.random(n)
    keep = u >= dele
    is_sub = (u >= dele) & (u < dele + sub)
    out = codes.copy()
    out[is_sub] = (out[is_sub] + rng.integers(1, 4, is_sub.sum(), dtype=np.uint8)) & 3
    is_ins = rng.random(n) < ins
    reps = keep.astype(np.int64) + is_ins
    res = np.repeat(out, reps)
    # the inserted copy gets a random base
    pos = np.cumsum(reps) - 1
    ins_pos = pos[is_ins & (reps > 0)]
    res[ins_pos] = rng.integers(0, 4, len(ins_pos), dtype=np.uint8)
    return res


def make_reads(genome_len=100_000, n_reads=200, mean_len=5000, seed=1, sub=0.03,
               ins=0.03, dele=0.04, min_len=500, genome=None):
    rng = np.random.default_rng(seed)
    if genome is None:
        genome = rng.integers(0, 4, genome_len, dtype=np.uint8)
    genome_len = len(genome)
    seqs = []
    for _ in range(n_reads):
        ln = int(np.clip(rng.normal(mean_len, 0.2 * mean_len), min_len, 5 * mean_len))
        ln = min(ln, genome_len)
        st = int(rng.integers(0, genome_len - ln + 1))
        s = genome[st:st + ln]
        if rng.random() < 0.5:
            s = (3 - s[::-1]).astype(np.uint8)
        seqs.append(mutate(s, rng, sub, ins, dele))
    return seqio.pack_codes(seqs)


def random_reads(lengths, seed=1):
    """Unrelated random reads of given lengths (edge-case tests)."""
    rng = np.random.default_rng(seed)
    return seqio.pack_codes([rng.integers(0, 4, n, dtype=np.uint8) for n in lengths])
