"""Sequence containers in the biosoup wire format (host side, numpy).

The reference stores every read as ``biosoup::NucleicAcid`` (fields pinned by
RavenLib/include/raven/graph/graph.h:13-18): 2-bit codes A0 C1 G2 T3, 32 bases
per ``uint64`` word, base ``i`` at bits ``[(i<<1)&63, +1]`` of word ``i>>5``,
plus one mean Phred byte per 64 bases.  ``ReadSet`` is that layout flattened
over a whole read set -- exactly what ``rvn_reads_upload`` (include/raven_b200.h)
takes, so no repacking happens between the parser and the GPU.

Parsing mirrors RavenLib/src/io.cc:7-41 (format chosen by file suffix).
"""
from __future__ import annotations

import dataclasses
import gzip
import os

import numpy as np

_CODE = np.full(256, 255, dtype=np.uint8)
for _c, _v in zip(b"ACGTU", (0, 1, 2, 3, 3)):
    _CODE[_c] = _v
    _CODE[ord(chr(_c).lower())] = _v
# IUPAC ambiguity codes collapse like include/biosoup/nucleic_acid.hpp
for _c, _v in zip(b"RYKMSWBDHVN-", (0, 3, 2, 1, 1, 0, 1, 0, 3, 2, 0, 0)):
    _CODE[_c] = _v
    _CODE[ord(chr(_c).lower())] = _v


@dataclasses.dataclass
class ReadSet:
    words: np.ndarray      # uint64, concatenated deflated_data
    word_off: np.ndarray   # uint64, n + 1
    lens: np.ndarray       # uint32, n (inflated_len)
    block_quality: np.ndarray | None = None   # uint8, concatenated
    bq_off: np.ndarray | None = None          # uint64, n + 1
    names: list | None = None

    @property
    def n(self) -> int:
        return int(self.lens.shape[0])

    @property
    def bases(self) -> int:
        return int(self.lens.sum(dtype=np.uint64))

    def codes(self, i: int) -> np.ndarray:
        """2-bit codes of read i as a uint8 array."""
        w = self.words[int(self.word_off[i]):int(self.word_off[i + 1])]
        shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))
        c = ((w[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8)
        return c.reshape(-1)[: int(self.lens[i])]

    def ascii(self, i: int) -> bytes:
        return np.frombuffer(b"ACGT", dtype=np.uint8)[self.codes(i)].tobytes()

    def subset(self, idx) -> "ReadSet":
        idx = list(idx)
        seqs = [self.codes(i) for i in idx]
        rs = pack_codes(seqs)
        if self.block_quality is not None:
            bq = [self.block_quality[int(self.bq_off[i]):int(self.bq_off[i + 1])]
                  for i in idx]
            rs.block_quality = (np.concatenate(bq) if bq else
                                np.zeros(0, np.uint8))
            rs.bq_off = np.concatenate(
                [[0], np.cumsum([len(b) for b in bq])]).astype(np.uint64)
        if self.names is not None:
            rs.names = [self.names[i] for i in idx]
        return rs

    def save(self, path: str) -> None:
        d = dict(words=self.words, word_off=self.word_off, lens=self.lens)
        if self.block_quality is not None:
            d.update(block_quality=self.block_quality, bq_off=self.bq_off)
        if self.names is not None:
            d.update(names=np.array(self.names))
        np.savez_compressed(path, **d)

    @staticmethod
    def load(path: str) -> "ReadSet":
        z = np.load(path, allow_pickle=False)
        rs = ReadSet(z["words"], z["word_off"], z["lens"])
        if "block_quality" in z:
            rs.block_quality, rs.bq_off = z["block_quality"], z["bq_off"]
        if "names" in z:
            rs.names = [str(x) for x in z["names"]]
        return rs


def pack_codes(seqs) -> ReadSet:
    """Pack a list of uint8 code arrays (values 0..3)."""
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    nwords = (lens.astype(np.uint64) + np.uint64(31)) >> np.uint64(5)
    word_off = np.concatenate([[0], np.cumsum(nwords)]).astype(np.uint64)
    words = np.zeros(int(word_off[-1]), dtype=np.uint64)
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))
    for i, s in enumerate(seqs):
        nw = int(nwords[i])
        if nw == 0:
            continue
        buf = np.zeros(nw * 32, dtype=np.uint64)
        buf[: len(s)] = np.asarray(s, dtype=np.uint64)
        words[int(word_off[i]):int(word_off[i]) + nw] = np.bitwise_or.reduce(
            buf.reshape(nw, 32) << shifts[None, :], axis=1)
    return ReadSet(words, word_off, lens)


def pack_ascii(seqs, quals=None, names=None) -> ReadSet:
    """Pack ASCII reads (bytes); optional FASTQ quality strings."""
    codes = []
    for s in seqs:
        c = _CODE[np.frombuffer(s, dtype=np.uint8)]
        if (c == 255).any():
            raise ValueError("not a nucleotide")
        codes.append(c)
    rs = pack_codes(codes)
    if quals is not None and any(q is not None and len(q) for q in quals):
        bqs = []
        for q in quals:
            q = np.frombuffer(q, dtype=np.uint8).astype(np.uint32) - 33
            nb = (len(q) + 63) // 64
            if nb == 0:
                bqs.append(np.zeros(0, np.uint8))
                continue
            sums = np.add.reduceat(q, np.arange(0, len(q), 64))
            cnt = np.minimum(64, len(q) - np.arange(nb) * 64)
            bqs.append((sums // cnt).astype(np.uint8))
        rs.block_quality = np.concatenate(bqs) if bqs else np.zeros(0, np.uint8)
        rs.bq_off = np.concatenate(
            [[0], np.cumsum([len(b) for b in bqs])]).astype(np.uint64)
    rs.names = list(names) if names is not None else None
    return rs


_FASTA = (".fasta", ".fa", ".fasta.gz", ".fa.gz")
_FASTQ = (".fastq", ".fq", ".fastq.gz", ".fq.gz")


def parse(path: str) -> ReadSet:
    """FASTA/FASTQ(.gz) -> ReadSet; suffix rules of RavenLib/src/io.cc:7-41."""
    low = path.lower()
    if not low.endswith(_FASTA + _FASTQ):
        raise ValueError(
            "[raven::CreateParser] error: file " + path +
            " has unsupported format extension (valid extensions: .fasta, "
            ".fasta.gz, .fa, .fa.gz, .fastq, .fastq.gz, .fq, .fq.gz)")
    if not os.path.exists(path):
        raise ValueError("[bioparser::Parser::Create] error: unable to open file " + path)
    opener = gzip.open if low.endswith(".gz") else open
    with opener(path, "rb") as f:
        data = f.read()
    names, seqs, quals = [], [], []
    lines = data.split(b"\n")
    if low.endswith(_FASTQ):
        i = 0
        while i + 3 < len(lines) + 0 and lines[i].startswith(b"@"):
            names.append(lines[i][1:].split()[0].decode())
            seqs.append(lines[i + 1].strip())
            quals.append(lines[i + 3].strip())
            i += 4
        return pack_ascii(seqs, quals, names)
    cur = []
    for ln in lines:
        if ln.startswith(b">"):
            if names:
                seqs.append(b"".join(cur))
            names.append(ln[1:].split()[0].decode())
            cur = []
        else:
            cur.append(ln.strip())
    if names:
        seqs.append(b"".join(cur))
    return pack_ascii(seqs, None, names)
