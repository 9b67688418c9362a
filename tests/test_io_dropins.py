"""CPU: the drop-in I/O dependencies of the reference's CLI - bioparser on zlib
(include/bioparser; RavenLib/src/io.cc:7-41, RavenExe/src/main.cc:258-272) and the
cereal archives (include/cereal; RavenLib/src/binary.cc:73-93)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from raven_b200 import seqio

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "cpp", "_build", "io_test")


@pytest.fixture(scope="module")
def io_test():
    subprocess.run(["make", "-C", os.path.join(HERE, "cpp"), "host"], check=True,
                   stdout=subprocess.DEVNULL)
    return BIN


def _records(out):
    recs = []
    for line in out.decode().splitlines():
        i, name, seq, bq = line.split("\t")
        recs.append((int(i), name, seq, [int(x) for x in bq.split(",") if x]))
    return recs


def test_bioparser_fasta_fastq_plain_and_gz(io_test, tmp_path):
    rng = np.random.default_rng(3)
    names = [f"read{i} some description {i}" for i in range(40)]
    seqs = ["".join(rng.choice(list("ACGTacgtN"), int(rng.integers(1, 700)))) for _ in names]
    quals = ["".join(chr(33 + int(q)) for q in rng.integers(0, 60, len(s))) for s in seqs]
    fa = tmp_path / "r.fasta"
    with open(fa, "w") as f:
        for n, s in zip(names, seqs):
            f.write(f">{n}\n")
            for k in range(0, len(s), 60):          # wrapped lines
                f.write(s[k:k + 60] + "\n")
    fq = tmp_path / "r.fastq.gz"
    with gzip.open(fq, "wt") as f:
        for n, s, q in zip(names, seqs, quals):
            f.write(f"@{n}\n{s}\n+\n{q}\n")
    for kind, path in (("fasta", fa), ("fastq", fq)):
        for chunk in ([], ["5000"]):
            out = subprocess.run([io_test, "parse", kind, str(path)] + chunk, check=True,
                                 stdout=subprocess.PIPE).stdout
            recs = _records(out)
            assert [r[1] for r in recs] == [n.split()[0] for n in names]   # cut at white space
            want = seqio.pack_ascii([s.encode() for s in seqs],
                                    [q.encode() for q in quals] if kind == "fastq" else None)
            for i, r in enumerate(recs):
                assert r[0] == i
                assert r[2] == want.ascii(i).decode()
                if kind == "fastq":
                    a, b = int(want.bq_off[i]), int(want.bq_off[i + 1])
                    assert r[3] == want.block_quality[a:b].tolist()
    # error behaviour: std::invalid_argument like the reference's parser
    bad = tmp_path / "bad.fastq"
    bad.write_text("@x\nACGT\n+\nII\n")
    p = subprocess.run([io_test, "parse", "fastq", str(bad)], stdout=subprocess.PIPE)
    assert p.returncode == 3 and b"invalid_argument" in p.stdout
    p = subprocess.run([io_test, "parse", "fasta", str(tmp_path / "missing.fa")],
                       stdout=subprocess.PIPE)
    assert p.returncode == 3 and b"unable to open file" in p.stdout
    nuc = tmp_path / "nuc.fasta"
    nuc.write_text(">x\nAC!T\n")
    p = subprocess.run([io_test, "parse", "fasta", str(nuc)], stdout=subprocess.PIPE)
    assert p.returncode == 3 and b"not a nucleotide" in p.stdout


def test_cereal_archives_round_trip(io_test, tmp_path):
    p = subprocess.run([io_test, "cereal", str(tmp_path / "x.cereal")], stdout=subprocess.PIPE)
    assert p.returncode == 0 and p.stdout.decode().strip().endswith("ok")
    import json
    doc = json.loads(p.stdout.decode().rsplit("\n", 2)[0])
    assert doc["42"]["id"] == 10 and doc["root"]["value0"] == -3
