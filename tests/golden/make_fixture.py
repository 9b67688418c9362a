"""Regenerate tests/golden/lambda_reads.npz and lambda_genome.npz.

Source: the reference's own test data (RavenTest/data/ERA476754.fastq.gz, 236
ONT reads of lambda phage; NC_001416.fasta.gz) read IN PLACE from
/root/reference and re-encoded in the biosoup wire format (2-bit words + one
mean-Phred byte per 64 bases) -- the exact bytes the reference would hold in
memory after parsing (RavenExe/src/main.cc:258-299).  /root/reference does not
exist on the GPU box, hence the committed re-encoding.

    python tests/golden/make_fixture.py
"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from raven_b200 import seqio  # noqa: E402

REF = "/root/reference/RavenTest/data"
HERE = os.path.dirname(os.path.abspath(__file__))

for src, dst in (("ERA476754.fastq.gz", "lambda_reads.npz"),
                 ("NC_001416.fasta.gz", "lambda_genome.npz")):
    p = os.path.join(REF, src)
    print(src, hashlib.sha256(open(p, "rb").read()).hexdigest())
    rs = seqio.parse(p)
    print("  reads", rs.n, "bases", rs.bases, "min", rs.lens.min(), "max", rs.lens.max())
    rs.save(os.path.join(HERE, dst))
