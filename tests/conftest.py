import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)


def write_fastq_gz(path, reads, members=1, crlf=False, final_newline=True):
    """reads: list of (header, seq, qual) -> gzipped FASTQ with `members` gzip members."""
    nl = "\r\n" if crlf else "\n"
    recs = ["@%s%s%s%s+%s%s%s" % (h, nl, s, nl, nl, q, nl) for h, s, q in reads]
    if not final_newline and recs:
        recs[-1] = recs[-1][:-len(nl)]
    chunks = [recs[i::members] for i in range(members)] if members > 1 else [recs]
    # keep input order: split contiguously instead of striding
    if members > 1:
        per = (len(recs) + members - 1) // members
        chunks = [recs[i * per:(i + 1) * per] for i in range(members)]
    with open(path, "wb") as f:
        for c in chunks:
            f.write(gzip.compress("".join(c).encode("latin-1")))
    return path


@pytest.fixture()
def toy_gz(tmp_path, golden):
    reads = [(r["header"], r["seq"], r["qual"]) for r in golden["toy"]["reads"]]
    return str(write_fastq_gz(tmp_path / "toy.fq.gz", reads))
