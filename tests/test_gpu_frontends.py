"""GPU tests of the callers either side of the hot path: per-read diagnostics as an
Arrow table, the progress callback behind show_progress, multi-file and plain-text front-ends."""
import gzip
import random

import numpy as np
import pytest

import oracle
from test_gpu_parity import PREFIX, SUFFIX, make_reads, oracle_run

pytestmark = pytest.mark.gpu


def _rows(out):
    cols = out.to_pydict() if hasattr(out, "to_pydict") else out.to_dict(as_series=False)
    return {k.encode(): v for k, v in zip(cols["sequence"], cols["count"])}


def _fastq(seqs, first=0):
    return "".join("@r%d\n%s\n+\n%s\n" % (first + i, s.decode(), "F" * len(s)) for i, s in enumerate(seqs)).encode()


def test_read_diagnostics_table_matches_the_oracle():
    from vfind_b200 import read_diagnostics
    rng = random.Random(41)
    seqs = make_reads(rng, PREFIX, SUFFIX, 3000, lead=(0, 30)) + [b"ACGT", PREFIX, PREFIX + SUFFIX]
    kw = dict(accept_prefix_alignment=0.7, accept_suffix_alignment=0.6)
    t = read_diagnostics(seqs, (PREFIX.decode(), SUFFIX.decode()), **kw)
    _, od, _ = oracle_run(seqs, (PREFIX, SUFFIX), **kw)
    assert t.num_rows == len(seqs)
    assert t.column_names == ["exact_prefix", "exact_suffix", "score_prefix", "len_prefix", "score_suffix", "len_suffix",
                              "accept_prefix", "accept_suffix", "start", "end", "region"]
    def col(name, fill):
        return np.array([fill if v is None else v for v in t.column(name).to_pylist()], dtype=np.int64)

    for f in ("exact_prefix", "exact_suffix", "start", "end"):
        assert (col(f, -1) == od[f]).all(), f                 # the oracle marks "none" with -1
    for side in ("prefix", "suffix"):
        ran = od["len_" + side] >= 0
        assert (np.array([v is not None for v in t.column("score_" + side).to_pylist()]) == ran).all()
        assert (col("score_" + side, 0)[ran] == od["score_" + side][ran]).all()
        assert (col("len_" + side, 0)[ran] == od["len_" + side][ran]).all()
    regions = t.column("region").to_pylist()
    for i, s in enumerate(seqs):
        st, en = od["start"][i], od["end"][i]
        want = s[st:en] if st >= 0 and en >= 0 and st < en else None
        assert regions[i] == want
    assert read_diagnostics([], (PREFIX.decode(), SUFFIX.decode())).num_rows == 0


def test_progress_callback_and_front_ends(tmp_path):
    from vfind_b200 import PanicException, find_variants, find_variants_multi
    rng = random.Random(42)
    ad = (PREFIX.decode(), SUFFIX.decode())
    a = make_reads(rng, PREFIX, SUFFIX, 12000, lib=300)
    b = make_reads(rng, PREFIX, SUFFIX, 9000, lib=300)
    pa_, pb_ = tmp_path / "a.fq.gz", tmp_path / "b.fq.gz"
    pa_.write_bytes(gzip.compress(_fastq(a)))
    pb_.write_bytes(gzip.compress(_fastq(b, len(a))))
    plain = tmp_path / "b.fq"
    plain.write_bytes(_fastq(b, len(a)))
    calls = []
    ta = _rows(find_variants(str(pa_), ad, progress=lambda r, d, t: calls.append((r, d, t))))
    assert calls and calls[-1][0] == len(a) and calls[-1][1] == calls[-1][2] == pa_.stat().st_size
    assert all(x[0] <= y[0] and x[1] <= y[1] for x, y in zip(calls, calls[1:]))
    tb = _rows(find_variants(str(pb_), ad, show_progress=False))
    both = _rows(find_variants_multi([str(pa_), pb_], ad))
    want = dict(ta)
    for k, v in tb.items():
        want[k] = want.get(k, 0) + v
    assert both == want == oracle_run(a + b, (PREFIX, SUFFIX))[0]
    # uncompressed text: the reference panics (gzip only); allow_text reads it
    with pytest.raises(PanicException):
        find_variants(str(plain), ad)
    assert _rows(find_variants(str(plain), ad, allow_text=True)) == tb
    assert _rows(find_variants_multi([str(pa_), str(plain)], ad, allow_text=True)) == want
    assert _rows(find_variants(str(pb_), ad, allow_text=True)) == tb          # gzip still recognised
    with pytest.raises(TypeError):
        find_variants_multi(str(pa_), ad)
    with pytest.raises(FileNotFoundError):
        find_variants_multi([str(pa_), str(tmp_path / "missing.fq.gz")], ad)
