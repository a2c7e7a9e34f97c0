"""Property tests of the CPU oracle (hypothesis): the parity anchor is checked against independent statements of the
same facts — Python's own bytes.find / str decoding, a dictionary codon table built from the reference's literals, and
invariants of the alignment that hold under every rule switch."""
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

import oracle  # noqa: E402
from oracle import twin  # noqa: E402

DNA = st.binary(min_size=0, max_size=60).map(lambda b: bytes(b"ACGT"[x & 3] for x in b))
ANY = st.binary(min_size=0, max_size=60)
RULES = st.builds(oracle.DpRules, st.integers(0, 1), st.integers(0, 1), st.integers(0, 3), st.integers(0, 1))
SCORING = st.sampled_from([(3, -2, 5, 2), (1, -1, 1, 1), (2, -3, 4, 1), (5, -4, 10, 1), (3, -2, 5, 5)])


@settings(max_examples=300, deadline=None)
@given(hay=ANY, needle=st.binary(min_size=0, max_size=6))
def test_memmem_is_bytes_find(hay, needle):
    # memchr::memmem::find, src/lib.rs:148: leftmost occurrence, empty needle at 0
    assert oracle.memmem(hay, needle) == hay.find(needle)


@settings(max_examples=300, deadline=None)
@given(b=st.binary(min_size=0, max_size=24))
def test_utf8_is_pythons_strict_decoder(b):
    # String::from_utf8, src/lib.rs:295
    try:
        b.decode("utf-8", errors="strict")
        ok = True
    except UnicodeDecodeError:
        ok = False
    assert oracle.is_utf8(b) == ok


@settings(max_examples=300, deadline=None)
@given(seq=st.binary(min_size=0, max_size=45))
def test_translate_against_a_dictionary(golden, seq):
    # translate, src/lib.rs:16-44, from the golden literals: None unless len % 3 == 0; 'X' for anything outside the table
    idx = golden["ascii_to_index"]
    aa = golden["aa_table_canonical"]
    if len(seq) % 3:
        assert oracle.translate(seq) is None
        return
    want = bytearray()
    for i in range(0, len(seq), 3):
        c = [idx[b] if b < 128 else 4 for b in seq[i:i + 3]]
        want.append(ord("X") if 4 in c else ord(aa[c[0] * 16 + c[1] * 4 + c[2]]))
    assert oracle.translate(seq) == bytes(want)


@settings(max_examples=200, deadline=None)
@given(ad=DNA.filter(lambda a: len(a) >= 1), lead=DNA, tail=DNA, rules=RULES, sc=SCORING)
def test_exact_occurrence_scores_full_marks(ad, lead, tail, rules, sc):
    # a read that holds the adapter verbatim aligns with A * match over A columns, under every rule switch
    read = lead + ad + tail
    score, length, ei, ej = oracle.sg_stats(ad, read, *sc, rules=rules)
    assert score == len(ad) * sc[0] and ei == len(ad)
    assert length >= len(ad)                       # the length statistic counts alignment columns
    if rules.end_rule != 1:                        # leftmost best end cell: no later than the first occurrence's end
        assert ej <= read.find(ad) + len(ad)


@settings(max_examples=200, deadline=None)
@given(ad=DNA.filter(lambda a: len(a) >= 1), read=DNA.filter(lambda r: len(r) >= 1), rules=RULES, sc=SCORING)
def test_alignment_bounds_and_twin(ad, read, rules, sc):
    score, length, ei, ej = oracle.sg_stats(ad, read, *sc, rules=rules)
    m = min(len(ad), len(read))
    assert score <= m * sc[0]                      # at most min(A, L) matched pairs (the end cell has i, j >= 1: it can be negative)
    assert 0 <= length <= len(ad) + len(read)
    assert (ei == len(ad) and 1 <= ej <= len(read)) or (ej == len(read) and 1 <= ei <= len(ad))
    assert twin.sg_stats(ad, read, *sc, gap_tie_open=rules.gap_tie_open, h_priority=rules.h_priority, end_rule=rules.end_rule,
                         wildcard_zero=rules.wildcard_zero) == (score, length, ei, ej)
    if oracle.simd_available():
        assert oracle.sg_stats_x16(ad, [read], *sc, rules=rules) == [(score, length, ei, ej)]


@settings(max_examples=100, deadline=None)
@given(reads=st.lists(ANY, min_size=0, max_size=12), skip=st.booleans(), threads=st.integers(1, 3))
def test_process_reads_is_the_sum_of_its_reads(reads, skip, threads):
    # worker + reducer closures, src/lib.rs:275-306: the table of a batch is the sum of the one-read tables
    ad = (b"ACGTAC", b"GTTGCA")
    p = oracle.make_params(ad, skip_translation=skip)
    reads = [r + ad[0] + r[:9] + ad[1] for r in reads]
    text, off, ln = oracle.pack_reads(reads)
    got, _, _ = oracle.process_reads(p, text, off, ln, n_threads=threads, simd=(threads == 2))
    want = {}
    for r in reads:
        t, o, l = oracle.pack_reads([r])
        for k, v in oracle.process_reads(p, t, o, l)[0].items():
            want[k] = want.get(k, 0) + v
    assert got == want
