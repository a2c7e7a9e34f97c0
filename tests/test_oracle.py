"""CPU tests: the oracle against the reference's golden vectors and the SURVEY §8(c) KATs.

The DP's gapped/tie behaviour is "parity unpinned" (parasail is not available); vectors
labelled rule-discriminating document what each recalled rule predicts.
"""
import random

import pytest

import oracle
from oracle import twin

PREFIX = b"GGGCCCAGCCGGCCGGAT"
SUFFIX = b"CCGGAGGCGGAGGTTCAG"


def test_tables_match_reference_literals(golden):
    # src/lib.rs:52-77 and :86-95
    assert oracle.aa_table() == golden["aa_table_canonical"]
    assert oracle.ascii_to_index() == golden["ascii_to_index"]
    assert twin._AA == golden["aa_table_canonical"]


def test_unit_vectors_accept_reject(golden):
    # src/lib.rs:339-394: assert score > min (pass) or min > score (fail)
    c = golden["unit_constants"]
    for v in golden["unit_vectors"]:
        a, s = v["adapter"].encode(), v["seq"].encode()
        score, length, _, _ = oracle.sg_stats(a, s, c["MATCH_SCORE"], c["MISMATCH_SCORE"],
                                              c["GAP_OPEN_PENALTY"], c["GAP_EXTEND_PENALTY"])
        mn = oracle.min_score(c["ACCEPT_ALIGNMENT"], c["MATCH_SCORE"], len(a))
        assert mn == 40.5
        if v["accept"]:
            assert float(score) > mn, v["name"]
        else:
            assert mn > float(score), v["name"]
        assert twin.sg_stats(a, s)[:2] == (score, length)


def test_unit_vector_known_scores(golden):
    # SURVEY §8(c) known-answer table (derived with the rule set; single best cell each)
    exp = {"test_good_prefix_alignment": (49, 18), "test_bad_prefix_alignment": (34, 18),
           "test_good_suffix_alignment": (49, 18), "test_bad_suffix_alignment": (29, 18)}
    for v in golden["unit_vectors"]:
        got = oracle.sg_stats(v["adapter"].encode(), v["seq"].encode())[:2]
        assert got == exp[v["name"]], (v["name"], got)


def test_toy_table_reads(golden):
    toy = golden["toy"]
    seqs = [r["seq"].encode() for r in toy["reads"]]
    text, off, ln = oracle.pack_reads(seqs)
    p = oracle.make_params(toy["adapters"])
    table, diag, cells = oracle.process_reads(p, text, off, ln, want_diag=True)
    assert table == {k.encode(): v for k, v in toy["table"]}
    assert twin.find_variants_reads(seqs, toy["adapters"]) == table
    # per-read expectations from SURVEY §8(c)
    names = [r["header"].split()[0] for r in toy["reads"]]
    d = {n: diag[i] for i, n in enumerate(names)}
    assert (d["s0"]["start"], d["s0"]["end"]) == (18, 39)
    assert d["s0"]["exact_prefix"] == 0 and d["s0"]["exact_suffix"] == 39
    for n in ("s1", "s2", "s3"):
        assert (d[n]["start"], d[n]["end"]) == (18, 39)
    assert d["s1"]["exact_prefix"] == -1 and d["s1"]["len_prefix"] == 18
    assert d["s2"]["exact_suffix"] == -1 and d["s2"]["len_suffix"] == 18
    assert (d["seq4"]["score_prefix"], d["seq4"]["score_suffix"]) == (34, 29)
    assert d["seq4"]["start"] == -1 and d["seq4"]["end"] == -1
    assert d["seq5"]["start"] == -1 and d["seq5"]["end"] == 39
    assert d["seq6"]["start"] == 18 and d["seq6"]["end"] == -1
    assert (d["s7"]["start"], d["s7"]["end"]) == (18, 41)     # 23 nt -> dropped
    assert cells == 18 * 57 * 8  # s1,s3 prefix; s2,s3 suffix; seq4 both; seq5 prefix; seq6 suffix


def test_toy_file(toy_gz, golden):
    toy = golden["toy"]
    table = oracle.find_variants_file(toy_gz, toy["adapters"], n_threads=3)
    assert table == {k.encode(): v for k, v in toy["table"]}


def test_multimember_crlf_nofinalnewline(tmp_path, golden):
    from conftest import write_fastq_gz
    toy = golden["toy"]
    reads = [(r["header"], r["seq"], r["qual"]) for r in toy["reads"]]
    exp = {k.encode(): v for k, v in toy["table"]}
    p = write_fastq_gz(tmp_path / "m.fq.gz", reads, members=3)       # CHANGELOG.md:30-31
    assert oracle.find_variants_file(str(p), toy["adapters"]) == exp
    p = write_fastq_gz(tmp_path / "c.fq.gz", reads, crlf=True)
    assert oracle.find_variants_file(str(p), toy["adapters"]) == exp
    p = write_fastq_gz(tmp_path / "n.fq.gz", reads, final_newline=False)
    assert oracle.find_variants_file(str(p), toy["adapters"]) == exp


def test_file_errors(tmp_path, toy_gz, golden):
    ad = golden["toy"]["adapters"]
    with pytest.raises(FileNotFoundError):
        oracle.find_variants_file(str(tmp_path / "missing.fq.gz"), ad)
    plain = tmp_path / "plain.fq"
    plain.write_text("@r\nACGT\n+\nFFFF\n")
    with pytest.raises(RuntimeError):
        oracle.find_variants_file(str(plain), ad)          # plain text is not gzip (Q11)
    import gzip
    bad = tmp_path / "bad.fq.gz"
    bad.write_bytes(gzip.compress(b"@r\nACGT\n+\nFFF\n"))
    with pytest.raises(RuntimeError):
        oracle.find_variants_file(str(bad), ad)            # seq/qual length mismatch
    empty = tmp_path / "empty.fq.gz"
    empty.write_bytes(gzip.compress(b""))
    assert oracle.find_variants_file(str(empty), ad) == {}  # Q12
    for thr in (0.0, -0.1, 1.5, float("nan")):
        with pytest.raises(ValueError, match=golden["threshold_error"]):
            oracle.find_variants_file(toy_gz, ad, accept_prefix_alignment=thr)


def test_threshold_floats():
    # IEEE-double behaviour of (thr*match)*len
    assert oracle.min_score(0.75, 3, 18) == 40.5
    assert oracle.min_score(0.6, 3, 40) == 72.0
    assert oracle.min_score(0.7, 3, 20) == 0.7 * 3.0 * 20.0 < 42.0
    assert oracle.min_score(0.8, 3, 20) == 0.8 * 3.0 * 20.0 > 48.0
    assert oracle.threshold_preflight(1.0) is True
    assert oracle.threshold_preflight(0.5) is False
    for bad in (0.0, 1.0000001, -1.0, float("nan"), float("inf")):
        with pytest.raises(ValueError):
            oracle.threshold_preflight(bad)


def test_translate_all_codons():
    aa = oracle.aa_table()
    for i, c1 in enumerate("ACGT"):
        for j, c2 in enumerate("ACGT"):
            for k, c3 in enumerate("ACGT"):
                cod = c1 + c2 + c3
                want = aa[i * 16 + j * 4 + k].encode()
                for variant in (cod, cod.lower(), cod.replace("T", "U"), cod.lower().replace("t", "u")):
                    assert oracle.translate(variant.encode()) == want
                    assert twin.translate(variant.encode()) == want
    assert oracle.translate(b"ATGNNNTAA") == b"MX*"
    assert oracle.translate(b"AT") is None and oracle.translate(b"ATGA") is None
    assert oracle.translate(b"") == b""
    assert oracle.translate(b"A\xc3\xa9ATG") == b"XM"
    assert oracle.translate(bytes([65, 200, 65])) == b"X"
    rng = random.Random(7)
    for _ in range(200):
        s = bytes(rng.randrange(256) for _ in range(3 * rng.randrange(1, 20)))
        assert oracle.translate(s) == twin.translate(s)


def test_utf8():
    good = [b"", b"ACGT", "é".encode(), "€".encode(), "😀".encode(), b"\xf4\x8f\xbf\xbf"]
    bad = [b"\x80", b"\xc0\x80", b"\xc1\xbf", b"\xe0\x80\x80", b"\xed\xa0\x80", b"\xf0\x80\x80\x80",
           b"\xf4\x90\x80\x80", b"\xf5\x80\x80\x80", b"\xc3", b"\xe2\x82", b"A\xffC"]
    for g in good:
        assert oracle.is_utf8(g), g
    for b in bad:
        assert not oracle.is_utf8(b), b
    rng = random.Random(3)
    for _ in range(2000):
        s = bytes(rng.choice([0x41, 0x80, 0xbf, 0xc2, 0xe0, 0xed, 0xf0, 0xf4, 0xa0, 0x90, 0x9f, 0x8f])
                  for _ in range(rng.randrange(0, 6)))
        try:
            s.decode("utf-8")
            ok = True
        except UnicodeDecodeError:
            ok = False
        assert oracle.is_utf8(s) == ok, s


def test_memmem():
    assert oracle.memmem(b"AAACGTACGT", b"ACGT") == 2
    assert oracle.memmem(b"ACGT", b"ACGTA") == -1
    assert oracle.memmem(b"acgt", b"ACGT") == -1          # case-sensitive (Q9)
    assert oracle.memmem(b"", b"A") == -1
    rng = random.Random(5)
    for _ in range(500):
        h = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(0, 60)))
        n = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(1, 5)))
        assert oracle.memmem(h, n) == h.find(n)


# ---- rule-discriminating KATs: document what each recalled rule predicts
def test_kat_end_cell_tie():
    a, r = b"TCTCCAGGTAAG", b"TCTCCAGGTAAAGCGGGCTCAT"
    assert oracle.sg_stats(a, r)[:2] == (31, 12)                                   # leftmost (assumed)
    assert oracle.sg_stats(a, r, rules=oracle.DpRules(end_rule=1))[:2] == (31, 13)  # rightmost


def test_kat_end_rule_column_first():
    # end_rule 3: a last-column cell of a row < A that TIES the best last-row cell wins.
    # Here the last row's best cell (12, 10) and the last-column cell (11, 11) both score 15 with lengths 10 / 11.
    a, r = b"CGACTAACGGGG", b"ACTAGCACGGT"
    assert oracle.sg_stats(a, r) == (15, 10, 12, 10)                                     # row first (assumed)
    assert oracle.sg_stats(a, r, rules=oracle.DpRules(end_rule=3)) == (15, 11, 11, 11)   # column first
    assert twin.sg_stats(a, r, end_rule=3) == (15, 11, 11, 11)


def test_end_rule_sensitivity_count():
    """How often would the two candidate end-cell orders (0 = row first, column only if strictly better or the corner;
    3 = column rows 1..A-1 first) give find_variants a different boundary?  Counted over accepted alignments of
    C3-like reads at the default scoring and at threshold 0.6 (parity unpinned without parasail)."""
    rng = random.Random(2024)
    diff = {0.75: 0, 0.6: 0}
    accepted = {0.75: 0, 0.6: 0}
    for it in range(3000):
        inst = bytearray(PREFIX)
        for _ in range(rng.randrange(1, 4)):
            k = rng.randrange(len(inst))
            op = rng.randrange(3)
            if op == 0:
                inst[k] = rng.choice(b"ACGT")
            elif op == 1:
                del inst[k]
            else:
                inst.insert(k, rng.choice(b"ACGT"))
        # suffix-side geometry too: the adapter instance at the very end of the read, possibly cut short
        tail_cut = rng.randrange(0, 4)
        body = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(20, 60)))
        read = body + bytes(inst)[:len(inst) - tail_cut] if rng.random() < 0.5 else bytes(inst) + body
        s0, l0 = oracle.sg_stats(PREFIX, read)[:2]
        s3, l3 = oracle.sg_stats(PREFIX, read, rules=oracle.DpRules(end_rule=3))[:2]
        assert s0 == s3
        for thr in (0.75, 0.6):
            if s0 > thr * 3 * len(PREFIX):
                accepted[thr] += 1
                diff[thr] += l0 != l3
    # the switch exists so that one run against real parasail can settle it; the counts document the exposure
    print("end_rule 0 vs 3: boundary differs in %d / %d accepted alignments at 0.75, %d / %d at 0.6"
          % (diff[0.75], accepted[0.75], diff[0.6], accepted[0.6]))
    assert accepted[0.75] > 500 and diff[0.75] <= accepted[0.75]


def test_kat_h_priority():
    a, r = b"CTTATATGCGAG", b"CTATATTGCGAGGCAACAGCAAGGAGA"
    assert oracle.sg_stats(a, r)[:2] == (23, 13)                                    # diag > F > E (assumed)
    assert oracle.sg_stats(a, r, rules=oracle.DpRules(h_priority=1))[:2] == (23, 12)


def test_kat_q1_lead_del_ins():
    # boundary comes from the alignment LENGTH (src/lib.rs:159-160)
    var = b"ATGGCGGGCATCTGTGCACTT"
    p1 = b"GGGCCCAGCCGGCGGGAT"                                     # 1 mismatch vs PREFIX
    lead2 = b"TT" + p1 + var + SUFFIX
    t = twin.find_variants_reads([lead2], (PREFIX, SUFFIX))
    text, off, ln = oracle.pack_reads([lead2])
    tab, diag, _ = oracle.process_reads(oracle.make_params((PREFIX, SUFFIX)), text, off, ln, want_diag=True)
    assert diag[0]["start"] == 18 and tab == {} == t            # true end is 20 -> 23 nt -> dropped
    del1 = PREFIX[:9] + PREFIX[10:] + var + SUFFIX                 # one adapter base deleted
    assert oracle.sg_stats(PREFIX, del1)[:2] == (46, 18)
    ins1 = PREFIX[:9] + b"A" + PREFIX[9:] + var + SUFFIX           # one base inserted
    assert oracle.sg_stats(PREFIX, ins1)[:2] == (49, 19)
    text, off, ln = oracle.pack_reads([del1, ins1])
    tab, diag, _ = oracle.process_reads(oracle.make_params((PREFIX, SUFFIX)), text, off, ln, want_diag=True)
    assert tab == {b"MAGICAL": 1}


def _rand_case(rng, scoring_pool):
    A = rng.randrange(1, 24)
    adapter = bytes(rng.choice(b"ACGT") for _ in range(A))
    # read = junk + mutated adapter + junk, occasionally with N / lower case
    inst = bytearray(adapter)
    for _ in range(rng.randrange(0, 4)):
        if not inst:
            break
        k = rng.randrange(len(inst))
        op = rng.randrange(3)
        if op == 0:
            inst[k] = rng.choice(b"ACGT")
        elif op == 1:
            del inst[k]
        else:
            inst.insert(k, rng.choice(b"ACGT"))
    alpha = b"ACGT" if rng.random() < 0.8 else b"ACGTNacgtn"
    read = bytes(rng.choice(alpha) for _ in range(rng.randrange(0, 12))) + bytes(inst) + \
        bytes(rng.choice(alpha) for _ in range(rng.randrange(0, 12)))
    if not read:
        read = b"A"
    return adapter, read, rng.choice(scoring_pool)


def test_c_oracle_vs_python_twin_random():
    rng = random.Random(1234)
    pool = [(3, -2, 5, 2), (1, -1, 0, 0), (2, -3, 4, 1), (5, -4, 10, 1), (1, -1, 1, 1), (2, -1, 3, 0)]
    for it in range(1500):
        a, r, (m, x, o, e) = _rand_case(rng, pool)
        for rules in ({}, {"gap_tie_open": 1}, {"h_priority": 1}, {"end_rule": 1}, {"end_rule": 2}, {"end_rule": 3},
                      {"wildcard_zero": 0}):
            base = dict(gap_tie_open=0, h_priority=0, end_rule=0, wildcard_zero=1)
            base.update(rules)
            got = oracle.sg_stats(a, r, m, x, o, e, rules=oracle.DpRules(**base))
            want = twin.sg_stats(a, r, m, x, o, e, **base)
            assert got == want, (a, r, (m, x, o, e), rules)


def test_process_reads_threads_and_skip_translation():
    rng = random.Random(99)
    var_lib = [bytes(rng.choice(b"ACGT") for _ in range(rng.choice([21, 21, 21, 22, 24]))) for _ in range(12)]
    seqs = []
    for _ in range(400):
        pre, suf = bytearray(PREFIX), bytearray(SUFFIX)
        for ad in (pre, suf):
            if rng.random() < 0.4:
                k = rng.randrange(len(ad))
                op = rng.randrange(3)
                if op == 0:
                    ad[k] = rng.choice(b"ACGT")
                elif op == 1:
                    del ad[k]
                else:
                    ad.insert(k, rng.choice(b"ACGT"))
        seqs.append(bytes(pre) + rng.choice(var_lib) + bytes(suf))
    text, off, ln = oracle.pack_reads(seqs)
    for skip in (False, True):
        for thr in (0.75, 0.6, 1.0):
            p = oracle.make_params((PREFIX, SUFFIX), accept_prefix_alignment=thr,
                                   accept_suffix_alignment=thr, skip_translation=skip)
            t1, _, c1 = oracle.process_reads(p, text, off, ln, n_threads=1)
            t4, _, c4 = oracle.process_reads(p, text, off, ln, n_threads=4)
            tw = twin.find_variants_reads(seqs, (PREFIX, SUFFIX), accept_prefix_alignment=thr,
                                          accept_suffix_alignment=thr, skip_translation=skip)
            assert t1 == t4 == tw and c1 == c4
            if thr == 1.0:
                assert c1 == 0                                  # Q4: alignment disabled
    # non-UTF-8 region is dropped only when not translating (src/lib.rs:295)
    bad = PREFIX + b"AC\xffGTA" + SUFFIX
    text, off, ln = oracle.pack_reads([bad])
    t, _, _ = oracle.process_reads(oracle.make_params((PREFIX, SUFFIX), skip_translation=True), text, off, ln)
    assert t == {}
    t, _, _ = oracle.process_reads(oracle.make_params((PREFIX, SUFFIX)), text, off, ln)
    assert t == {b"XV": 1}


# ---- the CPU baseline's fast leg (oracle/sg_stats_simd.c): sixteen alignments per AVX2 vector, checked against the
# scalar statement of the same recurrence, which stays the parity checker

def _mutated(rng, ad, alpha):
    m = bytearray(ad)
    for _ in range(rng.randint(0, 4)):
        op, pos = rng.random(), rng.randrange(len(m)) if m else 0
        if op < 0.4 and m:
            m[pos] = rng.choice(alpha)
        elif op < 0.7 and m:
            del m[pos]
        else:
            m.insert(pos, rng.choice(alpha))
    return bytes(m)


@pytest.mark.skipif(not oracle.simd_available(), reason="no AVX2 on this host")
def test_simd_kernel_equals_scalar_under_every_rule():
    rng = random.Random(20261019)
    scorings = [(3, -2, 5, 2), (1, -1, 1, 1), (2, -3, 4, 1), (5, -4, 10, 1), (1, -1, 0, 0), (3, -2, 5, 5), (2, 2, 1, 1)]
    compared = 0
    for it in range(700):
        A = rng.choice([1, 2, 5, 12, 20, 20, 33, 40, 64])
        alpha = rng.choice([b"ACGT", b"ACGTN", b"AC", b"ACGTacgtNx"])
        ad = bytes(rng.choice(alpha) for _ in range(A))
        reads = []
        for _ in range(rng.randint(1, 16)):                 # ragged lengths inside one vector
            L = rng.choice([1, 2, 3, A, A + 1, 30, 60, 150]) + rng.randint(0, 3)
            r = bytearray(rng.choice(alpha) for _ in range(L))
            if rng.random() < 0.7 and L > A:
                m = _mutated(rng, ad, alpha)
                p = rng.randint(0, max(0, L - len(m)))
                r[p:p + len(m)] = m
            reads.append(bytes(r))
        sc = rng.choice(scorings)
        rules = oracle.DpRules(rng.randint(0, 1), rng.randint(0, 1), it % 4, rng.randint(0, 1))
        got = oracle.sg_stats_x16(ad, reads, *sc, rules=rules)
        assert got is not None
        for k, r in enumerate(reads):
            assert got[k] == oracle.sg_stats(ad, r, *sc, rules=rules), (ad, r, sc, it % 4)
            compared += 1
    assert compared > 5000


@pytest.mark.skipif(not oracle.simd_available(), reason="no AVX2 on this host")
def test_simd_kernel_declines_what_16_bit_lanes_cannot_hold():
    ad, reads = b"ACGTACGTAC", [b"ACGTTCGTACGG"]
    assert oracle.sg_stats_x16(ad, reads) is not None
    assert oracle.sg_stats_x16(ad, reads, match=4000) is None               # 10 rows x 4000 > 16 bit
    assert oracle.sg_stats_x16(ad, reads, gap_open=20000) is None
    assert oracle.sg_stats_x16(ad, reads, gap_extend=-1) is None
    assert oracle.sg_stats_x16(ad, [b"A" * 31000]) is None                  # lengths beyond the lanes
    assert oracle.sg_stats_x16(ad, [b""]) is None                           # vfo_sg_stats refuses it too


@pytest.mark.parametrize("skip_translation", [False, True])
def test_process_reads_simd_leg_is_identical(skip_translation):
    rng = random.Random(77 + skip_translation)
    reads = []
    for i in range(5000):
        lead = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 12)))
        region = bytes(rng.choice(b"ACGTacgtN") for _ in range(3 * rng.randint(0, 20) + (rng.random() < 0.1)))
        pre = PREFIX if rng.random() < 0.5 else _mutated(rng, PREFIX, b"ACGT")
        suf = SUFFIX if rng.random() < 0.5 else _mutated(rng, SUFFIX, b"ACGT")
        tail = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 12)))
        r = lead + pre + region + suf + tail
        if i % 97 == 0:
            r = b""                                          # empty reads never reach the matrices
        elif i % 101 == 0:
            r = r[:rng.randint(1, 10)]                       # shorter than the adapters
        reads.append(r)
    text, off, ln = oracle.pack_reads(reads)
    for kw in (dict(), dict(match_score=4000, mismatch_score=-3000), dict(rules=oracle.DpRules(1, 1, 3, 0)),
               dict(accept_prefix_alignment=1.0), dict(accept_prefix_alignment=0.5, accept_suffix_alignment=0.5)):
        p = oracle.make_params((PREFIX, SUFFIX), skip_translation=skip_translation, **kw)
        want, wdiag, wcells = oracle.process_reads(p, text, off, ln, n_threads=1, want_diag=True)
        for threads in (1, 3):
            got, gdiag, gcells = oracle.process_reads(p, text, off, ln, n_threads=threads, want_diag=True, simd=True)
            assert got == want and gcells == wcells, kw
            for f in wdiag.dtype.names:
                assert (gdiag[f] == wdiag[f]).all(), (f, kw)
        offs, data, counts = oracle.process_reads(p, text, off, ln, n_threads=2, simd=True, as_dict=False)[0]
        raw = data.tobytes()
        assert {raw[int(offs[i]):int(offs[i + 1])]: int(counts[i]) for i in range(len(counts))} == want
