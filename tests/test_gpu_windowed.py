"""GPU parity tests of the windowed DP (kernels_dpw.cu: bit-parallel filter -> DP on windows -> resolve).

The windowed path is what `find_variants` runs by default (non-diagnostics contexts).  Its contract:
for every alignment the reference ACCEPTS (src/lib.rs:157) score, length and therefore the region
boundary are bit-identical to the full DP; every other alignment is rejected.  `dp_mode=2` runs it
with diagnostics on so both halves can be checked per read against the oracle; `dp_mode=1` is the
full DP.  Inputs lean on what could break a window argument: several adapter-like sites per read,
adapters hanging over either read end, low-complexity adapters, equal best scores at different
columns, wildcards, reads around the adapter length, a full window list.
"""
import random

import numpy as np
import pytest

import oracle
from vfind_b200 import api

from test_gpu_parity import PREFIX, SUFFIX, gpu_run, make_reads, mutate, oracle_run

pytestmark = pytest.mark.gpu

def accept_bound(thr, match, A):
    # score as f64 > (thr*match)*len  <=>  score >= floor(min)+1      (src/lib.rs:157, :260-261)
    return int(np.floor((thr * float(match)) * float(A))) + 1


def check_windowed(seqs, adapters, expect_windowed=True, **kw):
    """dp_mode=2 diagnostics vs the oracle (per read), and the default-mode table vs the oracle."""
    otable, od, cells = oracle_run(seqs, adapters, **kw)
    table, d, stats = gpu_run(seqs, adapters, dp_mode=2, **kw)
    if expect_windowed:
        assert stats["dp_kernel_kind"] == 3
    m = kw.get("match_score", 3)
    for side, thr_key, bound in (("prefix", "accept_prefix_alignment", "start"), ("suffix", "accept_suffix_alignment", "end")):
        T = accept_bound(kw.get(thr_key, 0.75), m, len(adapters[0 if side == "prefix" else 1]))
        aligned = od["exact_" + side] < 0                        # no exact hit: the reference aligns (-1 = none)
        acc = aligned & (od["score_" + side] >= T) & (od["len_" + side] >= 0)
        assert (d["exact_" + side] == od["exact_" + side]).all()
        assert (d[bound] == od[bound]).all(), (side, np.nonzero(d[bound] != od[bound])[0][:5])
        bad = np.nonzero(acc & ((d["score_" + side] != od["score_" + side]) | (d["len_" + side] != od["len_" + side])))[0]
        assert bad.size == 0, (side, bad[:5], [seqs[i] for i in bad[:2]])
        rej = aligned & ~acc
        assert (d["score_" + side][rej] < T).all()          # lower bounds, never an accept
    assert table == otable
    assert stats["dp_cells"] == cells                       # algorithmic cells (full matrices)
    # default mode, no diagnostics: what find_variants runs
    t2, _, s2 = gpu_run(seqs, adapters, want_diag=False, **kw)
    assert t2 == otable
    if expect_windowed:
        assert s2["dp_kernel_kind"] == 3 and s2["dp_cells_computed"] < max(1, s2["dp_cells"]) + 1
    # and the full DP agrees
    t3, _, s3 = gpu_run(seqs, adapters, want_diag=False, dp_mode=1, **kw)
    assert t3 == otable and s3["dp_kernel_kind"] in (1, 2)
    return stats


@pytest.mark.parametrize("scoring", [(3, -2, 5, 2), (2, -3, 4, 1), (5, -4, 10, 1), (1, -1, 1, 1), (4, -6, 1, 3),
                                     (3, -1, 2, 1), (10, -9, 7, 7)])
@pytest.mark.parametrize("thr", [0.75, 0.6, 0.9, 0.5])
def test_windowed_random_scoring(scoring, thr):
    rng = random.Random(hash((scoring, thr)) & 0xFFFF)
    seqs = make_reads(rng, PREFIX, SUFFIX, 3000, lead=(0, 40))
    kw = dict(match_score=scoring[0], mismatch_score=scoring[1], gap_open_penalty=scoring[2],
              gap_extend_penalty=scoring[3], accept_prefix_alignment=thr, accept_suffix_alignment=thr)
    check_windowed(seqs, (PREFIX, SUFFIX), expect_windowed=False, **kw)


@pytest.mark.parametrize("A", [4, 5, 7, 8, 9, 12, 16, 17, 19, 20, 21, 24, 27, 28, 29, 31, 32])
def test_windowed_adapter_lengths(A):
    rng = random.Random(2000 + A)
    pre = bytes(rng.choice(b"ACGT") for _ in range(A))
    suf = bytes(rng.choice(b"ACGT") for _ in range(A))
    seqs = make_reads(rng, pre, suf, 1500, lead=(0, 30))
    seqs += [b"A", b"AC", pre[: max(1, A - 1)], pre, suf, pre + suf, pre + b"ACG" + suf, pre[1:], suf[:-1]]
    check_windowed(seqs, (pre, suf), expect_windowed=False, accept_prefix_alignment=0.6, accept_suffix_alignment=0.7)


@pytest.mark.parametrize("A", [33, 34, 36, 39, 40, 41, 47, 48, 50, 56, 59, 63, 64])
def test_windowed_long_adapters(A):
    """Adapters of 33..64 bases: the filter runs Myers' recurrence on one 64-bit word per vector, the window kernel has
    up to 64 register-resident rows (BASELINE config 5's 40-nt adapters; src/lib.rs:155-160)."""
    rng = random.Random(4000 + A)
    pre = bytes(rng.choice(b"ACGT") for _ in range(A))
    suf = bytes(rng.choice(b"ACGT") for _ in range(A))
    seqs = make_reads(rng, pre, suf, 1500, lead=(0, 60))
    for _ in range(300):        # heavier damage, several sites, overhangs
        a = mutate(rng, pre, max_edits=8)
        b = mutate(rng, suf, max_edits=8)
        body = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(0, 80)))
        seqs.append((a + body + b)[rng.randrange(0, 10):])
        seqs.append(a + body + a + body + b)
    seqs += [b"A", pre[: A - 1], pre, suf, pre + suf, pre + b"ACG" + suf, pre[1:], suf[:-1]]
    st = check_windowed(seqs, (pre, suf), expect_windowed=True, accept_prefix_alignment=0.8, accept_suffix_alignment=0.75)
    assert st["dp_windows"] > 0
    # at 0.6 the budget admits too many edits for the filter to be selective: the full-matrix kernel runs, same answers
    check_windowed(seqs[:800], (pre, suf), expect_windowed=False, accept_prefix_alignment=0.6, accept_suffix_alignment=0.6)


def test_windowed_applies_at_defaults():
    rng = random.Random(3)
    seqs = make_reads(rng, PREFIX, SUFFIX, 2000, lead=(0, 100))
    st = check_windowed(seqs, (PREFIX, SUFFIX))
    assert st["dp_windows"] > 0 and st["dp_cells_computed"] < st["dp_cells"]


def test_windowed_multiple_sites_and_overhangs():
    # several adapter-like sites per read (equal and unequal scores: the leftmost best end cell must win
    # across windows), adapters cut by either read end (first column border, last-column rule)
    rng = random.Random(4)
    pre, suf = b"ACGTTGCATGCCGATAGCTA", b"TTGACCGGATATCCGTAGGA"
    seqs = []
    for _ in range(4000):
        parts = []
        for _s in range(rng.randrange(1, 5)):
            parts.append(bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(0, 60))))
            ad = pre if rng.random() < 0.5 else suf
            parts.append(mutate(rng, ad, max_edits=4))
        parts.append(bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(0, 60))))
        read = b"".join(parts)
        cut = rng.random()
        if cut < 0.25:
            read = read[rng.randrange(0, 12):]
        elif cut < 0.5:
            read = read[: max(1, len(read) - rng.randrange(0, 12))]
        seqs.append(read)
    # overhanging adapters at both ends, every cut length
    v = b"ATGGCGGGCATCTGTGCACTT"
    for k in range(0, 12):
        seqs.append(pre[k:] + v + suf[: len(suf) - k])
        seqs.append(mutate(rng, pre)[k:] + v + mutate(rng, suf)[: len(suf) - k])
    for thr in (0.75, 0.5):      # at 0.5 the filter is not selective enough and the full DP runs
        check_windowed(seqs, (pre, suf), expect_windowed=thr == 0.75, accept_prefix_alignment=thr,
                       accept_suffix_alignment=thr)


@pytest.mark.parametrize("pre,suf", [
    (b"ACACACACACACACACACAC", b"GTGTGTGTGTGTGTGTGTGT"),
    (b"AAAAAAAAAAAAAAAAAAAA", b"CCCCCCCCCCCCCCCCCCCC"),
    (b"AAAAAAAAAACCCCCCCCCC", b"ACGACGACGACGACGACGAC"),
    (b"ACGTNNACGTACGTACGTAC", b"ttgaccggatatccgtagga"),
])
def test_windowed_low_complexity_and_wildcards(pre, suf):
    rng = random.Random(hash((pre, suf)) & 0xFFFF)
    seqs = []
    for _ in range(3000):
        lead = bytes(rng.choice(b"ACGTN" if rng.random() < 0.1 else b"ACGT") for _ in range(rng.randrange(0, 40)))
        body = bytes(rng.choice(b"ACGT") for _ in range(rng.choice((21, 24, 30))))
        tail = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(0, 40)))
        a = mutate(rng, pre.upper(), max_edits=4, alphabet=b"ACGTNacgt")
        b = mutate(rng, suf.upper(), max_edits=4, alphabet=b"ACGTNacgt")
        seqs.append(lead + a + body + b + tail)
        if rng.random() < 0.2:      # long homopolymer / repeat runs: many flagged columns, merged windows
            seqs.append(pre.upper()[:10] * rng.randrange(1, 8) + body + suf.upper()[5:] * rng.randrange(1, 6))
    for thr in (0.75, 0.6):      # at 0.6 the filter is not selective enough and the full DP runs
        check_windowed(seqs, (pre, suf), expect_windowed=thr == 0.75, accept_prefix_alignment=thr,
                       accept_suffix_alignment=thr)


def test_windowed_ragged_lengths():
    rng = random.Random(6)
    seqs = []
    for _ in range(1500):
        L = rng.choice([1, 5, 17, 18, 19, 20, 21, 25, 40, 63, 64, 65, 127, 128, 129, 300, 700, 1500])
        body = bytearray(rng.choice(b"ACGT") for _ in range(L))
        if L > 60 and rng.random() < 0.8:
            ins = mutate(rng, PREFIX) + bytes(rng.choice(b"ACGT") for _ in range(21)) + mutate(rng, SUFFIX)
            p = rng.randrange(0, L - len(ins) + 1) if L > len(ins) else 0
            body[p:p + len(ins)] = ins
        elif L <= 25 and rng.random() < 0.7:
            src = mutate(rng, PREFIX if rng.random() < 0.5 else SUFFIX)
            k = rng.randrange(0, max(1, len(src) - L + 1))
            body = bytearray(src[k:k + L]) or body
        seqs.append(bytes(body))
    for thr in (0.75, 0.6):
        check_windowed(seqs, (PREFIX, SUFFIX), expect_windowed=thr == 0.75, accept_prefix_alignment=thr,
                       accept_suffix_alignment=thr)


def test_windowed_full_window_list_falls_back():
    # a window list too small for the batch: the overflowing reads go through the full kernel and
    # the table, the per-read boundaries and the cell accounting do not change
    rng = random.Random(7)
    seqs = make_reads(rng, PREFIX, SUFFIX, 3000, lead=(0, 60))
    otable, od, cells = oracle_run(seqs, (PREFIX, SUFFIX))
    for cap in (1, 17, 400):
        table, d, stats = gpu_run(seqs, (PREFIX, SUFFIX), dp_mode=2, debug_win_cap=cap)
        assert table == otable
        assert (d["start"] == od["start"]).all() and (d["end"] == od["end"]).all()
        assert stats["dp_cells"] == cells
        t2, _, _ = gpu_run(seqs, (PREFIX, SUFFIX), want_diag=False, debug_win_cap=cap, batch_reads=700)
        assert t2 == otable


def test_windowed_threshold_edges():
    # Q3 float edges decide K through the integer bound
    rng = random.Random(8)
    pre = bytes(rng.choice(b"ACGT") for _ in range(20))
    suf = bytes(rng.choice(b"ACGT") for _ in range(20))
    seqs = make_reads(rng, pre, suf, 2500, lead=(0, 30))
    for thr in (0.7, 0.8, 0.95, 0.999, 0.34):
        check_windowed(seqs, (pre, suf), expect_windowed=False, accept_prefix_alignment=thr, accept_suffix_alignment=thr)
