"""CPU model of the windowed alignment path (csrc/kernels_dpw.cu), checked against the oracle's full matrices.

The CUDA path does not fill the whole A x L matrix for a read whose adapter was not found verbatim: a unit-cost
edit-distance filter flags the columns where an ACCEPTED alignment can end, the Gotoh recurrence runs only on windows
around them from a -inf border, and the end-cell rule is applied to what the windows produced.  The claim is that
this is exact: every read gets the accept decision of the full matrices (/root/reference/src/lib.rs:155-160), and
every accepted read the same (score, length).  The GPU tests check the kernels per read (tests/test_gpu_windowed.py);
this file checks the SCHEME on the CPU, with the library's own accept bound T and edit budget K
(vfb_debug_window_plan — the host code vfb_create runs, no GPU needed) and a plain Python restatement of the steps:

  (1) accepted  =>  the end cell's column is flagged (ed <= K), and column L when the alignment ends in the last column;
  (2) windows [first - A - K, last] over groups of flagged columns, DP from a -inf left border (true border at column 1);
  (3) leftmost best last-row cell over all windows, last-column rule when a window reaches column L, accept test.
"""
import ctypes as C
import random

import numpy as np
import pytest

import oracle
from vfind_b200 import api

NEG = -(10 ** 9)
_CODE = {}
for _i, _c in enumerate("ATCG"):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i


@pytest.fixture(scope="module")
def lib():
    return api.load_library()


def plan(lib, sc, A, thr):
    t, k = C.c_int32(0), C.c_int32(0)
    lib.vfb_debug_window_plan.argtypes = [C.c_int32] * 4 + [C.c_uint32, C.c_double, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    assert lib.vfb_debug_window_plan(sc[0], sc[1], sc[2], sc[3], A, thr, C.byref(t), C.byref(k)) == 0
    return t.value, k.value


def edit_distance_columns(ad, read):
    """ed[j], j = 0..L: unit-cost edit distance between the whole adapter and the best read substring ending at column j
    (free start in the read) — what Myers' bit-vector recurrence tracks.  A byte outside ACGT (either case) matches nothing."""
    ca = np.array([_CODE.get(b, -1) for b in ad])
    cr = np.array([_CODE.get(b, -2) for b in read])
    L = len(read)
    ar = np.arange(L + 1)
    prev = np.zeros(L + 1, dtype=np.int64)                  # D[0][j] = 0
    for i in range(1, len(ad) + 1):
        cur = np.empty(L + 1, dtype=np.int64)
        cur[0] = i
        cur[1:] = np.minimum(prev[:-1] + (cr != ca[i - 1]), prev[1:] + 1)
        cur = np.minimum.accumulate(cur - ar) + ar         # ... and D[i][j-1] + 1, all the way along the row
        prev = cur
    return prev


def window_cells(ad, read, j0, j1, sc):
    """The recurrence of the oracle (default rules) on columns j0..j1 only.  Returns ({j: (score, len)} of the last row,
    [(score, len)] of rows 1..A in column j1)."""
    match, mismatch, go, ge = sc
    A = len(ad)
    left = (0, 0) if j0 == 1 else (NEG, 0)                  # column 0 is the true border; elsewhere -inf
    H = [(0, 0)] + [left] * A
    E = [(NEG, 0)] * (A + 1)
    last_row = {}
    for j in range(j0, j1 + 1):
        rb = _CODE.get(read[j - 1])
        newH = [(0, 0)] * (A + 1)                           # H[0][j] = 0: the read's start is free
        F = (NEG, 0)
        for i in range(1, A + 1):
            eo, ee = H[i][0] - go, E[i][0] - ge
            E[i] = (eo, H[i][1] + 1) if eo > ee else (ee, E[i][1] + 1)
            fo, fe = newH[i - 1][0] - go, F[0] - ge
            F = (fo, newH[i - 1][1] + 1) if fo > fe else (fe, F[1] + 1)
            ab = _CODE.get(ad[i - 1])
            w = 0 if (ab is None or rb is None) else (match if ab == rb else mismatch)
            d = H[i - 1][0] + w
            e, f = E[i][0], F[0]
            if d >= e and d >= f:
                newH[i] = (d, H[i - 1][1] + 1)
            else:
                newH[i] = F if f >= e else E[i]
        H = newH
        last_row[j] = H[A]
    return last_row, H[1:]


def windowed(ad, read, sc, T, K):
    """(accepted, score, length) as the filter + windows + resolve produce them; score / length are None when no window
    saw a cell (k2_resolve then leaves the read alone)."""
    A, L = len(ad), len(read)
    ed = edit_distance_columns(ad, read)
    flagged = [j for j in range(1, L + 1) if ed[j] <= K]
    span = A + K
    groups = []
    for j in flagged:                                       # k2_filter: a group closes when the next column is > span away
        if groups and j - groups[-1][1] <= span:
            groups[-1][1] = j
        else:
            groups.append([j, j])
    best = None                                             # (score, -j, len): max score, then the leftmost column
    col = None
    for first, last in groups:
        j0 = max(1, first - span)
        row, lastcol = window_cells(ad, read, j0, last, sc)
        for j, (s, ln) in row.items():
            if best is None or s > best[0] or (s == best[0] and j < -best[1]):
                best = (s, -j, ln)
        if last == L:
            cb = None
            for s, ln in lastcol:                           # rows 1..A, the first maximum
                if cb is None or s > cb[0]:
                    cb = (s, ln)
            col = cb
    if best is None:
        return False, None, None, ed, flagged
    score, bestj, length = best[0], -best[1], best[2]
    if col is not None and (col[0] > score or (col[0] == score and bestj == L)):
        score, length = col
    return score >= T, score, length, ed, flagged


def mutate(rng, ad, n_edits, alphabet):
    m = bytearray(ad)
    for _ in range(n_edits):
        op, pos = rng.random(), rng.randrange(len(m)) if m else 0
        if op < 0.4 and m:
            m[pos] = rng.choice(alphabet)
        elif op < 0.7 and m:
            del m[pos]
        else:
            m.insert(pos, rng.choice(alphabet))
    return bytes(m)


CONFIGS = [((3, -2, 5, 2), 20, 0.75), ((3, -2, 5, 2), 18, 0.75), ((3, -2, 5, 2), 40, 0.75), ((3, -2, 5, 2), 33, 0.8),
           ((1, -1, 1, 1), 24, 0.8), ((2, -3, 4, 1), 20, 0.85), ((5, -4, 10, 1), 16, 0.85), ((3, -2, 5, 5), 28, 0.75),
           ((1, -1, 2, 1), 64, 0.9), ((3, -2, 5, 2), 12, 0.9)]


def test_plan_matches_the_documented_shapes(lib):
    assert plan(lib, (3, -2, 5, 2), 20, 0.75) == (46, 5)             # C3: min = 45.0, five edits fit in the budget of 14
    assert plan(lib, (3, -2, 5, 2), 40, 0.75) == (91, 13)            # C5's adapters at the default thresholds
    assert plan(lib, (3, -2, 5, 2), 40, 0.6)[1] == -1                # C5 as configured: 3K > A, the full matrices run
    assert plan(lib, (3, 3, 5, 2), 20, 0.75)[1] == -1                # a mismatch that costs nothing: no edit bound
    assert plan(lib, (3, -2, 0, 2), 20, 0.75)[1] == -1 and plan(lib, (3, -2, 5, 2), 65, 0.75)[1] == -1


@pytest.mark.parametrize("sc,A,thr", CONFIGS)
def test_windowed_scheme_equals_full_matrices(lib, sc, A, thr):
    T, K = plan(lib, sc, A, thr)
    assert K >= 0, "pick configurations the filter applies to"
    rng = random.Random(A * 1000 + int(thr * 100) + sc[0])
    accepted = rejected = col_ends = 0
    n_reads = 900 if A <= 40 else 300
    for it in range(n_reads):
        alphabet = b"ACGT" if it % 5 else b"ACGTNacgt"
        ad = bytes(rng.choice(b"ACGT") for _ in range(A))
        L = rng.choice([A // 2, A, A + 3, 2 * A, 90, 140])
        read = bytearray(rng.choice(alphabet) for _ in range(L))
        for _ in range(rng.choice([0, 1, 1, 2])):           # planted copies with up to K + 3 edits, possibly over an end
            m = mutate(rng, ad, rng.randint(0, K + 3), alphabet)
            p = rng.randint(-len(m) // 3, L - (2 * len(m)) // 3)
            if rng.random() < 0.35:                          # a few bases hanging over the read's right (or left) end
                over = rng.randint(1, max(1, A // 6))
                p = L - len(m) + over if rng.random() < 0.7 else -over
            lo, hi = max(0, p), min(L, p + len(m))
            if hi > lo:
                read[lo:hi] = m[lo - p:hi - p]
        read = bytes(read[:L]) or b"A"
        fs, fl, fi, fj = oracle.sg_stats(ad, read, *sc)
        ok, ws, wl, ed, flagged = windowed(ad, read, sc, T, K)
        if fs >= T:
            accepted += 1
            # (1) the end cell of an accepted alignment sits in a flagged column
            assert ed[fj] <= K, (ad, read, fs, fj, ed[fj], K)
            if fi < A:
                col_ends += 1
                assert fj == len(read) and ed[len(read)] <= K
            # (2) + (3) the windows reproduce the full matrices' answer
            assert ok and (ws, wl) == (fs, fl), (ad, read, (fs, fl, fi, fj), (ws, wl))
        else:
            rejected += 1
            assert not ok, (ad, read, fs, ws)
            assert ws is None or ws <= fs                   # every window value is a lower bound
    assert accepted >= n_reads // 12 and rejected >= n_reads // 8 and col_ends >= 3, (accepted, rejected, col_ends)


@pytest.mark.parametrize("sc,A,thr", CONFIGS)
def test_edit_budget_holds_for_the_cheapest_edits(lib, sc, A, thr):
    """Random mutations rarely spend the budget the cheapest way.  The cheapest edits are the ones K is computed from:
    one long insertion (open + extend + extend ...), one long deletion, a run of substitutions / wildcards, an overhang —
    planted here at every length up to past the budget; whatever the full matrices accept must come out of the windows."""
    T, K = plan(lib, sc, A, thr)
    match, mismatch, go, ge = sc
    rng = random.Random(7 * A + sc[2])
    other = {65: b"CGT", 67: b"AGT", 71: b"ACT", 84: b"ACG"}
    worst = 0
    for it in range(40):
        ad = bytes(rng.choice(b"ACGT") for _ in range(A))
        lead = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 30)))
        tail = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 30)))
        cut = rng.randint(2, A - 2)
        variants = []
        for n in range(1, K + 4):
            ins = bytes(rng.choice(other[ad[cut]]) for _ in range(n))                       # n read bases the adapter lacks
            variants.append(lead + ad[:cut] + ins + ad[cut:] + tail)
            if cut + n < A:
                variants.append(lead + ad[:cut] + ad[cut + n:] + tail)                      # n adapter bases the read lacks
            sub = bytearray(ad)
            for q in rng.sample(range(A), min(n, A)):
                sub[q] = rng.choice(other[ad[q]]) if rng.random() < 0.7 else ord("N")       # substitutions and wildcards
            variants.append(lead + bytes(sub) + tail)
            variants.append(lead + ad[:A - n])                                               # n bases over the right end
            variants.append(ad[n:] + tail)                                                   # n bases over the left end
        for read in variants:
            if not read:
                continue
            fs, fl, fi, fj = oracle.sg_stats(ad, read, *sc)
            ok, ws, wl, ed, _ = windowed(ad, read, sc, T, K)
            if fs >= T:
                worst = max(worst, int(ed[fj]))
                assert ed[fj] <= K, (ad, read, fs, fj, int(ed[fj]), K)
                assert ok and (ws, wl) == (fs, fl), (ad, read, (fs, fl, fi, fj), (ws, wl))
            else:
                assert not ok
    # the budget is not just an upper bound: some accepted alignment needs all of it (K - 1 would lose that read)
    assert worst == K, (worst, K)
