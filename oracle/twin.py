"""Independent, slow, pure-Python twin of the oracle (small cases only).

TEST INFRASTRUCTURE ONLY.  Written separately from vfind_oracle.c (full matrices,
column-major traversal, tuple-valued cells) so that the two restatements check each other.
Same parity status as the C oracle: the DP tie rules are "parity unpinned".

Follows /root/reference/src/lib.rs:141-166 (find_adapter_match), :16-44 (translate),
:100-110 (threshold preflight), :260-261 (min score), :275-306 (worker + reducer).
"""
from __future__ import annotations

NEG = -(10 ** 9)

_AA = ("KNKN" "TTTT" "RSRS" "IIMI" "QHQH" "PPPP" "RRRR" "LLLL"
       "EDED" "AAAA" "GGGG" "VVVV" "*Y*Y" "SSSS" "*CWC" "LFLF")
_IDX = {ord("A"): 0, ord("a"): 0, ord("C"): 1, ord("c"): 1, ord("G"): 2, ord("g"): 2,
        ord("T"): 3, ord("t"): 3, ord("U"): 3, ord("u"): 3}
_DNA = {ord(c): i for i, c in enumerate("ACGT")}
_DNA.update({ord(c): i for i, c in enumerate("acgt")})


def translate(seq: bytes):
    if len(seq) % 3:
        return None
    out = bytearray()
    for k in range(0, len(seq), 3):
        cod = seq[k:k + 3]
        if any(b >= 128 for b in cod):
            out.append(ord("X"))
            continue
        idx = [_IDX.get(b, 4) for b in cod]
        out.append(ord("X") if 4 in idx else ord(_AA[idx[0] * 16 + idx[1] * 4 + idx[2]]))
    return bytes(out)


def _w(a, b, match, mismatch, wildcard_zero=True):
    ca, cb = _DNA.get(a), _DNA.get(b)
    if ca is None or cb is None:
        return 0 if wildcard_zero else mismatch
    return match if ca == cb else mismatch


def sg_stats(adapter: bytes, read: bytes, match=3, mismatch=-2, gap_open=5, gap_extend=2,
             gap_tie_open=0, h_priority=0, end_rule=0, wildcard_zero=1):
    """Returns (score, length, end_i, end_j).  Cells are (score, length) tuples."""
    A, L = len(adapter), len(read)
    if A == 0 or L == 0:
        raise ValueError("empty adapter or read")
    H = [[(0, 0)] * (L + 1) for _ in range(A + 1)]
    E = [[(NEG, 0)] * (L + 1) for _ in range(A + 1)]
    F = [[(NEG, 0)] * (L + 1) for _ in range(A + 1)]
    for j in range(1, L + 1):           # column-major on purpose (the C oracle is row-major)
        for i in range(1, A + 1):
            eo = H[i][j - 1][0] - gap_open
            ee = E[i][j - 1][0] - gap_extend
            if eo > ee or (gap_tie_open and eo == ee):
                E[i][j] = (eo, H[i][j - 1][1] + 1)
            else:
                E[i][j] = (ee, E[i][j - 1][1] + 1)
            fo = H[i - 1][j][0] - gap_open
            fe = F[i - 1][j][0] - gap_extend
            if fo > fe or (gap_tie_open and fo == fe):
                F[i][j] = (fo, H[i - 1][j][1] + 1)
            else:
                F[i][j] = (fe, F[i - 1][j][1] + 1)
            d = H[i - 1][j - 1][0] + _w(adapter[i - 1], read[j - 1], match, mismatch, wildcard_zero)
            e, f = E[i][j][0], F[i][j][0]
            if d >= e and d >= f:
                H[i][j] = (d, H[i - 1][j - 1][1] + 1)
            elif h_priority == 0:
                H[i][j] = F[i][j] if f >= e else E[i][j]
            else:
                H[i][j] = E[i][j] if e >= f else F[i][j]
    score, length, ei, ej = NEG, 0, A, 0
    if end_rule == 3:               # last-column cells of rows 1..A-1 are candidates before the last row
        for i in range(1, A):
            if H[i][L][0] > score:
                score, length, ei, ej = H[i][L][0], H[i][L][1], i, L
    for j in range(1, L + 1):
        s = H[A][j][0]
        if s > score or (end_rule == 1 and s >= score):
            score, length, ei, ej = s, H[A][j][1], A, j
    if end_rule not in (2, 3):
        cb, ci = NEG, 0
        for i in range(1, A + 1):
            if H[i][L][0] > cb:
                cb, ci = H[i][L][0], i
        if cb > score or (cb == score and ej == L):
            score, length, ei, ej = cb, H[ci][L][1], ci, L
    return score, length, ei, ej


def threshold_preflight(thr: float) -> bool:
    if 0.0 < thr < 1.0:
        return False
    if thr == 1.0:
        return True
    raise ValueError("Accept alignment threshold must be between 0 and 1.")


def find_adapter_match(seq: bytes, adapter: bytes, align_enabled: bool, min_score: float,
                       is_prefix: bool, scoring=(3, -2, 5, 2), **rules):
    pos = seq.find(adapter)
    if pos >= 0:
        return pos + len(adapter) if is_prefix else pos
    if not align_enabled:
        return None
    score, length, _, _ = sg_stats(adapter, seq, *scoring, **rules)
    if float(score) > min_score:
        if is_prefix:
            return length
        return len(seq) - length if length <= len(seq) else None
    return None


def find_variants_reads(seqs, adapters, match_score=3, mismatch_score=-2, gap_open_penalty=5,
                        gap_extend_penalty=2, accept_prefix_alignment=0.75,
                        accept_suffix_alignment=0.75, skip_translation=False, **rules):
    """{sequence bytes: count} over an iterable of read byte strings."""
    prefix, suffix = [a if isinstance(a, bytes) else a.encode() for a in adapters]
    pa = not threshold_preflight(accept_prefix_alignment)
    sa = not threshold_preflight(accept_suffix_alignment)
    minp = accept_prefix_alignment * float(match_score) * float(len(prefix))
    mins = accept_suffix_alignment * float(match_score) * float(len(suffix))
    sc = (match_score, mismatch_score, gap_open_penalty, gap_extend_penalty)
    table = {}
    for seq in seqs:
        start = find_adapter_match(seq, prefix, pa, minp, True, sc, **rules)
        end = find_adapter_match(seq, suffix, sa, mins, False, sc, **rules)
        if start is None or end is None or not start < end or end > len(seq):
            continue
        var = seq[start:end]
        if skip_translation:
            try:
                var.decode("utf-8")
            except UnicodeDecodeError:
                continue
            key = var
        else:
            key = translate(var)
            if key is None:
                continue
        table[key] = table.get(key, 0) + 1
    return table
