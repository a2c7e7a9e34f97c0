"""GPU tests of the packed host->device path (hostpack.cu / hostpack_cpu.cpp): host threads turn blocks of
pinned read text into 2-bit codes, the device expands them, other blocks travel as they are.  The device
text must be byte-identical to the caller's, so every per-read result and the table must equal the plain
copy's (VFB_HOST_PACK=0) and the oracle's."""
import numpy as np
import pytest

import oracle
from vfind_b200 import api

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

PRE, SUF = b"GGGCCCAGCCGGCCGGATTA", b"CCGGAGGCGGAGGTTCAGAC"


def _reads(n, L=250, seed=3):
    """n reads of L bases: lead | prefix | region | suffix | tail, some adapters damaged, N / lower-case / U
    bytes sprinkled in, a few reads entirely lower-case, and a 6 MB stretch of N (a block that cannot be packed)."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    lib = rng.choice(acgt, size=(5000, 198))
    t = rng.choice(acgt, size=(n, L))
    lead = rng.integers(0, 7, size=n)
    for ld in range(7):
        rows = np.nonzero(lead == ld)[0]
        t[rows, ld:ld + 20] = np.frombuffer(PRE, dtype=np.uint8)
        t[rows, ld + 20:ld + 218] = lib[rng.integers(0, len(lib), size=len(rows))]
        t[rows, ld + 218:ld + 238] = np.frombuffer(SUF, dtype=np.uint8)
    dmg = rng.random(n) < 0.2                                      # substitutions inside the adapters
    t[dmg, lead[dmg] + rng.integers(0, 20, size=int(dmg.sum()))] = ord("A")
    dmg = rng.random(n) < 0.2
    t[dmg, lead[dmg] + 218 + rng.integers(0, 20, size=int(dmg.sum()))] = ord("C")
    flat = t.reshape(-1)
    pos = rng.integers(0, flat.size, size=flat.size // 2000)
    flat[pos] = rng.choice(np.frombuffer(b"NnacgtU", dtype=np.uint8), size=len(pos))
    low = rng.choice(n, size=50, replace=False)
    t[low] |= 0x20
    a = (n // 3) * L
    flat[a:a + (6 << 20)] = ord("N")
    off = (np.arange(n, dtype=np.uint64) * L).astype(np.uint32)
    ln = np.full(n, L, dtype=np.uint32)
    return np.ascontiguousarray(flat), off, ln


def _run(monkeypatch, threads, text, off, ln, diagnostics):
    monkeypatch.setenv("VFB_HOST_PACK", str(threads))
    ht = torch.from_numpy(text).pin_memory()
    sp = np.zeros(len(off), dtype=api.SPAN_DTYPE)
    sp["off"], sp["len"] = off, ln
    hs = torch.from_numpy(sp.view(np.uint32).reshape(-1, 2).copy()).pin_memory()
    with api.Context((PRE, SUF), diagnostics=diagnostics) as ctx:
        ctx.submit_host_ptr(ht.data_ptr(), ht.numel(), hs.data_ptr(), len(off))
        diag = ctx.diag(len(off)) if diagnostics else None
        table = ctx.finish_dict()
        stats = ctx.stats()
    return table, diag, stats


def test_packed_copy_equals_plain_copy_and_oracle(monkeypatch):
    n = 120000                                                     # 30 MB of text: 8 blocks
    text, off, ln = _reads(n)
    plain, pdiag, pstats = _run(monkeypatch, 0, text, off, ln, True)
    packed, kdiag, kstats = _run(monkeypatch, 4, text, off, ln, True)
    for f in pdiag.dtype.names:
        assert (pdiag[f] == kdiag[f]).all(), f
    assert packed == plain and len(plain) > 1000
    # fewer bytes crossed the link, but not a quarter: the block of N's and whatever the copy engine took stay raw
    assert kstats["h2d_bytes"] < pstats["h2d_bytes"]
    assert pstats["h2d_bytes"] == text.size + 8 * n
    # oracle on a sample (the plain path is checked against it everywhere else)
    m = 6000
    want, odiag, _ = oracle.process_reads(oracle.make_params((PRE, SUF)), text[:m * 250], off[:m], ln[:m], n_threads=4,
                                          want_diag=True)
    for f in ("exact_prefix", "exact_suffix", "score_prefix", "len_prefix", "start"):
        assert (kdiag[f][:m] == odiag[f]).all(), f


def test_packed_copy_many_threads_and_repeated_batches(monkeypatch):
    # more threads than blocks, several submissions into one context (staging buffers and ring slots are reused)
    n = 100000
    text, off, ln = _reads(n, seed=9)
    monkeypatch.setenv("VFB_HOST_PACK", "0")
    ht = torch.from_numpy(text).pin_memory()
    sp = np.zeros(n, dtype=api.SPAN_DTYPE)
    sp["off"], sp["len"] = off, ln
    hs = torch.from_numpy(sp.view(np.uint32).reshape(-1, 2).copy()).pin_memory()
    tables = []
    for threads in (0, 12):
        monkeypatch.setenv("VFB_HOST_PACK", str(threads))
        with api.Context((PRE, SUF)) as ctx:
            for _ in range(5):
                ctx.submit_host_ptr(ht.data_ptr(), ht.numel(), hs.data_ptr(), n)
            tables.append(ctx.finish_dict())
    assert tables[0] == tables[1]
    assert sum(tables[0].values()) % 5 == 0
