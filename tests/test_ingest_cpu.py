"""CPU tests of the ingest front half (vfb_debug_inflate_file): serial gzip, multi-member gzip,
member-parallel BGZF, mixed files, chunk cutting at 4-line boundaries and carry-over.
No GPU needed: this is host logic of vfind_b200/csrc/ingest.cu."""
import ctypes
import gzip
import os
import random
import struct
import zlib

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib():
    from vfind_b200 import build
    build.build()
    from vfind_b200 import api
    L = api.load_library()
    L.vfb_debug_inflate_file.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint64,
                                         ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64),
                                         ctypes.POINTER(ctypes.c_uint64)]
    return L


def bgzf_member(data: bytes) -> bytes:
    """One BGZF block (SAM spec 4.1): gzip member with a 'BC' extra field holding BSIZE."""
    assert len(data) < 65536
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = co.compress(data) + co.flush()
    bsize = 12 + 6 + len(body) + 8 - 1
    hdr = struct.pack("<BBBBIBBH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6) + b"BC" + struct.pack("<HH", 2, bsize)
    return hdr + body + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))


def bgzf(data: bytes, block=40000, eof=True) -> bytes:
    out = b"".join(bgzf_member(data[i:i + block]) for i in range(0, len(data), block))
    return out + (bgzf_member(b"") if eof else b"")


def fastq_text(n, rng, crlf=False):
    nl = "\r\n" if crlf else "\n"
    recs = []
    for i in range(n):
        L = rng.randrange(1, 180)
        seq = "".join(rng.choice("ACGTN") for _ in range(L))
        recs.append("@read%d some description%s%s%s+%s%s%s" % (i, nl, seq, nl, nl, "F" * L, nl))
    return "".join(recs).encode()


def run(lib, path, threads, chunk=None):
    from vfind_b200 import api
    if chunk is not None:
        os.environ["VFB_INGEST_CHUNK"] = str(chunk)
    try:
        nb, nl, nc = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
        cap = 1 << 26
        out = np.zeros(cap, dtype=np.uint8)
        rc = lib.vfb_debug_inflate_file(os.fsencode(path), threads, out.ctypes.data, cap,
                                        ctypes.byref(nb), ctypes.byref(nl), ctypes.byref(nc))
        if rc != 0:
            raise RuntimeError("%d: %s" % (rc, lib.vfb_last_error().decode()))
        return out[:nb.value].tobytes(), nl.value, nc.value
    finally:
        if chunk is not None:
            del os.environ["VFB_INGEST_CHUNK"]


def normalised(text: bytes) -> bytes:
    t = text.rstrip(b"\r\n")
    return t + b"\n" if t else b""


@pytest.mark.parametrize("kind", ["gzip", "multi", "bgzf", "bgzf_noeof", "bgzf_then_gzip", "gzip_then_bgzf"])
@pytest.mark.parametrize("threads", [1, 4])
def test_inflate_paths_agree_with_zlib(lib, tmp_path, kind, threads):
    rng = random.Random(hash(kind) & 0xffff)
    text = fastq_text(3000, rng, crlf=(kind == "multi"))
    half = len(text) // 2
    half = text.rfind(b"\n", 0, half) + 1
    blob = {
        "gzip": gzip.compress(text),
        "multi": b"".join(gzip.compress(text[i:i + 50000]) for i in range(0, len(text), 50000)),
        "bgzf": bgzf(text),
        "bgzf_noeof": bgzf(text, block=65000, eof=False),
        "bgzf_then_gzip": bgzf(text[:half], eof=False) + gzip.compress(text[half:]),
        "gzip_then_bgzf": gzip.compress(text[:half]) + bgzf(text[half:]),
    }[kind]
    assert gzip.decompress(blob) == text
    p = tmp_path / "x.fq.gz"
    p.write_bytes(blob)
    want = normalised(text)
    for chunk in (None, 1 << 20, 150000, 70000):
        got, lines, chunks = run(lib, p, threads, chunk)
        assert got == want, (kind, chunk)
        assert lines == want.count(b"\n") and lines % 4 == 0
        if chunk == 70000:
            assert chunks > 3


def test_chunks_end_on_record_boundaries(lib, tmp_path):
    rng = random.Random(5)
    text = fastq_text(2000, rng)
    p = tmp_path / "b.fq.gz"
    p.write_bytes(bgzf(text, block=3000))
    # the debug entry concatenates chunks; re-run with a chunk small enough to force many cuts
    got, lines, chunks = run(lib, p, 3, 5000)
    assert got == normalised(text) and chunks > 50


def test_eof_rules(lib, tmp_path):
    rec = b"@r\nACGT\n+\nFFFF"
    for tail, ok in ((b"", True), (b"\n", True), (b"\n\n\n", True), (b"\r\n", True), (b"\n@q\nAC", False),
                     (b"\n@q\nAC\n+\n", False)):
        for wrap in (gzip.compress, bgzf):
            p = tmp_path / "e.fq.gz"
            p.write_bytes(wrap(rec + tail))
            if ok:
                got, lines, _ = run(lib, p, 2)
                assert got == rec + b"\n" and lines == 4
            else:
                with pytest.raises(RuntimeError, match="truncated"):
                    run(lib, p, 2)
    p = tmp_path / "empty.fq.gz"
    p.write_bytes(gzip.compress(b""))
    assert run(lib, p, 1) == (b"", 0, 1)
    p.write_bytes(b"")
    assert run(lib, p, 1) == (b"", 0, 1)


def test_corrupt_inputs(lib, tmp_path):
    rng = random.Random(6)
    text = fastq_text(500, rng)
    good = bgzf(text)
    bad = bytearray(good)
    bad[len(bad) // 2] ^= 0x55
    p = tmp_path / "c.fq.gz"
    p.write_bytes(bytes(bad))
    with pytest.raises(RuntimeError):
        run(lib, p, 4)
    p.write_bytes(good[:len(good) // 2])
    with pytest.raises(RuntimeError):
        run(lib, p, 4)
    p.write_bytes(gzip.compress(text)[:-20])
    with pytest.raises(RuntimeError):
        run(lib, p, 1)
    p.write_bytes(b"@r\nACGT\n+\nFFFF\n")          # plain text is not gzip (Q11)
    with pytest.raises(RuntimeError):
        run(lib, p, 1)
    with pytest.raises(RuntimeError, match="cannot open"):
        run(lib, tmp_path / "missing.gz", 1)


@pytest.mark.parametrize("seed", range(10))
def test_random_member_mixes_and_chunk_sizes(lib, tmp_path, seed):
    """Seeded fuzz of the front half: runs of block-gzip members of random sizes (some empty), plain gzip members in
    between, records that straddle every kind of boundary, CR LF or LF, trailing blank lines, random chunk sizes and
    thread counts.  The submitted text is the file's text, cut into chunks of whole records."""
    rng = random.Random(400 + seed)
    text = fastq_text(rng.randrange(1, 2500), rng, crlf=(seed % 4 == 1))
    blob, pos, max_blk = [], 0, 1
    while pos < len(text):
        n = min(len(text) - pos, rng.choice([1, 100, 3000, 40000, 65000, 200000]))
        piece = text[pos:pos + n]
        pos += n
        if rng.random() < 0.6:
            blk = rng.choice([1 if n <= 1000 else 2000, 500 if n <= 40000 else 5000, 20000, 65000])
            max_blk = max(max_blk, min(blk, len(piece)))
            blob.append(bgzf(piece, block=blk, eof=rng.random() < 0.3))          # an end-of-file member mid-file is legal
        else:
            blob.append(gzip.compress(piece, rng.randrange(0, 10)))
    tail = rng.choice([b"", b"\n", b"\n\n", b"\r\n\r\n"])
    if tail:
        blob.append(rng.choice([gzip.compress, lambda d: bgzf(d, eof=True)])(tail))
    blob = b"".join(blob)
    assert gzip.decompress(blob) == text + tail
    p = tmp_path / "z.fq.gz"
    p.write_bytes(blob)
    want = normalised(text)
    for _ in range(4):
        # (a chunk has to hold one member's text plus a carried record: the knob is for tests, the default is 128 MB)
        threads, chunk = rng.choice([1, 2, 5, 8]), rng.choice([c for c in (None, 3000, 20000, 70000, 1 << 20) if c is None or c >= max_blk + 2000])
        got, lines, chunks = run(lib, p, threads, chunk)
        assert got == want, (seed, threads, chunk)
        assert lines == want.count(b"\n") and lines % 4 == 0
