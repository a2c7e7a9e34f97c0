"""CPU tests of the parallel single-stream gunzip (vfind_b200/csrc/pgunzip.cu) through vfb_debug_inflate_file:
its output must equal zlib's for every way a gzip stream can be laid out, for any thread count and segment
size, and malformed streams must fail as the serial zlib path does."""
import ctypes
import gzip
import os
import random
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


def fastq(n, rng, qual="FFFFF:,#"):
    recs = []
    for i in range(n):
        L = rng.randrange(60, 260)
        recs.append("@read%d lane=%d\n%s\n+\n%s\n" % (i, i % 4, "".join(rng.choice("ACGT") for _ in range(L)),
                                                      "".join(rng.choice(qual) for _ in range(L))))
    return "".join(recs).encode()


def inflate(lib, path, threads, cap, segment=None, pgunzip=True):
    env = {"VFB_PGUNZIP_SEGMENT": str(segment) if segment else None, "VFB_PGUNZIP": None if pgunzip else "0"}
    old = {k: os.environ.get(k) for k in env}
    for k, v in env.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    try:
        nb, nl, nc = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
        out = np.zeros(cap + 64, dtype=np.uint8)
        rc = lib.vfb_debug_inflate_file(os.fsencode(path), threads, out.ctypes.data, cap + 64, ctypes.byref(nb),
                                        ctypes.byref(nl), ctypes.byref(nc))
        if rc != 0:
            raise RuntimeError(lib.vfb_last_error().decode())
        return out[:nb.value].tobytes()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def deflate_raw(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, flush_every=None):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    if not flush_every:
        return co.compress(data) + co.flush()
    out = []
    for i in range(0, len(data), flush_every):
        out.append(co.compress(data[i:i + flush_every]))
        out.append(co.flush(zlib.Z_FULL_FLUSH if (i // flush_every) % 2 else zlib.Z_SYNC_FLUSH))
    return b"".join(out) + co.flush()


def member(data, **kw):
    import struct
    return (b"\x1f\x8b\x08\x00\0\0\0\0\0\xff" + deflate_raw(data, **kw) +
            struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data) & 0xffffffff))


def normalised(text):
    t = text.rstrip(b"\r\n")
    return t + b"\n" if t else b""


@pytest.mark.parametrize("layout", ["level1", "level6", "level9", "stored", "fixed", "huffman_only", "rle", "sync_flushes",
                                    "two_members", "many_members", "named_header"])
def test_layouts_threads_and_segments(lib, tmp_path, layout):
    rng = random.Random(hash(layout) & 0xffff)
    text = fastq(12000, rng)
    blob = {
        "level1": lambda: member(text, level=1),
        "level6": lambda: member(text, level=6),
        "level9": lambda: member(text, level=9),
        "stored": lambda: member(text, level=0),
        "fixed": lambda: member(text, strategy=zlib.Z_FIXED),
        "huffman_only": lambda: member(text, strategy=zlib.Z_HUFFMAN_ONLY),
        "rle": lambda: member(text, strategy=zlib.Z_RLE),
        "sync_flushes": lambda: member(text, flush_every=70001),          # empty stored blocks between the blocks
        "two_members": lambda: member(text[:len(text) // 3]) + member(text[len(text) // 3:], level=1),
        "many_members": lambda: b"".join(member(text[i:i + 30011]) for i in range(0, len(text), 30011)),
        "named_header": lambda: gzip.compress(text, 6),                    # FNAME-less but MTIME set; python writer
    }[layout]()
    assert gzip.decompress(blob) == text
    p = tmp_path / "x.fq.gz"
    p.write_bytes(blob)
    want = normalised(text)
    for threads in (2, 3, 8):
        for segment in (None, 40000, 333333):
            assert inflate(lib, p, threads, len(text), segment) == want, (layout, threads, segment)
    assert inflate(lib, p, 1, len(text)) == want                          # zlib stream
    assert inflate(lib, p, 4, len(text), pgunzip=False) == want           # switched off


def test_binary_and_long_distance_content(lib, tmp_path):
    # incompressible stretches (stored blocks inside a level-6 stream), long repeats at distance ~32 KiB, runs
    rng = random.Random(9)
    block = bytes(rng.randrange(256) for _ in range(31000))
    parts = []
    for i in range(120):
        k = rng.randrange(4)
        if k == 0:
            parts.append(bytes(rng.randrange(256) for _ in range(rng.randrange(1, 70000))))
        elif k == 1:
            parts.append(block[:rng.randrange(1, 31000)])
        elif k == 2:
            parts.append(bytes([rng.randrange(256)]) * rng.randrange(1, 100000))
        else:
            parts.append(fastq(200, rng))
    # four lines in all: the harness cuts at 4-line boundaries and checks nothing else
    data = b"@h\n" + b"".join(parts).replace(b"\n", b" ").replace(b"\r", b" ") + b"\n+\nq\n"
    p = tmp_path / "b.gz"
    for level in (1, 6, 9):
        p.write_bytes(member(data, level=level))
        for threads, segment in ((2, 50000), (5, 200000), (8, None)):
            assert inflate(lib, p, threads, len(data), segment) == data


def test_small_and_empty_streams(lib, tmp_path):
    p = tmp_path / "s.gz"
    for data in (b"", b"@r\nA\n+\nF\n", b"@r\nACGT\n+\nFFFF\n" * 3):
        for blob in (member(data), member(data, level=0), member(b"") + member(data) + member(b""), gzip.compress(data)):
            p.write_bytes(blob)
            for threads in (2, 8):
                assert inflate(lib, p, threads, len(data) + 16, 4096) == normalised(data)
    p.write_bytes(b"")
    assert inflate(lib, p, 4, 16) == b""


def test_malformed_streams_fail_like_zlib(lib, tmp_path):
    rng = random.Random(11)
    text = fastq(6000, rng)
    good = member(text)
    p = tmp_path / "m.gz"
    cases = {
        "truncated_data": good[:len(good) // 2],
        "truncated_trailer": good[:-3],
        "bad_crc": good[:-8] + bytes([good[-8] ^ 1]) + good[-7:],
        "bad_isize": good[:-1] + bytes([good[-1] ^ 1]),
        "garbage_after": good + b"garbage that is not a gzip header",
        "bad_magic": b"\x1f\x8c" + good[2:],
        "flipped_bit_1": good[:5000] + bytes([good[5000] ^ 0x10]) + good[5001:],
        "flipped_bit_2": good[:len(good) - 2000] + bytes([good[len(good) - 2000] ^ 0x04]) + good[len(good) - 1999:],
    }
    for name, blob in cases.items():
        p.write_bytes(blob)
        for threads, segment in ((1, None), (4, None), (4, 50000), (8, 20000)):
            with pytest.raises(RuntimeError):
                inflate(lib, p, threads, len(text), segment)


def _piecewise_member(data, rng):
    """ONE gzip member whose deflate stream is a chain of independently compressed pieces (the way pigz writes): every piece
    has its own level / strategy and ends byte aligned with a sync or full flush, the last one with the final block."""
    import struct
    out = [b"\x1f\x8b\x08\x00\0\0\0\0\0\xff"]
    pos = 0
    while True:
        n = rng.choice([1, 7, 300, 5000, 70000, 200000])
        n = min(n + rng.randrange(n), len(data) - pos)
        last = pos + n >= len(data)
        co = zlib.compressobj(rng.randrange(0, 10), zlib.DEFLATED, -15, rng.choice([1, 8, 9]),
                              rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]))
        out.append(co.compress(data[pos:pos + n]))
        out.append(co.flush(zlib.Z_FINISH) if last else co.flush(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH])))
        pos += n
        if last:
            break
    out.append(struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data) & 0xffffffff))
    return b"".join(out)


@pytest.mark.parametrize("seed", range(12))
def test_random_layouts_against_zlib(lib, tmp_path, seed):
    """Seeded fuzz: random content (reads, incompressible stretches, long runs, repeats at window distance), random piece
    sizes / levels / strategies / flush kinds / member boundaries, random thread counts and segment sizes."""
    rng = random.Random(1000 + seed)
    if seed % 3 == 0:
        text = fastq(rng.randrange(1, 9000), rng)
    else:
        block = bytes(rng.randrange(256) for _ in range(33000))
        parts = []
        for _ in range(rng.randrange(5, 60)):
            k = rng.randrange(5)
            if k == 0:
                parts.append(bytes(rng.randrange(256) for _ in range(rng.randrange(1, 50000))))
            elif k == 1:
                parts.append(block[rng.randrange(0, 2000):rng.randrange(2000, 33000)])
            elif k == 2:
                parts.append(bytes([rng.randrange(256)]) * rng.randrange(1, 90000))
            elif k == 3:
                parts.append(bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(1, 40000))))
            else:
                parts.append(fastq(rng.randrange(1, 300), rng))
        text = b"@h\n" + b"".join(parts).replace(b"\n", b" ").replace(b"\r", b" ") + b"\n+\nq\n"      # one giant record
    cuts = sorted(rng.sample(range(1, len(text)), min(rng.choice([0, 0, 1, 3]), len(text) - 1))) if len(text) > 1 else []
    blob = b"".join(_piecewise_member(text[a:b], rng) for a, b in zip([0] + cuts, cuts + [len(text)]))
    assert gzip.decompress(blob) == text
    p = tmp_path / "f.gz"
    p.write_bytes(blob)
    want = normalised(text) if seed % 3 == 0 else text
    for _ in range(4):
        threads, segment = rng.randrange(2, 9), rng.choice([None, 20000, 65536, 150000, 400000])
        assert inflate(lib, p, threads, len(text), segment) == want, (seed, threads, segment)
