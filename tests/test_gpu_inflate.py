"""GPU inflater (kernels_inflate.cu) against zlib: every block type, compression levels,
random / repetitive / FASTQ payloads, corrupt members (CRC-32, ISIZE, deflate data)."""
import ctypes
import gzip
import os
import random
import struct
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def gz_member(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, bgzf=True) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    body = co.compress(data) + co.flush()
    if bgzf:
        hdr = struct.pack("<BBBBIBBH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6) + b"BC" + struct.pack("<HH", 2, (12 + 6 + len(body) + 8 - 1) & 0xffff)
    else:
        hdr = struct.pack("<BBBBIBB", 0x1f, 0x8b, 8, 8, 0, 0, 0xff) + b"name.fq\0"
    return hdr + body + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data) & 0xffffffff)


def gpu_inflate(members, payload_sizes):
    from vfind_b200 import api
    L = api.load_library()
    L.vfb_debug_gpu_inflate.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint32,
                                        ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_double)]
    z = np.frombuffer(b"".join(members), dtype=np.uint8)
    tab = np.zeros((len(members), 4), dtype=np.uint32)
    zo = oo = 0
    for i, (m, n) in enumerate(zip(members, payload_sizes)):
        tab[i] = (zo, len(m), oo, n)
        zo += len(m)
        oo += n
    out = np.zeros(max(oo, 1), dtype=np.uint8)
    bad = ctypes.c_uint32(0)
    ms = ctypes.c_double(0)
    rc = L.vfb_debug_gpu_inflate(z.ctypes.data, z.size, tab.ctypes.data, len(members), out.ctypes.data, oo, -1,
                                 ctypes.byref(bad), ctypes.byref(ms))
    assert rc == 0, L.vfb_last_error()
    return out[:oo].tobytes(), bad.value, ms.value


def payloads(rng):
    fq = []
    for i in range(150):
        L = rng.randrange(50, 300)
        fq.append("@r%d\n%s\n+\n%s\n" % (i, "".join(rng.choice("ACGT") for _ in range(L)), "F" * L))
    fq = "".join(fq).encode()
    return [b"", b"A", b"hello hello hello hello", bytes(range(256)) * 4, b"\0" * 65000, fq[:65000],
            bytes(rng.randrange(256) for _ in range(30000)),                      # incompressible -> stored blocks
            bytes(rng.choice(b"AC") for _ in range(65535)),
            (b"ACGT" * 7 + b"N") * 2000, fq[:1000], b"x" * 258 + b"y" * 259 + b"x" * 300]


def test_all_block_types_and_levels():
    rng = random.Random(31)
    members, sizes, want = [], [], []
    for data in payloads(rng):
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                for b in (True, False):
                    m = gz_member(data, level, strategy, bgzf=b)
                    assert gzip.decompress(m) == data
                    members.append(m); sizes.append(len(data)); want.append(data)
    got, bad, ms = gpu_inflate(members, sizes)
    assert bad == 0xFFFFFFFF
    assert got == b"".join(want)


def test_corrupt_members_are_reported():
    rng = random.Random(32)
    data = payloads(rng)[5]
    good = gz_member(data, 6)
    # CRC, ISIZE, a flipped bit in the deflate stream, truncated stream
    crc_bad = bytearray(good); crc_bad[-8] ^= 1
    isize_bad = bytearray(good); isize_bad[-1] ^= 1
    for mutate_at in (30, len(good) // 2, len(good) - 12):
        flip = bytearray(good); flip[mutate_at] ^= 0x10
        got, bad, _ = gpu_inflate([good, bytes(flip), good], [len(data)] * 3)
        assert bad == 1, mutate_at
        assert got[:len(data)] == data and got[2 * len(data):] == data     # neighbours are untouched
    assert gpu_inflate([good, bytes(crc_bad)], [len(data)] * 2)[1] == 1
    assert gpu_inflate([bytes(isize_bad), good], [len(data)] * 2)[1] == 0
    trunc = good[:len(good) // 2] + good[-8:]
    assert gpu_inflate([good, trunc], [len(data)] * 2)[1] == 1
    rnd = bytes(rng.randrange(256) for _ in range(500))
    assert gpu_inflate([good, good[:18] + rnd + good[-8:]], [len(data)] * 2)[1] == 1


def test_many_members_throughput():
    rng = random.Random(33)
    fq = []
    for i in range(260):
        L = 250
        fq.append("@r%09d\n%s\n+\n%s\n" % (i, "".join(rng.choice("ACGT") for _ in range(L)), "F" * L))
    block = "".join(fq).encode()[:65280]
    m = gz_member(block, 1)
    n = 4000
    got, bad, ms = gpu_inflate([m] * n, [len(block)] * n)
    assert bad == 0xFFFFFFFF and got == block * n
    print("\nGPU inflate: %d members, %.1f MB text in %.2f ms = %.1f GB/s" % (n, n * len(block) / 1e6, ms, n * len(block) / ms / 1e6))


def bgzf_file(data: bytes, block=30000, eof=True, level=6) -> bytes:
    out = b"".join(gz_member(data[i:i + block], level) for i in range(0, len(data), block))
    return out + (gz_member(b"") if eof else b"")


def _rows(out):
    cols = out.to_pydict() if hasattr(out, "to_pydict") else out.to_dict(as_series=False)
    return {k.encode(): v for k, v in zip(cols["sequence"], cols["count"])}


def test_find_variants_on_block_gzip_files(tmp_path):
    import oracle
    from test_gpu_parity import PREFIX, SUFFIX, make_reads
    from vfind_b200 import PanicException, find_variants
    rng = random.Random(34)
    seqs = make_reads(rng, PREFIX, SUFFIX, 20000, lib=400)
    ad = (PREFIX.decode(), SUFFIX.decode())
    text = "".join("@r%d\n%s\n+\n%s\n" % (i, s.decode(), "F" * len(s)) for i, s in enumerate(seqs)).encode()
    ref = tmp_path / "ref.fq.gz"
    ref.write_bytes(gzip.compress(text))
    want = oracle.find_variants_file(str(ref), (PREFIX, SUFFIX), n_threads=8)
    half = text.rfind(b"\n@r", 0, len(text) // 2) + 1
    variants = {
        "bgzf": bgzf_file(text),
        "bgzf_small_blocks_noeof": bgzf_file(text, block=777, eof=False),
        "bgzf_then_gzip": bgzf_file(text[:half], eof=False) + gzip.compress(text[half:]),
        "gzip_then_bgzf": gzip.compress(text[:half]) + bgzf_file(text[half:]),
        "bgzf_crlf_nofinal": bgzf_file(text.replace(b"\n", b"\r\n")[:-2]),
        "bgzf_trailing_blank": bgzf_file(text + b"\n\n\r\n"),
        "bgzf_split_mid_record": bgzf_file(text[:half + 7], eof=False) + bgzf_file(text[half + 7:]),
    }
    for name, blob in variants.items():
        p = tmp_path / (name + ".fq.gz")
        p.write_bytes(blob)
        for chunk in (None, "5000", "200000"):
            for gpu in ("1", "0"):
                if gpu == "0" and chunk == "5000":
                    continue        # the host-thread path needs a chunk that holds a whole member
                os.environ["VFB_GPU_INFLATE"] = gpu
                if chunk:
                    os.environ["VFB_INGEST_CHUNK"] = chunk
                try:
                    got = _rows(find_variants(str(p), ad, n_threads=4))
                finally:
                    os.environ.pop("VFB_GPU_INFLATE", None)
                    os.environ.pop("VFB_INGEST_CHUNK", None)
                assert got == want, (name, chunk, gpu)
    # errors: corrupt member, truncated record at the end, malformed record
    bad = bytearray(variants["bgzf"]); bad[len(bad) // 3] ^= 0x21
    p = tmp_path / "bad.fq.gz"
    p.write_bytes(bytes(bad))
    with pytest.raises(PanicException):
        find_variants(str(p), ad)
    # a file cut inside a member, and inside a member header
    whole = variants["bgzf"]
    second = whole.find(b"\x1f\x8b\x08\x04", 100)
    for cut in (len(whole) - 40, second + 9, second + 300):
        p.write_bytes(whole[:cut])
        for chunk in (None, "200000"):
            if chunk:
                os.environ["VFB_INGEST_CHUNK"] = chunk
            try:
                with pytest.raises(PanicException, match="truncated|invalid"):
                    find_variants(str(p), ad)
            finally:
                os.environ.pop("VFB_INGEST_CHUNK", None)
    p.write_bytes(bgzf_file(text + b"@q\nACGT\n"))
    with pytest.raises(PanicException, match="truncated"):
        find_variants(str(p), ad)
    p.write_bytes(bgzf_file(text + b"@q\nACGT\n+\nFF\n"))
    with pytest.raises(PanicException, match="record 20000"):
        find_variants(str(p), ad)
