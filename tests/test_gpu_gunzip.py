"""Single-stream gzip decoded on the device (kernels_inflate.cu k_gz_* + gunzip_gpu.cu) through vfb_debug_inflate_file with
VFB_GPU_GUNZIP=2 (the device decoder whatever the file size): its output must equal zlib's for every way a gzip stream can
be laid out, for any segment size / chunk gap / chunk capacity (small values force the cross-segment chain, dropped
candidates, and the hand-over to zlib in the middle of a stream), and malformed streams must fail as zlib's do."""
import ctypes
import gzip
import os
import random
import zlib

import numpy as np
import pytest

from test_pgunzip_cpu import fastq, member, normalised

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from vfind_b200 import api
    L = api.load_library()
    L.vfb_debug_inflate_file.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint64,
                                         ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64),
                                         ctypes.POINTER(ctypes.c_uint64)]
    return L


def gpu_inflate(lib, path, cap, segment=None, gap=None, chunk_cap=None, quota=None):
    env = {"VFB_GPU_GUNZIP": "2", "VFB_GUNZIP_SEGMENT": segment, "VFB_GUNZIP_GAP": gap, "VFB_GUNZIP_CAP": chunk_cap,
           "VFB_GUNZIP_HOST_QUOTA": quota}
    old = {k: os.environ.get(k) for k in env}
    for k, v in env.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    try:
        nb, nl, nc = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
        out = np.zeros(cap + 64, dtype=np.uint8)
        rc = lib.vfb_debug_inflate_file(os.fsencode(path), 4, out.ctypes.data, cap + 64, ctypes.byref(nb),
                                        ctypes.byref(nl), ctypes.byref(nc))
        if rc != 0:
            raise RuntimeError(lib.vfb_last_error().decode())
        text = out[:nb.value].tobytes()
        assert nl.value == text.count(b"\n")                # the device's newline counts are what frames the records
        return text
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("layout", ["level1", "level6", "level9", "stored", "fixed", "huffman_only", "rle", "sync_flushes",
                                    "two_members", "many_members", "named_header"])
def test_layouts_segments_gaps(lib, tmp_path, layout):
    rng = random.Random(hash(layout) & 0xffff)
    text = fastq(40000, rng)
    blob = {
        "level1": lambda: member(text, level=1),
        "level6": lambda: member(text, level=6),
        "level9": lambda: member(text, level=9),
        "stored": lambda: member(text, level=0),
        "fixed": lambda: member(text, strategy=zlib.Z_FIXED),
        "huffman_only": lambda: member(text, strategy=zlib.Z_HUFFMAN_ONLY),
        "rle": lambda: member(text, strategy=zlib.Z_RLE),
        "sync_flushes": lambda: member(text, flush_every=70001),
        "two_members": lambda: member(text[:len(text) // 3]) + member(text[len(text) // 3:], level=1),
        "many_members": lambda: b"".join(member(text[i:i + 300011]) for i in range(0, len(text), 300011)),
        "named_header": lambda: gzip.compress(text, 6),
    }[layout]()
    assert gzip.decompress(blob) == text
    p = tmp_path / "x.fq.gz"
    p.write_bytes(blob)
    want = normalised(text)
    # (small chunk capacities make chunks fail: zlib takes the stream over and — with a small quota — hands it back)
    for segment, gap, cc, quota in ((None, None, None, None), (262144, 4096, None, None), (100000, 1024, None, None),
                                    (None, 2048, 40000, None), (65536, 512, 70000, 100000), (300000, 4096, 30000, 1)):
        assert gpu_inflate(lib, p, len(text), segment, gap, cc, quota) == want, (layout, segment, gap, cc, quota)


def test_binary_and_long_distance_content(lib, tmp_path):
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
    data = b"@h\n" + b"".join(parts).replace(b"\n", b" ").replace(b"\r", b" ") + b"\n+\nq\n"
    p = tmp_path / "b.gz"
    for level in (1, 6, 9):
        p.write_bytes(member(data, level=level))
        for segment, gap in ((None, None), (200000, 2048), (70000, 1024)):
            assert gpu_inflate(lib, p, len(data), segment, gap) == data


def test_small_and_empty_streams(lib, tmp_path):
    p = tmp_path / "s.gz"
    for data in (b"", b"@r\nA\n+\nF\n", b"@r\nACGT\n+\nFFFF\n" * 3):
        for blob in (member(data), member(data, level=0), member(b"") + member(data) + member(b""), gzip.compress(data)):
            p.write_bytes(blob)
            assert gpu_inflate(lib, p, len(data) + 16, 4096) == normalised(data)


def test_malformed_streams_fail_like_zlib(lib, tmp_path):
    rng = random.Random(11)
    text = fastq(20000, rng)
    good = member(text)
    p = tmp_path / "m.gz"
    cases = {
        "truncated_data": good[:len(good) // 2],
        "truncated_trailer": good[:-3],
        "bad_crc": good[:-8] + bytes([good[-8] ^ 1]) + good[-7:],
        "bad_isize": good[:-1] + bytes([good[-1] ^ 1]),
        "garbage_after": good + b"garbage that is not a gzip header",
        "flipped_bit_1": good[:5000] + bytes([good[5000] ^ 0x10]) + good[5001:],
        "flipped_bit_2": good[:len(good) - 2000] + bytes([good[len(good) - 2000] ^ 0x04]) + good[len(good) - 1999:],
    }
    for name, blob in cases.items():
        p.write_bytes(blob)
        for segment, gap in ((None, None), (150000, 2048)):
            with pytest.raises(RuntimeError):
                gpu_inflate(lib, p, len(text), segment, gap)


def test_find_variants_on_a_plain_gzip_file_uses_the_device_decoder(tmp_path):
    """The user-facing call on a single-stream .gz: same table as the oracle."""
    import oracle
    from vfind_b200 import find_variants
    cfg = oracle.synth_cfg(seed=5, read_len=150, adapter_len=20, region_len=99, n_variants=500, p_err=0.2)
    ad = tuple(a.decode() for a in oracle.synth_adapters(cfg))
    path = str(tmp_path / "plain.fq")
    oracle.write_fastq(cfg, 0, 60000, path, bgzf=False)
    raw = open(path, "rb").read()
    gz = path + ".gz"
    with open(gz, "wb") as f:
        f.write(member(raw, level=1))
    want = oracle.find_variants_file(gz, ad, n_threads=4)
    os.environ["VFB_GPU_GUNZIP"] = "2"
    try:
        out = find_variants(gz, ad, show_progress=False, devices=[0])
    finally:
        os.environ.pop("VFB_GPU_GUNZIP", None)
    cols = out.to_pydict() if hasattr(out, "to_pydict") else out.to_dict(as_series=False)
    assert {k.encode(): v for k, v in zip(cols["sequence"], cols["count"])} == want
