"""Writes the synthetic read stream (vfind_b200.api.synth_host) as a FASTQ file, block-gzipped (BGZF members of
<= 65280 bytes of text, zlib level 1, like bgzip / sequencer pipelines) and optionally as one plain gzip stream.
Used by bench.py's ingest leg and the ingest probes; not part of the product."""
import os
import struct
import subprocess
import sys
import zlib
from concurrent.futures import ProcessPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def bgzf_member(data: bytes) -> bytes:
    co = zlib.compressobj(1, zlib.DEFLATED, -15)
    body = co.compress(data) + co.flush()
    hdr = struct.pack("<BBBBIBBH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(body) + 8 - 1)
    return hdr + body + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))


def _bgzf_span(args):
    p, lo, hi = args
    with open(p, "rb") as f:
        f.seek(lo)
        data = f.read(hi - lo)
    return b"".join(bgzf_member(data[i:i + 65280]) for i in range(0, len(data), 65280))


def write_fastq(path, cfg, first, n, api):
    """Plain-text FASTQ of reads [first, first+n): '@r<idx>' headers, quality 'F'."""
    text, _ = api.synth_host(cfg, first, n)
    L = cfg.read_len
    rec = np.empty((n, 2 * L + 17), dtype=np.uint8)
    rec[:, :13] = np.frombuffer(b"@r0000000000\n", np.uint8)
    idx = np.arange(first, first + n)
    for d in range(10):
        rec[:, 11 - d] = 48 + (idx // 10 ** d) % 10
    rec[:, 13:13 + L] = text.reshape(n, L)
    rec[:, 13 + L:16 + L] = np.frombuffer(b"\n+\n", np.uint8)
    rec[:, 16 + L:16 + 2 * L] = ord("F")
    rec[:, 16 + 2 * L] = 10
    rec.tofile(path)
    return rec.size


def write_bgzf_fastq(path_bgzf, cfg, n, api, keep_text=None, plain_gzip=None, procs=None):
    """BGZF file of n reads at path_bgzf; returns (text_bytes, compressed_bytes)."""
    txt = keep_text or path_bgzf + ".txt"
    size = write_fastq(txt, cfg, 0, n, api)
    step = 65280 * 256
    with ProcessPoolExecutor(procs or os.cpu_count()) as ex, open(path_bgzf, "wb") as out:
        for blob in ex.map(_bgzf_span, [(txt, lo, min(size, lo + step)) for lo in range(0, size, step)]):
            out.write(blob)
        out.write(bgzf_member(b""))
    if plain_gzip:
        with open(plain_gzip, "wb") as g:
            subprocess.check_call(["gzip", "-1", "-c", txt], stdout=g)
    if not keep_text:
        os.remove(txt)
    return size, os.path.getsize(path_bgzf)
