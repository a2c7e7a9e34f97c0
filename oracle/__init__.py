"""ctypes binding of the CPU oracle (oracle/vfind_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (vfind_b200) never imports it.

Parity status: see oracle/vfind_oracle.h — exact search, translate, thresholds and the
reference's own fixtures are pinned; the DP's gapped / tie behaviour is "parity unpinned"
(parasail cannot be built or imported in this image).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvfind_oracle.so")

NONE = -1
INT32_MIN = -(2 ** 31)


class DpRules(C.Structure):
    _fields_ = [("gap_tie_open", C.c_int), ("h_priority", C.c_int),
                ("end_rule", C.c_int), ("wildcard_zero", C.c_int)]

    def __init__(self, gap_tie_open=0, h_priority=0, end_rule=0, wildcard_zero=1):
        super().__init__(gap_tie_open, h_priority, end_rule, wildcard_zero)


class Params(C.Structure):
    _fields_ = [("prefix", C.c_char_p), ("prefix_len", C.c_size_t),
                ("suffix", C.c_char_p), ("suffix_len", C.c_size_t),
                ("match_score", C.c_int32), ("mismatch_score", C.c_int32),
                ("gap_open_penalty", C.c_int32), ("gap_extend_penalty", C.c_int32),
                ("accept_prefix_alignment", C.c_double), ("accept_suffix_alignment", C.c_double),
                ("skip_translation", C.c_int32), ("rules", DpRules)]


DIAG_DTYPE = np.dtype([("exact_prefix", "<i4"), ("exact_suffix", "<i4"),
                       ("score_prefix", "<i4"), ("len_prefix", "<i4"),
                       ("score_suffix", "<i4"), ("len_suffix", "<i4"),
                       ("start", "<i4"), ("end", "<i4")])


class SynthCfg(C.Structure):
    """vfb_synth_cfg (include/vfind_b200.h), restated here so that the reference arm never loads the product."""
    _fields_ = [("seed", C.c_uint64), ("read_len", C.c_uint32), ("adapter_len", C.c_uint32),
                ("region_len", C.c_uint32), ("n_variants", C.c_uint32), ("zipf", C.c_uint32),
                ("p_err_ppm", C.c_uint32), ("indel_ppm", C.c_uint32), ("force_indel", C.c_uint32),
                ("frameshift_ppm", C.c_uint32), ("noise_ppm", C.c_uint32), ("n_ppm", C.c_uint32),
                ("reserved", C.c_uint32)]


def synth_cfg(seed=1003, read_len=250, adapter_len=20, region_len=198, n_variants=1000000, zipf=1,
              p_err=0.30, indel=0.5, force_indel=1, frameshift=0.05, noise=0.001, n_rate=1e-4) -> SynthCfg:
    ppm = lambda x: int(round(x * 1e6))
    return SynthCfg(seed, read_len, adapter_len, region_len, n_variants, zipf, ppm(p_err), ppm(indel),
                    force_indel, ppm(frameshift), ppm(noise), ppm(n_rate), 0)


def _cfg(cfg) -> SynthCfg:
    """Accepts this module's SynthCfg or any ctypes struct with the same fields (vfind_b200.api.SynthCfg)."""
    if isinstance(cfg, SynthCfg):
        return cfg
    return SynthCfg(*[getattr(cfg, f) for f, _ in SynthCfg._fields_])


def synth_adapters(cfg):
    c = _cfg(cfg)
    a, b = C.create_string_buffer(64), C.create_string_buffer(64)
    if lib().vfo_synth_adapters(C.byref(c), a, b) != 0:
        raise ValueError("bad synth cfg")
    return a.raw[:c.adapter_len], b.raw[:c.adapter_len]


def synth_reads(cfg, first: int, n: int, threads: int = 0):
    """(text uint8[n*L], off uint32[n], len uint32[n]) of reads [first, first+n): the same bytes as vfb_synth_host."""
    c = _cfg(cfg)
    text = np.zeros(n * c.read_len, dtype=np.uint8)
    off = np.zeros(n, dtype=np.uint32)
    ln = np.zeros(n, dtype=np.uint32)
    if lib().vfo_synth_reads(C.byref(c), first, n, text.ctypes.data, off.ctypes.data, ln.ctypes.data,
                             threads or (os.cpu_count() or 1)) != 0:
        raise ValueError("bad synth cfg / size")
    return text, off, ln


def write_fastq(cfg, first: int, n: int, path: str, bgzf: bool = True, level: int = 1, threads: int = 0, append: bool = False):
    """The stream as a FASTQ file ('@r<index>' headers, quality 'F'); bgzf: block-gzip members of 65280 text bytes.
    Returns (text_bytes, file_bytes)."""
    c = _cfg(cfg)
    fb = C.c_uint64(0)
    tb = lib().vfo_write_fastq(C.byref(c), first, n, os.fsencode(path), 1 if bgzf else 0, level,
                               threads or (os.cpu_count() or 1), 1 if append else 0, C.byref(fb))
    if tb == 0 and n:
        raise OSError("cannot write %s" % path)
    return int(tb), int(fb.value)


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc).  Building the checker is not using it."""
    srcs = [os.path.join(_HERE, "vfind_oracle.c"), os.path.join(_HERE, "sg_stats_simd.c"), os.path.join(_HERE, "synth_host.c"),
            os.path.join(_HERE, "vfind_oracle.h"),
            os.path.join(os.path.dirname(_HERE), "vfind_b200", "csrc", "synth.h")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libvfind_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.vfo_synth_adapters.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.vfo_synth_reads.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.vfo_write_fastq.restype = C.c_uint64
        L.vfo_write_fastq.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_uint64)]
        L.vfo_memmem.restype = C.c_int64
        L.vfo_memmem.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.vfo_threshold_preflight.argtypes = [C.c_double, C.POINTER(C.c_int)]
        L.vfo_min_score.restype = C.c_double
        L.vfo_min_score.argtypes = [C.c_double, C.c_int32, C.c_size_t]
        L.vfo_sg_stats.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.POINTER(DpRules)] + [C.POINTER(C.c_int)] * 4
        L.vfo_translate.restype = C.c_int64
        L.vfo_translate.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.vfo_is_utf8.argtypes = [C.c_char_p, C.c_size_t]
        L.vfo_aa_table.restype = C.c_char_p
        L.vfo_ascii_to_index.restype = C.POINTER(C.c_uint8)
        L.vfo_table_new.restype = C.c_void_p
        L.vfo_table_free.argtypes = [C.c_void_p]
        L.vfo_table_rows.restype = C.c_uint64
        L.vfo_table_rows.argtypes = [C.c_void_p]
        L.vfo_table_key_bytes.restype = C.c_uint64
        L.vfo_table_key_bytes.argtypes = [C.c_void_p]
        L.vfo_table_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.vfo_table_export_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.vfo_process_reads.argtypes = [C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_uint64)]
        L.vfo_process_reads_ex.argtypes = L.vfo_process_reads.argtypes + [C.c_uint]
        L.vfo_sg_stats_x16.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(DpRules)] + [C.POINTER(C.c_int)] * 4
        L.vfo_find_variants_file.argtypes = [C.c_char_p, C.POINTER(Params), C.c_int, C.c_void_p,
                                             C.POINTER(C.c_uint64), C.c_char_p, C.c_size_t]
        L.vfo_find_variants_file_ex.argtypes = [C.c_char_p, C.POINTER(Params), C.c_int, C.c_void_p,
                                                C.POINTER(C.c_uint64), C.c_char_p, C.c_size_t, C.c_uint, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def memmem(hay: bytes, needle: bytes) -> int:
    return int(lib().vfo_memmem(hay, len(hay), needle, len(needle)))


def threshold_preflight(thr: float):
    """Returns skip_alignment (bool); raises ValueError like src/lib.rs:107-109."""
    skip = C.c_int(0)
    if lib().vfo_threshold_preflight(thr, C.byref(skip)) != 0:
        raise ValueError("Accept alignment threshold must be between 0 and 1.")
    return bool(skip.value)


def min_score(thr: float, match_score: int, adapter_len: int) -> float:
    return float(lib().vfo_min_score(thr, match_score, adapter_len))


def sg_stats(adapter: bytes, read: bytes, match=3, mismatch=-2, gap_open=5, gap_extend=2,
             rules: DpRules | None = None):
    """(score, length, end_i, end_j) of the semi-global alignment with statistics."""
    r = rules or DpRules()
    out = [C.c_int(0) for _ in range(4)]
    rc = lib().vfo_sg_stats(adapter, len(adapter), read, len(read), match, mismatch, gap_open,
                            gap_extend, C.byref(r), *[C.byref(o) for o in out])
    if rc != 0:
        raise ValueError("empty adapter or read")
    return tuple(o.value for o in out)


def simd_available() -> bool:
    return bool(lib().vfo_simd_available())


def sg_stats_x16(adapter: bytes, reads, match=3, mismatch=-2, gap_open=5, gap_extend=2, rules: DpRules | None = None):
    """The sixteen-lane kernel of the CPU baseline on up to 16 reads: a list of (score, length, end_i, end_j), or None
    when it declines (no AVX2, values outside its 16-bit lanes) and the caller has to use sg_stats."""
    r = rules or DpRules()
    n = len(reads)
    arr = (C.c_char_p * n)(*reads)
    lens = (C.c_int * n)(*[len(x) for x in reads])
    out = [(C.c_int * n)() for _ in range(4)]
    rc = lib().vfo_sg_stats_x16(adapter, len(adapter), arr, lens, n, match, mismatch, gap_open, gap_extend, C.byref(r),
                                *out)
    if rc != 0:
        return None
    return [tuple(o[k] for o in out) for k in range(n)]


def translate(seq: bytes):
    buf = C.create_string_buffer(len(seq) // 3 + 1)
    k = lib().vfo_translate(seq, len(seq), buf)
    return None if k < 0 else buf.raw[:k]


def is_utf8(b: bytes) -> bool:
    return bool(lib().vfo_is_utf8(b, len(b)))


def aa_table() -> str:
    return lib().vfo_aa_table().decode()


def ascii_to_index():
    p = lib().vfo_ascii_to_index()
    return [p[i] for i in range(128)]


def make_params(adapters, match_score=3, mismatch_score=-2, gap_open_penalty=5,
                gap_extend_penalty=2, accept_prefix_alignment=0.75, accept_suffix_alignment=0.75,
                skip_translation=False, rules: DpRules | None = None) -> Params:
    pre = adapters[0] if isinstance(adapters[0], bytes) else adapters[0].encode()
    suf = adapters[1] if isinstance(adapters[1], bytes) else adapters[1].encode()
    p = Params(pre, len(pre), suf, len(suf), match_score, mismatch_score, gap_open_penalty,
               gap_extend_penalty, accept_prefix_alignment, accept_suffix_alignment,
               1 if skip_translation else 0, rules or DpRules())
    p._keep = (pre, suf)
    return p


def _export(t, as_dict=True):
    L = lib()
    rows = int(L.vfo_table_rows(t))
    kb = int(L.vfo_table_key_bytes(t))
    offs = np.zeros(rows + 1, dtype=np.uint64)
    data = np.zeros(max(kb, 1), dtype=np.uint8)
    counts = np.zeros(max(rows, 1), dtype=np.uint64)
    L.vfo_table_export_ex(t, offs.ctypes.data, data.ctypes.data, counts.ctypes.data, 1 if as_dict else 0)
    if not as_dict:
        return offs, data[:kb], counts[:rows]          # the columns in table order, as the reference's `unzip` leaves them (src/lib.rs:312)
    raw = data.tobytes()
    return {raw[int(offs[i]):int(offs[i + 1])]: int(counts[i]) for i in range(rows)}


def process_reads(params: Params, text, off, length, n_threads=1, want_diag=False, simd=False, as_dict=True):
    """Run the worker+reducer closures over packed reads.

    text: bytes / uint8 array; off, length: uint32 arrays.  Returns (table dict, diag|None, cells).
    simd: gather the alignments of a block of reads and run them sixteen at a time (the CPU baseline's fast leg;
    identical outputs — the parity tests use the scalar default).  as_dict=False: the table as (offsets, data, counts)
    columns in table order instead of a Python dict (what a timed baseline run wants).
    """
    L = lib()
    text = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else \
        np.ascontiguousarray(text, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint32)
    length = np.ascontiguousarray(length, dtype=np.uint32)
    n = len(off)
    diag = np.zeros(n, dtype=DIAG_DTYPE) if want_diag else None
    cells = C.c_uint64(0)
    t = L.vfo_table_new()
    try:
        rc = L.vfo_process_reads_ex(C.byref(params), text.ctypes.data if len(text) else None,
                                    off.ctypes.data if n else None,
                                    length.ctypes.data if n else None, n, n_threads, t,
                                    diag.ctypes.data if want_diag and n else None, C.byref(cells), 1 if simd else 0)
        if rc == -1:
            raise ValueError("Accept alignment threshold must be between 0 and 1.")
        return _export(t, as_dict), diag, int(cells.value)
    finally:
        L.vfo_table_free(t)


def pack_reads(seqs):
    """Pack a list of byte strings back to back -> (text uint8, off uint32, len uint32)."""
    seqs = [s if isinstance(s, bytes) else s.encode() for s in seqs]
    length = np.array([len(s) for s in seqs], dtype=np.uint32)
    off = np.zeros(len(seqs), dtype=np.uint32)
    if len(seqs):
        off[1:] = np.cumsum(length[:-1], dtype=np.uint64).astype(np.uint32)
    text = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    return text, off, length


def find_variants_file(path: str, adapters, n_threads=1, **kw) -> dict:
    """Oracle end-to-end over a gzipped FASTQ: {sequence bytes: count}."""
    L = lib()
    p = make_params(adapters, **kw)
    t = L.vfo_table_new()
    err = C.create_string_buffer(256)
    n = C.c_uint64(0)
    try:
        rc = L.vfo_find_variants_file(os.fsencode(path), C.byref(p), n_threads, t, C.byref(n),
                                      err, 256)
        if rc == -1:
            raise ValueError(err.value.decode())
        if rc == -2:
            raise FileNotFoundError(err.value.decode())
        if rc != 0:
            raise RuntimeError(err.value.decode())
        return _export(t)
    finally:
        L.vfo_table_free(t)


def find_variants_file_timed(path: str, adapters, n_threads=1, simd=True, **kw):
    """The oracle end to end over a gzipped FASTQ file as a CPU baseline: (rows, reads, [inflate s, framing s, closures s]).
    The table is not turned into a Python dict."""
    L = lib()
    p = make_params(adapters, **kw)
    t = L.vfo_table_new()
    err = C.create_string_buffer(256)
    n = C.c_uint64(0)
    ph = (C.c_double * 3)()
    try:
        rc = L.vfo_find_variants_file_ex(os.fsencode(path), C.byref(p), n_threads, t, C.byref(n), err, 256,
                                         1 if simd else 0, ph)
        if rc != 0:
            raise RuntimeError(err.value.decode() or "oracle failed (%d)" % rc)
        return int(L.vfo_table_rows(t)), int(n.value), [float(x) for x in ph]
    finally:
        L.vfo_table_free(t)
