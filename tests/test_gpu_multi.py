"""Several devices behind one call (vfb_multi_*): the file's segments dealt round robin, the
per-device tables merged inside the library, ONE table out — against the CPU oracle, bit for bit.

The driver's test box has one GPU: the merge logic is then exercised with two contexts on the same device
(VFB_MULTI_ALLOW_DUP=1: peer copies device -> same device); with two or more GPUs the same tests use them."""
import gzip
import os
import random
import subprocess
import sys

import numpy as np
import pytest

import oracle
from test_gpu_inflate import _rows, bgzf_file
from test_gpu_parity import PREFIX, SUFFIX, make_reads, spans_of
from vfind_b200 import api

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    return api._visible_devices()


def two_devices():
    if n_gpus() >= 2:
        return [0, 1]
    os.environ["VFB_MULTI_ALLOW_DUP"] = "1"
    return [0, 0]


def fastq_text(seqs):
    return "".join("@r%d\n%s\n+\n%s\n" % (i, s.decode(), "F" * len(s)) for i, s in enumerate(seqs)).encode()


def test_multi_submit_merge_and_resubmit():
    """Halves of the reads on two contexts; merged table == oracle; then more reads and a second merge (rows that
    went to their owner keep their slot with a zero count and must not come back)."""
    rng = random.Random(71)
    seqs = make_reads(rng, PREFIX, SUFFIX, 30000, lib=3000)
    more = make_reads(rng, PREFIX, SUFFIX, 10000, lib=3000) + seqs[:5000]
    want1 = oracle.process_reads(oracle.make_params((PREFIX, SUFFIX)), *oracle.pack_reads(seqs), n_threads=8)[0]
    want2 = oracle.process_reads(oracle.make_params((PREFIX, SUFFIX)), *oracle.pack_reads(seqs + more), n_threads=8)[0]
    devs = two_devices()
    with api.MultiContext((PREFIX, SUFFIX), devices=devs) as m:
        assert m.n_devices == 2
        for k, part in enumerate((seqs[::2], seqs[1::2])):
            text, off, ln = oracle.pack_reads(part)
            m.contexts[k].submit_host(text, spans_of(off, ln))
        got1 = m.finish_dict()
        assert got1 == want1
        # every key sits on the context that owns it
        for k, c in enumerate(m.contexts):
            for key in c.finish_dict():
                assert api.key_owner(api.hash_key(key), 2) == k
        for k, part in enumerate((more[::2], more[1::2])):
            text, off, ln = oracle.pack_reads(part)
            m.contexts[1 - k].submit_host(text, spans_of(off, ln))
        assert m.finish_dict() == want2
        batch = m.finish_arrow()
        assert {k.encode(): v for k, v in zip(batch.column(0).to_pylist(), batch.column(1).to_pylist())} == want2
        assert m.stats()["reads"] == len(seqs) + len(more)


@pytest.mark.parametrize("raw_block", [None, "131072"])
def test_multi_file_variants(tmp_path, raw_block):
    """find_variants(devices=[..]) on block-gzip files whose records straddle members, segments and raw blocks."""
    from vfind_b200 import find_variants
    rng = random.Random(72)
    seqs = make_reads(rng, PREFIX, SUFFIX, 40000, lib=500, lead=(0, 40))
    ad = (PREFIX.decode(), SUFFIX.decode())
    text = fastq_text(seqs)
    ref = tmp_path / "ref.fq.gz"
    ref.write_bytes(gzip.compress(text))
    want = oracle.find_variants_file(str(ref), (PREFIX, SUFFIX), n_threads=8)
    half = text.rfind(b"\n@r", 0, len(text) // 2) + 1
    files = {
        "bgzf": bgzf_file(text, block=65280, level=1),
        "bgzf_odd_blocks": bgzf_file(text, block=12345, eof=False),
        "bgzf_then_gzip": bgzf_file(text[:half + 3], eof=False) + gzip.compress(text[half + 3:]),
        "plain_gzip": gzip.compress(text, 1),
    }
    devs = two_devices()
    if raw_block:
        os.environ["VFB_RAW_BLOCK"] = raw_block
    try:
        for name, blob in files.items():
            p = tmp_path / (name + ".fq.gz")
            p.write_bytes(blob)
            for chunk in ("40000", "300000", None):
                if chunk:
                    os.environ["VFB_INGEST_CHUNK"] = chunk
                try:
                    for d in ([0], devs):
                        assert _rows(find_variants(str(p), ad, n_threads=4, devices=d, show_progress=False)) == want, (name, chunk, d)
                        if "gzip" in name:
                            # the plain gzip part decoded on the device (forced: these files are small), its text handed on
                            # device to device — to the other context's device by a peer copy
                            os.environ["VFB_GPU_GUNZIP"] = "2"
                            try:
                                assert _rows(find_variants(str(p), ad, n_threads=4, devices=d, show_progress=False)) == want, (name, chunk, d, "device gunzip")
                            finally:
                                os.environ.pop("VFB_GPU_GUNZIP", None)
                finally:
                    os.environ.pop("VFB_INGEST_CHUNK", None)
    finally:
        os.environ.pop("VFB_RAW_BLOCK", None)


def test_multi_errors(tmp_path):
    from vfind_b200 import PanicException, find_variants
    rng = random.Random(73)
    seqs = make_reads(rng, PREFIX, SUFFIX, 20000, lib=100)
    ad = (PREFIX.decode(), SUFFIX.decode())
    text = fastq_text(seqs)
    devs = two_devices()
    os.environ["VFB_INGEST_CHUNK"] = "100000"
    try:
        whole = bgzf_file(text, block=20000)
        p = tmp_path / "bad.fq.gz"
        bad = bytearray(whole); bad[len(bad) // 2] ^= 0x21
        p.write_bytes(bytes(bad))
        with pytest.raises(PanicException):
            find_variants(str(p), ad, devices=devs)
        p.write_bytes(whole[:len(whole) - 40])
        with pytest.raises(PanicException, match="truncated|invalid"):
            find_variants(str(p), ad, devices=devs)
        # a malformed record deep in the file is reported with its global index, whichever device parsed it
        lines = text.split(b"\n")
        lines[4 * 15000 + 2] = b"-"
        p.write_bytes(bgzf_file(b"\n".join(lines), block=20000))
        with pytest.raises(PanicException, match="record 15000"):
            find_variants(str(p), ad, devices=devs)
        with pytest.raises(ValueError):
            find_variants(str(p), ad, device=0, devices=devs)
    finally:
        os.environ.pop("VFB_INGEST_CHUNK", None)


def test_multi_gpu_2m_reads_match_oracle(tmp_path):
    """merged table == oracle on 2 M C3-shaped reads from a block-gzip file, all visible GPUs
    (two contexts on one device when the box has one)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import synth_fastq
    from vfind_b200 import find_variants
    cfg = api.synth_cfg(n_variants=200000)
    n = 2_000_000
    fq = str(tmp_path / "c3.fq.gz")
    synth_fastq.write_bgzf_fastq(fq, cfg, n, api)
    adapters = api.synth_adapters(cfg)
    text, spans = api.synth_host(cfg, 0, n)
    want = oracle.process_reads(oracle.make_params(adapters), text, spans["off"], spans["len"], n_threads=os.cpu_count() or 1)[0]
    devs = list(range(n_gpus())) if n_gpus() >= 2 else two_devices()
    got = _rows(find_variants(fq, tuple(a.decode() for a in adapters), devices=devs, show_progress=False))
    assert got == want
    one = _rows(find_variants(fq, tuple(a.decode() for a in adapters), device=0, show_progress=False))
    assert one == want


def test_nccl_merge_two_ranks(tmp_path):
    """One process per GPU: vfb_merge_nccl inside the library == the single-GPU table (needs two GPUs)."""
    if n_gpus() < 2:
        pytest.skip("needs two GPUs")
    script = os.path.join(ROOT, "tests", "nccl_merge_worker.py")
    out = tmp_path / "out"
    out.mkdir()
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29611", script, str(out)], cwd=ROOT, timeout=600)
    parts = [np.load(str(out / ("rank%d.npz" % r)), allow_pickle=True) for r in range(2)]
    merged = {}
    for r, pz in enumerate(parts):
        for k, v in zip(pz["keys"], pz["counts"]):
            assert bytes(k) not in merged
            assert api.key_owner(api.hash_key(bytes(k)), 2) == r
            merged[bytes(k)] = int(v)
    cfg = api.synth_cfg(seed=5, n_variants=50000)
    text, spans = api.synth_host(cfg, 0, 400000)
    want = oracle.process_reads(oracle.make_params(api.synth_adapters(cfg)), text, spans["off"], spans["len"], n_threads=os.cpu_count() or 1)[0]
    assert merged == want
