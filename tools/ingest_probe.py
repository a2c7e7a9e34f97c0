"""Write synthetic FASTQ (C3 shape) as plain gzip and as BGZF, time find_variants end to end on both
(single-stream inflate vs member-parallel inflate on n_threads), and check against the oracle."""
import os, struct, subprocess, sys, time, zlib
from concurrent.futures import ProcessPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vfind_b200 import api, find_variants
import oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
path = "/tmp/synth_%d.fq" % n
cfg = api.synth_cfg()
ad = tuple(a.decode() for a in api.synth_adapters(cfg))


def bgzf_member(data):
    co = zlib.compressobj(1, zlib.DEFLATED, -15)
    body = co.compress(data) + co.flush()
    hdr = struct.pack("<BBBBIBBH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(body) + 8 - 1)
    return hdr + body + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))


def bgzf_span(args):
    p, lo, hi = args
    with open(p, "rb") as f:
        f.seek(lo)
        data = f.read(hi - lo)
    return b"".join(bgzf_member(data[i:i + 65280]) for i in range(0, len(data), 65280))


if not os.path.exists(path + ".gz"):
    t0 = time.time()
    text, spans = api.synth_host(cfg, 0, n)
    L = cfg.read_len
    rec = np.empty((n, 2 * L + 17), dtype=np.uint8)
    rec[:, :13] = np.frombuffer(b"@r0000000000\n", np.uint8)
    idx = np.arange(n)
    for d in range(10):
        rec[:, 11 - d] = 48 + (idx // 10 ** d) % 10
    rec[:, 13:13 + L] = text.reshape(n, L)
    rec[:, 13 + L:16 + L] = np.frombuffer(b"\n+\n", np.uint8)
    rec[:, 16 + L:16 + 2 * L] = ord("F")
    rec[:, 16 + 2 * L] = 10
    rec.tofile(path)
    size = os.path.getsize(path)
    step = 65280 * 256
    with ProcessPoolExecutor(os.cpu_count()) as ex, open(path + ".bgzf.gz", "wb") as out:
        for blob in ex.map(bgzf_span, [(path, lo, min(size, lo + step)) for lo in range(0, size, step)]):
            out.write(blob)
        out.write(bgzf_member(b""))
    if "--bgzf-only" in sys.argv:
        os.remove(path)
        open(path + ".gz", "wb").close()
    else:
        subprocess.check_call(["gzip", "-1", "-f", path])
    print("wrote %s.gz (%.0f MB) and .bgzf.gz (%.0f MB) in %.1f s" % (
        path, os.path.getsize(path + ".gz") / 1e6, os.path.getsize(path + ".bgzf.gz") / 1e6, time.time() - t0))

results = {}
for name, p, threads, gpu in (("gzip zlib x1", path + ".gz", 1, "1"), ("gzip pgunzip x3", path + ".gz", 3, "1"),
                              ("gzip pgunzip x%d" % os.cpu_count(), path + ".gz", os.cpu_count(), "1"),
                              ("bgzf host x3", path + ".bgzf.gz", 3, "0"),
                              ("bgzf host x%d" % os.cpu_count(), path + ".bgzf.gz", os.cpu_count(), "0"),
                              ("bgzf GPU inflate", path + ".bgzf.gz", 3, "1")):
    if "--bgzf-only" in sys.argv and "bgzf" not in name:
        continue
    os.environ["VFB_GPU_INFLATE"] = gpu
    os.environ["VFB_INGEST_THREADS"] = str(threads)
    if gpu == "1" and "bgzf" in name:
        pass
    best = 1e9
    for rep in range(4):
        t0 = time.time()
        out = find_variants(p, ad, n_threads=threads, show_progress=False)
        best = min(best, time.time() - t0)
    dt = best
    print("find_variants %-18s best of 4: %.2f s  %.2f M reads/s (whole call: context, ingest, kernels, table)" % (name, dt, n / dt / 1e6))
    cols = out.to_pydict() if hasattr(out, "to_pydict") else out.to_dict(as_series=False)
    results[name] = {k.encode(): v for k, v in zip(cols["sequence"], cols["count"])}
if "--bgzf-only" in sys.argv:
    assert len({tuple(sorted(r.items())) for r in results.values()}) == 1
    print("tables identical across inflate modes")
    sys.exit(0)
t0 = time.time()
tab = oracle.find_variants_file(path + ".gz", ad, n_threads=os.cpu_count())
dt = time.time() - t0
print("oracle (cpu, %d threads): %.2f s  %.2f M reads/s  rows %d" % (os.cpu_count(), dt, n / dt / 1e6, len(tab)))
assert all(r == tab for r in results.values())
print("tables identical")
