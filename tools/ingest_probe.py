"""Write a synthetic gzipped FASTQ (C3 shape) and time find_variants end to end on it."""
import gzip, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vfind_b200 import api, find_variants
import oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
path = "/tmp/synth_%d.fq" % n
cfg = api.synth_cfg()
ad = tuple(a.decode() for a in api.synth_adapters(cfg))
if not os.path.exists(path + ".gz"):
    t0 = time.time()
    text, spans = api.synth_host(cfg, 0, n)
    L = cfg.read_len
    reads = text.reshape(n, L)
    rec = np.empty((n, 2 * L + 16), dtype=np.uint8)
    hdr = np.frombuffer(b"@r0000000000\n", np.uint8)       # fixed-width header
    rec[:, :13] = hdr
    idx = np.arange(n)
    for d in range(10):
        rec[:, 11 - d] = 48 + (idx // 10 ** d) % 10
    rec[:, 13:13 + L] = reads
    rec[:, 13 + L:13 + L + 3] = np.frombuffer(b"\n+\n", np.uint8)
    rec[:, 16 + L:16 + 2 * L] = ord("F")
    rec = np.concatenate([rec[:, :16 + 2 * L], np.full((n, 1), 10, np.uint8)], axis=1)
    rec.tofile(path)
    subprocess.check_call(["gzip", "-1", "-f", path])
    print("wrote %s.gz in %.1f s (%.1f MB)" % (path, time.time() - t0, os.path.getsize(path + ".gz") / 1e6))
for rep in range(2):
    t0 = time.time()
    out = find_variants(path + ".gz", ad, show_progress=False)
    dt = time.time() - t0
    print("find_variants: %.2f s  %.2f M reads/s  rows %d" % (dt, n / dt / 1e6, out.num_rows if hasattr(out, "num_rows") else len(out)))
t0 = time.time()
tab = oracle.find_variants_file(path + ".gz", ad, n_threads=os.cpu_count())
dt = time.time() - t0
print("oracle (cpu, %d threads): %.2f s  %.2f M reads/s  rows %d" % (os.cpu_count(), dt, n / dt / 1e6, len(tab)))
cols = out.to_pydict() if hasattr(out, "to_pydict") else out.to_dict(as_series=False)
assert {k.encode(): v for k, v in zip(cols["sequence"], cols["count"])} == tab
print("tables identical")
