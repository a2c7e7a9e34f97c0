"""Ingest trace of one find_variants call on a block-gzip FASTQ file of N C3-shaped reads (standalone generator):
    python tools/ingest_trace3.py [reads] [devices]      (VFB_INGEST_TRACE / VFB_TRACE lines go to stderr)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from vfind_b200 import api, find_variants

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30_000_000
devs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
oracle.build()
cfg = oracle.synth_cfg()
ad = tuple(a.decode() for a in oracle.synth_adapters(cfg))
path = "/tmp/trace_%d.fq.gz" % n
block = min(n, 10_000_000)
t0 = time.time()
done = 0
open(path, "wb").close()
while done < n:
    m = min(block, n - done)
    oracle.write_fastq(cfg, 0, m, path, bgzf=True, level=1, append=True)
    done += m
oracle.write_fastq(cfg, 0, 0, path, bgzf=True, level=1, append=True)
print("wrote %s (%.0f MB) in %.1f s" % (path, os.path.getsize(path) / 1e6, time.time() - t0), flush=True)
for rep in range(3):
    if rep == 2:
        os.environ["VFB_INGEST_TRACE"] = "1"
        os.environ["VFB_TRACE"] = "1"
    t0 = time.time()
    out = find_variants(path, ad, show_progress=False, devices=devs)
    dt = time.time() - t0
    print("call %d: %.3f s  %.1f M reads/s" % (rep, dt, n / dt / 1e6), flush=True)
