"""find_variants(path, devices=[0..n-1]) on one big block-gzip file for n = 1, 2, 4, 8 (as many as visible), with
VFB_INGEST_TRACE: where the reader, the indexer, the per-device workers and the segment chain spend their time."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (generator only)
from vfind_b200 import api, find_variants  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = oracle.synth_cfg()
ads = tuple(a.decode() for a in oracle.synth_adapters(cfg))
path = "/tmp/ingest_scale.fq.gz"
t0 = time.time()
oracle.write_fastq(cfg, 0, min(n_reads, 10_000_000), path + ".b", bgzf=True, append=True)
with open(path, "wb") as out:
    for _ in range(max(1, n_reads // 10_000_000)):
        out.write(open(path + ".b", "rb").read())
os.remove(path + ".b")
oracle.write_fastq(cfg, 0, 0, path, bgzf=True, append=True)
print("file: %d reads, %.2f GB, %.1f s" % (n_reads, os.path.getsize(path) / 1e9, time.time() - t0), flush=True)
vis = api._visible_devices()
for n in (1, 2, 4, 8):
    if n > vis:
        break
    for rep in range(reps):
        if rep == reps - 1:
            os.environ["VFB_INGEST_TRACE"] = "1"
        t0 = time.time()
        out = find_variants(path, ads, show_progress=False, devices=list(range(n)))
        dt = time.time() - t0
        os.environ.pop("VFB_INGEST_TRACE", None)
        print("n=%d rep %d: %.3f s  %.1f M reads/s  rows %d" % (n, rep, dt, n_reads / dt / 1e6, out.num_rows), flush=True)
os.remove(path)
