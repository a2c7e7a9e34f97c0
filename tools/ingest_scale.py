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
        del out
    # the same call in pieces: where the time outside the ingest goes
    t0 = time.time()
    ctx = api.Context(ads, device=0) if n == 1 else api.MultiContext(ads, devices=list(range(n)))
    t1 = time.time()
    ctx.run_file(path)
    t2 = time.time()
    batch = ctx.finish_arrow()
    t3 = time.time()
    frame = api.batch_to_frame(batch)
    t4 = time.time()
    ctx.close()
    t5 = time.time()
    del batch, frame
    t6 = time.time()
    print("n=%d pieces: create %.1f ms, run_file %.1f ms, finish_arrow %.1f ms, to frame %.1f ms, close %.1f ms, release result %.1f ms"
          % (n, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4), 1e3 * (t6 - t5)), flush=True)
os.remove(path)
