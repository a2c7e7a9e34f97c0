"""Per-call overhead of find_variants: the reference's toy file (8 reads) and a small block-gzip file, first and later calls."""
import gzip
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json  # noqa: E402

import oracle  # noqa: E402
from vfind_b200 import find_variants  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))["toy"]
toy = "/tmp/toy.fq.gz"
with gzip.open(toy, "wb") as f:
    for r in g["reads"]:
        f.write(("@%s\n%s\n+\n%s\n" % (r["header"], r["seq"], r["qual"])).encode())
ads = tuple(g["adapters"])
for rep in range(6):
    t0 = time.perf_counter()
    out = find_variants(toy, ads, show_progress=False)
    print("toy call %d: %.2f ms  rows %d" % (rep, 1e3 * (time.perf_counter() - t0), out.num_rows), flush=True)
if "--toy-only" in sys.argv:
    sys.exit(0)
cfg = oracle.synth_cfg()
small = "/tmp/small.fq.gz"
oracle.write_fastq(cfg, 0, 4_000_000, small, bgzf=True)
ads2 = tuple(a.decode() for a in oracle.synth_adapters(cfg))
for rep in range(5):
    if rep == 4:
        os.environ["VFB_TRACE"] = "1"
        os.environ["VFB_INGEST_TRACE"] = "1"
    t0 = time.perf_counter()
    out = find_variants(small, ads2, show_progress=False, device=0)
    print("4 M-read BGZF call %d: %.1f ms  (%.1f M reads/s)  rows %d" % (rep, 1e3 * (time.perf_counter() - t0), 4e6 / (time.perf_counter() - t0) / 1e6, out.num_rows), flush=True)
os.environ.pop("VFB_TRACE", None)
os.environ.pop("VFB_INGEST_TRACE", None)
os.environ["VFB_TRACE"] = "1"
t0 = time.perf_counter()
find_variants(toy, ads, show_progress=False)
print("toy call (traced): %.2f ms" % (1e3 * (time.perf_counter() - t0)))
