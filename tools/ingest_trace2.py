"""VFB_TRACE=1 python tools/ingest_trace2.py N  — one traced find_variants over a BGZF file of N C3-shaped reads
(written by tools/ingest_probe.py N --bgzf-only), then three timed repeats."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vfind_b200 import api, find_variants
n = int(sys.argv[1])
p = "/tmp/synth_%d.fq.bgzf.gz" % n
cfg = api.synth_cfg()
ad = tuple(a.decode() for a in api.synth_adapters(cfg))
for rep in range(4):
    if rep == 1:
        os.environ.pop("VFB_TRACE", None)
    t0 = time.time()
    out = find_variants(p, ad, show_progress=False)
    dt = time.time() - t0
    print("rep %d: %.3f s  %.2f M reads/s" % (rep, dt, n / dt / 1e6), flush=True)
