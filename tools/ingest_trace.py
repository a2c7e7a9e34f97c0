import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vfind_b200 import api, find_variants
n = 4000000
path = "/tmp/synth_%d.fq" % n
cfg = api.synth_cfg()
ad = tuple(a.decode() for a in api.synth_adapters(cfg))
os.environ["VFB_INGEST_TRACE"] = "1"
for p, th in ((path + ".bgzf.gz", 16), (path + ".bgzf.gz", 16)):
    t0 = time.time()
    ctx = api.Context(ad, n_threads=th)
    t1 = time.time()
    ctx.run_file(p)
    t2 = time.time()
    o, d, c = ctx.finish_arrays()
    t3 = time.time()
    fr = api.table_to_frame(o, d, c)
    t4 = time.time()
    ctx.close()
    print("create %.2f run %.2f finish %.2f frame %.2f close %.2f" % (t1 - t0, t2 - t1, t3 - t2, t4 - t3, time.time() - t4), flush=True)
    t0 = time.time()
    out = find_variants(p, ad, n_threads=th, show_progress=False)
    print("find_variants total %.2f" % (time.time() - t0), flush=True)
