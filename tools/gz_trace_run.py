import os, sys, time
sys.path.insert(0, "/root/repo")
import oracle, subprocess
from vfind_b200 import find_variants
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
oracle.build()
cfg = oracle.synth_cfg()
ad = tuple(a.decode() for a in oracle.synth_adapters(cfg))
txt = "/tmp/gzt.fq"; gz = txt + ".gz"
oracle.write_fastq(cfg, 0, n, txt, bgzf=False)
subprocess.check_call([sys.executable, "/root/repo/tools/single_stream_gzip.py", txt, gz, "1"])
os.remove(txt)
os.environ["VFB_TRACE"] = "1"
os.environ["VFB_INGEST_TRACE"] = "1"
os.environ["VFB_GUNZIP_TRACE"] = "1"
for rep in range(3):
    print("=== call %d" % rep, file=sys.stderr, flush=True)
    t0 = time.time()
    find_variants(gz, ad, show_progress=False, devices=[0])
    print("call %d %.3f s" % (rep, time.time() - t0), flush=True)
