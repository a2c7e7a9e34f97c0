"""One find_variants call on a plain gzip file of N C3-shaped reads (for ncu captures of the k_gz_* kernels)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from vfind_b200 import find_variants
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
oracle.build()
cfg = oracle.synth_cfg()
ad = tuple(a.decode() for a in oracle.synth_adapters(cfg))
txt = "/tmp/gzk_%d.fq" % n
gz = txt + ".gz"
oracle.write_fastq(cfg, 0, n, txt, bgzf=False)
subprocess.check_call([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "single_stream_gzip.py"), txt, gz, "1"])
os.remove(txt)
out = find_variants(gz, ad, show_progress=False, devices=[0])
print("rows", out.num_rows if hasattr(out, "num_rows") else len(out))
os.remove(gz)
