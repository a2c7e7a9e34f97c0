"""How often a find_variants call on ONE plain gzip stream has to wait for the count table's counters (VFB_TRACE lines of
table_reserve), per call, with the call's time.   python tools/gz_reserve_trace.py [reads] 2> trace.err"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from vfind_b200 import find_variants
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
oracle.build()
cfg = oracle.synth_cfg()
ad = tuple(a.decode() for a in oracle.synth_adapters(cfg))
txt = "/tmp/gzr.fq"; gz = txt + ".gz"
oracle.write_fastq(cfg, 0, n, txt, bgzf=False)
subprocess.check_call([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "single_stream_gzip.py"), txt, gz, "1"])
os.remove(txt)
find_variants(gz, ad, show_progress=False, device=0)          # warm: context, pools (set VFB_TRACE=1 VFB_GUNZIP_TRACE=1 outside)
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    print("=== call %d" % rep, file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    out = find_variants(gz, ad, show_progress=False, device=0)
    print("call %d: %.3f s = %.1f M reads/s, rows %d" % (rep, time.perf_counter() - t0, n / (time.perf_counter() - t0) / 1e6, out.num_rows), flush=True)
