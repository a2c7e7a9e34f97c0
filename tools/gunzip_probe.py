"""Plain (single-stream) gzip FASTQ through find_variants: the device decoder (VFB_GPU_GUNZIP unset / 1) against the host
threads (VFB_GPU_GUNZIP=0).   python tools/gunzip_probe.py [reads] [level]"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from vfind_b200 import find_variants

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
level = sys.argv[2] if len(sys.argv) > 2 else "1"
oracle.build()
cfg = oracle.synth_cfg()
ad = tuple(a.decode() for a in oracle.synth_adapters(cfg))
txt = "/tmp/gzp_%d.fq" % n
t0 = time.time()
oracle.write_fastq(cfg, 0, n, txt, bgzf=False)
gz = txt + ".gz"
if "--gzip" in sys.argv:
    subprocess.check_call(["gzip", "-%s" % level, "-f", txt])
else:
    # ONE deflate stream compressed in parallel the way pigz does it: slices end with a sync flush (an empty stored block,
    # byte aligned), the last one with the final block; one gzip header in front, one trailer behind
    import struct, zlib
    from concurrent.futures import ProcessPoolExecutor
    size = os.path.getsize(txt)
    step = 64 << 20
    def piece(args):
        lo, hi, last = args
        with open(txt, "rb") as f:
            f.seek(lo)
            data = f.read(hi - lo)
        co = zlib.compressobj(int(level), zlib.DEFLATED, -15)
        return co.compress(data) + (co.flush(zlib.Z_FINISH) if last else co.flush(zlib.Z_SYNC_FLUSH)), zlib.crc32(data), len(data)
    import multiprocessing
    multiprocessing.set_start_method("fork", force=True)
    jobs = [(lo, min(size, lo + step), lo + step >= size) for lo in range(0, size, step)]
    crcs = []
    with ProcessPoolExecutor(os.cpu_count()) as ex, open(gz, "wb") as out:
        out.write(b"\x1f\x8b\x08\x00\0\0\0\0\0\xff")
        for blob, c, ln in ex.map(piece, jobs):
            out.write(blob)
            crcs.append((c, ln))
    crc = 0
    with open(txt, "rb") as f:
        while True:
            b = f.read(64 << 20)
            if not b:
                break
            crc = zlib.crc32(b, crc)
    with open(gz, "ab") as out:
        out.write(struct.pack("<II", crc & 0xffffffff, size & 0xffffffff))
    os.remove(txt)
print("wrote %s (%.0f MB) in %.1f s" % (gz, os.path.getsize(gz) / 1e6, time.time() - t0), flush=True)
tables = {}
for mode in ("1", "0", "1"):
    os.environ["VFB_GPU_GUNZIP"] = mode
    best = 1e9
    for rep in range(3):
        if rep == 2 and mode == "1":
            os.environ["VFB_GUNZIP_TRACE"] = "1"
        t0 = time.time()
        out = find_variants(gz, ad, show_progress=False, devices=[0])
        best = min(best, time.time() - t0)
        os.environ.pop("VFB_GUNZIP_TRACE", None)
    cols = out.to_pydict() if hasattr(out, "to_pydict") else out.to_dict(as_series=False)
    tables[mode] = dict(zip(cols["sequence"], cols["count"]))
    print("VFB_GPU_GUNZIP=%s: best of 3 %.3f s = %.1f M reads/s" % (mode, best, n / best / 1e6), flush=True)
assert tables["0"] == tables["1"]
print("tables identical (%d rows)" % len(tables["0"]))
os.remove(gz)
