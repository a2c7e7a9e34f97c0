"""GPU inflate throughput on distinct block-gzip members of a synthetic FASTQ file (generated here).
    python tools/inflate_probe.py [members]"""
import ctypes, os, struct, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from vfind_b200 import api
n_members = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
path = "/tmp/inflate_probe.fq.gz"
cfg = oracle.synth_cfg()
oracle.write_fastq(cfg, 0, n_members * 127, path, bgzf=True)        # ~126 records of 517 bytes per 65280-byte member
raw = np.fromfile(path, dtype=np.uint8)
blob = raw.tobytes()
tab, p, oo = [], 0, 0
while p < len(blob) and len(tab) < n_members:
    bsize = struct.unpack_from("<H", blob, p + 16)[0] + 1
    isize = struct.unpack_from("<I", blob, p + bsize - 4)[0]
    tab.append((p, bsize, oo, isize)); p += bsize; oo += isize
tab = np.array(tab, dtype=np.uint32)
L = api.load_library()
L.vfb_debug_gpu_inflate.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                    ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_double)]
out = np.zeros(oo, dtype=np.uint8)
for rep in range(3):
    bad, ms = ctypes.c_uint32(0), ctypes.c_double(0)
    rc = L.vfb_debug_gpu_inflate(raw.ctypes.data, p, tab.ctypes.data, len(tab), out.ctypes.data, oo, -1, ctypes.byref(bad), ctypes.byref(ms))
    assert rc == 0
print("members=%d text=%.0f MB compressed=%.0f MB kernel %.2f ms = %.1f GB/s of text, bad=%x" % (len(tab), oo / 1e6, p / 1e6, ms.value, oo / ms.value / 1e6, bad.value))
os.remove(path)
