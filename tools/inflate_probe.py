"""GPU inflate throughput on distinct block-gzip members (the BGZF file tools/ingest_probe.py writes)."""
import ctypes, os, struct, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vfind_b200 import api
path = sys.argv[2] if len(sys.argv) > 2 else "/tmp/synth_4000000.fq.bgzf.gz"
raw = np.fromfile(path, dtype=np.uint8)
blob = raw.tobytes()
tab, p, oo = [], 0, 0
while p < len(blob) and len(tab) < int(sys.argv[1]):
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
print("lanes=%s members=%d text=%.0f MB kernel %.1f ms = %.2f GB/s bad=%x" % (os.environ.get("VFB_INFLATE_LANES", "1"), len(tab), oo / 1e6, ms.value, oo / ms.value / 1e6, bad.value))
