import time, sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from vfind_b200 import api
cfg = api.synth_cfg()
ad = api.synth_adapters(cfg)
n = 12_500_000; L = 250
dev = torch.device('cuda', 0)
t = torch.empty(n*L, dtype=torch.uint8, device=dev); s = torch.empty(n*2, dtype=torch.int32, device=dev)
api.synth_device(cfg, 0, n, t.data_ptr(), s.data_ptr(), 0)
ht = torch.empty(n*L, dtype=torch.uint8, pin_memory=True); hs = torch.empty(n*2, dtype=torch.int32, pin_memory=True)
ht.copy_(t); hs.copy_(s); torch.cuda.synchronize()
# raw H2D bandwidth
d2 = torch.empty_like(t)
for _ in range(2):
    t0 = time.perf_counter(); d2.copy_(ht, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter()-t0
print("raw H2D GB/s", n*L/dt/1e9)
ctx = api.Context(ad, device=0, table_capacity_hint=40_000_000, batch_reads=n)
for rep in range(3):
    ctx.table_clear(); ctx.sync()
    t0 = time.perf_counter()
    ctx.submit_host_ptr(ht.data_ptr(), ht.numel(), hs.data_ptr(), n)
    t1 = time.perf_counter()
    ctx.sync()
    t2 = time.perf_counter()
    o, d, c = ctx.finish_arrays()
    t3 = time.perf_counter()
    print("submit %.1f ms  sync %.1f ms  finish %.1f ms  rows %d" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, len(c)))
