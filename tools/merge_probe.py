"""Time the pieces of the multi-GPU table merge (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from vfind_b200 import api
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = api.synth_cfg(); ad = api.synth_adapters(cfg)
n = 12_500_000; chunks = []
for c in range(8):
    t = torch.empty(n * 250, dtype=torch.uint8, device=dev); s = torch.empty(n * 2, dtype=torch.int32, device=dev)
    api.synth_device(cfg, rank * 100_000_000 + c * n, n, t.data_ptr(), s.data_ptr(), local); chunks.append((t, s))
ctx = api.Context(ad, device=local, table_capacity_hint=40_000_000, batch_reads=n)
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    ctx.table_clear()
    for t, s in chunks: ctx.submit_device(t.data_ptr(), t.numel(), s.data_ptr(), n)
    ctx.sync(); dist.barrier(); t0 = T()
    sizes = ctx.partition_sizes(world); t1 = T()
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    send = torch.empty(int(sum(sizes)), dtype=torch.uint8, device=dev); t2 = T()
    ctx.partition_fill(world, send.data_ptr(), offs); t3 = T()
    ss = torch.tensor(sizes, dtype=torch.int64, device=dev); rs = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rs, ss); rsl = [int(x) for x in rs.cpu().tolist()]; t4 = T()
    recv = torch.empty(sum(rsl), dtype=torch.uint8, device=dev)
    dist.all_to_all_single(recv, send, output_split_sizes=rsl, input_split_sizes=sizes); t5 = T()
    ctx.table_clear(); t6 = T()
    o = 0
    for r in rsl: ctx.absorb(recv.data_ptr() + o, r); o += r
    ctx.sync(); t7 = T()
    if rank == 0:
        print("sizes %.1f alloc %.1f fill %.1f a2a-sizes %.1f a2a %.1f clear %.1f absorb %.1f total %.1f ms (send %.0f MB)" % tuple(
            [1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t7 - t0)] + [sum(sizes) / 1e6]), flush=True)
dist.destroy_process_group()
