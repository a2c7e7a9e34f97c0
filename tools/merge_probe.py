"""Time the pieces of the in-library NCCL table merge (run under torchrun; VFB_TRACE=1 prints per-phase stamps)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from vfind_b200 import api
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ids = [api.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
comm = api.nccl_comm_init(ids[0], world, rank, local)
cfg = api.synth_cfg(); ad = api.synth_adapters(cfg)
n = 12_500_000; chunks = []
for c in range(8):
    t = torch.empty(n * 250, dtype=torch.uint8, device=dev); s = torch.empty(n * 2, dtype=torch.int32, device=dev)
    api.synth_device(cfg, rank * 100_000_000 + c * n, n, t.data_ptr(), s.data_ptr(), local); chunks.append((t, s))
ctx = api.Context(ad, device=local, table_capacity_hint=40_000_000, batch_reads=n)
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(4):
    ctx.table_clear()
    for t, s in chunks: ctx.submit_device(t.data_ptr(), t.numel(), s.data_ptr(), n)
    ctx.sync(); dist.barrier(); t0 = T()
    ctx.merge_nccl(comm, rank, world)
    ctx.sync(); t1 = T()
    if rank == 0:
        print("rep %d: merge %.2f ms" % (rep, 1e3 * (t1 - t0)), flush=True)
api.nccl_comm_destroy(comm)
dist.destroy_process_group()
