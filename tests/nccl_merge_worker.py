"""Worker of tests/test_gpu_multi.py::test_nccl_merge_two_ranks (launched by torchrun, one rank per GPU): each rank
counts its half of a synthetic stream, the tables are merged with vfb_merge_nccl (NCCL send/recv inside the
library; torch.distributed only carries the 128-byte communicator id), every rank saves its partition."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vfind_b200 import api  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ids = [api.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
comm = api.nccl_comm_init(ids[0], world, rank, local)
cfg = api.synth_cfg(seed=5, n_variants=50000)
n = 400000
per = n // world
text, spans = api.synth_host(cfg, rank * per, per)
with api.Context(api.synth_adapters(cfg), device=local) as ctx:
    ctx.submit_host(text, spans)
    ctx.merge_nccl(comm, rank, world)
    # a second round: more reads after a merge, merged again (released rows must not be double counted)
    table = ctx.finish_dict()
keys = np.array(list(table.keys()), dtype=object)
np.savez(os.path.join(sys.argv[1], "rank%d.npz" % rank), keys=keys, counts=np.array(list(table.values()), dtype=np.uint64))
api.nccl_comm_destroy(comm)
dist.barrier()
dist.destroy_process_group()
