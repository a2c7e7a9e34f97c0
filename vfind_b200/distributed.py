"""Multi-GPU merge of per-GPU count tables; one process per GPU.

The reference has no multi-process path.  Reads shard across ranks with no data-path
collective; the only exchange is this final merge: every key is owned by one rank
(owner = f(hash(key))), each rank exports its table as one chunk per owner, chunks are
exchanged with an all-to-all (NCCL over NVLink on GPUs, gloo in the CPU tests), and each
rank absorbs what it received.  Counts are integers, so the merged table is bit-identical
for any number of ranks.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def merge_tables(ctx, group=None, device=None):
    """All-to-all merge.  Afterwards `ctx` holds exactly the keys this rank owns, with
    global counts.  `ctx` needs partition_sizes / partition_fill / table_clear / absorb
    (vfind_b200.api.Context, or a host stand-in in the CPU tests)."""
    world = dist.get_world_size(group)
    if world == 1:
        return
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    sizes = ctx.partition_sizes(world)
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    send = torch.empty(max(int(sum(sizes)), 1), dtype=torch.uint8, device=device)
    ctx.partition_fill(world, send.data_ptr(), offs)
    send_sizes = torch.tensor(sizes, dtype=torch.int64, device=device)
    recv_sizes = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(recv_sizes, send_sizes, group=group)
    rs = [int(x) for x in recv_sizes.cpu().tolist()]
    recv = torch.empty(max(sum(rs), 1), dtype=torch.uint8, device=device)
    dist.all_to_all_single(recv, send[:sum(sizes)] if sum(sizes) else send[:0], output_split_sizes=rs,
                           input_split_sizes=sizes, group=group)
    if device.type == "cuda":
        torch.cuda.current_stream(device).synchronize()
    ctx.table_clear()
    o = 0
    for r in rs:
        ctx.absorb(recv.data_ptr() + o, r)
        o += r
    ctx._merge_keepalive = (send, recv)


def gather_table(offsets, data, counts, group=None, dst=0):
    """Concatenate every rank's disjoint partition on rank `dst` (host arrays, any backend)."""
    world = dist.get_world_size(group)
    if world == 1:
        return offsets, data, counts
    parts = [None] * world
    dist.all_gather_object(parts, (offsets, data, counts), group=group)
    if dist.get_rank(group) != dst:
        return None
    rows = sum(len(p[2]) for p in parts)
    offs = np.zeros(rows + 1, dtype=np.uint64)
    k, base = 0, 0
    for o, d, c in parts:
        n = len(c)
        offs[k:k + n + 1] = o.astype(np.uint64) + base
        k += n
        base += int(o[-1])
    return offs, np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts])
