"""CPU (gloo, world_size 2) tests of the multi-rank merge plumbing in vfind_b200/distributed.py.

The GPU table is replaced by a host stand-in that speaks the same chunk format
(vfind_b200/csrc/vfb_internal.cuh: ChunkHeader + hash/count/koff/klen sections + key area)
and uses the library's own host hash/owner functions, so the exchange logic — sizes
all-to-all, payload all-to-all with uneven splits, absorb, gather — runs exactly as on GPUs.
"""
import ctypes
import os
import random
import struct
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAGIC = 0x5646423230304B31


def a16(x):
    return (x + 15) & ~15


class HostTable:
    """Host stand-in for vfind_b200.api.Context's merge entry points."""

    def __init__(self, api):
        self.api = api
        self.t = {}
        self._parts = None

    def add(self, key: bytes, n: int = 1):
        self.t[key] = self.t.get(key, 0) + n

    def partition_sizes(self, n_parts):
        parts = [[] for _ in range(n_parts)]
        for k, c in self.t.items():
            h = self.api.hash_key(k)
            parts[self.api.key_owner(h, n_parts)].append((h, c, k))
        self._parts = parts
        return [32 + 3 * a16(8 * len(p)) + a16(4 * len(p)) + a16(sum(a16(len(k)) for _, _, k in p)) for p in parts]

    def partition_fill(self, n_parts, buf_ptr, offsets):
        for p, off in zip(self._parts, offsets):
            n = len(p)
            koffs, o = [], 0
            for _, _, k in p:
                koffs.append(o)
                o += a16(len(k))
            blob = struct.pack("<4Q", MAGIC, n, o, 0)
            blob += np.array([h for h, _, _ in p], np.uint64).tobytes().ljust(a16(8 * n), b"\0")
            blob += np.array([c for _, c, _ in p], np.uint64).tobytes().ljust(a16(8 * n), b"\0")
            blob += np.array(koffs, np.uint64).tobytes().ljust(a16(8 * n), b"\0")
            blob += np.array([len(k) for _, _, k in p], np.uint32).tobytes().ljust(a16(4 * n), b"\0")
            blob += b"".join(k.ljust(a16(len(k)), b"\0") for _, _, k in p).ljust(a16(o), b"\0")
            ctypes.memmove(buf_ptr + int(off), blob, len(blob))

    def table_clear(self):
        self.t = {}

    def absorb(self, ptr, nbytes):
        raw = ctypes.string_at(ptr, nbytes)
        magic, n, kb, _ = struct.unpack_from("<4Q", raw, 0)
        assert magic == MAGIC
        rows = ctypes.c_uint64(0)
        assert self.api.load_library().vfb_chunk_rows(raw, nbytes, ctypes.byref(rows)) == 0 and rows.value == n
        o = 32
        hashes = np.frombuffer(raw, np.uint64, n, o); o += a16(8 * n)
        counts = np.frombuffer(raw, np.uint64, n, o); o += a16(8 * n)
        koff = np.frombuffer(raw, np.uint64, n, o); o += a16(8 * n)
        klen = np.frombuffer(raw, np.uint32, n, o); o += a16(4 * n)
        for i in range(n):
            k = raw[o + int(koff[i]):o + int(koff[i]) + int(klen[i])]
            assert self.api.hash_key(k) == int(hashes[i])
            self.add(k, int(counts[i]))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vfind_b200 import api
    from vfind_b200.distributed import gather_table, merge_tables
    rng = random.Random(1234)            # same stream on every rank; each takes its shard
    keys = [bytes(rng.choice(b"ACDEFGHIKLMNPQRSTVWY*X") for _ in range(rng.randrange(1, 40))) for _ in range(400)]
    draws = [rng.choice(keys) for _ in range(6000)]
    tab = HostTable(api)
    for k in draws[rank::world]:
        tab.add(k)
    if rank == 1:
        tab.table_clear() if False else None
    merge_tables(tab, device=torch.device("cpu"))
    owned = dict(tab.t)
    for k in owned:                       # every key sits on its owner and only there
        assert api.key_owner(api.hash_key(k), world) == rank
    ks = sorted(owned)
    data = np.frombuffer(b"".join(ks), np.uint8)
    offs = np.concatenate([[0], np.cumsum([len(k) for k in ks])]).astype(np.uint64)
    cnts = np.array([owned[k] for k in ks], np.uint64)
    g = gather_table(offs, data, cnts)
    if rank == 0:
        o, d, c = g
        raw = d.tobytes()
        merged = {raw[int(o[i]):int(o[i + 1])]: int(c[i]) for i in range(len(c))}
        want = {}
        for k in draws:
            want[k] = want.get(k, 0) + 1
        q.put(merged == want)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_merge_tables_gloo(world):
    from vfind_b200 import build
    build.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_empty_rank_and_empty_table():
    # a rank with nothing to send and a chunk with zero rows are both legal
    sys.path.insert(0, ROOT)
    from vfind_b200 import api, build
    build.build()
    t = HostTable(api)
    sizes = t.partition_sizes(4)
    assert sizes == [32, 32, 32, 32]
    buf = (ctypes.c_uint8 * sum(sizes))()
    t.partition_fill(4, ctypes.addressof(buf), np.cumsum([0] + sizes[:-1]))
    t2 = HostTable(api)
    for i in range(4):
        t2.absorb(ctypes.addressof(buf) + 32 * i, 32)
    assert t2.t == {}
