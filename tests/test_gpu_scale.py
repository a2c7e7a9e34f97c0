"""GPU tests at BASELINE.json's shapes: oracle parity on samples of every config, and
size-independent properties at full size (where the oracle would take too long).

C2: 10 M x 150 bp, 20-nt adapters, 5 % adapter errors      (full size here)
C3: 250 bp, 30 % adapters with indels                        (parity sample; bench.py runs it at 100 M)
C4: high diversity, uniform library                          (count/merge stress, reduced to 8 M reads / 4 M variants)
C5: 300 bp, 40-nt adapters, thresholds 0.6/0.6               (parity sample)
"""
import numpy as np
import pytest

import oracle
from vfind_b200 import api

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

CFG = {
    "C2": dict(seed=1002, read_len=150, adapter_len=20, region_len=99, n_variants=100000, zipf=1, p_err=0.05,
               indel=0.2, force_indel=0),
    "C3": dict(seed=1003, read_len=250, adapter_len=20, region_len=198, n_variants=1000000, zipf=1, p_err=0.30,
               indel=0.5, force_indel=1),
    "C4": dict(seed=1004, read_len=250, adapter_len=20, region_len=198, n_variants=4000000, zipf=0, p_err=0.05,
               indel=0.2, force_indel=0),
    "C5": dict(seed=1005, read_len=300, adapter_len=40, region_len=210, n_variants=1000000, zipf=1, p_err=0.15,
               indel=0.4, force_indel=0),
}
THR = {"C2": 0.75, "C3": 0.75, "C4": 0.75, "C5": 0.6}


def device_reads(cfg, first, n):
    t = torch.empty(n * cfg.read_len, dtype=torch.uint8, device="cuda")
    s = torch.empty(n * 2, dtype=torch.int32, device="cuda")
    api.synth_device(cfg, first, n, t.data_ptr(), s.data_ptr())
    return t, s


def table_of(ctx):
    o, d, c = ctx.finish_arrays()
    return o, d, c


def as_dict(o, d, c):
    raw = d.tobytes()
    return {raw[int(o[i]):int(o[i + 1])]: int(c[i]) for i in range(len(c))}


@pytest.mark.parametrize("name", ["C2", "C3", "C4", "C5"])
def test_config_sample_matches_oracle(name):
    cfg = api.synth_cfg(**CFG[name])
    n = 150000
    ad = api.synth_adapters(cfg)
    t, s = device_reads(cfg, 5_000_000, n)
    with api.Context(ad, accept_prefix_alignment=THR[name], accept_suffix_alignment=THR[name],
                     diagnostics=True) as ctx:
        ctx.submit_device(t.data_ptr(), t.numel(), s.data_ptr(), n)
        diag = ctx.diag(n)
        got = ctx.finish_dict()
        st = ctx.stats()
    text = t.cpu().numpy()
    sp = s.cpu().numpy().view(np.uint32).reshape(n, 2)
    p = oracle.make_params(ad, accept_prefix_alignment=THR[name], accept_suffix_alignment=THR[name])
    want, odiag, cells = oracle.process_reads(p, text, sp[:, 0].copy(), sp[:, 1].copy(), n_threads=16, want_diag=True)
    for f in diag.dtype.names:
        assert (diag[f] == odiag[f]).all(), f
    assert got == want and st["dp_cells"] == cells
    assert st["dp_kernel_kind"] == 1
    # the default (non-diagnostics) context: windowed DP where its filter applies (20-nt adapters at 0.75)
    with api.Context(ad, accept_prefix_alignment=THR[name], accept_suffix_alignment=THR[name]) as ctx:
        ctx.submit_device(t.data_ptr(), t.numel(), s.data_ptr(), n)
        got2 = ctx.finish_dict()
        st2 = ctx.stats()
    assert got2 == want
    assert st2["dp_kernel_kind"] == (1 if name == "C5" else 3)
    if name != "C5":
        assert st2["dp_cells_computed"] < 0.3 * st2["dp_cells"]


def test_c3_windowed_dp_equals_full_dp_at_size():
    # 12.5 M alignment-heavy reads (one bench.py chunk): the windowed DP and the full DP give the same table
    cfg = api.synth_cfg(**CFG["C3"])
    n = 12_500_000
    ad = api.synth_adapters(cfg)
    t, s = device_reads(cfg, 40_000_000, n)
    tabs = []
    for mode in (0, 1):
        with api.Context(ad, dp_mode=mode, table_capacity_hint=8_000_000) as ctx:
            ctx.submit_device(t.data_ptr(), t.numel(), s.data_ptr(), n)
            o, d, c = table_of(ctx)
            st = ctx.stats()
            assert st["dp_kernel_kind"] == (3 if mode == 0 else 1)
            assert int(c.sum()) == st["counted"]
            tabs.append((as_dict(o, d, c), st["dp_cells"], st["dp_prefix"], st["dp_suffix"]))
    assert tabs[0] == tabs[1]


def test_c2_full_size_properties():
    cfg = api.synth_cfg(**CFG["C2"])
    n = 10_000_000
    ad = api.synth_adapters(cfg)
    t, s = device_reads(cfg, 0, n)
    with api.Context(ad, table_capacity_hint=4_000_000) as ctx:
        ctx.submit_device(t.data_ptr(), t.numel(), s.data_ptr(), n)
        o1, d1, c1 = table_of(ctx)
        st = ctx.stats()
        # every counted read is in exactly one row
        assert int(c1.sum()) == st["counted"] and st["reads"] == n
        assert len(c1) == st["unique"]
        assert 0.80 * n < st["counted"] < n                     # exact-match dominated
        assert st["dp_prefix"] < 0.08 * n
        # all keys are peptides of the expected length and distinct
        lens = np.diff(o1.astype(np.int64))
        assert lens.min() >= 1 and np.bincount(lens).argmax() == cfg.region_len // 3
        one = as_dict(o1, d1, c1)
        assert len(one) == len(c1)
        # counting is additive: the same reads again double every count
        ctx.submit_device(t.data_ptr(), t.numel(), s.data_ptr(), n)
        two = as_dict(*table_of(ctx))
        assert two.keys() == one.keys() and all(two[k] == 2 * v for k, v in one.items())
    # independent of batching and of the initial table size (growth + rehash)
    with api.Context(ad, batch_reads=1_300_000) as ctx:
        ctx.submit_device(t.data_ptr(), t.numel(), s.data_ptr(), n)
        assert as_dict(*table_of(ctx)) == one
    # host path (pinned, H2D inside) gives the same table as the resident path
    ht = torch.empty(t.numel(), dtype=torch.uint8, pin_memory=True).copy_(t)
    hs = torch.empty(s.numel(), dtype=torch.int32, pin_memory=True).copy_(s)
    torch.cuda.synchronize()
    with api.Context(ad, batch_reads=3_000_000) as ctx:
        ctx.submit_host_ptr(ht.data_ptr(), ht.numel(), hs.data_ptr(), n)
        assert as_dict(*table_of(ctx)) == one


def test_c4_high_diversity_partition_roundtrip():
    cfg = api.synth_cfg(**CFG["C4"])
    n = 8_000_000
    ad = api.synth_adapters(cfg)
    t, s = device_reads(cfg, 0, n)
    with api.Context(ad) as src:
        src.submit_device(t.data_ptr(), t.numel(), s.data_ptr(), n)
        o, d, c = table_of(src)
        st = src.stats()
        assert st["unique"] > 2_000_000 and int(c.sum()) == st["counted"]
        n_parts = 8
        sizes = src.partition_sizes(n_parts)
        offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
        buf = torch.empty(int(sum(sizes)), dtype=torch.uint8, device="cuda")
        src.partition_fill(n_parts, buf.data_ptr(), offs)
        rows = total = 0
        sums = np.zeros(2, dtype=np.uint64)
        with api.Context(ad) as dst:
            for p in range(n_parts):
                dst.table_clear()
                dst.absorb(buf.data_ptr() + int(offs[p]), sizes[p])
                po, pd, pc = table_of(dst)
                rows += len(pc)
                total += int(pc.sum())
                # a checksum of the keys: sum of (hash * count) is partition-independent
                raw = pd.tobytes()
                for i in range(0, len(pc), max(1, len(pc) // 2000)):
                    k = raw[int(po[i]):int(po[i + 1])]
                    assert api.key_owner(api.hash_key(k), n_parts) == p
        assert rows == len(c) and total == int(c.sum())
