#!/usr/bin/env python
"""Benchmark of the variant-recovery hot path (BASELINE.json metric: reads/sec, 250-bp synthetic).

    python bench.py --gpus N --steps K --warmup W            # this build, N GPUs (torchrun for N>1)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port)

Workload (config.workload): BASELINE config #3 — 100 M synthetic 250-bp merged amplicon reads
per GPU, 20-nt adapters, 30 % of adapters mutated with at least one indel (alignment-heavy),
seed 1003; weak scaling: every rank gets its own 100 M-read shard of the
same stream.  One step = one pass of the whole hot path (K1 scan -> K2 DP -> K3 translate ->
K4 count, then the NCCL table merge when N > 1) over that shard.

Prints ONE JSON line (rank 0).  `value` times the path with reads resident in HBM; `e2e`
times the same reads through the C ABI from pinned HOST buffers (H2D copies and the D2H of
the result table inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "C3: 250-bp synthetic merged amplicon reads, 20-nt adapters, 30% adapters mutated (>=1 indel), seed 1003"
ALU_OPS_PER_CELL = 4    # VIADDMNMX x2 + VIMNMX3 + LOP3 on the INT32 ALU pipe (3 IMAD ride the FMA pipe)
FILTER_ALU_OPS_PER_COL = 11   # k2_filter: 7 LOP3 + PRMT + 2 LEA.HI (score) + VIMNMX per read column


def c3_cfg(api):
    return api.synth_cfg(seed=1003, read_len=250, adapter_len=20, region_len=198, n_variants=1000000,
                         zipf=1, p_err=0.30, indel=0.5, force_indel=1, frameshift=0.05,
                         noise=0.001, n_rate=1e-4)


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML every
    10 ms from a thread (the timed region of the resident leg is ~0.15 s), `nvidia-smi -lms` if NVML is not importable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None
        self.nv = None
        self.stop_flag = threading.Event()
        self.t = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.nv = pynvml
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nv
        bits = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append(["", sm, self.mx, pw, ""] + ["Active" if rs & b else "Not Active" for _, b in bits])
            except Exception:
                pass
            self.stop_flag.wait(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.stop_flag.set()
            self.t.join(timeout=2)
        elif not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if str(v).lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvml, 10 ms" if self.nv is not None else "nvidia-smi -lms 50"}


FULL_AFFINITY = None


def bind_to_gpu_cpus(gpu_index):
    """Run this rank on the CPUs NVML names as local to its GPU (the socket its PCIe root hangs off), so that the
    pinned host buffers of the e2e leg are allocated on that NUMA node.  Best effort: containers may forbid it."""
    if os.environ.get("VFB_BENCH_NO_AFFINITY"):
        return "off"
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        global FULL_AFFINITY
        if FULL_AFFINITY is None:
            FULL_AFFINITY = os.sched_getaffinity(0)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = sorted(os.sched_getaffinity(0))
        return "nvml ideal CPUs of GPU %d: %d of %d allowed (%d..%d)" % (gpu_index, len(after), before, after[0], after[-1])
    except Exception as e:
        return "unchanged (%s)" % type(e).__name__


def oracle_sample(api, oracle, cfg, adapters, text, spans, threads, target_s):
    """Time the CPU oracle (all host threads) on a bounded prefix of the workload."""
    n_all = len(spans)
    probe = min(n_all, 20000 * threads)
    p = oracle.make_params(adapters)
    t0 = time.perf_counter()
    oracle.process_reads(p, text, spans["off"][:probe], spans["len"][:probe], n_threads=threads)
    dt = time.perf_counter() - t0
    n = int(min(n_all, max(probe, probe * target_s / max(dt, 1e-3))))
    t0 = time.perf_counter()
    table, _, cells = oracle.process_reads(p, text, spans["off"][:n], spans["len"][:n], n_threads=threads)
    dt = time.perf_counter() - t0
    return n, dt, cells, len(table)


def run_reference(args):
    """--impl reference: the reference's CPU path (the oracle port — the Rust/parasail
    reference cannot be built in this image), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from vfind_b200 import api
    oracle.build()
    cfg = c3_cfg(api)
    adapters = api.synth_adapters(cfg)
    threads = os.cpu_count() or 1
    n = args.ref_reads
    text, spans = api.synth_host(cfg, 0, n)
    p = oracle.make_params(adapters)
    times = []
    cells = 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, _, cells = oracle.process_reads(p, text, spans["off"], spans["len"], n_threads=threads)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    line = {
        "impl": "reference", "metric": "reads/sec", "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "reads_per_step": n, "read_len": cfg.read_len,
                   "adapter_len": cfg.adapter_len,
                   "note": "CPU oracle port of src/lib.rs (parasail/Rust cannot be built here); bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": "port",
                         "sample": "%d reads of the C3 stream per step" % n,
                         "gcups": cells / (total / len(times)) / 1e9},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else libraries print on fd 1 (NCCL's version banner,
    for one) was sent to stderr by main()."""
    os.write(REAL_STDOUT, (json.dumps(line) + "\n").encode())


REAL_STDOUT = 1


def main():
    global REAL_STDOUT
    sys.stdout.flush()
    REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=100_000_000, help="reads per GPU per step")
    ap.add_argument("--chunk-reads", type=int, default=12_500_000)
    ap.add_argument("--e2e-reads", type=int, default=-1, help="reads per e2e step (-1 = same as --reads)")
    ap.add_argument("--ref-reads", type=int, default=1_000_000, help="reads per step of --impl reference")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ingest-reads", type=int, default=4_000_000,
                    help="reads in the block-gzip FASTQ file of the ingest leg (N=1 only; 0 = skip)")
    ap.add_argument("--table-hint", type=int, default=40_000_000, help="expected distinct variants per GPU (table capacity hint)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from vfind_b200 import api
    from vfind_b200.distributed import merge_tables

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_cpus(local)          # before any pinned allocation: first touch decides the NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node = --gpus"

    cfg = c3_cfg(api)
    adapters = api.synth_adapters(cfg)
    R, L = args.reads, cfg.read_len
    # ---- synthetic shard of this rank, resident in HBM, in <= 4 GiB chunks
    chunks = []
    first = rank * R
    done = 0
    while done < R:
        n = min(args.chunk_reads, R - done)
        t = torch.empty(n * L, dtype=torch.uint8, device=dev)
        s = torch.empty(n * 2, dtype=torch.int32, device=dev)
        api.synth_device(cfg, first + done, n, t.data_ptr(), s.data_ptr(), local)
        chunks.append((t, s, n))
        done += n
    torch.cuda.synchronize()

    stream = torch.cuda.Stream(device=dev)
    ctx = api.Context(adapters, device=local, table_capacity_hint=min(R, args.table_hint), batch_reads=args.chunk_reads)
    ctx.set_compute_stream(stream.cuda_stream)

    merge_events = []

    def step_device():
        ctx.table_clear()
        for t, s, n in chunks:
            ctx.submit_device(t.data_ptr(), t.numel(), s.data_ptr(), n)
        if world > 1:
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record(stream)
            merge_tables(ctx, device=dev)
            m1.record(stream)
            merge_events.append((m0, m1))

    def barrier():
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
        barrier()
        merge_check = None
        if world > 1:
            # untimed sanity check of the merge: partitions are disjoint by construction, so the
            # counts summed over ranks must equal the reads counted over ranks before the merge
            ctx.table_clear()
            for t, s, n in chunks:
                ctx.submit_device(t.data_ptr(), t.numel(), s.data_ptr(), n)
            before = torch.tensor([ctx.stats()["counted"]], dtype=torch.int64, device=dev)
            merge_tables(ctx, device=dev)
            _, _, cnts = ctx.finish_arrays(copy=False)
            after = torch.tensor([int(cnts.sum()), len(cnts)], dtype=torch.int64, device=dev)
            dist.all_reduce(before)
            dist.all_reduce(after)
            merge_check = {"counted_before": int(before[0]), "counted_after": int(after[0]),
                           "rows_total": int(after[1]), "ok": int(before[0]) == int(after[0])}
            barrier()
        ctx.reset()
        ctx.set_profiling(True)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        merge_events.clear()
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        merge_ms = sum(a.elapsed_time(b) for a, b in merge_events) / max(1, len(merge_events)) if merge_events else 0.0
        clocks = sampler.stop() if rank == 0 else None
        st = ctx.stats()
        ctx.set_profiling(False)
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_max = float(tms.item())
        value = world * R * args.steps / (ms_max * 1e-3)

        # ---- e2e: the same reads from pinned host memory through vfb_submit_host + vfb_finish
        e2e = None
        if not args.no_e2e:
            er = R if args.e2e_reads < 0 else min(args.e2e_reads, R)
            host = []
            got = 0
            for t, s, n in chunks:
                if got >= er:
                    break
                m = min(n, er - got)
                ht = torch.empty(m * L, dtype=torch.uint8, pin_memory=True)
                hs = torch.empty(m * 2, dtype=torch.int32, pin_memory=True)
                ht.copy_(t[:m * L]); hs.copy_(s[:m * 2])
                host.append((ht, hs, m))
                got += m
            torch.cuda.synchronize()
            rows = 0

            def step_e2e():
                nonlocal rows
                ctx.table_clear()
                for ht, hs, m in host:
                    ctx.submit_host_ptr(ht.data_ptr(), ht.numel(), hs.data_ptr(), m)
                if world > 1:
                    merge_tables(ctx, device=dev)
                offsets, data, counts = ctx.finish_arrays(copy=False)
                rows = len(counts)
                return offsets.nbytes + data.nbytes + counts.nbytes

            step_e2e()
            barrier()
            # the link itself: one pinned chunk copied alone (best of 3) — what `e2e` is bounded by
            link_gbs = 0.0
            if host:
                ht0 = host[0][0]
                scratch = torch.empty(ht0.numel(), dtype=torch.uint8, device=dev)
                for _ in range(3):
                    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    c0.record(stream)
                    scratch.copy_(ht0, non_blocking=True)
                    c1.record(stream)
                    stream.synchronize()
                    link_gbs = max(link_gbs, ht0.numel() / (c0.elapsed_time(c1) * 1e-3) / 1e9)
                del scratch
            d2h = 0
            link0 = ctx.stats()["h2d_bytes"]
            t0 = time.perf_counter()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            esteps = max(1, min(args.steps, 3))
            for _ in range(esteps):
                d2h = step_e2e()
            f1.record(stream)
            barrier()
            wall = time.perf_counter() - t0
            link_bytes = (ctx.stats()["h2d_bytes"] - link0) / esteps
            ems = torch.tensor([max(f0.elapsed_time(f1), wall * 1e3)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ems, op=dist.ReduceOp.MAX)
            e2e = {"value": world * got * esteps / (float(ems.item()) * 1e-3), "unit": "reads/s",
                   "h2d_bytes_per_step": int(sum(h[0].numel() + h[1].numel() * 4 for h in host)),
                   "d2h_bytes_per_step": int(d2h), "reads_per_step": got, "steps": esteps,
                   "table_rows": rows, "ms_per_step": float(ems.item()) / esteps,
                   "h2d_gbs_achieved": world * sum(h[0].numel() + h[1].numel() * 4 for h in host) * esteps / (float(ems.item()) * 1e-3) / 1e9,
                   "h2d_gbs_link_alone": link_gbs,
                   "h2d_bytes_on_link_per_step": int(link_bytes),
                   "host_pack": os.environ.get("VFB_HOST_PACK", "off"),
                   "bound": "host DRAM -> PCIe: the caller hands over 258 B per read in pinned memory; reading it out of host "
                            "DRAM caps at ~48 GB/s on this box whether the copy engine or host threads (VFB_HOST_PACK, "
                            "2-bit packing) read it"}
            del host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K2 DP, INT32-ALU bound) and of the HBM-bound scan
    alu_gops, dual_gops = api.measure_int_peak(local)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
        if traffic.get("reads_per_launch") != min(args.chunk_reads, R):
            traffic = {}                    # captured at another batch size: not comparable per launch
    except (OSError, ValueError):
        pass
    dp_s = st["ms_dp"] * 1e-3
    gcups = st["dp_cells"] / dp_s / 1e9 if dp_s > 0 else 0.0      # effective: full A x L matrices / DP stage time
    windowed = st["dp_kernel_kind"] == 3
    if windowed:
        # the DP stage = k2_filter (Myers bit-vector, every column of every aligned read) + k2_dp_window
        # (packed Gotoh cells on the flagged windows only) + k2_resolve.  Roofline of the window kernel:
        # cells it actually evaluated x 4 ALU-pipe ops, against the measured ALU-pipe peak.
        win_s = st["ms_dp_window"] * 1e-3
        kernel_gcups = st["dp_cells_computed"] / win_s / 1e9 if win_s > 0 else 0.0
        dp_kernel, dp_ms = "k2_dp_window", st["ms_dp_window"]
    else:
        kernel_gcups, dp_kernel, dp_ms = gcups, "k2_dp_packed", st["ms_dp"]
    dp_achieved = kernel_gcups * ALU_OPS_PER_CELL     # G lane-ops/s on the ALU pipe
    scan_bytes = st["reads"] * (L + 12)               # N*(L+12)
    scan_gbs = scan_bytes / (st["ms_scan"] * 1e-3) / 1e9 if st["ms_scan"] > 0 else 0.0
    key_bytes = st["counted"] * args.steps * (cfg.region_len + cfg.region_len // 3 + 8)
    roofline = {"bound": "int32", "kernel": dp_kernel, "achieved": dp_achieved, "peak": alu_gops,
                "unit": "Gop/s", "frac": dp_achieved / alu_gops if alu_gops else None,
                "traffic": traffic.get(dp_kernel), "traffic_unit": "DRAM bytes per launch (ncu, profiles/r01_ncu_traffic.json)",
                "gcups": kernel_gcups, "alu_ops_per_cell": ALU_OPS_PER_CELL,
                "peak_source": "vfb_measure_int_peak (VIADDMNMX stream, measured in this run); ALU+FMA dual-issue peak %.0f Gop/s" % dual_gops,
                "share_of_step": dp_ms / st["ms_total"] if st["ms_total"] else None}
    roofline_filter = None
    if windowed and st["ms_dp_filter"] > 0:
        # Myers column: 7 LOP3 + 1 PRMT + 3 score/min ops on the ALU pipe (3 IMAD adds/shifts on the FMA pipe, 1 LDS)
        cols = st["dp_cells"] / cfg.adapter_len       # one column per read base of every aligned read
        f_ach = cols * FILTER_ALU_OPS_PER_COL / (st["ms_dp_filter"] * 1e-3) / 1e9
        roofline_filter = {"bound": "int32", "kernel": "k2_filter", "achieved": f_ach, "peak": alu_gops, "unit": "Gop/s",
                           "frac": f_ach / alu_gops if alu_gops else None, "traffic": traffic.get("k2_filter"),
                           "columns_per_s": cols / (st["ms_dp_filter"] * 1e-3), "alu_ops_per_column": FILTER_ALU_OPS_PER_COL,
                           "share_of_step": st["ms_dp_filter"] / st["ms_total"] if st["ms_total"] else None}
    roofline_hbm = {"bound": "hbm", "kernel": "k1_scan", "achieved": scan_gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": scan_gbs / hbm_peak, "traffic": traffic.get("k1_scan"),
                    "algorithmic_bytes_per_launch": min(args.chunk_reads, R) * (L + 12), "peak_source": hbm_src,
                    "share_of_step": st["ms_scan"] / st["ms_total"] if st["ms_total"] else None}
    stages = {k: st[k] / args.steps for k in ("ms_scan", "ms_worklist", "ms_dp", "ms_dp_filter", "ms_dp_window",
                                              "ms_translate", "ms_count", "ms_total")}
    stages["translate_gbs"] = key_bytes / (st["ms_translate"] * 1e-3) / 1e9 if st["ms_translate"] > 0 else 0.0

    cpu = None
    if world == 1 and not args.no_cpu:
        import oracle
        oracle.build()
        if FULL_AFFINITY is not None:
            os.sched_setaffinity(0, FULL_AFFINITY)      # the CPU baseline gets every host core
        threads = os.cpu_count() or 1
        t, s, n = chunks[0]
        m = min(n, 4_000_000)
        text = t[:m * L].cpu().numpy()
        spans = s[:m * 2].cpu().numpy().view(np.uint32).reshape(m, 2)
        sp = np.zeros(m, dtype=api.SPAN_DTYPE)
        sp["off"], sp["len"] = spans[:, 0], spans[:, 1]
        cn, cdt, ccells, _ = oracle_sample(api, oracle, cfg, adapters, text, sp, threads, args.cpu_seconds)
        cpu = {"value": cn / cdt, "unit": "reads/s", "cores": threads, "kind": "port",
               "sample": "first %d reads of the same C3 stream, %.1f s" % (cn, cdt),
               "gcups": ccells / cdt / 1e9}

    # ---- ingest leg (untimed by the contract; reported beside it): the user-facing call, vfind.find_variants(path),
    # on a block-gzip FASTQ file of the same stream — file read, H2D of the compressed members, GPU inflate, GPU
    # FASTQ parse, K1..K4, table hand-off; whole call, best of 3 after one warm-up call
    ingest = None
    if world == 1 and args.ingest_reads > 0:
        bind_to_gpu_cpus(local)
        # A user's process does not sit on the resident legs' 30 GB of inputs and tables: let them go before timing whole
        # find_variants calls (with them in place every call's cudaMalloc / cudaFree takes several times longer and the
        # leg measures the driver's allocator, see profiles/README.md).  Nothing below uses ctx or chunks.
        try:
            del t, s
        except NameError:
            pass
        chunks.clear()
        ctx.close()
        torch.cuda.empty_cache()
        try:
            import tempfile
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import synth_fastq
            from vfind_b200 import find_variants
            tmp = tempfile.mkdtemp(prefix="vfb_bench_")
            fq = os.path.join(tmp, "c3.fq.gz")
            tb, zb = synth_fastq.write_bgzf_fastq(fq, cfg, args.ingest_reads, api)
            ads = tuple(a.decode() for a in adapters)
            times = []
            rows = 0
            for rep in range(4):
                t0 = time.perf_counter()
                out = find_variants(fq, ads, show_progress=False)
                times.append(time.perf_counter() - t0)
                rows = out.num_rows if hasattr(out, "num_rows") else len(out)
            best = min(times[1:])
            ingest = {"value": args.ingest_reads / best, "unit": "reads/s", "call": "vfind.find_variants(path, adapters)",
                      "input": "block-gzip (BGZF, zlib level 1) FASTQ, %d reads, %.0f MB text, %.0f MB compressed"
                               % (args.ingest_reads, tb / 1e6, zb / 1e6),
                      "seconds_best_of_3": best, "seconds_first_call": times[0], "table_rows": rows,
                      "text_gbs": tb / best / 1e9}
            os.remove(fq)
            # the same call on ONE plain gzip stream (what `gzip` / fastp write): decoded by the parallel
            # host gunzip (pgunzip.cu) on the ingest threads
            n_gz = min(args.ingest_reads, 2_000_000)
            txt = os.path.join(tmp, "c3_plain.fq")
            gz = txt + ".gz"
            tb2 = synth_fastq.write_fastq(txt, cfg, 0, n_gz, api)
            with open(gz, "wb") as g:
                subprocess.check_call(["gzip", "-1", "-c", txt], stdout=g)
            os.remove(txt)
            times = []
            for rep in range(3):
                t0 = time.perf_counter()
                find_variants(gz, ads, show_progress=False)
                times.append(time.perf_counter() - t0)
            ingest["plain_gzip"] = {"value": n_gz / min(times), "unit": "reads/s",
                                    "input": "single gzip stream (gzip -1), %d reads, %.0f MB text, %.0f MB compressed"
                                             % (n_gz, tb2 / 1e6, os.path.getsize(gz) / 1e6),
                                    "seconds_best_of_3": min(times), "host_threads": min(os.cpu_count() or 1, 32)}
            os.remove(gz)
            os.rmdir(tmp)
        except Exception as e:          # the ingest leg never fails the bench line
            ingest = {"error": repr(e)}

    line = {
        "metric": "reads/sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reads_per_gpu": R, "read_len": L, "adapter_len": cfg.adapter_len,
                   "region_len": cfg.region_len, "library": cfg.n_variants, "thresholds": [0.75, 0.75],
                   "scoring": [3, -2, 5, 2], "l2": "inputs (%.1f GB per GPU) are larger than L2" % (R * L / 1e9),
                   "parallelism": "reads sharded over %d GPU(s), NCCL all-to-all table merge" % world,
                   "cpu_affinity": affinity},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(st["kernel_launches"]),
        "roofline": roofline, "roofline_filter": roofline_filter, "roofline_hbm": roofline_hbm,
        "stages_ms_per_step": stages,
        "dp": {"effective_gcups": gcups, "mode": "windowed (filter + windows)" if windowed else "full matrices",
               "cells_per_step": st["dp_cells"] // args.steps,
               "cells_computed_per_step": st["dp_cells_computed"] // args.steps,
               "windows_per_step": st["dp_windows"] // args.steps,
               "alignments_per_step": (st["dp_prefix"] + st["dp_suffix"]) // args.steps},
        "merge_ms_per_step": merge_ms, "table": {"unique": st["unique"], "counted_per_step": st["counted"], "merge_check": merge_check},
        "cpu_baseline": cpu, "ingest": ingest,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
