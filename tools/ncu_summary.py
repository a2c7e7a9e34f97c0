#!/usr/bin/env python
"""Summarise ncu output for profiles/: a launch list CSV (--metrics gpu__time_duration.sum)
and/or a --set full report (.ncu-rep).  Usage:
    python tools/ncu_summary.py --launches gpurun_out/launches.csv --rep gpurun_out/prof.ncu-rep > profiles/xyz.txt
"""
import argparse
import collections
import csv
import io
import subprocess

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if r[ui] == "ns":
            v /= 1e3
        elif r[ui] == "ms":
            v *= 1e3
        a = agg.setdefault(r[ki][:70], [0.0, 0])
        a[0] += v
        a[1] += 1
    tot = sum(v[0] for v in agg.values())
    print("== launch list (%s): device time per kernel, cold-cache & serialised: compare SHARES" % path)
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
        print("%-72s n=%4d %12.1f us %6.1f%%" % (k, v[1], v[0], 100 * v[0] / tot))


def rep(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    print("== ncu --set full (%s)" % path)
    for r in rows[2:]:
        print("-- kernel:", r[h.index("Kernel Name")][:90])
        for k in KEYS:
            if k in h:
                print("   %-66s %s %s" % (k, r[h.index(k)], units[h.index(k)]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    heads = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
    if not heads:
        return
    h = rows[heads[0]]
    end = heads[1] - 1 if len(heads) > 1 else len(rows)
    idx = {c: i for i, c in enumerate(h)}
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    tot, byop, ex = collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows[heads[0] + 1:end]:
        try:
            s = int(r[idx["# Samples"]])
        except (ValueError, IndexError):
            continue
        parts = r[idx["Source"]].split()
        if not parts:
            continue
        op = (parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]).split(".")[0]
        byop[op] += s
        try:
            ex[op] += int(r[idx["Instructions Executed"]])
        except ValueError:
            pass
        for c in stalls:
            try:
                tot[c] += int(r[idx[c]])
            except ValueError:
                pass
    print("-- warp stall samples (first captured launch):")
    ts = sum(tot.values()) or 1
    for c, v in tot.most_common(10):
        print("   %-28s %8d %5.1f%%" % (c, v, 100 * v / ts))
    print("-- SASS opcode mix (warp instructions executed / stall samples):")
    te, tsam = sum(ex.values()) or 1, sum(byop.values()) or 1
    for op, v in ex.most_common(14):
        print("   %-12s exec %12d %5.1f%%   samples %5.1f%%" % (op, v, 100 * v / te, 100 * byop[op] / tsam))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    a = ap.parse_args()
    if a.launches:
        launches(a.launches)
    if a.rep:
        rep(a.rep)
