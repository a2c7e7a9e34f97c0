#!/usr/bin/env python
"""Per-kernel source view of an ncu report: hottest CUDA-C lines and SASS instructions.
    python tools/ncu_src.py REP KERNEL_REGEX [--top N] [--sass]   (needs -lineinfo, --import-source on)
"""
import argparse
import csv
import io
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("kernel")
ap.add_argument("--top", type=int, default=25)
ap.add_argument("--sass", action="store_true", help="list every SASS instruction with >= --min-share of the executed instructions")
ap.add_argument("--min-share", type=float, default=0.004)
ap.add_argument("--launch", type=int, default=0, help="which matching launch (0 = first)")
a = ap.parse_args()
raw = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + a.kernel],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
if not heads:
    raise SystemExit("no source page for " + a.kernel)
# one section per source file of a launch; a launch's sections are consecutive and a new launch
# starts when a file path repeats
launches, cur_files = [[]], set()
for h in heads:
    f = rows[h - 2][1] if h >= 2 else "?"
    if f in cur_files:
        launches.append([]); cur_files = set()
    cur_files.add(f); launches[-1].append(h)
secs = launches[min(a.launch, len(launches) - 1)]
H = rows[secs[0]]
iS, iE, iT = H.index("# Samples"), H.index("Instructions Executed"), H.index("Thread Instructions Executed")
print("kernel:", rows[secs[0] - 1][1], " files:", len(secs))
lines, sass = [], []
for h0 in secs:
    nxt = [x for x in heads if x > h0]
    end = nxt[0] - 2 if nxt else len(rows)
    fname = rows[h0 - 2][1].split("/")[-1] if h0 >= 2 else "?"
    cur = None
    for r in rows[h0 + 1:end]:
        if len(r) <= iT or r[2] == "...":
            continue
        try:
            s, e, t = int(r[iS]), int(r[iE]), int(r[iT])
        except ValueError:
            continue
        if r[0]:
            cur = (fname[:10] + ":" + r[0], r[1].strip())
            lines.append((s, e, t, cur[0], r[1].strip()))
        else:
            sass.append((r[2], s, e, t, r[3].strip(), cur[0] if cur else "?"))
# the same address can appear under several source lines (inlining): keep each address once
seen, uniq = set(), []
for x in sass:
    if x[0] in seen:
        continue
    seen.add(x[0]); uniq.append(x)
totE = sum(x[2] for x in uniq); totS = sum(x[1] for x in uniq); totT = sum(x[3] for x in uniq)
print("warp instructions executed: %d   thread instructions: %d   samples: %d" % (totE, totT, totS))
print("-- hottest source lines (samples%, warp-inst%)")
for s, e, t, ln, src in sorted(lines, key=lambda x: -x[0])[:a.top]:
    print("  %5.1f%% %5.1f%%  %-16s %s" % (100.0 * s / max(1, totS), 100.0 * e / max(1, totE), ln, src[:110]))
if a.sass:
    print("-- SASS in address order (>= %.1f%% of executed warp instructions or of samples)" % (100 * a.min_share))
    for ad, s, e, t, txt, ln in sorted(uniq, key=lambda x: x[0]):
        if e >= a.min_share * totE or s >= a.min_share * totS:
            print("  %s  inst %5.2f%%  smp %5.2f%%  L%-4s %s" % (ad[-5:], 100.0 * e / totE, 100.0 * s / max(1, totS), ln, txt))
