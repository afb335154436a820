#!/usr/bin/env python
"""Per-source-line table of an .ncu-rep (warp instructions, lanes, stall samples), in line order,
for one file and line range.  Usage: python tools/ncu_lines.py REP rl_device.cuh 520 790"""
import collections, csv, io, os, subprocess, sys
rep, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = hd = None
per = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if "Instructions Executed" in r:
        hd = r; li, ii, si, ti = r.index("Line No"), r.index("Instructions Executed"), r.index("# Samples"), r.index("Thread Instructions Executed"); continue
    if hd and len(r) == len(hd) and r[li].isdigit():
        try:
            v = (int(r[ii]), int(r[si]), int(r[ti]))
        except ValueError:
            continue
        tot += v[0]
        a = per[(cur, int(r[li]))]; a[0] += v[0]; a[1] += v[1]; a[2] += v[2]
src = open(os.path.join(ROOT, "robigo-luculenta_b200", "csrc", fname)).read().splitlines()
samples = sum(v[1] for v in per.values()) or 1
acc = 0
for n in range(lo, hi + 1):
    v = per.get((fname, n))
    if not v or not v[0]:
        continue
    acc += v[0]
    print(f"{n:5d} {100*v[0]/tot:5.2f}% inst {100*v[1]/samples:5.2f}% smp lanes {v[2]/v[0]:4.1f}  {src[n-1].strip()[:100]}")
print(f"range total {100*acc/tot:.2f}% of {tot} warp instructions")
