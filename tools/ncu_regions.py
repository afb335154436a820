#!/usr/bin/env python
"""Share of executed warp instructions per device function (by source line range)."""
import collections, csv, io, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = hd = None
per_line = collections.defaultdict(lambda: [0, 0, 0])
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if "Instructions Executed" in r:
        hd = r; li, ii, si, ti = r.index("Line No"), r.index("Instructions Executed"), r.index("# Samples"), r.index("Thread Instructions Executed"); continue
    if hd and len(r) == len(hd) and r[li].isdigit():
        try:
            a = per_line[(cur, int(r[li]))]; a[0] += int(r[ii]); a[1] += int(r[si]); a[2] += int(r[ti])
        except ValueError:
            pass
# function ranges: lines starting a definition at column 0
def ranges(path):
    starts = []
    for n, line in enumerate(open(path), 1):
        m = re.match(r"^(?:static\s+)?(?:__device__|__global__|RL_HD|inline|template|struct Rng)\b.*?(\w+)\s*\(", line) or \
            re.match(r"^(?:trace_kernel|debug_\w+_kernel)\s*\(", line)
        if re.match(r"^struct Rng", line): starts.append((n, "Rng")); continue
        if m and not line.startswith(" "):
            name = re.findall(r"(\w+)\s*\(", line)
            starts.append((n, name[-1] if len(name) == 1 else name[0] if name[0] not in ("__launch_bounds__",) else name[-1]))
    return starts
files = {f: ranges(os.path.join(ROOT, "robigo-luculenta_b200", "csrc", f)) for f in ("rl_device.cuh", "rl_math.cuh", "rl_kernels.cu")}
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for (f, l), v in per_line.items():
    name = "?"
    for n, nm in files.get(f, []):
        if n <= l: name = nm
    key = f"{f}:{name}" if f in files else f
    agg[key][0] += v[0]; agg[key][1] += v[1]; agg[key][2] += v[2]; tot += v[0]
tots = sum(v[1] for v in agg.values()) or 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"{100*v[0]/tot:6.2f}% inst {100*v[1]/tots:6.2f}% samples lanes {v[2]/max(v[0],1):5.1f}  {k}")
