#!/usr/bin/env python
"""Sweep of the splat kernel's (K2) staging geometry -- stages in flight, blocks per SM
(environment knobs of launch_splat) -- on 2^25 records with the L2 flushed, as bench.py
times it.  Also checks every geometry against the default's image."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PROBE = os.path.join(ROOT, "build", "librl_no_red.so")
if "--build-no-red" in sys.argv:     # here, no GPU needed: a private copy of the library without the reductions
    import subprocess
    import __graft_entry__ as entry
    os.makedirs(os.path.dirname(PROBE), exist_ok=True)
    subprocess.run([entry._nvcc()] + entry.NVCC_FLAGS + ["-DRL_PROBE_NO_RED", "-shared", "-o", PROBE]
                   + [os.path.join(entry.PKG_DIR, f) for f in entry.SOURCES], check=True, cwd=entry.PKG_DIR)
    sys.exit(0)
if "--no-red" in sys.argv:
    os.environ["RL_B200_LIB"] = PROBE
import numpy as np
import torch
import __graft_entry__ as entry
pkg = entry.load_package()
W = H = 1024
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
scene = pkg.Scene(pkg.SceneBuilder(pkg.SCENE_C2))
n = 1 << 25
tr = pkg.TraceUnit(0, W, H, seed=0x5EED, batch=n); tr.set_stream(side.cuda_stream)
plot = pkg.PlotUnit(0, W, H); plot.set_stream(side.cuda_stream)
rec = tr.render_range(scene, 0, n, download=True)
n_lit = int(np.count_nonzero(rec["probability"])); del rec
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def time_ms(fn, reps=7):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        flush.fill_(1); flush[: 192 << 20].sum()
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

ref = None
for stages, per_sm in ((4, 0), (2, 0), (3, 0), (6, 0), (8, 0), (12, 0), (4, 1), (2, 2), (2, 3)):
    os.environ["RL_SPLAT_STAGES"] = str(stages)
    os.environ["RL_SPLAT_BLOCKS_PER_SM"] = str(per_sm)
    plot.clear(); plot.plot(tr); img = plot.download()
    if ref is None:
        ref = img
    err = float(np.abs(img - ref).max() / max(1e-30, np.abs(ref).max()))
    ms = time_ms(lambda: plot.plot(tr))
    print(json.dumps({"stages": stages, "blocks_per_sm": per_sm or "max", "ms": round(ms, 4),
                      "algorithmic_GBps": round((n * 16 + n_lit * 48) / (ms * 1e-3) / 1e9, 1),
                      "record_stream_GBps": round(n * 16 / (ms * 1e-3) / 1e9, 1), "rel_err_vs_default": err}), flush=True)
