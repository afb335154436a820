#!/usr/bin/env python
"""Gather kernel (K3) at 4096^2 against plain device copies of the same byte count, L2 flushed."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
pkg = entry.load_package()
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
gw = 4096
gp, gg = pkg.PlotUnit(0, gw, gw), pkg.GatherUnit(gw, gw)
gp.set_stream(side.cuda_stream); gg.set_stream(side.cuda_stream)
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

real = gw * gw * 80
for bps in (8, 32, 48, 64, 96, 128, 256):
    os.environ["RL_GATHER_BLOCKS_PER_SM"] = str(bps)
    ms = time_ms(lambda: gg.accumulate(gp, clear=True))
    print(json.dumps({"kernel": "gather+clear", "blocks_per_sm": bps, "ms": round(ms, 4),
                      "algorithmic_GBps": round(gw * gw * 72 / ms / 1e6, 1), "moved_GBps": round(real / ms / 1e6, 1)}), flush=True)
# the same bytes as a plain copy (half read, half written) and as a pure read
src = torch.empty(real // 2, dtype=torch.uint8, device="cuda"); dst = torch.empty_like(src)
ms = time_ms(lambda: dst.copy_(src))
print(json.dumps({"kernel": "torch copy_ of the same bytes", "ms": round(ms, 4), "moved_GBps": round(real / ms / 1e6, 1)}))
big = torch.empty(real // 4, dtype=torch.float32, device="cuda")
ms = time_ms(lambda: big.sum())
print(json.dumps({"kernel": "torch sum (pure read) of the same bytes", "ms": round(ms, 4), "moved_GBps": round(real / ms / 1e6, 1)}))
