#!/usr/bin/env python
"""Mrays/s of the fused trace+splat path on the BASELINE.json configs that are
parity cases rather than bench lines (C1, C3, C4), for DESIGN.md.  Not a bench
contract line: one warm-up, CUDA events around a few launches."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry

pkg = entry.load_package()
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
out = {}
for name, which, w, h, n in (("C1 sphere+plane 256^2", 1, 256, 256, 1 << 24), ("C2 built-in 1024^2", 2, 1024, 1024, 1 << 26),
                             ("C3 prism 1024^2", 3, 1024, 1024, 1 << 26), ("C4 4096 spheres 2048^2", 4, 2048, 2048, 1 << 24),
                             ("C5 built-in 4096^2 (one GPU's share)", 2, 4096, 4096, 1 << 26)):
    if os.environ.get("RL_RATES_ONLY") and not name.startswith(os.environ["RL_RATES_ONLY"]):
        continue
    sc = pkg.Scene(pkg.SceneBuilder(which))
    tu = pkg.TraceUnit(0, w, h, seed=0x5EED, batch=n); pl = pkg.PlotUnit(0, w, h)
    tu.set_stream(side.cuda_stream); pl.set_stream(side.cuda_stream)
    tu.render_fused(sc, pl, 0, n); torch.cuda.synchronize()
    r0 = tu.ray_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(3):
        tu.render_fused(sc, pl, (k + 1) * n, n)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b); rays = tu.ray_count() - r0
    out[name] = {"mrays_per_s": round(rays / ms / 1e3, 1), "mphotons_per_s": round(3 * n / ms / 1e3, 1),
                 "rays_per_photon": round(rays / (3 * n), 3)}
print(json.dumps(out))
