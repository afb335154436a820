#!/usr/bin/env python
"""The launches the committed ncu captures are taken from (run under ncu, see profiles/):
  0  warm-up
  1  fused trace+splat, built-in scene, 1024^2, 2^24 photons: the bench configuration (768 x 1)
  2  one reference batch (524 288 photons) into records: the small-launch configuration (256 x 3)
  3  splat of those records
  4  trace of 2^25 photons into records (512 MiB)
  5  splat of 2^25 records: the configuration bench.py times for K2
  6, 7  gather + clear of a 4096^2 frame: the configuration bench.py times for K3
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
# RL_PROFILE_SCENE=4 RL_PROFILE_CANVAS=2048: the same launches on another BASELINE config
W = H = int(os.environ.get("RL_PROFILE_CANVAS", "1024"))
scene = pkg.Scene(pkg.SceneBuilder(int(os.environ.get("RL_PROFILE_SCENE", str(pkg.SCENE_C2)))))
tu = pkg.TraceUnit(0, W, H, seed=0x5EED, batch=1 << 24)
pl = pkg.PlotUnit(0, W, H)
tu.render_fused(scene, pl, 0, 1 << 22)
tu.sync()
tu.render_fused(scene, pl, 1 << 24, 1 << 24)
tu.sync()
if os.environ.get("RL_PROFILE_TRACE_ONLY"):
    print("rays", tu.ray_count())
    sys.exit(0)
tu.render_range(scene, 1 << 26, 524288, download=False)
tu.sync()
pl.plot(tu)
pl.sync()
big = pkg.TraceUnit(1, W, H, seed=0x5EED, batch=1 << 25)
big.render_range(scene, 0, 1 << 25, download=False)
big.sync()
pl.plot(big)
pl.sync()
del big
gp, gg = pkg.PlotUnit(2, 4096, 4096), pkg.GatherUnit(4096, 4096)
for _ in range(2):
    gg.accumulate(gp, clear=True)
gg.sync()
print("rays", tu.ray_count())
