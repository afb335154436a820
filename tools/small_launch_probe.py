#!/usr/bin/env python
"""Rate of the reference's 524 288-photon batches (trace_unit.rs:67) as a function of how the
launches share the GPU: U units on their own streams, launches interleaved, each launch asking
for 1/min(U, share_max) of the SMs' block slots (launch_trace's small-launch rule) -- i.e. of the
photons per thread a block lives for.  Records stay on the device; nothing but trace kernels runs.
Not a bench contract line.   python tools/small_launch_probe.py  -> JSON lines"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
W = H = 1024
BATCH = 524288
scene = pkg.Scene(pkg.SceneBuilder(pkg.SCENE_C2))


def run(n_units, batches, batch=BATCH):
    units = [pkg.TraceUnit(i, W, H, seed=0x5EED, batch=batch) for i in range(n_units)]

    def go(first, count):
        for b in range(count):
            units[b % n_units].render_range(scene, (first + b) * batch, batch, download=False)
        for u in units:
            u.sync()

    go(0, n_units * 2)
    r0 = sum(u.ray_count() for u in units)
    t0 = time.perf_counter()
    go(100000, batches)
    dt = time.perf_counter() - t0
    rays = sum(u.ray_count() for u in units) - r0
    return round(rays / dt / 1e6, 1)


big = run(1, 4, 1 << 26)
print(json.dumps({"one launch of 2^26 photons": big}), flush=True)
for cta in [int(c) for c in os.environ.get("RL_PROBE_CTAS", "384,256").split(",")]:
    os.environ["RL_TRACE_SMALL_CTA"] = str(cta)
    for share in [int(c) for c in os.environ.get("RL_PROBE_SHARES", "1,3,6,12,24,48,96").split(",")]:
        os.environ["RL_TRACE_SHARE_MAX"] = str(share)
        units = max(16, 2 * share)
        r = run(units, 1536)
        print(json.dumps({"small_cta": cta, "share_max": share, "units": units, "photons_per_thread": round(4.6 * share, 1), "connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "default"),
                          "mrays_per_s": r, "of_one_launch": round(r / big, 3)}), flush=True)
