#!/usr/bin/env python
"""The small-launch rate of tools/small_launch_probe.py with the launches issued by T host threads
(each over its own units, as the worker threads of app.rs:95-111 do) instead of one.
python tools/small_launch_threads_probe.py -> JSON lines"""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
W = H = 1024
BATCH = 524288
scene = pkg.Scene(pkg.SceneBuilder(pkg.SCENE_C2))


def run(n_units, n_threads, batches, download):
    units = [pkg.TraceUnit(i, W, H, seed=0x5EED, batch=BATCH) for i in range(n_units)]
    outs = [pkg.pinned_records(BATCH) if download and hasattr(pkg, "pinned_records") else None for _ in units]

    def worker(t, first, count):
        mine = list(range(t, n_units, n_threads))
        for b in range(count):
            k = mine[b % len(mine)]
            units[k].render_range(scene, (first + t * count + b) * BATCH, BATCH, download=False)

    def go(first, per_thread):
        ts = [threading.Thread(target=worker, args=(t, first, per_thread)) for t in range(n_threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for u in units:
            u.sync()

    go(0, 2 * n_units // n_threads)
    r0 = sum(u.ray_count() for u in units)
    t0 = time.perf_counter()
    go(100000, batches // n_threads)
    dt = time.perf_counter() - t0
    rays = sum(u.ray_count() for u in units) - r0
    return round(rays / dt / 1e6, 1)


for share in (12, 24):
    os.environ["RL_TRACE_SHARE_MAX"] = str(share)
    for units, threads in ((48, 1), (48, 4), (48, 16), (16, 16), (96, 16)):
        r = run(units, threads, 1536, False)
        print(json.dumps({"share_max": share, "units": units, "host_threads": threads, "mrays_per_s": r,
                          "connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "default")}), flush=True)
