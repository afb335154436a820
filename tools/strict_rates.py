#!/usr/bin/env python
"""Throughput of the reference's own call pattern -- 524 288-photon batches
(trace_unit.rs:67) through TraceUnit::render / PlotUnit::plot / GatherUnit::accumulate
with host buffers -- for DESIGN.md and for choosing the small-batch launch policy.
Not a bench contract line.

  1. one unit, one stream: back-to-back render launches of one reference batch,
     for each small-batch policy (environment knobs of launch_trace);
  2. several units on their own streams, launches interleaved (what the C worker
     threads of app.rs produce), same policies;
  3. the scheduler replay (host/rl_replay.cpp), strict and device mode, pinned and
     pageable host buffers.
Writes JSON lines to stdout."""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry

pkg = entry.load_package()
W = H = 1024
BATCH = 524288
scene = pkg.Scene(pkg.SceneBuilder(pkg.SCENE_C2))


def set_policy(small_paths, small_cta, blocks_per_sm):
    os.environ["RL_TRACE_SMALL_PATHS"] = str(small_paths)
    os.environ["RL_TRACE_SMALL_CTA"] = str(small_cta)
    os.environ["RL_TRACE_BLOCKS_PER_SM"] = str(blocks_per_sm)


def run_units(n_units, batches, fused):
    units = [pkg.TraceUnit(i, W, H, seed=0x5EED, batch=BATCH) for i in range(n_units)]
    plots = [pkg.PlotUnit(i, W, H) for i in range(n_units)] if fused else None

    def go(first, count):
        for b in range(count):
            u = units[b % n_units]
            if fused:
                u.render_fused(scene, plots[b % n_units], (first + b) * BATCH, BATCH)
            else:
                u.render_range(scene, (first + b) * BATCH, BATCH, download=False)
        for u in units:
            u.sync()

    go(0, n_units * 2)
    r0 = sum(u.ray_count() for u in units)
    t0 = time.perf_counter()
    go(1000, batches)
    dt = time.perf_counter() - t0
    rays = sum(u.ray_count() for u in units) - r0
    return {"mrays_per_s": round(rays / dt / 1e6, 1), "batches_per_s": round(batches / dt, 1),
            "ms_per_batch": round(dt / batches * 1e3, 3)}


def clear_policy():
    for k in ("RL_TRACE_SMALL_PATHS", "RL_TRACE_SMALL_CTA", "RL_TRACE_BLOCKS_PER_SM"):
        os.environ.pop(k, None)


policies = [("default (adaptive share)", None), ("768x1 (large-batch config)", (0, 256, 0)),
            ("256x3", (1000, 256, 3)), ("256x1 of 3", (1000, 256, 1))]
for name, knobs in (policies if "--units" in sys.argv else []):
    clear_policy()
    if knobs:
        set_policy(*knobs)
    for n_units in (1, 2, 3, 8, 16):
        r = run_units(n_units, 192, False)
        r.update({"test": "units", "policy": name, "units": n_units})
        print(json.dumps(r), flush=True)

if "--replay" in sys.argv:
    exe = entry.build_replay()
    threads = min(16, os.cpu_count() or 8)
    batches = "2048"
    # (policy knobs, mode, page-locked host buffers, lazy host mirrors, worker threads)
    runs = [(None, "strict", 1, 1, threads), (None, "strict", 1, 0, threads), (None, "strict", 0, 0, threads),
            (None, "device", 1, 1, threads), (None, "strict", 1, 1, 8), (None, "strict", 1, 1, 4),
            (None, "strict", 1, 1, 2), (None, "device", 1, 1, 4), ((0, 256, 0), "strict", 1, 1, threads)]
    for knobs, mode, pin, lazy, thr in runs:
        clear_policy()
        if knobs:
            set_policy(*knobs)
        res = subprocess.run([exe, "--width", str(W), "--height", str(H), "--threads", str(thr), "--batches", batches,
                              "--batch", str(BATCH), "--mode", mode, "--scene", "2", "--out", "/tmp/replay_rate",
                              "--pin", str(pin), "--lazy", str(lazy)], capture_output=True, text=True, timeout=900)
        if res.returncode != 0:
            print(json.dumps({"test": "replay", "error": res.stderr[-300:]}), flush=True)
            continue
        r = json.loads(res.stdout.strip().splitlines()[-1])
        r.update({"test": "replay", "policy": "768x1 (large-batch config)" if knobs else "default (adaptive share)"})
        print(json.dumps(r), flush=True)
