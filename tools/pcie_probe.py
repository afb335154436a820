#!/usr/bin/env python
"""Aggregate host<->device copy bandwidth of the box with 1, 2, 4, ... GPUs copying at once
(page-locked host memory, one stream per GPU, one process).  The strict-mode replay moves every
MappedPhoton batch to the host (16 B per photon), so this is the ceiling of `e2e` at N GPUs.

  python tools/pcie_probe.py [--mib 256] [--iters 12]        -> JSON lines
"""
import argparse
import json
import time

import torch


def run(devs, direction, mib, iters):
    n = mib << 20
    bufs = []
    for d in devs:
        with torch.cuda.device(d):
            dev = torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}")
            host = torch.empty(n, dtype=torch.uint8).pin_memory()
            dev2 = torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}")
            host2 = torch.empty(n, dtype=torch.uint8).pin_memory()
            bufs.append((d, dev, host, dev2, host2, torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)))

    def issue():
        for d, dev, host, dev2, host2, s1, s2 in bufs:
            if direction in ("d2h", "both"):
                with torch.cuda.stream(s1):
                    host.copy_(dev, non_blocking=True)
            if direction in ("h2d", "both"):
                with torch.cuda.stream(s2):
                    dev2.copy_(host2, non_blocking=True)

    def sync():
        for d, *_ in bufs:
            torch.cuda.synchronize(d)

    issue(); sync()
    t0 = time.perf_counter()
    for _ in range(iters):
        issue()
    sync()
    dt = time.perf_counter() - t0
    per_dir = n * iters * len(devs) / dt / 1e9
    return per_dir * (2 if direction == "both" else 1), per_dir


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=256)
    ap.add_argument("--iters", type=int, default=12)
    args = ap.parse_args()
    g = torch.cuda.device_count()
    counts = [c for c in (1, 2, 4, 8) if c <= g]
    for c in counts:
        for direction in ("d2h", "h2d", "both"):
            total, per_dir = run(list(range(c)), direction, args.mib, args.iters)
            print(json.dumps({"gpus": c, "direction": direction, "aggregate_GBps": round(total, 1),
                              "per_direction_GBps": round(per_dir, 1), "per_gpu_GBps": round(total / c, 1)}), flush=True)


if __name__ == "__main__":
    main()
