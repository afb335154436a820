#!/usr/bin/env python
"""Aggregate host<->device copy bandwidth of the box with 1, 2, 4, ... GPUs copying at once
(one stream per GPU and direction, one process).  The strict-mode replay moves every
MappedPhoton batch to the host (16 B per photon), so this is the ceiling of `e2e` at N GPUs.

  python tools/pcie_probe.py [--mib 256] [--iters 12] [--alloc pinned|registered|wc] [--only 8]   -> JSON lines

--alloc: how the host buffer is page-locked.  pinned = cudaHostAlloc (default flags);
registered = malloc'ed memory + cudaHostRegister (what the units' Vecs get, rl_host_register);
wc = cudaHostAlloc(cudaHostAllocWriteCombined): not snooped by the CPU caches during the
transfer, slow for the CPU to read back.
"""
import argparse
import ctypes
import json
import time

import torch

cudart = ctypes.CDLL("libcudart.so.12")
cudart.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
cudart.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
cudart.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
libc = ctypes.CDLL("libc.so.6")
libc.aligned_alloc.restype = ctypes.c_void_p
libc.aligned_alloc.argtypes = [ctypes.c_size_t, ctypes.c_size_t]
H2D, D2H = 1, 2


def host_buffer(n, alloc):
    p = ctypes.c_void_p()
    if alloc == "registered":
        p = ctypes.c_void_p(libc.aligned_alloc(4096, n))
        ctypes.memset(p, 0, n)
        rc = cudart.cudaHostRegister(p, n, 1)                      # portable
    else:
        rc = cudart.cudaHostAlloc(ctypes.byref(p), n, 1 | (4 if alloc == "wc" else 0))
    if rc != 0:
        raise RuntimeError(f"host allocation ({alloc}) failed: cudaError {rc}")
    return p


def run(devs, direction, mib, iters, alloc):
    n = mib << 20
    bufs = []
    for d in devs:
        with torch.cuda.device(d):
            dev = torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}")
            dev2 = torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}")
            bufs.append((d, dev, host_buffer(n, alloc), dev2, host_buffer(n, alloc),
                         torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)))

    def issue():
        for d, dev, host, dev2, host2, s1, s2 in bufs:
            with torch.cuda.device(d):
                if direction in ("d2h", "both"):
                    cudart.cudaMemcpyAsync(host, dev.data_ptr(), n, D2H, s1.cuda_stream)
                if direction in ("h2d", "both"):
                    cudart.cudaMemcpyAsync(dev2.data_ptr(), host2, n, H2D, s2.cuda_stream)

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
    ap.add_argument("--alloc", default="pinned", choices=["pinned", "registered", "wc"])
    ap.add_argument("--only", type=int, default=0, help="only this many GPUs")
    ap.add_argument("--directions", default="d2h,h2d,both")
    args = ap.parse_args()
    g = torch.cuda.device_count()
    counts = [c for c in (1, 2, 4, 8) if c <= g and (not args.only or c == args.only)]
    for c in counts:
        for direction in args.directions.split(","):
            total, per_dir = run(list(range(c)), direction, args.mib, args.iters, args.alloc)
            print(json.dumps({"gpus": c, "alloc": args.alloc, "direction": direction, "aggregate_GBps": round(total, 1),
                              "per_direction_GBps": round(per_dir, 1), "per_gpu_GBps": round(total / c, 1)}), flush=True)


if __name__ == "__main__":
    main()
