"""The path's one exchange step across GPUs, and its parity check.

Photon paths are independent and the image is a linear sum of splats
(trace_unit.rs:152, plot_unit.rs:80-83), so rank r traces photon ids
[r*n, (r+1)*n) into its own full-frame accumulator and the frames meet once per
gather: rank 0's gather kernel reads every rank's frame (peer loads over
NVLink through CUDA IPC) and applies one Kahan step per frame in rank order --
what the reference does with N plot units (gather_unit.rs:49-64, app.rs:143-148).
`--reduce nccl` sums the frames with one NCCL reduce first instead.

One process per GPU; `dist` is an initialised torch.distributed (NCCL) module.
"""
from __future__ import annotations

import numpy as np


class DeviceView:
    """Zero-copy torch view of a unit's device buffer (for the NCCL reduce)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2}


class FrameExchange:
    """frames of all ranks -> rank 0's gather unit (the path's only collective step)"""

    def __init__(self, pkg, dist, torch, plot, gather, rank, world, mode="p2p"):
        self.pkg, self.dist, self.torch = pkg, dist, torch
        self.plot, self.gather, self.rank, self.world, self.mode = plot, gather, rank, world, mode
        self.peer_ptrs = None
        self.opened = []
        ptr, _ = plot.device_buffer()
        self.view = None
        if world > 1 and mode == "p2p":
            handles = [None] * world
            dist.all_gather_object(handles, plot.ipc_export())
            if rank == 0:
                self.opened = [pkg.ipc_open(handles[r]) for r in range(1, world)]
                self.peer_ptrs = [ptr] + self.opened
        elif world > 1:
            self.view = torch.as_tensor(DeviceView(ptr, (plot.height, plot.width, 4)), device="cuda")

    def combine(self):
        if self.world == 1:
            self.gather.accumulate(self.plot, clear=True)
        elif self.mode == "p2p":
            self.dist.barrier()                  # every rank's trace kernel has finished
            if self.rank == 0:
                self.gather.accumulate_device(self.peer_ptrs)
                self.gather.sync()
            self.dist.barrier()                  # frames consumed: owners may clear them
            self.plot.clear()
        else:
            self.dist.reduce(self.view, dst=0, op=self.dist.ReduceOp.SUM)
            if self.rank == 0:
                self.gather.accumulate(self.plot, clear=True)
            else:
                self.plot.clear()

    def close(self):
        for p in self.opened:
            self.pkg.ipc_close(p)
        self.opened = []


def parity_check(pkg, dist, torch, scene, rank, world, width=256, height=192, n_per_rank=1 << 18, seed=0x5EED):
    """Rank 0's frame is the sum of the ranks' frames: checked on every N > 1 run.

    (1) the peer-reading gather kernel's accumulator AND compensation buffer are bit-equal to a
        sequential Kahan accumulation of the all-gathered frames in rank order
        (gather_unit.rs:49-64 with N plot units);
    (2) the NCCL-reduce variant and (3) a single-GPU render of the union of the photon ids agree
        with it within 1e-5 * max|image| (float summation order is all that differs).
    Returns "ok" or the reason on rank 0 (None on the other ranks); collective: call on all ranks."""
    w, h, n = width, height, n_per_rank
    trace = pkg.TraceUnit(1000 + rank, w, h, seed=seed, batch=n)
    plot = pkg.PlotUnit(1000 + rank, w, h)
    trace.render_fused(scene, plot, rank * n, n)
    plot.sync()
    frame = plot.tristimulus_buffer
    ptr, _ = plot.device_buffer()
    why = []

    gather = pkg.GatherUnit(w, h)
    ex = FrameExchange(pkg, dist, torch, plot, gather, rank, world, "p2p")
    dist.barrier()
    if rank == 0:
        gather.accumulate_device(ex.peer_ptrs)
        img_p2p, comp_p2p = gather.download(with_compensation=True)
    dist.barrier()
    ex.close()

    frames = [torch.empty((h, w, 3), dtype=torch.float32, device="cuda") for _ in range(world)]
    dist.all_gather(frames, torch.from_numpy(frame).cuda())
    if rank == 0:
        ref = pkg.GatherUnit(w, h)
        for f in frames:
            ref.accumulate(f.cpu().numpy())
        img_ref, comp_ref = ref.download(with_compensation=True)
        if not (np.array_equal(img_p2p, img_ref) and np.array_equal(comp_p2p, comp_ref)):
            why.append("peer-reading gather differs from the sequential Kahan gather of the frames")
        if not img_ref.any():
            why.append("black frame")

    view = torch.as_tensor(DeviceView(ptr, (h, w, 4)), device="cuda")
    dist.reduce(view, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        g_nccl = pkg.GatherUnit(w, h)
        g_nccl.accumulate(plot, clear=True)
        tol = 1e-5 * float(np.abs(img_ref).max())
        e = float(np.abs(g_nccl.download() - img_ref).max())
        if e > tol:
            why.append(f"nccl reduce off by {e} > {tol}")
        single = pkg.PlotUnit(1999, w, h)
        t1 = pkg.TraceUnit(1999, w, h, seed=seed, batch=world * n)
        t1.render_fused(scene, single, 0, world * n)
        e = float(np.abs(single.tristimulus_buffer - img_ref).max())
        if e > tol:
            why.append(f"single-GPU render of the union off by {e} > {tol}")
    dist.barrier()
    if rank != 0:
        return None
    return "ok" if not why else "; ".join(why)
