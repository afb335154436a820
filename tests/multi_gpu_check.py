"""Run under torchrun (one rank per GPU): the multi-GPU path against a
single-GPU render of the same photon set, on rank 0.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


class DeviceView:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2}


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.load_package()
    w, h, n, seed = 256, 192, 1 << 18, 0x5EED
    builder = pkg.SceneBuilder(pkg.SCENE_C2)
    scene = pkg.Scene(builder)
    trace = pkg.TraceUnit(rank, w, h, seed=seed, batch=n)
    plot = pkg.PlotUnit(rank, w, h)
    trace.render_fused(scene, plot, rank * n, n)            # rank r: photon ids [r n, (r+1) n)
    plot.sync()
    frame = plot.tristimulus_buffer
    ptr, _ = plot.device_buffer()

    # (1) peer-to-peer: rank 0's gather kernel reads every rank's frame through CUDA IPC
    handles = [None] * world
    dist.all_gather_object(handles, plot.ipc_export())
    dist.barrier()
    ok = True
    if rank == 0:
        ptrs = [ptr] + [pkg.ipc_open(handles[r]) for r in range(1, world)]
        g_p2p = pkg.GatherUnit(w, h)
        g_p2p.accumulate_device(ptrs)
        img_p2p, comp_p2p = g_p2p.download(with_compensation=True)
    dist.barrier()

    # expected on rank 0: Kahan-accumulate the frames in rank order (frames gathered over gloo-free NCCL)
    frames = [torch.empty((h, w, 3), dtype=torch.float32, device="cuda") for _ in range(world)]
    dist.all_gather(frames, torch.from_numpy(frame).cuda())
    if rank == 0:
        g_ref = pkg.GatherUnit(w, h)
        for f in frames:
            g_ref.accumulate(f.cpu().numpy())
        img_ref, comp_ref = g_ref.download(with_compensation=True)
        ok &= np.array_equal(img_p2p, img_ref) and np.array_equal(comp_p2p, comp_ref)
        print("p2p gather == per-frame Kahan gather (bit-equal):", ok)

    # (2) NCCL reduce of the padded frames, then one gather step
    view = torch.as_tensor(DeviceView(ptr, (h, w, 4)), device="cuda")
    dist.reduce(view, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        g_nccl = pkg.GatherUnit(w, h)
        g_nccl.accumulate(plot, clear=True)
        img_nccl = g_nccl.download()
        tol = 1e-5 * float(np.abs(img_ref).max())
        e = float(np.abs(img_nccl - img_ref).max())
        print("nccl reduce vs p2p: max |diff|", e, "tol", tol)
        ok &= e <= tol
        # (3) the same photon set on one GPU
        single = pkg.PlotUnit(99, w, h)
        t1 = pkg.TraceUnit(99, w, h, seed=seed, batch=world * n)
        t1.render_fused(scene, single, 0, world * n)
        e = float(np.abs(single.tristimulus_buffer - img_ref).max())
        print("single-GPU render of the union vs multi-GPU: max |diff|", e, "tol", tol)
        ok &= e <= tol
        ok &= bool(img_ref.any())
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag[0]) != 1:
        sys.exit(1)
    if rank == 0:
        print("MULTI_GPU_CHECK_OK world", world)


if __name__ == "__main__":
    main()
