"""Run under torchrun (one rank per GPU): the multi-GPU path against a
single-GPU render of the same photon set, on rank 0.  The check itself is the one
bench.py runs in its warm-up on every N > 1 run and smoke() runs when it sees
two GPUs (robigo-luculenta_b200/multi_gpu.py: parity_check).

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.load_package()
    from robigo_luculenta_b200 import multi_gpu
    scene = pkg.Scene(pkg.SceneBuilder(pkg.SCENE_C2))
    verdict = multi_gpu.parity_check(pkg, dist, torch, scene, rank, world)
    if rank == 0:
        print("multi-GPU parity:", verdict)
    flag = torch.tensor([1 if (rank != 0 or verdict == "ok") else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag[0]) != 1:
        sys.exit(1)
    if rank == 0:
        print("MULTI_GPU_CHECK_OK world", world)


if __name__ == "__main__":
    main()
