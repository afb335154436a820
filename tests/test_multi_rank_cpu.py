"""World-size-2 checks of the multi-GPU host logic on CPU (gloo).

The path shards by photon id (DESIGN.md "Multi-GPU"): rank r traces ids
[r*n, (r+1)*n) into its own full-frame XYZ accumulator, one sum-reduce lands
the frames on rank 0, which runs the Kahan gather.  Without a GPU the per-rank
frames come from the oracle (the checker standing in for the kernel); what is
tested is the partition, the collective and the gather order."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def photon_range(rank, world, n_per_rank):
    """bench.py's partition: disjoint, contiguous, covers [0, world * n)."""
    return rank * n_per_rank, n_per_rank


def _worker(rank, world, port, n_per_rank, w, h, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as entry
    import oracle_lib as orc
    pkg = entry.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc = pkg.SceneBuilder(pkg.SCENE_C2).desc()
    first, n = photon_range(rank, world, n_per_rank)
    frame = orc.plot(w, h, orc.trace(desc, 0x5EED, w, h, first, n))
    t = torch.from_numpy(frame)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)          # the path's one collective
    if rank == 0:
        acc = np.zeros_like(frame)
        comp = np.zeros_like(frame)
        orc.gather_accumulate(acc, comp, t.numpy())       # rank 0: GatherUnit::accumulate
        np.save(out_path, acc)
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_a_disjoint_cover():
    for world in (1, 2, 4, 8):
        n = 1 << 28
        ranges = [photon_range(r, world, n) for r in range(world)]
        assert ranges[0][0] == 0
        for (a, na), (b, _) in zip(ranges[:-1], ranges[1:]):
            assert a + na == b
        assert ranges[-1][0] + ranges[-1][1] == world * n


@pytest.mark.timeout(300)
def test_two_rank_reduce_equals_single_rank(tmp_path, pkg, orc):
    w, h, n = 48, 32, 6000
    out = str(tmp_path / "acc.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, n, w, h, out), nprocs=2, join=True)
    got = np.load(out)
    desc = pkg.SceneBuilder(pkg.SCENE_C2).desc()
    want = orc.plot(w, h, orc.trace(desc, 0x5EED, w, h, 0, 2 * n))
    # same photon set, different summation order
    assert float(np.abs(got - want).max()) <= 1e-5 * float(np.abs(want).max()) + 1e-12
    assert got.any()


def test_host_cpu_shares_for_replay_processes():
    # bench.py pins one scheduler-replay process per GPU to its share of the host's cores; without
    # NUMA information (numa_node = -1, as on the VMs here) the allowed cores are split evenly
    import importlib.util
    spec = importlib.util.spec_from_file_location("rl_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    allowed = sorted(os.sched_getaffinity(0))
    world = 2 if len(allowed) >= 2 else 1
    shares = [bench.host_cpus_for_rank(r, world, "ffff:ff:1f.0")[0] for r in range(world)]
    assert all(shares) and all(set(s) <= set(allowed) for s in shares)
    if world == 2:
        assert not set(shares[0]) & set(shares[1])
    assert bench._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
