#!/usr/bin/env python
"""bench.py -- Mrays/s of the hot path on the BASELINE.json configs (default: configs[1],
the reference's built-in scene (app.rs:166-363), 1024x1024, 256 spp, 1xB200).

A step is one full pass of the fused trace+splat path over the workload (configs[1]: 2^28
photons = 256 spp x 1024^2 = 512 reference batches of 524 288, trace_unit.rs:67) plus the
gather.  A ray is one Scene::intersect call (scene.rs:39), counted on the device.

  python bench.py --gpus N --steps K --warmup W [--config c1|c2|c3|c4|c5]   # this engine
  python bench.py --impl reference ...                                      # the reference's CPU path (oracle port)

Under torchrun (N > 1) every rank traces its own photon-id range (weak scaling for c1-c4: the
job renders N x spp; c5 -- 4096^2, 4096 spp -- splits its 2^36 photons over the ranks), the XYZ
framebuffers meet on rank 0 once per step (gather_unit.rs:49-64).  The first thing an N > 1 run
does is check that exchange against a sequential gather and a single-GPU render
(`parity_multi_gpu`); a failed check fails the run.

`value` is the device-resident rate (one fused launch per step).  `e2e` is the same workload
pushed through the reference host's own call pattern with host buffers (strict mode:
host/rl_replay.cpp replays app.rs / task_scheduler.rs against the C ABI: 524 288-photon
TraceUnit::render calls from C worker threads, every MappedPhoton batch copied to the host,
GatherUnit::save called after every gather (device-side snapshot; the file is rewritten at
most every 0.1 s and once more at the end), the idle task sleeping the reference's 100 ms).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0x5EED
BATCH = 524288                         # trace_unit.rs:67

# BASELINE.json configs; scene ids are include/rl_host.h's rl_builtin_scene
CONFIGS = {
    "c1": {"scene": 1, "w": 256, "h": 256, "spp": 1, "split": False,
           "workload": "single diffuse sphere + emissive plane, 256x256, 1 spp = 65 536 photons per step"},
    "c2": {"scene": 2, "w": 1024, "h": 1024, "spp": 256, "split": False,
           "workload": "built-in scene (app.rs:166-363, 339 objects), 1024x1024, 256 spp = 2^28 photons per step"},
    "c3": {"scene": 3, "w": 1024, "h": 1024, "spp": 1024, "split": False,
           "workload": "dispersive SF10 prism + area emitter, 1024x1024, 1024 spp = 2^30 photons per step"},
    "c4": {"scene": 4, "w": 2048, "h": 2048, "spp": 512, "split": False,
           "workload": "synthetic 4096 random spheres, 2048x2048, 512 spp = 2^31 photons per step"},
    "c5": {"scene": 2, "w": 4096, "h": 4096, "spp": 4096, "split": True,
           "workload": "built-in scene at 4096x4096, 4096 spp = 2^36 photons per step, split over the GPUs"},
}

# algorithmic bytes (DESIGN.md "Kernels")
SPLAT_FUSED_BYTES_PER_PHOTON = 48      # 4 px x 3 ch x 4 B accumulator payload, no record round trip
SPLAT_BYTES_PER_PHOTON = 64            # + 16 B MappedPhoton read
GATHER_BYTES_PER_PIXEL = 72            # read px/acc/comp, write acc/comp, clear px (12 B each)
TONEMAP_BYTES_PER_PIXEL = 27           # moments pass reads 12 B, map pass reads 12 B and writes 3 B
# SURVEY 8d per-primitive costs of the reference's linear scan over the built-in scene (containment
# tests of the prisms excluded: a lower bound)
C2_FLOPS_PER_RAY = 311 * 19 + 3 * 45 + 3 * 15 + 22 * 8 * 15


def photons_per_rank(cfg, world):
    total = cfg["w"] * cfg["h"] * cfg["spp"]
    return total // world if cfg["split"] else total


def make_config(name, world, reduce_mode):
    """The `config` object of the JSON line -- the same for both arms."""
    cfg = CONFIGS[name]
    return {"workload": cfg["workload"], "name": name, "photons_per_step_per_gpu": photons_per_rank(cfg, world),
            "seed": SEED, "l2": "flushed (256 MiB write) between timed steps",
            "parallelism": (f"photon-id partition x{world}, XYZ frames combined on rank 0 per step by "
                            + ("the gather kernel reading peer frames over NVLink (CUDA IPC)" if reduce_mode == "p2p"
                               else "one NCCL reduce")) if world > 1 else "single GPU"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_hash():
    """Names the kernel sources a committed ncu capture belongs to."""
    h = hashlib.sha256()
    for f in ("rl_kernels.cu", "rl_device.cuh", "rl_math.cuh", "rl_scene_layout.h", "rl_kernels.h"):
        with open(os.path.join(ROOT, "robigo-luculenta_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def measured_traffic():
    """DRAM bytes per launch of the kernels, from the committed ncu capture
    (profiles/kernel_traffic.json, written by tools/ncu_traffic.py).  A capture of other kernel
    sources is not this build's traffic: it is reported as null, loudly."""
    path = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if not os.path.exists(path):
        return {}, "profiles/kernel_traffic.json is missing: run tools/ncu_traffic.py"
    with open(path) as f:
        data = json.load(f)
    if data.get("kernel_source_sha16") != kernel_source_hash():
        print(f"[bench] profiles/kernel_traffic.json was captured for kernel sources {data.get('kernel_source_sha16')}, "
              f"this build is {kernel_source_hash()}: `traffic` is reported as null (re-run tools/ncu_traffic.py)",
              file=sys.stderr, flush=True)
        return {}, "stale: the kernel sources changed since the ncu capture (re-run tools/ncu_traffic.py)"
    return data.get("kernels", {}), f"profiles/kernel_traffic.json ({data.get('captured', '?')})"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 9:
                self.samples.append(parts)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[1])); mx.append(float(s[2])); power.append(float(s[3]))
            except ValueError:
                continue
            for name, v in zip(names, s[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_pci_address(torch, index):
    """sysfs name of the GPU's PCI function (dddd:bb:dd.f), '' if it cannot be found."""
    props = torch.cuda.get_device_properties(index)
    try:
        return f"{int(props.pci_domain_id):04x}:{int(props.pci_bus_id):02x}:{int(props.pci_device_id):02x}.0"
    except (AttributeError, TypeError, ValueError):
        pass
    try:
        out = subprocess.run(["nvidia-smi", f"--id={index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if out.count(":") == 2:
            dom, bus, rest = out.split(":")
            return f"{dom[-4:]}:{bus}:{rest}"
    except (OSError, subprocess.SubprocessError):
        pass
    return ""


def host_cpus_for_rank(local_rank, world, pci_bus_id):
    """The CPUs rank `local_rank`'s host-side work should run on: its share of the cores of the
    NUMA node its GPU hangs off (so that the page-locked buffers it allocates, and the DMA to and
    from them, stay on that socket), else an even share of the allowed cores."""
    allowed = sorted(os.sched_getaffinity(0))
    node = -1
    try:
        with open(f"/sys/bus/pci/devices/{pci_bus_id.lower()}/numa_node") as f:
            node = int(f.read().strip())
    except (OSError, ValueError):
        pass
    if node >= 0:
        try:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                local = sorted(_parse_cpulist(f.read()) & set(allowed))
            n_nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
            per_node = max(1, (world + n_nodes - 1) // n_nodes)          # ranks sharing this node
            if len(local) >= per_node:
                share = len(local) // per_node
                k = local_rank % per_node
                return local[k * share:(k + 1) * share], node
        except OSError:
            pass
    share = max(1, len(allowed) // world)
    return allowed[local_rank * share:(local_rank + 1) * share] or allowed, node




def cpu_reference_run(orc, desc, cfg, n_photons, threads, first=0):
    """The reference's CPU pipeline shape on host threads (oracle port, glibc math)."""
    # the bounded sample is cut into 8 batches per thread (the reference's 524 288-photon batch
    # would leave most threads idle on a sample this small); throughput is batch-size invariant
    batch = max(1024, n_photons // (threads * 8))
    _, ct, secs = orc.render_mt(desc, SEED, cfg["w"], cfg["h"], first, n_photons, threads,
                                mode=orc.MATH_LIBM, batch=batch, want_image=False)
    return ct["rays"], secs


def cpu_sample_size(orc, desc, cfg, threads, target_seconds):
    probe = max(threads * 4096, 16384)
    rays, secs = cpu_reference_run(orc, desc, cfg, probe, threads, first=1 << 40)
    rate = probe / max(secs, 1e-6)
    n = int(rate * target_seconds)
    return max(probe, (n // 4096) * 4096)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.
    The reference is Rust and cannot be compiled here (no rustc/cargo in the
    image), so this is the oracle port of it, all host threads, glibc math.
    Only the oracle and the host-only scene builders (librl_host.so) are loaded."""
    if rank != 0:
        return
    import __graft_entry__ as entry
    entry.build_oracle()
    entry.build_host_library()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc
    pkg = entry.load_package()
    cfg = CONFIGS[args.config]
    desc = pkg.SceneBuilder(cfg["scene"]).desc()
    threads = orc.hardware_threads()
    n = cpu_sample_size(orc, desc, cfg, threads, 6.0)
    for i in range(args.warmup):
        cpu_reference_run(orc, desc, cfg, n, threads, first=i * n)
    rays = 0
    secs = 0.0
    for i in range(args.steps):
        r, s = cpu_reference_run(orc, desc, cfg, n, threads, first=(args.warmup + i) * n)
        rays += r
        secs += s
    value = rays / secs / 1e6
    sample = (f"{n} photons per step (a bounded sample of the workload's {photons_per_rank(cfg, world)}), "
              f"{threads} threads, 8 batches per thread")
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if cfg["split"] else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": make_config(args.config, world, args.reduce), "sample": sample,
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "photons_per_s": n * args.steps / secs,
    }
    emit(line)


_JSON_FD = None


def quiet_stdout():
    """Everything but the JSON line goes to stderr: NCCL prints its version banner on fd 1 of
    rank 0, and the driver reads one JSON line from stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS),
                    help="BASELINE.json config to run (default c2 = configs[1], the one the metric is quoted on)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the scheduler replays (kernel experiments)")
    ap.add_argument("--reduce", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: gather kernel reads the peers' frames over NVLink (p2p) or NCCL reduce first")
    ap.add_argument("--photons", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import __graft_entry__ as entry

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = entry.load_package()
    if pkg.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
    from robigo_luculenta_b200 import multi_gpu

    cfg = CONFIGS[args.config]
    is_default = args.config == "c2"
    W, H = cfg["w"], cfg["h"]
    n = args.photons or photons_per_rank(cfg, world)
    builder = pkg.SceneBuilder(cfg["scene"])
    desc = builder.desc()
    scene = pkg.Scene(builder)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- N > 1: is rank 0's frame the sum of the ranks' frames?  Checked before anything is timed --
    parity_multi_gpu = None
    if world > 1:
        verdict = multi_gpu.parity_check(pkg, dist, torch, scene, rank, world)
        flag = torch.tensor([1 if (rank != 0 or verdict == "ok") else 0], device="cuda")
        dist.broadcast(flag, src=0)
        parity_multi_gpu = verdict
        if int(flag[0]) != 1:
            if rank == 0:
                emit({"metric": "Mrays/s", "value": None, "n_gpus": world, "parity_multi_gpu": verdict,
                      "error": "the multi-GPU frame exchange failed its parity check; nothing was timed"})
            dist.barrier()
            dist.destroy_process_group()
            sys.exit(1)

    # a non-default torch stream: the units launch on it, so torch.cuda.Event records on the
    # same stream the kernels run on (handle 0, the legacy default stream, means "own stream"
    # to rl_*_set_stream)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    stream = side.cuda_stream
    assert stream != 0
    trace = pkg.TraceUnit(rank, W, H, seed=SEED, batch=n)
    plot = pkg.PlotUnit(rank, W, H)
    gather = pkg.GatherUnit(W, H)
    for u in (trace, plot, gather):
        u.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    first = rank * n                                                     # this rank's photon ids
    exchange = multi_gpu.FrameExchange(pkg, dist, torch, plot, gather, rank, world, args.reduce)
    use_p2p = world > 1 and args.reduce == "p2p"
    combine = exchange.combine

    def step():
        trace.render_fused(scene, plot, first, n)
        combine()

    for _ in range(args.warmup):
        step()
    barrier()
    rays0 = trace.ray_count()
    pkg.reset_kernel_launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b, c in ev:
        flush.fill_(1)                       # L2 flush between timed steps, outside the timed events
        a.record()
        trace.render_fused(scene, plot, first, n)
        b.record()                           # [a, b] = the trace+splat kernel alone
        combine()
        c.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = pkg.kernel_launch_count()
    step_ms = sum(a.elapsed_time(c) for a, _, c in ev)
    kernel_ms = sum(a.elapsed_time(b) for a, b, _ in ev)
    rays = trace.ray_count() - rays0
    t = torch.tensor([step_ms, kernel_ms], dtype=torch.float64, device="cuda")
    r = torch.tensor([rays], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    step_ms, kernel_ms = float(t[0]), float(t[1])
    total_rays = int(r[0])
    value = total_rays / (step_ms * 1e-3) / 1e6

    # ---- end to end, device mode: the C ABI with the records left on the GPU ------------------
    # every step: scene descriptor (host) -> rl_scene_create, fused trace+splat, gather, and the
    # XYZ framebuffer copied back into pinned host memory (the four-line edit of app.rs)
    host_xyz = torch.empty((H, W, 3), dtype=torch.float32).pin_memory().numpy()

    def e2e_step():
        sc = pkg.Scene(desc)
        trace.render_fused(sc, plot, first, n)
        combine()
        if rank == 0:
            gather.download(out=host_xyz)
        return sc

    e2e_step()
    barrier()
    rays1 = trace.ray_count()
    pkg.reset_transfer_counters()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    dev_h2d, dev_d2h = pkg.transfer_counters()
    e2e_rays = trace.ray_count() - rays1
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    r = torch.tensor([e2e_rays], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    e2e_device = {"value": int(r[0]) / float(t[0]) / 1e6, "unit": "Mrays/s",
                  "h2d_bytes_per_step": dev_h2d // e2e_steps, "d2h_bytes_per_step": dev_d2h // e2e_steps,
                  "steps": e2e_steps,
                  "path": "rl_scene_create(host descriptor) + rl_trace_unit_render_fused + gather + XYZ frame to "
                          "pinned host memory, one call per step (device mode, DESIGN.md 1)"}

    # ---- end to end, strict mode: the reference host's own call pattern with host buffers ------
    # host/rl_replay.cpp drives the C ABI exactly as app.rs:95-164 / task_scheduler.rs:91-182 do:
    # C worker threads, 3C trace units rendering 524 288-photon batches (TraceUnit::render fills the
    # host Vec `mapped_photons`), PlotUnit::plot(&unit.mapped_photons), GatherUnit::accumulate(
    # &plot_unit.tristimulus_buffer), buffer.raw saved after every gather, one tonemap at the end, the
    # idle task sleeping 100 ms (app.rs:128-130).  One process per GPU; rank r renders the batches
    # [r * B, (r + 1) * B); with N > 1 the ranks' gathered frames are then summed onto rank 0
    # (host -> device -> NCCL reduce -> host) inside the timed region.
    batch = min(BATCH, n)
    replay_batches = args.steps * max(1, n // batch)
    pci = gpu_pci_address(torch, local_rank)
    cpus, numa_node = (host_cpus_for_rank(local_rank, world, pci) if world > 1
                       else (sorted(os.sched_getaffinity(0)), -1))
    workers = max(2, min(16, len(cpus)))
    try:
        exe = entry.build_replay()
    except (OSError, subprocess.SubprocessError) as e:
        print(f"[bench] rank {rank}: building rl_replay failed: {e!r}", file=sys.stderr, flush=True)
        exe = "/nonexistent/rl_replay"
    out_prefix = f"/tmp/rl_bench_replay_{os.getpid()}"
    cmd = [exe, "--width", str(W), "--height", str(H), "--threads", str(workers), "--batches",
           str(replay_batches), "--batch", str(batch), "--seed", str(SEED), "--mode", "strict", "--scene",
           str(cfg["scene"]), "--out", out_prefix, "--first-batch", str(rank * replay_batches)]
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=visible.split(",")[local_rank] if visible else str(local_rank))

    def run_replay(extra):
        """One replay per rank; None on every rank if it failed on any (so that the ranks stay in
        step and the bench line is still printed, with the failure noted in it)."""
        barrier()
        out, why = None, ""
        try:
            res = subprocess.run(cmd + extra, capture_output=True, text=True, env=env, timeout=1200,
                                 preexec_fn=(lambda: os.sched_setaffinity(0, cpus)) if world > 1 else None)
            if res.returncode == 0:
                out = json.loads(res.stdout.strip().splitlines()[-1])
            else:
                why = res.stderr[-300:]
        except (OSError, ValueError, IndexError, subprocess.SubprocessError) as e:
            why = repr(e)
        ok = torch.tensor([1 if out is not None else 0], dtype=torch.int32, device="cuda")
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok[0]) == 0:
            if why:
                print(f"[bench] rank {rank}: rl_replay {' '.join(extra)} failed: {why}", file=sys.stderr, flush=True)
            return None
        return out

    def replay_failed(what):
        d = dict(e2e_device)
        d["note"] = f"{what} replay failed on this box (stderr has the reason): this entry repeats e2e_device"
        return d

    def replay_entry(extra, what, combine_frames):
        replay = run_replay(extra)
        combine_s = 0.0
        if replay is not None and world > 1 and combine_frames:
            t0 = time.perf_counter()
            frame = np.fromfile(out_prefix + ".raw", dtype="<f4", count=W * H * 3)
            dev_frame = torch.from_numpy(frame).cuda()
            dist.reduce(dev_frame, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                host_xyz[...] = dev_frame.cpu().numpy().reshape(H, W, 3)
            barrier()
            combine_s = time.perf_counter() - t0
        for suffix in (".raw", ".ppm"):
            try:
                os.remove(out_prefix + suffix)
            except OSError:
                pass
        if replay is None:
            return replay_failed(what)
        t = torch.tensor([replay["seconds"] + combine_s, replay["seconds"] + combine_s + replay.get("setup_seconds", 0.0)],
                         dtype=torch.float64, device="cuda")
        r = torch.tensor([replay["rays"], replay["h2d_bytes"], replay["d2h_bytes"]], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(r, op=dist.ReduceOp.SUM)
        frame_bytes = W * H * 12 if (world > 1 and combine_frames) else 0
        return {"value": int(r[0]) / float(t[0]) / 1e6, "unit": "Mrays/s",
                "h2d_bytes_per_step": (int(r[1]) + frame_bytes * world) // args.steps,
                "d2h_bytes_per_step": (int(r[2]) + frame_bytes) // args.steps,
                "steps": args.steps, "batches_per_step_per_gpu": max(1, n // batch), "worker_threads_per_gpu": workers,
                "host_cpus_rank0": f"{len(cpus)} cores" + (f" of NUMA node {numa_node}" if numa_node >= 0 else ""),
                "seconds": float(t[0]),
                "d2h_gb_per_s": (int(r[2]) + frame_bytes) / float(t[0]) / 1e9,
                "h2d_gb_per_s": (int(r[1]) + frame_bytes * world) / float(t[0]) / 1e9,
                "value_with_unit_creation": int(r[0]) / float(t[1]) / 1e6,
                "timing": "steady state: wall time from the first scheduler task to buffer.raw on disk, max over "
                          "ranks; TaskScheduler::new (unit creation, page-locking of the units' host buffers, once "
                          "per render) is outside it and inside `value_with_unit_creation`; the idle task sleeps "
                          "the reference's 100 ms (app.rs:128-130)"}

    if args.no_e2e:
        e2e = dict(e2e_device)
        e2e["note"] = "--no-e2e: the scheduler replay was skipped, this entry repeats e2e_device"
        e2e_deferred = None
    else:
        # the same unchanged call sites with the records left on the device until host code reads
        # them (never, in app.rs)
        e2e_deferred = replay_entry(["--records", "deferred"], "deferred-records", False)
        if "note" not in e2e_deferred:
            e2e_deferred["path"] = ("the strict-mode replay with `mapped_photons` copied out only when host code reads "
                                    "it (never, in app.rs); per-rank frames not combined")
        e2e = replay_entry([], "strict-mode", True)
        if "note" not in e2e:
            e2e["path"] = ("strict mode: the reference host's call pattern replayed against the C ABI "
                           "(host/rl_replay.cpp; app.rs:95-164, task_scheduler.rs:91-182): 524 288-photon "
                           "TraceUnit::render calls, each a launch that shares the SMs with the other units' launches, every batch of records copied "
                           "into the unit's host Vec, PlotUnit::plot / GatherUnit::accumulate consuming the units' "
                           "device copies, GatherUnit::save after every gather (device-side snapshot, file rewritten at most every 0.1 s and at the end), tonemap at the end; 16 B per photon "
                           "cross PCIe to the host, which is what bounds this mode on several GPUs of one host "
                           "(profiles/r2_pcie_probe_8gpu.jsonl: 92 GB/s device-to-host in all with 8 GPUs copying)"
                           + ("; ranks' frames summed onto rank 0" if world > 1 else ""))

    # ---- the other BASELINE.json configs, as measured lines -------------------------------------
    def rate_of(which, w, h, n_photons, scene_obj=None, reps=3, first_id=0):
        """fused trace+splat of `reps` x n_photons photons: CUDA events on the launching stream,
        max over ranks, whole-job rays"""
        sc = scene_obj or pkg.Scene(pkg.SceneBuilder(which))
        tu = pkg.TraceUnit(500 + rank, w, h, seed=SEED, batch=n_photons)
        pl = pkg.PlotUnit(500 + rank, w, h)
        tu.set_stream(stream); pl.set_stream(stream)
        tu.render_fused(sc, pl, first_id, min(n_photons, 1 << 22))       # warm-up
        barrier()
        r0 = tu.ray_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(1)
        a.record()
        for k in range(reps):
            tu.render_fused(sc, pl, first_id + (k + 1) * n_photons * world, n_photons)
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        rr = torch.tensor([tu.ray_count() - r0], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(rr, op=dist.ReduceOp.SUM)
        ms, rr = float(ms[0]), int(rr[0])
        return {"mrays_per_s": rr / ms / 1e3, "mphotons_per_s": reps * n_photons * world / ms / 1e3,
                "rays_per_photon": rr / (reps * n_photons * world), "photons": reps * n_photons * world,
                "ms": ms, "n_gpus": world}

    other_configs = None
    c5 = None
    if is_default:
        other_configs = []
        plan = [("c1", 1, 256, 256, 1 << 24), ("c3", 3, 1024, 1024, 1 << 26), ("c4", 4, 2048, 2048, 1 << 26),
                ("c5-canvas", 2, 4096, 4096, 1 << 26)]
        for name, which, w, h, n_cfg in plan:
            entry_ = rate_of(which, w, h, n_cfg, first_id=rank * n_cfg)
            entry_["config"] = name
            entry_["workload"] = (CONFIGS[name]["workload"] if name in CONFIGS
                                  else "built-in scene on configs[4]'s 4096x4096 canvas (268 MB accumulator, beyond L2)")
            entry_["note"] = ("fused trace+splat, weak scaling (every rank its own photon ids), CUDA events; "
                              "a sample of the config's photon count, the rate does not depend on it")
            other_configs.append(entry_)
        c5_spp = int(os.environ.get("RL_BENCH_C5_SPP", "4096" if world == 8 else "0"))
        if c5_spp > 0 and world > 1:
            # configs[4] in full (at 8 GPUs): 4096^2, 4096 spp = 2^36 photons over the GPUs, one exchange
            # of the XYZ frame, Kahan gather and tonemap on rank 0.  (RL_BENCH_C5_SPP runs the same
            # code at another world size / sample count: how this block is tested on two GPUs.)
            w5 = h5 = 4096
            n5 = (w5 * h5 * c5_spp) // world
            tu5 = pkg.TraceUnit(700 + rank, w5, h5, seed=SEED, batch=n5)
            pl5 = pkg.PlotUnit(700 + rank, w5, h5)
            g5 = pkg.GatherUnit(w5, h5)
            tm5 = pkg.TonemapUnit(w5, h5) if rank == 0 else None
            for u in (tu5, pl5, g5):
                u.set_stream(stream)
            ex5 = multi_gpu.FrameExchange(pkg, dist, torch, pl5, g5, rank, world, args.reduce)
            tu5.render_fused(scene, pl5, rank * n5, 1 << 22)
            ex5.combine()
            barrier()
            r0 = tu5.ray_count()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            tu5.render_fused(scene, pl5, rank * n5, n5)
            ex5.combine()
            if rank == 0:
                tm5.set_stream(stream)
                tm5.tonemap(g5, download=False)
            b.record()
            barrier()
            ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
            rr = torch.tensor([tu5.ray_count() - r0], dtype=torch.int64, device="cuda")
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(rr, op=dist.ReduceOp.SUM)
            c5 = {"config": "c5", "workload": CONFIGS["c5"]["workload"], "mrays_per_s": int(rr[0]) / float(ms[0]) / 1e3,
                  "seconds": float(ms[0]) / 1e3, "photons": n5 * world, "rays_per_photon": int(rr[0]) / (n5 * world),
                  "n_gpus": world, "spp": c5_spp,
                  "note": f"one full pass: {n5} photons per GPU, one frame exchange, gather + tonemap on rank 0"}
            ex5.close()
            del tu5, pl5, g5, tm5

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- secondary, bandwidth-shaped kernels (rank 0, N = 1 semantics) ----------------------
    peak, peak_src = measured_peaks()
    traffic, traffic_src = measured_traffic()

    def time_ms(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = 0.0
        for _ in range(reps):
            flush.fill_(1)
            flush[: 192 << 20].sum()                  # read pass: leaves L2 full of clean lines
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            total += a.elapsed_time(b)
        return total / reps

    kernel_s = kernel_ms * 1e-3 / args.steps
    achieved = n * SPLAT_FUSED_BYTES_PER_PHOTON / kernel_s / 1e9
    tk = traffic.get("trace_kernel")
    roofline = {
        "kernel": "trace_kernel (fused TraceUnit::render + PlotUnit::plot)",
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": tk["dram_bytes_per_photon"] * n if tk else None, "traffic_source": traffic_src,
        "peak_source": peak_src,
        "algorithmic_bytes_per_photon": SPLAT_FUSED_BYTES_PER_PHOTON,
        "note": ("the fused kernel is FP32-issue/divergence bound, not HBM bound: scene tables sit in shared "
                 "memory and the accumulator is L2-resident, so its HBM fraction is small by design; "
                 "the bandwidth-shaped kernels are listed under `also`, the FP32 figure under `compute`"),
        "kernel_ms_per_launch": kernel_s * 1e3,
    }
    if is_default:
        n_splat = 1 << 25                                   # 512 MiB of records > L2
        tr2 = pkg.TraceUnit(100, W, H, seed=SEED, batch=n_splat)
        tr2.set_stream(stream)
        records = tr2.render_range(scene, 0, n_splat, download=True)
        n_lit = int(np.count_nonzero(records["probability"]))   # only these carry accumulator payload
        del records
        splat_ms = time_ms(lambda: plot.plot(tr2), 5)
        plot.clear()
        splat_bytes = n_splat * 16 + n_lit * SPLAT_FUSED_BYTES_PER_PHOTON
        gw = 4096                                           # 4096^2: 192 MiB acc + 256 MiB plot > L2
        gp, gg = pkg.PlotUnit(101, gw, gw), pkg.GatherUnit(gw, gw)
        gp.set_stream(stream); gg.set_stream(stream)
        gather_ms = time_ms(lambda: gg.accumulate(gp, clear=True), 5)
        # K4 on the same canvas: a frame with structure (a few accumulated splats), image left on the device
        tm = pkg.TonemapUnit(gw, gw)
        tm.set_stream(stream)
        tr3 = pkg.TraceUnit(102, gw, gw, seed=SEED, batch=1 << 22)
        tr3.set_stream(stream)
        tr3.render_fused(scene, gp, 0, 1 << 22)
        gg.accumulate(gp, clear=True)
        tonemap_ms = time_ms(lambda: tm.tonemap(gg, download=False), 5)
        del gp, gg, tr2, tr3, tm
        sk, gk = traffic.get("splat_kernel"), traffic.get("gather_kernel")
        splat_rate = splat_bytes / (splat_ms * 1e-3) / 1e9
        splat_dram = sk["dram_bytes"] * (n_splat / sk["records"]) if sk else None
        gather_dram = gk["dram_bytes"] * (gw * gw / gk["pixels"]) if gk else None
        roofline["also"] = [
            {"kernel": "splat_kernel (PlotUnit::plot, 2^25 records)", "bound": "hbm",
             "achieved": splat_rate, "peak": peak, "unit": "GB/s", "frac": splat_rate / peak, "ms": splat_ms,
             "bytes": f"algorithmic: 16 B x {n_splat} records + 48 B of accumulator payload x {n_lit} contributing "
                      "photons (the payload lands in L2, not in DRAM)",
             "traffic": splat_dram, "traffic_source": traffic_src,
             "frac_by_dram_bytes": splat_dram / (splat_ms * 1e-3) / 1e9 / peak if splat_dram else None},
            {"kernel": f"gather_kernel (GatherUnit::accumulate + clear, {gw}^2)", "bound": "hbm",
             "achieved": gw * gw * GATHER_BYTES_PER_PIXEL / (gather_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
             "frac": gw * gw * GATHER_BYTES_PER_PIXEL / (gather_ms * 1e-3) / 1e9 / peak, "ms": gather_ms,
             "traffic": gather_dram, "traffic_source": traffic_src,
             "frac_by_dram_bytes": gather_dram / (gather_ms * 1e-3) / 1e9 / peak if gather_dram else None},
            {"kernel": f"tonemap kernels (TonemapUnit::tonemap: moments, exposure, map; {gw}^2)",
             "bound": "alu, not hbm: six specified ln, three exp and six IEEE divisions per pixel "
                      "(tonemap_unit.rs:82-86, srgb.rs:20-26); runs once per 30 s in the reference",
             "achieved": gw * gw * TONEMAP_BYTES_PER_PIXEL / (tonemap_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
             "frac": gw * gw * TONEMAP_BYTES_PER_PIXEL / (tonemap_ms * 1e-3) / 1e9 / peak, "ms": tonemap_ms,
             "bytes": "27 B/pixel: 12 B read for the moments, 12 B read + 3 B written by the map"},
        ]

    if cfg["scene"] == 2:
        # The trace kernel's own bound is FP32 issue.  Algorithmic flops of one Scene::intersect call =
        # the reference's linear scan (scene.rs:39-60) over the built-in scene with SURVEY 8d's per-
        # primitive costs.  The kernel reaches the same hits with fewer executed operations (culling),
        # so this is reference-equivalent work per second, next to the executed issue-slot
        # utilisation of the committed ncu capture.
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fp32_peak = torch.cuda.get_device_properties(local_rank).multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12
        rays_per_launch = total_rays / (world * args.steps)
        roofline["compute"] = {
            "bound": "fp32 issue (no tensor-core work on this path)",
            "achieved": rays_per_launch * C2_FLOPS_PER_RAY / kernel_s / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": rays_per_launch * C2_FLOPS_PER_RAY / kernel_s / 1e12 / fp32_peak,
            "algorithmic_flops_per_ray": C2_FLOPS_PER_RAY,
            "peak_source": f"SMs x 128 FP32 lanes x 2 x {sm_mhz:.0f} MHz (sampled SM clock)",
            "note": "algorithmic = the reference's brute-force scan per Scene::intersect call, containment tests of "
                    "the prisms excluded (lower bound); executed issue-slot utilisation is in profiles/",
        }

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        entry.build_oracle()
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as orc
        threads = orc.hardware_threads()
        n_cpu = cpu_sample_size(orc, desc, cfg, threads, 12.0)
        c_rays, c_secs = cpu_reference_run(orc, desc, cfg, n_cpu, threads)
        cpu_baseline = {"value": c_rays / c_secs / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                        "sample": f"{n_cpu} photons of the workload ({c_secs:.1f} s), 8 batches per thread, "
                                  "C++ restatement of the reference CPU path (no Rust toolchain), glibc math"}

    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if cfg["split"] else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args.config, world, args.reduce),
        "rays_per_photon": total_rays / (n * world * args.steps),
        "mphotons_per_s": n * world * args.steps / (step_ms * 1e-3) / 1e6,
        "batches_per_s": n * world * args.steps / (step_ms * 1e-3) / BATCH,
        "clocks": clocks,
        "e2e": e2e,
        "e2e_deferred_records": e2e_deferred,
        "e2e_device": e2e_device,
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    if parity_multi_gpu is not None:
        line["parity_multi_gpu"] = parity_multi_gpu
    if other_configs is not None:
        line["other_configs"] = other_configs
    if c5 is not None:
        line["c5"] = c5
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
