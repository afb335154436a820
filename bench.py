#!/usr/bin/env python
"""bench.py -- Mrays/s of the hot path on BASELINE.json configs[1]:
the reference's built-in scene (app.rs:166-363), 1024x1024, 256 spp, 1xB200.

A step is one full pass of the fused trace+splat path over the workload:
2^28 photons (= 256 spp x 1024^2; 512 reference batches of 524 288,
trace_unit.rs:67) traced and splatted into the XYZ accumulator.  A ray is one
Scene::intersect call (scene.rs:39), counted on the device.

  python bench.py --gpus N --steps K --warmup W          # this engine
  python bench.py --impl reference ...                    # the reference's CPU path (oracle port)

Under torchrun (N > 1) every rank traces its own 2^28-photon id range (weak
scaling: the job renders N x 256 spp), the XYZ framebuffers are summed onto
rank 0 with one NCCL reduce and gathered there (gather_unit.rs:49-64).

`value` is the device-resident rate (one fused launch per step).  `e2e` is the
same workload pushed through the reference host's own call pattern with host
buffers (strict mode: host/rl_replay.cpp, 524 288-photon batches, every
MappedPhoton and tristimulus buffer crossing PCIe, buffer.raw written after
every gather); `e2e_device` is the device-mode API with the frame copied back.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH = HEIGHT = 1024
SPP = 256
PHOTONS_PER_STEP = WIDTH * HEIGHT * SPP          # 2^28
SEED = 0x5EED
WORKLOAD = "built-in scene (app.rs:166-363, 339 objects), 1024x1024, 256 spp = 2^28 photons per step"

# algorithmic bytes (DESIGN.md "Kernels")
SPLAT_FUSED_BYTES_PER_PHOTON = 48      # 4 px x 3 ch x 4 B accumulator payload, no record round trip
SPLAT_BYTES_PER_PHOTON = 64            # + 16 B MappedPhoton read
GATHER_BYTES_PER_PIXEL = 72            # read px/acc/comp, write acc/comp, clear px (12 B each)
TONEMAP_BYTES_PER_PIXEL = 27           # moments pass reads 12 B, map pass reads 12 B and writes 3 B


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


class DeviceView:
    """Zero-copy torch view of a unit's device buffer (for the NCCL reduce)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2}


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


def cpu_reference_run(orc, desc, n_photons, threads, first=0):
    """The reference's CPU pipeline shape on host threads (oracle port, glibc math)."""
    # the bounded sample is cut into 8 batches per thread (the reference's 524 288-photon batch
    # would leave most threads idle on a sample this small); throughput is batch-size invariant
    batch = max(1024, n_photons // (threads * 8))
    _, ct, secs = orc.render_mt(desc, SEED, WIDTH, HEIGHT, first, n_photons, threads,
                                mode=orc.MATH_LIBM, batch=batch, want_image=False)
    return ct["rays"], secs


def cpu_sample_size(orc, desc, threads, target_seconds):
    probe = max(threads * 4096, 16384)
    rays, secs = cpu_reference_run(orc, desc, probe, threads, first=1 << 40)
    rate = probe / max(secs, 1e-6)
    n = int(rate * target_seconds)
    return max(probe, (n // 4096) * 4096)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.
    The reference is Rust and cannot be compiled here (no rustc/cargo in the
    image), so this is the oracle port of it, all host threads, glibc math."""
    if rank != 0:
        return
    import __graft_entry__ as entry
    entry.build_oracle()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc
    pkg = entry.load_package()
    desc = pkg.SceneBuilder(pkg.SCENE_C2).desc()
    threads = orc.hardware_threads()
    n = cpu_sample_size(orc, desc, threads, 6.0)
    for i in range(args.warmup):
        cpu_reference_run(orc, desc, n, threads, first=i * n)
    rays = 0
    secs = 0.0
    for i in range(args.steps):
        r, s = cpu_reference_run(orc, desc, n, threads, first=(args.warmup + i) * n)
        rays += r
        secs += s
    value = rays / secs / 1e6
    sample = f"{n} photons per step (of the workload's 2^28), {threads} threads, 8 batches per thread"
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reduce", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: gather kernel reads the peers' frames over NVLink (p2p) or NCCL reduce first")
    ap.add_argument("--photons", type=int, default=PHOTONS_PER_STEP, help=argparse.SUPPRESS)
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

    n = args.photons
    builder = pkg.SceneBuilder(pkg.SCENE_C2)
    desc = builder.desc()
    scene = pkg.Scene(builder)
    # a non-default torch stream: the units launch on it, so torch.cuda.Event records on the
    # same stream the kernels run on (handle 0, the legacy default stream, means "own stream"
    # to rl_*_set_stream)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    stream = side.cuda_stream
    assert stream != 0
    trace = pkg.TraceUnit(rank, WIDTH, HEIGHT, seed=SEED, batch=n)
    plot = pkg.PlotUnit(rank, WIDTH, HEIGHT)
    gather = pkg.GatherUnit(WIDTH, HEIGHT)
    for u in (trace, plot, gather):
        u.set_stream(stream)
    plot_ptr, plot_bytes = plot.device_buffer()
    plot_view = torch.as_tensor(DeviceView(plot_ptr, (HEIGHT, WIDTH, 4)), device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    first = rank * n                                                     # this rank's photon ids

    # N > 1, p2p: every rank exports its accumulator; rank 0 maps them and its gather kernel
    # reads all frames in one launch (peer loads over NVLink), one Kahan step per frame
    peer_ptrs = None
    use_p2p = world > 1 and args.reduce == "p2p"
    if use_p2p:
        handles = [None] * world
        dist.all_gather_object(handles, plot.ipc_export())
        if rank == 0:
            peer_ptrs = [plot_ptr] + [pkg.ipc_open(handles[r]) for r in range(1, world)]

    def combine():
        """the path's one exchange step: frames of all ranks -> rank 0's gather unit"""
        if world == 1:
            gather.accumulate(plot, clear=True)
        elif use_p2p:
            dist.barrier()                       # every rank's trace kernel has finished
            if rank == 0:
                gather.accumulate_device(peer_ptrs)
                gather.sync()
            dist.barrier()                       # frames consumed: owners may clear them
            plot.clear()
        else:
            dist.reduce(plot_view, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                gather.accumulate(plot, clear=True)
            else:
                plot.clear()

    def step():
        trace.render_fused(scene, plot, first, n)
        combine()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    host_xyz = torch.empty((HEIGHT, WIDTH, 3), dtype=torch.float32).pin_memory().numpy()

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
                          "pinned host memory, one 2^28-photon call per step (device mode, DESIGN.md 1)"}

    # ---- end to end, strict mode: the reference host's own call pattern with host buffers ------
    # host/rl_replay.cpp drives the C ABI exactly as app.rs:95-164 / task_scheduler.rs:91-182 do:
    # C worker threads, 3C trace units rendering 524 288-photon batches into host Vecs
    # (TraceUnit::render), PlotUnit::plot(&[MappedPhoton]) from host memory, GatherUnit::accumulate
    # (&[Vector3]) from host memory, buffer.raw saved after every gather, one tonemap at the end.
    # One process per GPU; rank r renders the batches [r * B, (r + 1) * B); with N > 1 the ranks'
    # gathered frames are then summed onto rank 0 (host -> device -> NCCL reduce -> host).
    # at least four steps' worth, so that the replay's ramp-up and drain (last gathers, tonemap,
    # buffer.raw flush) weigh as they do in a long render
    e2e_steps = min(max(args.steps, 4), 16)
    replay_batches = e2e_steps * (n // 524288)
    # one replay process per GPU, pinned to its share of the cores next to that GPU
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
    cmd = [exe, "--width", str(WIDTH), "--height", str(HEIGHT), "--threads", str(workers), "--batches",
           str(replay_batches), "--batch", "524288", "--seed", str(SEED), "--mode", "strict", "--scene", "2",
           "--out", out_prefix, "--first-batch", str(rank * replay_batches)]
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

    # the same unchanged call sites with the records left on the device until host code reads them
    # (PlotUnit::plot recognises `&unit.mapped_photons` by its type: INTEGRATION.md, rl_units.hpp)
    deferred = run_replay(["--records", "deferred"])
    if deferred is None:
        e2e_deferred = replay_failed("deferred-records")
    else:
        t = torch.tensor([deferred["seconds"]], dtype=torch.float64, device="cuda")
        r = torch.tensor([deferred["rays"], deferred["h2d_bytes"], deferred["d2h_bytes"]], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(r, op=dist.ReduceOp.SUM)
        e2e_deferred = {"value": int(r[0]) / float(t[0]) / 1e6, "unit": "Mrays/s",
                        "h2d_bytes_per_step": int(r[1]) // e2e_steps, "d2h_bytes_per_step": int(r[2]) // e2e_steps,
                        "steps": e2e_steps,
                        "path": "the strict-mode replay with `mapped_photons` copied out only when host code reads it "
                                "(never, in app.rs): frames and buffer.raw still cross PCIe; per-rank frames not combined"}
    replay = run_replay([])
    combine_s = 0.0
    if replay is not None and world > 1:
        t0 = time.perf_counter()
        frame = np.fromfile(out_prefix + ".raw", dtype="<f4", count=WIDTH * HEIGHT * 3)
        dev_frame = torch.from_numpy(frame).cuda()
        dist.reduce(dev_frame, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            host_xyz[...] = dev_frame.cpu().numpy().reshape(HEIGHT, WIDTH, 3)
        barrier()
        combine_s = time.perf_counter() - t0
    for suffix in (".raw", ".ppm"):
        try:
            os.remove(out_prefix + suffix)
        except OSError:
            pass
    if replay is None:
        replay = {"seconds": 1.0, "rays": 0, "h2d_bytes": 0, "d2h_bytes": 0}
        strict_failed = True
    else:
        strict_failed = False
    t = torch.tensor([replay["seconds"] + combine_s], dtype=torch.float64, device="cuda")
    r = torch.tensor([replay["rays"], replay["h2d_bytes"], replay["d2h_bytes"]], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    frame_bytes = WIDTH * HEIGHT * 12 if world > 1 else 0
    e2e = {"value": int(r[0]) / float(t[0]) / 1e6, "unit": "Mrays/s",
           "h2d_bytes_per_step": (int(r[1]) + frame_bytes * world) // e2e_steps,
           "d2h_bytes_per_step": (int(r[2]) + frame_bytes) // e2e_steps,
           "steps": e2e_steps, "batches_per_step_per_gpu": n // 524288, "worker_threads_per_gpu": workers,
           "host_cpus_rank0": f"{len(cpus)} cores" + (f" of NUMA node {numa_node}" if numa_node >= 0 else ""),
           "seconds": float(t[0]),
           "path": "strict mode: the reference host's call pattern replayed against the C ABI with host buffers "
                   "(host/rl_replay.cpp; app.rs:95-164, task_scheduler.rs:91-182): 524 288-photon TraceUnit::render "
                   "into host memory, PlotUnit::plot / GatherUnit::accumulate from host memory, buffer.raw saved "
                   "after every gather, tonemap at the end" + ("; ranks' frames summed onto rank 0" if world > 1 else "")}

    if strict_failed:
        e2e = replay_failed("strict-mode")

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- secondary, bandwidth-shaped kernels (rank 0, N = 1 semantics) ----------------------
    peak, peak_src = measured_peaks()

    def time_ms(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = 0.0
        for _ in range(reps):
            flush.fill_(1)
            flush_sum = flush[: 192 << 20].sum()      # read pass: leaves L2 full of clean lines
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            total += a.elapsed_time(b)
        return total / reps

    n_splat = 1 << 25                                   # 512 MiB of records > L2
    tr2 = pkg.TraceUnit(100, WIDTH, HEIGHT, seed=SEED, batch=n_splat)
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

    kernel_s = kernel_ms * 1e-3 / args.steps
    photons_per_launch = n
    achieved = photons_per_launch * SPLAT_FUSED_BYTES_PER_PHOTON / kernel_s / 1e9
    # DRAM traffic of the kernel from the committed ncu capture (profiles/, 2^24-photon launch),
    # scaled to this launch's photon count
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "trace_kernel_dram_bytes_per_photon.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)["dram_bytes_per_photon"] * photons_per_launch
    roofline = {
        "kernel": "trace_kernel (fused TraceUnit::render + PlotUnit::plot)",
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_photon": SPLAT_FUSED_BYTES_PER_PHOTON,
        "note": ("the fused kernel is FP32-issue/divergence bound, not HBM bound: scene tables sit in shared "
                 "memory and the 16.8 MB accumulator is L2-resident, so its HBM fraction is small by design; "
                 "the bandwidth-shaped kernels are listed under `also`"),
        "kernel_ms_per_launch": kernel_s * 1e3,
        "also": [
            {"kernel": "splat_kernel (PlotUnit::plot, 2^25 records)", "bound": "hbm",
             "achieved": splat_bytes / (splat_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
             "frac": splat_bytes / (splat_ms * 1e-3) / 1e9 / peak, "ms": splat_ms,
             "bytes": f"16 B x {n_splat} records + 48 B x {n_lit} contributing photons",
             "traffic": 562370560.0,
             "traffic_source": "profiles/r1_s2_splat_kernel_ncu.txt: dram read + write of one 2^25-record launch"},
            {"kernel": "gather_kernel (GatherUnit::accumulate + clear, 4096^2)", "bound": "hbm",
             "achieved": gw * gw * GATHER_BYTES_PER_PIXEL / (gather_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
             "frac": gw * gw * GATHER_BYTES_PER_PIXEL / (gather_ms * 1e-3) / 1e9 / peak, "ms": gather_ms,
             "traffic": 1357937152.0,
             "traffic_source": "profiles/r1_s2_gather_kernel_ncu.txt: dram read + write of one 4096^2 launch "
                               "(80 B/pixel are really moved: the source frame is padded to float4)"},
            {"kernel": "tonemap kernels (TonemapUnit::tonemap: moments, exposure, map; 4096^2)",
             "bound": "alu, not hbm: six specified ln, three exp and six IEEE divisions per pixel "
                      "(tonemap_unit.rs:82-86, srgb.rs:20-26); runs once per 30 s in the reference",
             "achieved": gw * gw * TONEMAP_BYTES_PER_PIXEL / (tonemap_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
             "frac": gw * gw * TONEMAP_BYTES_PER_PIXEL / (tonemap_ms * 1e-3) / 1e9 / peak, "ms": tonemap_ms,
             "bytes": "27 B/pixel: 12 B read for the moments, 12 B read + 3 B written by the map"},
        ],
    }

    # The trace kernel's own bound is FP32 issue.  Algorithmic flops of one Scene::intersect call =
    # the reference's linear scan (scene.rs:39-60) over the built-in scene with SURVEY 8d's per-
    # primitive costs: 311 spheres x 19 + 3 paraboloids x 45 + 3 planes/circles x 15 + 22 prisms x
    # 8 half-space tests x 15 (their containment tests, which depend on the ray, are left out: a
    # lower bound).  The kernel reaches the same hits with fewer executed operations (culling), so
    # this is reference-equivalent work per second, next to the executed issue-slot utilisation of
    # the committed ncu capture.
    flops_per_ray = 311 * 19 + 3 * 45 + 3 * 15 + 22 * 8 * 15
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp32_peak = torch.cuda.get_device_properties(local_rank).multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12
    rays_per_launch = total_rays / (world * args.steps)
    roofline["compute"] = {
        "bound": "fp32 issue (no tensor-core work on this path)",
        "achieved": rays_per_launch * flops_per_ray / kernel_s / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": rays_per_launch * flops_per_ray / kernel_s / 1e12 / fp32_peak,
        "algorithmic_flops_per_ray": flops_per_ray,
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
        n_cpu = cpu_sample_size(orc, desc, threads, 12.0)
        c_rays, c_secs = cpu_reference_run(orc, desc, n_cpu, threads)
        cpu_baseline = {"value": c_rays / c_secs / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                        "sample": f"{n_cpu} photons of the workload ({c_secs:.1f} s), 8 batches per thread, "
                                  "C++ restatement of the reference CPU path (no Rust toolchain), glibc math"}

    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "photons_per_step_per_gpu": n, "seed": SEED,
                   "l2": "flushed (256 MiB write) between timed steps",
                   "parallelism": f"photon-id partition x{world}, XYZ frames combined on rank 0 per step by "
                                  + ("the gather kernel reading peer frames over NVLink (CUDA IPC)" if use_p2p
                                     else "one NCCL reduce")
                   if world > 1 else "single GPU"},
        "rays_per_photon": total_rays / (n * world * args.steps),
        "mphotons_per_s": n * world * args.steps / (step_ms * 1e-3) / 1e6,
        "batches_per_s": n * world * args.steps / (step_ms * 1e-3) / 524288,
        "clocks": clocks,
        "e2e": e2e,
        "e2e_deferred_records": e2e_deferred,
        "e2e_device": e2e_device,
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
