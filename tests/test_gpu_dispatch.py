"""Group launches (csrc/rl_api.cu TraceDispatcher, csrc/rl_kernels.cu K1; RL_TRACE_GROUPS=1):
batches of the reference's size are queued to the scene's dispatcher and traced several to a
launch, each batch one segment of the launch's photon pool (task_scheduler.rs:95-96,127-182 is the caller that
produces them).  Whichever launch, CTA and lane traces a photon, its record is a function of
(scene, seed, photon id): everything here is compared bit for bit with the oracle, and with the
one-launch-per-batch path."""
import threading

import numpy as np
import pytest

from test_gpu_parity import SEED, assert_records_equal, image_tolerance

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def group_launches(monkeypatch):
    """The dispatcher is opt-in (one launch per batch measured faster on the scheduler replay)."""
    monkeypatch.setenv("RL_TRACE_GROUPS", "1")


def test_grouped_batches_bit_equal_and_counted(gpu, orc):
    b = gpu.SceneBuilder(2)
    sc = gpu.Scene(b)
    w, h, n, rounds = 320, 200, 5000, 6
    units = [gpu.TraceUnit(i, w, h, seed=SEED, batch=n) for i in range(7)]
    sc.reset_batch_counter(3)
    got = {}
    for r in range(rounds):
        for u in units:
            u.render(sc, wait=False)                # 7 batches queued: they share launches
        for i, u in enumerate(units):
            u.sync()
            got[3 + r * len(units) + i] = u.mapped_photons.copy()
    ct = orc.Counters()
    total = rounds * len(units)
    want = orc.trace(b.desc(), SEED, w, h, 3 * n, total * n, orc.MATH_SPEC, False, ct)
    for k in range(total):
        assert_records_equal(got[3 + k], want[k * n:(k + 1) * n], f"batch {3 + k}")
    assert sum(u.ray_count() for u in units) == ct.rays
    launches, batches = sc.dispatch_stats()
    assert batches == total and 1 <= launches <= total


@pytest.mark.parametrize("n", [1, 7, 1023, 1024, 1025, 4097])
def test_dispatched_ragged_batch_sizes(gpu, orc, n):
    b = gpu.SceneBuilder(3)
    sc = gpu.Scene(b)
    tu = gpu.TraceUnit(0, 64, 48, seed=11, batch=n)
    for first in (0, 5 * n, (1 << 34) + 3):
        got = tu.render_range(sc, first, n)
        assert_records_equal(got, orc.trace(b.desc(), 11, 64, 48, first, n), f"n={n} first={first}")


def test_group_launch_equals_one_launch_per_batch(gpu, monkeypatch):
    b = gpu.SceneBuilder(2)
    sc = gpu.Scene(b)
    n = 40000
    tu = gpu.TraceUnit(0, 512, 512, seed=SEED, batch=n)
    through_dispatcher = tu.render_range(sc, 77 * n, n).copy()
    rays_dispatcher = tu.ray_count()
    assert sc.dispatch_stats() == (1, 1)
    monkeypatch.setenv("RL_TRACE_GROUPS", "0")
    tu2 = gpu.TraceUnit(1, 512, 512, seed=SEED, batch=n)
    direct = tu2.render_range(sc, 77 * n, n)
    assert sc.dispatch_stats() == (1, 1)                # the second render was a launch of its own
    assert_records_equal(through_dispatcher, direct, "dispatcher vs launch")
    assert rays_dispatcher == tu2.ray_count()


def test_many_batches_from_two_host_threads(gpu, orc):
    # 1400 batches from several units and two host threads
    b = gpu.SceneBuilder(1)
    sc = gpu.Scene(b)
    n, per_thread = 96, 700
    results = [{}, {}]
    errors = []

    def drive(t):
        try:
            units = [gpu.TraceUnit(10 * t + i, 32, 32, seed=5, batch=n) for i in range(4)]
            for k in range(per_thread):
                u = units[k % 4]
                first = (t * per_thread + k) * n
                results[t][first] = u.render_range(sc, first, n).copy()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=drive, args=(t,)) for t in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    want = orc.trace(b.desc(), 5, 32, 32, 0, 2 * per_thread * n)
    for t in range(2):
        for first, rec in results[t].items():
            assert_records_equal(rec, want[first:first + n], f"batch at {first}")


def test_two_scenes_keep_separate_dispatchers(gpu, orc):
    # two scenes (two Apps) in one process: records through each scene's own dispatcher, splatted
    # from the device copy, and small fused batches (which stay launches) beside them
    w, h, n = 96, 64, 20000
    b2, b3 = gpu.SceneBuilder(2), gpu.SceneBuilder(3)
    s2, s3 = gpu.Scene(b2), gpu.Scene(b3)
    t2, t3 = gpu.TraceUnit(0, w, h, seed=SEED, batch=n), gpu.TraceUnit(1, w, h, seed=SEED, batch=n)
    p2, p3 = gpu.PlotUnit(0, w, h), gpu.PlotUnit(1, w, h)
    for k in range(4):
        if k % 2:
            t2.render_fused(s2, p2, k * n, n)
            t3.render_fused(s3, p3, k * n, n)
        else:
            t2.render_range(s2, k * n, n, download=False)
            t3.render_range(s3, k * n, n, download=False)
            p2.plot(t2)
            p3.plot(t3)
    for scene_b, plot in ((b2, p2), (b3, p3)):
        ref = orc.plot(w, h, orc.trace(scene_b.desc(), SEED, w, h, 0, 4 * n))
        assert float(np.abs(plot.tristimulus_buffer - ref).max()) <= image_tolerance(ref)


def test_group_of_units_with_different_seeds_and_sizes(gpu, orc):
    # one launch carries segments of different RNG streams and lengths; units on another canvas
    # are grouped separately
    b = gpu.SceneBuilder(2)
    sc = gpu.Scene(b)
    spec = [(0, 64, 48, 5, 3000), (1, 64, 48, 6, 1), (2, 64, 48, 7, 4097), (3, 32, 32, 8, 2500),
            (4, 64, 48, 5, 777), (5, 32, 32, 9, 1024)]
    units = [gpu.TraceUnit(i, w, h, seed=seed, batch=n) for i, w, h, seed, n in spec]
    for rounds in range(3):
        sc.reset_batch_counter(10 * rounds)
        for u in units:
            u.render(sc, wait=False)
        for k, (u, (i, w, h, seed, n)) in enumerate(zip(units, spec)):
            u.sync()
            want = orc.trace(b.desc(), seed, w, h, (10 * rounds + k) * n, n)
            assert_records_equal(u.mapped_photons, want, f"round {rounds} unit {i}")


def test_scene_destroyed_before_its_units(gpu):
    # the Python wrappers are collected in any order: a unit outliving the scene it rendered is fine
    b = gpu.SceneBuilder(1)
    sc = gpu.Scene(b)
    units = [gpu.TraceUnit(i, 32, 32, seed=3, batch=512) for i in range(5)]
    for u in units:
        u.render(sc, wait=False)
    del sc                                              # launches what is queued, stops the dispatcher
    for u in units:
        u.sync()
        assert np.isfinite(u.mapped_photons["x"]).all()
    del units
