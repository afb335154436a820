"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle.

Bar (DESIGN.md "Parity"): every per-photon record (x, y, wavelength,
probability) and every per-ray function is BIT-EQUAL to the oracle in
specified-math mode; the XYZ accumulator, whose float atomics commute only up
to rounding, is within 1e-5 * max|image| per channel; gather (Kahan) is
bit-equal; tonemap is bit-equal given the same exposure and within 1 LSB
end to end.
"""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 0x5EED


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def assert_bit_equal(got, want, what):
    g, w = bits(got), bits(want)
    if not np.array_equal(g, w):
        bad = np.flatnonzero((g != w).reshape(g.shape[0], -1).any(axis=1))
        raise AssertionError(f"{what}: {bad.size}/{g.shape[0]} differ, first at {bad[:5]}: "
                             f"got {np.asarray(got)[bad[0]]} want {np.asarray(want)[bad[0]]}")


def assert_records_equal(got, want, what):
    for f in ("x", "y", "wavelength", "probability"):
        assert_bit_equal(got[f], want[f], f"{what}.{f}")


# ------------------------------------------------------------- per-function
@pytest.mark.parametrize("fn,lo,hi", [(0, -20.0, 20.0), (1, -20.0, 20.0), (0, -7000.0, 7000.0),
                                      (2, -120.0, 5.0), (3, -1.0, 1.0), (6, 1e-6, 50.0), (8, -1.5, 1.5)])
def test_specified_math_bit_equal(gpu, orc, fn, lo, hi):
    rng = np.random.default_rng(fn)
    x = rng.uniform(lo, hi, 1 << 20).astype(np.float32)
    x[:8] = [lo, hi, 0.0, -0.0, 1.0, -1.0, 0.5, -0.5]
    assert_bit_equal(gpu.debug_math(fn, x), orc.math(fn, x), f"math fn {fn}")


def test_pow_boltzmann_ior_bit_equal(gpu, orc):
    rng = np.random.default_rng(7)
    x = rng.uniform(0.0, 1.5, 1 << 18).astype(np.float32)
    y = np.full_like(x, 1.0 / 2.4)
    assert_bit_equal(gpu.debug_math(7, x, y), orc.math(7, x, y), "pow")
    wl = rng.uniform(380.0, 780.0, 1 << 18).astype(np.float32)
    t = rng.choice(np.array([5000.0, 6504.0, 7600.0, 2700.0], np.float32), wl.size)
    assert_bit_equal(gpu.debug_math(4, wl, t), orc.math(4, wl, t), "boltzmann")
    assert_bit_equal(gpu.debug_math(5, wl), orc.math(5, wl), "sf10 ior")


def test_tristimulus_bit_equal(gpu, orc):
    wl = np.concatenate([np.linspace(370.0, 790.0, 100001), [380.0, 780.0, 555.0, 377.5]]).astype(np.float32)
    assert_bit_equal(gpu.debug_tristimulus(wl), orc.tristimulus(wl), "tristimulus")


@pytest.mark.parametrize("which,w,h", [(1, 256, 256), (2, 1024, 1024), (2, 1280, 720), (3, 640, 480)])
def test_camera_rays_bit_equal(gpu, orc, which, w, h):
    b = gpu.SceneBuilder(which)
    sc = gpu.Scene(b)
    n = 20000
    rays, xy = sc.camera_rays(SEED, w, h, 12345, n)
    orays, oxy = orc.camera_rays(b.desc(), SEED, w, h, 12345, n)
    for f in ("x", "y", "wavelength", "probability"):
        assert_bit_equal(xy[f], oxy[f], f"camera draws {f}")
    assert_bit_equal(rays["origin"], orays["origin"], "camera origin")
    assert_bit_equal(rays["direction"], orays["direction"], "camera direction")


def random_rays(rng, n, extent, towards_origin=0.5):
    rays = np.zeros(n, dtype=[("origin", "<f4", 3), ("direction", "<f4", 3), ("wavelength", "<f4"),
                              ("probability", "<f4")])
    o = rng.uniform(-extent, extent, (n, 3))
    d = rng.normal(size=(n, 3))
    aim = rng.uniform(size=n) < towards_origin
    target = rng.uniform(-extent * 0.4, extent * 0.4, (n, 3))
    d[aim] = (target - o)[aim]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["origin"], rays["direction"] = o, d
    rays["wavelength"] = rng.uniform(380, 780, n)
    rays["probability"] = 1.0
    return rays


def compare_hits(got, want, what):
    assert np.array_equal(got["object"], want["object"]), (
        f"{what}: {np.count_nonzero(got['object'] != want['object'])} object mismatches")
    for f in ("distance", "position", "normal", "tangent"):
        assert_bit_equal(got[f], want[f], f"{what}.{f}")


@pytest.mark.parametrize("which,param,extent", [(1, 0, 8.0), (2, 0, 60.0), (3, 0, 15.0), (4, 512, 30.0), (4, 4096, 30.0), (4, 1500, 30.0),
                                                (6, 0, 12.0)])
def test_scene_intersect_bit_equal(gpu, orc, which, param, extent):
    b = gpu.SceneBuilder(which, param)
    sc = gpu.Scene(b)
    rays = random_rays(np.random.default_rng(which), 200000, extent)
    got, want = sc.intersect(rays), orc.intersect(b.desc(), rays)
    assert np.count_nonzero(want["object"] >= 0) > 1000
    compare_hits(got, want, f"scene {which}")


def test_intersect_edge_cases(gpu, orc):
    # the quirks the oracle KATs pin: inside-sphere and tangent rays miss, t <= 0 misses,
    # parallel rays miss, ties go to the first object (scene.rs:51)
    b = gpu.SceneBuilder()
    grey = gpu.SceneBuilder.material(gpu.MATERIAL_DIFFUSE_GREY, 0.5)
    b.object(b.plane((0, 0, -1), (0, 0, 4)), grey)
    b.object(b.sphere((0, 0, 0), 1.0), grey)
    b.object(b.plane((0, 0, -1), (0, 0, 4)), grey)                    # coincident with object 0
    b.object(b.circle((0, 0, -1), (0, 0, 3), 2.0), grey)
    b.object(b.paraboloid((0, 0, 1), (0, 0, -9), 2.0), grey)
    b.object(b.hexagonal_prism((0, 0, 1), (6, 0, 0), 3.0, 1.0, 0.0, 8.0), grey)
    b.object(b.prism((0, 1, 0), (-6, 0, 0), 2.0, 0.5, 3.0), grey)
    sc = gpu.Scene(b)
    cases = [((0, -5, 0), (0, 1, 0)), ((0, 0, 0), (0, 1, 0)), ((0, 0.5, 0), (0, 0, 1)), ((1, -5, 0), (0, 1, 0)),
             ((0, 5, 0), (0, 1, 0)), ((0, 0, 0), (1, 0, 0)), ((0, 0, 4), (0, 0, 1)), ((9, 9, 0), (0, 0, 1)),
             ((2, 0, 0), (0, 0, 1)), ((2.001, 0, 0), (0, 0, 1)), ((0, 0, 10), (0, 0, -1)), ((3, 0, 10), (0, 0, -1)),
             ((6, 0, -2), (0, 0, 1)), ((6, 0, 4), (0, 0, 1)), ((6, 0, 4), (1, 0, 0)), ((6, 0, 4), (0, 1, 0)),
             ((11, -10, 4), (0, 1, 0)), ((-10, 0, 9), (1, 0, 0)), ((-6, -4, 0.2), (0, 1, 0)),
             ((-6, 1, 0.2), (0, 1, 0)), ((-6, 1, 0.2), (1, 0, 0)), ((-20, 1, 0.2), (1, 0, 0))]
    rays = np.zeros(len(cases), dtype=gpu.RAY)
    rays["origin"] = [c[0] for c in cases]
    rays["direction"] = [c[1] for c in cases]
    rays["wavelength"], rays["probability"] = 550.0, 1.0
    want = orc.intersect(b.desc(), rays)
    compare_hits(sc.intersect(rays), want, "edge cases")
    assert want["object"][0] == 1 and want["object"][6] != 0          # sanity of the cases
    assert set(want["object"]) >= {-1, 0, 1, 3, 4, 5, 6} and 2 not in set(want["object"])
    rng = np.random.default_rng(99)
    rays = random_rays(rng, 100000, 14.0)
    compare_hits(sc.intersect(rays), orc.intersect(b.desc(), rays), "mixed primitives")


# ------------------------------------------------------------- trace records
@pytest.mark.parametrize("which,param,w,h,n", [
    (1, 0, 256, 256, 65536),        # BASELINE configs[0] in full: 256x256, 1 spp
    (2, 0, 1024, 1024, 60000),      # built-in scene (configs[1] scene, reduced photon count)
    (2, 0, 1280, 720, 20000),       # the reference's own default canvas (main.rs:47-48)
    (3, 0, 1024, 1024, 60000),      # dispersive prism (configs[2] scene)
    (4, 256, 512, 512, 20000),      # random spheres (configs[3] family, reduced)
    (6, 0, 640, 480, 60000),        # compounds over spheres: lens, dome, clipped sphere (geometry.rs:263-267,361-407)
])
def test_trace_records_bit_equal(gpu, orc, which, param, w, h, n):
    b = gpu.SceneBuilder(which, param)
    sc = gpu.Scene(b)
    tu = gpu.TraceUnit(0, w, h, seed=SEED, batch=n)
    first = 1 << 33 if which == 2 else 0          # photon ids beyond 32 bits use the high counter word
    got = tu.render_range(sc, first, n)
    ct = orc.Counters()
    want = orc.trace(b.desc(), SEED, w, h, first, n, orc.MATH_SPEC, False, ct)
    assert_records_equal(got, want, f"scene {which}")
    assert tu.ray_count() == ct.rays               # Scene::intersect calls agree exactly
    assert np.count_nonzero(want["probability"]) > 0


def test_trace_full_c4_scene_small_batch(gpu, orc):
    b = gpu.SceneBuilder(4)                        # 4096 spheres: the full configs[3] scene
    sc = gpu.Scene(b)
    tu = gpu.TraceUnit(0, 2048, 2048, seed=3, batch=3000)
    got = tu.render_range(sc, 0, 3000)
    want = orc.trace(b.desc(), 3, 2048, 2048, 0, 3000)
    assert_records_equal(got, want, "C4 4096 spheres")


def test_render_uses_the_scene_batch_counter(gpu, orc):
    # TraceUnit::render draws fresh randomness every call (trace_unit.rs:151-168); here each
    # call takes the next batch index, whichever unit makes it
    b = gpu.SceneBuilder(1)
    sc = gpu.Scene(b)
    sc.reset_batch_counter(5)
    other = gpu.Scene(b)                            # a second App in the same process: its own ids
    u0 = gpu.TraceUnit(0, 64, 64, seed=9, batch=gpu.TEST_BATCH_PHOTONS)
    u1 = gpu.TraceUnit(1, 64, 64, seed=9, batch=gpu.TEST_BATCH_PHOTONS)
    u0.render(sc)
    u1.render(sc)
    u0_first = u0.mapped_photons.copy()
    u0.render(sc)
    n = gpu.TEST_BATCH_PHOTONS
    want = orc.trace(b.desc(), 9, 64, 64, 5 * n, 3 * n)
    assert_records_equal(u0_first, want[:n], "batch 5")
    assert_records_equal(u1.mapped_photons, want[n:2 * n], "batch 6")
    assert_records_equal(u0.mapped_photons, want[2 * n:], "batch 7")
    u1.render(other)
    assert_records_equal(u1.mapped_photons, orc.trace(b.desc(), 9, 64, 64, 0, n), "batch 0 of the other scene")


def test_empty_and_ragged_batches(gpu, orc):
    b = gpu.SceneBuilder(2)
    sc = gpu.Scene(b)
    tu = gpu.TraceUnit(0, 33, 17, seed=1, batch=1)
    tu.render_range(sc, 0, 0)                       # empty: no launch, no error
    for n in (1, 31, 33, 257, 1000):
        got = tu.render_range(sc, 77, n)
        want = orc.trace(b.desc(), 1, 33, 17, 77, n)
        assert_records_equal(got, want, f"ragged n={n}")


def test_page_locked_host_buffers_and_transfer_counters(gpu, orc):
    # the shim page-locks a unit's Vec once (rl_host_register); results are the same bytes, and
    # the ABI counts what crosses the host/device boundary (bench.py's h2d/d2h bytes per step)
    import ctypes as C
    b = gpu.SceneBuilder(1)
    sc = gpu.Scene(b)
    n = 8192
    tu = gpu.TraceUnit(0, 96, 64, seed=SEED, batch=n)
    pinned = np.zeros(n, dtype=gpu.MAPPED_PHOTON)
    ptr = pinned.ctypes.data_as(C.c_void_p)
    assert gpu.lib().rl_host_register(ptr, pinned.nbytes) == gpu.RL_OK
    assert gpu.lib().rl_host_register(ptr, pinned.nbytes) == gpu.RL_OK          # idempotent per address
    gpu.reset_transfer_counters()
    tu.render_range(sc, 0, n, out=pinned)
    plain = tu.render_range(sc, 0, n).copy()
    assert gpu.transfer_counters() == (0, 2 * n * 16)
    assert_records_equal(pinned, plain, "page-locked vs pageable destination")
    assert_records_equal(pinned, orc.trace(b.desc(), SEED, 96, 64, 0, n), "page-locked destination")
    pl = gpu.PlotUnit(0, 96, 64)
    pl.plot(pinned)
    h2d, d2h = gpu.transfer_counters()
    assert h2d == n * 16 and d2h == 2 * n * 16
    pl.download()
    assert gpu.transfer_counters() == (n * 16, 2 * n * 16 + 96 * 64 * 12)
    assert gpu.lib().rl_host_unregister(ptr) == gpu.RL_OK
    assert gpu.lib().rl_host_unregister(ptr) == gpu.RL_OK                        # not registered: not an error
    assert gpu.lib().rl_host_register(None, 16) == gpu.RL_ERR_INVALID


def test_render_async_fills_the_buffer_behind_the_call(gpu, orc):
    # rl_trace_unit_render_async queues kernel + copy and returns; after sync() the buffers hold
    # exactly what the blocking call delivers, whatever the order the units were queued in
    b = gpu.SceneBuilder(2)
    sc = gpu.Scene(b)
    n = 30000
    sc.reset_batch_counter(40)
    units = [gpu.TraceUnit(i, 200, 120, seed=SEED, batch=n) for i in range(5)]
    for u in units:
        u.render(sc, wait=False)                    # batches 40..44, one per unit, all in flight
    for u in reversed(units):
        u.sync()
    want = orc.trace(b.desc(), SEED, 200, 120, 40 * n, 5 * n)
    for i, u in enumerate(units):
        assert_records_equal(u.mapped_photons, want[i * n:(i + 1) * n], f"async unit {i}")


def test_launch_geometry_does_not_change_results(gpu, monkeypatch):
    # block size, share of the SMs' block slots and the order in which a block's pool hands
    # out photon ids are scheduling only: records and ray counts are bit-equal under every policy
    sc = gpu.Scene(gpu.SceneBuilder(2))
    n, first = 150000, (1 << 33) + 99
    want = None
    for knobs in ({}, {"RL_TRACE_SMALL_PATHS": "0"}, {"RL_TRACE_SMALL_CTA": "128"},
                  {"RL_TRACE_SMALL_CTA": "384", "RL_TRACE_BLOCKS_PER_SM": "1"}, {"RL_TRACE_BLOCKS_PER_SM": "2"}):
        for k in ("RL_TRACE_SMALL_PATHS", "RL_TRACE_SMALL_CTA", "RL_TRACE_BLOCKS_PER_SM"):
            monkeypatch.delenv(k, raising=False)
        for k, v in knobs.items():
            monkeypatch.setenv(k, v)
        tu = gpu.TraceUnit(0, 640, 360, seed=SEED, batch=n)
        got = (tu.render_range(sc, first, n).copy(), tu.ray_count())
        if want is None:
            want = got
        assert_records_equal(got[0], want[0], f"policy {knobs}")
        assert got[1] == want[1]


def test_units_driven_from_concurrent_threads(gpu, orc):
    # app.rs:95-111: C worker threads, each with its own trace unit, render at the same time
    # (the small launches then share the SMs); every unit still returns exactly its batch
    import threading
    b = gpu.SceneBuilder(2)
    sc = gpu.Scene(b)
    n, rounds, workers = 20000, 3, 6
    units = [gpu.TraceUnit(i, 320, 200, seed=SEED, batch=n) for i in range(workers)]
    got = {}

    def work(i):
        for r in range(rounds):
            got[(i, r)] = units[i].render_range(sc, (r * workers + i) * n, n).copy()

    threads = [threading.Thread(target=work, args=(i,)) for i in range(workers)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    want = orc.trace(b.desc(), SEED, 320, 200, 0, rounds * workers * n)
    for (i, r), rec in got.items():
        k = r * workers + i
        assert_records_equal(rec, want[k * n:(k + 1) * n], f"unit {i} round {r}")


# ------------------------------------------------------------------- splat
def image_tolerance(ref):
    return 1e-5 * float(np.abs(ref).max()) + 1e-12


@pytest.mark.parametrize("which,w,h,n", [(1, 256, 256, 65536), (2, 128, 96, 200000), (3, 64, 64, 100000)])
def test_plot_matches_oracle(gpu, orc, which, w, h, n):
    b = gpu.SceneBuilder(which)
    sc = gpu.Scene(b)
    tu = gpu.TraceUnit(0, w, h, seed=SEED, batch=n)
    photons = tu.render_range(sc, 0, n)
    want = orc.plot(w, h, orc.trace(b.desc(), SEED, w, h, 0, n))
    tol = image_tolerance(want)
    # (1) PlotUnit::plot on the host slice, exactly as app.rs:139 calls it
    p_host = gpu.PlotUnit(0, w, h)
    p_host.plot(photons)
    # (2) on the records left on the device
    p_dev = gpu.PlotUnit(1, w, h)
    p_dev.plot(tu)
    # (3) fused trace + splat
    p_fused = gpu.PlotUnit(2, w, h)
    tu.render_fused(sc, p_fused, 0, n)
    for name, p in (("host", p_host), ("device", p_dev), ("fused", p_fused)):
        img = p.tristimulus_buffer
        assert img.shape == (h, w, 3)
        err = float(np.abs(img - want).max())
        assert err <= tol, f"{name}: max |diff| {err} > {tol}"
    # plotting twice accumulates (plot_unit.rs:80-83); clear() resets (plot_unit.rs:98-102)
    p_host.plot(photons)
    assert float(np.abs(p_host.tristimulus_buffer - 2 * want).max()) <= 2 * tol
    p_host.clear()
    assert not p_host.tristimulus_buffer.any()


def test_plot_edge_photons(gpu, orc):
    # borders, exact pixel centres, zero and huge probabilities, lambda at the table ends
    w, h = 16, 8
    ph = np.zeros(12, dtype=gpu.MAPPED_PHOTON)
    ph["x"] = [-1, 1, -1, 1, 0, 0.123, -0.999999, 1, 0, 0, 0.5, -0.5]
    ph["y"] = [-0.5, 0.5, 0.5, -0.5, 0, 0.0371, 0.499999, 0.25, 0, 0, 0.1, -0.1]
    ph["probability"] = [1, 1, 1, 1, 1, 2.5, 0.1, 3, 0, 1e6, 1e-20, 1]
    ph["wavelength"] = [380, 780, 555, 555, 382.5, 600, 779.99, 400, 500, 450, 650, 777.5]
    p = gpu.PlotUnit(0, w, h)
    p.plot(ph)
    want = orc.plot(w, h, ph)
    assert np.allclose(p.tristimulus_buffer, want, rtol=1e-6, atol=1e-30)


# ------------------------------------------------------------------ gather
def test_gather_kahan_bit_equal(gpu, orc):
    w, h = 67, 31        # w*h not a multiple of 4: exercises the tail path
    rng = np.random.default_rng(5)
    g = gpu.GatherUnit(w, h)
    acc = np.zeros((h, w, 3), np.float32)
    comp = np.zeros((h, w, 3), np.float32)
    for i in range(6):
        px = (rng.uniform(0, 1, (h, w, 3)) * 10.0 ** rng.integers(-6, 4)).astype(np.float32)
        g.accumulate(px)
        orc.gather_accumulate(acc, comp, px)
    got_acc, got_comp = g.download(with_compensation=True)
    assert_bit_equal(got_acc.reshape(-1), acc.reshape(-1), "gather acc")
    assert_bit_equal(got_comp.reshape(-1), comp.reshape(-1), "gather comp")


def test_gather_from_plot_units_and_clear(gpu, orc):
    # app.rs:143-151: accumulate every done plot unit, clear it, save
    w, h, n = 64, 64, 50000
    b = gpu.SceneBuilder(2)
    sc = gpu.Scene(b)
    tu = gpu.TraceUnit(0, w, h, seed=SEED, batch=n)
    plots = [gpu.PlotUnit(i, w, h) for i in range(3)]
    for i, p in enumerate(plots):
        tu.render_fused(sc, p, i * n, n)
    imgs = [p.tristimulus_buffer for p in plots]
    g = gpu.GatherUnit(w, h)
    acc = np.zeros((h, w, 3), np.float32)
    comp = np.zeros_like(acc)
    for p, img in zip(plots, imgs):
        g.accumulate(p, clear=True)
        orc.gather_accumulate(acc, comp, img)
    got_acc, got_comp = g.download(with_compensation=True)
    assert_bit_equal(got_acc.reshape(-1), acc.reshape(-1), "gather acc")
    assert_bit_equal(got_comp.reshape(-1), comp.reshape(-1), "gather comp")
    assert all(not p.tristimulus_buffer.any() for p in plots)
    # the multi-buffer entry point (peer buffers in the multi-GPU case) gives the same sums
    for i, p in enumerate(plots):
        tu.render_fused(sc, p, i * n, n)
    imgs2 = [p.tristimulus_buffer for p in plots]
    g2 = gpu.GatherUnit(w, h)
    for p in plots:
        p.sync()
    g2.accumulate_device([p.device_buffer()[0] for p in plots])
    acc2 = np.zeros_like(acc)
    comp2 = np.zeros_like(acc)
    for img in imgs2:
        orc.gather_accumulate(acc2, comp2, img)
    assert_bit_equal(g2.download().reshape(-1), acc2.reshape(-1), "gather (device buffers)")


def test_buffer_raw_checkpoint_format(gpu, tmp_path):
    # gather_unit.rs:68-92: accumulator then compensation, 12 raw bytes per pixel, no header
    w, h = 40, 30
    rng = np.random.default_rng(3)
    g = gpu.GatherUnit(w, h)
    for _ in range(3):
        g.accumulate(rng.uniform(0, 5, (h, w, 3)).astype(np.float32))
    acc, comp = g.download(with_compensation=True)
    path = str(tmp_path / "buffer.raw")
    g.save(path)
    assert os.path.getsize(path) == 24 * w * h
    raw = np.fromfile(path, dtype="<f4")
    assert np.array_equal(raw[: 3 * w * h], acc.reshape(-1)) and np.array_equal(raw[3 * w * h:], comp.reshape(-1))
    # resume (GatherUnit::new reads the file if it exists, gather_unit.rs:43)
    g2 = gpu.GatherUnit(w, h, resume_path=path)
    acc2, comp2 = g2.download(with_compensation=True)
    assert np.array_equal(acc2, acc) and np.array_equal(comp2, comp)
    # a missing file is not an error for the constructor; a short file fills a prefix (read.rs:20-32)
    g3 = gpu.GatherUnit(w, h, resume_path=str(tmp_path / "absent.raw"))
    assert not g3.download().any()
    short = str(tmp_path / "short.raw")
    raw[: w * h].tofile(short)
    g3.load(short)
    got = g3.download().reshape(-1)
    assert np.array_equal(got[: w * h], acc.reshape(-1)[: w * h]) and not got[w * h:].any()
    with pytest.raises(gpu.RlError) as e:
        g3.load(str(tmp_path / "absent.raw"))
    assert e.value.code == gpu.RL_ERR_IO


def test_buffer_raw_background_writer(gpu, tmp_path):
    # app.rs:151 saves after every gather: only the first save to a path writes in the caller,
    # later ones hand a snapshot to the unit's writer thread (newest wins); the file on disk is
    # always one complete snapshot and flush / load / destroy make it current
    w, h = 256, 256
    rng = np.random.default_rng(11)
    path = str(tmp_path / "buffer.raw")
    other = gpu.GatherUnit(w, h)
    with pytest.raises(gpu.RlError) as e:
        other.save(str(tmp_path / "no_such_dir" / "buffer.raw"))    # "failed to open file" in the caller
    assert e.value.code == gpu.RL_ERR_IO
    g = gpu.GatherUnit(w, h)
    states = []
    for i in range(12):
        g.accumulate(rng.uniform(0, 5, (h, w, 3)).astype(np.float32))
        acc, comp = g.download(with_compensation=True)
        states.append(np.concatenate([acc.reshape(-1), comp.reshape(-1)]))
        g.save(path, wait=False)
        # whatever the writer is doing, a reader sees a whole snapshot of some earlier or the current state
        raw = np.fromfile(path, dtype="<f4")
        assert raw.size == 6 * w * h
        assert any(np.array_equal(raw, st) for st in states)
    g.flush()
    assert np.array_equal(np.fromfile(path, dtype="<f4"), states[-1])
    # two files in turn: a newer snapshot only replaces a queued one for the same file
    path2 = str(tmp_path / "second.raw")
    last = {}
    for i in range(6):
        g.accumulate(rng.uniform(0, 5, (h, w, 3)).astype(np.float32))
        acc, comp = g.download(with_compensation=True)
        target = path if i % 2 else path2
        last[target] = np.concatenate([acc.reshape(-1), comp.reshape(-1)])
        g.save(target, wait=False)
    g.flush()
    for target, want in last.items():
        assert np.array_equal(np.fromfile(target, dtype="<f4"), want)
    # load waits for a queued save of the same file; destroy writes out what is still queued
    g.accumulate(rng.uniform(0, 5, (h, w, 3)).astype(np.float32))
    acc, comp = g.download(with_compensation=True)
    g.save(path, wait=False)
    g.load(path)
    acc2, comp2 = g.download(with_compensation=True)
    assert np.array_equal(acc2, acc) and np.array_equal(comp2, comp)
    g.accumulate(rng.uniform(0, 5, (h, w, 3)).astype(np.float32))
    acc, comp = g.download(with_compensation=True)
    g.save(path, wait=False)
    del g
    raw = np.fromfile(path, dtype="<f4")
    assert np.array_equal(raw[: 3 * w * h], acc.reshape(-1)) and np.array_equal(raw[3 * w * h:], comp.reshape(-1))
    assert not os.path.exists(path + ".tmp")


def test_buffer_raw_save_interval(gpu, tmp_path):
    # the writer starts at most one file per save interval: a burst of gathers (a GPU gathers a
    # hundred times a second, app.rs:143-152) costs device-side snapshots, not host traffic, and
    # flush still leaves the newest state on disk
    w, h = 256, 256
    snapshot = 24 * w * h
    rng = np.random.default_rng(12)
    path = str(tmp_path / "buffer.raw")
    g = gpu.GatherUnit(w, h)
    frame = rng.uniform(0, 5, (h, w, 3)).astype(np.float32)
    g.save(path)                                     # proves the path (synchronous)
    g.set_save_interval(5.0)
    g.accumulate(frame)
    g.save(path, wait=False)                         # the first background write starts at once
    time.sleep(0.5)
    gpu.reset_transfer_counters()
    for _ in range(25):
        g.accumulate(frame)
        g.save(path, wait=False)
    h2d, d2h = gpu.transfer_counters()
    assert h2d == 25 * 12 * w * h and d2h == 0       # inside the interval: nothing came down
    acc, comp = g.download(with_compensation=True)
    t0 = time.time()
    g.flush()                                        # does not wait for the interval
    assert time.time() - t0 < 2.0
    raw = np.fromfile(path, dtype="<f4")
    assert np.array_equal(raw[: 3 * w * h], acc.reshape(-1)) and np.array_equal(raw[3 * w * h:], comp.reshape(-1))
    _, d2h = gpu.transfer_counters()
    assert d2h == snapshot + 2 * 12 * w * h          # one snapshot + the download above
    with pytest.raises(gpu.RlError):
        g.set_save_interval(-1.0)


# ----------------------------------------------------------------- tonemap
def test_tonemap_matches_oracle(gpu, orc):
    w, h, n = 96, 64, 300000
    b = gpu.SceneBuilder(2)
    sc = gpu.Scene(b)
    tu = gpu.TraceUnit(0, w, h, seed=SEED, batch=n)
    p = gpu.PlotUnit(0, w, h)
    tu.render_fused(sc, p, 0, n)
    g = gpu.GatherUnit(w, h)
    g.accumulate(p, clear=True)
    xyz = g.tristimulus_buffer
    t = gpu.TonemapUnit(w, h)
    rgb_dev = t.tonemap(g).copy()                   # tonemap(&gather.tristimulus_buffer), app.rs:157
    exposure = t.last_exposure
    rgb_host = t.tonemap(xyz).copy()                # the same through the host slice
    assert np.array_equal(rgb_dev, rgb_host)
    # exposure: the reference folds in f32 sequentially, the GPU reduces in f64
    ref_exposure = orc.find_exposure(xyz)
    assert abs(exposure / ref_exposure - 1.0) < 1e-4
    # given the same exposure the map is bit-equal to the oracle's specified-math mode ...
    assert np.array_equal(rgb_dev, orc.tonemap(xyz, orc.MATH_SPEC, exposure))
    # ... and end to end within 1 LSB of the reference's libm arithmetic
    ref = orc.tonemap(xyz, orc.MATH_LIBM)
    assert np.abs(rgb_dev.astype(int) - ref.astype(int)).max() <= 1
    assert rgb_dev.max() > 100                      # not a black frame


def test_tonemap_reference_fold_is_bit_equal(gpu, orc):
    # RL_EXPOSURE_REFERENCE_FOLD: find_exposure as the reference's two sequential f32 folds
    # (tonemap_unit.rs:55-69): the exposure, and with it the whole image, bit-equal to the oracle
    w, h, n = 320, 200, 400000
    b = gpu.SceneBuilder(2)
    sc = gpu.Scene(b)
    tu = gpu.TraceUnit(0, w, h, seed=SEED, batch=n)
    p = gpu.PlotUnit(0, w, h)
    tu.render_fused(sc, p, 0, n)
    g = gpu.GatherUnit(w, h)
    g.accumulate(p, clear=True)
    xyz = g.tristimulus_buffer
    t = gpu.TonemapUnit(w, h)
    t.set_exposure_mode(True)
    rgb = t.tonemap(g).copy()
    assert np.float32(t.last_exposure).view(np.uint32) == np.float32(orc.find_exposure(xyz)).view(np.uint32)
    assert np.array_equal(rgb, orc.tonemap(xyz, orc.MATH_SPEC))        # exposure found by the oracle's own fold
    # ... including the reference's accident: a near-constant image has a NaN exposure and goes black
    flat = np.full((1024, 1024, 3), 0.1, dtype=np.float32)
    assert np.isnan(orc.find_exposure(flat))
    t2 = gpu.TonemapUnit(1024, 1024)
    t2.set_exposure_mode(True)
    black = t2.tonemap(flat)
    assert np.isnan(t2.last_exposure) and not black.any()
    assert np.array_equal(black, orc.tonemap(flat, orc.MATH_SPEC))
    t2.set_exposure_mode(False)                                         # the default reduces in f64: finite
    assert np.isfinite(t2.tonemap(flat).astype(np.float32)).all() and np.isfinite(t2.last_exposure)


def test_tonemap_constant_image(gpu, orc):
    xyz = np.full((16, 16, 3), 0.5, dtype=np.float32)
    t = gpu.TonemapUnit(16, 16)
    rgb = t.tonemap(xyz)
    assert t.last_exposure == 0.5
    assert np.array_equal(rgb, orc.tonemap(xyz, orc.MATH_SPEC))


# ----------------------------------------------------- size-independent laws
def test_full_resolution_properties(gpu, orc):
    # BASELINE configs[1] canvas (1024x1024) with 2^22 photons: too many for the oracle to
    # trace in seconds, so check laws that hold at any size.
    w = h = 1024
    n = 1 << 22
    b = gpu.SceneBuilder(2)
    sc = gpu.Scene(b)
    tu = gpu.TraceUnit(0, w, h, seed=SEED, batch=n)
    whole = gpu.PlotUnit(0, w, h)
    tu.render_fused(sc, whole, 0, n)
    rays_whole = tu.ray_count()
    img = whole.tristimulus_buffer
    # (1) linearity / partition invariance: splitting the photon range changes only the
    #     order of float additions
    parts = gpu.PlotUnit(1, w, h)
    tu2 = gpu.TraceUnit(1, w, h, seed=SEED, batch=n)
    cuts = [0, 1000, 1 << 20, (1 << 21) + 12345, n]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        tu2.render_fused(sc, parts, lo, hi - lo)
    img2 = parts.tristimulus_buffer
    assert tu2.ray_count() == rays_whole
    assert float(np.abs(img - img2).max()) <= image_tolerance(img)
    # (2) checksum of checksums: the image total equals the sum over photons of
    #     cie(lambda) * probability (bilinear weights sum to 1, plot_unit.rs:72-75)
    photons = tu.render_range(sc, 0, n)
    cie = orc.tristimulus(photons["wavelength"]).astype(np.float64)
    total = (cie * photons["probability"].astype(np.float64)[:, None]).sum(axis=0)
    assert np.allclose(img.astype(np.float64).sum(axis=(0, 1)), total, rtol=2e-4)
    # (3) a sample of the records is bit-equal to the oracle
    idx = slice(3_000_000, 3_000_000 + 4096)
    want = orc.trace(b.desc(), SEED, w, h, 3_000_000, 4096)
    assert_records_equal(photons[idx], want, "sample of the full batch")
    # (4) records-then-plot equals fused
    p3 = gpu.PlotUnit(2, w, h)
    p3.plot(tu)
    assert float(np.abs(p3.tristimulus_buffer - img).max()) <= image_tolerance(img)
    assert np.all(img >= 0.0) and np.isfinite(img).all()


def test_request_larger_than_one_launch(gpu):
    # the kernel indexes the photons of a launch with 32 bits; a request of more than 2^31 photons
    # (C3 is 2^30 per frame, C4 2^31, C5 2^36) is split into launches that continue the id range
    w = h = 256
    n = (1 << 31) + 4099
    sc = gpu.Scene(gpu.SceneBuilder(1))            # sphere + emissive plane: 1.09 rays per photon
    one = gpu.PlotUnit(0, w, h)
    tu = gpu.TraceUnit(0, w, h, seed=SEED, batch=1)
    tu.render_fused(sc, one, 7, n)
    rays_one = tu.ray_count()
    two = gpu.PlotUnit(1, w, h)
    tu2 = gpu.TraceUnit(1, w, h, seed=SEED, batch=1)
    tu2.render_fused(sc, two, 7, 1 << 31)
    tu2.render_fused(sc, two, 7 + (1 << 31), 4099)
    assert tu2.ray_count() == rays_one
    a, b = one.tristimulus_buffer, two.tristimulus_buffer
    # 2^31 float atomics per image in different orders: sums of ~3e4 terms per pixel
    assert float(np.abs(a - b).max()) <= 2e-4 * float(np.abs(a).max())
    assert rays_one > n


def test_c5_canvas_properties(gpu, orc):
    # BASELINE configs[4] canvas (4096x4096, built-in scene): 201 MB frames, far more pixels than
    # photons here -- the laws that do not depend on size, plus gather + clear + tonemap at that size
    w = h = 4096
    n = 1 << 21
    b = gpu.SceneBuilder(2)
    sc = gpu.Scene(b)
    tu = gpu.TraceUnit(0, w, h, seed=SEED, batch=n)
    fused = gpu.PlotUnit(0, w, h)
    tu.render_fused(sc, fused, 5 * n, n)
    photons = tu.render_range(sc, 5 * n, n)
    img = fused.tristimulus_buffer
    cie = orc.tristimulus(photons["wavelength"]).astype(np.float64)
    total = (cie * photons["probability"].astype(np.float64)[:, None]).sum(axis=0)
    assert np.allclose(img.astype(np.float64).sum(axis=(0, 1)), total, rtol=2e-4)
    want = orc.trace(b.desc(), SEED, w, h, 5 * n + 777, 2048)
    assert_records_equal(photons[777:777 + 2048], want, "sample at 4096^2")
    # the oracle's splat of the same records, compared where it matters: every touched pixel
    ref = orc.plot(w, h, photons)
    assert float(np.abs(img - ref).max()) <= image_tolerance(ref)
    # gather the frame twice (Kahan), clearing the plot unit the second time
    g = gpu.GatherUnit(w, h)
    g.accumulate(fused)
    g.accumulate(fused, clear=True)
    acc = np.zeros_like(ref)
    comp = np.zeros_like(ref)
    orc.gather_accumulate(acc, comp, img)
    orc.gather_accumulate(acc, comp, img)
    assert_bit_equal(g.download().reshape(-1), acc.reshape(-1), "gather at 4096^2")
    assert not fused.tristimulus_buffer.any()
    rgb = gpu.TonemapUnit(w, h).tonemap(g)
    assert rgb.shape == (h, w, 3) and rgb.any()


def test_scheduler_call_pattern(gpu, orc):
    # The smoke sequence of the reference's only integration test (main.rs:69-74, app.rs:75-90,
    # task_scheduler.rs:127-182 with concurrency 1): Trace(0), Trace(1), Plot(plot0, [0, 1]),
    # Trace(2), Trace(0) -- then, beyond what the reference exercises, gather + tonemap.
    w, h = 1280, 720
    n = gpu.TEST_BATCH_PHOTONS
    b = gpu.SceneBuilder(2)
    sc = gpu.Scene(b)
    traces = [gpu.TraceUnit(i, w, h, seed=SEED, batch=n) for i in range(3)]
    plot0 = gpu.PlotUnit(0, w, h)
    traces[0].render(sc)
    traces[1].render(sc)
    for u in (traces[0], traces[1]):
        plot0.plot(u.mapped_photons)
    traces[2].render(sc)
    traces[0].render(sc)
    gather = gpu.GatherUnit(w, h)
    gather.accumulate(plot0.tristimulus_buffer)
    plot0.clear()
    tm = gpu.TonemapUnit(w, h)
    rgb = tm.tonemap(gather.tristimulus_buffer)
    want = orc.trace(b.desc(), SEED, w, h, 0, 4 * n)
    assert_records_equal(traces[2].mapped_photons, want[2 * n:3 * n], "trace unit 2")
    assert_records_equal(traces[0].mapped_photons, want[3 * n:], "trace unit 0, second batch")
    ref_img = orc.plot(w, h, want[:2 * n])
    assert float(np.abs(gather.tristimulus_buffer - ref_img).max()) <= image_tolerance(ref_img)
    assert rgb.shape == (h, w, 3)


def test_error_behaviour(gpu):
    b = gpu.SceneBuilder(1)
    sc = gpu.Scene(b)
    tu = gpu.TraceUnit(0, 32, 32)
    with pytest.raises(gpu.RlError) as e:
        tu.render_fused(sc, gpu.PlotUnit(0, 16, 16), 0, 10)     # canvas mismatch
    assert e.value.code == gpu.RL_ERR_INVALID
    with pytest.raises(gpu.RlError):
        gpu.TraceUnit(0, 0, 32)
    # the children of a compound are Volumes (geometry.rs:379): a plane is a Surface only
    bad = gpu.SceneBuilder()
    bad.object(bad.compound(bad.sphere((0, 0, 0), 1.0), bad.plane((0, 0, 1), (0, 0, 0))),
               gpu.SceneBuilder.material(gpu.MATERIAL_SF10_GLASS))
    with pytest.raises(gpu.RlError) as e:
        gpu.Scene(bad)
    assert e.value.code == gpu.RL_ERR_UNSUPPORTED


# --------------------------------------------------------------------- culling
@pytest.mark.parametrize("which,param,w,h,n", [(2, 0, 1024, 1024, 1 << 21), (3, 0, 1024, 1024, 1 << 20), (6, 0, 640, 480, 1 << 19),
                                               (4, 0, 2048, 2048, 1 << 17), (1, 0, 256, 256, 1 << 18)])
def test_culls_are_result_preserving(gpu, which, param, w, h, n):
    # every ray of n traced paths: culled Scene::intersect == brute force over all primitives
    # (object, distance bits, primitive), evaluated side by side on the device
    sc = gpu.Scene(gpu.SceneBuilder(which, param))
    rays, bad = sc.cull_check(SEED + which, w, h, 7 * n, n)
    assert rays >= n
    assert bad == 0, f"{bad} of {rays} rays differ between culled and brute-force intersect"


# ------------------------------------------- the reference's arithmetic, image level
def test_image_matches_reference_arithmetic(gpu, orc):
    """north_star: "matches the reference CPU path on the same scene and RNG seed within a stated
    float tolerance on the XYZ accumulator".  GPU frame (specified arithmetic) against the
    oracle in LIBM mode (glibc math = the arithmetic of the Rust binary), built-in scene, 256^2,
    2^24 photons of the same ids; tolerance and rationale in tests/stat_parity.py."""
    import stat_parity
    b = gpu.SceneBuilder(2)
    sc = gpu.Scene(b)
    w = h = 256
    threads = orc.hardware_threads()
    n = (1 << 24) if threads >= 8 else (1 << 22)
    k = 16
    tu = gpu.TraceUnit(0, w, h, seed=SEED, batch=n // k)
    subs = []
    for i in range(k):
        pl = gpu.PlotUnit(i, w, h)
        tu.render_fused(sc, pl, i * (n // k), n // k)
        subs.append(pl.tristimulus_buffer)
    ref, _, _ = orc.render_mt(b.desc(), SEED, w, h, 0, n, threads, mode=orc.MATH_LIBM, batch=n // (threads * 8))
    print(stat_parity.check(subs, ref, f"GPU vs LIBM oracle, {n} photons"))
    # per-photon: the RNG-only fields are bit-equal, the probabilities differ for a bounded share
    m = 100000
    got = tu.render_range(sc, 0, m)
    want = orc.trace(b.desc(), SEED, w, h, 0, m, orc.MATH_LIBM)
    for f in ("x", "y", "wavelength"):
        assert_bit_equal(got[f], want[f], f"libm.{f}")
    differ = np.count_nonzero(~np.isclose(got["probability"], want["probability"], rtol=1e-4, atol=1e-7)) / m
    print(f"photons whose probability differs from the LIBM oracle's by more than 1e-4: {differ:.4f}")
    assert differ < 0.03


def test_keyframe_camera_bit_equal(gpu, orc):
    # the third camera model (rl_b200.h: RL_CAMERA_KEYFRAMES): rays and whole paths against the oracle
    from test_oracle_kat import keyframe_scene
    b = keyframe_scene(gpu)
    sc = gpu.Scene(b)
    n = 50000
    got_r, got_xy = sc.camera_rays(SEED, 320, 200, 123, n)
    want_r, want_xy = orc.camera_rays(b.desc(), SEED, 320, 200, 123, n)
    assert_bit_equal(got_r["origin"], want_r["origin"], "keyframe camera origin")
    assert_bit_equal(got_r["direction"], want_r["direction"], "keyframe camera direction")
    tu = gpu.TraceUnit(0, 320, 200, seed=SEED, batch=n)
    assert_records_equal(tu.render_range(sc, 123, n), orc.trace(b.desc(), SEED, 320, 200, 123, n), "keyframe scene")


def test_scene_beyond_shared_memory(gpu, orc):
    # 20 000 spheres: their pre-test records (320 KB) do not fit an SM's shared memory and stay in
    # global memory; same records, same ray count, culls still result-preserving
    b = gpu.SceneBuilder(4, 20000)
    sc = gpu.Scene(b)
    n = 3000
    tu = gpu.TraceUnit(0, 256, 256, seed=SEED, batch=n)
    got = tu.render_range(sc, 0, n)
    ct = orc.Counters()
    want = orc.trace(b.desc(), SEED, 256, 256, 0, n, orc.MATH_SPEC, False, ct)
    assert_records_equal(got, want, "20000 spheres")
    assert tu.ray_count() == ct.rays
    rays, bad = sc.cull_check(SEED, 256, 256, 0, 1 << 15)
    assert bad == 0 and rays >= 1 << 15
