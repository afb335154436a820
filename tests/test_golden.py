"""Committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py):
regression pins of the oracle in specified-math mode -- not reference outputs,
which cannot be produced here (see the script's docstring).  The oracle is
checked on CPU, the CUDA path on the GPU, both bit-for-bit."""
import glob
import os

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def load(path):
    z = np.load(path)
    which, param, w, h, seed, first, n = (int(v) for v in z["meta"])
    return z, which, param, w, h, seed, first, n


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def test_fixtures_exist():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_golden(pkg, orc, path):
    z, which, param, w, h, seed, first, n = load(path)
    desc = pkg.SceneBuilder(which, param).desc()
    ct = orc.Counters()
    photons = orc.trace(desc, seed, w, h, first, n, orc.MATH_SPEC, False, ct)
    assert same_bits(photons, z["photons"]) and ct.rays == int(z["rays"])
    rays, _ = orc.camera_rays(desc, seed, w, h, first, 32)
    assert same_bits(rays, z["camera_rays"])
    assert same_bits(orc.intersect(desc, rays), z["hits"])
    img = z["image"]
    assert same_bits(orc.plot(img.shape[1], img.shape[0], photons), img)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_gpu_matches_golden(gpu, path):
    z, which, param, w, h, seed, first, n = load(path)
    sc = gpu.Scene(gpu.SceneBuilder(which, param))
    tu = gpu.TraceUnit(0, w, h, seed=seed, batch=n)
    assert same_bits(tu.render_range(sc, first, n), z["photons"])
    assert tu.ray_count() == int(z["rays"])
    rays, _ = sc.camera_rays(seed, w, h, first, 32)
    assert same_bits(rays, z["camera_rays"])
    assert same_bits(sc.intersect(rays), z["hits"])
    img = z["image"]
    p = gpu.PlotUnit(0, img.shape[1], img.shape[0])
    p.plot(z["photons"])
    assert np.allclose(p.tristimulus_buffer, img, rtol=1e-5, atol=1e-7)
