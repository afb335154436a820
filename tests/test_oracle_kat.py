"""Known-answer tests that pin the oracle to the reference source.

The reference ships no golden vectors for this path (src/main.rs:69-74 asserts
nothing; its RNG is unseeded), so the pins are closed-form values derived from
the reference's own formulas and tables, cited per test.
"""
import math

import numpy as np
import pytest


def ray(o, d, wl=550.0):
    r = np.zeros(1, dtype=[("origin", "<f4", 3), ("direction", "<f4", 3), ("wavelength", "<f4"),
                           ("probability", "<f4")])
    r["origin"][0] = o
    r["direction"][0] = d
    r["wavelength"] = wl
    r["probability"] = 1.0
    return r


def one_object_scene(pkg, make_surface, material=None):
    b = pkg.SceneBuilder()
    s = make_surface(b)
    b.object(s, material or pkg.SceneBuilder.material(pkg.MATERIAL_DIFFUSE_GREY, 0.8))
    return b


# ------------------------------------------------------------------------ RNG
def test_philox4x32_10_known_answers(orc):
    # Random123 kat_vectors for philox4x32-10
    assert orc.philox((0, 0), (0, 0, 0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox((0xffffffff,) * 2, (0xffffffff,) * 4) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox((0xa4093822, 0x299f31d0), (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)) == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_draw_stream_layout_and_distributions(orc):
    # sequential draws (what precedes the first bounce): draw i = word i%4 of block i/4, counter
    # (id_lo, id_hi, block, 0), key = seed; the draws of bounce j restart at block 2 + j -- that
    # part of the layout is pinned by tests/test_oracle_path_cross_check.py
    seed, photon = 0x0123456789ABCDEF, 0x1_0000_0007
    words = []
    for block in range(3):
        words += orc.philox((seed & 0xffffffff, seed >> 32), (photon & 0xffffffff, photon >> 32, block, 0))
    closed = orc.draws(seed, photon, np.zeros(12, np.uint8))
    half = orc.draws(seed, photon, np.ones(12, np.uint8))
    for i, w in enumerate(words):
        # monte_carlo.rs:25-28 (Closed01<f32>) and :37 (f32), rand 0.3.11 semantics
        assert closed[i] == np.float32(np.float32(w >> 8) / np.float32(16777215.0))
        assert half[i] == np.float32(w >> 8) * np.float32(2.0 ** -24)
    big = orc.draws(7, 7, np.zeros(4096, np.uint8))
    assert 0.0 <= big.min() and big.max() <= 1.0 and abs(big.mean() - 0.5) < 0.03
    # the closed distribution reaches 1.0 exactly, the half-open one cannot
    assert np.float32(16777215) / np.float32(16777215.0) == np.float32(1.0)
    assert np.float32(16777215) * np.float32(2.0 ** -24) < np.float32(1.0)


# --------------------------------------------------------------------- cie1931
def test_tristimulus_table_and_lerp(orc):
    # cie1931.rs:20-48 and the tables at :53-305
    t = orc.tristimulus([380.0, 555.0, 780.0, 382.5, 777.5])
    f32 = lambda v: np.asarray(v, dtype=np.float32)
    assert np.array_equal(t[0], f32([0.001368, 0.000039, 0.006450]))
    assert np.array_equal(t[1], f32([0.512050, 1.0, 0.005750]))
    assert np.array_equal(t[2], f32([0.000042, 0.000015, 0.0]))   # index == 80 branch
    assert np.allclose(t[3], [(0.001368 + 0.002236) / 2, (0.000039 + 0.000064) / 2, (0.006450 + 0.010550) / 2],
                       rtol=1e-6)
    # out of range -> zero (cie1931.rs:25-27)
    assert np.all(orc.tristimulus([300.0, 900.0]) == 0.0)
    # index == -1 branch (cie1931.rs:28-30): 377.5 nm -> X[0] * 0.5
    assert np.allclose(orc.tristimulus([377.5])[0], [0.001368 * 0.5, 0.000039 * 0.5, 0.006450 * 0.5], rtol=1e-6)


# ------------------------------------------------------------------- materials
@pytest.mark.parametrize("mode", [0, 1])
def test_blackbody_normalised_at_wien_peak(pkg, orc, mode):
    # material.rs:92-105: get_intensity(Wien peak) == intensity; the Planck law is the
    # per-frequency form, so the spectrum rises towards red (material.rs:61-74)
    for kelvins, intensity in ((6504.0, 1.0), (7600.0, 0.6), (5000.0, 0.6)):
        m = pkg.SceneBuilder.blackbody(kelvins, intensity)
        peak_nm = 2.897772126e-3 / kelvins * 1e9
        v = orc.blackbody_intensity(kelvins, m.p1, [peak_nm, 380.0, 780.0], mode)
        assert abs(v[0] / intensity - 1.0) < 1e-4
    m = pkg.SceneBuilder.blackbody(6504.0, 1.0)
    v = orc.blackbody_intensity(6504.0, m.p1, [380.0, 780.0], mode)
    assert abs(v[0] - 0.682) < 2e-3 and abs(v[1] - 1.653) < 2e-3
    # independent evaluation of the same formula in Python floats
    h, k, c = 6.62606957e-34, 1.3806488e-23, 299792458.0
    f = c / (600.0 * 1.0e-9)
    want = (2.0 * h * f * f * f) / (c * c * (math.exp(h * f / (k * 6504.0)) - 1.0))
    got = orc.math(4, [600.0], [6504.0], mode)[0]
    assert abs(got / want - 1.0) < 1e-6


def test_sf10_sellmeier(orc):
    # material.rs:203-213
    n = orc.math(5, [380.0, 486.1, 589.3, 656.3, 780.0])
    assert np.allclose(n, [1.86074, 1.80652, 1.78446, 1.77595, 1.76583], atol=2e-5)
    w2 = float(np.float32(np.float32(589.3) * np.float32(589.3)) * np.float32(1.0e-6))
    want = math.sqrt(1.0 + 1.737596950 * w2 / (w2 - 0.0131887070) + 0.313747346 * w2 / (w2 - 0.0623068142)
                     + 1.898781010 * w2 / (w2 - 155.23629000))
    assert n[2] == np.float32(want)


# -------------------------------------------------------------------- geometry
def test_sphere_hit_inside_and_tangent(pkg, orc):
    # geometry.rs:204-261: outside -> t = (b - sqrt(disc)) / 2; a ray starting inside never
    # hits (t1 <= 0 and the t2 branch is unreachable); a tangent ray (t1 == t2) misses
    d = one_object_scene(pkg, lambda b: b.sphere((0, 0, 0), 1.0)).desc()
    hit = orc.intersect(d, ray((0, -5, 0), (0, 1, 0)))[0]
    assert hit["object"] == 0 and hit["distance"] == np.float32(4.0)
    assert np.allclose(hit["position"], [0, -1, 0]) and np.allclose(hit["normal"], [0, -1, 0])
    # tangent = normalise(cross((0,1,0), normal)); zero for this normal (vector3.rs:58-59)
    assert np.all(hit["tangent"] == 0.0)
    hit = orc.intersect(d, ray((-5, 0, 0), (1, 0, 0)))[0]
    assert np.allclose(hit["normal"], [-1, 0, 0]) and np.allclose(hit["tangent"], [0, 0, 1])
    assert orc.intersect(d, ray((0, 0, 0), (0, 1, 0)))[0]["object"] == -1       # inside
    assert orc.intersect(d, ray((0, 0.5, 0), (0, 0, 1)))[0]["object"] == -1     # inside, off-centre
    assert orc.intersect(d, ray((1, -5, 0), (0, 1, 0)))[0]["object"] == -1      # tangent
    assert orc.intersect(d, ray((0, 5, 0), (0, 1, 0)))[0]["object"] == -1       # behind


def test_plane_circle_halfspace(pkg, orc):
    # geometry.rs:55-71: t <= 0 and d == 0 miss; plane/circle normals face the ray (:80,:178)
    d = one_object_scene(pkg, lambda b: b.plane((0, 0, -1), (0, 0, 4))).desc()
    up = orc.intersect(d, ray((0, 0, 0), (0, 0, 1)))[0]
    assert up["object"] == 0 and up["distance"] == np.float32(4.0) and np.allclose(up["normal"], [0, 0, -1])
    down = orc.intersect(d, ray((0, 0, 8), (0, 0, -1)))[0]
    assert down["distance"] == np.float32(4.0) and np.allclose(down["normal"], [0, 0, 1])
    assert orc.intersect(d, ray((0, 0, 0), (1, 0, 0)))[0]["object"] == -1       # parallel, d == 0
    assert orc.intersect(d, ray((0, 0, 4), (0, 0, 1)))[0]["object"] == -1       # t == 0
    assert orc.intersect(d, ray((0, 0, 5), (0, 0, 1)))[0]["object"] == -1       # behind
    d = one_object_scene(pkg, lambda b: b.circle((0, 0, -1), (0, 0, 4), 2.0)).desc()
    assert orc.intersect(d, ray((2, 0, 0), (0, 0, 1)))[0]["object"] == 0        # on the rim: <= r^2 (:171)
    assert orc.intersect(d, ray((2.001, 0, 0), (0, 0, 1)))[0]["object"] == -1


def test_paraboloid(pkg, orc):
    # geometry.rs:286-358.  Paraboloid::new((0,0,1), 0, f): points with |p - focus| = distance
    # to the plane z = -f, i.e. z = (x^2 + y^2) / (4 f)
    f = 2.0
    d = one_object_scene(pkg, lambda b: b.paraboloid((0, 0, 1), (0, 0, 0), f)).desc()
    for x in (0.0, 1.0, 3.0):
        hit = orc.intersect(d, ray((x, 0, 10), (0, 0, -1)))[0]
        assert hit["object"] == 0
        assert abs(hit["position"][2] - x * x / (4 * f)) < 1e-5
    # a == 0 branch (:316-320): ray parallel to the axis; above: exact vertex hit
    hit = orc.intersect(d, ray((0, 0, 10), (0, 0, -1)))[0]
    assert abs(hit["distance"] - 10.0) < 1e-6 and np.allclose(hit["normal"], [0, 0, 1])
    # general branch picks the nearest positive root (:335-340)
    hit = orc.intersect(d, ray((-10, 0, 2), (1, 0, 0)))[0]
    assert abs(hit["position"][0] + 4.0) < 1e-4
    assert orc.intersect(d, ray((0, 0, -1), (1, 0, 0)))[0]["object"] == -1


def test_hexagonal_prism_is_eight_halfspaces(pkg, orc):
    # geometry.rs:409-416, :495-515; lies_inside is strict (:126)
    b = pkg.SceneBuilder()
    s = b.hexagonal_prism((0, 0, 1), (0, 0, 0), 3.0, 1.0, 0.0, 8.0)
    b.object(s, pkg.SceneBuilder.material(pkg.MATERIAL_SF10_GLASS))
    d = b.desc()
    assert d.n_surfaces == 15
    # along the axis from below: enters through the bottom cap z = 0, normal -z
    hit = orc.intersect(d, ray((0, 0, -2), (0, 0, 1)))[0]
    assert hit["object"] == 0 and abs(hit["distance"] - 2.0) < 1e-6 and np.allclose(hit["normal"], [0, 0, -1])
    # from inside: leaves through the top cap z = 8 with the outward normal (half-spaces are one-sided, :115)
    hit = orc.intersect(d, ray((0, 0, 4), (0, 0, 1)))[0]
    assert abs(hit["distance"] - 4.0) < 1e-6 and np.allclose(hit["normal"], [0, 0, 1])
    # sideways from inside: exits through a side face; inradius of the triangle = sqrt(3)/6 * 3
    hit = orc.intersect(d, ray((0, 0, 4), (1, 0, 0)))[0]
    assert abs(hit["distance"] - math.sqrt(3.0) / 6.0 * 3.0) < 1e-5
    # a ray passing beside the prism misses
    assert orc.intersect(d, ray((5, -10, 4), (0, 1, 0)))[0]["object"] == -1
    # a ray above the top cap misses even though it crosses the infinite prism
    assert orc.intersect(d, ray((-10, 0, 9), (1, 0, 0)))[0]["object"] == -1


def test_scene_intersect_first_object_wins_ties(pkg, orc):
    # scene.rs:51: strict <, so of two coincident surfaces the first in the list is kept
    b = pkg.SceneBuilder()
    b.object(b.plane((0, 0, -1), (0, 0, 4)), pkg.SceneBuilder.material(pkg.MATERIAL_DIFFUSE_GREY, 0.1))
    b.object(b.sphere((0, 0, 5), 1.0), pkg.SceneBuilder.material(pkg.MATERIAL_DIFFUSE_GREY, 0.2))
    b.object(b.plane((0, 0, -1), (0, 0, 4)), pkg.SceneBuilder.material(pkg.MATERIAL_DIFFUSE_GREY, 0.3))
    hit = orc.intersect(b.desc(), ray((0, 0, 0), (0, 0, 1)))[0]
    assert hit["object"] == 0 and hit["distance"] == np.float32(4.0)


# ----------------------------------------------------------------------- trace
@pytest.mark.parametrize("mode", [0, 1])
def test_c1_trace_statistics(pkg, orc, mode):
    # C1: 256x256, 1 spp (BASELINE.json configs[0]).  x, y, wavelength are pure RNG
    # (trace_unit.rs:154-158): uniform in [-1,1], [-1/aspect,1/aspect], [380,780]
    d = pkg.SceneBuilder(pkg.SCENE_C1).desc()
    ct = orc.Counters()
    ph = orc.trace(d, 0x5EED, 256, 256, 0, 65536, mode, False, ct)
    assert ct.photons == 65536 and ct.rays >= 65536
    assert -1.0 <= ph["x"].min() and ph["x"].max() <= 1.0 and abs(ph["x"].mean()) < 0.01
    assert 380.0 <= ph["wavelength"].min() and ph["wavelength"].max() <= 780.0
    assert abs(ph["wavelength"].mean() - 580.0) < 2.0
    # the emissive plane covers the upper half of the view: roughly half the photons see light
    lit = np.count_nonzero(ph["probability"])
    assert 0.3 < lit / 65536 < 0.7
    # direct hits of the emitter return intensity 1 * blackbody(lambda) (trace_unit.rs:99-101)
    m = d.objects[1].material
    direct = orc.blackbody_intensity(m.p0, m.p1, ph["wavelength"], mode)
    is_direct = ph["probability"] == direct
    assert np.count_nonzero(is_direct) > 0.25 * 65536
    # paths that bounced off the grey sphere carry reflectance^k < 1
    others = ph["probability"][(~is_direct) & (ph["probability"] > 0)]
    assert others.size > 0 and np.all(others < direct.max())


def test_spec_and_libm_modes_agree_statistically(pkg, orc):
    # Same estimator, different libm.  The RNG-only fields are bit-equal; the probabilities differ
    # where an ulp of libm moved a direction or flipped a branch (measured: 1.5 % of the photons,
    # 18 % of the lit ones -- paths are chaotic); the images agree within the stated fraction of
    # the Monte-Carlo noise (tests/stat_parity.py; the GPU-sized version is in
    # tests/test_gpu_parity.py::test_image_matches_reference_arithmetic).
    import stat_parity
    d = pkg.SceneBuilder(pkg.SCENE_C2).desc()
    w = h = 64
    k, n = 16, 12000
    subs, differ = [], 0
    whole_libm = np.zeros((h, w, 3), dtype=np.float32)
    for i in range(k):
        spec = orc.trace(d, 11, w, h, i * n, n, orc.MATH_SPEC)
        libm = orc.trace(d, 11, w, h, i * n, n, orc.MATH_LIBM)
        for f in ("x", "y", "wavelength"):
            assert np.array_equal(spec[f], libm[f])    # RNG only: bit-equal
        differ += int(np.count_nonzero(~np.isclose(spec["probability"], libm["probability"], rtol=1e-4, atol=1e-7)))
        subs.append(orc.plot(w, h, spec))
        orc.plot(w, h, libm, whole_libm)
    assert differ / (k * n) < 0.03, f"{differ / (k * n):.4f} of the photons differ"
    stat_parity.check(subs, whole_libm, "oracle SPEC vs LIBM")
    # and the statistic does tell two independent renders apart
    other = orc.plot(w, h, orc.trace(d, 12, w, h, 0, k * n, orc.MATH_SPEC))
    rms, _, _ = stat_parity.compare(subs, other)
    assert rms > 1.0


# ------------------------------------------------------------------------ plot
def test_plot_weights_and_borders(orc):
    # plot_unit.rs:56-84: four weights sum to 1; px/py clamp to the canvas
    w, h = 8, 4
    ph = np.zeros(3, dtype=orc.MAPPED_PHOTON)
    ph["x"], ph["y"] = [-1.0, 1.0, 0.123], [-0.5, 0.5, 0.0371]       # aspect = 2: y in [-0.5, 0.5]
    ph["probability"], ph["wavelength"] = 1.0, 555.0
    img = orc.plot(w, h, ph)
    cie = orc.tristimulus([555.0])[0]
    assert np.allclose(img.sum(axis=(0, 1)), 3 * cie, rtol=1e-5)
    assert np.allclose(img[0, 0], cie) and np.allclose(img[h - 1, w - 1], cie)
    # zero-probability photons leave the buffer unchanged (they add +0.0)
    ph["probability"] = 0.0
    assert np.array_equal(orc.plot(w, h, ph, img.copy()), img)


# ---------------------------------------------------------------------- gather
def test_kahan_step(orc):
    # gather_unit.rs:55-63
    acc = np.array([1.0, 1.0e8, 0.0], dtype=np.float32)
    comp = np.zeros(3, dtype=np.float32)
    px = np.array([1.0e-8, 1.0, 2.5], dtype=np.float32)
    want_acc, want_comp = acc.copy(), comp.copy()
    for _ in range(10):
        extra = px - want_comp
        s = want_acc + extra
        want_comp = (s - want_acc) - extra
        want_acc = s
        orc.gather_accumulate(acc, comp, px)
    assert np.array_equal(acc, want_acc) and np.array_equal(comp, want_comp)
    # compensated: 1e8 + 10 * 1 is recovered although 1e8 + 1 == 1e8 in f32
    assert float(acc[1]) - float(comp[1]) == 1.0e8 + 10.0


# --------------------------------------------------------------------- tonemap
@pytest.mark.parametrize("mode", [0, 1])
def test_tonemap_constant_image(orc, mode):
    # tonemap_unit.rs:55-85: constant 0.5 -> sigma = 0 -> white = mean -> ln 2 / ln 4 = 0.5 per
    # channel before the sRGB matrix (srgb.rs:29-41)
    xyz = np.full((16, 16, 3), 0.5, dtype=np.float32)
    assert orc.find_exposure(xyz) == 0.5
    rgb = orc.tonemap(xyz, mode)
    lin = np.array([3.2406 - 1.5372 - 0.4986, -0.9689 + 1.8758 + 0.0415, 0.0557 - 0.2040 + 1.0570]) * 0.5
    want = np.floor(np.clip(1.055 * lin ** (1 / 2.4) - 0.055, 0, 1) * 255.0)
    assert np.all(np.abs(rgb[0, 0].astype(int) - want.astype(int)) <= 1)
    assert np.all(rgb == rgb[0, 0])


def test_tonemap_nan_exposure_quirk(orc):
    # tonemap_unit.rs:65-68: sequential f32 sums can make the variance negative -> sqrt -> NaN
    # -> every channel NaN -> `as u8` saturates NaN to 0: a black frame
    xyz = np.full((1024, 1024, 3), 0.1, dtype=np.float32)
    e = orc.find_exposure(xyz)
    assert math.isnan(e)
    rgb = orc.tonemap(xyz, 0)           # exposure computed the reference's way -> NaN
    assert np.all(rgb == 0)


def test_spec_math_close_to_libm(orc):
    # the specified polynomial math tracks glibc to a few ulp on the path's argument ranges
    def ulps(a, b):
        return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
    x = np.linspace(-20.0, 20.0, 200001, dtype=np.float32)
    for fn in (0, 1):
        a, b = orc.math(fn, x, mode=0), orc.math(fn, x, mode=1)
        assert np.max(np.abs(a - b)) < 3e-7
    x = np.linspace(-30.0, 0.0, 100001, dtype=np.float32)
    assert ulps(orc.math(2, x, mode=0), orc.math(2, x, mode=1)).max() <= 2
    x = np.linspace(-0.999, 0.999, 100001, dtype=np.float32)
    assert ulps(orc.math(3, x, mode=0), orc.math(3, x, mode=1)).max() <= 4
    x = np.linspace(1.0, 3.0, 100001, dtype=np.float32)
    assert ulps(orc.math(6, x, mode=0), orc.math(6, x, mode=1))[1:].max() <= 4
    x = np.linspace(0.0032, 1.0, 100001, dtype=np.float32)
    y = np.full_like(x, 1.0 / 2.4)
    assert ulps(orc.math(7, x, y, mode=0), orc.math(7, x, y, mode=1)).max() <= 16
    wl = np.linspace(380.0, 780.0, 4001, dtype=np.float32)
    t = np.full_like(wl, 6504.0)
    assert ulps(orc.math(4, wl, t, mode=0), orc.math(4, wl, t, mode=1)).max() <= 1


def test_closed_draw_is_successor_of_scaled_integer():
    # The CUDA path computes Closed01<f32> = n / (2^24 - 1) as the float successor of n * 2^-24
    # (csrc/rl_device.cuh, Rng::unit); exhaustive check of that identity against IEEE division.
    n = np.arange(0, 1 << 24, dtype=np.uint32)
    quotient = n.astype(np.float32) / np.float32(16777215.0)
    succ = (n.astype(np.float32) * np.float32(2.0 ** -24)).view(np.uint32) + np.uint32(1)
    succ[0] = 0
    assert np.array_equal(quotient.view(np.uint32), succ)


def test_compound_of_spheres(pkg, orc):
    # Sphere is the reference's other Volume (geometry.rs:263-267): Compound<Sphere, Sphere> is a
    # lens, Compound<Sphere, SpacePartitioning> a dome.  geometry.rs:380-401: each child's own hit,
    # kept iff it lies inside the other child; Sphere::intersect misses from inside (:236-240), so
    # a ray that starts inside the lens never sees its far face.
    b = pkg.SceneBuilder()
    grey = pkg.SceneBuilder.material(pkg.MATERIAL_DIFFUSE_GREY, 0.5)
    b.object(b.compound(b.sphere((-1.2, 0, 2), 2.0), b.sphere((1.2, 0, 2), 2.0)), grey)           # lens, 0
    b.object(b.compound(b.sphere((5, 6, 0.5), 2.0), b.halfspace((0, 0, -1), (0, 0, 0.5))), grey)  # dome, 1
    rays = np.zeros(7, dtype=orc.RAY)
    rays["origin"] = [(-10, 0, 2), (10, 0, 2), (0, 0, 2), (0, -10, 2), (5, 6, 10), (5, 6, -10), (5, -10, 0.4)]
    rays["direction"] = [(1, 0, 0), (-1, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1), (0, 1, 0)]
    rays["wavelength"], rays["probability"] = 550.0, 1.0
    hit = orc.intersect(b.desc(), rays)
    # from -x: the first sphere's near face (x = -3.2) lies outside the second sphere; the second
    # sphere's near face (x = -0.8) lies inside the first: that is the lens's face
    assert hit["object"][0] == 0 and abs(hit["distance"][0] - 9.2) < 1e-5
    assert np.allclose(hit["normal"][0], (-1, 0, 0)) and np.allclose(hit["tangent"][0], (0, 0, 1))
    assert hit["object"][1] == 0 and abs(hit["distance"][1] - 9.2) < 1e-5 and np.allclose(hit["normal"][1], (1, 0, 0))
    assert hit["object"][2] == -1                       # from inside: both spheres miss
    # along y through the lens's rim region: |x| = 0 plane, the lens's half-height is sqrt(4 - 1.44) = 1.6
    assert hit["object"][3] == 0 and abs(hit["distance"][3] - (10 - 1.6)) < 1e-5
    # dome from above: the spherical cap at z = 2.5; from below: the flat face at z = 0.5
    assert hit["object"][4] == 1 and abs(hit["distance"][4] - 7.5) < 1e-5 and np.allclose(hit["normal"][4], (0, 0, 1))
    assert hit["object"][5] == 1 and abs(hit["distance"][5] - 10.5) < 1e-5 and np.allclose(hit["normal"][5], (0, 0, -1))
    assert np.allclose(hit["tangent"][5], 0.0)          # only spheres set a tangent (geometry.rs:250-251)
    assert hit["object"][6] == -1                       # below the cut: the sphere's faces there are outside the half-space


def keyframe_scene(pkg):
    """C1's sphere and emissive plane under a tabulated camera function: three frames that differ in
    position, orientation, field of view, depth of field and chromatic aberration."""
    b = pkg.SceneBuilder()
    b.object(b.sphere((0, 0, 0), 1.0), pkg.SceneBuilder.material(pkg.MATERIAL_DIFFUSE_GREY, 0.8))
    b.object(b.plane((0, 0, -1), (0, 0, 4)), pkg.SceneBuilder.blackbody(6504.0, 1.0))
    s = float(np.sin(0.2)), float(np.cos(0.2))
    b.keyframe_camera([((0, -5, 0), (0, 0, 0, 1), 0.35 * np.pi, 5.0, 1.0e9, 0.0),
                       ((1, -6, 0.5), (0, 0, s[0], s[1]), 0.30 * np.pi, 6.0, 3.0, 0.01),
                       ((-2, -4, 1), (s[0], 0, 0, s[1]), 0.40 * np.pi, 4.5, 2.0, 0.02)])
    return b


def test_keyframe_camera_picks_floor_of_t_times_n(pkg, orc):
    # scene.rs:34 `fn(f32) -> Camera`, tabulated: frame min(floor(t n), n - 1); t is the fourth draw
    b = keyframe_scene(pkg)
    n = 3000
    rays, xy = orc.camera_rays(b.desc(), 5, 64, 64, 0, n)
    t = xy["probability"]                               # the probe returns t in this field
    frame = np.minimum(np.floor(t * np.float32(3)).astype(int), 2)
    assert set(frame) == {0, 1, 2}
    pos = np.array([(0, -5, 0), (1, -6, 0.5), (-2, -4, 1)], dtype=np.float32)
    pinhole = frame == 0                                # depth_of_field 1e9: the lens is a point
    assert np.allclose(rays["origin"][pinhole], pos[0], atol=1e-6)
    for k in (1, 2):                                    # lens radius <= 1 / depth_of_field
        d = np.linalg.norm(rays["origin"][frame == k] - pos[k], axis=1)
        assert d.max() <= 1.0 / (3.0, 2.0)[k - 1] + 1e-5 and d.max() > 0.05
