"""A second, independent restatement of Scene::intersect -- vectorised numpy in IEEE f32, written
from the reference source (scene.rs:39-60, geometry.rs:53-407, vector3.rs:27-67) without looking
at oracle/oracle.cpp's control flow -- checked bit for bit against the C++ oracle on random rays.

The oracle is what the GPU is bit-equal to; the reference ships no golden vectors for this path
and cannot be built here, so this narrows the "did the oracle misread the source" risk for the
part of the path that is pure arithmetic (+ - * / sqrt are correctly rounded in both)."""
import numpy as np
import pytest

F = np.float32


def v3(x, y, z):
    return np.stack([x, y, z], axis=-1)


def dot(a, b):                       # vector3.rs:35-37: left to right, no fusing
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]


def cross(a, b):                     # vector3.rs:27-33
    return v3(a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
              a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
              a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0])


def normalise(a):                    # vector3.rs:56-67: three divisions, zero vector unchanged
    m = np.sqrt(dot(a, a))
    safe = np.where(m == 0, F(1), m)
    return np.where((m == 0)[..., None], a, a / safe[..., None])


class Hit:
    """Option<Intersection> for N rays: `ok` says Some."""

    def __init__(self, ok, t, pos, normal, tangent):
        self.ok, self.t, self.pos, self.normal, self.tangent = ok, t, pos, normal, tangent

    @staticmethod
    def select(cond, a, b):
        c3 = cond[..., None]
        return Hit(np.where(cond, a.ok, b.ok), np.where(cond, a.t, b.t), np.where(c3, a.pos, b.pos),
                   np.where(c3, a.normal, b.normal), np.where(c3, a.tangent, b.tangent))


def vec(c):
    return np.array([c.x, c.y, c.z], dtype=F)


def intersect_plane(normal, offset, o, d):          # geometry.rs:55-71
    origin = o - offset
    dn = dot(normal[None, :], d)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = -dot(normal[None, :], origin) / dn
        ok = (dn != 0) & ~(t <= 0)                    # `if t <= 0.0 { None }`: a NaN distance passes, as there
    pos = o + d * t[..., None]
    return ok, t, pos, dn


def surface(desc, idx, o, d, pkg):
    """Surface::intersect of node `idx` for all rays."""
    s = desc.surfaces[idx]
    n = o.shape[0]
    zero3 = np.zeros((n, 3), dtype=F)
    if s.kind in (pkg.SURFACE_PLANE, pkg.SURFACE_CIRCLE):         # geometry.rs:73-87, :166-184
        normal, offset = vec(s.a), vec(s.b)
        ok, t, pos, dn = intersect_plane(normal, offset, o, d)
        if s.kind == pkg.SURFACE_CIRCLE:
            rel = pos - offset
            ok = ok & (dot(rel, rel) <= F(s.s))
        two_sided = np.where((dn < 0)[..., None], normal[None, :], -normal[None, :])
        return Hit(ok, t, pos, two_sided, zero3)
    if s.kind == pkg.SURFACE_HALFSPACE:                            # geometry.rs:109-121
        normal, offset = vec(s.a), vec(s.b)
        ok, t, pos, _ = intersect_plane(normal, offset, o, d)
        return Hit(ok, t, pos, np.broadcast_to(normal, (n, 3)).copy(), zero3)
    if s.kind == pkg.SURFACE_SPHERE:                               # geometry.rs:204-261
        centre, r2 = vec(s.a), F(s.s)
        co = centre[None, :] - o
        b = F(2) * dot(d, co)
        c = dot(co, co) - r2
        disc = b * b - F(4) * F(1) * c
        with np.errstate(invalid="ignore"):
            root = np.sqrt(disc)
            t1 = F(-0.5) * (-b + root) / F(1)
            t2 = F(-0.5) * (-b - root) / F(1)
            take1 = (t1 > 0) & (t1 < t2)
            take2 = ~take1 & (t2 > 0) & (t2 < t1)
        ok = (disc >= 0) & (take1 | take2)
        t = np.where(take1, t1, t2)
        pos = o + d * t[..., None]
        normal = normalise(pos - centre[None, :])
        up = np.broadcast_to(np.array([0, 1, 0], dtype=F), (n, 3))
        tangent = normalise(cross(up, normal))
        return Hit(ok, t, pos, normal, tangent)
    if s.kind == pkg.SURFACE_PARABOLOID:                           # geometry.rs:299-358
        offset, normal, focal = vec(s.a), vec(s.b), vec(s.c)
        origin = o - offset
        fo = origin - focal
        ndd = dot(normal[None, :], d)
        ndo = dot(normal[None, :], origin)
        ddf = dot(d, fo)
        a = ndd * ndd - F(1)
        b = F(2) * ndd * ndo - F(2) * ddf
        c = ndo * ndo - dot(fo, fo)
        with np.errstate(divide="ignore", invalid="ignore"):
            lin = -c / b
            disc = b * b - F(4) * a * c
            root = np.sqrt(disc)
            p = F(0.5) * (-b + root) / a
            q = F(0.5) * (-b - root) / a
            pick_p = (p > 0) & ((p < q) | (q < 0))
            pick_q = ~pick_p & (q > 0)
            quad_ok = (disc >= 0) & (pick_p | pick_q)
            t = np.where(a == 0, lin, np.where(pick_p, p, q))
            ok = np.where(a == 0, ~(lin < 0), quad_ok)
        pos = o + d * t[..., None]
        local = pos - offset
        plane_pr = local - normal[None, :] * dot(local, normal[None, :])[..., None]
        nrm = normalise(focal[None, :] - plane_pr)
        return Hit(ok, t, pos, nrm, zero3)
    if s.kind == pkg.SURFACE_COMPOUND:                             # geometry.rs:380-401
        c1, c2 = int(s.child[0]), int(s.child[1])
        i1, i2 = surface(desc, c1, o, d, pkg), surface(desc, c2, o, d, pkg)
        ok1 = i1.ok & inside(desc, c2, i1.pos, pkg)
        ok2 = i2.ok & inside(desc, c1, i2.pos, pkg)
        with np.errstate(invalid="ignore"):
            first = ok1 & (~ok2 | (i1.t < i2.t))                   # both valid: strictly nearer, else the second
        out = Hit.select(first, i1, i2)
        out.ok = ok1 | ok2
        return out
    raise AssertionError(f"surface kind {s.kind}")


def inside(desc, idx, p, pkg):
    """Volume::lies_inside (geometry.rs:123-128, :404-407)."""
    s = desc.surfaces[idx]
    if s.kind == pkg.SURFACE_HALFSPACE:
        return dot(p - vec(s.b)[None, :], vec(s.a)[None, :]) < 0
    if s.kind == pkg.SURFACE_COMPOUND:
        return inside(desc, int(s.child[0]), p, pkg) & inside(desc, int(s.child[1]), p, pkg)
    raise AssertionError(f"volume kind {s.kind}")


def scene_intersect(desc, o, d, pkg):                               # scene.rs:39-60
    n = o.shape[0]
    best = Hit(np.zeros(n, bool), np.full(n, 1.0e12, dtype=F), np.zeros((n, 3), F), np.zeros((n, 3), F),
               np.zeros((n, 3), F))
    winner = np.full(n, -1, dtype=np.int32)
    for k in range(desc.n_objects):
        h = surface(desc, int(desc.objects[k].surface), o, d, pkg)
        with np.errstate(invalid="ignore"):
            nearer = h.ok & (h.t < best.t)
        best = Hit.select(nearer, h, best)
        winner = np.where(nearer, np.int32(k), winner)
    return winner, best


def random_rays(rng, n, extent, towards_origin):
    o = rng.uniform(-extent, extent, (n, 3)).astype(F)
    d = rng.normal(size=(n, 3)).astype(F)
    if towards_origin:
        d = (-o + rng.normal(scale=extent * 0.3, size=(n, 3))).astype(F)
    d = normalise(d)
    return o, d


def compare(pkg, orc, desc, o, d, what, min_hit_share=0.1):
    n = o.shape[0]
    rays = np.zeros(n, dtype=pkg.RAY)
    rays["origin"], rays["direction"], rays["wavelength"] = o, d, 550.0
    got = orc.intersect(desc, rays)
    winner, best = scene_intersect(desc, o, d, pkg)
    assert np.count_nonzero(winner >= 0) > n * min_hit_share, what  # the sample does hit things
    assert np.array_equal(got["object"], winner), what
    hit = winner >= 0
    for name, mine in (("distance", best.t), ("position", best.pos), ("normal", best.normal), ("tangent", best.tangent)):
        a = np.ascontiguousarray(got[name][hit]).view(np.uint32)
        b = np.ascontiguousarray(mine[hit].astype(F)).view(np.uint32)
        assert np.array_equal(a, b), f"{what} {name}: {np.count_nonzero(a != b)} words differ"
    return winner, best


@pytest.mark.parametrize("which,param,extent,n", [(1, 0, 8.0, 20000), (2, 0, 60.0, 12000), (3, 0, 15.0, 20000),
                                                  (4, 96, 30.0, 12000)])
def test_oracle_intersect_matches_numpy_restatement(pkg, orc, which, param, extent, n):
    desc = pkg.SceneBuilder(which, param).desc()
    rng = np.random.default_rng(1000 + which)
    o1, d1 = random_rays(rng, n // 2, extent, False)
    o2, d2 = random_rays(rng, n - n // 2, extent, True)
    o, d = np.concatenate([o1, o2]), np.concatenate([d1, d2])
    winner, best = compare(pkg, orc, desc, o, d, "primary rays")
    # second and third generation: rays that leave a surface the way render_ray continues a path
    # (trace_unit.rs:104-114: new direction, origin nudged 1e-5 along it) -- grazing starts, origins
    # on prism faces and inside spheres' epsilon shells
    for generation in (2, 3):
        hit = winner >= 0
        pos = best.pos[hit]
        new_d = normalise(rng.normal(size=pos.shape).astype(F))
        flip = dot(new_d, best.normal[hit]) < 0
        new_d = np.where(flip[..., None] & (rng.random(pos.shape[0]) < 0.7)[..., None], -new_d, new_d)
        new_o = pos + new_d * F(0.00001)
        winner, best = compare(pkg, orc, desc, new_o, new_d, f"generation {generation}", 0.005)


def test_oracle_plot_matches_scalar_restatement(orc):
    # PlotUnit::plot + plot_pixel + get_tristimulus (plot_unit.rs:56-95, cie1931.rs:20-48) restated
    # photon by photon in numpy f32 scalars; the table itself is read back from the oracle at the
    # grid wavelengths (its values are pinned by test_tristimulus_table_and_lerp)
    grid = (380.0 + 5.0 * np.arange(81)).astype(F)
    table = orc.tristimulus(grid).astype(F)                        # remainder 0: X[i] * 1 + X[i+1] * 0 = X[i]
    w, h = 37, 23
    aspect = F(w) / F(h)
    rng = np.random.default_rng(77)
    n = 6000
    photons = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("probability", "<f4"), ("wavelength", "<f4")])
    photons["x"] = rng.uniform(-1.02, 1.02, n)                     # a little outside: the clamps
    photons["y"] = rng.uniform(-1.02, 1.02, n) / aspect
    photons["probability"] = np.where(rng.random(n) < 0.3, 0.0, rng.uniform(0, 3, n))
    photons["wavelength"] = rng.uniform(380.0, 780.0, n)
    photons["wavelength"][:4] = [380.0, 780.0, 779.99994, 382.5]
    photons["x"][:3], photons["y"][:3] = [-1.0, 1.0, 0.0], [F(-1) / aspect, F(1) / aspect, 0.0]
    buf = np.zeros((h, w, 3), dtype=F)
    one, half = F(1), F(0.5)
    for ph in photons:
        indexf = (F(ph["wavelength"]) - F(380)) / F(5)
        index = int(np.floor(indexf))
        rem = indexf - F(index)
        if index < -1 or index > 80:
            cie = np.zeros(3, dtype=F)
        elif index == -1:
            cie = table[0] * rem
        elif index == 80:
            cie = table[80] * (one - rem)
        else:
            cie = table[index] * (one - rem) + table[index + 1] * rem
        cie = cie * F(ph["probability"])
        px = (F(ph["x"]) * half + half) * (F(w) - one)
        py = (F(ph["y"]) * aspect * half + half) * (F(h) - one)
        px1 = max(0, min(w - 1, int(np.floor(px)))); px2 = max(0, min(w - 1, int(np.ceil(px))))
        py1 = max(0, min(h - 1, int(np.floor(py)))); py2 = max(0, min(h - 1, int(np.ceil(py))))
        cx, cy = px - F(px1), py - F(py1)
        c11, c12, c21, c22 = (one - cx) * (one - cy), (one - cx) * cy, cx * (one - cy), cx * cy
        buf[py1, px1] = buf[py1, px1] + cie * c11
        buf[py1, px2] = buf[py1, px2] + cie * c21
        buf[py2, px1] = buf[py2, px1] + cie * c12
        buf[py2, px2] = buf[py2, px2] + cie * c22
    got = orc.plot(w, h, photons)
    assert np.array_equal(got.view(np.uint32), buf.view(np.uint32))
    assert buf.any()


def test_oracle_gather_and_tonemap_match_numpy_restatement(orc):
    # GatherUnit::accumulate (gather_unit.rs:49-64), TonemapUnit::find_exposure / tonemap
    # (tonemap_unit.rs:55-100), srgb::transform / gamma_correct (srgb.rs:20-41) restated in numpy f32;
    # ln and powf go through the oracle's specified functions (orc.math), the rest is arithmetic
    rng = np.random.default_rng(5)
    w, h = 48, 31
    acc = np.zeros((h, w, 3), dtype=F)
    comp = np.zeros((h, w, 3), dtype=F)
    mine_acc, mine_comp = acc.copy(), comp.copy()
    for _ in range(7):
        px = (rng.uniform(0, 1, (h, w, 3)) ** 6 * 40).astype(F)     # a few bright pixels, many dim ones
        orc.gather_accumulate(acc, comp, px)
        extra = px - mine_comp
        total = mine_acc + extra
        mine_comp = (total - mine_acc) - extra
        mine_acc = total
    assert np.array_equal(acc.view(np.uint32), mine_acc.view(np.uint32))
    assert np.array_equal(comp.view(np.uint32), mine_comp.view(np.uint32))

    # find_exposure: two sequential f32 folds over the pixels in order
    y = acc[..., 1].reshape(-1)
    n = F(w * h)
    mean = np.cumsum(y, dtype=F)[-1] / n
    sqr_mean = np.cumsum(y * y, dtype=F)[-1] / n
    white = mean + np.sqrt(sqr_mean - mean * mean)
    assert F(orc.find_exposure(acc)) == white

    def fn(code, x, x2=None):
        return orc.math(code, np.ascontiguousarray(x, dtype=F).reshape(-1), None if x2 is None else
                        np.ascontiguousarray(x2, dtype=F).reshape(-1), mode=orc.MATH_SPEC).reshape(np.shape(x))

    ln4 = fn(6, np.array([4.0], dtype=F))[0]
    cie = fn(6, acc / white + F(1)) / ln4
    x, yy, z = cie[..., 0], cie[..., 1], cie[..., 2]
    lin = np.stack([F(3.2406) * x - F(1.5372) * yy - F(0.4986) * z,
                    F(-0.9689) * x + F(1.8758) * yy + F(0.0415) * z,
                    F(0.0557) * x - F(0.2040) * yy + F(1.0570) * z], axis=-1)
    with np.errstate(invalid="ignore"):
        powed = fn(7, lin, np.full(lin.shape, F(1) / F(2.4), dtype=F))
        gamma = np.where(lin <= F(0.0031308), F(12.92) * lin, F(1.055) * powed - F(0.055))
    clamped = np.where(gamma < 0, F(0), np.where(gamma > 1, F(1), gamma))
    rgb = np.floor(clamped * F(255)).astype(np.uint8)                # `as u8` truncates; all values are in [0, 255]
    got = orc.tonemap(acc, orc.MATH_SPEC)
    assert np.array_equal(got, rgb)
    assert len(np.unique(rgb)) > 100                                 # a real image, not a flat one


def test_oracle_blackbody_and_sellmeier_match_numpy_restatement(orc):
    # boltzmann / BlackBodyMaterial (material.rs:61-74, :92-105, constants.rs:19-25) in numpy f64; numpy's
    # exp is not glibc's bit for bit, so the f32 result may differ in the last place -- and the SF10
    # Sellmeier index (material.rs:203-213), which is arithmetic only: bit-equal
    h, k, c, wien = 6.62606957e-34, 1.3806488e-23, 299792458.0, 2.897772126e-3

    def boltzmann(wavelength_nm, temperature):
        f = c / (wavelength_nm * 1.0e-9)
        return (2.0 * h * f * f * f) / (c * c * (np.exp(h * f / (k * temperature)) - 1.0))

    wl = np.linspace(380.0, 780.0, 2001).astype(F)
    for temperature, intensity in ((6504.0, 1.0), (7600.0, 0.6), (5000.0, 0.6), (2700.0, 3.0)):
        t32 = F(temperature)
        norm = F(intensity) / F(boltzmann((wien / float(t32)) * 1.0e9, float(t32)))
        want = (boltzmann(wl.astype(np.float64), float(t32)).astype(F) * norm).astype(F)
        got = orc.blackbody_intensity(float(t32), float(norm), wl, orc.MATH_LIBM)
        ulps = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
        assert ulps.max() <= 1, f"T={temperature}: {ulps.max()} ulp"
        spec = orc.blackbody_intensity(float(t32), float(norm), wl, orc.MATH_SPEC)
        assert np.abs(spec.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64)).max() <= 2

    w2 = (wl * wl * F(1.0e-6)).astype(np.float64)
    n2 = (1.0 + 1.737596950 * w2 / (w2 - 0.0131887070) + 0.313747346 * w2 / (w2 - 0.0623068142)
          + 1.898781010 * w2 / (w2 - 155.23629000))
    ior = np.sqrt(n2).astype(F)
    got = orc.math(5, wl)
    assert np.array_equal(got.view(np.uint32), ior.view(np.uint32))
