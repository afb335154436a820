"""A second, independent restatement of the whole per-photon path -- TraceUnit::render,
render_camera_ray, Camera::get_ray / get_screen_ray, make_camera, render_ray, every material,
monte_carlo, rotate_towards, quaternion rotation (trace_unit.rs:81-168, camera.rs:47-108,
app.rs:327-357, material.rs:38-306, monte_carlo.rs:25-58, vector3.rs:69-93, quaternion.rs:34-110)
-- written in numpy f32 from the reference source, with its own control flow (a wavefront over all
photons, intersections by tests/test_oracle_numpy_cross_check.py), and compared bit for bit with
the C++ oracle's MappedPhoton records in SPEC math mode.

Shared with the oracle on purpose: Philox (pinned by the Random123 vectors), the elementary
functions sin/cos/tan/exp/acos and the black-body spectrum (called through orc.math /
orc.blackbody_intensity: those are *specified*), and the scene descriptor.
Everything the reference's source decides -- draw order, formulas, operation order, branch
conditions, the roulette, the origin nudge -- is restated here independently."""
import numpy as np
import pytest

from test_oracle_numpy_cross_check import F, cross, dot, normalise, scene_intersect, vec

PI = F(3.14159274)          # std::f32::consts::PI


class Draws:
    """monte_carlo.rs:25-43 on the specified stream: the draws in front of the first bounce are the
    words of Philox blocks 0 and 1 in order, the draws of bounce j the first words of block 2 + j."""

    def __init__(self, orc, seed, photon):
        self.orc, self.key = orc, (seed & 0xffffffff, seed >> 32)
        self.ctr = (photon & 0xffffffff, photon >> 32)
        self.block, self.words = 0, []

    def begin_bounce(self, j):
        self.block, self.words = 2 + j, []

    def _u24(self):
        if not self.words:
            self.words = list(self.orc.philox(self.key, (self.ctr[0], self.ctr[1], self.block, 0)))
            self.block += 1
        return self.words.pop(0) >> 8

    def unit(self):                                  # Closed01<f32>
        return F(self._u24()) / F(16777215.0)

    def half_open(self):                             # f32 in [0, 1)
        return F(self._u24()) * F(2.0 ** -24)

    def bi_unit(self):
        return self.unit() * F(2) - F(1)

    def longitude(self):
        return self.half_open() * PI * F(2)

    def wavelength(self):
        return self.unit() * F(400) + F(380)


class Math:
    def __init__(self, orc, mode):
        self.orc, self.mode = orc, mode

    def _f(self, fn, x):
        return F(self.orc.math(fn, np.array([x], dtype=F), mode=self.mode)[0])

    def sin(self, x): return self._f(0, x)
    def cos(self, x): return self._f(1, x)
    def exp(self, x): return self._f(2, x)
    def acos(self, x): return self._f(3, x)
    def tan(self, x): return self._f(8, x)


def qmul(a, b):                                      # quaternion.rs:100-110; (x, y, z, w)
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return (aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
            aw * bw - ax * bx - ay * by - az * bz)


def rotation(m, x, y, z, angle):                     # quaternion.rs:36-45
    s, c = m.sin(angle * F(0.5)), m.cos(angle * F(0.5))
    return (s * F(x), s * F(y), s * F(z), c)


def rotate(v, q):                                    # vector3.rs:85-89
    p = (v[0], v[1], v[2], F(0))
    conj = (-q[0], -q[1], -q[2], q[3])
    r = qmul(qmul(q, p), conj)
    return np.array([r[0], r[1], r[2]], dtype=F)


def camera_at(desc, m, t, pkg):                      # app.rs:327-357 / a static camera
    cm = desc.camera
    fixed = cm.fixed
    if cm.kind == pkg.CAMERA_STATIC:
        q = fixed.orientation
        return dict(position=vec(fixed.position), fov=F(fixed.field_of_view), focal=F(fixed.focal_distance),
                    dof=F(fixed.depth_of_field), ca=F(fixed.chromatic_abberation),
                    orientation=(F(q.x), F(q.y), F(q.z), F(q.w)))
    phi = PI * (F(cm.phi_base) + F(cm.phi_rate) * t)
    alpha = PI * (F(cm.alpha_base) + F(cm.alpha_rate) * t)
    distance = F(cm.distance_base) + F(cm.distance_rate) * t
    position = np.array([m.cos(alpha) * m.sin(phi) * distance, m.cos(alpha) * m.cos(phi) * distance,
                         m.sin(alpha) * distance], dtype=F)
    orientation = qmul(rotation(m, 0.0, 0.0, -1.0, phi + PI), rotation(m, 1.0, 0.0, 0.0, -alpha))
    return dict(position=position, fov=F(fixed.field_of_view), focal=distance * F(cm.focal_factor),
                dof=F(fixed.depth_of_field), ca=F(fixed.chromatic_abberation), orientation=orientation)


def camera_ray(cam, m, x, y, wavelength, rng):       # camera.rs:94-108, :47-90
    dof_angle = rng.longitude()
    dof_radius = rng.unit() / cam["dof"]
    d = (wavelength - F(580)) / F(200)
    zoom = F(1) + d * cam["ca"]
    screen_distance = F(1) / m.tan(cam["fov"] * F(0.5))
    xs, ys = x * zoom, y * zoom
    direction = normalise(np.array([xs, screen_distance, -ys], dtype=F))
    focus_point = direction * (cam["focal"] / direction[1])
    lens_point = np.array([m.cos(dof_angle) * dof_radius, F(0), m.sin(dof_angle) * dof_radius], dtype=F)
    origin = cam["position"] + rotate(lens_point, cam["orientation"])
    direction = normalise(rotate(focus_point - lens_point, cam["orientation"]))
    return origin, direction


def reflect(v, n):                                   # vector3.rs:91-93
    return v - n * F(2) * dot(n, v)


def rotate_towards(v, n):                            # vector3.rs:69-83
    if n[2] > F(0.9999):
        return v
    if n[2] < F(-0.9999):
        return np.array([v[0], v[1], -v[2]], dtype=F)
    up = np.array([0, 0, 1], dtype=F)
    a1 = normalise(cross(up, n))
    a2 = normalise(cross(a1, n))
    return a1 * v[0] + a2 * v[1] + n * v[2]


def diffuse_direction(m, rng, incoming, normal):     # material.rs:38-58, monte_carlo.rs:47-58
    phi = rng.longitude()
    rq = rng.unit()
    r = np.sqrt(rq)
    hemi = np.array([m.cos(phi) * r, m.sin(phi) * r, np.sqrt(F(1) - rq)], dtype=F)
    facing = normal if dot(incoming, normal) < 0 else -normal
    return rotate_towards(hemi, facing)


def sf10_ior(wavelength):                            # material.rs:203-213: f32 square, f64 Sellmeier
    w2 = np.float64(wavelength * wavelength * F(1.0e-6))
    n2 = (np.float64(1.0) + np.float64(1.737596950) * w2 / (w2 - np.float64(0.0131887070))
          + np.float64(0.313747346) * w2 / (w2 - np.float64(0.0623068142))
          + np.float64(1.898781010) * w2 / (w2 - np.float64(155.23629000)))
    return F(np.sqrt(n2))


def clamp999(x):
    return F(-0.999) if x < F(-0.999) else (F(0.999) if x > F(0.999) else x)


def bounce(pkg, m, rng, mat, direction, wavelength, normal, tangent):
    """Material::get_new_ray: (new direction, probability)."""
    kind = mat.kind
    if kind == pkg.MATERIAL_DIFFUSE_GREY:            # material.rs:122-130
        return diffuse_direction(m, rng, direction, normal), F(mat.p0)
    if kind == pkg.MATERIAL_DIFFUSE_COLOURED:        # material.rs:155-168
        p = (F(mat.p1) - wavelength) / F(mat.p2)
        q = m.exp(F(-0.5) * p * p)
        return diffuse_direction(m, rng, direction, normal), F(mat.p0) * q
    if kind == pkg.MATERIAL_GLOSSY_MIRROR:           # material.rs:185-196
        diffuse = diffuse_direction(m, rng, direction, normal)
        g = F(mat.p0)
        return normalise(diffuse * g + reflect(direction, normal) * (F(1) - g)), F(1)
    if kind == pkg.MATERIAL_SF10_GLASS:              # material.rs:216-261
        cos_i = -dot(direction, normal)
        ior = sf10_ior(wavelength)
        n = normal
        if cos_i > 0:
            ior = F(1) / ior
        else:
            n, cos_i = -normal, -cos_i
        sin_t_sqr = ior * ior * (F(1) - cos_i * cos_i)
        if sin_t_sqr > F(1):
            return reflect(direction, n), F(1)
        cos_t = np.sqrt(F(1) - sin_t_sqr)
        return direction * ior + n * (ior * cos_i - cos_t), F(1)
    if kind == pkg.MATERIAL_SOAP_BUBBLE:             # material.rs:267-306
        cos_alpha = dot(direction, normal)
        new = reflect(direction, normal) if rng.unit() - F(0.3) > abs(cos_alpha) else direction
        phase = (wavelength - F(380)) / F(200) * PI
        cos_phi, cos_theta = clamp999(dot(new, normal)), clamp999(dot(new, tangent))
        p = m.cos(phase - m.acos(cos_phi) * F(3) - m.acos(cos_theta) * F(2) + PI * F(0.5))
        return new, p * F(0.1) + F(0.9)
    raise AssertionError(f"material kind {kind}")


def trace(pkg, orc, desc, seed, width, height, first, n, mode):
    """TraceUnit::render for photon ids [first, first + n), all paths advanced together."""
    m = Math(orc, mode)
    aspect = F(width) / F(height)
    out = np.zeros(n, dtype=pkg.MAPPED_PHOTON)
    paths = []
    for k in range(n):                               # trace_unit.rs:151-158, :136-145
        rng = Draws(orc, seed, first + k)
        wavelength = rng.wavelength()
        x = rng.bi_unit()
        y = rng.bi_unit() / aspect
        out[k]["wavelength"], out[k]["x"], out[k]["y"] = wavelength, x, y
        t = rng.unit()
        origin, direction = camera_ray(camera_at(desc, m, t, pkg), m, x, y, wavelength, rng)
        paths.append(dict(k=k, rng=rng, o=origin, d=direction, wl=wavelength, intensity=F(1), chance=F(1), bounce=0))
    rays = 0
    while paths:                                     # trace_unit.rs:81-132
        o = np.stack([p["o"] for p in paths]).astype(F)
        d = np.stack([p["d"] for p in paths]).astype(F)
        winner, hit = scene_intersect(desc, o, d, pkg)
        rays += len(paths)
        alive = []
        for j, p in enumerate(paths):
            if winner[j] < 0:
                out[p["k"]]["probability"] = 0.0     # The Void
                continue
            mat = desc.objects[int(winner[j])].material
            if mat.kind == pkg.MATERIAL_BLACKBODY:   # material.rs:99-105
                light = F(orc.blackbody_intensity(mat.p0, mat.p1, np.array([p["wl"]], dtype=F), mode)[0])
                out[p["k"]]["probability"] = p["intensity"] * light
                continue
            p["rng"].begin_bounce(p["bounce"])
            p["bounce"] += 1
            new_d, probability = bounce(pkg, m, p["rng"], mat, p["d"], p["wl"], hit.normal[j], hit.tangent[j])
            p["intensity"] = p["intensity"] * probability
            p["d"] = new_d.astype(F)
            p["o"] = (hit.pos[j] + p["d"] * F(0.00001)).astype(F)
            p["chance"] = p["chance"] * F(0.96)
            if p["rng"].unit() * F(0.85) > p["chance"] * (F(1) - m.exp(p["intensity"] * F(-20))):
                out[p["k"]]["probability"] = 0.0     # roulette
                continue
            alive.append(p)
        paths = alive
    return out, rays


@pytest.mark.parametrize("which,param,w,h,first,n", [
    (1, 0, 256, 256, 0, 3000),                  # sphere + emissive plane, static camera
    (2, 0, 1024, 1024, 12345, 4000),            # built-in scene, orbiting camera, every material
    (2, 0, 1280, 720, (1 << 33) + 5, 1000),     # ids above 2^32, the reference's default canvas
    (3, 0, 640, 480, 777, 4000),                # SF10 prism: refraction, total internal reflection
    (4, 64, 512, 512, 99, 2500),                # random spheres: grey, coloured, glossy, soap
])
def test_oracle_trace_matches_numpy_restatement(pkg, orc, which, param, w, h, first, n):
    check(pkg, orc, which, param, w, h, first, n, orc.MATH_SPEC)


def test_oracle_trace_matches_numpy_restatement_libm_mode(pkg, orc):
    # the mode the CPU baseline runs in (glibc's functions, what the Rust binary calls)
    check(pkg, orc, 2, 0, 1024, 1024, 5_000_000, 1500, orc.MATH_LIBM)


def check(pkg, orc, which, param, w, h, first, n, mode):
    desc = pkg.SceneBuilder(which, param).desc()
    seed = 0x5EED
    ct = orc.Counters()
    want = orc.trace(desc, seed, w, h, first, n, mode, False, ct)
    with np.errstate(all="ignore"):
        got, rays = trace(pkg, orc, desc, seed, w, h, first, n, mode)
    assert rays == ct.rays
    for field in ("wavelength", "x", "y", "probability"):
        a, b = got[field].view(np.uint32), want[field].view(np.uint32)
        assert np.array_equal(a, b), f"{field}: {np.count_nonzero(a != b)} of {n} photons differ"
    assert np.count_nonzero(want["probability"]) >= 20      # enough lit paths to compare intensities on
