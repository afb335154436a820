"""The built-in scene (BASELINE configs[1]; App::set_up_scene, app.rs:166-325, with the geometry
constructors geometry.rs:195-201, :286-295, :420-515) restated in numpy f32 and compared, object by
object and field by field, with the descriptor host/rl_scene_builder.cpp emits -- the scene both the
oracle and the GPU render in every bench line.  sin and cos are the *specified* functions (the
builder uses them so that the scene is the same bytes on every platform; the Rust binary would
call the platform libm, whose values differ from these by an ulp here and there); everything else
is arithmetic."""
import numpy as np

from test_oracle_numpy_cross_check import F, dot, normalise, vec
from test_oracle_path_cross_check import PI, rotate_towards

GOLDEN_RATIO = 1.6180339887498948482045868343656381177203091798057628


class SpecTrig:
    def __init__(self, orc):
        self.orc = orc

    def sin(self, x):
        return F(self.orc.math(0, np.array([x], dtype=F), mode=self.orc.MATH_SPEC)[0])

    def cos(self, x):
        return F(self.orc.math(1, np.array([x], dtype=F), mode=self.orc.MATH_SPEC)[0])


def v(x, y, z):
    return np.array([x, y, z], dtype=F)


# ---- geometry constructors: each returns a tree ("kind", params..., children)
def sphere(position, radius):                       # geometry.rs:195-201
    return ("sphere", position, radius * radius)


def circle(normal, position, radius):               # geometry.rs:142-150
    return ("circle", normal, position, radius * radius)


def paraboloid(normal, offset, focal_distance):     # geometry.rs:286-295
    return ("paraboloid", offset - normal * focal_distance, normal, normal * (focal_distance * F(2)))


def halfspace(normal, offset):
    return ("halfspace", normal, offset)


def infinite_prism(m, axis, offset, edge_length, angle):     # geometry.rs:420-451
    radius = np.sqrt(F(3)) / F(6) * edge_length
    a1, a2, a3 = angle, angle + PI * F(2) / F(3), angle + PI * F(4) / F(3)
    planes = []
    for a in (a1, a2, a3):
        p = rotate_towards(v(m.cos(a), m.sin(a), F(0)), axis)
        planes.append(halfspace(p, p * radius + offset))
    return ("compound", ("compound", planes[0], planes[1]), planes[2])


def thick_plane(normal, offset, thickness):         # geometry.rs:456-470
    return ("compound", halfspace(-normal, offset), halfspace(normal, offset + normal * thickness))


def prism(m, axis, offset, edge_length, angle, height):      # geometry.rs:476-487
    return ("compound", infinite_prism(m, axis, offset, edge_length, angle), thick_plane(axis, offset, height))


def hexagonal_prism(m, axis, offset, edge_length, bevel_size, angle, height):   # geometry.rs:495-515
    bevel = infinite_prism(m, axis, offset, edge_length * F(2) - bevel_size * F(3), angle + PI)
    return ("compound", bevel, prism(m, axis, offset, edge_length, angle, height))


def paraboloid_hit(shape, origin, direction):       # geometry.rs:299-358 for one ray: (position, normal) or None
    _, offset, normal, focal = shape
    o = origin - offset
    fo = o - focal
    ndd, ndo, ddf = dot(normal, direction), dot(normal, o), dot(direction, fo)
    a = ndd * ndd - F(1)
    b = F(2) * ndd * ndo - F(2) * ddf
    c = ndo * ndo - dot(fo, fo)
    if a == 0:
        t = -c / b
        if t < 0:
            return None
    else:
        d = b * b - F(4) * a * c
        if d < 0:
            return None
        root = np.sqrt(d)
        p, q = F(0.5) * (-b + root) / a, F(0.5) * (-b - root) / a
        if p > 0 and (p < q or q < 0):
            t = p
        elif q > 0:
            t = q
        else:
            return None
    pos = origin + direction * t
    local = pos - offset
    plane_pr = local - normal * dot(local, normal)
    return pos, normalise(focal - plane_pr)


def built_in_scene(m):
    """[(shape tree, (material kind name, p0, p1, p2))] in list order (app.rs:166-325)."""
    objects = []
    sun_radius = F(5)
    sun_position = v(0, 0, 0)
    objects.append((sphere(sun_position, sun_radius), ("blackbody", F(6504), F(1.0), None)))
    floor_normal = v(0, 0, -1)
    floor = paraboloid(floor_normal, v(0, 0, -sun_radius), sun_radius * sun_radius)
    objects.append((floor, ("grey", F(0.8), None, None)))
    objects.append((paraboloid(v(0, 0, 1), v(1, 0, -(sun_radius * sun_radius)), sun_radius * sun_radius),
                    ("coloured", F(0.9), F(550), F(40))))
    objects.append((paraboloid(v(0, 0, 1), v(-1, 0, -(sun_radius * sun_radius)), sun_radius * sun_radius),
                    ("coloured", F(0.9), F(660), F(60))))
    sky_height = F(30)
    objects.append((circle(floor_normal, v(-sun_radius, 0, sky_height), F(5)), ("blackbody", F(7600), F(0.6), None)))
    sky2_radius = F(15)
    objects.append((circle(floor_normal, v(-sun_radius * F(0.5), sun_radius * F(2) + sky2_radius, sky_height), sky2_radius),
                    ("blackbody", F(5000), F(0.6), None)))
    objects.append((("plane", floor_normal, v(0, 0, sky_height * F(2))), ("coloured", F(0.5), F(470), F(25))))

    gamma = PI * F(2) * (F(1) - F(1) / F(GOLDEN_RATIO))
    seed_size, seed_scale = F(0.8), F(1.5)
    first_seed = int((sun_radius / seed_scale + F(1)) * (sun_radius / seed_scale + F(1)) + F(0.5))
    seeds = 100
    for i in range(first_seed, first_seed + seeds):              # spiral sunflower seeds
        phi = F(i) * gamma
        r = np.sqrt(F(i)) * seed_scale
        position = v(m.cos(phi) * r, m.sin(phi) * r, (r - sun_radius) * F(-0.5)) + sun_position
        wavelength = F(i - first_seed) / F(seeds) * F(130) + F(600)
        objects.append((sphere(position, seed_size), ("coloured", F(0.9), wavelength, F(60))))
    for i in range(first_seed, first_seed + seeds):              # seeds in between
        phi = (F(i) + F(0.5)) * gamma
        r = np.sqrt(F(i) + F(0.5)) * seed_scale
        position = v(m.cos(phi) * r, m.sin(phi) * r, (r - sun_radius) * F(-0.25)) + sun_position
        objects.append((sphere(position, seed_size * F(0.5)), ("glossy", F(0.1), None, None)))
    for i in range(first_seed // 2, first_seed + seeds):         # soap bubbles above
        phi = F(-i) * gamma
        r = np.sqrt(F(i)) * seed_scale * F(1.5)
        position = v(m.cos(phi) * r, m.sin(phi) * r, (r - sun_radius) * F(1.5) + sun_radius * F(2)) + sun_position
        objects.append((sphere(position, seed_size * (F(0.5) + np.sqrt(F(i)) * F(0.2))), ("soap", None, None, None)))

    prisms = 11                                                  # prisms along the walls
    prism_angle = PI * F(2) / F(prisms)
    prism_radius, prism_height = F(17), F(8)
    for i in range(prisms):
        for ofs, radius, phi_ofs, h in ((F(0), F(1), F(0), F(1)), (F(0.5) * prism_angle, F(1.2), PI * F(0.5), F(1.5))):
            phi = F(i) * prism_angle + ofs
            position = v(m.cos(phi) * prism_radius * radius, m.sin(phi) * prism_radius * radius, 0)
            normal = v(0, 0, -1)
            hit = paraboloid_hit(floor, position, normal)
            if hit is not None:
                normal = -hit[1]
                position = hit[0] + normal * F(2) * h
            objects.append((hexagonal_prism(m, normal, position, F(3), F(1), phi + phi_ofs, prism_height * h),
                            ("glass", None, None, None)))
    return objects


def tree_of(desc, idx, pkg):
    s = desc.surfaces[idx]
    if s.kind == pkg.SURFACE_PLANE:
        return ("plane", vec(s.a), vec(s.b))
    if s.kind == pkg.SURFACE_HALFSPACE:
        return ("halfspace", vec(s.a), vec(s.b))
    if s.kind == pkg.SURFACE_CIRCLE:
        return ("circle", vec(s.a), vec(s.b), F(s.s))
    if s.kind == pkg.SURFACE_SPHERE:
        return ("sphere", vec(s.a), F(s.s))
    if s.kind == pkg.SURFACE_PARABOLOID:
        return ("paraboloid", vec(s.a), vec(s.b), vec(s.c))
    return ("compound", tree_of(desc, int(s.child[0]), pkg), tree_of(desc, int(s.child[1]), pkg))


def same(a, b, where):
    assert a[0] == b[0], f"{where}: {a[0]} vs {b[0]}"
    if a[0] == "compound":
        same(a[1], b[1], where + ".1")
        same(a[2], b[2], where + ".2")
        return
    for k, (x, y) in enumerate(zip(a[1:], b[1:])):
        x, y = np.asarray(x, dtype=F), np.asarray(y, dtype=F)
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), f"{where} {a[0]} field {k}: {x} vs {y}"


def test_built_in_scene_descriptor_matches_numpy_restatement(pkg, orc):
    desc = pkg.SceneBuilder(pkg.SCENE_C2).desc()
    mine = built_in_scene(SpecTrig(orc))
    assert desc.n_objects == len(mine) == 339
    names = {pkg.MATERIAL_BLACKBODY: "blackbody", pkg.MATERIAL_DIFFUSE_GREY: "grey", pkg.MATERIAL_DIFFUSE_COLOURED: "coloured",
             pkg.MATERIAL_GLOSSY_MIRROR: "glossy", pkg.MATERIAL_SF10_GLASS: "glass", pkg.MATERIAL_SOAP_BUBBLE: "soap"}
    for k, (shape, material) in enumerate(mine):
        obj = desc.objects[k]
        same(tree_of(desc, int(obj.surface), pkg), shape, f"object {k}")
        assert names[obj.material.kind] == material[0], f"object {k}"
        if material[0] == "blackbody":
            # BlackBodyMaterial::new (material.rs:92-97): intensity / boltzmann(Wien peak) in f64, rounded once
            h, kb, c, wien = 6.62606957e-34, 1.3806488e-23, 299792458.0, 2.897772126e-3
            t = float(material[1])
            f = c / ((wien / t) * 1.0e9 * 1.0e-9)
            peak = (2.0 * h * f * f * f) / (c * c * (np.exp(h * f / (kb * t)) - 1.0))
            assert F(obj.material.p0) == material[1]
            assert abs(float(obj.material.p1) / (float(material[2]) / float(F(peak))) - 1.0) < 1e-6
        else:
            for got, want in zip((obj.material.p0, obj.material.p1, obj.material.p2), material[1:]):
                if want is not None:
                    assert F(got) == want, f"object {k} material parameter"
