// rl_scene_builder.cpp -- host-side mirrors of the reference's geometry and
// material constructors and of App::set_up_scene, emitting the POD scene
// descriptor of include/rl_b200.h.  Input generation only: nothing here runs
// on the path.  Trigonometry uses the specified sincos (csrc/rl_math.cuh) so
// that descriptors are identical on every host.
#include <cmath>
#include <cstdint>
#include <new>
#include <vector>

#include "../../include/rl_host.h"
#include "../csrc/rl_math.cuh"

using rl::V3;
using rl::mk;

struct rl_scene_builder {
    std::vector<rl_surface> surfaces;
    std::vector<rl_object> objects;
    std::vector<rl_camera> keyframes;
    rl_camera_model camera;
};

namespace {

const float PI = RL_PI;

rl_vec3 rv(V3 v) { return rl_vec3{v.x, v.y, v.z}; }
V3 vr(rl_vec3 v) { return mk(v.x, v.y, v.z); }

int push(rl_scene_builder *b, const rl_surface &s) {
    b->surfaces.push_back(s);
    return (int)b->surfaces.size() - 1;
}

rl_surface blank(uint32_t kind) {
    rl_surface s;
    s.kind = kind;
    s.a = s.b = s.c = rl_vec3{0.f, 0.f, 0.f};
    s.s = 0.f;
    s.child[0] = s.child[1] = 0;
    return s;
}

int halfspace(rl_scene_builder *b, V3 normal, V3 offset) {  // geometry.rs:99-106
    rl_surface s = blank(RL_SURFACE_HALFSPACE);
    s.a = rv(normal); s.b = rv(offset);
    return push(b, s);
}

int compound(rl_scene_builder *b, int s1, int s2) {          // geometry.rs:369-378
    rl_surface s = blank(RL_SURFACE_COMPOUND);
    s.child[0] = (uint32_t)s1; s.child[1] = (uint32_t)s2;
    return push(b, s);
}

int infinite_prism(rl_scene_builder *b, V3 axis, V3 offset, float edge_length, float angle) {
    // geometry.rs:421-450
    const float radius = std::sqrt(3.0f) / 6.0f * edge_length;
    const float a1 = angle;
    const float a2 = angle + PI * 2.0f / 3.0f;
    const float a3 = angle + PI * 4.0f / 3.0f;
    float s, c;
    rl::spec_sincos(a1, s, c); V3 p1 = mk(c, s, 0.0f);
    rl::spec_sincos(a2, s, c); V3 p2 = mk(c, s, 0.0f);
    rl::spec_sincos(a3, s, c); V3 p3 = mk(c, s, 0.0f);
    p1 = rl::rotate_towards(p1, axis);
    p2 = rl::rotate_towards(p2, axis);
    p3 = rl::rotate_towards(p3, axis);
    const int sp1 = halfspace(b, p1, p1 * radius + offset);
    const int sp2 = halfspace(b, p2, p2 * radius + offset);
    const int sp3 = halfspace(b, p3, p3 * radius + offset);
    return compound(b, compound(b, sp1, sp2), sp3);
}

int thick_plane(rl_scene_builder *b, V3 normal, V3 offset, float thickness) {  // geometry.rs:455-469
    const int sp1 = halfspace(b, -normal, offset);
    const int sp2 = halfspace(b, normal, offset + normal * thickness);
    return compound(b, sp1, sp2);
}

int prism(rl_scene_builder *b, V3 axis, V3 offset, float edge_length, float angle, float height) {
    // geometry.rs:474-487
    const int p = infinite_prism(b, axis, offset, edge_length, angle);
    const int plane = thick_plane(b, axis, offset, height);
    return compound(b, p, plane);
}

int hexagonal_prism(rl_scene_builder *b, V3 axis, V3 offset, float edge_length, float bevel_size,
                    float angle, float height) {
    // geometry.rs:495-515
    const int ip = infinite_prism(b, axis, offset, edge_length * 2.0f - bevel_size * 3.0f, angle + PI);
    const int p = prism(b, axis, offset, edge_length, angle, height);
    return compound(b, ip, p);
}

rl_surface paraboloid(V3 normal, V3 offset, float focal_distance) {  // geometry.rs:286-295
    rl_surface s = blank(RL_SURFACE_PARABOLOID);
    s.a = rv(offset - normal * focal_distance);
    s.b = rv(normal);
    s.c = rv(normal * (focal_distance * 2.0f));
    return s;
}

// Paraboloid::intersect on the host (geometry.rs:299-358), used by
// set_up_scene to seat the prisms on the floor (app.rs:309-313).
bool paraboloid_hit(const rl_surface &p, V3 ray_origin, V3 ray_direction, V3 &pos, V3 &nrm) {
    const V3 offset = vr(p.a), normal = vr(p.b), focal_point = vr(p.c);
    const V3 origin = ray_origin - offset;
    const V3 focal_offset = origin - focal_point;
    const float n_dot_d = rl::dot(normal, ray_direction);
    const float n_dot_o = rl::dot(normal, origin);
    const float d_dot_f = rl::dot(ray_direction, focal_offset);
    const float a = n_dot_d * n_dot_d - 1.0f;
    const float b = 2.0f * n_dot_d * n_dot_o - 2.0f * d_dot_f;
    const float c = n_dot_o * n_dot_o - rl::magnitude_squared(focal_offset);
    float t;
    if (a == 0.0f) {
        const float t1 = -c / b;
        if (t1 < 0.0f) return false;
        t = t1;
    } else {
        const float d = b * b - 4.0f * a * c;
        if (d < 0.0f) return false;
        const float sqrt_d = std::sqrt(d);
        const float p1 = 0.5f * (-b + sqrt_d) / a;
        const float q1 = 0.5f * (-b - sqrt_d) / a;
        if (p1 > 0.0f && (p1 < q1 || q1 < 0.0f)) t = p1;
        else if (q1 > 0.0f) t = q1;
        else return false;
    }
    pos = ray_origin + ray_direction * t;
    const V3 local_pos = pos - offset;
    const V3 plane_pr = local_pos - normal * rl::dot(local_pos, normal);
    nrm = rl::normalise(focal_point - plane_pr);
    return true;
}

rl_material material(uint32_t kind, float p0 = 0.f, float p1 = 0.f, float p2 = 0.f) {
    rl_material m;
    m.kind = kind; m.p0 = p0; m.p1 = p1; m.p2 = p2;
    return m;
}

rl_material blackbody(float kelvins, float intensity) {      // material.rs:92-97
    const double wien = 2.897772126e-3;                       // constants.rs:25
    const float peak = (float)rl::boltzmann((wien / (double)kelvins) * 1.0e9, (double)kelvins);
    return material(RL_MATERIAL_BLACKBODY, kelvins, intensity / peak);
}

int add_object(rl_scene_builder *b, int surface, rl_material m) {
    rl_object o;
    o.surface = (uint32_t)surface;
    o.material = m;
    b->objects.push_back(o);
    return (int)b->objects.size() - 1;
}

int add_sphere(rl_scene_builder *b, V3 position, float radius) {  // geometry.rs:195-200
    rl_surface s = blank(RL_SURFACE_SPHERE);
    s.a = rv(position); s.s = radius * radius;
    return push(b, s);
}

int add_circle(rl_scene_builder *b, V3 normal, V3 position, float radius) {  // geometry.rs:142-148
    rl_surface s = blank(RL_SURFACE_CIRCLE);
    s.a = rv(normal); s.b = rv(position); s.s = radius * radius;
    return push(b, s);
}

int add_plane(rl_scene_builder *b, V3 normal, V3 offset) {        // geometry.rs:46-51
    rl_surface s = blank(RL_SURFACE_PLANE);
    s.a = rv(normal); s.b = rv(offset);
    return push(b, s);
}

rl_camera_model orbit_camera() {  // app.rs:327-357
    rl_camera_model cm;
    cm.kind = RL_CAMERA_ORBIT;
    cm.fixed.position = rl_vec3{0.f, 0.f, 0.f};
    cm.fixed.field_of_view = PI * 0.35f;
    cm.fixed.focal_distance = 0.0f;
    cm.fixed.depth_of_field = 2.0f;
    cm.fixed.chromatic_abberation = 0.012f;
    cm.fixed.orientation = rl_quat{0.f, 0.f, 0.f, 1.f};
    cm.phi_base = 1.0f; cm.phi_rate = 0.01f;
    cm.alpha_base = 0.3f; cm.alpha_rate = -0.01f;
    cm.distance_base = 50.0f; cm.distance_rate = -0.5f;
    cm.focal_factor = 0.9f;
    cm.keyframes = nullptr; cm.n_keyframes = 0;
    return cm;
}

rl_camera_model static_camera(V3 position, rl::Quat q, float fov, float focal, float dof, float ca) {
    rl_camera_model cm;
    cm.kind = RL_CAMERA_STATIC;
    cm.fixed.position = rv(position);
    cm.fixed.field_of_view = fov;
    cm.fixed.focal_distance = focal;
    cm.fixed.depth_of_field = dof;
    cm.fixed.chromatic_abberation = ca;
    cm.fixed.orientation = rl_quat{q.x, q.y, q.z, q.w};
    cm.phi_base = cm.phi_rate = cm.alpha_base = cm.alpha_rate = 0.f;
    cm.distance_base = cm.distance_rate = cm.focal_factor = 0.f;
    cm.keyframes = nullptr; cm.n_keyframes = 0;
    return cm;
}

// C1: one diffuse sphere + emissive plane, pinhole-like static camera.
void scene_c1(rl_scene_builder *b) {
    add_object(b, add_sphere(b, mk(0.f, 0.f, 0.f), 1.0f), material(RL_MATERIAL_DIFFUSE_GREY, 0.8f));
    add_object(b, add_plane(b, mk(0.f, 0.f, -1.f), mk(0.f, 0.f, 4.f)), blackbody(6504.0f, 1.0f));
    b->camera = static_camera(mk(0.f, -5.f, 0.f), rl::mkq(0.f, 0.f, 0.f, 1.f), PI * 0.35f, 5.0f, 1.0e9f, 0.0f);
}

// C2: App::set_up_scene (app.rs:166-363): 339 objects.
void scene_c2(rl_scene_builder *b) {
    const float sun_radius = 5.0f;
    const V3 sun_position = mk(0.f, 0.f, 0.f);
    add_object(b, add_sphere(b, sun_position, sun_radius), blackbody(6504.0f, 1.0f));    // :172-177

    const V3 floor_normal = mk(0.f, 0.f, -1.f);
    const V3 floor_position = mk(0.f, 0.f, -sun_radius);
    const rl_surface floor_paraboloid = paraboloid(floor_normal, floor_position, sun_radius * sun_radius);
    add_object(b, push(b, floor_paraboloid), material(RL_MATERIAL_DIFFUSE_GREY, 0.8f));   // :180-186

    const float sr2 = sun_radius * sun_radius;
    add_object(b, push(b, paraboloid(mk(0.f, 0.f, 1.f), mk(1.f, 0.f, -sr2), sr2)),
               material(RL_MATERIAL_DIFFUSE_COLOURED, 0.9f, 550.0f, 40.0f));              // :189-196
    add_object(b, push(b, paraboloid(mk(0.f, 0.f, 1.f), mk(-1.f, 0.f, -sr2), sr2)),
               material(RL_MATERIAL_DIFFUSE_COLOURED, 0.9f, 660.0f, 60.0f));              // :199-206

    const float sky_height = 30.0f;
    const float sky1_radius = 5.0f;
    add_object(b, add_circle(b, floor_normal, mk(-sun_radius, 0.f, sky_height), sky1_radius),
               blackbody(7600.0f, 0.6f));                                                 // :209-215
    const float sky2_radius = 15.0f;
    add_object(b, add_circle(b, floor_normal,
                             mk(-sun_radius * 0.5f, sun_radius * 2.0f + sky2_radius, sky_height),
                             sky2_radius),
               blackbody(5000.0f, 0.6f));                                                 // :217-224
    add_object(b, add_plane(b, floor_normal, mk(0.f, 0.f, sky_height * 2.0f)),
               material(RL_MATERIAL_DIFFUSE_COLOURED, 0.5f, 470.0f, 25.0f));              // :227-231

    const float golden_ratio = (float)1.6180339887498948482;                              // constants.rs:17
    const float gamma = PI * 2.0f * (1.0f - 1.0f / golden_ratio);                         // :234
    const float seed_size = 0.8f;
    const float seed_scale = 1.5f;
    const float fs = sun_radius / seed_scale + 1.0f;
    const long first_seed = (long)(fs * fs + 0.5f);                                       // :237
    const long seeds = 100;
    for (long i = first_seed; i < first_seed + seeds; i++) {                              // :239-253
        const float phi = (float)i * gamma;
        const float r = std::sqrt((float)i) * seed_scale;
        float s, c;
        rl::spec_sincos(phi, s, c);
        const V3 position = mk(c * r, s * r, (r - sun_radius) * -0.5f) + sun_position;
        add_object(b, add_sphere(b, position, seed_size),
                   material(RL_MATERIAL_DIFFUSE_COLOURED, 0.9f,
                            (float)(i - first_seed) / (float)seeds * 130.0f + 600.0f, 60.0f));
    }
    for (long i = first_seed; i < first_seed + seeds; i++) {                              // :256-268
        const float phi = ((float)i + 0.5f) * gamma;
        const float r = std::sqrt((float)i + 0.5f) * seed_scale;
        float s, c;
        rl::spec_sincos(phi, s, c);
        const V3 position = mk(c * r, s * r, (r - sun_radius) * -0.25f) + sun_position;
        add_object(b, add_sphere(b, position, seed_size * 0.5f), material(RL_MATERIAL_GLOSSY_MIRROR, 0.1f));
    }
    for (long i = first_seed / 2; i < first_seed + seeds; i++) {                          // :271-284
        const float phi = (float)(-i) * gamma;
        const float r = std::sqrt((float)i) * seed_scale * 1.5f;
        float s, c;
        rl::spec_sincos(phi, s, c);
        const V3 position = mk(c * r, s * r, (r - sun_radius) * 1.5f + sun_radius * 2.0f) + sun_position;
        add_object(b, add_sphere(b, position, seed_size * (0.5f + std::sqrt((float)i) * 0.2f)),
                   material(RL_MATERIAL_SOAP_BUBBLE));
    }

    const long prisms = 11;                                                               // :287-325
    const float prism_angle = PI * 2.0f / (float)prisms;
    const float prism_radius = 17.0f;
    const float prism_height = 8.0f;
    for (long i = 0; i < prisms; i++) {
        const float variants[2][4] = {{0.0f, 1.0f, 0.0f, 1.0f},
                                      {0.5f * prism_angle, 1.2f, PI * 0.5f, 1.5f}};
        for (int v = 0; v < 2; v++) {
            const float ofs = variants[v][0], radius = variants[v][1];
            const float phi_ofs = variants[v][2], h = variants[v][3];
            const float phi = (float)i * prism_angle + ofs;
            float s, c;
            rl::spec_sincos(phi, s, c);
            V3 position = mk(c * prism_radius * radius, s * prism_radius * radius, 0.0f);
            V3 normal = mk(0.f, 0.f, -1.f);
            V3 hit_pos, hit_nrm;
            if (paraboloid_hit(floor_paraboloid, position, normal, hit_pos, hit_nrm)) {
                normal = -hit_nrm;
                position = hit_pos + normal * 2.0f * h;
            }
            add_object(b, hexagonal_prism(b, normal, position, 3.0f, 1.0f, phi + phi_ofs, prism_height * h),
                       material(RL_MATERIAL_SF10_GLASS));
        }
    }
    b->camera = orbit_camera();
}

// C3: SF10 prism + emissive circle + grey floor, static camera tilted down.
void scene_c3(rl_scene_builder *b) {
    add_object(b, add_plane(b, mk(0.f, 0.f, 1.f), mk(0.f, 0.f, 0.f)), material(RL_MATERIAL_DIFFUSE_GREY, 0.8f));
    add_object(b, prism(b, mk(0.f, 0.f, 1.f), mk(0.f, 0.f, 0.5f), 4.0f, 0.3f, 3.0f),
               material(RL_MATERIAL_SF10_GLASS));
    add_object(b, add_circle(b, mk(0.f, 0.f, -1.f), mk(0.f, 2.f, 8.f), 3.0f), blackbody(7600.0f, 1.0f));
    const rl::Quat q = rl::rotation(1.0f, 0.0f, 0.0f, -0.3f);
    b->camera = static_camera(mk(0.f, -12.f, 5.f), q, PI * 0.35f, 12.0f, 40.0f, 0.0f);
}

struct Pcg32 {  // O'Neill's PCG-XSH-RR 64/32
    uint64_t state, inc;
    explicit Pcg32(uint64_t seed) : state(0), inc((54u << 1) | 1u) { next(); state += seed; next(); }
    uint32_t next() {
        const uint64_t old = state;
        state = old * 6364136223846793005ULL + inc;
        const uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        const uint32_t rot = (uint32_t)(old >> 59u);
        return (xs >> rot) | (xs << ((32 - rot) & 31));
    }
    float unit() { return (float)(next() >> 8) * 5.9604644775390625e-8f; }
};

// C4: n random spheres (intersection-bound stress), orbit camera.
void scene_c4(rl_scene_builder *b, uint32_t n) {
    if (n == 0) n = 4096;
    add_object(b, add_sphere(b, mk(0.f, 0.f, 0.f), 5.0f), blackbody(6504.0f, 1.0f));
    add_object(b, add_plane(b, mk(0.f, 0.f, -1.f), mk(0.f, 0.f, 60.f)), blackbody(5000.0f, 0.6f));
    Pcg32 rng(4096);
    for (uint32_t i = 0; i < n; i++) {
        const float x = rng.unit() * 40.0f - 20.0f;
        const float y = rng.unit() * 40.0f - 20.0f;
        const float z = rng.unit() * 40.0f - 20.0f;
        const float r = rng.unit() * 0.4f + 0.2f;
        rl_material m;
        switch (i & 3u) {
        case 0: m = material(RL_MATERIAL_DIFFUSE_GREY, 0.8f); break;
        case 1: m = material(RL_MATERIAL_DIFFUSE_COLOURED, 0.9f, 450.0f + (float)(i % 300u), 60.0f); break;
        case 2: m = material(RL_MATERIAL_GLOSSY_MIRROR, 0.1f); break;
        default: m = material(RL_MATERIAL_SOAP_BUBBLE); break;
        }
        add_object(b, add_sphere(b, mk(x, y, z), r), m);
    }
    b->camera = orbit_camera();
}

// C6: compounds over the reference's other Volume, the sphere (geometry.rs:263-267, :361-407):
// a biconvex SF10 lens = Compound<Sphere, Sphere>, a soap-bubble dome = Compound<Sphere,
// SpacePartitioning> (its spherical face carries the tangent of geometry.rs:250-251), a glossy
// lens cut by a slab = Compound<Compound<Sphere, Sphere>, ThickPlane>, a diffuse sphere clipped to
// a prism = Compound<Prism, Sphere>; grey floor, emissive circle and sphere, static camera.
void scene_c6(rl_scene_builder *b) {
    add_object(b, add_plane(b, mk(0.f, 0.f, 1.f), mk(0.f, 0.f, 0.f)), material(RL_MATERIAL_DIFFUSE_GREY, 0.8f));
    add_object(b, compound(b, add_sphere(b, mk(-1.2f, 0.f, 2.f), 2.0f), add_sphere(b, mk(1.2f, 0.f, 2.f), 2.0f)),
               material(RL_MATERIAL_SF10_GLASS));
    add_object(b, compound(b, add_sphere(b, mk(5.f, 1.f, 0.5f), 2.0f), halfspace(b, mk(0.f, 0.f, -1.f), mk(0.f, 0.f, 0.5f))),
               material(RL_MATERIAL_SOAP_BUBBLE));
    add_object(b, compound(b, compound(b, add_sphere(b, mk(-5.f, 0.f, 1.f), 2.5f), add_sphere(b, mk(-5.f, 0.f, 4.f), 2.5f)),
                           thick_plane(b, mk(1.f, 0.f, 0.f), mk(-5.5f, 0.f, 0.f), 1.0f)),
               material(RL_MATERIAL_GLOSSY_MIRROR, 0.1f));
    add_object(b, compound(b, prism(b, mk(0.f, 0.f, 1.f), mk(0.f, 5.f, 0.2f), 4.0f, 0.7f, 3.0f),
                           add_sphere(b, mk(0.f, 5.f, 1.5f), 1.8f)),
               material(RL_MATERIAL_DIFFUSE_COLOURED, 0.9f, 620.0f, 60.0f));
    add_object(b, add_circle(b, mk(0.f, 0.f, -1.f), mk(0.f, 2.f, 9.f), 4.0f), blackbody(7600.0f, 1.0f));
    add_object(b, add_sphere(b, mk(2.f, -3.f, 0.8f), 0.8f), blackbody(5000.0f, 0.5f));
    const rl::Quat q = rl::rotation(1.0f, 0.0f, 0.0f, -0.3f);
    b->camera = static_camera(mk(0.f, -14.f, 6.f), q, PI * 0.35f, 14.0f, 40.0f, 0.01f);
}

}  // namespace

extern "C" {

int rl_scene_builder_create(rl_scene_builder **out) {
    if (!out) return RL_ERR_INVALID;
    rl_scene_builder *b = new (std::nothrow) rl_scene_builder();
    if (!b) return RL_ERR_NOMEM;
    b->camera = static_camera(mk(0.f, 0.f, 0.f), rl::mkq(0.f, 0.f, 0.f, 1.f), PI * 0.35f, 1.0f, 1.0e9f, 0.0f);
    *out = b;
    return RL_OK;
}

int rl_scene_builder_destroy(rl_scene_builder *b) {
    delete b;
    return RL_OK;
}

int rl_scene_builder_builtin(rl_scene_builder *b, int which, uint32_t param) {
    if (!b) return RL_ERR_INVALID;
    b->surfaces.clear();
    b->objects.clear();
    switch (which) {
    case RL_SCENE_C1_SPHERE_PLANE: scene_c1(b); break;
    case RL_SCENE_C2_BUILTIN: scene_c2(b); break;
    case RL_SCENE_C3_PRISM: scene_c3(b); break;
    case RL_SCENE_C4_SPHERES: scene_c4(b, param); break;
    case RL_SCENE_C6_LENSES: scene_c6(b); break;
    default: return RL_ERR_INVALID;
    }
    return RL_OK;
}

int rl_scene_builder_plane(rl_scene_builder *b, rl_vec3 normal, rl_vec3 offset) {
    if (!b) return RL_ERR_INVALID;
    return add_plane(b, vr(normal), vr(offset));
}
int rl_scene_builder_circle(rl_scene_builder *b, rl_vec3 normal, rl_vec3 position, float radius) {
    if (!b) return RL_ERR_INVALID;
    return add_circle(b, vr(normal), vr(position), radius);
}
int rl_scene_builder_halfspace(rl_scene_builder *b, rl_vec3 normal, rl_vec3 offset) {
    if (!b) return RL_ERR_INVALID;
    return halfspace(b, vr(normal), vr(offset));
}

int rl_scene_builder_compound(rl_scene_builder *b, uint32_t surface1, uint32_t surface2) {
    if (!b || surface1 >= b->surfaces.size() || surface2 >= b->surfaces.size()) return RL_ERR_INVALID;
    return compound(b, (int)surface1, (int)surface2);
}

int rl_scene_builder_sphere(rl_scene_builder *b, rl_vec3 position, float radius) {
    if (!b) return RL_ERR_INVALID;
    return add_sphere(b, vr(position), radius);
}
int rl_scene_builder_paraboloid(rl_scene_builder *b, rl_vec3 normal, rl_vec3 offset, float focal_distance) {
    if (!b) return RL_ERR_INVALID;
    return push(b, paraboloid(vr(normal), vr(offset), focal_distance));
}
int rl_scene_builder_prism(rl_scene_builder *b, rl_vec3 axis, rl_vec3 offset, float edge_length,
                           float angle, float height) {
    if (!b) return RL_ERR_INVALID;
    return prism(b, vr(axis), vr(offset), edge_length, angle, height);
}
int rl_scene_builder_hexagonal_prism(rl_scene_builder *b, rl_vec3 axis, rl_vec3 offset,
                                     float edge_length, float bevel_size, float angle, float height) {
    if (!b) return RL_ERR_INVALID;
    return hexagonal_prism(b, vr(axis), vr(offset), edge_length, bevel_size, angle, height);
}
int rl_material_blackbody(float kelvins, float intensity, rl_material *out) {
    if (!out) return RL_ERR_INVALID;
    *out = blackbody(kelvins, intensity);
    return RL_OK;
}
int rl_scene_builder_object(rl_scene_builder *b, uint32_t surface, rl_material m) {
    if (!b || surface >= b->surfaces.size()) return RL_ERR_INVALID;
    return add_object(b, (int)surface, m);
}
int rl_scene_builder_camera(rl_scene_builder *b, const rl_camera_model *camera) {
    if (!b || !camera) return RL_ERR_INVALID;
    b->camera = *camera;
    b->keyframes.clear();
    if (camera->kind == RL_CAMERA_KEYFRAMES) {                 // the builder keeps its own copy of the table
        if (!camera->keyframes || camera->n_keyframes == 0) return RL_ERR_INVALID;
        b->keyframes.assign(camera->keyframes, camera->keyframes + camera->n_keyframes);
    }
    b->camera.keyframes = b->keyframes.empty() ? nullptr : b->keyframes.data();
    b->camera.n_keyframes = (uint32_t)b->keyframes.size();
    return RL_OK;
}
int rl_scene_builder_desc(rl_scene_builder *b, rl_scene_desc *out) {
    if (!b || !out) return RL_ERR_INVALID;
    out->surfaces = b->surfaces.data();
    out->n_surfaces = (uint32_t)b->surfaces.size();
    out->objects = b->objects.data();
    out->n_objects = (uint32_t)b->objects.size();
    out->camera = b->camera;
    return RL_OK;
}

}  // extern "C"
