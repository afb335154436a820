/*
 * oracle.cpp -- CPU restatement of the Robigo Luculenta hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this; the product
 * (robigo-luculenta_b200/) never links, imports or calls it.
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference ships no golden vectors or
 * numeric tests for this path (src/main.rs:69-74 asserts nothing) and draws
 * every random number from an unseeded OS-seeded generator
 * (src/monte_carlo.rs:22-28, crate rand 0.3.11 per Cargo.lock:80-81, source
 * not vendored), and no Rust toolchain exists in this environment, so the
 * reference binary cannot produce fixtures either.  What pins this file is
 * (a) the closed-form known-answer tests derived from the reference source
 * (tests/test_oracle_kat.py), (b) a second, independent restatement in numpy
 * that it agrees with bit for bit -- Scene::intersect, whole photon paths,
 * plot, gather, tonemap, and the built-in scene (tests/test_oracle_numpy_cross_check.py,
 * tests/test_oracle_path_cross_check.py, tests/test_scene_builder_cross_check.py) --
 * and (c) line-by-line citation below.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference's src/).  It is a restatement over a flattened POD scene
 * (include/rl_b200.h), not a transliteration: no trait objects, one recursive
 * evaluator for compound surfaces, a counter-based RNG.
 *
 * Arithmetic.  f32 everywhere the reference is f32, f64 where it is f64, each
 * operation rounded separately in the reference's evaluation order (build
 * with -ffp-contract=off; rustc does not contract).  Two math modes:
 *   ORC_MATH_LIBM (0): sin/cos/tan/exp/acos/ln/pow from glibc -- what the
 *       Rust binary calls on x86-64 Linux (f32::sin -> sinf, ...).
 *   ORC_MATH_SPEC (1): the same code with those functions replaced by the
 *       fully specified polynomial versions in struct SpecMath (explicit
 *       fused multiply-adds, IEEE + - * / sqrt only).  The CUDA path
 *       implements the same specification, so GPU results are compared
 *       bit-for-bit against this mode; LIBM vs SPEC is compared statistically.
 *
 * RNG.  rand::random has no seed, so both modes draw from Philox4x32-10
 * (Salmon et al., SC'11) keyed by (seed, photon id), counter (id_lo, id_hi,
 * block, 0): the six draws in front of the first bounce (wavelength, x, y, t,
 * lens angle, lens radius) are the words of blocks 0 and 1 in order; the draws
 * of bounce j (the material's, then the roulette's -- at most three) are the
 * first words of block 2 + j.  A bounce therefore needs exactly one block,
 * known before the ray is traced.  The two
 * distributions are rand 0.3.11's: f32 = (u32 >> 8) * 2^-24 in [0,1);
 * Closed01<f32> = that * 2^24 / (2^24 - 1) in [0,1].
 */
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/rl_b200.h"
#include "oracle.h"

namespace {

// ---------------------------------------------------------------- bit casts
inline float f32_from_bits(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline double f64_from_bits(uint64_t u) { double f; std::memcpy(&f, &u, 8); return f; }

// ------------------------------------------------------------------- math
// The reference calls libm through std (f32::sin etc.).
struct LibmMath {
    static inline void sincos(float x, float &s, float &c) { s = ::sinf(x); c = ::cosf(x); }
    static inline float tan(float x) { return ::tanf(x); }
    static inline float exp(float x) { return ::expf(x); }
    static inline float acos(float x) { return ::acosf(x); }
    static inline float ln(float x) { return ::logf(x); }
    static inline float pow(float x, float y) { return ::powf(x, y); }
    static inline double exp64(double x) { return ::exp(x); }
};

// The specification the CUDA path shares (DESIGN.md "Specified math").
struct SpecMath {
    // Cody-Waite reduction by pi/2 (two-term, fused), Cephes sinf/cosf minimax
    // polynomials on [-pi/4, pi/4], quadrant select.
    static inline void sincos(float x, float &s, float &c) {
        float kf = ::rintf(x * 0.636619747f);
        float r = ::fmaf(kf, -1.57079637f, x);
        r = ::fmaf(kf, 4.37113883e-8f, r);
        float z = r * r;
        float ps = ::fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
        ps = ::fmaf(ps, z, -1.6666654611e-1f);
        float sn = ::fmaf(ps * z, r, r);
        float pc = ::fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
        pc = ::fmaf(pc, z, 4.166664568298827e-2f);
        float cs = ::fmaf(pc * z, z, ::fmaf(-0.5f, z, 1.0f));
        int q = (int)kf & 3;
        float s0 = (q & 1) ? cs : sn;
        float c0 = (q & 1) ? sn : cs;
        s = (q & 2) ? -s0 : s0;
        c = ((q + 1) & 2) ? -c0 : c0;
    }
    static inline float tan(float x) { float s, c; sincos(x, s, c); return s / c; }
    // Cephes expf: k = rint(x log2 e), two-term ln2 reduction, degree-5 core.
    static inline float exp(float x) {
        if (!(x == x)) return x;
        if (x < -104.0f) return 0.0f;
        if (x > 88.0f) x = 88.0f;
        float kf = ::rintf(x * 1.44269502f);
        float r = ::fmaf(kf, -0.693359375f, x);
        r = ::fmaf(kf, 2.12194442e-4f, r);
        float z = r * r;
        float p = 1.9875691500e-4f;
        p = ::fmaf(p, r, 1.3981999507e-3f);
        p = ::fmaf(p, r, 8.3334519073e-3f);
        p = ::fmaf(p, r, 4.1665795894e-2f);
        p = ::fmaf(p, r, 1.6666665459e-1f);
        p = ::fmaf(p, r, 5.0000001201e-1f);
        float y = ::fmaf(p, z, r) + 1.0f;
        int k = (int)kf;
        if (k >= -126) return y * f32_from_bits((uint32_t)(k + 127) << 23);
        return (y * f32_from_bits((uint32_t)(k + 227) << 23)) * f32_from_bits(27u << 23);
    }
    // Cephes asinf core; acos by the usual two-range identities.
    static inline float acos(float x) {
        float a = ::fabsf(x);
        bool big = a > 0.5f;
        float zz, w;
        if (big) { zz = (1.0f - a) * 0.5f; w = ::sqrtf(zz); }
        else { zz = a * a; w = a; }
        float p = 4.2163199048e-2f;
        p = ::fmaf(p, zz, 2.4181311049e-2f);
        p = ::fmaf(p, zz, 4.5470025998e-2f);
        p = ::fmaf(p, zz, 7.4953002686e-2f);
        p = ::fmaf(p, zz, 1.6666752422e-1f);
        float as = ::fmaf(p * zz, w, w);
        if (big) { float t = as + as; return x < 0.0f ? 3.14159274f - t : t; }
        return x < 0.0f ? 1.57079637f + as : 1.57079637f - as;
    }
    // ln x = k ln2 + ln m, m in [sqrt(1/2), sqrt 2); Cephes logf core.
    static inline float ln(float x) {
        if (!(x > 0.0f)) return x == 0.0f ? -INFINITY : NAN;
        if (x == INFINITY) return x;
        uint32_t u; std::memcpy(&u, &x, 4);
        int e = 0;
        if (u < 0x00800000u) { x *= 8388608.0f; std::memcpy(&u, &x, 4); e = -23; }
        e += (int)(u >> 23) - 126;
        float m = f32_from_bits((u & 0x007fffffu) | 0x3f000000u);  // [0.5, 1)
        if (m < 0.707106769f) { e -= 1; m = m + m; }
        float t = m - 1.0f;
        float z = t * t;
        float p = 7.0376836292e-2f;
        p = ::fmaf(p, t, -1.1514610310e-1f);
        p = ::fmaf(p, t, 1.1676998740e-1f);
        p = ::fmaf(p, t, -1.2420140846e-1f);
        p = ::fmaf(p, t, 1.4249322787e-1f);
        p = ::fmaf(p, t, -1.6668057665e-1f);
        p = ::fmaf(p, t, 2.0000714765e-1f);
        p = ::fmaf(p, t, -2.4999993993e-1f);
        p = ::fmaf(p, t, 3.3333331174e-1f);
        float y = (t * z) * p;
        float ef = (float)e;
        y = ::fmaf(ef, -2.12194440e-4f, y);
        y = ::fmaf(-0.5f, z, y);
        float r = t + y;
        return ::fmaf(ef, 0.693359375f, r);
    }
    // x^y = exp(y ln x) for x > 0 (the only use is gamma, srgb.rs:24).
    static inline float pow(float x, float y) { return exp(y * ln(x)); }
    // f64 exp: k = rint(x log2 e), fdlibm ln2 hi/lo, Taylor degree 13.
    static inline double exp64(double x) {
        if (!(x == x)) return x;
        if (x > 709.0) return INFINITY;
        if (x < -708.0) return 0.0;
        double kf = ::rint(x * 1.4426950408889634);
        double r = ::fma(kf, -0.6931471803691238, x);
        r = ::fma(kf, -1.9082149292705877e-10, r);
        double p = 1.6059043836821613e-10;
        p = ::fma(p, r, 2.08767569878681e-09);
        p = ::fma(p, r, 2.505210838544172e-08);
        p = ::fma(p, r, 2.755731922398589e-07);
        p = ::fma(p, r, 2.7557319223985893e-06);
        p = ::fma(p, r, 2.48015873015873e-05);
        p = ::fma(p, r, 0.0001984126984126984);
        p = ::fma(p, r, 0.001388888888888889);
        p = ::fma(p, r, 0.008333333333333333);
        p = ::fma(p, r, 0.041666666666666664);
        p = ::fma(p, r, 0.16666666666666666);
        p = ::fma(p, r, 0.5);
        p = ::fma(p, r, 1.0);
        p = ::fma(p, r, 1.0);
        int64_t k = (int64_t)kf;
        return p * f64_from_bits((uint64_t)(k + 1023) << 52);
    }
};

// ----------------------------------------------------------------- vectors
struct V3 { float x, y, z; };
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 v3(const rl_vec3 &v) { return V3{v.x, v.y, v.z}; }
inline rl_vec3 to_rl(const V3 &v) { return rl_vec3{v.x, v.y, v.z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }   // vector3.rs:96-106
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }   // vector3.rs:108-118
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }                        // vector3.rs:120-130
inline V3 operator*(V3 a, float f) { return {a.x * f, a.y * f, a.z * f}; }      // vector3.rs:132-142
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }      // vector3.rs:35-37
inline V3 cross(V3 a, V3 b) {                                                   // vector3.rs:27-33
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float magnitude_squared(V3 a) { return dot(a, a); }                      // vector3.rs:48-50
inline V3 normalise(V3 a) {                                                     // vector3.rs:56-67
    float m = std::sqrt(magnitude_squared(a));
    if (m == 0.0f) return a;
    return {a.x / m, a.y / m, a.z / m};
}
inline V3 rotate_towards(V3 v, V3 n) {                                          // vector3.rs:69-83
    float d = n.z;
    if (d > 0.9999f) return v;
    if (d < -0.9999f) return {v.x, v.y, -v.z};
    V3 up = {0.0f, 0.0f, 1.0f};
    V3 a1 = normalise(cross(up, n));
    V3 a2 = normalise(cross(a1, n));
    return a1 * v.x + a2 * v.y + n * v.z;
}
inline V3 reflect(V3 v, V3 n) { return v - n * 2.0f * dot(n, v); }              // vector3.rs:91-93

struct Quat { float x, y, z, w; };
inline Quat conjugate(Quat q) { return {-q.x, -q.y, -q.z, q.w}; }               // quaternion.rs:47-49
inline Quat operator*(Quat a, Quat b) {                                         // quaternion.rs:100-110
    return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
            a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w,
            a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
template <class M>
inline Quat rotation(float x, float y, float z, float angle) {                  // quaternion.rs:36-45
    float s, c;
    M::sincos(angle * 0.5f, s, c);
    return {s * x, s * y, s * z, c};
}
inline V3 rotate(V3 v, Quat q) {                                                // vector3.rs:85-89
    Quat p = {v.x, v.y, v.z, 0.0f};
    Quat r = q * p * conjugate(q);
    return {r.x, r.y, r.z};
}

struct Ray { V3 origin, direction; float wavelength, probability; };            // ray.rs:19-33
struct Isect { V3 position, normal, tangent; float distance; };                 // intersection.rs:19-32

const float PI = 3.14159265358979323846f;                                       // std::f32::consts::PI

// --------------------------------------------------------------------- RNG
struct Philox {
    uint32_t key0, key1;
    uint32_t c0, c1;      // photon id
    uint32_t block;
    uint32_t buf[4];
    int idx;
    Philox(uint64_t seed, uint64_t photon)
        : key0((uint32_t)seed), key1((uint32_t)(seed >> 32)), c0((uint32_t)photon),
          c1((uint32_t)(photon >> 32)), block(0), idx(4) {}
    static inline void block4x32_10(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1,
                                    uint32_t x2, uint32_t x3, uint32_t out[4]) {
        for (int r = 0; r < 10; r++) {
            uint64_t p0 = (uint64_t)0xD2511F53u * x0;
            uint64_t p1 = (uint64_t)0xCD9E8D57u * x2;
            uint32_t y0 = (uint32_t)(p1 >> 32) ^ x1 ^ k0;
            uint32_t y1 = (uint32_t)p1;
            uint32_t y2 = (uint32_t)(p0 >> 32) ^ x3 ^ k1;
            uint32_t y3 = (uint32_t)p0;
            x0 = y0; x1 = y1; x2 = y2; x3 = y3;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
    }
    inline uint32_t next_u32() {
        if (idx == 4) { block4x32_10(key0, key1, c0, c1, block, 0u, buf); block++; idx = 0; }
        return buf[idx++];
    }
    // The draws of bounce j (the material's, then the roulette's: at most three) are the first
    // words of block 2 + j; the six draws in front of the first bounce (wavelength, x, y, t, lens
    // angle and radius) are blocks 0 and 1.  Unused words of a block are dropped.
    inline void begin_bounce(uint32_t j) { block = 2u + j; idx = 4; }
    // monte_carlo.rs:25-28 -- Closed01<f32> of rand 0.3.11
    inline float unit() { return (float)(next_u32() >> 8) / 16777215.0f; }
    // monte_carlo.rs:37 -- rand::random::<f32>()
    inline float half_open() { return (float)(next_u32() >> 8) * 5.9604644775390625e-8f; }
    inline float bi_unit() { return unit() * 2.0f - 1.0f; }                     // monte_carlo.rs:31-33
    inline float longitude() { return half_open() * PI * 2.0f; }                // monte_carlo.rs:36-38
    inline float wavelength() { return unit() * 400.0f + 380.0f; }              // monte_carlo.rs:41-43
};

template <class M>
inline V3 hemisphere_vector(Philox &rng) {                                      // monte_carlo.rs:47-58
    float phi = rng.longitude();
    float rq = rng.unit();
    float r = std::sqrt(rq);
    float s, c;
    M::sincos(phi, s, c);
    return {c * r, s * r, std::sqrt(1.0f - rq)};
}

// ---------------------------------------------------------------- geometry
// geometry.rs:55-71
inline bool intersect_plane(V3 normal, V3 offset, const Ray &ray, V3 &pos, float &t, float &d) {
    V3 origin = ray.origin - offset;
    d = dot(normal, ray.direction);
    if (d == 0.0f) return false;
    t = -dot(normal, origin) / d;
    if (t <= 0.0f) return false;
    pos = ray.origin + ray.direction * t;
    return true;
}

struct SceneView {
    const rl_surface *surfaces;
    uint32_t n_surfaces;
    const rl_object *objects;
    uint32_t n_objects;
    rl_camera_model camera;
};

// Volume::lies_inside: geometry.rs:124-128 (half-space), :263-267 (sphere), :403-407 (compound)
bool lies_inside(const SceneView &sc, uint32_t node, V3 p) {
    const rl_surface &s = sc.surfaces[node];
    if (s.kind == RL_SURFACE_HALFSPACE) return dot(p - v3(s.b), v3(s.a)) < 0.0f;
    if (s.kind == RL_SURFACE_SPHERE) return magnitude_squared(p - v3(s.a)) < s.s;
    return lies_inside(sc, s.child[0], p) && lies_inside(sc, s.child[1], p);
}

// Surface::intersect for every node kind
bool surface_intersect(const SceneView &sc, uint32_t node, const Ray &ray, Isect &out,
                       uint64_t *tests) {
    const rl_surface &s = sc.surfaces[node];
    switch (s.kind) {
    case RL_SURFACE_PLANE: {                                                    // geometry.rs:73-87
        if (tests) ++*tests;
        V3 pos; float t, d;
        if (!intersect_plane(v3(s.a), v3(s.b), ray, pos, t, d)) return false;
        out.position = pos;
        out.normal = d < 0.0f ? v3(s.a) : -v3(s.a);
        out.tangent = {0.0f, 0.0f, 0.0f};
        out.distance = t;
        return true;
    }
    case RL_SURFACE_HALFSPACE: {                                                // geometry.rs:109-122
        if (tests) ++*tests;
        V3 pos; float t, d;
        if (!intersect_plane(v3(s.a), v3(s.b), ray, pos, t, d)) return false;
        out.position = pos;
        out.normal = v3(s.a);
        out.tangent = {0.0f, 0.0f, 0.0f};
        out.distance = t;
        return true;
    }
    case RL_SURFACE_CIRCLE: {                                                   // geometry.rs:166-184
        if (tests) ++*tests;
        V3 pos; float t, d;
        if (!intersect_plane(v3(s.a), v3(s.b), ray, pos, t, d)) return false;
        if (!(magnitude_squared(pos - v3(s.b)) <= s.s)) return false;
        out.position = pos;
        out.normal = d < 0.0f ? v3(s.a) : -v3(s.a);
        out.tangent = {0.0f, 0.0f, 0.0f};
        out.distance = t;
        return true;
    }
    case RL_SURFACE_SPHERE: {                                                   // geometry.rs:204-261
        if (tests) ++*tests;
        float a = 1.0f;
        V3 centre_offset = v3(s.a) - ray.origin;
        float b = 2.0f * dot(ray.direction, centre_offset);
        float c = magnitude_squared(centre_offset) - s.s;
        float discriminant = b * b - 4.0f * a * c;
        if (discriminant < 0.0f) return false;
        float d = std::sqrt(discriminant);
        float t1 = -0.5f * (-b + d) / a;
        float t2 = -0.5f * (-b - d) / a;
        float t;
        if (t1 > 0.0f && t1 < t2) t = t1;
        else if (t2 > 0.0f && t2 < t1) t = t2;
        else return false;
        V3 position = ray.origin + ray.direction * t;
        V3 normal = normalise(position - v3(s.a));
        V3 up = {0.0f, 1.0f, 0.0f};
        out.position = position;
        out.normal = normal;
        out.tangent = normalise(cross(up, normal));
        out.distance = t;
        return true;
    }
    case RL_SURFACE_PARABOLOID: {                                               // geometry.rs:299-358
        if (tests) ++*tests;
        V3 offset = v3(s.a), normal = v3(s.b), focal_point = v3(s.c);
        V3 origin = ray.origin - offset;
        V3 focal_offset = origin - focal_point;
        float n_dot_d = dot(normal, ray.direction);
        float n_dot_o = dot(normal, origin);
        float d_dot_f = dot(ray.direction, focal_offset);
        float a = n_dot_d * n_dot_d - 1.0f;
        float b = 2.0f * n_dot_d * n_dot_o - 2.0f * d_dot_f;
        float c = n_dot_o * n_dot_o - magnitude_squared(focal_offset);
        float t;
        if (a == 0.0f) {
            float t1 = -c / b;
            if (t1 < 0.0f) return false;
            t = t1;
        } else {
            float d = b * b - 4.0f * a * c;
            if (d < 0.0f) return false;
            float sqrt_d = std::sqrt(d);
            float p = 0.5f * (-b + sqrt_d) / a;
            float q = 0.5f * (-b - sqrt_d) / a;
            if (p > 0.0f && (p < q || q < 0.0f)) t = p;
            else if (q > 0.0f) t = q;
            else return false;
        }
        V3 pos = ray.origin + ray.direction * t;
        V3 local_pos = pos - offset;
        V3 plane_pr = local_pos - normal * dot(local_pos, normal);
        out.position = pos;
        out.normal = normalise(focal_point - plane_pr);
        out.tangent = {0.0f, 0.0f, 0.0f};
        out.distance = t;
        return true;
    }
    case RL_SURFACE_COMPOUND: {                                                 // geometry.rs:380-401
        Isect i1, i2;
        bool h1 = surface_intersect(sc, s.child[0], ray, i1, tests);
        bool h2 = surface_intersect(sc, s.child[1], ray, i2, tests);
        if (h1) h1 = lies_inside(sc, s.child[1], i1.position);
        if (h2) h2 = lies_inside(sc, s.child[0], i2.position);
        if (h1 && h2) { out = (i1.distance < i2.distance) ? i1 : i2; return true; }
        if (h1) { out = i1; return true; }
        if (h2) { out = i2; return true; }
        return false;
    }
    }
    return false;
}

// scene.rs:39-60
int scene_intersect(const SceneView &sc, const Ray &ray, Isect &best, uint64_t *tests) {
    int result = -1;
    float distance = 1.0e12f;
    for (uint32_t i = 0; i < sc.n_objects; i++) {
        Isect isect;
        if (surface_intersect(sc, sc.objects[i].surface, ray, isect, tests)) {
            if (isect.distance < distance) {
                best = isect;
                result = (int)i;
                distance = isect.distance;
            }
        }
    }
    return result;
}

// --------------------------------------------------------------- materials
template <class M>
inline Ray get_diffuse_ray(const Ray &in, const Isect &is, Philox &rng) {       // material.rs:38-58
    V3 hemi = hemisphere_vector<M>(rng);
    V3 normal = dot(in.direction, is.normal) < 0.0f ? is.normal : -is.normal;
    return Ray{is.position, rotate_towards(hemi, normal), in.wavelength, 1.0f};
}

template <class M>
inline double boltzmann(double wavelength, double temperature) {                // material.rs:61-74
    const double h = 6.62606957e-34;   // constants.rs:19
    const double k = 1.3806488e-23;    // constants.rs:21
    const double c = 299792458.0;      // constants.rs:23
    double f = c / (wavelength * 1.0e-9);
    return (2.0 * h * f * f * f) / (c * c * (M::exp64(h * f / (k * temperature)) - 1.0));
}

template <class M>
inline float blackbody_intensity(const rl_material &m, float wavelength) {      // material.rs:101-105
    return (float)boltzmann<M>((double)wavelength, (double)m.p0) * m.p1;
}

inline float sf10_index_of_refraction(float wavelength) {                       // material.rs:203-213
    double w2 = (double)(wavelength * wavelength * 1.0e-6f);
    return (float)std::sqrt(1.0 + 1.737596950 * w2 / (w2 - 0.0131887070)
                            + 0.313747346 * w2 / (w2 - 0.0623068142)
                            + 1.898781010 * w2 / (w2 - 155.23629000));
}

template <class M>
Ray material_new_ray(const rl_material &m, const Ray &in, const Isect &is, Philox &rng) {
    switch (m.kind) {
    case RL_MATERIAL_DIFFUSE_GREY: {                                            // material.rs:122-130
        Ray ray = get_diffuse_ray<M>(in, is, rng);
        ray.probability = m.p0;
        return ray;
    }
    case RL_MATERIAL_DIFFUSE_COLOURED: {                                        // material.rs:155-168
        float p = (m.p1 - in.wavelength) / m.p2;
        float q = M::exp(-0.5f * p * p);
        Ray ray = get_diffuse_ray<M>(in, is, rng);
        ray.probability = m.p0 * q;
        return ray;
    }
    case RL_MATERIAL_GLOSSY_MIRROR: {                                           // material.rs:185-196
        Ray ray = get_diffuse_ray<M>(in, is, rng);
        V3 reflection = reflect(in.direction, is.normal);
        ray.direction = normalise(ray.direction * m.p0 + reflection * (1.0f - m.p0));
        return ray;
    }
    case RL_MATERIAL_SF10_GLASS: {                                              // material.rs:216-261
        float cos_i = -dot(in.direction, is.normal);
        float ior = sf10_index_of_refraction(in.wavelength);
        V3 normal = is.normal;
        if (cos_i > 0.0f) {
            ior = 1.0f / ior;
        } else {
            normal = -normal;
            cos_i = -cos_i;
        }
        float sin_t_sqr = ior * ior * (1.0f - cos_i * cos_i);
        V3 dir;
        if (sin_t_sqr > 1.0f) {
            dir = reflect(in.direction, normal);
        } else {
            float cos_t = std::sqrt(1.0f - sin_t_sqr);
            dir = in.direction * ior + normal * (ior * cos_i - cos_t);
        }
        return Ray{is.position, dir, in.wavelength, 1.0f};
    }
    case RL_MATERIAL_SOAP_BUBBLE: {                                             // material.rs:267-306
        float cos_alpha = dot(in.direction, is.normal);
        V3 direction = (rng.unit() - 0.3f > std::fabs(cos_alpha))
                           ? reflect(in.direction, is.normal)
                           : in.direction;
        float phase_shift = (in.wavelength - 380.0f) / 200.0f * PI;
        auto clamp = [](float x) { return x < -0.999f ? -0.999f : (x > 0.999f ? 0.999f : x); };
        float cos_phi = clamp(dot(direction, is.normal));
        float cos_theta = clamp(dot(direction, is.tangent));
        float s, p;
        M::sincos(phase_shift - M::acos(cos_phi) * 3.0f - M::acos(cos_theta) * 2.0f + PI * 0.5f,
                  s, p);
        (void)s;
        return Ray{is.position, direction, in.wavelength, p * 0.1f + 0.9f};
    }
    }
    return in;  // unreachable: descriptors are validated
}

// ------------------------------------------------------------------ camera
struct Camera {                                                                 // camera.rs:21-44
    V3 position;
    float field_of_view, focal_distance, depth_of_field, chromatic_abberation;
    Quat orientation;
};

template <class M>
inline Camera camera_at_time(const rl_camera_model &cm, float t) {
    Camera cam;
    if (cm.kind == RL_CAMERA_KEYFRAMES) {
        // a tabulated `fn(f32) -> Camera` (scene.rs:34): frame floor(t n), the last one for t == 1
        uint32_t k = (uint32_t)std::floor(t * (float)cm.n_keyframes);
        if (k > cm.n_keyframes - 1u) k = cm.n_keyframes - 1u;
        const rl_camera &f = cm.keyframes[k];
        cam.position = v3(f.position);
        cam.field_of_view = f.field_of_view;
        cam.focal_distance = f.focal_distance;
        cam.depth_of_field = f.depth_of_field;
        cam.chromatic_abberation = f.chromatic_abberation;
        cam.orientation = {f.orientation.x, f.orientation.y, f.orientation.z, f.orientation.w};
        return cam;
    }
    cam.field_of_view = cm.fixed.field_of_view;
    cam.depth_of_field = cm.fixed.depth_of_field;
    cam.chromatic_abberation = cm.fixed.chromatic_abberation;
    if (cm.kind == RL_CAMERA_STATIC) {
        cam.position = v3(cm.fixed.position);
        cam.focal_distance = cm.fixed.focal_distance;
        cam.orientation = {cm.fixed.orientation.x, cm.fixed.orientation.y,
                           cm.fixed.orientation.z, cm.fixed.orientation.w};
        return cam;
    }
    // app.rs:327-357 (make_camera), constants lifted into the descriptor
    float phi = PI * (cm.phi_base + cm.phi_rate * t);
    float alpha = PI * (cm.alpha_base + cm.alpha_rate * t);
    float distance = cm.distance_base + cm.distance_rate * t;
    float sa, ca, sp, cp;
    M::sincos(alpha, sa, ca);
    M::sincos(phi, sp, cp);
    cam.position = {ca * sp * distance, ca * cp * distance, sa * distance};
    cam.orientation = rotation<M>(0.0f, 0.0f, -1.0f, phi + PI) * rotation<M>(1.0f, 0.0f, 0.0f, -alpha);
    cam.focal_distance = distance * cm.focal_factor;
    return cam;
}

template <class M>
inline Ray camera_get_ray(const Camera &cam, float x, float y, float wavelength, Philox &rng) {
    // camera.rs:94-108
    float dof_angle = rng.longitude();
    float dof_radius = rng.unit() / cam.depth_of_field;
    float d = (wavelength - 580.0f) / 200.0f;
    float chromatic_zoom = 1.0f + d * cam.chromatic_abberation;
    // camera.rs:47-90
    float screen_distance = 1.0f / M::tan(cam.field_of_view * 0.5f);
    float xs = x * chromatic_zoom;
    float ys = y * chromatic_zoom;
    V3 direction = normalise(v3(xs, screen_distance, -ys));
    V3 focus_point = direction * (cam.focal_distance / direction.y);
    float sd, cd;
    M::sincos(dof_angle, sd, cd);
    V3 lens_point = {cd * dof_radius, 0.0f, sd * dof_radius};
    Ray r;
    r.origin = cam.position + rotate(lens_point, cam.orientation);
    r.direction = normalise(rotate(focus_point - lens_point, cam.orientation));
    r.wavelength = wavelength;
    r.probability = 1.0f;
    return r;
}

// ------------------------------------------------------------------- trace
struct Counters { uint64_t photons = 0, rays = 0, tests = 0, hits_emissive = 0, bounces_max = 0; };

template <class M>
float render_ray(const SceneView &sc, Ray ray, Philox &rng, Counters &ct, bool count_tests) {
    // trace_unit.rs:81-132
    float continue_chance = 1.0f;
    float intensity = 1.0f;
    uint64_t bounces = 0;
    for (;;) {
        rng.begin_bounce((uint32_t)bounces);
        Isect is;
        ct.rays++;
        int obj = scene_intersect(sc, ray, is, count_tests ? &ct.tests : nullptr);
        if (obj < 0) return 0.0f;
        const rl_material &mat = sc.objects[obj].material;
        if (mat.kind == RL_MATERIAL_BLACKBODY) {
            ct.hits_emissive++;
            return intensity * blackbody_intensity<M>(mat, ray.wavelength);
        }
        ray = material_new_ray<M>(mat, ray, is, rng);
        intensity = intensity * ray.probability;
        ray.origin = ray.origin + ray.direction * 0.00001f;
        continue_chance = continue_chance * 0.96f;
        bounces++;
        if (bounces > ct.bounces_max) ct.bounces_max = bounces;
        if (rng.unit() * 0.85f > continue_chance * (1.0f - M::exp(intensity * -20.0f))) break;
    }
    return 0.0f;
}

template <class M>
void trace_range(const SceneView &sc, uint64_t seed, uint64_t first, uint64_t n, float aspect,
                 rl_mapped_photon *out, Counters &ct, bool count_tests) {
    // trace_unit.rs:151-168 + :136-148
    for (uint64_t i = 0; i < n; i++) {
        Philox rng(seed, first + i);
        float wavelength = rng.wavelength();
        float x = rng.bi_unit();
        float y = rng.bi_unit() / aspect;
        float t = rng.unit();
        Camera cam = camera_at_time<M>(sc.camera, t);
        Ray ray = camera_get_ray<M>(cam, x, y, wavelength, rng);
        out[i].wavelength = wavelength;
        out[i].x = x;
        out[i].y = y;
        out[i].probability = render_ray<M>(sc, ray, rng, ct, count_tests);
        ct.photons++;
    }
}

// -------------------------------------------------------------------- plot
const float CIE[81][3] = {
#include "cie1931_data.inc"
};

inline V3 tristimulus(float wavelength) {                                       // cie1931.rs:20-48
    float indexf = (wavelength - 380.0f) / 5.0f;
    long index = (long)std::floor(indexf);
    float remainder = indexf - (float)index;
    if (index < -1 || index > 80) return {0.0f, 0.0f, 0.0f};
    if (index == -1) return {CIE[0][0] * remainder, CIE[0][1] * remainder, CIE[0][2] * remainder};
    if (index == 80)
        return {CIE[80][0] * (1.0f - remainder), CIE[80][1] * (1.0f - remainder),
                CIE[80][2] * (1.0f - remainder)};
    long i = index;
    return {CIE[i][0] * (1.0f - remainder) + CIE[i + 1][0] * remainder,
            CIE[i][1] * (1.0f - remainder) + CIE[i + 1][1] * remainder,
            CIE[i][2] * (1.0f - remainder) + CIE[i + 1][2] * remainder};
}

inline void plot_pixel(float *buffer, uint32_t width, uint32_t height, float aspect, float x,
                       float y, V3 cie) {
    // plot_unit.rs:56-84
    long w = width, h = height;
    float px = (x * 0.5f + 0.5f) * ((float)w - 1.0f);
    float py = (y * aspect * 0.5f + 0.5f) * ((float)h - 1.0f);
    auto clampi = [](long v, long hi) { return v < 0 ? 0 : (v > hi ? hi : v); };
    long px1 = clampi((long)std::floor(px), w - 1);
    long px2 = clampi((long)std::ceil(px), w - 1);
    long py1 = clampi((long)std::floor(py), h - 1);
    long py2 = clampi((long)std::ceil(py), h - 1);
    float cx = px - (float)px1;
    float cy = py - (float)py1;
    float c11 = (1.0f - cx) * (1.0f - cy);
    float c12 = (1.0f - cx) * cy;
    float c21 = cx * (1.0f - cy);
    float c22 = cx * cy;
    auto add = [&](long yy, long xx, float c) {
        float *p = buffer + 3 * (yy * w + xx);
        p[0] = p[0] + cie.x * c;
        p[1] = p[1] + cie.y * c;
        p[2] = p[2] + cie.z * c;
    };
    add(py1, px1, c11);
    add(py1, px2, c21);
    add(py2, px1, c12);
    add(py2, px2, c22);
}

void plot(float *buffer, uint32_t width, uint32_t height, const rl_mapped_photon *photons,
          uint64_t n) {
    // plot_unit.rs:87-95
    float aspect = (float)width / (float)height;
    for (uint64_t i = 0; i < n; i++) {
        V3 cie = tristimulus(photons[i].wavelength);
        plot_pixel(buffer, width, height, aspect, photons[i].x, photons[i].y,
                   cie * photons[i].probability);
    }
}

// ------------------------------------------------------------------ gather
void gather_accumulate(float *acc, float *comp, const float *px, uint64_t n_floats) {
    // gather_unit.rs:49-64, per component
    for (uint64_t i = 0; i < n_floats; i++) {
        float extra = px[i] - comp[i];
        float sum = acc[i] + extra;
        comp[i] = (sum - acc[i]) - extra;
        acc[i] = sum;
    }
}

// ----------------------------------------------------------------- tonemap
template <class M>
inline float gamma_correct(float f) {                                           // srgb.rs:20-26
    if (f <= 0.0031308f) return 12.92f * f;
    return 1.055f * M::pow(f, 1.0f / 2.4f) - 0.055f;
}

inline float clamp01(float x) {                                                 // tonemap_unit.rs:34-38
    if (x < 0.0f) return 0.0f;
    if (1.0f < x) return 1.0f;
    return x;
}

inline uint8_t to_u8(float v) {  // Rust `as u8`: saturating, NaN -> 0
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}

float find_exposure(const float *xyz, uint32_t width, uint32_t height) {
    // tonemap_unit.rs:55-69: two sequential f32 folds
    float n = (float)(width * height);
    uint64_t count = (uint64_t)width * height;
    float sum = 0.0f, sqr = 0.0f;
    for (uint64_t i = 0; i < count; i++) sum += xyz[3 * i + 1];
    float mean = sum / n;
    for (uint64_t i = 0; i < count; i++) sqr += xyz[3 * i + 1] * xyz[3 * i + 1];
    float sqr_mean = sqr / n;
    float variance = sqr_mean - mean * mean;
    return mean + std::sqrt(variance);
}

template <class M>
void tonemap(const float *xyz, uint32_t width, uint32_t height, float max_intensity,
             uint8_t *rgb) {
    // tonemap_unit.rs:73-100 + srgb.rs:29-41
    float ln_4 = M::ln(4.0f);
    uint64_t count = (uint64_t)width * height;
    for (uint64_t i = 0; i < count; i++) {
        float cx = M::ln(xyz[3 * i + 0] / max_intensity + 1.0f) / ln_4;
        float cy = M::ln(xyz[3 * i + 1] / max_intensity + 1.0f) / ln_4;
        float cz = M::ln(xyz[3 * i + 2] / max_intensity + 1.0f) / ln_4;
        float r = 3.2406f * cx - 1.5372f * cy - 0.4986f * cz;
        float g = -0.9689f * cx + 1.8758f * cy + 0.0415f * cz;
        float b = 0.0557f * cx - 0.2040f * cy + 1.0570f * cz;
        rgb[3 * i + 0] = to_u8(clamp01(gamma_correct<M>(r)) * 255.0f);
        rgb[3 * i + 1] = to_u8(clamp01(gamma_correct<M>(g)) * 255.0f);
        rgb[3 * i + 2] = to_u8(clamp01(gamma_correct<M>(b)) * 255.0f);
    }
}

// -------------------------------------------------------------- validation
bool is_volume(const SceneView &sc, uint32_t node, int depth) {
    if (node >= sc.n_surfaces || depth > 16) return false;
    const rl_surface &s = sc.surfaces[node];
    if (s.kind == RL_SURFACE_HALFSPACE || s.kind == RL_SURFACE_SPHERE) return true;    // the two Volume leaves
    if (s.kind == RL_SURFACE_COMPOUND)
        return is_volume(sc, s.child[0], depth + 1) && is_volume(sc, s.child[1], depth + 1);
    return false;
}

int make_view(const rl_scene_desc *desc, SceneView &sc) {
    if (!desc || (!desc->surfaces && desc->n_surfaces) || (!desc->objects && desc->n_objects))
        return RL_ERR_INVALID;
    sc.surfaces = desc->surfaces;
    sc.n_surfaces = desc->n_surfaces;
    sc.objects = desc->objects;
    sc.n_objects = desc->n_objects;
    sc.camera = desc->camera;
    if (sc.camera.kind != RL_CAMERA_STATIC && sc.camera.kind != RL_CAMERA_ORBIT
        && sc.camera.kind != RL_CAMERA_KEYFRAMES)
        return RL_ERR_INVALID;
    if (sc.camera.kind == RL_CAMERA_KEYFRAMES && (!sc.camera.keyframes || sc.camera.n_keyframes == 0))
        return RL_ERR_INVALID;
    for (uint32_t i = 0; i < sc.n_objects; i++) {
        const rl_object &o = sc.objects[i];
        if (o.surface >= sc.n_surfaces) return RL_ERR_INVALID;
        if (o.material.kind < RL_MATERIAL_BLACKBODY || o.material.kind > RL_MATERIAL_SOAP_BUBBLE)
            return RL_ERR_INVALID;
        const rl_surface &s = sc.surfaces[o.surface];
        if (s.kind < RL_SURFACE_PLANE || s.kind > RL_SURFACE_COMPOUND) return RL_ERR_INVALID;
        if (s.kind == RL_SURFACE_COMPOUND && !is_volume(sc, o.surface, 0))
            return RL_ERR_UNSUPPORTED;
    }
    return RL_OK;
}

void export_counters(const Counters &c, orc_counters *out) {
    if (!out) return;
    out->photons += c.photons;
    out->rays += c.rays;
    out->primitive_tests += c.tests;
    out->emissive_hits += c.hits_emissive;
    if (c.bounces_max > out->max_bounces) out->max_bounces = c.bounces_max;
}

}  // namespace

// ====================================================================== C API
extern "C" {

int orc_trace(const rl_scene_desc *desc, uint64_t seed, uint32_t width, uint32_t height,
              uint64_t first_photon, uint64_t n, int math_mode, int count_tests,
              rl_mapped_photon *out, orc_counters *counters) {
    SceneView sc;
    int rc = make_view(desc, sc);
    if (rc != RL_OK) return rc;
    if (!out || width == 0 || height == 0) return RL_ERR_INVALID;
    float aspect = (float)width / (float)height;  // trace_unit.rs:73
    Counters ct;
    if (math_mode == ORC_MATH_SPEC)
        trace_range<SpecMath>(sc, seed, first_photon, n, aspect, out, ct, count_tests != 0);
    else
        trace_range<LibmMath>(sc, seed, first_photon, n, aspect, out, ct, count_tests != 0);
    export_counters(ct, counters);
    return RL_OK;
}

int orc_plot(uint32_t width, uint32_t height, const rl_mapped_photon *photons, uint64_t n,
             float *xyz) {
    if (!xyz || (!photons && n) || width == 0 || height == 0) return RL_ERR_INVALID;
    plot(xyz, width, height, photons, n);
    return RL_OK;
}

int orc_gather_accumulate(float *acc, float *comp, const float *px, uint64_t n_pixels) {
    if (!acc || !comp || !px) return RL_ERR_INVALID;
    gather_accumulate(acc, comp, px, n_pixels * 3);
    return RL_OK;
}

int orc_find_exposure(uint32_t width, uint32_t height, const float *xyz, float *out) {
    if (!xyz || !out) return RL_ERR_INVALID;
    *out = find_exposure(xyz, width, height);
    return RL_OK;
}

int orc_tonemap(uint32_t width, uint32_t height, const float *xyz, int math_mode,
                float exposure_or_nan, uint8_t *rgb) {
    if (!xyz || !rgb) return RL_ERR_INVALID;
    float e = exposure_or_nan;
    if (!(e == e)) e = find_exposure(xyz, width, height);
    if (math_mode == ORC_MATH_SPEC) tonemap<SpecMath>(xyz, width, height, e, rgb);
    else tonemap<LibmMath>(xyz, width, height, e, rgb);
    return RL_OK;
}

int orc_intersect(const rl_scene_desc *desc, const rl_ray *rays, uint64_t n, rl_hit *out) {
    SceneView sc;
    int rc = make_view(desc, sc);
    if (rc != RL_OK) return rc;
    for (uint64_t i = 0; i < n; i++) {
        Ray r{v3(rays[i].origin), v3(rays[i].direction), rays[i].wavelength, rays[i].probability};
        Isect is{};
        int obj = scene_intersect(sc, r, is, nullptr);
        out[i].object = obj;
        if (obj >= 0) {
            out[i].distance = is.distance;
            out[i].position = to_rl(is.position);
            out[i].normal = to_rl(is.normal);
            out[i].tangent = to_rl(is.tangent);
        } else {
            out[i].distance = 0.0f;
            out[i].position = out[i].normal = out[i].tangent = rl_vec3{0, 0, 0};
        }
    }
    return RL_OK;
}

int orc_math(int fn, int math_mode, const float *in, const float *in2, uint64_t n, float *out) {
    bool spec = math_mode == ORC_MATH_SPEC;
    for (uint64_t i = 0; i < n; i++) {
        float x = in[i], s, c;
        switch (fn) {
        case 0: if (spec) SpecMath::sincos(x, s, c); else LibmMath::sincos(x, s, c); out[i] = s; break;
        case 1: if (spec) SpecMath::sincos(x, s, c); else LibmMath::sincos(x, s, c); out[i] = c; break;
        case 2: out[i] = spec ? SpecMath::exp(x) : LibmMath::exp(x); break;
        case 3: out[i] = spec ? SpecMath::acos(x) : LibmMath::acos(x); break;
        case 4:
            out[i] = spec ? (float)boltzmann<SpecMath>((double)x, (double)in2[i])
                          : (float)boltzmann<LibmMath>((double)x, (double)in2[i]);
            break;
        case 5: out[i] = sf10_index_of_refraction(x); break;
        case 6: out[i] = spec ? SpecMath::ln(x) : LibmMath::ln(x); break;
        case 7: out[i] = spec ? SpecMath::pow(x, in2[i]) : LibmMath::pow(x, in2[i]); break;
        case 8: out[i] = spec ? SpecMath::tan(x) : LibmMath::tan(x); break;
        default: return RL_ERR_INVALID;
        }
    }
    return RL_OK;
}

int orc_blackbody_intensity(float temperature, float normalisation, int math_mode,
                            const float *wavelengths, uint64_t n, float *out) {
    rl_material m{RL_MATERIAL_BLACKBODY, temperature, normalisation, 0.0f};
    for (uint64_t i = 0; i < n; i++)
        out[i] = math_mode == ORC_MATH_SPEC ? blackbody_intensity<SpecMath>(m, wavelengths[i])
                                            : blackbody_intensity<LibmMath>(m, wavelengths[i]);
    return RL_OK;
}

int orc_tristimulus(const float *wavelengths, uint64_t n, float *out_xyz) {
    for (uint64_t i = 0; i < n; i++) {
        V3 t = tristimulus(wavelengths[i]);
        out_xyz[3 * i] = t.x; out_xyz[3 * i + 1] = t.y; out_xyz[3 * i + 2] = t.z;
    }
    return RL_OK;
}

int orc_camera_rays(const rl_scene_desc *desc, uint64_t seed, uint32_t width, uint32_t height,
                    uint64_t first_photon, uint64_t n, int math_mode, rl_ray *out_rays,
                    rl_mapped_photon *out_xy) {
    SceneView sc;
    int rc = make_view(desc, sc);
    if (rc != RL_OK) return rc;
    float aspect = (float)width / (float)height;
    for (uint64_t i = 0; i < n; i++) {
        Philox rng(seed, first_photon + i);
        float wavelength = rng.wavelength();
        float x = rng.bi_unit();
        float y = rng.bi_unit() / aspect;
        float t = rng.unit();
        Ray ray;
        if (math_mode == ORC_MATH_SPEC) {
            Camera cam = camera_at_time<SpecMath>(sc.camera, t);
            ray = camera_get_ray<SpecMath>(cam, x, y, wavelength, rng);
        } else {
            Camera cam = camera_at_time<LibmMath>(sc.camera, t);
            ray = camera_get_ray<LibmMath>(cam, x, y, wavelength, rng);
        }
        out_rays[i].origin = to_rl(ray.origin);
        out_rays[i].direction = to_rl(ray.direction);
        out_rays[i].wavelength = ray.wavelength;
        out_rays[i].probability = ray.probability;
        if (out_xy) { out_xy[i].x = x; out_xy[i].y = y; out_xy[i].wavelength = wavelength; out_xy[i].probability = t; }
    }
    return RL_OK;
}

int orc_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
               uint32_t *out4) {
    Philox::block4x32_10(k0, k1, c0, c1, c2, c3, out4);
    return RL_OK;
}

int orc_draws(uint64_t seed, uint64_t photon, uint32_t n, const uint8_t *half_open, float *out) {
    Philox rng(seed, photon);
    for (uint32_t i = 0; i < n; i++) out[i] = half_open && half_open[i] ? rng.half_open() : rng.unit();
    return RL_OK;
}

/*
 * The reference's CPU pipeline shape (app.rs:55-70, :132-151): `threads`
 * workers each trace 524 288-photon batches and plot them into a private
 * buffer; the buffers are then Kahan-gathered.  This is the CPU baseline.
 */
int orc_render_mt(const rl_scene_desc *desc, uint64_t seed, uint32_t width, uint32_t height,
                  uint64_t first_photon, uint64_t n_photons, uint64_t batch, int threads,
                  int math_mode, float *xyz_out, orc_counters *counters,
                  double *seconds_trace_plot) {
    SceneView sc;
    int rc = make_view(desc, sc);
    if (rc != RL_OK) return rc;
    if (threads < 1) threads = 1;
    if (batch == 0) batch = RL_BATCH_PHOTONS;
    uint64_t n_batches = (n_photons + batch - 1) / batch;
    size_t n_floats = (size_t)width * height * 3;
    std::vector<std::vector<float>> buffers(threads);
    std::vector<Counters> cts(threads);
    std::atomic<uint64_t> next{0};
    float aspect = (float)width / (float)height;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int ti = 0; ti < threads; ti++) {
        pool.emplace_back([&, ti]() {
            buffers[ti].assign(n_floats, 0.0f);
            std::vector<rl_mapped_photon> photons(batch);
            for (;;) {
                uint64_t b = next.fetch_add(1);
                if (b >= n_batches) break;
                uint64_t lo = b * batch;
                uint64_t cnt = std::min<uint64_t>(batch, n_photons - lo);
                if (math_mode == ORC_MATH_SPEC)
                    trace_range<SpecMath>(sc, seed, first_photon + lo, cnt, aspect, photons.data(),
                                          cts[ti], false);
                else
                    trace_range<LibmMath>(sc, seed, first_photon + lo, cnt, aspect, photons.data(),
                                          cts[ti], false);
                plot(buffers[ti].data(), width, height, photons.data(), cnt);
            }
        });
    }
    for (auto &t : pool) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (seconds_trace_plot) *seconds_trace_plot = std::chrono::duration<double>(t1 - t0).count();
    if (xyz_out) {
        std::vector<float> comp(n_floats, 0.0f);
        std::memset(xyz_out, 0, n_floats * sizeof(float));
        for (int ti = 0; ti < threads; ti++)
            gather_accumulate(xyz_out, comp.data(), buffers[ti].data(), n_floats);
    }
    for (int ti = 0; ti < threads; ti++) export_counters(cts[ti], counters);
    return RL_OK;
}

/* Dump the first `cap` rays (Scene::intersect inputs) of photons [first, first+n): analysis aid. */
int orc_dump_rays(const rl_scene_desc *desc, uint64_t seed, uint32_t width, uint32_t height,
                  uint64_t first_photon, uint64_t n, uint64_t cap, rl_ray *out, uint64_t *n_out) {
    SceneView sc;
    int rc = make_view(desc, sc);
    if (rc != RL_OK) return rc;
    float aspect = (float)width / (float)height;
    uint64_t k = 0;
    for (uint64_t i = 0; i < n && k < cap; i++) {
        Philox rng(seed, first_photon + i);
        float wavelength = rng.wavelength();
        float x = rng.bi_unit();
        float y = rng.bi_unit() / aspect;
        float t = rng.unit();
        Camera cam = camera_at_time<SpecMath>(sc.camera, t);
        Ray ray = camera_get_ray<SpecMath>(cam, x, y, wavelength, rng);
        float continue_chance = 1.0f, intensity = 1.0f;
        for (uint32_t bounce = 0;; bounce++) {
            rng.begin_bounce(bounce);
            if (k < cap) {
                out[k].origin = to_rl(ray.origin); out[k].direction = to_rl(ray.direction);
                out[k].wavelength = ray.wavelength; out[k].probability = intensity; k++;
            }
            Isect is;
            int obj = scene_intersect(sc, ray, is, nullptr);
            if (obj < 0) break;
            const rl_material &mat = sc.objects[obj].material;
            if (mat.kind == RL_MATERIAL_BLACKBODY) break;
            ray = material_new_ray<SpecMath>(mat, ray, is, rng);
            intensity = intensity * ray.probability;
            ray.origin = ray.origin + ray.direction * 0.00001f;
            continue_chance = continue_chance * 0.96f;
            if (rng.unit() * 0.85f > continue_chance * (1.0f - SpecMath::exp(intensity * -20.0f))) break;
        }
    }
    *n_out = k;
    return RL_OK;
}

int orc_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
