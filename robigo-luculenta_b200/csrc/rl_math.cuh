// rl_math.cuh -- f32 vector algebra and the specified math of the path.
//
// Everything here is IEEE + - * / sqrt plus *explicit* fused multiply-adds, so
// that the device result of a path is a pure function of (scene, seed, photon
// id): the translation unit is compiled with -fmad=false (device) and
// -ffp-contract=off (host), the only FMAs are the ones spelled fmaf()/fma().
// The evaluation order of every expression follows the reference (cited per
// function); the transcendental functions the reference takes from libm
// (f32::sin, cos, tan, exp, acos, ln, powf, f64::exp) are replaced by the
// polynomial versions specified in DESIGN.md "Specified math".
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RL_HD __host__ __device__ __forceinline__
#else
#define RL_HD inline
#endif

namespace rl {

RL_HD float bits_to_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
RL_HD uint32_t f32_to_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
RL_HD double bits_to_f64(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double f; memcpy(&f, &u, 8); return f;
#endif
}

#define RL_PI 3.14159274f  /* std::f32::consts::PI */

// ------------------------------------------------------------ specified math
// sin and cos together: k = rint(x * 2/pi); r = x - k*pi/2 in two fused steps;
// Cephes single-precision minimax polynomials on [-pi/4, pi/4].
RL_HD void spec_sincos(float x, float &s, float &c) {
    const float k = rintf(x * 0.636619747f);
    float r = fmaf(k, -1.57079637f, x);
    r = fmaf(k, 4.37113883e-8f, r);
    const float z = r * r;
    float ps = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    const float sn = fmaf(ps * z, r, r);
    float pc = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    const float cs = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
    const int q = (int)k;
    const float s0 = (q & 1) ? cs : sn;
    const float c0 = (q & 1) ? sn : cs;
    s = (q & 2) ? -s0 : s0;
    c = ((q + 1) & 2) ? -c0 : c0;
}

RL_HD float spec_tan(float x) {
    float s, c;
    spec_sincos(x, s, c);
    return s / c;
}

RL_HD float spec_exp(float x) {
    if (!(x == x)) return x;
    if (x < -104.0f) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    const float k = rintf(x * 1.44269502f);
    float r = fmaf(k, -0.693359375f, x);
    r = fmaf(k, 2.12194442e-4f, r);
    const float z = r * r;
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    const float y = fmaf(p, z, r) + 1.0f;
    const int e = (int)k;
    if (e >= -126) return y * bits_to_f32((uint32_t)(e + 127) << 23);
    return (y * bits_to_f32((uint32_t)(e + 227) << 23)) * bits_to_f32(27u << 23);
}

RL_HD float spec_acos(float x) {
    const float a = fabsf(x);
    const bool big = a > 0.5f;
    float zz, w;
    if (big) { zz = (1.0f - a) * 0.5f; w = sqrtf(zz); }
    else { zz = a * a; w = a; }
    float p = 4.2163199048e-2f;
    p = fmaf(p, zz, 2.4181311049e-2f);
    p = fmaf(p, zz, 4.5470025998e-2f);
    p = fmaf(p, zz, 7.4953002686e-2f);
    p = fmaf(p, zz, 1.6666752422e-1f);
    const float as = fmaf(p * zz, w, w);
    if (big) { const float t = as + as; return x < 0.0f ? 3.14159274f - t : t; }
    return x < 0.0f ? 1.57079637f + as : 1.57079637f - as;
}

RL_HD float spec_ln(float x) {
    if (!(x > 0.0f)) return x == 0.0f ? -INFINITY : NAN;
    if (x == INFINITY) return x;
    uint32_t u = f32_to_bits(x);
    int e = 0;
    if (u < 0x00800000u) { x *= 8388608.0f; u = f32_to_bits(x); e = -23; }
    e += (int)(u >> 23) - 126;
    float m = bits_to_f32((u & 0x007fffffu) | 0x3f000000u);
    if (m < 0.707106769f) { e -= 1; m = m + m; }
    const float t = m - 1.0f;
    const float z = t * t;
    float p = 7.0376836292e-2f;
    p = fmaf(p, t, -1.1514610310e-1f);
    p = fmaf(p, t, 1.1676998740e-1f);
    p = fmaf(p, t, -1.2420140846e-1f);
    p = fmaf(p, t, 1.4249322787e-1f);
    p = fmaf(p, t, -1.6668057665e-1f);
    p = fmaf(p, t, 2.0000714765e-1f);
    p = fmaf(p, t, -2.4999993993e-1f);
    p = fmaf(p, t, 3.3333331174e-1f);
    float y = (t * z) * p;
    const float ef = (float)e;
    y = fmaf(ef, -2.12194440e-4f, y);
    y = fmaf(-0.5f, z, y);
    const float r = t + y;
    return fmaf(ef, 0.693359375f, r);
}

RL_HD float spec_pow(float x, float y) { return spec_exp(y * spec_ln(x)); }

RL_HD double spec_exp64(double x) {
    if (!(x == x)) return x;
    if (x > 709.0) return (double)INFINITY;
    if (x < -708.0) return 0.0;
    const double k = rint(x * 1.4426950408889634);
    double r = fma(k, -0.6931471803691238, x);
    r = fma(k, -1.9082149292705877e-10, r);
    double p = 1.6059043836821613e-10;
    p = fma(p, r, 2.08767569878681e-09);
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 0.0001984126984126984);
    p = fma(p, r, 0.001388888888888889);
    p = fma(p, r, 0.008333333333333333);
    p = fma(p, r, 0.041666666666666664);
    p = fma(p, r, 0.16666666666666666);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const long long e = (long long)k;
    return p * bits_to_f64((uint64_t)(e + 1023) << 52);
}

// ------------------------------------------------------------------ vectors
struct V3 { float x, y, z; };

RL_HD V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
RL_HD V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }  // vector3.rs:96-106
RL_HD V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }  // vector3.rs:108-118
RL_HD V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }                       // vector3.rs:120-130
RL_HD V3 operator*(V3 a, float f) { return mk(a.x * f, a.y * f, a.z * f); }     // vector3.rs:132-142
RL_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }       // vector3.rs:35-37
RL_HD V3 cross(V3 a, V3 b) {                                                    // vector3.rs:27-33
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
RL_HD float magnitude_squared(V3 a) { return dot(a, a); }
RL_HD V3 normalise(V3 a) {                                                      // vector3.rs:56-67
    const float m = sqrtf(magnitude_squared(a));
    if (m == 0.0f) return a;
    return mk(a.x / m, a.y / m, a.z / m);
}
RL_HD V3 rotate_towards(V3 v, V3 n) {                                           // vector3.rs:69-83
    if (n.z > 0.9999f) return v;
    if (n.z < -0.9999f) return mk(v.x, v.y, -v.z);
    // cross((0,0,1), n) = (-n.y, n.x, 0) up to the sign of zero, which no later
    // operation observes; written out in full to keep the reference's roundings.
    const V3 a1 = normalise(cross(mk(0.0f, 0.0f, 1.0f), n));
    const V3 a2 = normalise(cross(a1, n));
    return a1 * v.x + a2 * v.y + n * v.z;
}
RL_HD V3 reflect(V3 v, V3 n) { return v - n * 2.0f * dot(n, v); }               // vector3.rs:91-93

struct Quat { float x, y, z, w; };
RL_HD Quat mkq(float x, float y, float z, float w) { Quat q; q.x = x; q.y = y; q.z = z; q.w = w; return q; }
RL_HD Quat conjugate(Quat q) { return mkq(-q.x, -q.y, -q.z, q.w); }             // quaternion.rs:47-49
RL_HD Quat operator*(Quat a, Quat b) {                                          // quaternion.rs:100-110
    return mkq(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
               a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
               a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w,
               a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
RL_HD Quat rotation(float x, float y, float z, float angle) {                   // quaternion.rs:36-45
    float s, c;
    spec_sincos(angle * 0.5f, s, c);
    return mkq(s * x, s * y, s * z, c);
}
RL_HD V3 rotate(V3 v, Quat q) {                                                 // vector3.rs:85-89
    const Quat r = q * mkq(v.x, v.y, v.z, 0.0f) * conjugate(q);
    return mk(r.x, r.y, r.z);
}

// --------------------------------------------------------- material physics
// material.rs:61-74 with constants.rs:19-23
RL_HD double boltzmann(double wavelength, double temperature) {
    const double h = 6.62606957e-34;
    const double k = 1.3806488e-23;
    const double c = 299792458.0;
    const double f = c / (wavelength * 1.0e-9);
    return (2.0 * h * f * f * f) / (c * c * (spec_exp64(h * f / (k * temperature)) - 1.0));
}

// material.rs:203-213
RL_HD float sf10_index_of_refraction(float wavelength) {
    const double w2 = (double)(wavelength * wavelength * 1.0e-6f);
    return (float)sqrt(1.0 + 1.737596950 * w2 / (w2 - 0.0131887070)
                       + 0.313747346 * w2 / (w2 - 0.0623068142)
                       + 1.898781010 * w2 / (w2 - 155.23629000));
}

}  // namespace rl
