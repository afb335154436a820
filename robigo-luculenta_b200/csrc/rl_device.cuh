// rl_device.cuh -- device-side functions of the path: RNG draws, camera ray,
// Scene::intersect, materials, splat.  Included by rl_kernels.cu only.
//
// Layout in HBM / shared memory.  rl_scene_create flattens the descriptor into
// one "primitive blob" of 16-byte records, grouped by primitive type so that a
// warp walks each list with uniform (broadcast) shared-memory loads:
//
//   spheres      n_spheres     x float4   {cx, cy, cz, r^2}            (global memory only)
//   sphere_k     n_spheres     x float4   {cx, cy, cz, |c|^2 - r^2}, record of the sphere pre-test
//   clusters     n_clusters    x float4   {mx, my, mz, |m|^2 - R^2}, bounding sphere of a group of spheres
//   cluster_range n_clusters   x uint32   first member | count << 16 (members are contiguous)
//   supers       n_supers      x float4   bounding sphere of clusters [8k, 8k + 8) (scenes of >= 1024 spheres)
//   planes       n_planes      x 2 float4 {n.xyz, kind} {offset.xyz, r^2}
//   paraboloids  n_paraboloids x 3 float4 {offset,0} {normal,0} {focal_point,0}
//   leaves       n_leaves      x 2 float4 {n.xyz, 0} {offset.xyz, 0}   half-spaces of compounds
//   compounds    n_compounds   x 2 float4 {first_leaf, n_leaves, first_op, n_ops} {bound c.xyz, bound r^2 (< 0: unbounded)}
//   ops          n_ops         x uint32   post-order program of the compound trees
//   *_obj                      x uint32   object index of each sphere/plane/paraboloid/compound
//
// The kernel copies the blob -- all but the exact sphere records and their
// object indices, which sit at its end -- into shared memory once per CTA.  Per-object
// material records {kind, p0, p1, p2} stay in global memory (one read per
// bounce, L1-resident).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rl_b200.h"
#include "rl_math.cuh"
#include "rl_scene_layout.h"

namespace rl {

// Dynamic shared memory of every kernel that traces:
//   [ PrimTables header | primitive blob | scratch: ray table, candidate queues, pair lists, counters ]
extern __shared__ float4 rl_smem[];

// Views into the blob once it sits in shared memory, as offsets from rl_smem in
// float4 units (32-bit shared-memory addressing: generic pointers kept in
// shared memory would turn every table read into a 64-bit generic load).  The
// struct lives at the start of shared memory so that every device function
// reaches the tables without carrying them in registers.
struct PrimTables {
    // exact sphere records and their object indices are read for the few candidates only:
    // they stay in global memory (L1/L2-resident), which halves the shared memory of big scenes
    const float4 *spheres;
    const uint32_t *sphere_obj;
    const float4 *sphere_k_global;   // the pre-test records when they are too many for shared memory, else null
    uint32_t sphere_k, clusters, cluster_range, supers, planes, paraboloids, leaves, compounds, ops;
    uint32_t body_bounds, body_always;
    uint32_t plane_obj, paraboloid_obj, compound_obj;
    uint32_t scratch;         // per-block scratch behind the blob (see Scratch)
    uint32_t n_spheres, n_clusters, n_supers, n_planes, n_paraboloids, n_compounds;
    uint32_t sphere_leaves;   // some compound has a sphere leaf
    float sphere_cmax2, cluster_rmax, super_rmax, leaf_off_max, body_rmax;
};

#define RL_TABLES_VEC4 ((sizeof(PrimTables) + 15) / 16)
#define RL_CAND_SLOTS 8        // queued sphere candidates per lane
#define RL_COMPOUND_SLOTS 4    // body results per lane, and body tasks per thread of the block list
#define RL_PAIR_CAP 768        // (lane, cluster) or (lane, body) pairs per warp and round
// three-level scan: the (lane, cluster) pairs in flight sit at the end of the pair list; a step of the
// group level runs while at most RL_PAIR2_FILL are waiting and adds at most 32 x 8
#define RL_PAIR2_FILL 64
#define RL_PAIR2_CAP (RL_PAIR2_FILL + 256)
// lanes that share one (lane, cluster) pair in the member tests: 1 = every lane takes a pair of the
// warp's list and walks the cluster's members itself (the list is what balances the lanes; the
// per-pair set-up is then paid once per 32 pairs), 8 = eight lanes split the members of a pair
#ifndef RL_PAIR_LANES
#define RL_PAIR_LANES 1
#endif
#ifndef RL_SLAB_LANES
#define RL_SLAB_LANES 1        // the same choice for the slab test of the (lane, body) pairs
#endif
#define RL_BODIES_PER_ROUND 64  // bodies whose bounds one scan covers (one bit each of a lane's candidate mask)
// scratch bytes per thread: ray table 48; sphere queue 2 per slot; two counters 8; pair list
// 2 * RL_PAIR_CAP / 32; body results 8 per slot.  All of it is private to a warp (its threads'
// columns, its pair list): the warps of a block never exchange anything in Scene::intersect.
#define RL_SCRATCH_BYTES_PER_THREAD (48 + 2 * RL_CAND_SLOTS + 8 + 2 * RL_PAIR_CAP / 32 + 8 * RL_COMPOUND_SLOTS)
static_assert(RL_PAIR_INDEX_BITS + 5 <= 16, "a pair record is a lane (5 bits) and a table index in 16 bits");

__device__ __forceinline__ const PrimTables &tables() {
    return *reinterpret_cast<const PrimTables *>(rl_smem);
}
__device__ __forceinline__ const float4 *sm_vec(uint32_t off) { return rl_smem + off; }
__device__ __forceinline__ const uint32_t *sm_u32(uint32_t off) {
    return reinterpret_cast<const uint32_t *>(rl_smem + off);
}

// Block-wide: copy the blob into shared memory and publish the table views.
__device__ __forceinline__ void setup_tables(const DevScene &sc) {
    const uint32_t base = RL_TABLES_VEC4;
    for (uint32_t i = threadIdx.x; i < sc.smem_vec4; i += blockDim.x) rl_smem[base + i] = sc.blob[i];
    if (threadIdx.x == 0) {
        PrimTables t;
        t.spheres = sc.blob + sc.off_spheres;
        t.sphere_obj = reinterpret_cast<const uint32_t *>(sc.blob + sc.off_sphere_obj);
        t.sphere_k = base + sc.off_sphere_k;
        t.sphere_k_global = sc.sphere_k_global ? sc.blob + sc.off_sphere_k : nullptr;
        t.clusters = base + sc.off_clusters;
        t.cluster_range = base + sc.off_cluster_range;
        t.supers = base + sc.off_supers;
        t.planes = base + sc.off_planes;
        t.paraboloids = base + sc.off_paraboloids;
        t.leaves = base + sc.off_leaves;
        t.compounds = base + sc.off_compounds;
        t.ops = base + sc.off_ops;
        t.body_bounds = base + sc.off_body_bounds;
        t.body_always = base + sc.off_body_always;
        t.plane_obj = base + sc.off_plane_obj;
        t.paraboloid_obj = base + sc.off_paraboloid_obj;
        t.compound_obj = base + sc.off_compound_obj;
        t.scratch = base + sc.smem_vec4;
        t.n_spheres = sc.n_spheres;
        t.n_clusters = sc.n_clusters;
        t.n_supers = sc.n_supers;
        t.n_planes = sc.n_planes;
        t.n_paraboloids = sc.n_paraboloids;
        t.n_compounds = sc.n_compounds;
        t.sphere_leaves = sc.sphere_leaves;
        t.sphere_cmax2 = sc.sphere_cmax2;
        t.cluster_rmax = sc.cluster_rmax;
        t.super_rmax = sc.super_rmax;
        t.leaf_off_max = sc.leaf_off_max;
        t.body_rmax = sc.body_rmax;
        *reinterpret_cast<PrimTables *>(rl_smem) = t;
    }
    // zero the scratch counters (same layout as in intersect_scene)
    {
        float4 *ray_tab = rl_smem + base + sc.smem_vec4;
        uint16_t *sq_base = reinterpret_cast<uint16_t *>(ray_tab + 3 * blockDim.x);
        uint32_t *counters = reinterpret_cast<uint32_t *>(sq_base + RL_CAND_SLOTS * blockDim.x);
        for (uint32_t k = threadIdx.x; k < 2 * blockDim.x; k += blockDim.x) counters[k] = 0u;
    }
    __syncthreads();
}

// A scene without compound bodies never touches the body results, the last area of the
// scratch: it does not pay for them (4096 spheres keep 768 threads).
__host__ __device__ __forceinline__ uint32_t scratch_bytes_per_thread(uint32_t n_compounds) {
    return RL_SCRATCH_BYTES_PER_THREAD - (n_compounds ? 0u : 8u * RL_COMPOUND_SLOTS);
}

// Shared memory a tracing kernel needs with `threads` threads per block.
inline size_t tracing_smem_bytes(const DevScene &sc, int threads) {
    return (RL_TABLES_VEC4 + (size_t)sc.smem_vec4) * sizeof(float4)
           + (size_t)scratch_bytes_per_thread(sc.n_compounds) * threads;
}

// ---------------------------------------------------------------------- RNG
// Philox4x32-10, counter (photon_lo, photon_hi, block, 0), key (seed_lo,
// seed_hi).  Stands in for rand::random (monte_carlo.rs:22-28).
// One block is computed out of line so that the call sites share it.
static __device__ __noinline__ uint4 philox_block(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                                  uint32_t block) {
    uint32_t x0 = c0, x1 = c1, x2 = block, x3 = 0u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = __umulhi(0xD2511F53u, x0), l0 = 0xD2511F53u * x0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, x2), l1 = 0xCD9E8D57u * x2;
        const uint32_t y0 = h1 ^ x1 ^ k0, y2 = h0 ^ x3 ^ k1;
        x0 = y0; x1 = l1; x2 = y2; x3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(x0, x1, x2, x3);
}

// The stream a path draws from is named by (seed, photon id); the generator itself only
// carries the position in that stream (block, words left) and the unread words, so that a
// path's state stays small: the key is rebuilt from the launch arguments where a draw is made.
struct RngKey { uint64_t seed, photon; };

struct Rng {
    uint32_t block;
    uint32_t b0, b1, b2, b3;
    uint32_t left;

    __device__ __forceinline__ void init() { block = 0; left = 0; b0 = b1 = b2 = b3 = 0; }
    __device__ __forceinline__ uint32_t next_u32(const RngKey &k) {
        if (left == 0) {
            const uint4 v = philox_block((uint32_t)k.seed, (uint32_t)(k.seed >> 32), (uint32_t)k.photon,
                                         (uint32_t)(k.photon >> 32), block);
            b0 = v.x; b1 = v.y; b2 = v.z; b3 = v.w;
            block++;
            left = 4;
        }
        const uint32_t v = b0;
        b0 = b1; b1 = b2; b2 = b3;
        left--;
        return v;
    }
    // Closed01<f32> of rand 0.3.11 (monte_carlo.rs:25-28): n / (2^24 - 1) for the 24-bit n,
    // in [0, 1].  n / (2^24 - 1) = n 2^-24 (1 + 2^-24 + ...) exceeds the exactly representable
    // n 2^-24 by between half and one ulp, so the correctly rounded quotient is its successor:
    // one integer add on the bit pattern instead of an IEEE division (identity checked for
    // all 2^24 values by the CPU test suite, whose checker does the division).
    __device__ __forceinline__ float unit(const RngKey &k) {
        const uint32_t n = next_u32(k) >> 8;
        const float scaled = (float)n * 5.9604644775390625e-8f;
        return n == 0u ? 0.0f : __uint_as_float(__float_as_uint(scaled) + 1u);
    }
    // rand::random::<f32>() (monte_carlo.rs:37): in [0, 1)
    __device__ __forceinline__ float half_open(const RngKey &k) { return (float)(next_u32(k) >> 8) * 5.9604644775390625e-8f; }
    __device__ __forceinline__ float bi_unit(const RngKey &k) { return unit(k) * 2.0f - 1.0f; }             // monte_carlo.rs:31-33
    __device__ __forceinline__ float longitude(const RngKey &k) { return half_open(k) * RL_PI * 2.0f; }     // monte_carlo.rs:36-38
    __device__ __forceinline__ float wavelength(const RngKey &k) { return unit(k) * 400.0f + 380.0f; }      // monte_carlo.rs:41-43
};

// The draws of one bounce: the material's, then the roulette's (at most three), are the first
// words of block 2 + j of the path's stream (j = bounces so far; blocks 0 and 1 belong to the
// draws in front of the first bounce, see Rng).  One block per bounce, known before the ray is
// traced: the trace loop computes it for all its live lanes at one place, with full warps,
// instead of inside whichever material code happens to run out of words.
struct BounceRng {
    uint32_t d0, d1, d2;
    __device__ __forceinline__ void load(const RngKey &k, uint32_t bounce) {
        const uint4 v = philox_block((uint32_t)k.seed, (uint32_t)(k.seed >> 32), (uint32_t)k.photon,
                                     (uint32_t)(k.photon >> 32), 2u + bounce);
        d0 = v.x; d1 = v.y; d2 = v.z;
    }
    __device__ __forceinline__ uint32_t next_u32() {
        const uint32_t v = d0;
        d0 = d1; d1 = d2;
        return v;
    }
    __device__ __forceinline__ float unit() {                       // Closed01<f32>, see Rng::unit
        const uint32_t n = next_u32() >> 8;
        const float scaled = (float)n * 5.9604644775390625e-8f;
        return n == 0u ? 0.0f : __uint_as_float(__float_as_uint(scaled) + 1u);
    }
    __device__ __forceinline__ float half_open() { return (float)(next_u32() >> 8) * 5.9604644775390625e-8f; }
    __device__ __forceinline__ float longitude() { return half_open() * RL_PI * 2.0f; }     // monte_carlo.rs:36-38
};

// sin and cos through one shared out-of-line copy (eight call sites).
static __device__ __noinline__ float2 sincos_call(float x) {
    float s, c;
    spec_sincos(x, s, c);
    return make_float2(s, c);
}
__device__ __forceinline__ Quat rotation_dev(float x, float y, float z, float angle) {   // quaternion.rs:36-45
    const float2 sc = sincos_call(angle * 0.5f);
    return mkq(sc.x * x, sc.x * y, sc.x * z, sc.y);
}

// Vector3::normalise (vector3.rs:56-67) through one shared out-of-line copy:
// a square root and three IEEE divisions, nine call sites.
static __device__ __noinline__ float3 normalise_call(float x, float y, float z) {
    const V3 n = normalise(mk(x, y, z));
    return make_float3(n.x, n.y, n.z);
}
__device__ __forceinline__ V3 normalise_dev(V3 a) {
    const float3 n = normalise_call(a.x, a.y, a.z);
    return mk(n.x, n.y, n.z);
}
__device__ __forceinline__ V3 rotate_towards_dev(V3 v, V3 n) {                            // vector3.rs:69-83
    if (n.z > 0.9999f) return v;
    if (n.z < -0.9999f) return mk(v.x, v.y, -v.z);
    const V3 a1 = normalise_dev(cross(mk(0.0f, 0.0f, 1.0f), n));
    const V3 a2 = normalise_dev(cross(a1, n));
    return a1 * v.x + a2 * v.y + n * v.z;
}

struct Ray { V3 origin, direction; float wavelength; };

// A ray that hits nothing and queues nothing: lanes without a live path trace
// it so that the warp-wide votes inside intersect_scene stay convergent.
__device__ __forceinline__ Ray idle_ray() {
    Ray r;
    r.origin = mk(1.0e6f, 1.0e6f, 1.0e6f);
    r.direction = mk(1.0f, 0.0f, 0.0f);
    r.wavelength = 0.0f;
    return r;
}

// ------------------------------------------------------------------- camera
// app.rs:327-357 (make_camera) in closed form, camera.rs:94-108 + :47-90.
__device__ __forceinline__ Ray camera_ray(const DevCamera &cm, float x, float y, float wavelength, float t,
                                          Rng &rng, const RngKey &key) {
    V3 position;
    Quat orientation;
    float focal_distance;
    float depth_of_field = cm.depth_of_field, chromatic_abberation = cm.chromatic_abberation;
    // camera.rs:60: 1 / tan(fov / 2) is a constant of the camera, evaluated once on the host with the
    // same specified functions (rl_api.cu, dev_camera)
    float screen_distance = cm.screen_distance;
    if (cm.kind == RL_CAMERA_STATIC) {
        position = mk(cm.px, cm.py, cm.pz);
        orientation = mkq(cm.qx, cm.qy, cm.qz, cm.qw);
        focal_distance = cm.focal_distance;
    } else if (cm.kind == RL_CAMERA_KEYFRAMES) {
        // the tabulated camera function: frame floor(t n), the last one for t == 1
        const uint32_t n = cm.n_keyframes;
        uint32_t k = (uint32_t)floorf(t * (float)n);
        if (k > n - 1u) k = n - 1u;
        const float4 f0 = __ldg(cm.keyframes + 3u * k), f1 = __ldg(cm.keyframes + 3u * k + 1u),
                     f2 = __ldg(cm.keyframes + 3u * k + 2u);
        position = mk(f0.x, f0.y, f0.z);
        focal_distance = f0.w;
        orientation = mkq(f1.x, f1.y, f1.z, f1.w);
        depth_of_field = f2.x; chromatic_abberation = f2.y; screen_distance = f2.z;
    } else {
        const float phi = RL_PI * (cm.phi_base + cm.phi_rate * t);
        const float alpha = RL_PI * (cm.alpha_base + cm.alpha_rate * t);
        const float distance = cm.distance_base + cm.distance_rate * t;
        const float2 a2 = sincos_call(alpha), p2 = sincos_call(phi);
        const float sa = a2.x, ca = a2.y, sp = p2.x, cp = p2.y;
        position = mk(ca * sp * distance, ca * cp * distance, sa * distance);
        orientation = rotation_dev(0.0f, 0.0f, -1.0f, phi + RL_PI) * rotation_dev(1.0f, 0.0f, 0.0f, -alpha);
        focal_distance = distance * cm.focal_factor;
    }
    const float dof_angle = rng.longitude(key);
    const float dof_radius = rng.unit(key) / depth_of_field;
    const float d = (wavelength - 580.0f) / 200.0f;
    const float chromatic_zoom = 1.0f + d * chromatic_abberation;
    const float xs = x * chromatic_zoom;
    const float ys = y * chromatic_zoom;
    const V3 direction = normalise_dev(mk(xs, screen_distance, -ys));
    const V3 focus_point = direction * (focal_distance / direction.y);
    const float2 dof = sincos_call(dof_angle);
    const float sd = dof.x, cd = dof.y;
    const V3 lens_point = mk(cd * dof_radius, 0.0f, sd * dof_radius);
    Ray r;
    r.origin = position + rotate(lens_point, orientation);
    r.direction = normalise_dev(rotate(focus_point - lens_point, orientation));
    r.wavelength = wavelength;
    return r;
}

// ------------------------------------------------------------- intersection
#define RL_HIT_NONE 0u
#define RL_HIT_SPHERE 1u
#define RL_HIT_PLANE 2u
#define RL_HIT_PARABOLOID 3u
#define RL_HIT_LEAF 4u

struct Hit {
    float t;
    int obj;        // object index, -1 = miss
    uint32_t code;  // (type << 28) | index of the primitive (sphere / plane / paraboloid / leaf)
};

// scene.rs:51 keeps the first object of the list on equal distance; the typed
// lists are walked out of list order, so ties resolve on the object index.
__device__ __forceinline__ void consider(Hit &best, float t, int obj, uint32_t code) {
    if (t < best.t || (t == best.t && obj < best.obj)) {
        best.t = t; best.obj = obj; best.code = code;
    }
}

// geometry.rs:55-71; returns t (> 0) or a negative number for "no hit"
__device__ __forceinline__ float plane_t(V3 n, V3 off, const Ray &ray, float &d) {
    const V3 origin = ray.origin - off;
    d = dot(n, ray.direction);
    if (d == 0.0f) return -1.0f;
    const float t = -dot(n, origin) / d;
    return t <= 0.0f ? -1.0f : t;
}

// geometry.rs:204-240: the distance of Sphere::intersect, or negative.
__device__ __forceinline__ float sphere_t(float4 s, const Ray &ray) {
    const V3 co = mk(s.x, s.y, s.z) - ray.origin;
    const float b = 2.0f * dot(ray.direction, co);
    const float c = magnitude_squared(co) - s.w;
    const float disc = b * b - 4.0f * c;
    if (disc < 0.0f) return -1.0f;
    const float d = sqrtf(disc);
    const float t1 = -0.5f * (-b + d);
    const float t2 = -0.5f * (-b - d);
    // `t2 > 0 && t2 < t1` (geometry.rs:238) cannot hold since d >= 0
    return (t1 > 0.0f && t1 < t2) ? t1 : -1.0f;
}

// geometry.rs:299-341
__device__ __forceinline__ float paraboloid_t(const float4 *p, const Ray &ray) {
    const V3 offset = mk(p[0].x, p[0].y, p[0].z);
    const V3 normal = mk(p[1].x, p[1].y, p[1].z);
    const V3 focal_point = mk(p[2].x, p[2].y, p[2].z);
    const V3 origin = ray.origin - offset;
    const V3 focal_offset = origin - focal_point;
    const float n_dot_d = dot(normal, ray.direction);
    const float n_dot_o = dot(normal, origin);
    const float d_dot_f = dot(ray.direction, focal_offset);
    const float a = n_dot_d * n_dot_d - 1.0f;
    const float b = 2.0f * n_dot_d * n_dot_o - 2.0f * d_dot_f;
    const float c = n_dot_o * n_dot_o - magnitude_squared(focal_offset);
    if (a == 0.0f) {
        const float t1 = -c / b;
        return t1 < 0.0f ? -1.0f : t1;   // NaN (b == 0, c == 0) passes `t1 < 0` like the reference
    }
    const float d = b * b - 4.0f * a * c;
    if (d < 0.0f) return -1.0f;
    const float sqrt_d = sqrtf(d);
    const float p1 = 0.5f * (-b + sqrt_d) / a;
    const float q1 = 0.5f * (-b - sqrt_d) / a;
    if (p1 > 0.0f && (p1 < q1 || q1 < 0.0f)) return p1;
    if (q1 > 0.0f) return q1;
    return -1.0f;
}

// geometry.rs:380-407 for trees whose leaves are the reference's two Volume
// types, half-spaces (geometry.rs:90-128) and spheres (geometry.rs:186-267): the
// post-order program replays the reference's recursion with an explicit stack.
// A leaf is two float4: half-space {n, max(1, |n|)} {offset, 0}; sphere
// {0, 0, 0, 0} {centre, r^2} -- the w of the first record tells them apart, and
// the zero "normal" makes a sphere leaf drop out of the slab test by itself.
// op word: bits 0-1 kind (0 = leaf, 1 = compound); leaf: bits 8.. = leaf index
// relative to the compound's first leaf; compound: lo = bits 8-15, mid = bits
// 16-23, hi = bits 24-31 (children own leaves [lo, mid) and [mid, hi)).
// Volume::lies_inside of a leaf: geometry.rs:124-128 (half-space), :263-267 (sphere)
template <bool SPHERES>
__device__ __forceinline__ bool leaf_contains(float4 n4, float4 o4, V3 pos) {
    if (SPHERES && n4.w == 0.0f) return magnitude_squared(pos - mk(o4.x, o4.y, o4.z)) < o4.w;
    return dot(pos - mk(o4.x, o4.y, o4.z), mk(n4.x, n4.y, n4.z)) < 0.0f;
}

template <bool SPHERES>
static __device__ __forceinline__ float2 compound_call(uint32_t first_leaf, uint32_t first_op, uint32_t n_ops,
                                                    float ox, float oy, float oz, float dx, float dy, float dz) {
    Ray ray;
    ray.origin = mk(ox, oy, oz);
    ray.direction = mk(dx, dy, dz);
    ray.wavelength = 0.0f;
    const PrimTables &tb = tables();
    const float4 *leaves = sm_vec(tb.leaves);
    const uint32_t *ops = sm_u32(tb.ops);
    // evaluation stack of (distance, leaf) pairs in registers, top at index 0; the host
    // rejects programs that need more than RL_MAX_COMPOUND_STACK entries
    float t0 = -1.0f, t1 = -1.0f, t2 = -1.0f, t3 = -1.0f, t4 = -1.0f;
    uint32_t l0 = 0, l1 = 0, l2 = 0, l3 = 0, l4 = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < n_ops; i++) {
        const uint32_t op = ops[first_op + i];
        if ((op & 3u) == 0u) {
            const uint32_t leaf = first_leaf + (op >> 8);
            const float4 n4 = leaves[2 * leaf], o4 = leaves[2 * leaf + 1];
            float d;
            const float t = (SPHERES && n4.w == 0.0f) ? sphere_t(o4, ray)
                                         : plane_t(mk(n4.x, n4.y, n4.z), mk(o4.x, o4.y, o4.z), ray, d);
            t4 = t3; l4 = l3; t3 = t2; l3 = l2; t2 = t1; l2 = l1; t1 = t0; l1 = l0;
            t0 = t; l0 = leaf;
        } else {
            const uint32_t lo = first_leaf + ((op >> 8) & 0xffu);
            const uint32_t mid = first_leaf + ((op >> 16) & 0xffu);
            const uint32_t hi = first_leaf + ((op >> 24) & 0xffu);
            float b2 = t0; const uint32_t k2 = l0;              // surface2's hit
            float b1 = t1; const uint32_t k1 = l1;              // surface1's hit
            if (b1 > 0.0f) {  // surface2.lies_inside(i1.position)
                const V3 pos = ray.origin + ray.direction * b1;
#pragma unroll 1
                for (uint32_t k = mid; k < hi; k++) {
                    if (!leaf_contains<SPHERES>(leaves[2 * k], leaves[2 * k + 1], pos)) { b1 = -1.0f; break; }
                }
            }
            if (b2 > 0.0f) {  // surface1.lies_inside(i2.position)
                const V3 pos = ray.origin + ray.direction * b2;
#pragma unroll 1
                for (uint32_t k = lo; k < mid; k++) {
                    if (!leaf_contains<SPHERES>(leaves[2 * k], leaves[2 * k + 1], pos)) { b2 = -1.0f; break; }
                }
            }
            // both valid: the nearer, ties to surface2 (geometry.rs:391-396); else whichever is valid
            const bool first = b1 > 0.0f && (!(b2 > 0.0f) || b1 < b2);
            t0 = first ? b1 : b2; l0 = first ? k1 : k2;
            t1 = t2; l1 = l2; t2 = t3; l2 = l3; t3 = t4; l3 = l4;
        }
    }
    return make_float2(t0, __uint_as_float(l0));
}

// Compound::intersect for the body described by the record c4 = {first_leaf, n_leaves, first_op,
// n_ops}.  The evaluation sits on the critical path of the whole block (every warp waits at a
// barrier for the threads that evaluate bodies), and the leaf-kind tests cost 4 % of the kernel's
// time on the built-in scene when they are compiled in: a scene without sphere leaves (block-
// uniform flag) runs the copy without them.
__device__ __forceinline__ float compound_t(float4 c4, const Ray &ray, uint32_t &leaf_out) {
    const uint32_t first_leaf = __float_as_uint(c4.x), first_op = __float_as_uint(c4.z), n_ops = __float_as_uint(c4.w);
    const float2 r = tables().sphere_leaves
                         ? compound_call<true>(first_leaf, first_op, n_ops, ray.origin.x, ray.origin.y, ray.origin.z,
                                               ray.direction.x, ray.direction.y, ray.direction.z)
                         : compound_call<false>(first_leaf, first_op, n_ops, ray.origin.x, ray.origin.y, ray.origin.z,
                                                ray.direction.x, ray.direction.y, ray.direction.z);
    leaf_out = __float_as_uint(r.y);
    return r.x;
}

// Plane, Circle, top-level SpacePartitioning and Paraboloid objects: few per
// scene, evaluated exactly for every ray (shared by both intersect variants).
__device__ __forceinline__ void intersect_flat_surfaces(const PrimTables &tb, const Ray &ray, Hit &best) {
    const float4 *planes = sm_vec(tb.planes);
    const uint32_t *plane_obj = sm_u32(tb.plane_obj);
    for (uint32_t k = 0; k < tb.n_planes; k++) {
        const float4 n4 = planes[2 * k], o4 = planes[2 * k + 1];
        float dn;
        const float t = plane_t(mk(n4.x, n4.y, n4.z), mk(o4.x, o4.y, o4.z), ray, dn);
        if (t > 0.0f) {
            bool ok = true;
            if (__float_as_uint(n4.w) == RL_SURFACE_CIRCLE) {  // geometry.rs:169-172
                const V3 pos = ray.origin + ray.direction * t;
                ok = magnitude_squared(pos - mk(o4.x, o4.y, o4.z)) <= o4.w;
            }
            if (ok) consider(best, t, (int)plane_obj[k], (RL_HIT_PLANE << 28) | k);
        }
    }
    const float4 *paraboloids = sm_vec(tb.paraboloids);
    const uint32_t *paraboloid_obj = sm_u32(tb.paraboloid_obj);
    for (uint32_t k = 0; k < tb.n_paraboloids; k++) {
        const float t = paraboloid_t(paraboloids + 3 * k, ray);
        // the a == 0 branch admits t == 0 (geometry.rs:319: only t1 < 0 is rejected)
        if (t >= 0.0f) consider(best, t, (int)paraboloid_obj[k], (RL_HIT_PARABOLOID << 28) | k);
    }
}

// Scene::intersect (scene.rs:39-60): closest hit over all objects, every
// primitive evaluated with the reference's arithmetic.  Kept as the in-kernel
// reference the culled version below is checked against (rl_debug_cull_check).
__device__ __forceinline__ Hit intersect_scene_brute(const Ray &ray) {
    const PrimTables &tb = tables();
    Hit best;
    best.t = 1.0e12f; best.obj = -1; best.code = RL_HIT_NONE;
    const float4 *spheres = tb.spheres;
    const uint32_t *sphere_obj = tb.sphere_obj;
    for (uint32_t i = 0; i < tb.n_spheres; i++) {
        const float t = sphere_t(__ldg(spheres + i), ray);
        if (t > 0.0f) consider(best, t, (int)__ldg(sphere_obj + i), (RL_HIT_SPHERE << 28) | i);
    }
    intersect_flat_surfaces(tb, ray, best);
    const float4 *compounds = sm_vec(tb.compounds);
    const uint32_t *compound_obj = sm_u32(tb.compound_obj);
    for (uint32_t i = 0; i < tb.n_compounds; i++) {
        const float4 c4 = compounds[2 * i];
        uint32_t leaf;
        const float t = compound_t(c4, ray, leaf);
        if (t > 0.0f) consider(best, t, (int)compound_obj[i], (RL_HIT_LEAF << 28) | leaf);
    }
    return best;
}

// Uniform scan of `count` bounding-sphere records {m, |m|^2 - R^2} (count a multiple of 8, at most
// 64; tables are padded with records no ray selects): bit k of the result is set iff the sphere
// pre-test keeps record k for this lane's ray -- 8 fused operations, two compares and one
// predicated OR per record, the candidates of a lane stay in two registers.
// B'^2 - C' >= thr is evaluated as B'^2 - (C' - |o|^2) >= thr + |o|^2: the |o|^2 term is the same for
// every record of a ray, so it is added to the threshold once instead of to every C' (one rounding
// of a term of the same magnitude either way: the error bound of the pre-test is unchanged).
__device__ __forceinline__ uint64_t scan_bounds(const float4 *tab, uint32_t count, V3 d, float ndo, float m2ox,
                                                float m2oy, float m2oz, float oo, float thr, float bthr) {
    const float thr_oo = thr + oo;
    uint64_t mask = 0ull;
#pragma unroll 1
    for (uint32_t g = 0; g < count; g += 8) {
        uint32_t m8 = 0u;
#pragma unroll
        for (uint32_t j = 0; j < 8; j++) {
            const float4 s = tab[g + j];
            const float b = fmaf(d.x, s.x, fmaf(d.y, s.y, fmaf(d.z, s.z, ndo)));
            const float c = fmaf(m2ox, s.x, fmaf(m2oy, s.y, fmaf(m2oz, s.z, s.w)));   // C' - |o|^2
            const float disc = fmaf(b, b, -c);
            if (disc >= thr_oo && b >= bthr) m8 |= 1u << j;
        }
        mask |= (uint64_t)m8 << g;
    }
    return mask;
}

// Lists the warp's candidates as (lane, base + bit) records in `pairs` (RL_PAIR_CAP entries) and
// clears the listed bits of `todo`; returns the number of records (warp-uniform).  When the
// warp's candidates do not fit, only a window of 24 bit positions is listed (at most 24 x 32
// records; cap / 32 positions for a smaller cap) and the caller comes back for the rest.
__device__ __forceinline__ uint32_t emit_pairs(uint64_t &todo, uint32_t base, uint16_t *pairs, uint32_t lane,
                                               uint32_t cap = RL_PAIR_CAP) {
    uint64_t take = todo;
    uint32_t cnt = (uint32_t)__popcll(take);
    uint32_t total = __reduce_add_sync(0xffffffffu, cnt);
    if (total > cap) {                                              // warp-uniform
        const uint32_t low = __reduce_min_sync(0xffffffffu, take ? (uint32_t)__ffsll((long long)take) - 1u : 64u);
        take &= ((1ull << (cap >> 5)) - 1ull) << low;
        cnt = (uint32_t)__popcll(take);
        total = __reduce_add_sync(0xffffffffu, cnt);
    }
    todo &= ~take;
    uint32_t incl = cnt;
#pragma unroll
    for (uint32_t sh = 1; sh < 32; sh <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, sh);
        if (lane >= sh) incl += v;
    }
    uint16_t *dst = pairs + (incl - cnt);
    const uint32_t tag = (lane << RL_PAIR_INDEX_BITS) | base;       // base is a multiple of 64: OR = add
#pragma unroll 1
    while (take) {
        const uint32_t bit = (uint32_t)__ffsll((long long)take) - 1u;
        take &= take - 1ull;
        *dst++ = (uint16_t)(tag | bit);
    }
    __syncwarp();
    return total;
}

// Compound::intersect (geometry.rs:380-401) of a tree of at most eight half-spaces, decided by
// eight lanes at once (lane `sub` of its group of eight holds leaf `sub`), without walking the tree.
//
// What the recursion returns.  A leaf's own hit (plane_t > 0) moves up the tree; at every
// compound node each child's surviving hit is first FILTERED -- kept iff its position lies
// inside every leaf of the other child (lies_inside, geometry.rs:403-407) -- and then the nearer
// of the two survivors is kept (ties: the second child's).  So (a) a hit that reaches the root
// has passed the containment test of every other leaf; (b) a hit that fails the test of some
// leaf m is dropped at the lowest common ancestor with m at the latest, BEFORE it is compared
// with the survivor of m's side.  Let "sound" mean: valid and inside every other leaf (the
// reference's own f32 predicates, evaluated here on the same operands).  Then:
//   - no sound leaf: nothing reaches the root -- a miss;
//   - otherwise let L be the nearest sound leaf.  If every other valid leaf k with t_k <= t_L
//     fails the test of L's half-space, k is dropped no later than where it would meet L's
//     survivor, so L wins every comparison on its way up (the survivors it is compared with are
//     farther) and passes every filter (it is sound): the recursion returns L's hit.
//   - anything else (a leaf at or before t_L that lies inside L's half-space: ties, or hits within
//     rounding of an edge) is not decided here: the caller runs the recursion itself.
// Returns t_L (> 0), -1 for a miss, -2 for "undecided"; leaf_out = L (relative to the body's first
// leaf).  Called by all 32 lanes; `valid` and the body are uniform within a group of eight.
__device__ __forceinline__ float eval_body_exact(const float4 *leaves, uint32_t n_leaves, const Ray &ray, bool use,
                                                 uint32_t sub, uint32_t group_bits, uint32_t &leaf_out) {
    // every lane of the warp runs through here (warp-wide votes and shuffles below); a group
    // whose pair is not evaluated this way has use == false and ignores what it computes
    const uint32_t nl = use ? n_leaves : 1u;
    const bool mine = use && sub < nl;
    const float4 n4 = leaves[2 * (mine ? sub : 0u)], o4 = leaves[2 * (mine ? sub : 0u) + 1];
    float dn;
    const float t = plane_t(mk(n4.x, n4.y, n4.z), mk(o4.x, o4.y, o4.z), ray, dn);
    const bool hit = mine && t > 0.0f;
    const V3 pos = ray.origin + ray.direction * t;              // Intersection::position (geometry.rs:108-110)
    uint32_t inside = 0;                                        // bit j: this leaf's hit lies inside leaf j
#pragma unroll
    for (uint32_t j = 0; j < 8; j++) {
        const float4 nj = leaves[2 * (j < nl ? j : 0u)], oj = leaves[2 * (j < nl ? j : 0u) + 1];
        if (dot(pos - mk(oj.x, oj.y, oj.z), mk(nj.x, nj.y, nj.z)) < 0.0f) inside |= 1u << j;   // geometry.rs:124-128
    }
    const uint32_t all = (1u << nl) - 1u;
    const bool sound = hit && ((inside | (1u << sub)) & all) == all;
    float t_near = sound ? t : 3.0e38f;
#pragma unroll
    for (uint32_t sh = 1; sh < 8; sh <<= 1) t_near = fminf(t_near, __shfl_xor_sync(0xffffffffu, t_near, sh));
    const uint32_t nearest = __ballot_sync(0xffffffffu, sound && t == t_near) & group_bits;
    const uint32_t L = nearest ? (uint32_t)(__ffs(nearest) - 1) & 7u : 0u;
    leaf_out = L;
    // a valid leaf at or before t_L that lies inside L's half-space (this includes a second nearest sound leaf)
    const uint32_t spoils = __ballot_sync(0xffffffffu, hit && sub != L && t <= t_near && ((inside >> L) & 1u)) & group_bits;
    if (nearest == 0u) return -1.0f;                            // no sound leaf
    return spoils ? -2.0f : t_near;
}

// Conservative slab test of a convex body against the ray, with every
// half-space moved outwards by RL_SLAB_INFLATE (evaluated one lane per (lane,
// body) pair inside intersect_scene).  A hit the reference returns lies on one leaf plane
// and passes the f32 containment test of every other leaf (each sits in a
// sibling subtree on the way to the root, geometry.rs:386-388), so it is inside
// the inflated body up to rounding (~1e-5 at the scene's coordinate
// magnitudes); if the inflated body's [enter, exit] interval is empty, ends
// before the origin, or starts beyond the best hit so far, the exact evaluation
// cannot change the result and is skipped.  The f32 evaluation of n.(o - off)
// is off by up to ~4 eps (|o| + |off|) |n|, so the inflation grows with the
// magnitudes involved: RL_SLAB_INFLATE + 20 eps (|o| + max |off|), times
// max(1, |n|) of the leaf (kept in the w lane of its normal record) -- for a ray
// that starts far from the scene (a path leaving a distant plane) the slab test
// keeps everything, which is the safe direction.
#define RL_SLAB_INFLATE 2.0e-3f
#define RL_SLAB_INFLATE_REL 1.2e-6f

// Scene::intersect with result-preserving culls.
//
// Spheres.  The reference accepts a sphere only if fl(b^2 - 4c) >= 0 and
// t1 = (b - sqrt(disc)) / 2 > 0, which needs b > 0 (geometry.rs:204-240).
// With B = d.(c - o) and C = |c - o|^2 - r^2 (so disc / 4 = B^2 - C), the
// pre-test evaluates B and C from per-ray constants in 6 + 2 fused operations,
//     B' = d.c - d.o,     C' = (|c|^2 - r^2) - 2 o.c + |o|^2,
// and keeps the sphere iff B'^2 - C' >= -e1 and B' >= -e2, where e1, e2 bound
// the rounding error of BOTH evaluations (reference and pre-test):
//     |B'^2 - C' - disc_ref/4| <= 44 eps (cmax2 + |o|^2) max(1,|d|^2)   <  e1 = 2^-17 (cmax2 + |o|^2) max(1,|d|^2)
//     |B' - b_ref/2|           <= 12 eps sqrt((cmax2 + |o|^2) |d|^2)    <  e2 = 2^-19 sqrt((cmax2 + |o|^2) |d|^2)
// (eps = 2^-24, cmax2 = max |c|^2 + r^2; derivation in DESIGN.md "Culling").
// Every survivor is then evaluated with the reference's exact arithmetic
// (sphere_t), so the hit distance, the winner and every later rounding are
// unchanged; culling only removes spheres the reference rejects.
//
// The spheres are grouped into clusters with bounding spheres (host side):
// level 1 runs the pre-test against the cluster bounds in a uniform loop, each
// lane queueing its candidate clusters privately; the queues are compacted into
// one (lane, cluster) list per warp; at level 2 every lane takes ONE pair of
// that list and tests the cluster's members with the owner's constants from a
// shared ray table, so that lanes with many candidates do not hold the warp
// back and the per-pair set-up is paid once per 32 pairs; level 3 is the exact
// sphere_t, each lane for the few candidates queued for it.
//
// Compounds.  A bounded convex body can only be hit where the ray passes its
// bounding sphere (inflated on the host well beyond rounding); survivors go
// through the slab test above (one lane per listed pair), and what remains is
// evaluated exactly by eight lanes per body, one leaf each (eval_body_exact).
// Unbounded bodies always pass the bounding test.
//
// Nothing in here crosses a warp: every warp of the block runs through Scene::intersect on its
// own (with the body evaluation shared by the whole block through a task list and two block
// barriers, 22 % of all warp cycles were spent waiting at those barriers, and the bodies -- 22 of
// the built-in scene's 339 objects -- cost a third of the kernel's time).
//
// Must be called by all 32 lanes of a warp together (warp votes and shuffles inside); lanes
// without a live path pass idle_ray() and live = false.  A warp none of whose lanes is live
// skips everything (its rays hit nothing anyway).
//
// GLOBAL_K: the spheres' pre-test records are read from global memory (scenes with more than 6144
// spheres) instead of shared memory.  A template parameter, not a run-time test: the read sits in
// the innermost loop of the member test, where a uniform select costs 2.7 % of the kernel's time
// on the built-in scene.
//
// DEEP: scenes of a thousand spheres or more carry a third level, bounds over groups of eight
// clusters (of eight spheres each).  The uniform scan then runs over the groups; a lane's
// candidate groups are listed as (lane, group) pairs and every lane tests the eight cluster bounds
// of one pair with the owner's constants, exactly as the members of a cluster are tested one level
// down; the surviving (lane, cluster) pairs collect in a short list that is drained 32 pairs at a
// time into the member tests.  4096 random spheres: 64 + ~80 + ~150 pre-tests per ray instead of
// 256 + ~300 with two levels (1842 -> 2500 Mrays/s together with the one-lane-per-pair walks).
template <bool GLOBAL_K, bool DEEP>
__device__ __forceinline__ Hit intersect_scene(const Ray &ray, bool live = true) {
    const PrimTables &tb = tables();
    Hit best;
    best.t = 1.0e12f; best.obj = -1; best.code = RL_HIT_NONE;

    const V3 o = ray.origin, d = ray.direction;
    const float oo = fmaf(o.z, o.z, fmaf(o.y, o.y, o.x * o.x));
    const float dd = fmaf(d.z, d.z, fmaf(d.y, d.y, d.x * d.x));
    const float ndo = -fmaf(d.z, o.z, fmaf(d.y, o.y, d.x * o.x));
    const float m2ox = -2.0f * o.x, m2oy = -2.0f * o.y, m2oz = -2.0f * o.z;
    const float scale = (tb.sphere_cmax2 + oo) * fmaxf(1.0f, dd);
    const float thr = -7.6293945e-6f * scale;                                          // -2^-17 * scale
    const float bthr = -1.9073486e-6f * sqrtf((tb.sphere_cmax2 + oo) * dd) - 1.0e-30f;  // -2^-19 * ...

    // Scratch views (per block, see RL_SCRATCH_BYTES_PER_THREAD): ray table [3 float4 per
    // thread]; sphere queue [slot][thread]; counters [2][thread]; pair list [RL_PAIR_CAP per warp];
    // body results [slot][thread] (distance, code).
    const uint32_t nthreads = blockDim.x, tid = threadIdx.x;
    const uint32_t lane = tid & 31u, wbase = tid & ~31u;
    float4 *ray_tab = rl_smem + tb.scratch;
    uint16_t *sq_base = reinterpret_cast<uint16_t *>(ray_tab + 3 * nthreads);
    uint32_t *sq_cnt = reinterpret_cast<uint32_t *>(sq_base + RL_CAND_SLOTS * nthreads);
    uint16_t *pairs = reinterpret_cast<uint16_t *>(sq_cnt + 2 * nthreads) + (wbase >> 5) * RL_PAIR_CAP;
    float2 *results = reinterpret_cast<float2 *>(reinterpret_cast<uint16_t *>(sq_cnt + 2 * nthreads)
                                                 + (nthreads >> 5) * RL_PAIR_CAP);
    // publish this lane's pre-test constants so that any lane of the block can test for it
    ray_tab[3 * tid + 0] = make_float4(m2ox, m2oy, m2oz, oo);
    ray_tab[3 * tid + 1] = make_float4(d.x, d.y, d.z, ndo);
    uint32_t *res_cnt = sq_cnt + nthreads;                      // body results other lanes computed for this one
    sq_cnt[tid] = 0u;

    const float slack = -2.0f * thr + 2.0f * fabsf(dd - 1.0f) * (tb.sphere_cmax2 + oo);
    const bool warp_live = __any_sync(0xffffffffu, live);       // warp-uniform

    if (warp_live) intersect_flat_surfaces(tb, ray, best);
    const float slab_inflate = fmaf(RL_SLAB_INFLATE_REL, sqrtf(oo) + tb.leaf_off_max, RL_SLAB_INFLATE);
    if (DEEP)                                                   // .zw: the owner's thresholds of the cluster-bound test
        ray_tab[3 * tid + 2] = make_float4(thr + oo, bthr, -(2.0f * tb.cluster_rmax * sqrtf(slack) + 2.0f * slack) + oo,
                                           bthr - sqrtf(dd) * tb.cluster_rmax);
    else
        ray_tab[3 * tid + 2] = make_float4(thr + oo, bthr, best.t,   // .x: threshold + |o|^2 (see scan_bounds); .z: nearest hit so far, bodies beyond it are skipped
                                           slab_inflate);

    // Two-level sphere scan.  Level 1, uniform over the warp: the same pre-test against the bounding
    // sphere {m, R} of each cluster of spheres, with thresholds widened so that a cluster is
    // kept whenever the reference could accept one of its members: a member i the reference
    // accepts has r_i^2 - dist(line, c_i)^2 >= -S/2 with S = 2 e1 + 2 |dd - 1| (cmax2 + |o|^2)
    // (rounding of the reference's discriminant, and its use of B^2 - C for a direction that
    // is only unit to rounding), hence dist(line, m) <= R + sqrt(S/2) and
    // R^2 - dist(line, m)^2 >= -(2 R sqrt(S/2) + S/2); B_cluster >= B_i - |d| R.
    // A lane's candidates are bits of a register (scan_bounds); the warp's (lane, cluster) pairs
    // are then listed in shared memory (emit_pairs).
    auto sphere_phase = [&]() {
        const float4 *spheres = tb.spheres;
        const float4 *sphere_k = sm_vec(tb.sphere_k);
        const float4 *sphere_k_global = tb.sphere_k_global;         // block-uniform
        const float4 *clusters = sm_vec(tb.clusters);
        const uint32_t *cluster_range = sm_u32(tb.cluster_range);
        const uint32_t *sphere_obj = tb.sphere_obj;
        const uint32_t n_clusters = tb.n_clusters;
        const float thr_c = -(2.0f * tb.cluster_rmax * sqrtf(slack) + 2.0f * slack);
        const float bthr_c = bthr - sqrtf(dd) * tb.cluster_rmax;
        if constexpr (DEEP) {
            const float4 *supers = sm_vec(tb.supers);
            const uint32_t n_supers = tb.n_supers;
            const float thr_s = -(2.0f * tb.super_rmax * sqrtf(slack) + 2.0f * slack);
            const float bthr_s = bthr - sqrtf(dd) * tb.super_rmax;
            uint16_t *pairs2 = pairs + (RL_PAIR_CAP - RL_PAIR2_CAP);
#pragma unroll 1
            for (uint32_t base = 0; warp_live && base < n_supers; base += 64) {
                uint64_t todo = scan_bounds(supers + base, min(64u, n_supers - base), d, ndo, m2ox, m2oy, m2oz, oo,
                                            thr_s, bthr_s);
                while (__any_sync(0xffffffffu, todo != 0ull)) {
                    const uint32_t npairs = emit_pairs(todo, base, pairs, lane, RL_PAIR_CAP - RL_PAIR2_CAP);
                    uint32_t pb = 0, n2 = 0;                        // warp-uniform
#pragma unroll 1
                    for (;;) {
                        if (pb < npairs && n2 <= RL_PAIR2_FILL) {
                            // group level: every lane takes one (lane, group) pair of the list and tests
                            // the group's eight cluster bounds with the owner's constants
                            const uint32_t p = pb + lane;
                            uint32_t m8 = 0u, tag = 0u;
                            if (p < npairs) {
                                const uint32_t pair = pairs[p];
                                const uint32_t owner = wbase + (pair >> RL_PAIR_INDEX_BITS);
                                const float4 *cl = clusters + ((pair & RL_PAIR_INDEX_MAX) << 3);
                                tag = (pair & ~RL_PAIR_INDEX_MAX) | ((pair & RL_PAIR_INDEX_MAX) << 3);
                                const float4 ro = ray_tab[3 * owner], rd = ray_tab[3 * owner + 1];
                                const float2 th = *reinterpret_cast<const float2 *>(&ray_tab[3 * owner + 2].z);
                                // groups start at multiples of eight records = 128 bytes: lane l begins
                                // with record l mod 8, so that the 32 reads of a step spread over all banks
#pragma unroll
                                for (uint32_t j = 0; j < 8; j++) {
                                    const uint32_t jj = (j + lane) & 7u;
                                    const float4 s = cl[jj];         // padding records are never kept
                                    const float b = fmaf(rd.x, s.x, fmaf(rd.y, s.y, fmaf(rd.z, s.z, rd.w)));
                                    const float cc = fmaf(ro.x, s.x, fmaf(ro.y, s.y, fmaf(ro.z, s.z, s.w)));
                                    const float disc = fmaf(b, b, -cc);
                                    if (disc >= th.x && b >= th.y) m8 |= 1u << jj;
                                }
                            }
                            const uint32_t cnt = (uint32_t)__popc(m8);
                            uint32_t incl = cnt;
#pragma unroll
                            for (uint32_t sh = 1; sh < 32; sh <<= 1) {
                                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, sh);
                                if (lane >= sh) incl += v;
                            }
                            uint16_t *dst = pairs2 + n2 + (incl - cnt);
#pragma unroll 1
                            while (m8) {
                                const uint32_t bit = (uint32_t)__ffs((int)m8) - 1u;
                                m8 &= m8 - 1u;
                                *dst++ = (uint16_t)(tag | bit);
                            }
                            n2 += __shfl_sync(0xffffffffu, incl, 31);
                            pb += 32u;
                            __syncwarp();
                        } else if (n2 != 0u) {
                            // member level: every lane takes one of the last 32 (lane, cluster) pairs
                            const uint32_t take = n2 < 32u ? n2 : 32u;
                            if (lane < take) {
                                const uint32_t pair = pairs2[n2 - 1u - lane];
                                const uint32_t owner = wbase + (pair >> RL_PAIR_INDEX_BITS);
                                const uint32_t r = cluster_range[pair & RL_PAIR_INDEX_MAX];
                                const uint32_t first = r & 0xffffu, count = r >> 16;
                                const float4 ro = ray_tab[3 * owner], rd = ray_tab[3 * owner + 1];
                                const float2 rt = *reinterpret_cast<const float2 *>(&ray_tab[3 * owner + 2]);
                                // clusters of eight start at multiples of 128 bytes: the lanes walk them
                                // from different members on (see the group level)
                                const uint32_t start = (lane & 7u) < count ? (lane & 7u) : 0u;
#pragma unroll 1
                                for (uint32_t j = 0; j < count; j++) {
                                    uint32_t jj = j + start;
                                    if (jj >= count) jj -= count;
                                    const uint32_t m = first + jj;
                                    const float4 s = GLOBAL_K ? __ldg(sphere_k_global + m) : sphere_k[m];
                                    const float b = fmaf(rd.x, s.x, fmaf(rd.y, s.y, fmaf(rd.z, s.z, rd.w)));
                                    const float c = fmaf(ro.x, s.x, fmaf(ro.y, s.y, fmaf(ro.z, s.z, s.w)));
                                    const float disc = fmaf(b, b, -c);
                                    if (disc >= rt.x && b >= rt.y) {
                                        const uint32_t slot = atomicAdd(&sq_cnt[owner], 1u);
                                        if (slot < RL_CAND_SLOTS) sq_base[slot * nthreads + owner] = (uint16_t)m;
                                    }
                                }
                            }
                            n2 -= take;
                            __syncwarp();
                        } else {
                            break;
                        }
                    }
                    // exact Sphere::intersect for the candidates queued for this lane (as below)
                    const uint32_t cnt = sq_cnt[tid];
                    if (cnt > RL_CAND_SLOTS) {
#pragma unroll 1
                        for (uint32_t k = 0; k < tb.n_spheres; k++) {
                            const float t = sphere_t(__ldg(spheres + k), ray);
                            if (t > 0.0f) consider(best, t, (int)__ldg(sphere_obj + k), (RL_HIT_SPHERE << 28) | k);
                        }
                    } else {
#pragma unroll 1
                        for (uint32_t k = 0; k < cnt; k++) {
                            const uint32_t idx = sq_base[k * nthreads + tid];
                            const float t = sphere_t(__ldg(spheres + idx), ray);
                            if (t > 0.0f) consider(best, t, (int)__ldg(sphere_obj + idx), (RL_HIT_SPHERE << 28) | idx);
                        }
                    }
                    sq_cnt[tid] = 0u;
                    __syncwarp();
                }
            }
        } else {
#pragma unroll 1
        for (uint32_t base = 0; warp_live && base < n_clusters; base += 64) {
            uint64_t todo = scan_bounds(clusters + base, min(64u, n_clusters - base), d, ndo, m2ox, m2oy, m2oz, oo,
                                        thr_c, bthr_c);
            while (__any_sync(0xffffffffu, todo != 0ull)) {
                const uint32_t npairs = emit_pairs(todo, base, pairs, lane);
                // Level 2: every lane takes one (lane, cluster) pair of the list and tests the
                // cluster's members with the owner's constants, so the work of lanes with many
                // candidate clusters is spread over the warp; survivors go to the owner's sphere queue.
#pragma unroll 1
                for (uint32_t pb = 0; pb < npairs; pb += 32 / RL_PAIR_LANES) {
                    const uint32_t p = pb + lane / RL_PAIR_LANES;
                    if (p < npairs) {
                        const uint32_t pair = pairs[p];
                        const uint32_t owner = wbase + (pair >> RL_PAIR_INDEX_BITS);
                        const uint32_t r = cluster_range[pair & RL_PAIR_INDEX_MAX];
                        const uint32_t end = (r & 0xffffu) + (r >> 16);
                        const float4 ro = ray_tab[3 * owner], rd = ray_tab[3 * owner + 1];
                        const float2 rt = *reinterpret_cast<const float2 *>(&ray_tab[3 * owner + 2]);
#pragma unroll 1
                        for (uint32_t m = (r & 0xffffu) + lane % RL_PAIR_LANES; m < end; m += RL_PAIR_LANES) {
                            const float4 s = GLOBAL_K ? __ldg(sphere_k_global + m) : sphere_k[m];   // {cx, cy, cz, |c|^2 - r^2}
                            const float b = fmaf(rd.x, s.x, fmaf(rd.y, s.y, fmaf(rd.z, s.z, rd.w)));
                            const float c = fmaf(ro.x, s.x, fmaf(ro.y, s.y, fmaf(ro.z, s.z, s.w)));
                            const float disc = fmaf(b, b, -c);
                            if (disc >= rt.x && b >= rt.y) {
                                const uint32_t slot = atomicAdd(&sq_cnt[owner], 1u);
                                if (slot < RL_CAND_SLOTS) sq_base[slot * nthreads + owner] = (uint16_t)m;
                            }
                        }
                    }
                }
                __syncwarp();
                // Level 3, per lane: exact Sphere::intersect for the candidates queued for this lane
                const uint32_t cnt = sq_cnt[tid];
                if (cnt > RL_CAND_SLOTS) {
                    // more candidates than slots (pathological): evaluate every sphere exactly
#pragma unroll 1
                    for (uint32_t k = 0; k < tb.n_spheres; k++) {
                        const float t = sphere_t(__ldg(spheres + k), ray);
                        if (t > 0.0f) consider(best, t, (int)__ldg(sphere_obj + k), (RL_HIT_SPHERE << 28) | k);
                    }
                } else {
#pragma unroll 1
                    for (uint32_t k = 0; k < cnt; k++) {
                        const uint32_t idx = sq_base[k * nthreads + tid];
                        const float t = sphere_t(__ldg(spheres + idx), ray);
                        if (t > 0.0f) consider(best, t, (int)__ldg(sphere_obj + idx), (RL_HIT_SPHERE << 28) | idx);
                    }
                }
                sq_cnt[tid] = 0u;
                __syncwarp();
            }
        }
        }
    };

    const uint32_t n_compounds = tb.n_compounds;
    sphere_phase();
    if (n_compounds == 0u) return best;                         // block-uniform
    if (DEEP) *reinterpret_cast<float2 *>(&ray_tab[3 * tid + 2].z) = make_float2(best.t, slab_inflate);
    else ray_tab[3 * tid + 2].z = best.t;                       // sphere and flat hits bound the bodies worth evaluating
    res_cnt[tid] = 0u;

    // Compound bodies, entirely within the warp (no block barrier: the warps of a block run
    // through Scene::intersect independently of each other).
    //  1. bounding spheres: the same uniform scan over the bodies' bound records; a hit the
    //     reference returns lies inside the (host-inflated) bound, so the line passes it:
    //     exactly B^2 - dd C >= 0, and the evaluated B'^2 - C' can fall short of that by the
    //     rounding of both (e1) and by |dd - 1| |C| only: threshold -slack; B >= -|d| R.
    //     Unbounded bodies are flagged in a mask that is OR-ed in;
    //  2. slab test, one lane per (lane, body) pair of the warp's list (the list is what balances
    //     the lanes); the surviving pairs are compacted in place;
    //  3. exact evaluation, again eight lanes per pair and one leaf per lane (eval_bodies_exact):
    //     the reference's recursion is decided from the leaves' own hits and containment tests,
    //     all independent of each other, or handed to the owner lane when it cannot be;
    //  4. every lane merges the results computed for its ray.
    const float4 *compounds = sm_vec(tb.compounds);
    const float4 *body_bounds = sm_vec(tb.body_bounds);
    const uint64_t *body_always = reinterpret_cast<const uint64_t *>(rl_smem + tb.body_always);
    const float4 *leaves = sm_vec(tb.leaves);
    const uint32_t *compound_obj = sm_u32(tb.compound_obj);
    const uint32_t group = lane >> 3, sub = lane & 7u;
    const uint32_t group_bits = 0xffu << (8u * group), lanes_below = (1u << lane) - 1u;
    const float thr_b = -slack;
    const float bthr_b = bthr - sqrtf(dd) * tb.body_rmax;
    const bool fast_bodies = tb.sphere_leaves == 0u;            // block-uniform
#pragma unroll 1
    for (uint32_t round = 0; warp_live && round < n_compounds; round += RL_BODIES_PER_ROUND) {
        const uint32_t round_end = min(round + RL_BODIES_PER_ROUND, n_compounds);
        uint64_t todo = scan_bounds(body_bounds + round, (round_end - round + 7u) & ~7u, d, ndo, m2ox, m2oy, m2oz, oo,
                                    thr_b, bthr_b) | body_always[round / RL_BODIES_PER_ROUND];
        while (__any_sync(0xffffffffu, todo != 0ull)) {
            const uint32_t npairs = emit_pairs(todo, round, pairs, lane);
            uint32_t nsurv = 0;                                 // warp-uniform: pairs that pass the slab test, compacted in place
#pragma unroll 1
#if RL_SLAB_LANES == 1
            // every lane takes one (lane, body) pair of the list and walks the body's half-spaces
            // itself, from a start staggered by lane (bodies of eight leaves start at multiples of
            // 256 bytes: the 32 reads of a step would otherwise meet in the same banks); max and min
            // are exact, so the interval does not depend on the order
            for (uint32_t pb = 0; pb < npairs; pb += 32) {
                const uint32_t p = pb + lane;
                const bool valid = p < npairs;
                const uint32_t pair = valid ? pairs[p] : 0u;
                const uint32_t owner = wbase + (pair >> RL_PAIR_INDEX_BITS), body = pair & RL_PAIR_INDEX_MAX;
                const float4 c4 = compounds[2 * body];
                const uint32_t first_leaf = __float_as_uint(c4.x);
                const uint32_t n_leaves = valid ? __float_as_uint(c4.y) : 0u;
                const float4 ro = ray_tab[3 * owner], rd = ray_tab[3 * owner + 1];
                const float2 bi = *reinterpret_cast<const float2 *>(&ray_tab[3 * owner + 2].z);
                const float best_t = bi.x, inflate = bi.y;
                const float ox = -0.5f * ro.x, oy = -0.5f * ro.y, oz = -0.5f * ro.z;   // ro = -2 o, exactly
                float t_enter = 0.0f, t_exit = 3.0e38f;
                bool outside_parallel = false;
                const uint32_t first_j = sub < n_leaves ? sub : 0u;
#pragma unroll 1
                for (uint32_t j = 0; j < n_leaves; j++) {
                    uint32_t jj = j + first_j;
                    if (jj >= n_leaves) jj -= n_leaves;
                    const uint32_t k = first_leaf + jj;
                    const float4 n4 = leaves[2 * k], o4 = leaves[2 * k + 1];
                    const float dn = fmaf(n4.z, rd.z, fmaf(n4.y, rd.y, n4.x * rd.x));
                    const float s0 = fmaf(n4.z, oz - o4.z, fmaf(n4.y, oy - o4.y, n4.x * (ox - o4.x))) - inflate * n4.w;
                    const float tk = __fdividef(-s0, dn);
                    if (dn < 0.0f) t_enter = fmaxf(t_enter, tk);
                    else if (dn > 0.0f) t_exit = fminf(t_exit, tk);
                    else if (s0 > 0.0f) outside_parallel = true;
                }
                const float start = t_enter * 0.9999f - 1.0e-3f;
                const bool may_hit = valid && !outside_parallel && !(t_exit < 0.0f) && !(start > t_exit) && !(start > best_t);
                const uint32_t keep = __ballot_sync(0xffffffffu, may_hit);
                __syncwarp();                                   // every lane has read its pair: the list may be overwritten
                if (may_hit) pairs[nsurv + __popc(keep & lanes_below)] = (uint16_t)pair;
                nsurv += __popc(keep);
            }
#else
            for (uint32_t pb = 0; pb < npairs; pb += 4) {
                const uint32_t p = pb + group;
                const bool valid = p < npairs;                  // uniform within a group of eight
                const uint32_t pair = valid ? pairs[p] : 0u;
                const uint32_t owner = wbase + (pair >> RL_PAIR_INDEX_BITS), body = pair & RL_PAIR_INDEX_MAX;
                const float4 c4 = compounds[2 * body];
                const uint32_t first_leaf = __float_as_uint(c4.x);
                const uint32_t n_leaves = valid ? __float_as_uint(c4.y) : 0u;
                const float4 ro = ray_tab[3 * owner], rd = ray_tab[3 * owner + 1];
                const float2 bi = *reinterpret_cast<const float2 *>(&ray_tab[3 * owner + 2].z);
                const float best_t = bi.x, inflate = bi.y;
                const float ox = -0.5f * ro.x, oy = -0.5f * ro.y, oz = -0.5f * ro.z;   // ro = -2 o, exactly
                // the slab test, one leaf per lane
                float t_enter = 0.0f, t_exit = 3.0e38f;
                bool outside_parallel = false;
#pragma unroll 1
                for (uint32_t k = first_leaf + sub; k < first_leaf + n_leaves; k += 8) {
                    const float4 n4 = leaves[2 * k], o4 = leaves[2 * k + 1];
                    const float dn = fmaf(n4.z, rd.z, fmaf(n4.y, rd.y, n4.x * rd.x));
                    const float s0 = fmaf(n4.z, oz - o4.z, fmaf(n4.y, oy - o4.y, n4.x * (ox - o4.x))) - inflate * n4.w;
                    const float tk = __fdividef(-s0, dn);
                    if (dn < 0.0f) t_enter = fmaxf(t_enter, tk);
                    else if (dn > 0.0f) t_exit = fminf(t_exit, tk);
                    else if (s0 > 0.0f) outside_parallel = true;
                }
#pragma unroll
                for (uint32_t sh = 1; sh < 8; sh <<= 1) {
                    t_enter = fmaxf(t_enter, __shfl_xor_sync(0xffffffffu, t_enter, sh));
                    t_exit = fminf(t_exit, __shfl_xor_sync(0xffffffffu, t_exit, sh));
                    outside_parallel |= (__shfl_xor_sync(0xffffffffu, (int)outside_parallel, sh) != 0);
                }
                const float start = t_enter * 0.9999f - 1.0e-3f;
                const bool may_hit = valid && !outside_parallel && !(t_exit < 0.0f) && !(start > t_exit) && !(start > best_t);
                const uint32_t keep = __ballot_sync(0xffffffffu, may_hit && sub == 0u);
                __syncwarp();                                   // every group has read its pair: the list may be overwritten
                if (may_hit && sub == 0u) pairs[nsurv + __popc(keep & lanes_below)] = (uint16_t)pair;
                nsurv += __popc(keep);
            }
#endif
            __syncwarp();
            // 3. exact evaluation of the survivors
#pragma unroll 1
            for (uint32_t sb = 0; sb < nsurv; sb += 4) {
                const uint32_t p = sb + group;
                const bool valid = p < nsurv;
                const uint32_t pair = valid ? pairs[p] : 0u;
                const uint32_t owner = wbase + (pair >> RL_PAIR_INDEX_BITS), body = pair & RL_PAIR_INDEX_MAX;
                const float4 c4 = compounds[2 * body];
                const uint32_t first_leaf = __float_as_uint(c4.x), n_leaves = __float_as_uint(c4.y);
                const bool use = valid && fast_bodies && n_leaves <= 8u;   // uniform within the group
                const float4 ro = ray_tab[3 * owner], rd = ray_tab[3 * owner + 1];
                Ray oray;                                       // the owner's ray, bit for bit
                oray.origin = mk(-0.5f * ro.x, -0.5f * ro.y, -0.5f * ro.z);
                oray.direction = mk(rd.x, rd.y, rd.z);
                oray.wavelength = 0.0f;
                uint32_t leaf_hit = 0;
                float t_hit = eval_body_exact(leaves + 2 * first_leaf, n_leaves, oray, use, sub, group_bits, leaf_hit);
                leaf_hit += first_leaf;
                if (!use) t_hit = -2.0f;                        // -2: the owner runs the reference's recursion itself
                if (valid && sub == 0u && t_hit != -1.0f) {     // -1: the body is missed
                    const uint32_t slot = atomicAdd(&res_cnt[owner], 1u);
                    if (slot < RL_COMPOUND_SLOTS)
                        results[slot * nthreads + owner] = make_float2(t_hit, __uint_as_float((body << 16) | leaf_hit));
                }
            }
            __syncwarp();
        }
        // 4. merge
        const uint32_t nres = res_cnt[tid];
        if (nres > RL_COMPOUND_SLOTS) {
            // more results than slots (pathological): evaluate every body of the round here
#pragma unroll 1
            for (uint32_t k = round; k < round_end; k++) {
                uint32_t leaf;
                const float t = compound_t(compounds[2 * k], ray, leaf);
                if (t > 0.0f) consider(best, t, (int)compound_obj[k], (RL_HIT_LEAF << 28) | leaf);
            }
        } else {
#pragma unroll 1
            for (uint32_t k = 0; k < nres; k++) {
                const float2 r = results[k * nthreads + tid];
                const uint32_t code = __float_as_uint(r.y);
                if (r.x == -2.0f) {                             // handed over: the reference's recursion, by the owner
                    uint32_t leaf;
                    const float t = compound_t(compounds[2 * (code >> 16)], ray, leaf);
                    if (t > 0.0f) consider(best, t, (int)compound_obj[code >> 16], (RL_HIT_LEAF << 28) | leaf);
                } else {
                    consider(best, r.x, (int)compound_obj[code >> 16], (RL_HIT_LEAF << 28) | (code & 0xffffu));
                }
            }
        }
        res_cnt[tid] = 0u;
        __syncwarp();
    }
    return best;
}

// for the probes: whichever instance the scene needs
__device__ __forceinline__ Hit intersect_scene_any(const Ray &ray, bool live = true) {
    if (tables().n_supers != 0u)
        return tables().sphere_k_global ? intersect_scene<true, true>(ray, live) : intersect_scene<false, true>(ray, live);
    return tables().sphere_k_global ? intersect_scene<true, false>(ray, live) : intersect_scene<false, false>(ray, live);
}

struct Surf { V3 position, normal, tangent; };

// The Intersection record of the winning primitive (intersection.rs:19-32);
// the reference builds it for every candidate, the result only needs the winner.
__device__ __forceinline__ Surf surface_at(const Ray &ray, const Hit &hit) {
    const PrimTables &tb = tables();
    Surf s;
    s.position = ray.origin + ray.direction * hit.t;
    s.tangent = mk(0.0f, 0.0f, 0.0f);
    const uint32_t type = hit.code >> 28, idx = hit.code & 0x0fffffffu;
    if (type == RL_HIT_SPHERE) {                                       // geometry.rs:243-251
        const float4 sp = __ldg(tb.spheres + idx);
        s.normal = normalise_dev(s.position - mk(sp.x, sp.y, sp.z));
        // the tangent (geometry.rs:250-251) is read by one material only: sphere_tangent() below
    } else if (type == RL_HIT_PLANE) {
        const float4 n4 = sm_vec(tb.planes)[2 * idx];
        const V3 n = mk(n4.x, n4.y, n4.z);
        if (__float_as_uint(n4.w) == RL_SURFACE_HALFSPACE) {
            s.normal = n;                                              // geometry.rs:115
        } else {
            const float d = dot(n, ray.direction);
            s.normal = d < 0.0f ? n : -n;                              // geometry.rs:80, :178
        }
    } else if (type == RL_HIT_PARABOLOID) {                            // geometry.rs:343-347
        const float4 *p = sm_vec(tb.paraboloids) + 3 * idx;
        const V3 offset = mk(p[0].x, p[0].y, p[0].z);
        const V3 normal = mk(p[1].x, p[1].y, p[1].z);
        const V3 focal_point = mk(p[2].x, p[2].y, p[2].z);
        const V3 local_pos = s.position - offset;
        const V3 plane_pr = local_pos - normal * dot(local_pos, normal);
        s.normal = normalise_dev(focal_point - plane_pr);
    } else {
        const float4 n4 = sm_vec(tb.leaves)[2 * idx];
        if (n4.w == 0.0f) {                                            // a sphere leaf, geometry.rs:243-248
            const float4 o4 = sm_vec(tb.leaves)[2 * idx + 1];
            s.normal = normalise_dev(s.position - mk(o4.x, o4.y, o4.z));
        } else {
            s.normal = mk(n4.x, n4.y, n4.z);                           // geometry.rs:115
        }
    }
    return s;
}

// Sphere::intersect's tangent, normalise(cross((0,1,0), normal)) (geometry.rs:250-251); every
// other surface leaves it zero.  Computed where it is read (SoapBubbleMaterial, material.rs:297).
__device__ __forceinline__ V3 sphere_tangent(const Hit &hit, const Surf &s) {
    const uint32_t type = hit.code >> 28;
    const bool sphere = type == RL_HIT_SPHERE
                        || (type == RL_HIT_LEAF && sm_vec(tables().leaves)[2 * (hit.code & 0x0fffffffu)].w == 0.0f);
    if (!sphere) return mk(0.0f, 0.0f, 0.0f);
    return normalise_dev(cross(mk(0.0f, 1.0f, 0.0f), s.normal));
}

// ---------------------------------------------------------------- materials
// material.rs:38-58 with monte_carlo.rs:47-58
__device__ __forceinline__ V3 diffuse_direction(const Ray &in, const Surf &s, BounceRng &rng) {
    const float phi = rng.longitude();
    const float rq = rng.unit();
    const float r = sqrtf(rq);
    const float2 sc = sincos_call(phi);
    const V3 hemi = mk(sc.y * r, sc.x * r, sqrtf(1.0f - rq));
    const V3 normal = dot(in.direction, s.normal) < 0.0f ? s.normal : -s.normal;
    return rotate_towards_dev(hemi, normal);
}

__device__ __forceinline__ float soap_clamp(float x) {                 // material.rs:290-294
    return x < -0.999f ? -0.999f : (x > 0.999f ? 0.999f : x);
}

// Material::get_new_ray for the five reflective materials; returns the new
// direction and the ray's probability (origin = intersection position).  The
// three diffuse-based materials share one copy of get_diffuse_ray.
__device__ __forceinline__ V3 material_bounce(float4 m, const Ray &in, const Hit &hit, const Surf &s, BounceRng &rng,
                                              float &probability) {
    const uint32_t kind = __float_as_uint(m.x);
    if (kind <= RL_MATERIAL_GLOSSY_MIRROR) {
        // grey (material.rs:122-130), coloured (:155-168), glossy (:185-196)
        probability = kind == RL_MATERIAL_DIFFUSE_GREY ? m.y : 1.0f;
        if (kind == RL_MATERIAL_DIFFUSE_COLOURED) {
            const float p = (m.z - in.wavelength) / m.w;
            probability = m.y * spec_exp(-0.5f * p * p);
        }
        V3 dir = diffuse_direction(in, s, rng);
        if (kind == RL_MATERIAL_GLOSSY_MIRROR) {
            const V3 reflection = reflect(in.direction, s.normal);
            dir = normalise_dev(dir * m.y + reflection * (1.0f - m.y));
        }
        return dir;
    }
    if (kind == RL_MATERIAL_SF10_GLASS) {                              // material.rs:216-261
        float cos_i = -dot(in.direction, s.normal);
        float ior = sf10_index_of_refraction(in.wavelength);
        V3 normal = s.normal;
        if (cos_i > 0.0f) {
            ior = 1.0f / ior;
        } else {
            normal = -normal;
            cos_i = -cos_i;
        }
        const float sin_t_sqr = ior * ior * (1.0f - cos_i * cos_i);
        probability = 1.0f;
        if (sin_t_sqr > 1.0f) return reflect(in.direction, normal);
        const float cos_t = sqrtf(1.0f - sin_t_sqr);
        return in.direction * ior + normal * (ior * cos_i - cos_t);
    }
    // RL_MATERIAL_SOAP_BUBBLE                                            material.rs:267-306
    const float cos_alpha = dot(in.direction, s.normal);
    const V3 direction = (rng.unit() - 0.3f > fabsf(cos_alpha)) ? reflect(in.direction, s.normal)
                                                                : in.direction;
    const float phase_shift = (in.wavelength - 380.0f) / 200.0f * RL_PI;
    const float cos_phi = soap_clamp(dot(direction, s.normal));
    const float cos_theta = soap_clamp(dot(direction, sphere_tangent(hit, s)));
    const float2 sc = sincos_call(phase_shift - spec_acos(cos_phi) * 3.0f - spec_acos(cos_theta) * 2.0f
                                  + RL_PI * 0.5f);
    probability = sc.y * 0.1f + 0.9f;
    return direction;
}

// material.rs:101-105
__device__ __forceinline__ float blackbody_intensity(float4 m, float wavelength) {
    return (float)boltzmann((double)wavelength, (double)m.y) * m.z;
}

// --------------------------------------------------------------------- plot
__constant__ float c_cie[81][3] = {
#include "rl_cie1931_data.inc"
};

// Where the CIE table is read from.  The constant bank serves one address per instruction: right
// for the trace kernel, where one or two lanes of a warp end a lit path at a time, wrong for the
// splat kernel, where 32 lit photons with 32 different wavelengths look the table up together
// (each of the six loads would be replayed once per distinct address).  The splat kernel keeps a
// float4 copy of the table in shared memory.
struct CieConstant {
    __device__ __forceinline__ V3 operator()(int i) const { return mk(c_cie[i][0], c_cie[i][1], c_cie[i][2]); }
};
struct CieShared {
    const float4 *tab;                                          // [81] {X, Y, Z, 0}
    __device__ __forceinline__ V3 operator()(int i) const { const float4 v = tab[i]; return mk(v.x, v.y, v.z); }
};
// fills a CieShared table; call with every thread of the block, then __syncthreads()
__device__ __forceinline__ void fill_cie_shared(float4 *tab) {
    for (uint32_t i = threadIdx.x; i < 81; i += blockDim.x) tab[i] = make_float4(c_cie[i][0], c_cie[i][1], c_cie[i][2], 0.0f);
}

// cie1931.rs:20-48
template <typename Table>
__device__ __forceinline__ V3 tristimulus_from(const Table &cie, float wavelength) {
    const float indexf = (wavelength - 380.0f) / 5.0f;
    const int index = (int)floorf(indexf);
    const float remainder = indexf - (float)index;
    if (index < -1 || index > 80) return mk(0.0f, 0.0f, 0.0f);
    if (index == -1) { const V3 a = cie(0); return mk(a.x * remainder, a.y * remainder, a.z * remainder); }
    if (index == 80) {
        const float w = 1.0f - remainder;
        const V3 a = cie(80);
        return mk(a.x * w, a.y * w, a.z * w);
    }
    const float w = 1.0f - remainder;
    const V3 a = cie(index), b = cie(index + 1);
    return mk(a.x * w + b.x * remainder, a.y * w + b.y * remainder, a.z * w + b.z * remainder);
}
__device__ __forceinline__ V3 tristimulus(float wavelength) { return tristimulus_from(CieConstant(), wavelength); }

__device__ __forceinline__ void red_add_v4(float4 *addr, float x, float y, float z) {
#ifdef RL_PROBE_NO_RED
    // experiments only (tools/splat_probe.py --no-red): everything but the reduction itself
    asm volatile("" :: "l"(addr), "f"(x), "f"(y), "f"(z) : "memory");
#else
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(x), "f"(y), "f"(z), "f"(0.0f) : "memory");
#endif
}

// PlotUnit::plot for one photon (plot_unit.rs:56-95) into the padded
// accumulator (one float4 {X, Y, Z, 0} per pixel): four vector reductions.
template <typename Table>
__device__ __forceinline__ void splat_photon_from(const Table &table, float4 *accum, int w, int h, float aspect,
                                                  float x, float y, float wavelength, float probability) {
    const V3 cie = tristimulus_from(table, wavelength) * probability;
    const float px = (x * 0.5f + 0.5f) * ((float)w - 1.0f);
    const float py = (y * aspect * 0.5f + 0.5f) * ((float)h - 1.0f);
    const int px1 = max(0, min(w - 1, (int)floorf(px)));
    const int px2 = max(0, min(w - 1, (int)ceilf(px)));
    const int py1 = max(0, min(h - 1, (int)floorf(py)));
    const int py2 = max(0, min(h - 1, (int)ceilf(py)));
    const float cx = px - (float)px1;
    const float cy = py - (float)py1;
    const float c11 = (1.0f - cx) * (1.0f - cy);
    const float c12 = (1.0f - cx) * cy;
    const float c21 = cx * (1.0f - cy);
    const float c22 = cx * cy;
    red_add_v4(accum + ((size_t)py1 * w + px1), cie.x * c11, cie.y * c11, cie.z * c11);
    red_add_v4(accum + ((size_t)py1 * w + px2), cie.x * c21, cie.y * c21, cie.z * c21);
    red_add_v4(accum + ((size_t)py2 * w + px1), cie.x * c12, cie.y * c12, cie.z * c12);
    red_add_v4(accum + ((size_t)py2 * w + px2), cie.x * c22, cie.y * c22, cie.z * c22);
}
__device__ __forceinline__ void splat_photon(float4 *accum, int w, int h, float aspect, float x,
                                             float y, float wavelength, float probability) {
    splat_photon_from(CieConstant(), accum, w, h, aspect, x, y, wavelength, probability);
}

}  // namespace rl
