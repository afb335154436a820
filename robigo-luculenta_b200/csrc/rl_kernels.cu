// rl_kernels.cu -- the sm_100a kernels of the path.
//
//   K1 trace_kernel    TraceUnit::render (trace_unit.rs:81-168), optionally
//                      fused with PlotUnit::plot (plot_unit.rs:56-95)
//   K2 splat_kernel    PlotUnit::plot over MappedPhoton records
//   K3 gather_kernel   GatherUnit::accumulate (gather_unit.rs:49-64) + PlotUnit::clear
//   K4 tonemap_*       TonemapUnit::tonemap (tonemap_unit.rs:55-100, srgb.rs:20-41)
//
// Compile with -fmad=false: see rl_math.cuh.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "rl_kernels.h"
#include "rl_device.cuh"

namespace rl {

static std::atomic<uint64_t> g_launches{0};
uint64_t kernel_launches() { return g_launches.load(); }
void kernel_launches_reset() { g_launches.store(0); }

#ifndef RL_TRACE_THREADS
#define RL_TRACE_THREADS 768
#endif
#ifndef RL_TRACE_MIN_BLOCKS
#define RL_TRACE_MIN_BLOCKS 1
#endif
// launches with fewer than RL_TRACE_SMALL_PATHS photons per thread of a full grid use CTAs of
// RL_TRACE_SMALL_CTA threads (0 paths: never); both can be overridden from the environment
#ifndef RL_TRACE_SMALL_CTA
#define RL_TRACE_SMALL_CTA 384
#endif
#ifndef RL_TRACE_SMALL_PATHS
#define RL_TRACE_SMALL_PATHS 64
#endif

// ------------------------------------------------------------------ K1 trace
// One launch traces up to RL_MAX_SEGMENTS batches ("segments": the photon ranges of several
// TraceUnit::render calls on one scene and canvas, rl_api.cu groups them) as one pool of photons.
struct TraceArgs {
    uint32_t n_seg;
    uint32_t seg_shift;        // a path remembers its photon as (segment << seg_shift) | index within the segment
    uint32_t ring_cap;         // block ring: camera rays it holds (a power of two); warp rings hold RL_WARP_RING_ENTRIES
    int width, height;
    float aspect;
    float4 *accum;
    TraceSegment seg[RL_MAX_SEGMENTS];
};

// Per-CTA bookkeeping in shared memory, behind the intersection scratch; the threads' parked
// screen positions and the warps' camera-ray rings follow it.
struct TraceCta {
    uint64_t seed[RL_MAX_SEGMENTS];
    uint64_t first_photon[RL_MAX_SEGMENTS];
    rl_mapped_photon *records[RL_MAX_SEGMENTS];
    uint32_t seg_start[RL_MAX_SEGMENTS + 1];   // first index of each segment in the launch's concatenated photon range
    uint32_t seg_rays[RL_MAX_SEGMENTS];        // Scene::intersect calls of the paths this CTA finished, per segment
    uint32_t pool_next;                        // warp rings: next photon of the block's range that no warp has taken yet
    uint32_t warp_dead[2][32];                 // block ring: lanes without a path, per warp (double-buffered by iteration parity)
    uint32_t pad[2];
};
// behind TraceCta: the screen position of every thread's current photon (float2 per thread; only
// read again when the path ends), then the camera-ray ring(s): one per block or one per warp
#define RL_TRACE_PARK_BYTES_PER_THREAD 8
#define RL_WARP_RING_ENTRIES 32u               // per warp, three float4 each
static_assert(sizeof(TraceCta) % 16 == 0, "the rings behind TraceCta hold float4");

__device__ __forceinline__ TraceCta *trace_cta(const DevScene &sc) {
    char *base = reinterpret_cast<char *>(rl_smem + RL_TABLES_VEC4 + sc.smem_vec4);
    return reinterpret_cast<TraceCta *>(base + (size_t)scratch_bytes_per_thread(sc.n_compounds) * blockDim.x);
}

// trace_unit.rs:151-158 and :136-145 for the photon at index `gidx` of the launch: the draws of
// the wavelength, the screen position and the time, and the camera ray (which draws the lens
// sample): the six draws of Philox blocks 0 and 1; the bounces draw from block 2 on (BounceRng).
__device__ __forceinline__ void generate_camera_entry(const DevScene &sc, const TraceArgs &a, const TraceCta *cta,
                                                      float4 *entry, uint32_t gidx) {
    uint32_t seg = 0;
    while (seg + 1 < a.n_seg && gidx >= cta->seg_start[seg + 1]) seg++;
    const uint32_t local = gidx - cta->seg_start[seg];
    const RngKey key = {cta->seed[seg], cta->first_photon[seg] + local};
    Rng rng;
    rng.init();
    const float wavelength = rng.wavelength(key);
    const float sx = rng.bi_unit(key);
    const float sy = rng.bi_unit(key) / a.aspect;
    const float t = rng.unit(key);
    const Ray ray = camera_ray(sc.camera, sx, sy, wavelength, t, rng, key);
    entry[0] = make_float4(ray.origin.x, ray.origin.y, ray.origin.z, ray.direction.x);
    entry[1] = make_float4(ray.direction.y, ray.direction.z, wavelength, sx);
    entry[2] = make_float4(sy, 0.0f, 0.0f, __uint_as_float((seg << a.seg_shift) | local));
}

// Persistent threads with path regeneration.  A block owns a contiguous range of the launch's
// photons.  Camera rays are produced in bulk into a ring of ready rays: whenever the ring cannot
// serve the lanes whose paths have just ended, ALL threads that share it generate one (the RNG
// draws, three sines and cosines, two normalisations and two quaternion rotations of
// trace_unit.rs:136-145 / camera.rs:47-108 run with full warps instead of for the two lanes in
// seven that need a new path in a given iteration), and a lane takes its next photon from the ring
// the moment the path it holds ends -- so a warp never idles on its longest path (1 ... ~150
// bounces) and no lane idles while the block has photons left.  The result of a photon depends
// on its id alone, so the deal changes nothing but the order of the accumulator atomics.  The
// primitive tables live in shared memory; each loop iteration is one Scene::intersect plus one
// material interaction for every live lane.
//
// Nothing in an iteration crosses a warp (Scene::intersect does not either), and the kernel
// exists in two forms:
//   BLOCK_RING = false: one ring per warp, no block barrier between set-up and the final ray
//     count; a warp with a long iteration holds nobody back.
//   BLOCK_RING = true: one ring per block, refilled by all its threads together, and the block's
//     warps meet at ONE barrier per iteration.  That barrier is there for the instruction cache:
//     the loop's code (sphere scan, bodies, six materials, two f64 functions) is larger than the
//     cache, and warps that drift apart fetch different parts of it -- free-running, the built-in
//     scene is 22 % slower (4.5 instruction-fetch stall cycles per issued instruction instead of
//     0.3).  Scenes without compound bodies execute a small enough part of the loop and are
//     10-20 % faster free-running: launch_trace picks the form by that.
template <bool BLOCK_RING, bool GLOBAL_K, bool DEEP>
__global__ void __launch_bounds__(RL_TRACE_THREADS, RL_TRACE_MIN_BLOCKS)
trace_kernel(const DevScene sc, const TraceArgs a) {
    setup_tables(sc);
    TraceCta *cta = trace_cta(sc);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    float2 *park = reinterpret_cast<float2 *>(cta + 1) + threadIdx.x;
    float4 *ring = reinterpret_cast<float4 *>(reinterpret_cast<float2 *>(cta + 1) + ((blockDim.x + 1u) & ~1u))
                   + (BLOCK_RING ? 0u : 3u * RL_WARP_RING_ENTRIES * warp);
    if (threadIdx.x < RL_MAX_SEGMENTS) {
        const uint32_t k = threadIdx.x;
        cta->seg_rays[k] = 0u;
        if (k < a.n_seg) {
            cta->seed[k] = a.seg[k].seed;
            cta->first_photon[k] = a.seg[k].first_photon;
            cta->records[k] = a.seg[k].records;
        }
    }
    if (threadIdx.x == 0) {
        uint32_t at = 0;
        for (uint32_t k = 0; k < a.n_seg; k++) { cta->seg_start[k] = at; at += (uint32_t)a.seg[k].n_photons; }
        cta->seg_start[a.n_seg] = at;
        cta->pool_next = (uint32_t)((uint64_t)at * blockIdx.x / gridDim.x);
    }
    __syncthreads();
    const uint32_t pool_begin = (uint32_t)((uint64_t)cta->seg_start[a.n_seg] * blockIdx.x / gridDim.x);
    const uint32_t pool_end = (uint32_t)((uint64_t)cta->seg_start[a.n_seg] * (blockIdx.x + 1ull) / gridDim.x);

    // the ring the thread takes its camera rays from, as every thread that shares it sees it:
    // entries produced / handed out so far; block ring: next photon to generate
    uint32_t tail = 0, head = 0, gen_next = pool_begin, parity = 0;
    bool pool_empty = false;
    bool alive = false;
    uint32_t cur = 0;                                               // (segment << seg_shift) | photon index in it
    Ray ray;
    ray.origin = mk(0.f, 0.f, 0.f); ray.direction = mk(0.f, 0.f, 0.f); ray.wavelength = 0.f;
    float intensity = 1.0f, continue_chance = 1.0f;
    BounceRng rng;
    rng.d0 = rng.d1 = rng.d2 = 0u;
    uint32_t rays = 0;                                              // of the current path = its bounces so far

    for (;;) {
        const uint32_t want = __ballot_sync(0xffffffffu, !alive);
        uint32_t slot = 0xffffffffu;                                // ring entry this lane takes
        if (BLOCK_RING) {
            // lanes without a path, per warp and in the block
            if (lane == 0) cta->warp_dead[parity][warp] = (uint32_t)__popc(want);
            const uint32_t dead = (uint32_t)__syncthreads_count(!alive);
            const uint32_t T = blockDim.x, cap_mask = a.ring_cap - 1u;
            uint32_t avail = tail - head;
            if (avail < dead && gen_next < pool_end) {
                uint32_t n_new = a.ring_cap - avail;
                if (n_new > T) n_new = T;
                if (n_new > pool_end - gen_next) n_new = pool_end - gen_next;
                if (threadIdx.x < n_new)
                    generate_camera_entry(sc, a, cta, ring + 3u * ((tail + threadIdx.x) & cap_mask), gen_next + threadIdx.x);
                __syncthreads();
                tail += n_new; gen_next += n_new; avail += n_new;
            }
            if (dead == T && avail == 0u) break;                    // no path alive, no photon left
            if (avail != 0u && want != 0u) {
                // the lanes without a path take the oldest ready camera rays, in thread order: a
                // warp's first entry is the number of such lanes in the warps before it
                const uint32_t before = __reduce_add_sync(0xffffffffu, lane < warp ? cta->warp_dead[parity][lane] : 0u);
                const uint32_t idx = head + before + __popc(want & ((1u << lane) - 1u));
                if (!alive && idx < tail) slot = idx & cap_mask;
            }
            head += dead < avail ? dead : avail;
            parity ^= 1u;
        } else {
            const uint32_t dead = (uint32_t)__popc(want);
            uint32_t avail = tail - head;
            if (avail < dead && !pool_empty) {
                // refill: as many camera rays as the ring has room for, from the block's photon range
                uint32_t n_new = RL_WARP_RING_ENTRIES - avail, base = 0;
                if (lane == 0) base = atomicAdd(&cta->pool_next, n_new);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= pool_end) { n_new = 0; pool_empty = true; }
                else if (n_new >= pool_end - base) { n_new = pool_end - base; pool_empty = true; }
                if (lane < n_new)
                    generate_camera_entry(sc, a, cta, ring + 3u * ((tail + lane) & (RL_WARP_RING_ENTRIES - 1u)), base + lane);
                __syncwarp();
                tail += n_new; avail += n_new;
            }
            if (dead == 32u && avail == 0u) break;                  // no path alive, no photon left for this warp
            const uint32_t rank = (uint32_t)__popc(want & ((1u << lane) - 1u));
            if (!alive && rank < avail) slot = (head + rank) & (RL_WARP_RING_ENTRIES - 1u);
            head += dead < avail ? dead : avail;
        }
        if (slot != 0xffffffffu) {
            const float4 *e = ring + 3u * slot;
            const float4 e0 = e[0], e1 = e[1], e2 = e[2];
            ray.origin = mk(e0.x, e0.y, e0.z);
            ray.direction = mk(e0.w, e1.x, e1.y);
            ray.wavelength = e1.z;
            *park = make_float2(e1.w, e2.x);                        // MappedPhoton x, y
            cur = __float_as_uint(e2.w);
            intensity = 1.0f;
            continue_chance = 1.0f;
            rays = 0;
            alive = true;
        }
        if (!BLOCK_RING) __syncwarp();                              // the entries are read before a refill overwrites them
        const uint32_t seg = cur >> a.seg_shift, index = cur & ((1u << a.seg_shift) - 1u);
        if (alive) {
            // the draws of this bounce, for every live lane at once (independent of the intersection
            // below, which hides the latency of the ten Philox rounds)
            const RngKey key = {cta->seed[seg], cta->first_photon[seg] + index};
            rng.load(key, rays);
        }
        const Hit hit = intersect_scene<GLOBAL_K, DEEP>(alive ? ray : idle_ray(), alive);
        if (alive) {
            rays++;                                                 // Scene::intersect calls (scene.rs:39)
            // trace_unit.rs:91-131
            bool done = false;
            float result = 0.0f;
            if (hit.obj < 0) {
                done = true;                                            // trace_unit.rs:94
            } else {
                const float4 m = __ldg(sc.materials + hit.obj);
                if (__float_as_uint(m.x) == RL_MATERIAL_BLACKBODY) {
                    result = intensity * blackbody_intensity(m, ray.wavelength);  // :99-101
                    done = true;
                } else {
                    const Surf s = surface_at(ray, hit);
                    float probability;
                    const V3 dir = material_bounce(m, ray, hit, s, rng, probability);       // :104-107
                    intensity = intensity * probability;
                    ray.direction = dir;
                    ray.origin = s.position + dir * 0.00001f;                      // :114
                    continue_chance = continue_chance * 0.96f;                    // :117
                    if (rng.unit() * 0.85f
                        > continue_chance * (1.0f - spec_exp(intensity * -20.0f)))  // :122-125
                        done = true;
                }
            }
            if (done) {
                const float2 screen = *park;
                rl_mapped_photon *records = cta->records[seg];
                if (records)
                    *reinterpret_cast<float4 *>(records + index) = make_float4(screen.x, screen.y, result, ray.wavelength);
                // adding cie * 0 leaves the accumulator unchanged (plot_unit.rs:80-83)
                if (a.accum && result != 0.0f)
                    splat_photon(a.accum, a.width, a.height, a.aspect, screen.x, screen.y, ray.wavelength, result);
                atomicAdd(&cta->seg_rays[seg], rays);
                alive = false;
            }
        }
    }

    // rays traced = Scene::intersect calls (scene.rs:39), per segment (= per trace unit)
    __syncthreads();
    if (threadIdx.x < a.n_seg) {
        const uint32_t r = cta->seg_rays[threadIdx.x];
        unsigned long long *counter = a.seg[threadIdx.x].ray_counter;
        if (r != 0u && counter) atomicAdd(counter, (unsigned long long)r);
    }
}

// block ring entries for CTAs of `threads` threads: the power of two at or above the CTA size (a
// refill is never cut short), or a half / a quarter of it (at least 128) when shared memory is
// short -- lanes that find the ring empty wait one iteration for the next refill
static uint32_t block_ring_entries(int threads, int shrink) {
    uint32_t cap = 128;
    while ((int)cap < threads) cap <<= 1;
    for (int k = 0; k < shrink && cap > 128; k++) cap >>= 1;
    return cap;
}
// CTAs of `threads` threads that fit an SM (228 KB of shared memory, 1 KB reserved per CTA; 2048
// threads; 64 K registers at 80 per thread)
static size_t trace_ctas_per_sm(size_t smem, int threads) {
    const size_t by_smem = (228u * 1024u) / (smem + 1024u);
    const size_t regs = RL_TRACE_THREADS > 768 ? 64u : 80u;      // what __launch_bounds__ leaves the compiler
    const size_t by_threads = 2048u / (size_t)threads, by_regs = 65536u / (regs * (size_t)threads);
    const size_t cap = by_threads < by_regs ? by_threads : by_regs;
    return by_smem < cap ? by_smem : cap;
}
// ring_cap > 0: one ring of that many entries per block; 0: one ring per warp
static size_t trace_kernel_smem_bytes(const DevScene &sc, int threads, uint32_t ring_cap) {
    const size_t entries = ring_cap ? (size_t)ring_cap : (size_t)(threads / 32) * RL_WARP_RING_ENTRIES;
    return tracing_smem_bytes(sc, threads) + sizeof(TraceCta) + (size_t)RL_TRACE_PARK_BYTES_PER_THREAD * ((threads + 1) & ~1)
           + entries * 3 * sizeof(float4);
}
size_t trace_smem_bytes(const DevScene &sc, int threads) { return tracing_smem_bytes(sc, threads); }


// Small launches in flight: one mark per stream, an event re-recorded behind that stream's latest
// small launch.  launch_trace (under its lock) counts the OTHER streams whose mark has not
// completed -- the concurrency the host's worker threads are producing right now (app.rs:95-111
// runs C of them over 3C trace units) -- and stops at the number that already gives the smallest
// share, scanning from where it last found one in flight: two or three event queries per launch
// in the steady state, however many units there are.
struct StreamMark { cudaStream_t stream; int device; cudaEvent_t done; };
static std::vector<StreamMark> g_marks;
static size_t g_scan_from = 0;

static int other_streams_in_flight(int device, cudaStream_t mine, int enough) {
    const size_t n = g_marks.size();
    int found = 0;
    for (size_t k = 0; k < n && found < enough; k++) {
        const size_t i = (g_scan_from + k) % n;
        const StreamMark &m = g_marks[i];
        if (m.device != device || m.stream == mine) continue;
        if (cudaEventQuery(m.done) == cudaErrorNotReady) {
            if (found == 0) g_scan_from = i;
            found++;
        }
    }
    cudaGetLastError();   // cudaErrorNotReady is not an error
    return found;
}

static void note_small_launch(int device, cudaStream_t st) {
    for (StreamMark &m : g_marks)
        if (m.stream == st && m.device == device) {
            if (cudaEventRecord(m.done, st) != cudaSuccess) cudaGetLastError();
            return;
        }
    if (g_marks.size() >= 4096) {                       // streams of units long gone: forget the finished ones
        size_t keep = 0;
        for (size_t i = 0; i < g_marks.size(); i++) {
            if (cudaEventQuery(g_marks[i].done) == cudaErrorNotReady) g_marks[keep++] = g_marks[i];
            else cudaEventDestroy(g_marks[i].done);
        }
        g_marks.resize(keep);
        g_scan_from = 0;
        cudaGetLastError();
    }
    StreamMark m{st, device, nullptr};
    if (cudaEventCreateWithFlags(&m.done, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaEventRecord(m.done, st) == cudaSuccess) g_marks.push_back(m);
    else { cudaGetLastError(); cudaEventDestroy(m.done); }
}

static int env_int(const char *name, int fallback) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : fallback;
}

// What the attribute calls last set, per device and kernel: thousands of identical launches per
// second should not each pay for driver round trips.  Guarded by the caller's lock.
struct KernelCache { int max_smem = 0, threads = 0, per_sm = 0, pct = -2; size_t smem = 0; };

template <typename Kernel>
static cudaError_t prepare_kernel(Kernel kernel, KernelCache &cached, int dev) {
    if (cached.max_smem != 0) return cudaSuccess;
    cudaError_t err = cudaDeviceGetAttribute(&cached.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (err == cudaSuccess)
        err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cached.max_smem);
    if (err != cudaSuccess) cached.max_smem = 0;
    return err;
}

// carve out only the shared memory the resident CTAs need (+1 KB each that the system
// reserves); the rest of the 228 KB stays L1 for the material records, the exact sphere
// records and the few spilled registers
template <typename Kernel>
static void set_carveout(Kernel kernel, KernelCache &cached, int per_sm, size_t smem) {
    const char *env = getenv("RL_TRACE_CARVEOUT");
    int pct = env ? atoi(env) : (int)((per_sm * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    if (pct >= 0 && pct != cached.pct) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cached.pct = pct;
    }
}

cudaError_t launch_trace(const DevScene &sc, const TraceLaunch &p, int sm_count, cudaStream_t st) {
    if (p.n_segments == 0 || p.n_segments > RL_MAX_SEGMENTS) return p.n_segments ? cudaErrorInvalidValue : cudaSuccess;
    uint64_t n_photons = 0;
    for (uint32_t k = 0; k < p.n_segments; k++) {
        if (p.n_segments > 1 && p.seg[k].n_photons >= (1ull << RL_SEGMENT_INDEX_BITS)) return cudaErrorInvalidValue;
        n_photons += p.seg[k].n_photons;
    }
    if (n_photons == 0) return cudaSuccess;
    // the function attributes below are process-wide: launches from the threads of different
    // units (each on its own stream) take turns setting them and launching
    static std::mutex launch_lock;
    std::lock_guard<std::mutex> guard(launch_lock);
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    KernelCache scratch_entry;
    // which instance of the kernel: see trace_kernel and intersect_scene.  RL_TRACE_LOCKSTEP
    // overrides the ring form (experiments).
    const bool block_ring = env_int("RL_TRACE_LOCKSTEP", sc.n_compounds != 0u ? 1 : 0) != 0;
    const bool global_k = sc.sphere_k_global != 0u;
    typedef void (*TraceKernel)(const DevScene, const TraceArgs);
    const bool deep = sc.n_supers != 0u;
    static const TraceKernel instances[8] = {
        trace_kernel<false, false, false>, trace_kernel<false, true, false>, trace_kernel<true, false, false>,
        trace_kernel<true, true, false>,   trace_kernel<false, false, true>, trace_kernel<false, true, true>,
        trace_kernel<true, false, true>,   trace_kernel<true, true, true>};
    const int which = (deep ? 4 : 0) + (block_ring ? 2 : 0) + (global_k ? 1 : 0);
    const TraceKernel kernel = instances[which];
    static KernelCache caches[8][16];
    KernelCache &cached = dev >= 0 && dev < 16 ? caches[which][dev] : scratch_entry;
    err = prepare_kernel(kernel, cached, dev);
    if (err != cudaSuccess) return err;
    const int max_smem = cached.max_smem;
    int threads = env_int("RL_TRACE_THREADS_MAX", RL_TRACE_THREADS);
    if (threads > RL_TRACE_THREADS || threads < 128 || threads % 128) threads = RL_TRACE_THREADS;
    // the largest CTA that fits; block ring: the roomiest one that does not cost a resident CTA
    auto pick_ring = [&](int t) -> int {
        if (!block_ring) return trace_kernel_smem_bytes(sc, t, 0) <= (size_t)max_smem ? 0 : -1;
        int best = -1;
        size_t best_ctas = 0;
        const int forced = env_int("RL_TRACE_RING_SHRINK", -1);      // experiments: a half (1) or a quarter (2) of the ring
        for (int shrink = forced >= 0 && forced <= 2 ? forced : 0; shrink <= 2; shrink++) {
            const size_t bytes = trace_kernel_smem_bytes(sc, t, block_ring_entries(t, shrink));
            if (bytes > (size_t)max_smem) continue;
            const size_t ctas = trace_ctas_per_sm(bytes, t);
            if (ctas > best_ctas) { best_ctas = ctas; best = shrink; }
        }
        return best;
    };
    int shrink = pick_ring(threads);
    while (shrink < 0) {
        if (threads <= 128) return cudaErrorInvalidValue;
        threads -= 128;
        shrink = pick_ring(threads);
    }
    // Small batches (the reference's 524 288 photons are 4.6 per thread of a full grid) would spend
    // most of their time in the tail, where a block waits for its last paths.  They are launched
    // as half-size CTAs, two per SM, and as few of them as the concurrency allows (below), so that
    // a block lives for tens of photons per thread and the block beside it on the SM -- another
    // unit's batch, from another stream -- covers its tail.  Scheduler replay, 6144 reference
    // batches, built-in scene (tools/replay_knobs.sh, profiles/r2_replay_knobs.txt), Mrays/s:
    // CTAs of 128: 2554, 256: 2954, 384: 3152, 512: 2703, 768 (no small launches): 2734; one
    // 2^28-photon launch: 3450.
    const int small_cta = env_int("RL_TRACE_SMALL_CTA", RL_TRACE_SMALL_CTA);
    const uint64_t small_paths = (uint64_t)env_int("RL_TRACE_SMALL_PATHS", RL_TRACE_SMALL_PATHS);
    const bool small = small_cta >= 128 && small_cta < threads && small_cta % 32 == 0
                       && n_photons < small_paths * (uint64_t)sm_count * (uint64_t)threads;
    if (small) {
        threads = small_cta;
        shrink = pick_ring(threads);
        if (shrink < 0) return cudaErrorInvalidValue;
        // two CTAs per SM leave 16 KB of the 228 for L1 (spills, material and exact sphere records):
        // half a ring each gives some of it back (+1.2 % on the replay, profiles/r2_ring_knobs.txt)
        if (block_ring && shrink == 0 && env_int("RL_TRACE_RING_SHRINK", -1) < 0) shrink = 1;
    }
    const uint32_t ring_cap = block_ring ? block_ring_entries(threads, shrink) : 0u;
    const size_t smem = trace_kernel_smem_bytes(sc, threads, ring_cap);
    if (cached.threads != threads || cached.smem != smem) {
        // the occupancy the runtime reports counts with the carve-out in force, i.e. the one the
        // PREVIOUS launch shape asked for (after one big launch -- a single CTA per SM, 67 % -- the
        // small launches were told that only one of their CTAs fits, and ran at two thirds of their
        // rate): ask with the largest carve-out, set_carveout below lowers it to what is needed
        if (cached.pct != 100) {
            cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            cached.pct = 100;
        }
        int occ = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
        if (err != cudaSuccess) return err;
        cached.threads = threads; cached.smem = smem; cached.per_sm = occ < 1 ? 1 : occ;
    }
    const int per_sm = cached.per_sm;
    set_carveout(kernel, cached, per_sm, smem);
    uint64_t full = (uint64_t)sm_count * per_sm;
    if (small) {
        // The launch's share of the GPU's block slots, by the concurrency seen: with k other
        // streams' small launches queued or running it asks for 1 / (k + 1) of the slots, counting
        // up to RL_TRACE_SHARE_MAX launches.  The host's worker threads run far ahead of the GPU,
        // so in the steady state a batch goes out as a few dozen blocks with tens of photons per
        // thread, a dozen batches side by side on every SM: a block then spends most of its life
        // with full warps, and its tail is covered by the blocks of other batches beside it.
        // RL_TRACE_BLOCKS_PER_SM > 0 fixes the share per SM instead (experiments).
        const int fixed = env_int("RL_TRACE_BLOCKS_PER_SM", 0);
        if (fixed > 0) {
            if (fixed < per_sm) full = (uint64_t)sm_count * fixed;
        } else {
            const int share_max = std::max(1, env_int("RL_TRACE_SHARE_MAX", 12));
            const int others = share_max > 1 ? other_streams_in_flight(dev, st, share_max - 1) : 0;
            full = (full + others) / (uint64_t)(others + 1);
            if (full < 1) full = 1;
        }
    }
    TraceArgs a;
    a.n_seg = p.n_segments;
    a.seg_shift = p.n_segments > 1 ? RL_SEGMENT_INDEX_BITS : 31u;
    a.ring_cap = ring_cap;
    a.width = (int)p.width;
    a.height = (int)p.height;
    a.aspect = (float)p.width / (float)p.height;      // trace_unit.rs:73, plot_unit.rs:49
    a.accum = p.accum;
    for (uint32_t k = 0; k < RL_MAX_SEGMENTS; k++) a.seg[k] = p.seg[k < p.n_segments ? k : 0];
    // a launch indexes the photons of a segment with seg_shift bits: a larger single request takes several launches
    const uint64_t chunk = 1ull << 31;
    const uint64_t passes = p.n_segments > 1 ? 1 : (n_photons + chunk - 1) / chunk;
    for (uint64_t pass = 0; pass < passes; pass++) {
        uint64_t n_launch = n_photons;
        if (p.n_segments == 1) {
            const uint64_t done = pass * chunk;
            n_launch = n_photons - done < chunk ? n_photons - done : chunk;
            a.seg[0].first_photon = p.seg[0].first_photon + done;
            a.seg[0].n_photons = n_launch;
            a.seg[0].records = p.seg[0].records ? p.seg[0].records + done : nullptr;
        }
        const uint64_t want = (n_launch + threads - 1) / threads;
        const unsigned grid = (unsigned)(want < full ? want : full);
        kernel<<<grid, threads, smem, st>>>(sc, a);
        g_launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (small) note_small_launch(dev, st);
    return cudaSuccess;
}

// ------------------------------------------------------------------ K2 splat
// Record stream through shared memory by bulk async copies (TMA, 1-D), warp-specialised: a
// producer warp keeps RL_SPLAT_STAGES copies of 1024 records (16 KB) in flight into a ring
// (full/empty mbarrier per stage); eight consumer warps each lift their 128-record slice of a
// stage into registers, hand the stage back at once and then compact and splat at their own pace.
// The bytes in flight therefore do not depend on how far the warps have got with the records they
// hold (with plain loads in registers the ballot chain and the occasional splat of one batch
// delayed the loads of the next: 59 % of the HBM peak on the record stream), and no warp waits
// for another warp's splat.
#ifndef RL_SPLAT_STAGES
#define RL_SPLAT_STAGES 4
#endif
#define RL_SPLAT_CHUNK 1024            // records per stage: 8 consumer warps x 4 rows x 32 lanes
#define RL_SPLAT_CONSUMERS 8           // warps
#define RL_SPLAT_THREADS (32 * (RL_SPLAT_CONSUMERS + 1))
#define RL_SPLAT_LIT_SLOTS 160         // per warp: up to 31 left over + 4 x 32 new

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "RL_MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra RL_MBAR_DONE;\n\t"
        "bra RL_MBAR_WAIT;\n\t"
        "RL_MBAR_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, bytes a multiple of 16, both addresses 16-byte aligned; streamed
// data: evict-first in L2 so the record stream does not push the accumulator out
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(RL_SPLAT_THREADS)
splat_kernel(const float4 *__restrict__ records, uint64_t n, float4 *accum, int width, int height,
             float aspect, uint32_t stages) {
    // Only ~8 % of the photons of the built-in scene carry light, so splatting in place would run
    // the splat code for two or three lanes of a warp at a time: contributing records are
    // compacted into a per-warp staging buffer with ballots and splatted 32 at a time with every
    // lane busy.  The four ballots of a round are independent of each other.
    extern __shared__ __align__(128) unsigned char splat_smem[];
    float4 *ring = reinterpret_cast<float4 *>(splat_smem);                       // [stages][RL_SPLAT_CHUNK]
    float4 *lit_all = ring + (size_t)stages * RL_SPLAT_CHUNK;                    // [consumers][RL_SPLAT_LIT_SLOTS]
    float4 *cie_tab = lit_all + RL_SPLAT_CONSUMERS * RL_SPLAT_LIT_SLOTS;         // [81 (+3 pad)] CIE table
    uint64_t *full = reinterpret_cast<uint64_t *>(cie_tab + 84);
    uint64_t *empty = full + stages;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, lanes_below = (1u << lane) - 1u;
    const uint64_t n_chunks = (n + RL_SPLAT_CHUNK - 1) / RL_SPLAT_CHUNK;
    if (tid == 0) {
        for (uint32_t s = 0; s < stages; s++) {
            mbar_init(full + s, 1);                       // the producer's arrive + the copy's bytes
            mbar_init(empty + s, RL_SPLAT_CONSUMERS);     // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fill_cie_shared(cie_tab);
    const CieShared cie{cie_tab};
    __syncthreads();

    if (warp == RL_SPLAT_CONSUMERS) {
        // producer: one lane refills a stage as soon as every consumer warp has lifted its slice
        if (lane == 0) {
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            uint32_t s = 0, parity = 0;
            for (uint64_t c = blockIdx.x, it = 0; c < n_chunks; c += gridDim.x, it++) {
                if (it >= stages) mbar_wait(empty + s, parity ^ 1u);     // the stage's previous use
                const uint64_t first = c * RL_SPLAT_CHUNK;
                const uint32_t count = (uint32_t)(n - first < RL_SPLAT_CHUNK ? n - first : RL_SPLAT_CHUNK);
                mbar_expect_tx(full + s, count * 16u);
                bulk_load(ring + (size_t)s * RL_SPLAT_CHUNK, records + first, count * 16u, full + s, policy);
                if (++s == stages) { s = 0; parity ^= 1u; }
            }
        }
        return;
    }

    float4 *mine = lit_all + warp * RL_SPLAT_LIT_SLOTS;
    uint32_t count = 0;                                                          // warp-uniform
    uint32_t s = 0, parity = 0;
    for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        mbar_wait(full + s, parity);
        const uint64_t first = c * RL_SPLAT_CHUNK;
        const uint32_t valid = (uint32_t)(n - first < RL_SPLAT_CHUNK ? n - first : RL_SPLAT_CHUNK);
        const float4 *src = ring + (size_t)s * RL_SPLAT_CHUNK + warp * 128u;
        float4 ph[4];
#pragma unroll
        for (int k = 0; k < 4; k++) ph[k] = src[k * 32 + lane];                  // {x, y, probability, wavelength}
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);                                   // slice lifted: the stage may be refilled
        if (++s == stages) { s = 0; parity ^= 1u; }
        uint32_t mask[4];
#pragma unroll
        for (int k = 0; k < 4; k++)      // adding cie * 0 changes nothing (plot_unit.rs:80-83)
            mask[k] = __ballot_sync(0xffffffffu, warp * 128u + k * 32 + lane < valid && ph[k].z != 0.0f);
        uint32_t at = count;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (mask[k] >> lane & 1u) mine[at + __popc(mask[k] & lanes_below)] = ph[k];
            at += __popc(mask[k]);
        }
        count = at;
        __syncwarp();
        while (count >= 32) {
            const float4 r = mine[count - 32 + lane];
            count -= 32;
            splat_photon_from(cie, accum, width, height, aspect, r.x, r.y, r.w, r.z);
        }
        __syncwarp();
    }
    if (lane < count) {
        const float4 r = mine[lane];
        splat_photon_from(cie, accum, width, height, aspect, r.x, r.y, r.w, r.z);
    }
}

cudaError_t launch_splat(const rl_mapped_photon *records, uint64_t n, float4 *accum, uint32_t width,
                         uint32_t height, int sm_count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    uint32_t stages = (uint32_t)env_int("RL_SPLAT_STAGES", RL_SPLAT_STAGES);
    if (stages < 1 || stages > 12) stages = RL_SPLAT_STAGES;
    const size_t smem = (size_t)stages * RL_SPLAT_CHUNK * sizeof(float4)
                        + (RL_SPLAT_CONSUMERS * RL_SPLAT_LIT_SLOTS + 84) * sizeof(float4) + 2 * stages * sizeof(uint64_t);
    static std::mutex attr_lock;
    int per_sm = 0;
    {
        std::lock_guard<std::mutex> guard(attr_lock);
        // what the attribute calls last set, per device (thousands of plot() calls per second)
        static size_t cached_smem[16] = {};
        static int cached_per_sm[16] = {};
        int dev = 0;
        cudaError_t err = cudaGetDevice(&dev);
        if (err != cudaSuccess) return err;
        if (dev < 0 || dev >= 16 || cached_smem[dev] != smem) {
            err = cudaFuncSetAttribute(splat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            // the splat runs beside trace blocks that hold the SM at its largest shared-memory
            // configuration: ask for the same one, so that placing a splat block never waits
            // for the SM to change its split
            const int carve = env_int("RL_SPLAT_CARVEOUT", 100);
            if (err == cudaSuccess && carve >= 0)
                err = cudaFuncSetAttribute(splat_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            if (err == cudaSuccess)
                err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, splat_kernel, RL_SPLAT_THREADS, smem);
            if (err != cudaSuccess) return err;
            if (dev >= 0 && dev < 16) { cached_smem[dev] = smem; cached_per_sm[dev] = per_sm; }
        } else {
            per_sm = cached_per_sm[dev];
        }
        if (per_sm < 1) per_sm = 1;
        // one block per SM: eight consumer warps keep up with the stream, and fewer reductions in
        // flight at once measured slightly faster than two or three blocks (tools/splat_probe.py)
        const int cap = env_int("RL_SPLAT_BLOCKS_PER_SM", 1);
        if (cap > 0 && cap < per_sm) per_sm = cap;
        const uint64_t want = (n + RL_SPLAT_CHUNK - 1) / RL_SPLAT_CHUNK;
        const uint64_t full = (uint64_t)sm_count * per_sm;
        const unsigned grid = (unsigned)(want < full ? want : full);
        splat_kernel<<<grid, RL_SPLAT_THREADS, smem, st>>>(reinterpret_cast<const float4 *>(records), n, accum,
                                                           (int)width, (int)height, (float)width / (float)height,
                                                           stages);
    }
    g_launches++;
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
pack_xyz_kernel(const float4 *__restrict__ accum, float *__restrict__ xyz, uint64_t n_pixels) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += stride) {
        const float4 v = accum[i];
        xyz[3 * i + 0] = v.x; xyz[3 * i + 1] = v.y; xyz[3 * i + 2] = v.z;
    }
}

cudaError_t launch_pack_xyz(const float4 *accum, float *xyz, uint64_t n_pixels, cudaStream_t st) {
    if (n_pixels == 0) return cudaSuccess;
    uint64_t want = (n_pixels + 255) / 256;
    unsigned grid = (unsigned)(want < 148 * 16 ? want : 148 * 16);
    pack_xyz_kernel<<<grid, 256, 0, st>>>(accum, xyz, n_pixels);
    g_launches++;
    return cudaGetLastError();
}

// ----------------------------------------------------------------- K3 gather
#define RL_GATHER_MAX_SRC 8
struct GatherSrcs { const float4 *p[RL_GATHER_MAX_SRC]; };

// gather_unit.rs:55-63, one component
__device__ __forceinline__ void kahan(float &acc, float &comp, float px) {
    const float extra = px - comp;
    const float sum = acc + extra;
    comp = (sum - acc) - extra;
    acc = sum;
}

// One thread owns 4 consecutive pixels = 12 floats = three 16-byte vectors of
// the packed accumulator / compensation arrays, so every access is 128-bit.
__global__ void __launch_bounds__(256)
gather_kernel(float *__restrict__ acc, float *__restrict__ comp, GatherSrcs srcs, uint32_t n_src,
              const float *__restrict__ packed_src, float4 *clear, uint64_t n_pixels) {
    const uint64_t n_quads = n_pixels / 4;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t q = tid; q < n_quads; q += stride) {
        float4 *acc4 = reinterpret_cast<float4 *>(acc) + 3 * q;
        float4 *comp4 = reinterpret_cast<float4 *>(comp) + 3 * q;
        float a[12], c[12];
        *reinterpret_cast<float4 *>(a + 0) = acc4[0];
        *reinterpret_cast<float4 *>(a + 4) = acc4[1];
        *reinterpret_cast<float4 *>(a + 8) = acc4[2];
        *reinterpret_cast<float4 *>(c + 0) = comp4[0];
        *reinterpret_cast<float4 *>(c + 4) = comp4[1];
        *reinterpret_cast<float4 *>(c + 8) = comp4[2];
        if (packed_src) {
            const float4 *s4 = reinterpret_cast<const float4 *>(packed_src) + 3 * q;
            float p[12];
            *reinterpret_cast<float4 *>(p + 0) = s4[0];
            *reinterpret_cast<float4 *>(p + 4) = s4[1];
            *reinterpret_cast<float4 *>(p + 8) = s4[2];
#pragma unroll
            for (int k = 0; k < 12; k++) kahan(a[k], c[k], p[k]);
        }
        for (uint32_t s = 0; s < n_src; s++) {
            const float4 *src = srcs.p[s] + 4 * q;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float4 v = src[k];
                kahan(a[3 * k + 0], c[3 * k + 0], v.x);
                kahan(a[3 * k + 1], c[3 * k + 1], v.y);
                kahan(a[3 * k + 2], c[3 * k + 2], v.z);
            }
        }
        acc4[0] = *reinterpret_cast<float4 *>(a + 0);
        acc4[1] = *reinterpret_cast<float4 *>(a + 4);
        acc4[2] = *reinterpret_cast<float4 *>(a + 8);
        comp4[0] = *reinterpret_cast<float4 *>(c + 0);
        comp4[1] = *reinterpret_cast<float4 *>(c + 4);
        comp4[2] = *reinterpret_cast<float4 *>(c + 8);
        if (clear) {  // PlotUnit::clear (plot_unit.rs:98-102)
            float4 *cl = clear + 4 * q;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            cl[0] = z; cl[1] = z; cl[2] = z; cl[3] = z;
        }
    }
    // up to 3 tail pixels
    for (uint64_t px = n_quads * 4 + tid; px < n_pixels; px += stride) {
        for (int k = 0; k < 3; k++) {
            float av = acc[3 * px + k], cv = comp[3 * px + k];
            if (packed_src) kahan(av, cv, packed_src[3 * px + k]);
            for (uint32_t s = 0; s < n_src; s++) {
                const float4 v = srcs.p[s][px];
                kahan(av, cv, k == 0 ? v.x : (k == 1 ? v.y : v.z));
            }
            acc[3 * px + k] = av; comp[3 * px + k] = cv;
        }
        if (clear) clear[px] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

cudaError_t launch_gather(float *acc, float *comp, const float4 *const *srcs, uint32_t n_src,
                          const float *packed_src, float4 *clear_or_null, uint64_t n_pixels,
                          int sm_count, cudaStream_t st) {
    if (n_pixels == 0) return cudaSuccess;
    if (n_src > RL_GATHER_MAX_SRC) return cudaErrorInvalidValue;
    GatherSrcs g;
    for (uint32_t i = 0; i < RL_GATHER_MAX_SRC; i++) g.p[i] = i < n_src ? srcs[i] : nullptr;
    uint64_t want = (n_pixels / 4 + 255) / 256 + 1;
    // many short blocks (one or two quads per thread) rather than a resident grid looping: measured
    // 0.219 ms against 0.245 ms at 4096^2, 94-96 % of a plain device copy of the same bytes
    // (tools/gather_probe.py)
    uint64_t full = (uint64_t)sm_count * (uint64_t)env_int("RL_GATHER_BLOCKS_PER_SM", 64);
    unsigned grid = (unsigned)(want < full ? want : full);
    gather_kernel<<<grid, 256, 0, st>>>(acc, comp, g, n_src, packed_src, clear_or_null, n_pixels);
    g_launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------- K4 tonemap
// tonemap_unit.rs:55-69.  The reference folds the two sums sequentially in
// f32; here they are reduced in f64 (order-independent to ~1e-16), then the
// reference's f32 formula is applied to the rounded means.
__global__ void __launch_bounds__(256)
tonemap_moments_kernel(const float *__restrict__ xyz, uint64_t n_pixels, double *moments) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double s1 = 0.0, s2 = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += stride) {
        const float y = xyz[3 * i + 1];
        s1 += (double)y;
        s2 += (double)(y * y);
    }
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    __shared__ double w1[8], w2[8];
    if ((threadIdx.x & 31) == 0) { w1[threadIdx.x >> 5] = s1; w2[threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int i = 0; i < 8; i++) { t1 += w1[i]; t2 += w2[i]; }
        atomicAdd(moments + 0, t1);
        atomicAdd(moments + 1, t2);
    }
}

// tonemap_unit.rs:55-69 to the letter: two sequential f32 folds over the pixels in row-major order
// (Rust's Iterator::sum), by ONE thread -- the block only stages the Y values in shared memory.
// 1 M pixels take a few milliseconds; the reference tone-maps once per 30 s
// (task_scheduler.rs:43-46).  Reproduces the reference's rounding and with it the accident that
// near-constant images get a NaN exposure (the variance rounds negative).
#define RL_FOLD_CHUNK 4096
__global__ void __launch_bounds__(256)
tonemap_fold_kernel(const float *__restrict__ xyz, uint64_t n_pixels, uint32_t width, uint32_t height, float *exposure) {
    __shared__ float ys[RL_FOLD_CHUNK];
    float sum = 0.0f, sqr = 0.0f;
    for (uint64_t first = 0; first < n_pixels; first += RL_FOLD_CHUNK) {
        const uint32_t count = (uint32_t)(n_pixels - first < RL_FOLD_CHUNK ? n_pixels - first : RL_FOLD_CHUNK);
        for (uint32_t k = threadIdx.x; k < count; k += blockDim.x) ys[k] = xyz[3 * (first + k) + 1];
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll 8
            for (uint32_t k = 0; k < count; k++) {
                const float y = ys[k];
                sum += y;                                // :61
                sqr += y * y;                            // :64
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float n = (float)(width * height);         // :56
        const float mean = sum / n;
        const float sqr_mean = sqr / n;
        const float variance = sqr_mean - mean * mean;   // :65
        *exposure = mean + sqrtf(variance);              // :68
    }
}

__global__ void tonemap_exposure_kernel(const double *moments, uint32_t width, uint32_t height,
                                        float *exposure) {
    const float n = (float)(width * height);               // tonemap_unit.rs:56
    const float mean = (float)moments[0] / n;              // :61
    const float sqr_mean = (float)moments[1] / n;          // :64
    const float variance = sqr_mean - mean * mean;         // :65
    *exposure = mean + sqrtf(variance);                    // :68
}

__device__ __forceinline__ float gamma_correct(float f) {  // srgb.rs:20-26
    if (f <= 0.0031308f) return 12.92f * f;
    return 1.055f * spec_pow(f, 1.0f / 2.4f) - 0.055f;
}
__device__ __forceinline__ float clamp01(float x) {        // tonemap_unit.rs:34-38
    if (x < 0.0f) return 0.0f;
    if (1.0f < x) return 1.0f;
    return x;
}
__device__ __forceinline__ uint32_t to_u8(float v) {       // Rust `as u8`: saturating, NaN -> 0
    if (!(v == v)) return 0u;
    if (v <= 0.0f) return 0u;
    if (v >= 255.0f) return 255u;
    return (uint32_t)v;
}

// tonemap_unit.rs:73-100 + srgb.rs:29-41, one pixel per thread-iteration.
__global__ void __launch_bounds__(256)
tonemap_map_kernel(const float *__restrict__ xyz, uint64_t n_pixels, const float *exposure,
                   uint8_t *__restrict__ rgb) {
    const float max_intensity = *exposure;
    const float ln_4 = spec_ln(4.0f);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += stride) {
        const float cx = spec_ln(xyz[3 * i + 0] / max_intensity + 1.0f) / ln_4;
        const float cy = spec_ln(xyz[3 * i + 1] / max_intensity + 1.0f) / ln_4;
        const float cz = spec_ln(xyz[3 * i + 2] / max_intensity + 1.0f) / ln_4;
        const float r = 3.2406f * cx - 1.5372f * cy - 0.4986f * cz;
        const float g = -0.9689f * cx + 1.8758f * cy + 0.0415f * cz;
        const float b = 0.0557f * cx - 0.2040f * cy + 1.0570f * cz;
        rgb[3 * i + 0] = (uint8_t)to_u8(clamp01(gamma_correct(r)) * 255.0f);
        rgb[3 * i + 1] = (uint8_t)to_u8(clamp01(gamma_correct(g)) * 255.0f);
        rgb[3 * i + 2] = (uint8_t)to_u8(clamp01(gamma_correct(b)) * 255.0f);
    }
}

cudaError_t launch_tonemap(const float *xyz, uint32_t width, uint32_t height, double *moments,
                           float *exposure_out, uint8_t *rgb, int sm_count, bool reference_fold, cudaStream_t st) {
    const uint64_t n_pixels = (uint64_t)width * height;
    if (n_pixels == 0) return cudaSuccess;
    uint64_t want = (n_pixels + 255) / 256;
    uint64_t full = (uint64_t)sm_count * 8;
    unsigned grid = (unsigned)(want < full ? want : full);
    if (reference_fold) {
        tonemap_fold_kernel<<<1, 256, 0, st>>>(xyz, n_pixels, width, height, exposure_out);
        g_launches += 1;
    } else {
        cudaError_t err = cudaMemsetAsync(moments, 0, 2 * sizeof(double), st);
        if (err != cudaSuccess) return err;
        tonemap_moments_kernel<<<grid, 256, 0, st>>>(xyz, n_pixels, moments);
        tonemap_exposure_kernel<<<1, 1, 0, st>>>(moments, width, height, exposure_out);
        g_launches += 2;
    }
    tonemap_map_kernel<<<grid, 256, 0, st>>>(xyz, n_pixels, exposure_out, rgb);
    g_launches += 1;
    return cudaGetLastError();
}

// -------------------------------------------------------------------- probes
__global__ void debug_intersect_kernel(const DevScene sc, const rl_ray *rays, uint64_t n, rl_hit *out) {
    setup_tables(sc);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    // warp-uniform trip count: intersect_scene votes across the whole warp
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += stride) {
        const uint64_t i = base + threadIdx.x;
        const bool live = i < n;
        Ray r = idle_ray();
        if (live) {
            r.origin = mk(rays[i].origin.x, rays[i].origin.y, rays[i].origin.z);
            r.direction = mk(rays[i].direction.x, rays[i].direction.y, rays[i].direction.z);
            r.wavelength = rays[i].wavelength;
        }
        const Hit h = intersect_scene_any(r);
        // intersect_scene leaves its block-wide scratch (task counter, result slots) to be reset
        // behind a barrier that the next call must not overtake
        __syncthreads();
        if (!live) continue;
        rl_hit o;
        o.object = h.obj;
        o.distance = 0.f;
        o.position = o.normal = o.tangent = rl_vec3{0.f, 0.f, 0.f};
        if (h.obj >= 0) {
            Surf s = surface_at(r, h);
            s.tangent = sphere_tangent(h, s);
            o.distance = h.t;
            o.position = rl_vec3{s.position.x, s.position.y, s.position.z};
            o.normal = rl_vec3{s.normal.x, s.normal.y, s.normal.z};
            o.tangent = rl_vec3{s.tangent.x, s.tangent.y, s.tangent.z};
        }
        out[i] = o;
    }
}

cudaError_t launch_debug_intersect(const DevScene &sc, const rl_ray *rays, uint64_t n, rl_hit *out,
                                   cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const size_t smem = trace_smem_bytes(sc, 128);
    cudaError_t err = cudaFuncSetAttribute(debug_intersect_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    uint64_t want = (n + 127) / 128;
    unsigned grid = (unsigned)(want < 148 * 4 ? want : 148 * 4);
    debug_intersect_kernel<<<grid, 128, smem, st>>>(sc, rays, n, out);
    g_launches++;
    return cudaGetLastError();
}

// Traces paths with the brute-force Scene::intersect as the driver and
// evaluates the culled one on every ray beside it; counts disagreements.
__global__ void __launch_bounds__(128)
debug_cull_check_kernel(const DevScene sc, uint64_t seed, float aspect, uint64_t first, uint64_t n,
                        unsigned long long *rays_out, unsigned long long *mismatches) {
    setup_tables(sc);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t next = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long rays = 0, bad = 0;
    bool alive = false;
    Ray ray = idle_ray();
    Rng rng;
    rng.init();
    RngKey key = {seed, first};
    float intensity = 1.0f, continue_chance = 1.0f;
    uint32_t bounce = 0;
    for (;;) {
        if (!alive && next < n) {
            key.photon = first + next;
            rng.init();
            next += stride;
            const float wavelength = rng.wavelength(key);
            const float x = rng.bi_unit(key);
            const float y = rng.bi_unit(key) / aspect;
            const float t = rng.unit(key);
            ray = camera_ray(sc.camera, x, y, wavelength, t, rng, key);
            intensity = 1.0f;
            continue_chance = 1.0f;
            bounce = 0;
            alive = true;
        }
        if (!__syncthreads_or(alive)) break;
        const Ray r = alive ? ray : idle_ray();
        const Hit culled = intersect_scene_any(r);
        if (!alive) continue;
        const Hit hit = intersect_scene_brute(r);
        BounceRng brng;
        brng.load(key, bounce);
        bounce++;
        rays++;
        if (hit.obj != culled.obj || __float_as_uint(hit.t) != __float_as_uint(culled.t) ||
            hit.code != culled.code)
            bad++;
        alive = false;
        if (hit.obj < 0) continue;
        const float4 m = __ldg(sc.materials + hit.obj);
        if (__float_as_uint(m.x) == RL_MATERIAL_BLACKBODY) continue;
        const Surf s = surface_at(ray, hit);
        float probability;
        const V3 dir = material_bounce(m, ray, hit, s, brng, probability);
        intensity = intensity * probability;
        ray.direction = dir;
        ray.origin = s.position + dir * 0.00001f;
        continue_chance = continue_chance * 0.96f;
        if (brng.unit() * 0.85f > continue_chance * (1.0f - spec_exp(intensity * -20.0f))) continue;
        alive = true;
    }
    atomicAdd(rays_out, rays);
    if (bad) atomicAdd(mismatches, bad);
}

cudaError_t launch_debug_cull_check(const DevScene &sc, uint64_t seed, uint32_t width, uint32_t height,
                                    uint64_t first, uint64_t n, unsigned long long *rays,
                                    unsigned long long *mismatches, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const size_t smem = trace_smem_bytes(sc, 128);
    cudaError_t err = cudaFuncSetAttribute(debug_cull_check_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    uint64_t want = (n + 127) / 128;
    unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);
    debug_cull_check_kernel<<<grid, 128, smem, st>>>(sc, seed, (float)width / (float)height, first, n,
                                                     rays, mismatches);
    g_launches++;
    return cudaGetLastError();
}

__global__ void debug_math_kernel(int fn, const float *in, const float *in2, uint64_t n, float *out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float x = in[i];
        float s, c, r = 0.f;
        switch (fn) {
        case 0: spec_sincos(x, s, c); r = s; break;
        case 1: spec_sincos(x, s, c); r = c; break;
        case 2: r = spec_exp(x); break;
        case 3: r = spec_acos(x); break;
        case 4: r = (float)boltzmann((double)x, (double)in2[i]); break;
        case 5: r = sf10_index_of_refraction(x); break;
        case 6: r = spec_ln(x); break;
        case 7: r = spec_pow(x, in2[i]); break;
        case 8: r = spec_tan(x); break;
        }
        out[i] = r;
    }
}

cudaError_t launch_debug_math(int fn, const float *in, const float *in2, uint64_t n, float *out,
                              cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    uint64_t want = (n + 255) / 256;
    unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);
    debug_math_kernel<<<grid, 256, 0, st>>>(fn, in, in2, n, out);
    g_launches++;
    return cudaGetLastError();
}

__global__ void debug_tristimulus_kernel(const float *wl, uint64_t n, float *out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const V3 t = tristimulus(wl[i]);
        out[3 * i] = t.x; out[3 * i + 1] = t.y; out[3 * i + 2] = t.z;
    }
}

cudaError_t launch_debug_tristimulus(const float *wl, uint64_t n, float *out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    uint64_t want = (n + 255) / 256;
    unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);
    debug_tristimulus_kernel<<<grid, 256, 0, st>>>(wl, n, out);
    g_launches++;
    return cudaGetLastError();
}

__global__ void debug_camera_kernel(const DevScene sc, uint64_t seed, float aspect, uint64_t first,
                                    uint64_t n, rl_ray *rays, rl_mapped_photon *xy) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Rng rng;
        rng.init();
        const RngKey key = {seed, first + i};
        const float wavelength = rng.wavelength(key);
        const float x = rng.bi_unit(key);
        const float y = rng.bi_unit(key) / aspect;
        const float t = rng.unit(key);
        const Ray r = camera_ray(sc.camera, x, y, wavelength, t, rng, key);
        rl_ray o;
        o.origin = rl_vec3{r.origin.x, r.origin.y, r.origin.z};
        o.direction = rl_vec3{r.direction.x, r.direction.y, r.direction.z};
        o.wavelength = wavelength;
        o.probability = 1.0f;
        rays[i] = o;
        if (xy) { xy[i].x = x; xy[i].y = y; xy[i].wavelength = wavelength; xy[i].probability = t; }
    }
}

cudaError_t launch_debug_camera(const DevScene &sc, uint64_t seed, uint32_t width, uint32_t height,
                                uint64_t first, uint64_t n, rl_ray *rays, rl_mapped_photon *xy,
                                cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    uint64_t want = (n + 127) / 128;
    unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);
    debug_camera_kernel<<<grid, 128, 0, st>>>(sc, seed, (float)width / (float)height, first, n, rays, xy);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace rl
