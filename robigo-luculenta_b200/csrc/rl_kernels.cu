// rl_kernels.cu -- the sm_100a kernels of the path.
//
//   K1 trace_kernel    TraceUnit::render (trace_unit.rs:81-168), optionally
//                      fused with PlotUnit::plot (plot_unit.rs:56-95)
//   K2 splat_kernel    PlotUnit::plot over MappedPhoton records
//   K3 gather_kernel   GatherUnit::accumulate (gather_unit.rs:49-64) + PlotUnit::clear
//   K4 tonemap_*       TonemapUnit::tonemap (tonemap_unit.rs:55-100, srgb.rs:20-41)
//
// Compile with -fmad=false: see rl_math.cuh.
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
#ifndef RL_SORT_PATHS
#define RL_SORT_PATHS 0          // 1: re-deal the paths of a CTA by next action every bounce (see rl_device.cuh)
#endif
#ifndef RL_TRACE_MIN_BLOCKS
#define RL_TRACE_MIN_BLOCKS 1
#endif
// launches with fewer than RL_TRACE_SMALL_PATHS photons per thread of a full grid use CTAs of
// RL_TRACE_SMALL_CTA threads (0 paths: never); both can be overridden from the environment
#ifndef RL_TRACE_SMALL_CTA
#define RL_TRACE_SMALL_CTA 256
#endif
#ifndef RL_TRACE_SMALL_PATHS
#define RL_TRACE_SMALL_PATHS 64
#endif

// ------------------------------------------------------------------ K1 trace
struct TraceArgs {
    uint64_t seed;
    uint64_t first_photon;
    uint64_t n_photons;
    int width, height;
    float aspect;
    rl_mapped_photon *records;
    float4 *accum;
    unsigned long long *ray_counter;
};

// Persistent threads with path regeneration: a block owns a contiguous range
// of the launch's photons and hands them out from a counter in shared memory
// (one warp-aggregated atomic per warp and iteration); a lane takes its next
// photon the moment its current path ends, so a warp never idles on its longest
// path and no lane idles while the block's pool is not empty -- with the
// reference's 524 288-photon batches a thread sees only a handful of paths and
// a static deal would leave most lanes waiting for the unluckiest one.  The
// result of a photon depends on its id alone, so the deal changes nothing but
// the order of the accumulator atomics.  The primitive tables live in shared
// memory; each loop iteration is one Scene::intersect plus one material
// interaction for every live lane.
__global__ void __launch_bounds__(RL_TRACE_THREADS, RL_TRACE_MIN_BLOCKS)
trace_kernel(const DevScene sc, const TraceArgs a) {
    setup_tables(sc);

    // photon indices within the launch fit 32 bits (launch_trace splits larger requests)
    const uint32_t pool_end = (uint32_t)(a.n_photons * (blockIdx.x + 1ull) / gridDim.x);
    uint32_t *pool = photon_pool();
    if (threadIdx.x == 0) *pool = (uint32_t)(a.n_photons * (uint64_t)blockIdx.x / gridDim.x);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    bool pool_empty = false;

    bool alive = false;
    uint32_t cur = 0;
    Ray ray;
    ray.origin = mk(0.f, 0.f, 0.f); ray.direction = mk(0.f, 0.f, 0.f); ray.wavelength = 0.f;
    float sx = 0.f, sy = 0.f;
    float intensity = 1.0f, continue_chance = 1.0f;
    Rng rng;
    rng.init();
    uint32_t rays = 0;
#if RL_SORT_PATHS
    // staging of the path exchange and the permutation table, behind the intersection scratch
    uint32_t *stage = reinterpret_cast<uint32_t *>(
        reinterpret_cast<char *>(rl_smem + tables().scratch) + (size_t)RL_SCRATCH_BYTES_PER_THREAD * blockDim.x);
    uint16_t *perm = reinterpret_cast<uint16_t *>(stage + RL_PATH_WORDS * blockDim.x);
    uint32_t parity = 0;
#endif

    for (;;) {
        // lanes without a path draw the next photon ids of the block's pool; a warp that has seen
        // the pool empty never touches the counter again, so it overshoots the end by at most
        // 32 per warp and cannot wrap
        uint32_t mine = 0xffffffffu;
        if (!pool_empty) {
            const uint32_t want = __ballot_sync(0xffffffffu, !alive);
            if (want != 0u) {
                const uint32_t leader = __ffs(want) - 1u;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(pool, (uint32_t)__popc(want));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (!alive) mine = base + __popc(want & ((1u << lane) - 1u));
                pool_empty = base + (uint32_t)__popc(want) >= pool_end;      // warp-uniform
            }
        }
        if (mine < pool_end) {
            // trace_unit.rs:151-158 and :136-145
            cur = mine;
            const RngKey key = {a.seed, a.first_photon + cur};
            rng.init();
            const float wavelength = rng.wavelength(key);
            sx = rng.bi_unit(key);
            sy = rng.bi_unit(key) / a.aspect;
            const float t = rng.unit(key);
            ray = camera_ray(sc.camera, sx, sy, wavelength, t, rng, key);
            intensity = 1.0f;
            continue_chance = 1.0f;
            alive = true;
        }
        // every thread of the block takes part in the intersection (block barriers and warp
        // votes inside); idle lanes trace a ray that hits nothing
        if (!__syncthreads_or(alive)) break;
        Hit hit = intersect_scene(alive ? ray : idle_ray(), alive);
        if (alive) rays++;                                              // Scene::intersect calls (scene.rs:39)
#if RL_SORT_PATHS
        {
            uint32_t cls = RL_CLASS_IDLE;
            if (alive) {
                cls = RL_CLASS_MISS;
                if (hit.obj >= 0) {
                    const uint32_t kind = __float_as_uint(__ldg(sc.materials + hit.obj).x);
                    cls = kind == RL_MATERIAL_BLACKBODY ? RL_CLASS_EMITTER
                          : (kind <= RL_MATERIAL_GLOSSY_MIRROR ? RL_CLASS_DIFFUSE
                             : (kind == RL_MATERIAL_SF10_GLASS ? RL_CLASS_GLASS : RL_CLASS_SOAP));
                }
            }
            path_sort_publish(cls, parity, perm);
            const uint32_t T = blockDim.x;
            uint32_t *p = stage + threadIdx.x;
            p[0 * T] = __float_as_uint(ray.origin.x); p[1 * T] = __float_as_uint(ray.origin.y);
            p[2 * T] = __float_as_uint(ray.origin.z); p[3 * T] = __float_as_uint(ray.direction.x);
            p[4 * T] = __float_as_uint(ray.direction.y); p[5 * T] = __float_as_uint(ray.direction.z);
            p[6 * T] = __float_as_uint(ray.wavelength); p[7 * T] = __float_as_uint(intensity);
            p[8 * T] = __float_as_uint(continue_chance); p[9 * T] = __float_as_uint(sx);
            p[10 * T] = __float_as_uint(sy); p[11 * T] = cur;
            p[12 * T] = (rng.block << 3) | rng.left;
            p[13 * T] = rng.b0; p[14 * T] = rng.b1; p[15 * T] = rng.b2;
            p[16 * T] = __float_as_uint(hit.t); p[17 * T] = (uint32_t)hit.obj; p[18 * T] = hit.code;
            __syncthreads();
            p = stage + path_sort_fetch(parity, perm, cls);
            parity ^= 1u;
            ray.origin = mk(__uint_as_float(p[0 * T]), __uint_as_float(p[1 * T]), __uint_as_float(p[2 * T]));
            ray.direction = mk(__uint_as_float(p[3 * T]), __uint_as_float(p[4 * T]), __uint_as_float(p[5 * T]));
            ray.wavelength = __uint_as_float(p[6 * T]); intensity = __uint_as_float(p[7 * T]);
            continue_chance = __uint_as_float(p[8 * T]); sx = __uint_as_float(p[9 * T]);
            sy = __uint_as_float(p[10 * T]); cur = p[11 * T];
            rng.block = p[12 * T] >> 3; rng.left = p[12 * T] & 7u;
            rng.b0 = p[13 * T]; rng.b1 = p[14 * T]; rng.b2 = p[15 * T];
            hit.t = __uint_as_float(p[16 * T]); hit.obj = (int)p[17 * T]; hit.code = p[18 * T];
            alive = cls != RL_CLASS_IDLE;
            // the staging area is written again only behind the next loop-top barrier
        }
#endif
        if (alive) {
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
                    const RngKey key = {a.seed, a.first_photon + cur};
                    float probability;
                    const V3 dir = material_bounce(m, ray, hit, s, rng, key, probability);  // :104-107
                    intensity = intensity * probability;
                    ray.direction = dir;
                    ray.origin = s.position + dir * 0.00001f;                      // :114
                    continue_chance = continue_chance * 0.96f;                    // :117
                    if (rng.unit(key) * 0.85f
                        > continue_chance * (1.0f - spec_exp(intensity * -20.0f)))  // :122-125
                        done = true;
                }
            }
            if (done) {
                if (a.records) {
                    rl_mapped_photon ph;
                    ph.x = sx; ph.y = sy; ph.probability = result; ph.wavelength = ray.wavelength;
                    *reinterpret_cast<float4 *>(a.records + cur) =
                        make_float4(ph.x, ph.y, ph.probability, ph.wavelength);
                }
                // adding cie * 0 leaves the accumulator unchanged (plot_unit.rs:80-83)
                if (a.accum && result != 0.0f)
                    splat_photon(a.accum, a.width, a.height, a.aspect, sx, sy, ray.wavelength, result);
                alive = false;
            }
        }
    }

    // rays traced = Scene::intersect calls (scene.rs:39)
    for (int o = 16; o > 0; o >>= 1) rays += __shfl_xor_sync(0xffffffffu, rays, o);
    if ((threadIdx.x & 31) == 0 && a.ray_counter) atomicAdd(a.ray_counter, (unsigned long long)rays);
}

// --------------------------------------------------------- K1b trace service
// The reference's host hands the engine 524 288-photon batches from C worker threads
// (task_scheduler.rs:95-96,127-182; app.rs:104-109): 4.6 photons per thread of a full grid.  One
// launch per batch spends most of its life in its tail.  The trace service keeps ONE resident
// kernel busy with whatever batches are queued: render() pushes a {photon range, records,
// completion flag} entry into a ring in device memory (service_push_kernel, in stream order on
// the unit's stream) and launches a worker (service_worker_kernel) behind it; a worker's CTAs
// claim chunks of photons from the oldest entries of the ring, whichever unit they belong to,
// and leave only when the ring has nothing left to claim -- so while the host keeps batches
// queued the same CTAs stay resident and go from one batch to the next without a tail, and a
// worker that finds the ring empty (an older worker took its entry) retires in microseconds.
// A finished entry stores its sequence number to the unit's completion word; the unit's stream
// waits on that word (cuStreamWaitValue32) before the copy of the records to the host.
//
// Same path code as trace_kernel (camera_ray, intersect_scene, material_bounce, roulette): a
// photon's record depends on (scene, seed, photon id) only, so which CTA of which worker traces
// it changes nothing (tests/test_gpu_service.py).
#define RL_SERVICE_CHUNK 1024u         // photons a CTA claims at a time
#define RL_SERVICE_OPEN 4u             // entries a CTA can hold paths of at once
#define RL_SERVICE_POOLS 2u            // chunks a CTA deals photons from
#ifndef RL_SERVICE_LINGER_CYCLES
#define RL_SERVICE_LINGER_CYCLES 60000 // a CTA that has worked waits this long for new entries (~30 us)
#endif

struct ServiceCta {                    // per-CTA bookkeeping in shared memory, behind the intersection scratch
    struct Open {
        uint64_t seed, first_photon;
        rl_mapped_photon *records;
        float4 *accum;
        ServiceEntry *slot;
        float aspect;
        int width, height;
        uint32_t ticket;               // ring ticket + 1 of the entry; 0 = free
        uint32_t n_photons;
        uint32_t unfinished;           // photons this CTA claimed and has not finished (dealt or not)
        uint32_t claimed;              // photons claimed since the last flush to the entry
        uint32_t rays;                 // Scene::intersect calls since the last flush
    } open[RL_SERVICE_OPEN];
    struct Pool { uint32_t next, end, open; } pool[RL_SERVICE_POOLS];
    uint32_t active;                   // some photon of this CTA is alive or waits to be dealt
    uint32_t done;                     // nothing left and nothing arrived: the CTA retires
    uint32_t worked;                   // the CTA has claimed at least one chunk
    long long idle_since;
};

__device__ __forceinline__ ServiceCta *service_cta(const DevScene &sc) {
    char *base = reinterpret_cast<char *>(rl_smem + RL_TABLES_VEC4 + sc.smem_vec4);
    return reinterpret_cast<ServiceCta *>(base + (size_t)RL_SCRATCH_BYTES_PER_THREAD * blockDim.x);
}

__global__ void service_push_kernel(ServiceQueue *q, ServiceEntry e) {
    if (threadIdx.x != 0) return;
    const uint32_t t = atomicAdd(&q->tail, 1u);
    ServiceEntry *s = &q->slots[t % RL_SERVICE_CAP];
    // the slot's previous entry (RL_SERVICE_CAP tickets ago) may still be in flight
    while (atomicCAS(&s->busy, 0u, 1u) != 0u) __nanosleep(200);
    s->seed = e.seed; s->first_photon = e.first_photon;
    s->records = e.records; s->accum = e.accum; s->ray_counter = e.ray_counter;
    s->done_flag = e.done_flag; s->done_value = e.done_value; s->n_photons = e.n_photons;
    s->width = e.width; s->height = e.height; s->aspect = e.aspect;
    s->finished = 0u;
    __threadfence();
    // publishes the entry: claimers compare the ticket half before they touch the slot
    atomicExch(&s->ticket_next, (unsigned long long)(t + 1u) << 32);
}

// Thread 0, between two block barriers: hand finished photons back to their entries, refill the
// pools from the ring, decide whether the CTA stays.
__device__ __forceinline__ void service_manage(ServiceQueue *q, ServiceCta *cta) {
    uint32_t pending = 0;
#pragma unroll 1
    for (uint32_t o = 0; o < RL_SERVICE_OPEN; o++) {
        ServiceCta::Open &op = cta->open[o];
        if (op.ticket == 0u) continue;
        if (op.unfinished != 0u) { pending += op.unfinished; continue; }
        // every photon this CTA took from the entry has ended; the records were written before the
        // barrier this thread has just passed
        ServiceEntry *s = op.slot;
        if (op.rays && s->ray_counter) atomicAdd(s->ray_counter, (unsigned long long)op.rays);
        __threadfence();
        const uint32_t before = atomicAdd(&s->finished, op.claimed);
        if (before + op.claimed == op.n_photons) {
            // the last photons of the entry: everything the other CTAs wrote is ordered before
            // their own additions to `finished`
            // the flag is a word in mapped host memory: the records must be visible to the copy
            // engine and to peers before the host sees it
            __threadfence_system();
            uint32_t *flag = s->done_flag;
            const uint32_t value = s->done_value;
            *reinterpret_cast<volatile uint32_t *>(flag) = value;
            __threadfence();
            atomicExch(&s->busy, 0u);
        }
        op.ticket = 0u; op.claimed = 0u; op.rays = 0u;
    }
#pragma unroll 1
    for (uint32_t p = 0; p < RL_SERVICE_POOLS; p++) {
        ServiceCta::Pool &pl = cta->pool[p];
        if (pl.next < pl.end) continue;                         // still dealing
        pl.next = pl.end = 0u;
        // claim the next chunk of the oldest entry that has photons left
        uint32_t h = *reinterpret_cast<volatile uint32_t *>(&q->head);
#pragma unroll 1
        for (;;) {
            const uint32_t t = *reinterpret_cast<volatile uint32_t *>(&q->tail);
            if (h == t) break;
            ServiceEntry *s = &q->slots[h % RL_SERVICE_CAP];
            unsigned long long word = *reinterpret_cast<volatile unsigned long long *>(&s->ticket_next);
            if ((uint32_t)(word >> 32) != h + 1u) {
                // not published yet, or the ring has moved on
                const uint32_t h2 = *reinterpret_cast<volatile uint32_t *>(&q->head);
                if (h2 == h) break;
                h = h2;
                continue;
            }
            __threadfence();                                    // the entry's fields were written before its ticket
            const uint32_t n = *reinterpret_cast<volatile uint32_t *>(&s->n_photons);
            const uint32_t c = (uint32_t)word;
            if (c >= n) {                                       // fully claimed: move the head on
                atomicCAS(&q->head, h, h + 1u);
                h = *reinterpret_cast<volatile uint32_t *>(&q->head);
                continue;
            }
            // an open record for this entry: the one it already has, else a free one
            uint32_t o = RL_SERVICE_OPEN, free_o = RL_SERVICE_OPEN;
            for (uint32_t k = 0; k < RL_SERVICE_OPEN; k++) {
                if (cta->open[k].ticket == h + 1u) o = k;
                else if (cta->open[k].ticket == 0u && free_o == RL_SERVICE_OPEN) free_o = k;
            }
            if (o == RL_SERVICE_OPEN) o = free_o;
            if (o == RL_SERVICE_OPEN) break;                    // paths of four entries alive: wait
            const uint32_t take = n - c < RL_SERVICE_CHUNK ? n - c : RL_SERVICE_CHUNK;
            if (atomicCAS(&s->ticket_next, word, word + take) != word) continue;   // another CTA was faster
            ServiceCta::Open &op = cta->open[o];
            if (op.ticket == 0u) {
                op.seed = s->seed; op.first_photon = s->first_photon;
                op.records = s->records; op.accum = s->accum; op.slot = s;
                op.aspect = s->aspect; op.width = (int)s->width; op.height = (int)s->height;
                op.n_photons = n;
                op.ticket = h + 1u; op.unfinished = 0u; op.claimed = 0u; op.rays = 0u;
            }
            op.unfinished += take;
            op.claimed += take;
            pl.next = c; pl.end = c + take; pl.open = o;
            pending += take;
            cta->worked = 1u;
            break;
        }
    }
    cta->active = pending != 0u;
    if (pending != 0u) {
        cta->idle_since = 0;
    } else if (!cta->worked) {
        cta->done = 1u;                                         // an older worker took everything
    } else {
        const long long now = clock64();
        if (cta->idle_since == 0) cta->idle_since = now;
        else if (now - cta->idle_since > RL_SERVICE_LINGER_CYCLES) cta->done = 1u;
        __nanosleep(500);
    }
}

__global__ void __launch_bounds__(RL_TRACE_THREADS, RL_TRACE_MIN_BLOCKS)
service_worker_kernel(const DevScene sc, ServiceQueue *q) {
    ServiceCta *cta = service_cta(sc);
    if (threadIdx.x == 0) {
        for (uint32_t o = 0; o < RL_SERVICE_OPEN; o++) {
            cta->open[o].ticket = 0u; cta->open[o].unfinished = 0u; cta->open[o].claimed = 0u; cta->open[o].rays = 0u;
        }
        for (uint32_t p = 0; p < RL_SERVICE_POOLS; p++) cta->pool[p].next = cta->pool[p].end = cta->pool[p].open = 0u;
        cta->active = 0u; cta->done = 0u; cta->worked = 0u; cta->idle_since = 0;
        service_manage(q, cta);
    }
    __syncthreads();
    if (cta->done) return;                                      // nothing to claim: no table set-up either
    setup_tables(sc);

    const uint32_t lane = threadIdx.x & 31u;
    bool alive = false;
    uint32_t cur = 0;                                           // photon index within its entry | open record << 28
    Ray ray;
    ray.origin = mk(0.f, 0.f, 0.f); ray.direction = mk(0.f, 0.f, 0.f); ray.wavelength = 0.f;
    float sx = 0.f, sy = 0.f;
    float intensity = 1.0f, continue_chance = 1.0f;
    Rng rng;
    rng.init();
    uint32_t rays = 0;                                          // of the current path

    for (bool first = true;; first = false) {
        if (!first) {
            __syncthreads();                                    // the path ends of the last iteration are counted
            if (threadIdx.x == 0) service_manage(q, cta);
            __syncthreads();
        }
        if (cta->done) break;
        if (!cta->active) continue;
        // lanes without a path draw photons from the CTA's pools (one warp-aggregated atomic per pool)
        uint32_t mine = 0xffffffffu, mine_open = 0;
#pragma unroll
        for (uint32_t p = 0; p < RL_SERVICE_POOLS; p++) {
            const uint32_t want = __ballot_sync(0xffffffffu, !alive && mine == 0xffffffffu);
            ServiceCta::Pool &pl = cta->pool[p];
            const uint32_t end = pl.end;
            if (want != 0u && pl.next < end) {                  // warp-uniform: one address, one load
                const uint32_t leader = __ffs(want) - 1u;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(&pl.next, (uint32_t)__popc(want));
                base = __shfl_sync(0xffffffffu, base, leader);
                const uint32_t idx = base + __popc(want & ((1u << lane) - 1u));
                if (!alive && mine == 0xffffffffu && idx < end) { mine = idx; mine_open = pl.open; }
            }
        }
        if (mine != 0xffffffffu) {
            // trace_unit.rs:151-158 and :136-145
            const ServiceCta::Open &op = cta->open[mine_open];
            cur = mine | (mine_open << 28);
            const RngKey key = {op.seed, op.first_photon + mine};
            rng.init();
            const float wavelength = rng.wavelength(key);
            sx = rng.bi_unit(key);
            sy = rng.bi_unit(key) / op.aspect;
            const float t = rng.unit(key);
            ray = camera_ray(sc.camera, sx, sy, wavelength, t, rng, key);
            intensity = 1.0f;
            continue_chance = 1.0f;
            rays = 0;
            alive = true;
        }
        const Hit hit = intersect_scene(alive ? ray : idle_ray(), alive);
        if (alive) {
            rays++;
            ServiceCta::Open &op = cta->open[cur >> 28];
            const uint32_t index = cur & 0x0fffffffu;
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
                    const RngKey key = {op.seed, op.first_photon + index};
                    float probability;
                    const V3 dir = material_bounce(m, ray, hit, s, rng, key, probability);  // :104-107
                    intensity = intensity * probability;
                    ray.direction = dir;
                    ray.origin = s.position + dir * 0.00001f;                      // :114
                    continue_chance = continue_chance * 0.96f;                    // :117
                    if (rng.unit(key) * 0.85f
                        > continue_chance * (1.0f - spec_exp(intensity * -20.0f)))  // :122-125
                        done = true;
                }
            }
            if (done) {
                if (op.records)
                    *reinterpret_cast<float4 *>(op.records + index) = make_float4(sx, sy, result, ray.wavelength);
                if (op.accum && result != 0.0f)
                    splat_photon(op.accum, op.width, op.height, op.aspect, sx, sy, ray.wavelength, result);
                atomicAdd(&op.rays, rays);
                atomicSub(&op.unfinished, 1u);
                alive = false;
            }
        }
    }
}

size_t trace_smem_bytes(const DevScene &sc, int threads) { return tracing_smem_bytes(sc, threads); }
// the trace kernel proper: plus the path-sort area when enabled
static size_t service_smem_bytes(const DevScene &sc, int threads) {
    return tracing_smem_bytes(sc, threads) + sizeof(ServiceCta);
}
static size_t trace_kernel_smem_bytes(const DevScene &sc, int threads) {
    return tracing_smem_bytes(sc, threads) + (RL_SORT_PATHS ? (size_t)RL_SORT_BYTES_PER_THREAD * threads : 0);
}

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

cudaError_t launch_service_push(ServiceQueue *q, const ServiceEntry &e, cudaStream_t st) {
    if (e.n_photons == 0 || e.n_photons > (1u << 28)) return cudaErrorInvalidValue;
    service_push_kernel<<<1, 32, 0, st>>>(q, e);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_service_worker(const DevScene &sc, ServiceQueue *q, int sm_count, int reserved_sms,
                                  cudaStream_t st) {
    static std::mutex lock;
    static KernelCache cache[16];
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    int threads = RL_TRACE_THREADS;
    size_t smem = 0;
    {
        std::lock_guard<std::mutex> guard(lock);
        KernelCache scratch_entry;
        KernelCache &cached = dev >= 0 && dev < 16 ? cache[dev] : scratch_entry;
        err = prepare_kernel(service_worker_kernel, cached, dev);
        if (err != cudaSuccess) return err;
        while (threads > 128 && service_smem_bytes(sc, threads) > (size_t)cached.max_smem) threads -= 128;
        smem = service_smem_bytes(sc, threads);
        if (smem > (size_t)cached.max_smem) return cudaErrorInvalidValue;
        set_carveout(service_worker_kernel, cached, 1, smem);
    }
    int grid = sm_count - reserved_sms;
    if (grid < 1) grid = 1;
    service_worker_kernel<<<(unsigned)grid, threads, smem, st>>>(sc, q);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_trace(const DevScene &sc, const TraceLaunch &p, int sm_count, cudaStream_t st) {
    if (p.n_photons == 0) return cudaSuccess;
    // the function attributes below are process-wide: launches from the threads of different
    // units (each on its own stream) take turns setting them and launching
    static std::mutex launch_lock;
    std::lock_guard<std::mutex> guard(launch_lock);
    // the largest CTA (up to RL_TRACE_THREADS) whose tables + scratch fit the shared memory of an SM:
    // big CTAs fill the per-CTA task list of the body evaluation best
    static KernelCache cache[16];
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    KernelCache scratch_entry;
    KernelCache &cached = dev >= 0 && dev < 16 ? cache[dev] : scratch_entry;
    err = prepare_kernel(trace_kernel, cached, dev);
    if (err != cudaSuccess) return err;
    const int max_smem = cached.max_smem;
    int threads = RL_TRACE_THREADS;
    while (threads > 128 && trace_kernel_smem_bytes(sc, threads) > (size_t)max_smem) threads -= 128;
    // Small batches (the reference's 524 288 photons are 4.6 per thread of a full grid) spend most
    // of their time in the tail, where a block waits for its last paths.  They are launched as
    // small CTAs, several per SM, so that the blocks of a launch retire one by one and the blocks
    // of the next unit's launch (another stream) move in beside the ones still in their tail.
    // Each such launch also takes only its share of the SM's block slots: with k other units'
    // small launches in flight it asks for ceil(per_sm / (k + 1)) blocks per SM, so k + 1 launches
    // run side by side with three times the paths per thread (fewer table copies and tails per
    // photon) instead of one after the other with every block spending most of its life in its
    // tail.  Measured on the reference's batch, built-in scene, 8 units (tools/strict_rates.py):
    // 768 x 1: 2200 Mrays/s, 256 x 3 full grid: 2310, 256 x 1-of-3: 2600 (large batches: 2750).
    const int small_cta = env_int("RL_TRACE_SMALL_CTA", RL_TRACE_SMALL_CTA);
    const uint64_t small_paths = (uint64_t)env_int("RL_TRACE_SMALL_PATHS", RL_TRACE_SMALL_PATHS);
    const bool small = small_cta >= 128 && small_cta < threads && small_cta % 32 == 0
                       && p.n_photons < small_paths * (uint64_t)sm_count * (uint64_t)threads;
    if (small) threads = small_cta;
    const size_t smem = trace_kernel_smem_bytes(sc, threads);
    if (cached.threads != threads || cached.smem != smem) {
        int occ = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, trace_kernel, threads, smem);
        if (err != cudaSuccess) return err;
        cached.threads = threads; cached.smem = smem; cached.per_sm = occ < 1 ? 1 : occ;
    }
    const int per_sm = cached.per_sm;
    set_carveout(trace_kernel, cached, per_sm, smem);
    uint64_t want = (p.n_photons + threads - 1) / threads;
    uint64_t full = (uint64_t)sm_count * per_sm;
    if (small) {
        // RL_TRACE_BLOCKS_PER_SM > 0 fixes the share (experiments); default: by the concurrency seen
        int share = env_int("RL_TRACE_BLOCKS_PER_SM", 0);
        if (share <= 0) {
            const int others = other_streams_in_flight(dev, st, per_sm - 1 > 1 ? per_sm - 1 : 1);
            share = (per_sm + others) / (others + 1);
        }
        if (share < per_sm) full = (uint64_t)sm_count * (share < 1 ? 1 : share);
    }
    unsigned grid = (unsigned)(want < full ? want : full);
    // the kernel indexes photons of a launch with 32 bits: larger requests take several launches
    const uint64_t chunk = 1ull << 31;
    for (uint64_t done = 0; done < p.n_photons; done += chunk) {
        TraceArgs a;
        a.seed = p.seed;
        a.first_photon = p.first_photon + done;
        a.n_photons = p.n_photons - done < chunk ? p.n_photons - done : chunk;
        a.width = (int)p.width;
        a.height = (int)p.height;
        a.aspect = (float)p.width / (float)p.height;  // trace_unit.rs:73, plot_unit.rs:49
        a.records = p.records ? p.records + done : nullptr;
        a.accum = p.accum;
        a.ray_counter = p.ray_counter;
        trace_kernel<<<grid, threads, smem, st>>>(sc, a);
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
        cudaError_t err = cudaFuncSetAttribute(splat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err == cudaSuccess)
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, splat_kernel, RL_SPLAT_THREADS, smem);
        if (err != cudaSuccess) return err;
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
                           float *exposure_out, uint8_t *rgb, int sm_count, cudaStream_t st) {
    const uint64_t n_pixels = (uint64_t)width * height;
    if (n_pixels == 0) return cudaSuccess;
    cudaError_t err = cudaMemsetAsync(moments, 0, 2 * sizeof(double), st);
    if (err != cudaSuccess) return err;
    uint64_t want = (n_pixels + 255) / 256;
    uint64_t full = (uint64_t)sm_count * 8;
    unsigned grid = (unsigned)(want < full ? want : full);
    tonemap_moments_kernel<<<grid, 256, 0, st>>>(xyz, n_pixels, moments);
    tonemap_exposure_kernel<<<1, 1, 0, st>>>(moments, width, height, exposure_out);
    tonemap_map_kernel<<<grid, 256, 0, st>>>(xyz, n_pixels, exposure_out, rgb);
    g_launches += 3;
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
        const Hit h = intersect_scene(r);
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
            alive = true;
        }
        if (!__syncthreads_or(alive)) break;
        const Ray r = alive ? ray : idle_ray();
        const Hit culled = intersect_scene(r);
        if (!alive) continue;
        const Hit hit = intersect_scene_brute(r);
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
        const V3 dir = material_bounce(m, ray, hit, s, rng, key, probability);
        intensity = intensity * probability;
        ray.direction = dir;
        ray.origin = s.position + dir * 0.00001f;
        continue_chance = continue_chance * 0.96f;
        if (rng.unit(key) * 0.85f > continue_chance * (1.0f - spec_exp(intensity * -20.0f))) continue;
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
