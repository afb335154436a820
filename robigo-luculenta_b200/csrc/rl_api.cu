// rl_api.cu -- the extern "C" boundary (include/rl_b200.h): handles, scene
// flattening, stream ordering between units.  No compute happens on the host;
// without a CUDA device every compute entry point returns RL_ERR_CUDA.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rl_b200.h"
#include "rl_kernels.h"
#include "rl_math.cuh"

using namespace rl;

namespace {

thread_local std::string g_error;

int fail(int code, const std::string &msg) {
    g_error = msg;
    return code;
}

// bytes that crossed the host/device boundary through this ABI (rl_transfer_counters)
std::atomic<uint64_t> g_h2d_bytes{0}, g_d2h_bytes{0};

cudaError_t copy_async(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t stream) {
    if (kind == cudaMemcpyHostToDevice) g_h2d_bytes += bytes;
    else if (kind == cudaMemcpyDeviceToHost) g_d2h_bytes += bytes;
    return cudaMemcpyAsync(dst, src, bytes, kind, stream);
}

#define RL_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(RL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));      \
    } while (0)

struct Device {
    int index = 0;
    int sm_count = 148;
    size_t max_smem = 0;
};

int current_device(Device &d) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(RL_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    }
    RL_CUDA(cudaGetDevice(&d.index));
    RL_CUDA(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, d.index));
    int smem = 0;
    RL_CUDA(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, d.index));
    d.max_smem = (size_t)smem;
    return RL_OK;
}

// Entry points run on the handle's device and leave the calling thread's current device as
// they found it (one process may drive several GPUs, or share the thread with torch).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched && prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

// A stream owned by the handle unless the caller bound its own.
struct StreamSlot {
    cudaStream_t stream = nullptr;
    bool owned = false;
    cudaEvent_t event = nullptr;  // ordering point for other units

    int create() {
        RL_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        owned = true;
        RL_CUDA(cudaEventCreateWithFlags(&event, cudaEventDisableTiming));
        return RL_OK;
    }
    // Work queued on the old stream finishes before anything is queued on the new one (a render
    // still writing the unit's buffers must not race with launches on the next stream).
    int bind(void *external) {
        cudaStream_t next = nullptr;
        if (external) next = (cudaStream_t)external;
        else RL_CUDA(cudaStreamCreateWithFlags(&next, cudaStreamNonBlocking));
        if (stream) {
            cudaError_t e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) {
                if (!external) cudaStreamDestroy(next);
                return fail(RL_ERR_CUDA, std::string("set_stream: ") + cudaGetErrorString(e));
            }
            if (owned) cudaStreamDestroy(stream);
        }
        stream = next;
        owned = external == nullptr;
        return RL_OK;
    }
    void destroy() {
        if (owned && stream) cudaStreamDestroy(stream);
        if (event) cudaEventDestroy(event);
        stream = nullptr; event = nullptr;
    }
};

// Make `later` wait for everything queued so far on `earlier`.
int order_after(StreamSlot &earlier, StreamSlot &later) {
    if (earlier.stream == later.stream) return RL_OK;
    RL_CUDA(cudaEventRecord(earlier.event, earlier.stream));
    RL_CUDA(cudaStreamWaitEvent(later.stream, earlier.event, 0));
    return RL_OK;
}

}  // namespace

// ------------------------------------------------------------------ handles
// The trace dispatcher of a scene.  The reference's host hands the engine 524 288-photon batches
// from C worker threads over 3C trace units (task_scheduler.rs:95-96,127-182; app.rs:104-109):
// 4.6 photons per thread of a full grid, so a launch per batch spends a third of its life in its
// tail (the last paths of every CTA).  TraceUnit::render calls of that size are therefore queued
// here, and a dispatcher thread traces everything that is queued with ONE launch (up to
// RL_MAX_SEGMENTS batches, csrc/rl_kernels.cu K1) whenever one of its two launch slots is free:
// while two launches are in flight the queue grows, so under load a launch carries a dozen or
// more batches and its tail is a few per cent, overlapped by the other slot's launch; a lone
// batch is launched at once.  Everything is ordered with CUDA events -- the launch waits for
// what each unit's stream held when render() was called, each unit's stream waits for the launch
// and then copies the records to the host -- so no stream or thread ever polls.
// Back-pressure comes with it: a unit's next call waits (on a condition variable) until its
// previous batch has been launched, so worker threads wait inside render()/plot(), as they do in
// the reference, instead of in the scheduler's 100 ms sleep task (app.rs:128-130).
struct TraceDispatcher {
    struct Request {
        rl_trace_unit *unit;
        uint64_t first_photon, n_photons;
        rl_mapped_photon *out;               // host buffer for the records, or nullptr
    };
    std::mutex m;
    std::condition_variable cv_work;         // dispatcher: a request is queued / stop
    std::condition_variable cv_done;         // callers: requests have been launched
    std::deque<Request> queue;
    std::thread thread;
    bool started = false, stop = false, failed = false;
    static const int MAX_SLOTS = 8;
    int slots = 2;                           // launches in flight at once
    cudaStream_t streams[MAX_SLOTS] = {};
    cudaEvent_t done[MAX_SLOTS] = {};
    uint64_t launches = 0, batches = 0;      // statistics (rl_scene_dispatch_stats)
};

struct rl_scene {
    Device dev;
    DevScene ds;
    void *d_blob = nullptr;
    void *d_materials = nullptr;
    void *d_keyframes = nullptr;
    size_t smem = 0;
    // TraceUnit::render takes its photon ids from the scene's batch counter: one scene is one
    // App (app.rs:63), two renders in one process do not share ids
    mutable std::atomic<uint64_t> next_batch{0};
    mutable TraceDispatcher dispatcher;
};

struct rl_trace_unit {
    uint64_t id = 0;
    uint32_t width = 0, height = 0;
    uint64_t seed = 0;
    uint64_t batch = RL_BATCH_PHOTONS;
    Device dev;
    StreamSlot ss;
    rl_mapped_photon *d_records = nullptr;
    uint64_t capacity = 0;
    uint64_t n_valid = 0;  // records left on the device by the last render
    unsigned long long *d_rays = nullptr;
    // trace dispatcher: the unit's latest batch sits in a scene's queue until `queued` drops; from
    // then on everything it needs is ordered on the unit's stream
    std::atomic<bool> queued{false};
    TraceDispatcher *dispatcher = nullptr;   // valid while `queued`
    cudaEvent_t ready = nullptr;             // what the unit's stream held when the batch was queued
    cudaEvent_t traced = nullptr;            // the latest records are complete (before their copy to the host)
    int dispatch_rc = RL_OK;                 // outcome of the launch, reported by the unit's next call
    std::string dispatch_error;
};

struct rl_plot_unit {
    uint64_t id = 0;
    uint32_t width = 0, height = 0;
    Device dev;
    StreamSlot ss;
    float4 *d_accum = nullptr;
    float *d_packed = nullptr;
    rl_mapped_photon *d_staging = nullptr;
    uint64_t staging_capacity = 0;
};

// buffer.raw writer of a gather unit (gather_unit.rs:68-78 rewrites the whole file on every
// gather, app.rs:151; at 1024^2 that is 25 MB of file I/O per gather and was the critical path
// of the whole pipeline -- and with the GPU gathering a hundred times a second it would be
// 2.5 GB/s of device-to-host traffic per GPU for checkpoints nobody can tell apart).
// save() snapshots the two device buffers into a device-side copy (25 MB at HBM speed, in the
// unit's stream order) and returns; the unit's writer thread brings the newest snapshot to
// page-locked host memory on its own stream, writes it to `path.tmp` and renames it over
// `path`, so the file on disk is always one complete snapshot.  A save that arrives while an
// older snapshot still waits replaces it (every save rewrites the whole file, only the newest
// state matters), and the writer starts at most one file per `min_interval` (0.1 s unless
// rl_gather_unit_set_save_interval says otherwise; flush / load / destroy do not wait for it):
// the file is never more than that interval plus one write behind the latest save.
// The first save to a path is written synchronously so that an unwritable path fails in the
// caller ("failed to open file", gather_unit.rs:69); a later background failure is returned by
// the next save / flush.
struct SaveWriter {
    std::mutex m;
    std::condition_variable cv;
    std::thread thread;
    bool started = false, stop = false;
    int hurry = 0;                                  // flush() callers waiting: no pacing
    float *host = nullptr;                          // page-locked; the writer's (and the synchronous first save's)
    float *dsnap[2] = {nullptr, nullptr};           // device-side snapshots
    cudaEvent_t taken[2] = {nullptr, nullptr};      // dsnap[i] is complete (recorded on the unit's stream)
    cudaStream_t copy_stream = nullptr;             // the writer's device-to-host copies
    int device = 0;
    size_t floats = 0;
    int latest = -1, reading = -1;                  // newest snapshot not yet taken by the writer / the one it is copying out
    std::string latest_path, error;
    std::vector<std::string> proven;
    double min_interval = 0.1;                      // seconds between the starts of two background writes
    std::chrono::steady_clock::time_point last_start{};

    static bool write_file(const std::string &path, const float *data, size_t floats, std::string &err) {
        const std::string tmp = path + ".tmp";
        FILE *f = fopen(tmp.c_str(), "wb");
        if (!f) { err = "failed to open file " + path; return false; }
        const size_t w = fwrite(data, sizeof(float), floats, f);
        const int c = fclose(f);
        if (w != floats || c != 0 || rename(tmp.c_str(), path.c_str()) != 0) {
            remove(tmp.c_str());
            err = "failed to write raw buffer " + path;
            return false;
        }
        return true;
    }
    cudaError_t allocate(int dev, size_t n_floats) {
        device = dev;
        floats = n_floats;
        cudaError_t e = cudaMallocHost(&host, floats * sizeof(float));
        for (int i = 0; i < 2 && e == cudaSuccess; i++) {
            e = cudaMalloc(&dsnap[i], floats * sizeof(float));
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&taken[i], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) release();
        return e;
    }
    void release() {
        if (host) cudaFreeHost(host);
        host = nullptr;
        for (float *&b : dsnap) { if (b) cudaFree(b); b = nullptr; }
        for (cudaEvent_t &e : taken) { if (e) cudaEventDestroy(e); e = nullptr; }
        if (copy_stream) cudaStreamDestroy(copy_stream);
        copy_stream = nullptr;
    }
    void run() {
        cudaSetDevice(device);
        std::unique_lock<std::mutex> lock(m);
        for (;;) {
            cv.wait(lock, [&] { return latest >= 0 || stop; });
            if (latest < 0) return;
            if (min_interval > 0.0)                 // pacing; a newer save may replace `latest` meanwhile
                cv.wait_until(lock, last_start + std::chrono::duration_cast<std::chrono::steady_clock::duration>(
                                                     std::chrono::duration<double>(min_interval)),
                              [&] { return hurry > 0 || stop; });
            const int idx = latest;
            const std::string path = latest_path;
            latest = -1;
            reading = idx;
            last_start = std::chrono::steady_clock::now();
            lock.unlock();
            std::string err;
            // the snapshot was only queued by save(): wait here, not in the caller, for it
            cudaError_t ce = cudaStreamWaitEvent(copy_stream, taken[idx], 0);
            if (ce == cudaSuccess) {
                g_d2h_bytes += floats * sizeof(float);
                ce = cudaMemcpyAsync(host, dsnap[idx], floats * sizeof(float), cudaMemcpyDeviceToHost, copy_stream);
            }
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(copy_stream);
            bool ok = ce == cudaSuccess;
            if (!ok) err = std::string("buffer.raw snapshot: ") + cudaGetErrorString(ce);
            else ok = write_file(path, host, floats, err);
            lock.lock();
            reading = -1;
            if (!ok && error.empty()) error = err;
            cv.notify_all();
        }
    }
    // waits until nothing is queued or being written; returns the first background failure
    std::string flush() {
        std::unique_lock<std::mutex> lock(m);
        hurry++;
        cv.notify_all();
        cv.wait(lock, [&] { return latest < 0 && reading < 0; });
        hurry--;
        std::string e;
        e.swap(error);
        return e;
    }
    void shutdown() {
        if (started) {
            {
                std::lock_guard<std::mutex> lock(m);
                stop = true;
            }
            cv.notify_all();
            thread.join();
            started = false;
        }
        release();
    }
};

struct rl_gather_unit {
    uint32_t width = 0, height = 0;
    Device dev;
    StreamSlot ss;
    float *d_acc = nullptr;
    float *d_comp = nullptr;
    float *d_staging = nullptr;
    SaveWriter writer;
};

struct rl_tonemap_unit {
    uint32_t width = 0, height = 0;
    Device dev;
    StreamSlot ss;
    float *d_xyz = nullptr;
    uint8_t *d_rgb = nullptr;
    double *d_moments = nullptr;
    float *d_exposure = nullptr;
    float last_exposure = 0.0f;
    bool reference_fold = false;   // find_exposure by the reference's sequential f32 folds
};

// ------------------------------------------------------------ scene flatten
namespace {

struct Flat {
    std::vector<float4> spheres, sphere_k, planes, paraboloids, leaves, compounds;
    double cmax2 = 0.0, leaf_off_max = 0.0;
    bool sphere_leaves = false;
    std::vector<uint32_t> ops, sphere_obj, plane_obj, paraboloid_obj, compound_obj;
};

float4 f4(rl_vec3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// Emits the post-order program of a compound tree; returns false on a
// malformed or unsupported tree.  [lo, hi) = leaf range of the subtree.
bool emit_compound(const rl_scene_desc *d, uint32_t node, Flat &fl, uint32_t first_leaf, int depth,
                   uint32_t &lo, uint32_t &hi) {
    if (node >= d->n_surfaces || depth > 16) return false;
    const rl_surface &s = d->surfaces[node];
    if (s.kind == RL_SURFACE_HALFSPACE) {
        uint32_t rel = (uint32_t)(fl.leaves.size() / 2) - first_leaf;
        if (rel > 254) return false;
        // w of the normal record: max(1, |n|), the scale of the slab test's inflation
        const double nlen = sqrt((double)s.a.x * s.a.x + (double)s.a.y * s.a.y + (double)s.a.z * s.a.z);
        fl.leaves.push_back(f4(s.a, (float)(nlen > 1.0 ? nlen * 1.000001 : 1.0)));
        fl.leaves.push_back(f4(s.b, 0.f));
        const double olen = sqrt((double)s.b.x * s.b.x + (double)s.b.y * s.b.y + (double)s.b.z * s.b.z);
        if (olen > fl.leaf_off_max) fl.leaf_off_max = olen;
        fl.ops.push_back(0u | (rel << 8));
        lo = rel; hi = rel + 1;
        return true;
    }
    if (s.kind == RL_SURFACE_SPHERE) {                      // the other Volume leaf (geometry.rs:263-267)
        uint32_t rel = (uint32_t)(fl.leaves.size() / 2) - first_leaf;
        if (rel > 254) return false;
        fl.leaves.push_back(make_float4(0.f, 0.f, 0.f, 0.f));   // w == 0 marks a sphere leaf; n = 0 keeps it out of the slab test
        fl.leaves.push_back(f4(s.a, s.s));
        fl.sphere_leaves = true;
        fl.ops.push_back(0u | (rel << 8));
        lo = rel; hi = rel + 1;
        return true;
    }
    if (s.kind != RL_SURFACE_COMPOUND) return false;
    uint32_t lo1, hi1, lo2, hi2;
    if (!emit_compound(d, s.child[0], fl, first_leaf, depth + 1, lo1, hi1)) return false;
    if (!emit_compound(d, s.child[1], fl, first_leaf, depth + 1, lo2, hi2)) return false;
    fl.ops.push_back(1u | (lo1 << 8) | (hi1 << 16) | (hi2 << 24));
    lo = lo1; hi = hi2;
    return true;
}

// Bounding sphere of the convex body cut out by half-spaces n_i.(x - p_i) < 0:
// vertices = triple-plane intersections that satisfy every half-space, computed
// in double; centre = their mean, radius = the largest distance, inflated by
// 1 % + 0.05 so that any point the reference's f32 containment tests can accept
// (geometry.rs:124-128 on positions rounded at magnitude ~1e2: slack ~1e-4)
// lies strictly inside.  Returns false for unbounded or degenerate bodies.
bool convex_bound(const std::vector<float4> &leaves, size_t first, size_t n, float4 &out) {
    struct D3 { double x, y, z; };
    auto dot3 = [](D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; };
    auto cross3 = [](D3 a, D3 b) { return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; };
    std::vector<D3> nrm, off;
    const size_t n_all = n;
    for (size_t i = 0; i < n_all; i++) {
        const float4 a = leaves[2 * (first + i)], b = leaves[2 * (first + i) + 1];
        if (a.w == 0.0f) continue;                          // sphere leaves bound the body by themselves (sphere_leaf_bound)
        nrm.push_back(D3{a.x, a.y, a.z});
        off.push_back(D3{b.x, b.y, b.z});
    }
    n = nrm.size();
    // unbounded iff the recession cone {u : n_i.u <= 0} is non-trivial; its extreme rays are
    // cross products of normal pairs (a cone containing a line leaves no vertices at all)
    for (size_t i = 0; i < n; i++)
        for (size_t j = i + 1; j < n; j++) {
            D3 u = cross3(nrm[i], nrm[j]);
            double len = sqrt(dot3(u, u));
            if (len < 1e-9) continue;
            u = D3{u.x / len, u.y / len, u.z / len};
            for (int sign = -1; sign <= 1; sign += 2) {
                bool in_cone = true;
                for (size_t m = 0; m < n && in_cone; m++)
                    if (sign * dot3(nrm[m], u) > 1e-9) in_cone = false;
                if (in_cone) return false;
            }
        }
    std::vector<D3> verts;
    for (size_t i = 0; i < n; i++)
        for (size_t j = i + 1; j < n; j++)
            for (size_t k = j + 1; k < n; k++) {
                const D3 cjk = cross3(nrm[j], nrm[k]);
                const double det = dot3(nrm[i], cjk);
                if (fabs(det) < 1e-9) continue;
                const double di = dot3(nrm[i], off[i]), dj = dot3(nrm[j], off[j]), dk = dot3(nrm[k], off[k]);
                const D3 cki = cross3(nrm[k], nrm[i]), cij = cross3(nrm[i], nrm[j]);
                const D3 x{(di * cjk.x + dj * cki.x + dk * cij.x) / det,
                           (di * cjk.y + dj * cki.y + dk * cij.y) / det,
                           (di * cjk.z + dj * cki.z + dk * cij.z) / det};
                bool inside = true;
                for (size_t m = 0; m < n && inside; m++) {
                    const D3 r{x.x - off[m].x, x.y - off[m].y, x.z - off[m].z};
                    if (dot3(r, nrm[m]) > 1e-6 * (1.0 + sqrt(dot3(x, x)))) inside = false;
                }
                if (inside) verts.push_back(x);
            }
    if (verts.size() < 4) return false;
    D3 c{0, 0, 0};
    for (const D3 &v : verts) { c.x += v.x; c.y += v.y; c.z += v.z; }
    c = D3{c.x / verts.size(), c.y / verts.size(), c.z / verts.size()};
    double r = 0.0;
    for (const D3 &v : verts) {
        const D3 e{v.x - c.x, v.y - c.y, v.z - c.z};
        r = fmax(r, sqrt(dot3(e, e)));
    }
    if (!(r > 0.0) || !std::isfinite(r)) return false;
    r = r * 1.01 + 0.05;
    out = make_float4((float)c.x, (float)c.y, (float)c.z, (float)(r * r));
    return true;
}

// A body with sphere leaves lies inside each of them: the smallest (inflated like the polytope
// bound) replaces `out` when it is tighter or when the half-spaces alone leave the body unbounded.
void sphere_leaf_bound(const std::vector<float4> &leaves, size_t first, size_t n, float4 &out) {
    for (size_t i = 0; i < n; i++) {
        const float4 a = leaves[2 * (first + i)], b = leaves[2 * (first + i) + 1];
        if (a.w != 0.0f) continue;
        const double r = sqrt(fmax(0.0, (double)b.w)) * 1.01 + 0.05;
        if (!std::isfinite(r)) continue;
        if (out.w < 0.0f || r * r < (double)out.w) out = make_float4(b.x, b.y, b.z, (float)(r * r));
    }
}

// Sphere clusters for the two-level pre-test: recursive median split of the
// centres along the widest axis until a group has at most `leaf` members;
// returns the permutation (cluster members contiguous) and, per cluster, its
// range and bounding sphere {centre m, radius R >= max |c_i - m| + r_i}.
struct Cluster { uint32_t first, count; double m[3], R; };

// `full_leaves`: split at multiples of `leaf`, so that every cluster but the last has exactly
// `leaf` members (the three-level scan walks clusters of eight).
void split_spheres(const std::vector<float4> &sph, std::vector<uint32_t> &idx, size_t lo, size_t hi,
                   size_t leaf, std::vector<Cluster> &out, bool full_leaves = false) {
    if (hi - lo <= leaf) {
        Cluster c;
        c.first = (uint32_t)lo; c.count = (uint32_t)(hi - lo);
        double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
        for (size_t k = lo; k < hi; k++) {
            const float4 s = sph[idx[k]];
            const double p[3] = {s.x, s.y, s.z};
            for (int a = 0; a < 3; a++) { mn[a] = fmin(mn[a], p[a]); mx[a] = fmax(mx[a], p[a]); }
        }
        for (int a = 0; a < 3; a++) c.m[a] = 0.5 * (mn[a] + mx[a]);
        c.R = 0.0;
        for (size_t k = lo; k < hi; k++) {
            const float4 s = sph[idx[k]];
            const double dx = s.x - c.m[0], dy = s.y - c.m[1], dz = s.z - c.m[2];
            c.R = fmax(c.R, sqrt(dx * dx + dy * dy + dz * dz) + sqrt((double)s.w));
        }
        c.R = c.R * (1.0 + 1e-6) + 1e-6;
        out.push_back(c);
        return;
    }
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (size_t k = lo; k < hi; k++) {
        const float4 s = sph[idx[k]];
        const double p[3] = {s.x, s.y, s.z};
        for (int a = 0; a < 3; a++) { mn[a] = fmin(mn[a], p[a]); mx[a] = fmax(mx[a], p[a]); }
    }
    int axis = 0;
    for (int a = 1; a < 3; a++) if (mx[a] - mn[a] > mx[axis] - mn[axis]) axis = a;
    size_t mid = lo + (hi - lo) / 2;
    if (full_leaves) {
        mid = lo + ((hi - lo) / 2 + leaf - 1) / leaf * leaf;
        if (mid >= hi) mid = hi - leaf;
    }
    auto key = [&](uint32_t i) { const float4 s = sph[i]; return axis == 0 ? s.x : (axis == 1 ? s.y : s.z); };
    std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi,
                     [&](uint32_t a, uint32_t b) { return key(a) < key(b) || (key(a) == key(b) && a < b); });
    split_spheres(sph, idx, lo, mid, leaf, out, full_leaves);
    split_spheres(sph, idx, mid, hi, leaf, out, full_leaves);
}

int max_stack(const std::vector<uint32_t> &ops, size_t first, size_t n) {
    int sp = 0, mx = 0;
    for (size_t i = 0; i < n; i++) {
        if ((ops[first + i] & 3u) == 0u) sp++; else sp--;
        if (sp > mx) mx = sp;
    }
    return mx;
}

// ---- trace dispatcher plumbing ----------------------------------------------------------------
int env_int(const char *name, int fallback) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : fallback;
}

// With RL_TRACE_GROUPS=1, batches below this size are queued to the scene's dispatcher.  The same
// bound as the small-launch rule of launch_trace: fewer than 64 photons per thread of a full
// grid -- larger requests fill the GPU on their own.  Off by default: measured on the scheduler
// replay (profiles/r2_dispatch_sweep*.txt), one launch per batch with the launches sharing the
// SMs' block slots (launch_trace) keeps the pipeline finer-grained -- a batch's records start
// their way to the host the moment ITS launch ends, not when the whole group's does -- and
// reaches 98.7 % of the one-launch rate in the steady state, against 75-90 % with groups.
bool use_dispatcher(const rl_scene *scene, uint64_t n_photons) {
    if (n_photons == 0 || n_photons >= (1ull << RL_SEGMENT_INDEX_BITS)) return false;
    if (scene->dispatcher.failed || !env_int("RL_TRACE_GROUPS", 0)) return false;
    return n_photons < 64ull * (uint64_t)scene->dev.sm_count * 768ull;
}

void dispatcher_loop(const rl_scene *scene) {
    TraceDispatcher &d = scene->dispatcher;
    cudaSetDevice(scene->dev.index);
    const size_t group_max = (size_t)std::min(RL_MAX_SEGMENTS, std::max(1, env_int("RL_TRACE_GROUP_MAX", RL_MAX_SEGMENTS)));
    int slot = 0;
    std::vector<TraceDispatcher::Request> group;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(d.m);
            d.cv_work.wait(lk, [&] { return d.stop || !d.queue.empty(); });
            if (d.queue.empty()) return;                          // stop, and nothing left to launch
        }
        // at most two launches in flight: wait for the one that last used this slot; what is
        // queued meanwhile joins this launch
        cudaEventSynchronize(d.done[slot]);
        group.clear();
        {
            std::lock_guard<std::mutex> lk(d.m);
            const uint32_t w = d.queue.front().unit->width, h = d.queue.front().unit->height;
            for (auto it = d.queue.begin(); it != d.queue.end() && group.size() < group_max;) {
                if (it->unit->width == w && it->unit->height == h) { group.push_back(*it); it = d.queue.erase(it); }
                else ++it;
            }
        }
        cudaStream_t st = d.streams[slot];
        TraceLaunch p;
        memset(&p, 0, sizeof(p));
        p.width = group[0].unit->width; p.height = group[0].unit->height;
        p.accum = nullptr;
        p.n_segments = (uint32_t)group.size();
        cudaError_t e = cudaSuccess;
        for (size_t k = 0; k < group.size() && e == cudaSuccess; k++) {
            rl_trace_unit *u = group[k].unit;
            p.seg[k].seed = u->seed;
            p.seg[k].first_photon = group[k].first_photon;
            p.seg[k].n_photons = group[k].n_photons;
            p.seg[k].records = u->d_records;
            p.seg[k].ray_counter = u->d_rays;
            e = cudaStreamWaitEvent(st, u->ready, 0);
        }
        if (e == cudaSuccess) e = launch_trace(scene->ds, p, scene->dev.sm_count, st);
        if (e == cudaSuccess) e = cudaEventRecord(d.done[slot], st);
        for (size_t k = 0; k < group.size() && e == cudaSuccess; k++) {
            rl_trace_unit *u = group[k].unit;
            e = cudaStreamWaitEvent(u->ss.stream, d.done[slot], 0);
            if (e == cudaSuccess) e = cudaEventRecord(u->traced, u->ss.stream);
            if (e == cudaSuccess && group[k].out)
                e = copy_async(group[k].out, u->d_records, group[k].n_photons * sizeof(rl_mapped_photon),
                               cudaMemcpyDeviceToHost, u->ss.stream);
        }
        if (e != cudaSuccess) cudaGetLastError();
        {
            std::lock_guard<std::mutex> lk(d.m);
            d.launches++; d.batches += group.size();
            for (auto &r : group) {
                if (e != cudaSuccess) {
                    r.unit->dispatch_rc = RL_ERR_CUDA;
                    r.unit->dispatch_error = std::string("trace dispatcher: ") + cudaGetErrorString(e);
                }
                r.unit->queued.store(false, std::memory_order_release);
            }
        }
        d.cv_done.notify_all();
        slot = (slot + 1) % d.slots;
    }
}

int dispatcher_start(const rl_scene *scene) {
    TraceDispatcher &d = scene->dispatcher;
    std::lock_guard<std::mutex> lk(d.m);
    if (d.started) return RL_OK;
    cudaError_t e = cudaSuccess;
    d.slots = std::min((int)TraceDispatcher::MAX_SLOTS, std::max(1, env_int("RL_TRACE_SLOTS", 2)));
    for (int k = 0; k < d.slots && e == cudaSuccess; k++) {
        e = cudaStreamCreateWithFlags(&d.streams[k], cudaStreamNonBlocking);
        // blocking sync: the dispatcher thread sleeps while it waits for a slot
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.done[k], cudaEventDisableTiming | cudaEventBlockingSync);
    }
    if (e != cudaSuccess) {
        d.failed = true;
        return fail(RL_ERR_CUDA, std::string("trace dispatcher: ") + cudaGetErrorString(e));
    }
    try {
        d.thread = std::thread(dispatcher_loop, scene);
    } catch (...) {
        d.failed = true;                                  // no thread: batches are launched one by one
        return fail(RL_ERR_NOMEM, "trace dispatcher: cannot create a thread");
    }
    d.started = true;
    return RL_OK;
}

void dispatcher_stop(rl_scene *scene) {
    TraceDispatcher &d = scene->dispatcher;
    {
        std::lock_guard<std::mutex> lk(d.m);
        d.stop = true;
    }
    d.cv_work.notify_all();
    if (d.thread.joinable()) d.thread.join();             // launches what is still queued first
    for (int k = 0; k < TraceDispatcher::MAX_SLOTS; k++) {
        if (d.streams[k]) { cudaStreamSynchronize(d.streams[k]); cudaStreamDestroy(d.streams[k]); }
        if (d.done[k]) cudaEventDestroy(d.done[k]);
    }
}

template <class T>
uint32_t append(std::vector<float4> &blob, const std::vector<T> &v) {
    uint32_t off = (uint32_t)blob.size();
    size_t bytes = v.size() * sizeof(T);
    size_t vec4 = (bytes + 15) / 16;
    blob.resize(blob.size() + vec4, make_float4(0.f, 0.f, 0.f, 0.f));
    if (bytes) memcpy(blob.data() + off, v.data(), bytes);
    return off;
}

}  // namespace

extern "C" {

int rl_abi_version(void) { return RL_ABI_VERSION; }
const char *rl_last_error(void) { return g_error.c_str(); }

int rl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

uint64_t rl_kernel_launch_count(void) { return kernel_launches(); }
void rl_kernel_launch_count_reset(void) { kernel_launches_reset(); }

// -------------------------------------------------------------------- scene
int rl_scene_create(const rl_scene_desc *desc, rl_scene **out) {
    if (!desc || !out) return fail(RL_ERR_INVALID, "rl_scene_create: null argument");
    if ((!desc->surfaces && desc->n_surfaces) || (!desc->objects && desc->n_objects))
        return fail(RL_ERR_INVALID, "rl_scene_create: null table");
    if (desc->camera.kind != RL_CAMERA_STATIC && desc->camera.kind != RL_CAMERA_ORBIT
        && desc->camera.kind != RL_CAMERA_KEYFRAMES)
        return fail(RL_ERR_INVALID, "rl_scene_create: unknown camera kind");
    if (desc->camera.kind == RL_CAMERA_KEYFRAMES && (!desc->camera.keyframes || desc->camera.n_keyframes == 0
                                                      || desc->camera.n_keyframes > (1u << 20)))
        return fail(RL_ERR_INVALID, "rl_scene_create: a keyframe camera needs 1 .. 2^20 keyframes");

    Flat fl;
    std::vector<float4> materials;
    for (uint32_t i = 0; i < desc->n_objects; i++) {
        const rl_object &o = desc->objects[i];
        if (o.surface >= desc->n_surfaces) return fail(RL_ERR_INVALID, "object surface index out of range");
        if (o.material.kind < RL_MATERIAL_BLACKBODY || o.material.kind > RL_MATERIAL_SOAP_BUBBLE)
            return fail(RL_ERR_INVALID, "unknown material kind");
        materials.push_back(make_float4(as_float(o.material.kind), o.material.p0, o.material.p1,
                                        o.material.p2));
        const rl_surface &s = desc->surfaces[o.surface];
        switch (s.kind) {
        case RL_SURFACE_SPHERE: {
            fl.spheres.push_back(f4(s.a, s.s));
            const double c2 = (double)s.a.x * s.a.x + (double)s.a.y * s.a.y + (double)s.a.z * s.a.z;
            fl.sphere_k.push_back(f4(s.a, (float)(c2 - (double)s.s)));
            if (c2 + (double)s.s > fl.cmax2) fl.cmax2 = c2 + (double)s.s;
            fl.sphere_obj.push_back(i);
            break;
        }
        case RL_SURFACE_PLANE:
        case RL_SURFACE_HALFSPACE:
        case RL_SURFACE_CIRCLE:
            fl.planes.push_back(f4(s.a, as_float(s.kind)));
            fl.planes.push_back(f4(s.b, s.s));
            fl.plane_obj.push_back(i);
            break;
        case RL_SURFACE_PARABOLOID:
            fl.paraboloids.push_back(f4(s.a, 0.f));
            fl.paraboloids.push_back(f4(s.b, 0.f));
            fl.paraboloids.push_back(f4(s.c, 0.f));
            fl.paraboloid_obj.push_back(i);
            break;
        case RL_SURFACE_COMPOUND: {
            uint32_t first_leaf = (uint32_t)(fl.leaves.size() / 2);
            uint32_t first_op = (uint32_t)fl.ops.size();
            uint32_t lo, hi;
            if (!emit_compound(desc, o.surface, fl, first_leaf, 0, lo, hi))
                return fail(RL_ERR_UNSUPPORTED,
                            "compound surfaces must be trees of half-spaces and spheres (<= 255 leaves)");
            uint32_t n_ops = (uint32_t)fl.ops.size() - first_op;
            if (max_stack(fl.ops, first_op, n_ops) > RL_MAX_COMPOUND_STACK)
                return fail(RL_ERR_UNSUPPORTED, "compound tree too deep");
            fl.compounds.push_back(make_float4(as_float(first_leaf), as_float(hi - lo),
                                               as_float(first_op), as_float(n_ops)));
            float4 bound = make_float4(0.f, 0.f, 0.f, -1.0f);          // r^2 < 0: unbounded, never culled
            convex_bound(fl.leaves, first_leaf, hi - lo, bound);
            sphere_leaf_bound(fl.leaves, first_leaf, hi - lo, bound);
            fl.compounds.push_back(bound);
            fl.compound_obj.push_back(i);
            break;
        }
        default:
            return fail(RL_ERR_INVALID, "unknown surface kind");
        }
    }

    rl_scene *sc = new (std::nothrow) rl_scene();
    if (!sc) return fail(RL_ERR_NOMEM, "out of memory");
    int rc = current_device(sc->dev);
    if (rc != RL_OK) { delete sc; return rc; }

    std::vector<float4> blob;
    DevScene &ds = sc->ds;
    memset(&ds, 0, sizeof(ds));
    // ---- spheres: cluster, reorder so that cluster members are contiguous, emit the pre-test records
    if (fl.spheres.size() > 65535) {
        delete sc;
        return fail(RL_ERR_UNSUPPORTED, "more than 65535 spheres");
    }
    std::vector<float4> clusters;
    std::vector<uint32_t> cluster_range;
    std::vector<float4> supers;
    double cluster_rmax = 0.0, super_rmax = 0.0;
    bool deep = false;
    {
        const size_t n = fl.spheres.size();
        std::vector<uint32_t> idx(n);
        for (size_t k = 0; k < n; k++) idx[k] = (uint32_t)k;
        std::vector<Cluster> cl;
        // clusters of about eight members, more for larger scenes (balances the uniform scan of
        // the cluster bounds against the member tests; measured flat between 5 and 16 on the
        // built-in scene, profiles/r2_cluster_leaf_sweep.txt)
        size_t leaf = (size_t)(0.45 * sqrt((double)n) + 0.5);
        leaf = leaf < 8 ? 8 : (leaf > 32 ? 32 : leaf);
        // a thousand spheres or more: three levels (groups of eight clusters of eight spheres), so
        // that neither the uniform scan of the top level nor the member tests grow with sqrt(n)
        deep = env_int("RL_DEEP_CLUSTERS", n >= 1024 ? 1 : 0) != 0;
        if (deep) leaf = 8;
        if (env_int("RL_CLUSTER_LEAF", 0) > 0) leaf = (size_t)env_int("RL_CLUSTER_LEAF", 0);   // experiments
        if (leaf < (n + 999) / 1000) leaf = (n + 999) / 1000;   // pair records index clusters with 11 bits
        if (n) split_spheres(fl.spheres, idx, 0, n, leaf, cl, deep);
        std::vector<float4> spheres(n), sphere_k(n);
        std::vector<uint32_t> sphere_obj(n);
        for (size_t k = 0; k < n; k++) {
            spheres[k] = fl.spheres[idx[k]];
            sphere_k[k] = fl.sphere_k[idx[k]];
            sphere_obj[k] = fl.sphere_obj[idx[k]];
        }
        fl.spheres.swap(spheres); fl.sphere_k.swap(sphere_k); fl.sphere_obj.swap(sphere_obj);
        for (const Cluster &c : cl) {
            const double m2 = c.m[0] * c.m[0] + c.m[1] * c.m[1] + c.m[2] * c.m[2];
            const float4 rec = make_float4((float)c.m[0], (float)c.m[1], (float)c.m[2], 0.f);
            // the record's centre is the f32-rounded one: bound the radius from it
            double R = 0.0;
            for (uint32_t k = c.first; k < c.first + c.count; k++) {
                const float4 sp = fl.spheres[k];
                const double dx = sp.x - rec.x, dy = sp.y - rec.y, dz = sp.z - rec.z;
                R = fmax(R, sqrt(dx * dx + dy * dy + dz * dz) + sqrt((double)sp.w));
            }
            R = R * (1.0 + 1e-6) + 1e-6;
            const double mr2 = (double)rec.x * rec.x + (double)rec.y * rec.y + (double)rec.z * rec.z;
            (void)m2;
            clusters.push_back(make_float4(rec.x, rec.y, rec.z, (float)(mr2 - R * R)));
            cluster_range.push_back(c.first | (c.count << 16));
            if (R > cluster_rmax) cluster_rmax = R;
            if (mr2 + R * R > fl.cmax2) fl.cmax2 = mr2 + R * R;
        }
        // the cluster scan takes eight records per step: pad with records no ray selects
        while (clusters.size() % 8 != 0) {
            clusters.push_back(make_float4(0.f, 0.f, 0.f, 1.0e30f));
            cluster_range.push_back(0u);
        }
        // the level above: clusters come out of the median split in tree order, so eight
        // consecutive ones are neighbours; their bound is taken over the member spheres themselves
        deep = deep && cl.size() > 8;
        for (size_t g = 0; deep && g < cl.size(); g += 8) {
            const size_t g_end = g + 8 < cl.size() ? g + 8 : cl.size();
            const uint32_t first = cl[g].first, last = cl[g_end - 1].first + cl[g_end - 1].count;
            double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
            for (uint32_t k = first; k < last; k++) {
                const float4 sp = fl.spheres[k];
                const double p[3] = {sp.x, sp.y, sp.z};
                for (int a = 0; a < 3; a++) { mn[a] = fmin(mn[a], p[a]); mx[a] = fmax(mx[a], p[a]); }
            }
            const float4 rec = make_float4((float)(0.5 * (mn[0] + mx[0])), (float)(0.5 * (mn[1] + mx[1])),
                                           (float)(0.5 * (mn[2] + mx[2])), 0.f);
            double R = 0.0;
            for (uint32_t k = first; k < last; k++) {
                const float4 sp = fl.spheres[k];
                const double dx = sp.x - rec.x, dy = sp.y - rec.y, dz = sp.z - rec.z;
                R = fmax(R, sqrt(dx * dx + dy * dy + dz * dz) + sqrt((double)sp.w));
            }
            R = R * (1.0 + 1e-6) + 1e-6;
            const double mr2 = (double)rec.x * rec.x + (double)rec.y * rec.y + (double)rec.z * rec.z;
            supers.push_back(make_float4(rec.x, rec.y, rec.z, (float)(mr2 - R * R)));
            if (R > super_rmax) super_rmax = R;
            if (mr2 + R * R > fl.cmax2) fl.cmax2 = mr2 + R * R;
        }
        while (supers.size() % 8 != 0) supers.push_back(make_float4(0.f, 0.f, 0.f, 1.0e30f));
    }
    ds.n_spheres = (uint32_t)fl.spheres.size();
    ds.off_planes = append(blob, fl.planes);            ds.n_planes = (uint32_t)fl.planes.size() / 2;
    ds.off_paraboloids = append(blob, fl.paraboloids);  ds.n_paraboloids = (uint32_t)fl.paraboloids.size() / 3;
    if (fl.leaves.size() / 2 > 65535u) {
        delete sc;
        return fail(RL_ERR_UNSUPPORTED, "more than 65535 half-spaces in compound surfaces");
    }
    ds.off_leaves = append(blob, fl.leaves);            ds.n_leaves = (uint32_t)fl.leaves.size() / 2;
    // (lane, cluster) and (lane, body) pair records index their table with RL_PAIR_INDEX_BITS bits
    if (fl.compounds.size() / 2 > RL_PAIR_INDEX_MAX || clusters.size() > RL_PAIR_INDEX_MAX) {
        delete sc;
        return fail(RL_ERR_UNSUPPORTED, "more than 2047 compound surfaces (or sphere clusters) in one scene");
    }
    ds.off_compounds = append(blob, fl.compounds);      ds.n_compounds = (uint32_t)fl.compounds.size() / 2;
    ds.off_ops = append(blob, fl.ops);                  ds.n_ops = (uint32_t)fl.ops.size();
    {
        // the compounds' bounding spheres once more, in the form the sphere pre-test scans:
        // {c, |c|^2 - r^2}, centre as stored (f32), padded to eight records; unbounded bodies are
        // flagged per group of 64 and always evaluated
        std::vector<float4> body_bounds;
        std::vector<uint64_t> body_always((fl.compounds.size() / 2 + 63) / 64 + 1, 0ull);
        double body_rmax = 0.0;
        for (size_t k = 0; k < fl.compounds.size() / 2; k++) {
            const float4 b = fl.compounds[2 * k + 1];
            if (b.w < 0.0f) {
                body_always[k / 64] |= 1ull << (k % 64);
                body_bounds.push_back(make_float4(0.f, 0.f, 0.f, 1.0e30f));
                continue;
            }
            const double c2 = (double)b.x * b.x + (double)b.y * b.y + (double)b.z * b.z;
            const double r2 = (double)b.w * (1.0 + 1e-6);          // the f32 record of |c|^2 - r^2 rounds: keep it outside
            body_bounds.push_back(make_float4(b.x, b.y, b.z, (float)(c2 - r2)));
            if (c2 + r2 > fl.cmax2) fl.cmax2 = c2 + r2;
            if (sqrt(r2) > body_rmax) body_rmax = sqrt(r2);
        }
        while (body_bounds.size() % 8 != 0 || body_bounds.empty()) body_bounds.push_back(make_float4(0.f, 0.f, 0.f, 1.0e30f));
        ds.off_body_bounds = append(blob, body_bounds);
        ds.off_body_always = append(blob, body_always);
        ds.body_rmax = (float)(body_rmax * 1.0001);
    }
    // the pre-test records of the spheres go to shared memory with the rest -- unless there are so
    // many of them (> 6144: 96 KB) that the CTAs would shrink or the scene would not fit at all:
    // then they stay in global memory (read through L1 by the cooperative member test) and the
    // scene is bounded by the 65 535 spheres of the 16-bit candidate queues only
    const bool sphere_k_global = fl.sphere_k.size() > 6144 && !env_int("RL_SPHERE_K_SHARED", 0);
    if (!sphere_k_global) ds.off_sphere_k = append(blob, fl.sphere_k);
    ds.off_clusters = append(blob, clusters);           ds.n_clusters = (uint32_t)clusters.size();
    ds.off_cluster_range = append(blob, cluster_range);
    ds.off_supers = append(blob, supers);               ds.n_supers = (uint32_t)supers.size();
    ds.super_rmax = (float)(super_rmax * 1.0001);
    ds.sphere_cmax2 = (float)(fl.cmax2 * 1.0001);
    ds.cluster_rmax = (float)(cluster_rmax * 1.0001);
    ds.leaf_off_max = (float)(fl.leaf_off_max * 1.0001);
    ds.sphere_leaves = fl.sphere_leaves ? 1u : 0u;
    ds.off_plane_obj = append(blob, fl.plane_obj);
    ds.off_paraboloid_obj = append(blob, fl.paraboloid_obj);
    ds.off_compound_obj = append(blob, fl.compound_obj);
    if (blob.empty()) blob.push_back(make_float4(0.f, 0.f, 0.f, 0.f));
    ds.smem_vec4 = (uint32_t)blob.size();               // everything above goes to shared memory
    ds.sphere_k_global = sphere_k_global ? 1u : 0u;
    if (sphere_k_global) ds.off_sphere_k = append(blob, fl.sphere_k);
    ds.off_spheres = append(blob, fl.spheres);          // exact records: global memory only
    ds.off_sphere_obj = append(blob, fl.sphere_obj);
    if (blob.empty()) blob.push_back(make_float4(0.f, 0.f, 0.f, 0.f));
    ds.blob_vec4 = (uint32_t)blob.size();
    ds.n_objects = desc->n_objects;
    sc->smem = trace_smem_bytes(ds, 128);   // the smallest CTA the kernels are launched with
    if (sc->smem > sc->dev.max_smem) {
        delete sc;
        return fail(RL_ERR_UNSUPPORTED, "scene primitive tables exceed shared memory per CTA");
    }

    const rl_camera_model &cm = desc->camera;
    DevCamera &c = ds.camera;
    c.kind = cm.kind;
    c.px = cm.fixed.position.x; c.py = cm.fixed.position.y; c.pz = cm.fixed.position.z;
    c.field_of_view = cm.fixed.field_of_view;
    {
        // camera.rs:60, in the specified arithmetic the kernels use (1 / tan, tan = sin / cos)
        float fs, fc;
        spec_sincos(c.field_of_view * 0.5f, fs, fc);
        c.screen_distance = 1.0f / (fs / fc);
    }
    c.focal_distance = cm.fixed.focal_distance;
    c.depth_of_field = cm.fixed.depth_of_field;
    c.chromatic_abberation = cm.fixed.chromatic_abberation;
    c.qx = cm.fixed.orientation.x; c.qy = cm.fixed.orientation.y;
    c.qz = cm.fixed.orientation.z; c.qw = cm.fixed.orientation.w;
    c.phi_base = cm.phi_base; c.phi_rate = cm.phi_rate;
    c.alpha_base = cm.alpha_base; c.alpha_rate = cm.alpha_rate;
    c.distance_base = cm.distance_base; c.distance_rate = cm.distance_rate;
    c.focal_factor = cm.focal_factor;

    // the tabulated camera function: three float4 per frame, 1 / tan(fov / 2) evaluated per frame
    std::vector<float4> keyframes;
    if (cm.kind == RL_CAMERA_KEYFRAMES) {
        for (uint32_t k = 0; k < cm.n_keyframes; k++) {
            const rl_camera &f = cm.keyframes[k];
            float fs, fc;
            spec_sincos(f.field_of_view * 0.5f, fs, fc);
            keyframes.push_back(make_float4(f.position.x, f.position.y, f.position.z, f.focal_distance));
            keyframes.push_back(make_float4(f.orientation.x, f.orientation.y, f.orientation.z, f.orientation.w));
            keyframes.push_back(make_float4(f.depth_of_field, f.chromatic_abberation, 1.0f / (fs / fc), 0.f));
        }
        c.n_keyframes = cm.n_keyframes;
    }

    if (materials.empty()) materials.push_back(make_float4(0.f, 0.f, 0.f, 0.f));
    cudaError_t e = cudaMalloc(&sc->d_blob, blob.size() * sizeof(float4));
    if (e == cudaSuccess && !keyframes.empty()) {
        e = cudaMalloc(&sc->d_keyframes, keyframes.size() * sizeof(float4));
        if (e == cudaSuccess)
            e = cudaMemcpy(sc->d_keyframes, keyframes.data(), keyframes.size() * sizeof(float4), cudaMemcpyHostToDevice);
        g_h2d_bytes += keyframes.size() * sizeof(float4);
    }
    if (e == cudaSuccess) e = cudaMalloc(&sc->d_materials, materials.size() * sizeof(float4));
    if (e == cudaSuccess)
        e = cudaMemcpy(sc->d_blob, blob.data(), blob.size() * sizeof(float4), cudaMemcpyHostToDevice);
    g_h2d_bytes += blob.size() * sizeof(float4) + materials.size() * sizeof(float4);
    if (e == cudaSuccess)
        e = cudaMemcpy(sc->d_materials, materials.data(), materials.size() * sizeof(float4),
                       cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(sc->d_blob); cudaFree(sc->d_materials); cudaFree(sc->d_keyframes);
        delete sc;
        return fail(RL_ERR_CUDA, std::string("rl_scene_create: ") + cudaGetErrorString(e));
    }
    ds.camera.keyframes = (const float4 *)sc->d_keyframes;
    ds.blob = (const float4 *)sc->d_blob;
    ds.materials = (const float4 *)sc->d_materials;
    *out = sc;
    return RL_OK;
}

int rl_scene_destroy(rl_scene *scene) {
    if (!scene) return RL_OK;
    DeviceGuard guard(scene->dev.index);
    dispatcher_stop(scene);
    cudaFree(scene->d_blob);
    cudaFree(scene->d_materials);
    cudaFree(scene->d_keyframes);
    delete scene;
    return RL_OK;
}

// --------------------------------------------------------------- TraceUnit
int rl_trace_unit_create(uint64_t id, uint32_t width, uint32_t height, uint64_t seed,
                         rl_trace_unit **out) {
    if (!out || width == 0 || height == 0) return fail(RL_ERR_INVALID, "rl_trace_unit_create: bad argument");
    rl_trace_unit *u = new (std::nothrow) rl_trace_unit();
    if (!u) return fail(RL_ERR_NOMEM, "out of memory");
    u->id = id; u->width = width; u->height = height; u->seed = seed;
    int rc = current_device(u->dev);
    if (rc == RL_OK) rc = u->ss.create();
    if (rc != RL_OK) { delete u; return rc; }
    cudaError_t e = cudaMalloc(&u->d_rays, sizeof(unsigned long long));
    // cleared on the unit's own stream: everything the unit does later is ordered behind it
    if (e == cudaSuccess) e = cudaMemsetAsync(u->d_rays, 0, sizeof(unsigned long long), u->ss.stream);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&u->ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&u->traced, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        cudaFree(u->d_rays);
        if (u->ready) cudaEventDestroy(u->ready);
        if (u->traced) cudaEventDestroy(u->traced);
        u->ss.destroy(); delete u;
        return fail(RL_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = u;
    return RL_OK;
}

// The unit's latest batch has been launched: everything it needs (the launch, the copy of its
// records to the host) is ordered on the unit's stream from here on.  A short host wait on the
// dispatcher's condition variable when the batch is still queued.
static int trace_settle(rl_trace_unit *u) {
    if (u->queued.load(std::memory_order_acquire)) {
        TraceDispatcher *d = u->dispatcher;
        std::unique_lock<std::mutex> lk(d->m);
        d->cv_done.wait(lk, [&] { return !u->queued.load(std::memory_order_acquire); });
    }
    if (u->dispatch_rc != RL_OK) {
        const int rc = u->dispatch_rc;
        u->dispatch_rc = RL_OK;
        return fail(rc, u->dispatch_error);
    }
    return RL_OK;
}

int rl_trace_unit_destroy(rl_trace_unit *u) {
    if (!u) return RL_OK;
    DeviceGuard guard(u->dev.index);
    trace_settle(u);
    if (u->ss.stream) cudaStreamSynchronize(u->ss.stream);
    cudaFree(u->d_records);
    cudaFree(u->d_rays);
    if (u->ready) cudaEventDestroy(u->ready);
    if (u->traced) cudaEventDestroy(u->traced);
    u->ss.destroy();
    delete u;
    return RL_OK;
}

int rl_trace_unit_set_batch_size(rl_trace_unit *u, uint64_t n) {
    if (!u || n == 0) return fail(RL_ERR_INVALID, "rl_trace_unit_set_batch_size: bad argument");
    u->batch = n;
    return RL_OK;
}

int rl_trace_unit_set_stream(rl_trace_unit *u, void *s) {
    if (!u) return fail(RL_ERR_INVALID, "null unit");
    DeviceGuard guard(u->dev.index);
    int rc = trace_settle(u);
    if (rc != RL_OK) return rc;
    return u->ss.bind(s);
}

static int ensure_records(rl_trace_unit *u, uint64_t n) {
    if (u->capacity >= n) return RL_OK;
    RL_CUDA(cudaStreamSynchronize(u->ss.stream));
    cudaFree(u->d_records);
    u->d_records = nullptr; u->capacity = 0;
    RL_CUDA(cudaMalloc(&u->d_records, n * sizeof(rl_mapped_photon)));
    u->capacity = n;
    return RL_OK;
}

// Photons [first, first + n) of `scene` into the unit's record buffer (and `out`, a host buffer,
// when given): through the scene's dispatcher when the batch is small (the reference's 524 288
// photons), as a launch on the unit's stream otherwise.  Returns once the work is queued.
static int queue_records(rl_trace_unit *u, const rl_scene *scene, uint64_t first_photon, uint64_t n_photons,
                         rl_mapped_photon *out) {
    if (n_photons == 0) return RL_OK;
    if (use_dispatcher(scene, n_photons) && dispatcher_start(scene) == RL_OK) {
        TraceDispatcher &d = scene->dispatcher;
        // the launch waits for what the unit's stream holds now (the copy and the splat of the
        // unit's previous records)
        RL_CUDA(cudaEventRecord(u->ready, u->ss.stream));
        {
            std::lock_guard<std::mutex> lk(d.m);
            u->dispatcher = &d;
            u->queued.store(true, std::memory_order_release);
            d.queue.push_back(TraceDispatcher::Request{u, first_photon, n_photons, out});
        }
        d.cv_work.notify_one();
        return RL_OK;
    }
    TraceLaunch p;
    memset(&p, 0, sizeof(p));
    p.width = u->width; p.height = u->height;
    p.accum = nullptr;
    p.n_segments = 1;
    p.seg[0].seed = u->seed; p.seg[0].first_photon = first_photon; p.seg[0].n_photons = n_photons;
    p.seg[0].records = u->d_records; p.seg[0].ray_counter = u->d_rays;
    RL_CUDA(launch_trace(scene->ds, p, u->dev.sm_count, u->ss.stream));
    RL_CUDA(cudaEventRecord(u->traced, u->ss.stream));
    if (out)
        RL_CUDA(copy_async(out, u->d_records, n_photons * sizeof(rl_mapped_photon), cudaMemcpyDeviceToHost,
                           u->ss.stream));
    return RL_OK;
}

static int render_range(rl_trace_unit *u, const rl_scene *scene, uint64_t first_photon,
                        uint64_t n_photons, rl_mapped_photon *out, bool wait) {
    if (!u || !scene) return fail(RL_ERR_INVALID, "rl_trace_unit_render: null argument");
    if (scene->dev.index != u->dev.index) return fail(RL_ERR_INVALID, "scene and unit on different devices");
    DeviceGuard guard(u->dev.index);
    int rc = trace_settle(u);                    // the previous batch has been launched
    if (rc == RL_OK) rc = ensure_records(u, n_photons);
    if (rc == RL_OK) rc = queue_records(u, scene, first_photon, n_photons, out);
    if (rc != RL_OK) return rc;
    u->n_valid = n_photons;
    if (wait && out && n_photons) {
        rc = trace_settle(u);
        if (rc != RL_OK) return rc;
        RL_CUDA(cudaStreamSynchronize(u->ss.stream));
    }
    return RL_OK;
}

int rl_trace_unit_render_range(rl_trace_unit *u, const rl_scene *scene, uint64_t first_photon,
                               uint64_t n_photons, rl_mapped_photon *out) {
    return render_range(u, scene, first_photon, n_photons, out, true);
}

int rl_trace_unit_render(rl_trace_unit *u, const rl_scene *scene, rl_mapped_photon *out) {
    if (!u || !scene) return fail(RL_ERR_INVALID, "rl_trace_unit_render: null argument");
    const uint64_t b = scene->next_batch.fetch_add(1);
    return render_range(u, scene, b * u->batch, u->batch, out, true);
}

int rl_trace_unit_render_async(rl_trace_unit *u, const rl_scene *scene, rl_mapped_photon *out) {
    if (!u || !scene) return fail(RL_ERR_INVALID, "rl_trace_unit_render_async: null argument");
    const uint64_t b = scene->next_batch.fetch_add(1);
    return render_range(u, scene, b * u->batch, u->batch, out, false);
}

int rl_trace_unit_download(rl_trace_unit *u, rl_mapped_photon *out, uint64_t capacity, uint64_t *out_count) {
    if (!u || (!out && capacity)) return fail(RL_ERR_INVALID, "rl_trace_unit_download: null argument");
    if (u->n_valid > capacity)
        return fail(RL_ERR_INVALID, "rl_trace_unit_download: the last render left more records than `capacity`");
    DeviceGuard guard(u->dev.index);
    int rc = trace_settle(u);
    if (rc != RL_OK) return rc;
    if (u->n_valid)
        RL_CUDA(copy_async(out, u->d_records, u->n_valid * sizeof(rl_mapped_photon), cudaMemcpyDeviceToHost,
                           u->ss.stream));
    RL_CUDA(cudaStreamSynchronize(u->ss.stream));
    if (out_count) *out_count = u->n_valid;
    return RL_OK;
}

int rl_trace_unit_render_fused(rl_trace_unit *u, const rl_scene *scene, rl_plot_unit *plot,
                               uint64_t first_photon, uint64_t n_photons) {
    if (!u || !scene || !plot) return fail(RL_ERR_INVALID, "rl_trace_unit_render_fused: null argument");
    if (scene->dev.index != u->dev.index || plot->dev.index != u->dev.index)
        return fail(RL_ERR_INVALID, "scene, trace unit and plot unit must share a device");
    if (plot->width != u->width || plot->height != u->height)
        return fail(RL_ERR_INVALID, "trace and plot unit canvas sizes differ");
    DeviceGuard guard(u->dev.index);
    int rc = trace_settle(u);
    if (rc == RL_OK) rc = order_after(plot->ss, u->ss);
    if (rc != RL_OK) return rc;
    // a fused batch is consumed by the plot unit's stream: always a launch of its own
    TraceLaunch p;
    memset(&p, 0, sizeof(p));
    p.width = u->width; p.height = u->height;
    p.accum = plot->d_accum;
    p.n_segments = 1;
    p.seg[0].seed = u->seed; p.seg[0].first_photon = first_photon; p.seg[0].n_photons = n_photons;
    p.seg[0].records = nullptr; p.seg[0].ray_counter = u->d_rays;
    RL_CUDA(launch_trace(scene->ds, p, u->dev.sm_count, u->ss.stream));
    u->n_valid = 0;
    return order_after(u->ss, plot->ss);
}

int rl_trace_unit_ray_count(rl_trace_unit *u, uint64_t *out_rays) {
    if (!u || !out_rays) return fail(RL_ERR_INVALID, "null argument");
    DeviceGuard guard(u->dev.index);
    int rc = trace_settle(u);
    if (rc != RL_OK) return rc;
    unsigned long long v = 0;
    RL_CUDA(copy_async(&v, u->d_rays, sizeof(v), cudaMemcpyDeviceToHost, u->ss.stream));
    RL_CUDA(cudaStreamSynchronize(u->ss.stream));
    *out_rays = v;
    return RL_OK;
}

int rl_trace_unit_sync(rl_trace_unit *u) {
    if (!u) return fail(RL_ERR_INVALID, "null unit");
    DeviceGuard guard(u->dev.index);
    int rc = trace_settle(u);
    if (rc != RL_OK) return rc;
    RL_CUDA(cudaStreamSynchronize(u->ss.stream));
    return RL_OK;
}

int rl_scene_batch_counter_reset(const rl_scene *scene, uint64_t next_batch) {
    if (!scene) return fail(RL_ERR_INVALID, "null scene");
    scene->next_batch.store(next_batch);
    return RL_OK;
}

int rl_scene_dispatch_stats(const rl_scene *scene, uint64_t *out_launches, uint64_t *out_batches) {
    if (!scene) return fail(RL_ERR_INVALID, "null scene");
    std::lock_guard<std::mutex> lk(scene->dispatcher.m);
    if (out_launches) *out_launches = scene->dispatcher.launches;
    if (out_batches) *out_batches = scene->dispatcher.batches;
    return RL_OK;
}

void rl_transfer_counters(uint64_t *h2d_bytes, uint64_t *d2h_bytes) {
    if (h2d_bytes) *h2d_bytes = g_h2d_bytes.load();
    if (d2h_bytes) *d2h_bytes = g_d2h_bytes.load();
}
void rl_transfer_counters_reset(void) { g_h2d_bytes.store(0); g_d2h_bytes.store(0); }

// ------------------------------------------------------- host buffer pinning
int rl_host_register(void *ptr, size_t bytes) {
    if (!ptr || bytes == 0) return fail(RL_ERR_INVALID, "rl_host_register: null or empty buffer");
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return RL_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(RL_ERR_CUDA, cudaGetErrorString(e)); }
    return RL_OK;
}

int rl_host_unregister(void *ptr) {
    if (!ptr) return RL_OK;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e == cudaErrorHostMemoryNotRegistered) { cudaGetLastError(); return RL_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(RL_ERR_CUDA, cudaGetErrorString(e)); }
    return RL_OK;
}

// ---------------------------------------------------------------- PlotUnit
int rl_plot_unit_create(uint64_t id, uint32_t width, uint32_t height, rl_plot_unit **out) {
    if (!out || width == 0 || height == 0) return fail(RL_ERR_INVALID, "rl_plot_unit_create: bad argument");
    rl_plot_unit *u = new (std::nothrow) rl_plot_unit();
    if (!u) return fail(RL_ERR_NOMEM, "out of memory");
    u->id = id; u->width = width; u->height = height;
    int rc = current_device(u->dev);
    if (rc == RL_OK) rc = u->ss.create();
    if (rc != RL_OK) { delete u; return rc; }
    const size_t n = (size_t)width * height;
    cudaError_t e = cudaMalloc(&u->d_accum, n * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&u->d_packed, n * 3 * sizeof(float));
    // cleared on the unit's own stream: everything the unit does later is ordered behind it (a
    // memset on the default stream would not be, the units' streams are non-blocking)
    if (e == cudaSuccess) e = cudaMemsetAsync(u->d_accum, 0, n * sizeof(float4), u->ss.stream);
    if (e != cudaSuccess) {
        cudaFree(u->d_accum); cudaFree(u->d_packed); u->ss.destroy(); delete u;
        return fail(RL_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = u;
    return RL_OK;
}

int rl_plot_unit_destroy(rl_plot_unit *u) {
    if (!u) return RL_OK;
    DeviceGuard guard(u->dev.index);
    if (u->ss.stream) cudaStreamSynchronize(u->ss.stream);
    cudaFree(u->d_accum); cudaFree(u->d_packed); cudaFree(u->d_staging);
    u->ss.destroy();
    delete u;
    return RL_OK;
}

int rl_plot_unit_set_stream(rl_plot_unit *u, void *s) {
    if (!u) return fail(RL_ERR_INVALID, "null unit");
    DeviceGuard guard(u->dev.index);
    return u->ss.bind(s);
}

int rl_plot_unit_plot(rl_plot_unit *u, const rl_mapped_photon *photons, uint64_t n) {
    if (!u || (!photons && n)) return fail(RL_ERR_INVALID, "rl_plot_unit_plot: null argument");
    if (n == 0) return RL_OK;
    DeviceGuard guard(u->dev.index);
    if (u->staging_capacity < n) {
        RL_CUDA(cudaStreamSynchronize(u->ss.stream));
        cudaFree(u->d_staging);
        u->d_staging = nullptr; u->staging_capacity = 0;
        RL_CUDA(cudaMalloc(&u->d_staging, n * sizeof(rl_mapped_photon)));
        u->staging_capacity = n;
    }
    RL_CUDA(copy_async(u->d_staging, photons, n * sizeof(rl_mapped_photon),
                            cudaMemcpyHostToDevice, u->ss.stream));
    RL_CUDA(launch_splat(u->d_staging, n, u->d_accum, u->width, u->height, u->dev.sm_count,
                         u->ss.stream));
    // the caller may reuse `photons` as soon as we return
    RL_CUDA(cudaStreamSynchronize(u->ss.stream));
    return RL_OK;
}

int rl_plot_unit_plot_device(rl_plot_unit *u, rl_trace_unit *trace) {
    if (!u || !trace) return fail(RL_ERR_INVALID, "rl_plot_unit_plot_device: null argument");
    if (u->dev.index != trace->dev.index) return fail(RL_ERR_INVALID, "units on different devices");
    if (trace->n_valid == 0) return RL_OK;
    DeviceGuard guard(u->dev.index);
    // the batch has been launched (short host wait if it is still in the dispatcher's queue): the
    // trace unit's stream now holds the launch's event and the copy of the records to the host;
    // the splat is ordered behind the event only -- it runs beside the copy -- and the unit's next
    // batch behind both
    int rc = trace_settle(trace);
    if (rc != RL_OK) return rc;
    RL_CUDA(cudaStreamWaitEvent(u->ss.stream, trace->traced, 0));
    RL_CUDA(launch_splat(trace->d_records, trace->n_valid, u->d_accum, u->width, u->height,
                         u->dev.sm_count, u->ss.stream));
    return order_after(u->ss, trace->ss);
}

int rl_plot_unit_clear(rl_plot_unit *u) {
    if (!u) return fail(RL_ERR_INVALID, "null unit");
    DeviceGuard guard(u->dev.index);
    RL_CUDA(cudaMemsetAsync(u->d_accum, 0, (size_t)u->width * u->height * sizeof(float4), u->ss.stream));
    return RL_OK;
}

int rl_plot_unit_download(rl_plot_unit *u, float *xyz) {
    if (!u || !xyz) return fail(RL_ERR_INVALID, "rl_plot_unit_download: null argument");
    DeviceGuard guard(u->dev.index);
    const uint64_t n = (uint64_t)u->width * u->height;
    RL_CUDA(launch_pack_xyz(u->d_accum, u->d_packed, n, u->ss.stream));
    RL_CUDA(copy_async(xyz, u->d_packed, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, u->ss.stream));
    RL_CUDA(cudaStreamSynchronize(u->ss.stream));
    return RL_OK;
}

int rl_plot_unit_device_buffer(rl_plot_unit *u, void **out_ptr, size_t *out_bytes) {
    if (!u || !out_ptr) return fail(RL_ERR_INVALID, "null argument");
    *out_ptr = u->d_accum;
    if (out_bytes) *out_bytes = (size_t)u->width * u->height * sizeof(float4);
    return RL_OK;
}

int rl_plot_unit_ipc_export(rl_plot_unit *u, void *handle_out) {
    if (!u || !handle_out) return fail(RL_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == RL_IPC_HANDLE_BYTES, "IPC handle size");
    DeviceGuard guard(u->dev.index);
    cudaIpcMemHandle_t h;
    RL_CUDA(cudaIpcGetMemHandle(&h, u->d_accum));
    memcpy(handle_out, &h, sizeof(h));
    return RL_OK;
}

int rl_ipc_open(const void *handle, void **out_ptr) {
    if (!handle || !out_ptr) return fail(RL_ERR_INVALID, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    RL_CUDA(cudaIpcOpenMemHandle(out_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RL_OK;
}

int rl_ipc_close(void *ptr) {
    if (!ptr) return RL_OK;
    RL_CUDA(cudaIpcCloseMemHandle(ptr));
    return RL_OK;
}

int rl_plot_unit_sync(rl_plot_unit *u) {
    if (!u) return fail(RL_ERR_INVALID, "null unit");
    DeviceGuard guard(u->dev.index);
    RL_CUDA(cudaStreamSynchronize(u->ss.stream));
    return RL_OK;
}

// -------------------------------------------------------------- GatherUnit
int rl_gather_unit_create(uint32_t width, uint32_t height, const char *resume_path,
                          rl_gather_unit **out) {
    if (!out || width == 0 || height == 0) return fail(RL_ERR_INVALID, "rl_gather_unit_create: bad argument");
    rl_gather_unit *u = new (std::nothrow) rl_gather_unit();
    if (!u) return fail(RL_ERR_NOMEM, "out of memory");
    u->width = width; u->height = height;
    int rc = current_device(u->dev);
    if (rc == RL_OK) rc = u->ss.create();
    if (rc != RL_OK) { delete u; return rc; }
    const size_t bytes = (size_t)width * height * 3 * sizeof(float);
    cudaError_t e = cudaMalloc(&u->d_acc, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&u->d_comp, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&u->d_staging, bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(u->d_acc, 0, bytes, u->ss.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(u->d_comp, 0, bytes, u->ss.stream);
    if (e != cudaSuccess) {
        cudaFree(u->d_acc); cudaFree(u->d_comp); cudaFree(u->d_staging); u->ss.destroy(); delete u;
        return fail(RL_ERR_CUDA, cudaGetErrorString(e));
    }
    if (resume_path) {
        // gather_unit.rs:43,82: resume silently iff the file exists
        FILE *f = fopen(resume_path, "rb");
        if (f) {
            fclose(f);
            rc = rl_gather_unit_load(u, resume_path);
            if (rc != RL_OK) { rl_gather_unit_destroy(u); return rc; }
        }
    }
    *out = u;
    return RL_OK;
}

int rl_gather_unit_destroy(rl_gather_unit *u) {
    if (!u) return RL_OK;
    DeviceGuard guard(u->dev.index);
    if (u->ss.stream) cudaStreamSynchronize(u->ss.stream);
    u->writer.shutdown();          // writes out a snapshot that is still queued
    cudaFree(u->d_acc); cudaFree(u->d_comp); cudaFree(u->d_staging);
    u->ss.destroy();
    delete u;
    return RL_OK;
}

int rl_gather_unit_set_stream(rl_gather_unit *u, void *s) {
    if (!u) return fail(RL_ERR_INVALID, "null unit");
    DeviceGuard guard(u->dev.index);
    return u->ss.bind(s);
}

int rl_gather_unit_accumulate(rl_gather_unit *u, const float *xyz) {
    if (!u || !xyz) return fail(RL_ERR_INVALID, "rl_gather_unit_accumulate: null argument");
    DeviceGuard guard(u->dev.index);
    const uint64_t n = (uint64_t)u->width * u->height;
    RL_CUDA(copy_async(u->d_staging, xyz, n * 3 * sizeof(float), cudaMemcpyHostToDevice, u->ss.stream));
    RL_CUDA(launch_gather(u->d_acc, u->d_comp, nullptr, 0, u->d_staging, nullptr, n, u->dev.sm_count,
                          u->ss.stream));
    RL_CUDA(cudaStreamSynchronize(u->ss.stream));
    return RL_OK;
}

int rl_gather_unit_accumulate_plot(rl_gather_unit *u, rl_plot_unit *plot, int clear_plot) {
    if (!u || !plot) return fail(RL_ERR_INVALID, "rl_gather_unit_accumulate_plot: null argument");
    if (u->dev.index != plot->dev.index) return fail(RL_ERR_INVALID, "units on different devices");
    if (u->width != plot->width || u->height != plot->height)
        return fail(RL_ERR_INVALID, "gather and plot unit canvas sizes differ");
    DeviceGuard guard(u->dev.index);
    int rc = order_after(plot->ss, u->ss);
    if (rc != RL_OK) return rc;
    const float4 *src = plot->d_accum;
    RL_CUDA(launch_gather(u->d_acc, u->d_comp, &src, 1, nullptr, clear_plot ? plot->d_accum : nullptr,
                          (uint64_t)u->width * u->height, u->dev.sm_count, u->ss.stream));
    return order_after(u->ss, plot->ss);
}

int rl_gather_unit_accumulate_device(rl_gather_unit *u, const void *const *bufs, uint32_t n_buffers) {
    if (!u || (!bufs && n_buffers)) return fail(RL_ERR_INVALID, "rl_gather_unit_accumulate_device: null argument");
    DeviceGuard guard(u->dev.index);
    const uint64_t n = (uint64_t)u->width * u->height;
    for (uint32_t i = 0; i < n_buffers; i += 8) {
        const float4 *srcs[8];
        uint32_t k = n_buffers - i < 8 ? n_buffers - i : 8;
        for (uint32_t j = 0; j < k; j++) srcs[j] = (const float4 *)bufs[i + j];
        RL_CUDA(launch_gather(u->d_acc, u->d_comp, srcs, k, nullptr, nullptr, n, u->dev.sm_count,
                              u->ss.stream));
    }
    return RL_OK;
}

int rl_gather_unit_save(rl_gather_unit *u, const char *path) {
    if (!u || !path) return fail(RL_ERR_INVALID, "rl_gather_unit_save: null argument");
    DeviceGuard guard(u->dev.index);
    SaveWriter &wr = u->writer;
    const size_t n = (size_t)u->width * u->height * 3;
    if (!wr.host) {
        const cudaError_t e = wr.allocate(u->dev.index, 2 * n);
        if (e != cudaSuccess) return fail(RL_ERR_CUDA, cudaGetErrorString(e));
    }
    bool proven = false;
    for (const std::string &p : wr.proven) proven |= p == path;
    bool no_thread = false;
    if (proven) {
        std::unique_lock<std::mutex> lock(wr.m);
        // a snapshot queued for another file is not this one's to replace: let it land first
        if (wr.latest >= 0 && wr.latest_path != path) {
            wr.hurry++;
            wr.cv.notify_all();
            wr.cv.wait(lock, [&] { return wr.latest < 0; });
            wr.hurry--;
        }
        if (!wr.error.empty()) {
            std::string e;
            e.swap(wr.error);
            return fail(RL_ERR_IO, e);
        }
        if (!wr.started) {
            try {
                wr.thread = std::thread([&wr] { wr.run(); });
                wr.started = true;
            } catch (const std::exception &) {      // no exception crosses the C boundary
                no_thread = true;
            }
        }
        if (!no_thread) {
            // The snapshot is taken in stream order, on the device; nobody waits for it here.  The
            // gather task holds the gather unit and every finished plot unit while it runs
            // (task_scheduler.rs:209-219): a host wait at this point is as long as the GPU's queue
            // is deep, and the scheduler has no plot unit to hand out meanwhile.  The lock is held
            // while the copies are queued so that the writer cannot pick this buffer half-queued;
            // the buffer it is copying out (`reading`) is never the one written here.
            const int idx = wr.latest >= 0 ? wr.latest : (wr.reading == 0 ? 1 : 0);
            RL_CUDA(cudaMemcpyAsync(wr.dsnap[idx], u->d_acc, n * sizeof(float), cudaMemcpyDeviceToDevice, u->ss.stream));
            RL_CUDA(cudaMemcpyAsync(wr.dsnap[idx] + n, u->d_comp, n * sizeof(float), cudaMemcpyDeviceToDevice, u->ss.stream));
            RL_CUDA(cudaEventRecord(wr.taken[idx], u->ss.stream));
            wr.latest = idx;
            wr.latest_path = path;
            lock.unlock();
            wr.cv.notify_all();
            return RL_OK;
        }
    }
    // first save to this path, or no writer thread to be had: written here, in the caller
    {
        const std::string e = wr.flush();
        if (!e.empty()) return fail(RL_ERR_IO, e);
    }
    RL_CUDA(copy_async(wr.host, u->d_acc, n * sizeof(float), cudaMemcpyDeviceToHost, u->ss.stream));
    RL_CUDA(copy_async(wr.host + n, u->d_comp, n * sizeof(float), cudaMemcpyDeviceToHost, u->ss.stream));
    RL_CUDA(cudaStreamSynchronize(u->ss.stream));
    std::string err;
    if (!SaveWriter::write_file(path, wr.host, 2 * n, err)) return fail(RL_ERR_IO, err);
    if (!proven) wr.proven.push_back(path);
    return RL_OK;
}

int rl_gather_unit_set_save_interval(rl_gather_unit *u, double seconds) {
    if (!u || !(seconds >= 0.0)) return fail(RL_ERR_INVALID, "rl_gather_unit_set_save_interval: invalid argument");
    std::lock_guard<std::mutex> lock(u->writer.m);
    u->writer.min_interval = seconds;
    u->writer.cv.notify_all();
    return RL_OK;
}

int rl_gather_unit_flush(rl_gather_unit *u) {
    if (!u) return fail(RL_ERR_INVALID, "null unit");
    const std::string e = u->writer.flush();
    if (!e.empty()) return fail(RL_ERR_IO, e);
    return RL_OK;
}

int rl_gather_unit_load(rl_gather_unit *u, const char *path) {
    if (!u || !path) return fail(RL_ERR_INVALID, "rl_gather_unit_load: null argument");
    DeviceGuard guard(u->dev.index);
    {
        const std::string e = u->writer.flush();     // a queued save of the same file lands first
        if (!e.empty()) return fail(RL_ERR_IO, e);
    }
    const size_t n = (size_t)u->width * u->height * 3;
    FILE *f = fopen(path, "rb");
    if (!f) return fail(RL_ERR_IO, std::string("failed to open file ") + path);
    // read.rs:20-32: a short file is not an error; what was not read keeps
    // its current value.
    std::vector<float> host(2 * n);
    RL_CUDA(copy_async(host.data(), u->d_acc, n * sizeof(float), cudaMemcpyDeviceToHost, u->ss.stream));
    RL_CUDA(copy_async(host.data() + n, u->d_comp, n * sizeof(float), cudaMemcpyDeviceToHost, u->ss.stream));
    RL_CUDA(cudaStreamSynchronize(u->ss.stream));
    size_t got = fread(host.data(), 1, 2 * n * sizeof(float), f);
    (void)got;
    fclose(f);
    RL_CUDA(copy_async(u->d_acc, host.data(), n * sizeof(float), cudaMemcpyHostToDevice, u->ss.stream));
    RL_CUDA(copy_async(u->d_comp, host.data() + n, n * sizeof(float), cudaMemcpyHostToDevice, u->ss.stream));
    RL_CUDA(cudaStreamSynchronize(u->ss.stream));
    return RL_OK;
}

int rl_gather_unit_download(rl_gather_unit *u, float *xyz, float *comp) {
    if (!u || !xyz) return fail(RL_ERR_INVALID, "rl_gather_unit_download: null argument");
    DeviceGuard guard(u->dev.index);
    const size_t bytes = (size_t)u->width * u->height * 3 * sizeof(float);
    RL_CUDA(copy_async(xyz, u->d_acc, bytes, cudaMemcpyDeviceToHost, u->ss.stream));
    if (comp) RL_CUDA(copy_async(comp, u->d_comp, bytes, cudaMemcpyDeviceToHost, u->ss.stream));
    RL_CUDA(cudaStreamSynchronize(u->ss.stream));
    return RL_OK;
}

int rl_gather_unit_sync(rl_gather_unit *u) {
    if (!u) return fail(RL_ERR_INVALID, "null unit");
    DeviceGuard guard(u->dev.index);
    RL_CUDA(cudaStreamSynchronize(u->ss.stream));
    return RL_OK;
}

// ------------------------------------------------------------- TonemapUnit
int rl_tonemap_unit_create(uint32_t width, uint32_t height, rl_tonemap_unit **out) {
    if (!out || width == 0 || height == 0) return fail(RL_ERR_INVALID, "rl_tonemap_unit_create: bad argument");
    rl_tonemap_unit *u = new (std::nothrow) rl_tonemap_unit();
    if (!u) return fail(RL_ERR_NOMEM, "out of memory");
    u->width = width; u->height = height;
    int rc = current_device(u->dev);
    if (rc == RL_OK) rc = u->ss.create();
    if (rc != RL_OK) { delete u; return rc; }
    const size_t n = (size_t)width * height;
    cudaError_t e = cudaMalloc(&u->d_xyz, n * 3 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&u->d_rgb, n * 3);
    if (e == cudaSuccess) e = cudaMalloc(&u->d_moments, 2 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&u->d_exposure, sizeof(float));
    if (e != cudaSuccess) {
        cudaFree(u->d_xyz); cudaFree(u->d_rgb); cudaFree(u->d_moments); cudaFree(u->d_exposure);
        u->ss.destroy(); delete u;
        return fail(RL_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = u;
    return RL_OK;
}

int rl_tonemap_unit_set_exposure_mode(rl_tonemap_unit *u, int mode) {
    if (!u || (mode != RL_EXPOSURE_F64_REDUCTION && mode != RL_EXPOSURE_REFERENCE_FOLD))
        return fail(RL_ERR_INVALID, "rl_tonemap_unit_set_exposure_mode: bad argument");
    u->reference_fold = mode == RL_EXPOSURE_REFERENCE_FOLD;
    return RL_OK;
}

int rl_tonemap_unit_destroy(rl_tonemap_unit *u) {
    if (!u) return RL_OK;
    DeviceGuard guard(u->dev.index);
    if (u->ss.stream) cudaStreamSynchronize(u->ss.stream);
    cudaFree(u->d_xyz); cudaFree(u->d_rgb); cudaFree(u->d_moments); cudaFree(u->d_exposure);
    u->ss.destroy();
    delete u;
    return RL_OK;
}

int rl_tonemap_unit_set_stream(rl_tonemap_unit *u, void *s) {
    if (!u) return fail(RL_ERR_INVALID, "null unit");
    DeviceGuard guard(u->dev.index);
    return u->ss.bind(s);
}

static int tonemap_device(rl_tonemap_unit *u, const float *d_xyz, uint8_t *rgb) {
    RL_CUDA(launch_tonemap(d_xyz, u->width, u->height, u->d_moments, u->d_exposure, u->d_rgb,
                           u->dev.sm_count, u->reference_fold, u->ss.stream));
    RL_CUDA(copy_async(&u->last_exposure, u->d_exposure, sizeof(float), cudaMemcpyDeviceToHost,
                            u->ss.stream));
    if (rgb)
        RL_CUDA(copy_async(rgb, u->d_rgb, (size_t)u->width * u->height * 3, cudaMemcpyDeviceToHost,
                                u->ss.stream));
    RL_CUDA(cudaStreamSynchronize(u->ss.stream));
    return RL_OK;
}

int rl_tonemap_unit_tonemap(rl_tonemap_unit *u, const float *xyz, uint8_t *rgb) {
    if (!u || !xyz || !rgb) return fail(RL_ERR_INVALID, "rl_tonemap_unit_tonemap: null argument");
    DeviceGuard guard(u->dev.index);
    RL_CUDA(copy_async(u->d_xyz, xyz, (size_t)u->width * u->height * 3 * sizeof(float),
                            cudaMemcpyHostToDevice, u->ss.stream));
    return tonemap_device(u, u->d_xyz, rgb);
}

int rl_tonemap_unit_tonemap_gather(rl_tonemap_unit *u, rl_gather_unit *g, uint8_t *rgb) {
    if (!u || !g) return fail(RL_ERR_INVALID, "rl_tonemap_unit_tonemap_gather: null argument");
    if (u->dev.index != g->dev.index) return fail(RL_ERR_INVALID, "units on different devices");
    if (u->width != g->width || u->height != g->height)
        return fail(RL_ERR_INVALID, "tonemap and gather unit canvas sizes differ");
    DeviceGuard guard(u->dev.index);
    int rc = order_after(g->ss, u->ss);
    if (rc != RL_OK) return rc;
    rc = tonemap_device(u, g->d_acc, rgb);
    if (rc != RL_OK) return rc;
    return order_after(u->ss, g->ss);
}

int rl_tonemap_unit_last_exposure(rl_tonemap_unit *u, float *out) {
    if (!u || !out) return fail(RL_ERR_INVALID, "null argument");
    *out = u->last_exposure;
    return RL_OK;
}

// ------------------------------------------------------------------ probes
int rl_debug_intersect(const rl_scene *scene, const rl_ray *rays, uint64_t n, rl_hit *out) {
    if (!scene || (!rays && n) || (!out && n)) return fail(RL_ERR_INVALID, "null argument");
    if (n == 0) return RL_OK;
    DeviceGuard guard(scene->dev.index);
    rl_ray *d_rays = nullptr; rl_hit *d_out = nullptr;
    RL_CUDA(cudaMalloc(&d_rays, n * sizeof(rl_ray)));
    cudaError_t e = cudaMalloc(&d_out, n * sizeof(rl_hit));
    if (e == cudaSuccess) e = cudaMemcpy(d_rays, rays, n * sizeof(rl_ray), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_debug_intersect(scene->ds, d_rays, n, d_out, 0);
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, n * sizeof(rl_hit), cudaMemcpyDeviceToHost);
    cudaFree(d_rays); cudaFree(d_out);
    if (e != cudaSuccess) return fail(RL_ERR_CUDA, cudaGetErrorString(e));
    return RL_OK;
}

int rl_debug_cull_check(const rl_scene *scene, uint64_t seed, uint32_t width, uint32_t height,
                        uint64_t first_photon, uint64_t n, uint64_t *out_rays, uint64_t *out_mismatches) {
    if (!scene || !out_rays || !out_mismatches || width == 0 || height == 0)
        return fail(RL_ERR_INVALID, "bad argument");
    DeviceGuard guard(scene->dev.index);
    unsigned long long *d = nullptr, h[2] = {0, 0};
    RL_CUDA(cudaMalloc(&d, 2 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemset(d, 0, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = launch_debug_cull_check(scene->ds, seed, width, height, first_photon, n, d, d + 1, 0);
    if (e == cudaSuccess) e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(RL_ERR_CUDA, cudaGetErrorString(e));
    *out_rays = h[0];
    *out_mismatches = h[1];
    return RL_OK;
}

int rl_debug_math(int fn, const float *in, const float *in2, uint64_t n, float *out) {
    if ((!in || !out) && n) return fail(RL_ERR_INVALID, "null argument");
    if (fn < 0 || fn > 8) return fail(RL_ERR_INVALID, "unknown function");
    if ((fn == 4 || fn == 7) && !in2) return fail(RL_ERR_INVALID, "second operand required");
    if (n == 0) return RL_OK;
    Device dev;
    int rc = current_device(dev);
    if (rc != RL_OK) return rc;
    float *d_in = nullptr, *d_in2 = nullptr, *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_in, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, n * sizeof(float));
    if (e == cudaSuccess && in2) e = cudaMalloc(&d_in2, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(d_in, in, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && in2) e = cudaMemcpy(d_in2, in2, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_debug_math(fn, d_in, d_in2, n, d_out, 0);
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_in2); cudaFree(d_out);
    if (e != cudaSuccess) return fail(RL_ERR_CUDA, cudaGetErrorString(e));
    return RL_OK;
}

int rl_debug_tristimulus(const float *wl, uint64_t n, float *out_xyz) {
    if ((!wl || !out_xyz) && n) return fail(RL_ERR_INVALID, "null argument");
    if (n == 0) return RL_OK;
    Device dev;
    int rc = current_device(dev);
    if (rc != RL_OK) return rc;
    float *d_in = nullptr, *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_in, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, 3 * n * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(d_in, wl, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_debug_tristimulus(d_in, n, d_out, 0);
    if (e == cudaSuccess) e = cudaMemcpy(out_xyz, d_out, 3 * n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) return fail(RL_ERR_CUDA, cudaGetErrorString(e));
    return RL_OK;
}

int rl_debug_camera_rays(const rl_scene *scene, uint64_t seed, uint32_t width, uint32_t height,
                         uint64_t first_photon, uint64_t n, rl_ray *out_rays, rl_mapped_photon *out_xy) {
    if (!scene || (!out_rays && n) || width == 0 || height == 0) return fail(RL_ERR_INVALID, "bad argument");
    if (n == 0) return RL_OK;
    DeviceGuard guard(scene->dev.index);
    rl_ray *d_rays = nullptr; rl_mapped_photon *d_xy = nullptr;
    cudaError_t e = cudaMalloc(&d_rays, n * sizeof(rl_ray));
    if (e == cudaSuccess && out_xy) e = cudaMalloc(&d_xy, n * sizeof(rl_mapped_photon));
    if (e == cudaSuccess) e = launch_debug_camera(scene->ds, seed, width, height, first_photon, n, d_rays, d_xy, 0);
    if (e == cudaSuccess) e = cudaMemcpy(out_rays, d_rays, n * sizeof(rl_ray), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && out_xy) e = cudaMemcpy(out_xy, d_xy, n * sizeof(rl_mapped_photon), cudaMemcpyDeviceToHost);
    cudaFree(d_rays); cudaFree(d_xy);
    if (e != cudaSuccess) return fail(RL_ERR_CUDA, cudaGetErrorString(e));
    return RL_OK;
}

}  // extern "C"
